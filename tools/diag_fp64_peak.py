import sys; sys.path.insert(0,'/root/repo')
from nexoclom_b200.engine import Engine
e=Engine(0)
for i in range(3): print('fp64 microbenchmark', e.measure_fp64_peak(), 'TFLOP/s')
