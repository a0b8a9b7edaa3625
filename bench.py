#!/usr/bin/env python
"""Benchmark of the nexoclom hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--packets P] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic packets:
restore the resident initial state, K2 adaptive integration, K4 radiance image
(800x800) -- BASELINE.json configs[1] (Na at Mercury, Maxwellian surface source,
radiation pressure + photoionisation, 1e7 packets per GPU).  Metric:
packet-steps/s = attempted Dormand-Prince steps / time, whole job over all GPUs.
Packets are sharded over ranks by global id (weak scaling, no data-path
collective); per-GPU images are combined with one NCCL all-reduce per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

WORKLOAD = 'Na.maxwellian.radpres.input'
FLOP_PER_STEP = 764          # algorithmic flop per attempted step
METRIC = 'packet-steps/s (FP64)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smmax)) if smmax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


# ---------------------------------------------------------------------------
# CPU arms (oracle port of the reference's NumPy path)
# ---------------------------------------------------------------------------
def _cpu_chunk(args):
    name, X0 = args
    from common import workload, oracle_constants
    from nexoclom_b200.runsetup import RunSetup
    from oracle import tracking
    rc = oracle_constants(RunSetup(workload(name)))
    t0 = time.time()
    _, att, _ = tracking.integrate_adaptive(X0, rc)
    return int(att.sum()), time.time() - t0


def cpu_steps_per_s(X0, procs):
    """Oracle adaptive driver over X0 split into `procs` chunks; packet-steps/s."""
    import multiprocessing as mp
    chunks = np.array_split(X0, procs)
    t0 = time.time()
    if procs == 1:
        res = [_cpu_chunk((WORKLOAD, chunks[0]))]
    else:
        with mp.get_context('spawn').Pool(procs) as pool:
            res = pool.map(_cpu_chunk, [(WORKLOAD, c) for c in chunks])
    wall = time.time() - t0
    steps = sum(r[0] for r in res)
    return steps / wall, steps, wall


def synth_x0_host(n, seed=0):
    """Host-side synthetic initial state of the bench workload for the CPU arm
    (NumPy draws through the oracle's restatement of the reference's samplers)."""
    from common import workload
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state
    setup = RunSetup(workload(WORKLOAD))
    return initial_state.draw_x0(setup, n, seed)[:, :8]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n = args.ref_packets * procs
    X0 = synth_x0_host(n)
    vals = []
    for it in range(args.warmup + args.steps):
        sps, steps, wall = cpu_steps_per_s(X0, procs)
        if it >= args.warmup:
            vals.append((sps, steps, wall))
    sps = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[2] for v in vals])) * 1e3
    sample = f'{n} packets of {WORKLOAD} per step ({procs} processes x {args.ref_packets})'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': sps, 'unit': 'packet-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'Na at Mercury, Maxwellian 1200 K surface source, radiation '
                               'pressure + photoionisation, adaptive RK5(4) (configs[1]); '
                               'bounded CPU sample', 'inputfile': WORKLOAD, 'packets': n},
        'cpu_baseline': {'value': sps, 'unit': 'packet-steps/s', 'cores': procs, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': sps, 'unit': 'packet-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from common import workload
    from nexoclom_b200._lib import ImageParams
    from nexoclom_b200.engine import Engine
    from nexoclom_b200.runsetup import RunSetup
    from nexoclom_b200.ModelImage import image_rotation

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: nexoclom_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    try:
        # run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU:
        # at 8 ranks the end-to-end path moves 8 x 640 MB per step out of host memory
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    eng = Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    n = args.packets
    setup = RunSetup(workload(WORKLOAD))
    setup.upload(eng)
    gt = setup.gtables([5891, 5897])
    eng.upload_gtables(gt)
    sp = setup.source_params(eng)
    first_id = rank * n
    eng.init_state(sp, args.seed, first_id, n)           # resident inputs
    eng.sync()
    init_ms = eng.last_kernel_ms()
    # keep a pristine device copy of the 8 state columns
    cap_cols = [eng.state_device_ptr(k) for k in range(9)]
    x0_dev = torch.empty((8, n), dtype=torch.float64, device='cuda')

    class _DevArray:
        """Zero-copy torch view of a library-owned device column."""

        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {'shape': (count,), 'typestr': '<f8',
                                             'data': (ptr, False), 'version': 2}

    state_cols = [torch.as_tensor(_DevArray(p, n), device='cuda') for p in cap_cols]

    def restore_state():
        for k in range(8):
            state_cols[k].copy_(x0_dev[k])
        state_cols[8].fill_(1000.0)

    for k in range(8):
        x0_dev[k].copy_(state_cols[k])

    M = np.asarray(image_rotation(0.0, np.pi / 2))
    ip = ImageParams()
    for k in range(9):
        ip.M[k] = float(M.flat[k])
    ip.x0, ip.x1, ip.z0, ip.z1 = -4., 4., -4., 4.
    ip.nx = ip.nz = 800
    rcm = setup.radius_km * 1e5
    ip.apix = (8 / 800 * rcm) * (8 / 800 * rcm)
    ip.vrplanet = setup.vrplanet
    ip.quantity = 1
    ip.round_f32 = 1
    ip.skip_dead = 1
    image = torch.zeros((800, 800), dtype=torch.float64, device='cuda')
    counts = torch.zeros((800, 800), dtype=torch.int64, device='cuda')

    k2_ms, k4_ms, steps_total = [], [], []

    def one_step(record):
        restore_state()
        image.zero_()
        counts.zero_()
        att, _ = eng.integrate_adaptive(n)
        if record:
            k2_ms.append(eng.last_kernel_ms())
        eng.image_accumulate_dev(ip, image.data_ptr(), counts.data_ptr(), n)
        if record:
            eng.sync()
            k4_ms.append(eng.last_kernel_ms())
        if world > 1:
            dist.all_reduce(image)
            dist.all_reduce(counts)
        if record:
            steps_total.append(att)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(False)
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        one_step(True)
    ev1.record(stream)
    fence()
    launches = eng.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms, float(sum(steps_total))], dtype=torch.float64, device='cuda')
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, all_steps = float(tmax[0]), float(tsum[1])
    else:
        all_steps = float(t[1])
    value = all_steps / (elapsed_ms * 1e-3)

    # ---- end-to-end through the host-buffer C ABI (H2D + kernels + D2H) ----
    host_in = torch.empty((8, n), dtype=torch.float64).pin_memory()
    host_in.copy_(x0_dev.cpu())
    host_np = host_in.numpy()
    cols = [host_np[k] for k in range(8)]
    img_host = torch.empty((800, 800), dtype=torch.float64).pin_memory()
    e2e_steps, e2e_ms = 0, 0.0
    for it in range(0 if args.no_e2e else 1 + max(1, min(args.steps, 3))):
        fence()
        t0 = time.perf_counter()
        # reference-facing host-buffer call: pinned H2D (64 B/packet) pipelined with K2
        att, _ = eng.integrate_adaptive_host(cols, nchunks=args.e2e_chunks)
        image.zero_()
        counts.zero_()
        eng.image_accumulate_dev(ip, image.data_ptr(), counts.data_ptr(), n)
        if world > 1:
            dist.all_reduce(image)
        img_host.copy_(image, non_blocking=False)          # D2H result
        fence()
        if it > 0:
            e2e_ms += (time.perf_counter() - t0) * 1e3
            e2e_steps += att
    te = torch.tensor([e2e_ms, float(e2e_steps)], dtype=torch.float64, device='cuda')
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = te.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        e2e_ms, e2e_steps = float(tm[0]), float(ts[1])
    e2e_value = e2e_steps / (e2e_ms * 1e-3) if e2e_ms > 0 else None

    # ---- line-of-sight sweep over the resident final state (not part of `value`) ----
    los_info = None
    if not args.no_los:
        from nexoclom_b200._lib import LosParams
        nlos = args.los
        g = torch.Generator(device='cpu').manual_seed(1)
        th = torch.rand(nlos, generator=g, dtype=torch.float64) * 2 * np.pi
        rr = 1.1 + 4.9 * torch.rand(nlos, generator=g, dtype=torch.float64)
        x_sc = torch.stack([0.3 * rr * torch.cos(th), 0.2 * rr * torch.cos(th) - 0.5,
                            rr * torch.sin(th)], dim=0)
        x_sc *= torch.clamp(x_sc.norm(dim=0), min=1.1) / x_sc.norm(dim=0)
        tgt = torch.randn(3, nlos, generator=g, dtype=torch.float64)
        tgt *= (1 + 3 * torch.rand(nlos, generator=g, dtype=torch.float64)) / tgt.norm(dim=0)
        bore = tgt - x_sc
        bore /= bore.norm(dim=0)
        dplan = x_sc.norm(dim=0)
        ang = torch.arccos(-(x_sc * bore).sum(dim=0) / dplan)
        dplan = torch.where(ang > torch.arcsin(1. / dplan), torch.full_like(dplan, 1e30), dplan)
        los_host = torch.cat([x_sc, bore], dim=0).T.contiguous().numpy()
        los_dev = torch.cat([x_sc, bore], dim=0).contiguous().cuda()
        dist_dev = dplan.contiguous().cuda()
        rad_dev = torch.zeros(nlos, dtype=torch.float64, device='cuda')
        npk_dev = torch.zeros(nlos, dtype=torch.int64, device='cuda')
        inc_dev = torch.zeros(n, dtype=torch.uint8, device='cuda')
        lp = LosParams()
        lp.dphi, lp.outeredge = float(np.radians(1.0)), 25.0
        lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
        lp.quantity, lp.round_f32, lp.skip_dead = 1, 1, 0
        los_ms = []
        for it in range(3):
            rad_dev.zero_(); npk_dev.zero_(); inc_dev.zero_()
            eng.los_accumulate_dev(nlos, los_dev.data_ptr(), dist_dev.data_ptr(), lp,
                                   rad_dev.data_ptr(), npk_dev.data_ptr(), inc_dev.data_ptr(), n)
            if world > 1:
                dist.all_reduce(rad_dev)
                dist.all_reduce(npk_dev)
            torch.cuda.synchronize()
            los_ms.append(eng.last_kernel_ms())
        los_info = {'lines_of_sight': nlos, 'packets_per_gpu': n, 'dphi_deg': 1.0,
                    'ms': float(min(los_ms)), 'ms_per_1e8_packets': float(min(los_ms)) * 1e8 / n,
                    'hits': int(npk_dev.sum().item()),
                    'note': 'K5 cell-grid path over all resident packets (skip_dead=0), '
                            'grid build included'}

    if rank == 0:
        fp64_peak = eng.measure_fp64_peak()
        peaks, peak_kind = measured_peaks()
        k2 = float(np.mean(k2_ms))
        k4 = float(np.mean(k4_ms))
        steps_per_launch = float(np.mean(steps_total))
        achieved = steps_per_launch * FLOP_PER_STEP / (k2 * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(REPO, 'profiles', 'k2_traffic_bytes.json')
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        line = {
            'metric': METRIC, 'value': value, 'unit': 'packet-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': 'Na at Mercury, Maxwellian 1200 K surface source, radiation '
                            'pressure + photoionisation, adaptive RK5(4), + 800x800 radiance '
                            'image (BASELINE configs[1])',
                'inputfile': WORKLOAD, 'packets_per_gpu': n, 'packets_total': n * world,
                'attempted_steps_per_packet': steps_per_launch / n,
                'l2': 'inputs (640 MB state per 1e7 packets) exceed the 126 MB L2; no flush',
                'step': 'restore resident X0 -> K2 adaptive integrate -> K4 image'
                        + (' -> NCCL all-reduce(image, counts)' if world > 1 else ''),
                'k1_init_ms': init_ms, 'k2_integrate_ms': k2, 'k4_image_ms': k4,
                'image_ms_per_1e8_packets': k4 * 1e8 / n,
                'image_hbm_gbs': 40.0 * n / (k4 * 1e-3) / 1e9,
                'image_hbm_frac_of_' + peak_kind: 40.0 * n / (k4 * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'los_sweep': los_info,
            },
            'roofline': {
                'bound': 'fp64', 'kernel': 'k_integrate_adaptive', 'achieved': achieved,
                'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': achieved / fp64_peak,
                'traffic': traffic,
                'peak_source': 'DFMA-chain microbenchmark run live by bench.py '
                               '(MEASURED_PEAKS.json has no FP64 entry); nominal 37.2',
                'flop_per_packet_step': FLOP_PER_STEP,
            },
            # second kernel of the step, in the task's own schema (HBM-bound image accumulation;
            # 40 B per packet = x, y, z, vy, frac; ncu: DRAM traffic == algorithmic bytes)
            'roofline_hbm': {
                'bound': 'hbm', 'kernel': 'k_image_accumulate',
                'achieved': 40.0 * n / (k4 * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                'unit': 'GB/s', 'frac': 40.0 * n / (k4 * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'traffic': int(40.18 * n), 'peak_source': peak_kind + ' (MEASURED_PEAKS.json)',
                'bytes_per_packet': 40},
            'e2e': {'value': e2e_value, 'unit': 'packet-steps/s',
                    'h2d_bytes_per_step': 64 * n, 'd2h_bytes_per_step': 8 * 800 * 800,
                    'ms_per_step': e2e_ms / max(1, min(args.steps, 3)),
                    'call': f'nx_integrate_adaptive_host(nchunks={args.e2e_chunks}) -> '
                            'nx_image_accumulate_dev -> D2H image'},
            'gpu_launches': int(launches),
            'clocks': clocks,
        }
        if world == 1 and not args.no_cpu:
            X0h = np.ascontiguousarray(host_np[:, :args.cpu_packets].T)
            sps, steps, wall = cpu_steps_per_s(X0h, 1)
            line['cpu_baseline'] = {
                'value': sps, 'unit': 'packet-steps/s', 'cores': 1, 'kind': 'port',
                'sample': f'first {args.cpu_packets} packets of the same resident X0, '
                          f'{steps} steps in {wall:.1f} s, NumPy oracle port of the '
                          'reference driver, 1 process (host has '
                          f'{os.cpu_count()} cores)'}
            # the image / line-of-sight products on the same final state (SURVEY 8d ii-iii)
            m = min(n, args.cpu_product_packets)
            fin = torch.stack([state_cols[k][:m] for k in (1, 2, 3, 5, 7)]).cpu().numpy()
            line['config']['cpu_products'] = cpu_products(
                fin, setup, gt, M, ip.apix,
                los_host[:args.cpu_product_los] if los_info is not None else None)
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_products(fin, setup, gtables, M, apix, los):
    """CPU timings of the two products of the path on a bounded sample of the bench's final
    state: the NumPy oracle port of create_image (two np.histogram2d passes) and of the
    compute_iteration loop (sklearn KDTree), one process each."""
    from oracle import imaging
    x, y, z, vy, frac = (np.ascontiguousarray(fin[k].astype(np.float32).astype(np.float64))
                         for k in range(5))
    m = len(x)
    t0 = time.time()
    imaging.create_image(x, y, z, vy, frac, vrplanet=setup.vrplanet, M=M, dims=[800, 800],
                         xrange=(-4., 4.), zrange=(-4., 4.), apix=apix, quantity='radiance',
                         gtables=gtables)
    t_img = time.time() - t0
    out = {'image': {'packets': m, 'ms': t_img * 1e3, 'ms_per_1e8_packets': t_img * 1e3 * 1e8 / m,
                     'kind': 'port', 'cores': 1}}
    if los is not None and len(los):
        t0 = time.time()
        _, npk, _, _ = imaging.los_iteration(
            x, y, z, vy, frac, los, vrplanet=setup.vrplanet, dphi=float(np.radians(1.0)),
            outeredge=25.0, rp_cm=setup.radius_km * 1e5, gtables=gtables)
        t_los = time.time() - t0
        out['los'] = {'packets': m, 'lines_of_sight': int(len(los)), 's': t_los,
                      'ms_per_line_of_sight': t_los * 1e3 / len(los), 'hits': int(npk.sum()),
                      'kind': 'port', 'cores': 1,
                      'note': 'KD-tree build included; the GPU los_sweep above is '
                              '1e5 lines of sight over 10x the packets'}
    return out


def run_gpu_config3(args):
    """Extra measurement (not the driver's contract line): BASELINE configs[2] -- Na at
    Mercury, surface-bound packets with temperature-dependent sticking, bounce and thermal
    accommodation, constant 30 s step, the 800x800 radiance image accumulated per step
    INSIDE the integrator (K1 -> K3 fused), packets sharded over the ranks, one NCCL
    all-reduce of the image per step.  `--packets` per GPU (1.25e7 x 8 = the 1e8 run)."""
    import torch
    import torch.distributed as dist
    from common import workload
    from nexoclom_b200._lib import ImageParams
    from nexoclom_b200.engine import Engine
    from nexoclom_b200.runsetup import RunSetup
    from nexoclom_b200.ModelImage import image_rotation

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    n = args.packets
    setup = RunSetup(workload('Na.bounce.input'))
    setup.upload(eng)
    eng.upload_gtables(setup.gtables([5891, 5897]))
    sp = setup.source_params(eng)
    M = np.asarray(image_rotation(0.0, np.pi / 2))
    ip = ImageParams()
    for k in range(9):
        ip.M[k] = float(M.flat[k])
    ip.x0, ip.x1, ip.z0, ip.z1 = -4., 4., -4., 4.
    ip.nx = ip.nz = 800
    rcm = setup.radius_km * 1e5
    ip.apix = (8 / 800 * rcm) * (8 / 800 * rcm)
    ip.vrplanet = setup.vrplanet
    ip.quantity, ip.round_f32, ip.skip_dead = 1, 1, 1
    image = torch.zeros((800, 800), dtype=torch.float64, device='cuda')
    counts = torch.zeros((800, 800), dtype=torch.int64, device='cuda')
    steps_done = []

    def one_step(record):
        image.zero_()
        counts.zero_()
        eng.init_state(sp, args.seed, rank * n, n)                    # K1
        _, nsteps, psteps = eng.integrate_constant(                   # K3 + fused image
            seed=args.seed + 1, first_id=rank * n, image_params=ip, image_dev=image.data_ptr(),
            counts_dev=counts.data_ptr())
        if world > 1:
            dist.all_reduce(image)
            dist.all_reduce(counts)
        if record:
            steps_done.append(psteps)
        return nsteps

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        nsteps = one_step(False)
    fence()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        one_step(True)
    ev1.record(stream)
    fence()
    t = torch.tensor([ev0.elapsed_time(ev1), float(sum(steps_done))], dtype=torch.float64,
                     device='cuda')
    rows = counts.sum().item()
    if world > 1:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, all_steps = float(tmax[0]), float(tsum[1])
    else:
        elapsed_ms, all_steps = float(t[0]), float(t[1])
    if rank == 0:
        emit({
            'metric': 'packet-steps/s (FP64), constant step + bounce + fused image',
            'value': all_steps / (elapsed_ms * 1e-3), 'unit': 'packet-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[2]: Na at Mercury, T-dependent sticking + '
                                   'bounce + accommodation, 30 s step, image fused into K3',
                       'inputfile': 'Na.bounce.input', 'packets_per_gpu': n,
                       'packets_total': n * world, 'nsteps': int(nsteps),
                       'rows_in_image_per_step': int(rows),
                       'step': 'K1 init -> K3 constant-step integrate with fused radiance '
                               'image' + (' -> NCCL all-reduce(image, counts)' if world > 1 else '')}})
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: whatever libraries print there while the bench runs
    (NCCL writes its version banner to stdout) is sent to stderr instead."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--packets', type=int, default=10_000_000, help='packets per GPU')
    ap.add_argument('--cpu-packets', type=int, default=100000,
                    help='packets of the bounded CPU sample (about 15 s on one core)')
    ap.add_argument('--cpu-product-packets', type=int, default=1_000_000,
                    help='packets of the CPU image / line-of-sight timing sample')
    ap.add_argument('--cpu-product-los', type=int, default=200,
                    help='lines of sight of the CPU line-of-sight timing sample')
    ap.add_argument('--ref-packets', type=int, default=20000,
                    help='packets per host process and step of the --impl reference arm')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-los', action='store_true')
    ap.add_argument('--no-e2e', action='store_true',
                    help='skip the end-to-end leg (its streaming kernel waits for concurrent '
                         'copies, which a serialising profiler never runs)')
    ap.add_argument('--e2e-chunks', type=int, default=16,
                    help='segments of the streamed H2D copy of the end-to-end path')
    ap.add_argument('--los', type=int, default=100_000, help='lines of sight of the K5 sweep')
    ap.add_argument('--config3', action='store_true',
                    help='extra: the constant-step / bounce / fused-image workload (configs[2])')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.config3:
        return run_gpu_config3(args)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
