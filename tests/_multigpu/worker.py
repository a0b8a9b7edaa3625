"""One rank of a sharded run through the public API (launched by torchrun from
tests/test_multigpu.py, or with WORLD_SIZE unset for the single-process answer).

    worker.py <workload> <npackets> <seed> <out.npz>

Input.run -> ModelImage -> LOSResult; rank 0 writes the products to <out.npz>."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def synthetic_scdata(nlos, species):
    import pandas as pd
    g = np.random.default_rng(5)
    th = g.random(nlos) * 2 * np.pi
    rr = 1.3 + 3 * g.random(nlos)
    x_sc = np.stack([0.4 * rr * np.cos(th), 0.3 * rr * np.cos(th) - 0.4, rr * np.sin(th)])
    tgt = g.standard_normal((3, nlos))
    tgt *= (1 + 2 * g.random(nlos)) / np.linalg.norm(tgt, axis=0)
    bore = tgt - x_sc
    bore /= np.linalg.norm(bore, axis=0)

    class SC:
        pass
    sc = SC()
    sc.data = pd.DataFrame(np.concatenate([x_sc, bore]).T,
                           columns=['x', 'y', 'z', 'xbore', 'ybore', 'zbore'])
    sc.data['radiance'] = np.linspace(1.0, 3.0, nlos)
    sc.data['sigma'] = 0.1
    sc.data['alttan'] = 1.0
    sc.species = species
    sc.query = 'synthetic'
    sc.subslong = sc.data.x * 0
    sc.set_frame = lambda frame: None
    SC.__len__ = lambda self: len(self.data)
    return sc


def main():
    name, npackets, seed, dest = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]), sys.argv[4]
    packs_per_it = int(float(sys.argv[5])) if len(sys.argv) > 5 else None
    from common import workload
    from nexoclom_b200 import ModelImage, LOSResult, sharding
    from nexoclom_b200.units import Quantity
    rank, world = sharding.init()
    inputs = workload(name)
    inputs.delete_files()
    inputs.run(npackets, seed=seed, packs_per_it=packs_per_it)
    _, files, mine, _ = inputs.search()
    image = ModelImage(inputs, {'quantity': 'radiance', 'dims': '300,300'})
    column = ModelImage(inputs, {'quantity': 'column', 'dims': '300,300'})
    sc = synthetic_scdata(150, inputs.options.species)
    los = LOSResult(sc, inputs, dphi=Quantity(2.0, 'deg'))
    los.simulate_data_from_inputs(sc)
    if rank == 0:
        np.savez(dest, image=image.image, packet_image=image.packet_image,
                 column=column.image, totalsource=image.totalsource,
                 atoms_per_packet=image.atoms_per_packet,
                 radiance=los.radiance.values, npackets_los=los.npackets_los.values,
                 los_totalsource=los.totalsource, sourcerate=float(los.sourcerate),
                 world=world, files=len(files), mine=mine)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
