"""``Input``: reads a nexoclom inputfile and runs the model.

Drop-in for the reference ``initial_state/Input.py:27-272``: same inputfile
grammar (``section.parameter = value``; ``;`` comments take priority over ``#``,
quirk Q11), same seven attribute groups, same ``run / search / produce_image /
delete_files`` signatures.  The PostgreSQL catalogue is replaced by
``nexoclom_b200.catalogue``.
"""
import os
import time

import numpy as np

from . import catalogue
from .input_classes import (Geometry, SurfaceInteraction, Forces, SpatialDist, SpeedDist,
                            AngularDist, Options)


class Input:
    def __init__(self, infile):
        self._inputfile = infile
        params = []
        if os.path.isfile(infile):
            with open(infile, 'r') as fh:
                for line in fh:
                    # strip comments: ';' wins over '#' (Input.py:63-68)
                    if ';' in line:
                        line = line[:line.find(';')]
                    elif '#' in line:
                        line = line[:line.find('#')]
                    if line.count('=') == 1:
                        param_, val_ = line.split('=')
                        if param_.count('.') == 1:
                            sec_, par_ = param_.split('.')
                            params.append((sec_.casefold().strip(), par_.casefold().strip(),
                                           val_.strip()))
        else:
            raise FileNotFoundError(infile)

        def extract_param(tag):
            return {b: c for (a, b, c) in params if a == tag}

        self.geometry = Geometry(extract_param('geometry'))
        self.surfaceinteraction = SurfaceInteraction(extract_param('surfaceinteraction'))
        self.forces = Forces(extract_param('forces'))
        self.spatialdist = SpatialDist(extract_param('spatialdist'))
        self.speeddist = SpeedDist(extract_param('speeddist'))
        self.angulardist = AngularDist(extract_param('angulardist'))
        self.options = Options(extract_param('options'))

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return False
        return all([self.geometry == other.geometry,
                    self.surfaceinteraction == other.surfaceinteraction,
                    self.forces == other.forces,
                    self.spatialdist == other.spatialdist,
                    self.speeddist == other.speeddist,
                    self.angulardist == other.angulardist,
                    self.options == other.options])

    def __repr__(self):
        return self.__str__()

    def __str__(self):
        return '\n'.join(str(g) for g in (self.geometry, self.surfaceinteraction, self.forces,
                                          self.spatialdist, self.speeddist, self.angulardist,
                                          self.options))

    def search(self):
        """(ids, filenames, npackets, totalsource) of the outputs already run for
        these inputs (reference Input.py:121-172, SQL replaced by the local catalogue)."""
        return catalogue.search(self)

    def run(self, npackets, packs_per_it=None, overwrite=False, compress=True,
            distribute=False, seed=None):
        """Run the model (reference Input.py:175-268): packets already in the catalogue are
        not re-run; the rest is integrated in chunks of ``packs_per_it``.

        Under ``torch.distributed`` (``sharding.init()``, one process per GPU) ``npackets`` is
        the size of the WHOLE run: the packets still to do are split into contiguous
        global-id ranges, one per rank, and every rank integrates its own range on its own
        GPU.  A packet is identified by its global id (the Philox counter), so the products
        do not depend on the number of GPUs nor on ``packs_per_it``; the reference hands the
        same ``seed`` to every chunk (Input.py:247), which repeats the packets when a seed is
        given -- here chunks continue the id sequence instead.

        Default ``packs_per_it``: the rank's whole share (the slabs of 1e8 packets are 20 GB
        of the 180 GB; the reference's 1e6 / 1 GiB-dense-tensor defaults are host-memory
        limits that do not apply)."""
        from .Output import Output
        from . import sharding
        t0_ = time.time()
        distribute = distribute in (True, 'delay', 'delayed')
        rank, world = sharding.rank_world()

        def global_packets():
            _, files, n_local, _ = self.search()
            if world == 1:
                return files, n_local
            tot = np.array([n_local], dtype=np.int64)
            sharding.allreduce_sum(tot)
            return files, int(tot[0])
        if overwrite:
            self.delete_files()
            totalpackets = 0
        else:
            outputfiles, totalpackets = global_packets()
            print(f'Found {len(outputfiles)} files with {totalpackets} packets.')

        npackets = int(npackets)
        ntodo = npackets - totalpackets
        if seed is None and ntodo > 0:
            seed = sharding.broadcast_object(
                int(np.random.SeedSequence().entropy & ((1 << 63) - 1)))
        while ntodo > 0:
            first, mine = sharding.shard_range(ntodo, rank, world)
            first += totalpackets                       # ids continue after earlier runs
            chunk = mine if packs_per_it is None else int(min(mine, packs_per_it))
            nits = int(np.ceil(mine / chunk)) if mine > 0 else 0
            print('Running Model')
            print(f'Will complete {nits} iterations of {chunk} packets.')
            if distribute:
                assert False, 'Dont do this'          # Input.py:235-236
            done = 0
            for it in range(nits):
                tit0_ = time.time()
                print(f'Starting iteration #{it + 1} of {nits}')
                m = min(chunk, mine - done)
                Output(self, m, compress=compress, seed=seed, first_id=first + done)
                done += m
                print(f'Completed iteration #{it + 1} in {time.time() - tit0_} seconds.')
            outputfiles, totalpackets = global_packets()
            print(f'Found {len(outputfiles)} files with {totalpackets} packets.')
            ntodo = npackets - totalpackets
        print(f'Model run completed in {time.time() - t0_} sec.')

    def produce_image(self, format_, overwrite=False, distribute=None):
        from .ModelImage import ModelImage
        return ModelImage(self, format_, overwrite=overwrite, distribute=distribute)

    def delete_files(self, filename=None):
        """Delete output files and remove them from the catalogue (Input.py:274-)."""
        catalogue.delete(self, filename)
