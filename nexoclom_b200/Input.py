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
        """Run the model (reference Input.py:175-268): packets already in the
        catalogue are not re-run; the rest is integrated in chunks of
        ``packs_per_it`` (default 1e6 adaptive, 1 GiB / nsteps / 8 constant-step)."""
        from .Output import Output
        t0_ = time.time()
        distribute = distribute in (True, 'delay', 'delayed')
        if overwrite:
            self.delete_files()
            totalpackets = 0
        else:
            _, outputfiles, totalpackets, _ = self.search()
            print(f'Found {len(outputfiles)} files with {totalpackets} packets.')

        npackets = int(npackets)
        ntodo = npackets - totalpackets
        while ntodo > 0:
            if (packs_per_it is None) and (self.options.step_size == 0):
                packs_per_it = 1000000
            elif packs_per_it is None:
                nsteps = int(np.ceil(self.options.endtime.value / self.options.step_size) + 1)
                packs_per_it = np.ceil(1024**3 / nsteps / 8)
            packs_per_it = int(np.min([ntodo, packs_per_it]))
            nits = int(np.ceil(ntodo / packs_per_it))
            print('Running Model')
            print(f'Will complete {nits} iterations of {packs_per_it} packets.')
            if distribute:
                assert False, 'Dont do this'          # Input.py:235-236
            for it in range(nits):
                tit0_ = time.time()
                print(f'Starting iteration #{it + 1} of {nits}')
                # distinct packets per chunk: offset the seed like a fresh default_rng would
                chunk_seed = None if seed is None else int(seed) + it
                Output(self, packs_per_it, compress=compress, seed=chunk_seed)
                print(f'Completed iteration #{it + 1} in {time.time() - tit0_} seconds.')
            _, outputfiles, totalpackets, _ = self.search()
            print(f'Found {len(outputfiles)} files with {totalpackets} packets.')
            ntodo = npackets - totalpackets
        print(f'Model run completed in {time.time() - t0_} sec.')

    def produce_image(self, format_, overwrite=False, distribute=None):
        from .ModelImage import ModelImage
        return ModelImage(self, format_, overwrite=overwrite, distribute=distribute)

    def delete_files(self, filename=None):
        """Delete output files and remove them from the catalogue (Input.py:274-)."""
        catalogue.delete(self, filename)
