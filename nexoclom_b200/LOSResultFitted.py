"""``LOSResultFitted``: re-weight the packets of an existing ``LOSResult`` so that the modelled
radiances reproduce the data (reference ``data_simulation/LOSResultFitted.py:18-262``).

The reference walks, in Python, over every spectrum and every packet that spectrum `used`
(``iteration_unfit.used_packets``, dictionaries of lists).  Here the `used` sets arrive from
the K5 kernel as one CSR structure (``nx_los_used``; ``IterationResult.used_csr``) and both
passes -- the weighted mean of data/model ratios per initial packet, and the fitted radiance
per spectrum -- are segment reductions over that CSR (``np.add.at`` / ``np.bincount``)."""
import numpy as np
import pandas as pd

from . import catalogue
from .LOSResult import IterationResult, LOSResult
from .Output import Output
from .runsetup import RunSetup
from .units import Quantity, value_of


class IterationResultFitted(IterationResult):
    """reference compute_iteration.py:77-85."""

    def __init__(self, iteration, losresult):
        super().__init__(iteration, losresult)
        self.unfit_outputfile = iteration['unfit_outputfile']
        self.unfit_outid = iteration['unfit_outid']
        self.unfit_modelfile = iteration['unfit_modelfile']
        self.fitted = True


def fit_packet_weights(off, rows, index0, n0, spectrum_xyz, packet_xyz, ratio, mask, sigma,
                       use_weight=None):
    """Weighted mean of the data / model ratios over the spectra that used each initial
    packet (reference LOSResultFitted.py:123-180).

    off, rows: CSR of the `used` sets (per spectrum: positions into the packet table);
    index0[row] = X0 row of that packet ('Index' column); n0 = len(X0).
    Returns `weighting` (n0,): multiplier of X0.frac, normalised to mean 1 over the packets
    that were used by at least one masked spectrum (0 for the others)."""
    nspec = len(off) - 1
    sp = np.repeat(np.arange(nspec), np.diff(off))
    keep = np.asarray(mask, dtype=bool)[sp]
    sp, r = sp[keep], np.asarray(rows)[keep]
    if use_weight in ('dist2', 'dist'):
        d = np.sqrt(((packet_xyz[r] - spectrum_xyz[sp])**2).sum(axis=1))
        w = 1 / d**2 if use_weight == 'dist2' else 1 / d
    elif use_weight == 'sigma':
        w = np.ones(len(r)) / np.asarray(sigma, dtype=float)[sp] * 2      # sic (:161)
    else:
        w = np.ones(len(r))
    ind0 = np.asarray(index0)[r]
    ratio_x_sigma, sig = np.zeros(n0), np.zeros(n0)
    np.add.at(ratio_x_sigma, ind0, np.asarray(ratio, dtype=float)[sp] * w)
    np.add.at(sig, ind0, w)
    used = sig > 0
    ratio_x_sigma[used] = ratio_x_sigma[used] / sig[used]
    return ratio_x_sigma / ratio_x_sigma[used].mean()


def fitted_radiance(off, rows, spectrum_xyz, packet_xyz, weight, dphi, rp_cm):
    """Radiance of every spectrum over its `used` packets with the re-weighted packets
    (reference LOSResultFitted.py:188-203; no shadow test here: `used` packets are sunlit)."""
    nspec = len(off) - 1
    sp = np.repeat(np.arange(nspec), np.diff(off))
    r = np.asarray(rows)
    d = np.sqrt(((packet_xyz[r] - spectrum_xyz[sp])**2).sum(axis=1))
    apix = np.pi * (d * np.sin(dphi))**2 * rp_cm**2
    return np.bincount(sp, weights=np.asarray(weight)[r] / apix, minlength=nspec)


def select_one_step(X, npackets, rng):
    """`use_selected`: keep, of every trajectory, the row of ONE step time drawn at random
    from the step times present in the output (reference LOSResultFitted.py:95-113: the
    set of (packet, drawn time) pairs intersected with the (Index, time) rows).  Row labels
    are preserved, so `used` sets recorded on the full output still address the rows."""
    times = pd.unique(X['time'].values)
    pick = rng.choice(times, npackets)
    ind0 = X['Index'].values.astype(np.int64)
    return X[X['time'].values == pick[ind0]]


def restrict_csr(off, rows):
    """Drop the entries of a CSR `used` structure whose row is gone (rows < 0): the
    reference's ``[x for x in used_packets if x in packets.index]`` (:135-136, :189-190)."""
    rows = np.asarray(rows)
    ok = rows >= 0
    if ok.all():
        return np.asarray(off), rows
    nspec = len(off) - 1
    sp = np.repeat(np.arange(nspec), np.diff(off))
    cnt = np.bincount(sp[ok], minlength=nspec)
    return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64), rows[ok]


class LOSResultFitted(LOSResult):
    def __init__(self, scdata, label_for_fitted, params=None, dphi=Quantity(1., 'deg'),
                 **kwargs):
        import copy
        # the reference flips `fitted` on the unfitted result's own Input object
        # (LOSResultFitted.py:21-22); a copy keeps that result searchable afterwards
        inputs = copy.deepcopy(scdata.model_result[label_for_fitted].inputs)
        inputs.options.fitted = True
        super().__init__(scdata, inputs, params=params, dphi=dphi, **kwargs)
        self.unfitted_label = label_for_fitted
        self.unfit_outid = None
        self.unfit_outputfiles = None

    def determine_source_from_data(self, scdata, overwrite=False, use_selected=False,
                                   use_weight=None):
        """Determine the source using a previous LOSResult (reference :69-262)."""
        unfit = scdata.model_result[self.unfitted_label]
        data = scdata.data
        if overwrite:
            self.inputs.delete_files()
        setup = RunSetup(self.inputs)
        gtables = setup.gtables(self.wavelength) if self.g is None else None
        rp_cm = setup.radius_km * 1e5
        sc_xyz = data[['x', 'y', 'z']].values.astype(float)
        ratio = (data.radiance / unfit.radiance).fillna(0).values
        mcol = f'mask_{self.unfitted_label}'
        mask = data[mcol].values if mcol in data else np.asarray(unfit.mask, dtype=bool)

        print(f'LOSResultFitted: {len(unfit.outid)} unfitted files.')
        results = []
        for ufit_id, ufit_outfile in zip(unfit.outid, unfit.outputfiles):
            output = Output.restore(ufit_outfile)
            if 'Index' not in output.X.columns:
                output.X['Index'] = output.X.index
            it_unfit = unfit._iterations[ufit_outfile]
            off, idx0, labels = it_unfit.used_csr
            print(f'use_selected = {use_selected}')
            if use_selected:
                rng = getattr(output, 'randgen', None)
                if rng is None:
                    rng = np.random.default_rng(getattr(output, 'seed', None))
                output.X = select_one_step(output.X, output.npackets, rng)
            rows = output.X.index.get_indexer(labels)          # positions of the used rows in X
            assert use_selected or np.all(rows >= 0)
            off, rows = restrict_csr(off, rows)
            xyz = output.X[['x', 'y', 'z']].values
            weighting = fit_packet_weights(off, rows, output.X['Index'].values,
                                           len(output.X0), sc_xyz, xyz, ratio, mask,
                                           data.sigma.values if 'sigma' in data else None,
                                           use_weight)
            multiplier = weighting[output.X['Index'].values]
            output.X.loc[:, 'frac'] = output.X['frac'].values * multiplier
            output.X0.loc[:, 'frac'] = output.X0['frac'].values * weighting
            nsteps = getattr(output, 'nsteps', 1)
            output.totalsource = output.X0['frac'].sum() * nsteps

            radvel_sun = output.X['vy'].values + setup.vrplanet
            if self.g is not None:
                gg = np.zeros(len(output.X)) + float(value_of(self.g))
            else:
                gg = np.zeros(len(output.X))
                for v, g in gtables:
                    gg += np.interp(radvel_sun, v, g)
            weight = output.X['frac'].values * gg / 1e6         # ModelResult.py:161
            radiance = fitted_radiance(off, rows, sc_xyz, xyz, weight, self.dphi, rp_cm)

            output.inputs = self.inputs
            output.save()                                        # the fitted output
            iteration = {'radiance': pd.Series(radiance, index=data.index),
                         'npackets': output.X0.frac.sum(), 'totalsource': output.totalsource,
                         'outputfile': output.filename, 'out_idnum': output.idnum,
                         'unfit_outputfile': ufit_outfile, 'unfit_outid': ufit_id,
                         'unfit_modelfile': it_unfit.modelfile, 'included': True,
                         'query': getattr(scdata, 'query', None)}
            res = IterationResultFitted(iteration, self)
            res.weighting = weighting
            results.append(res)

        self.modelfiles = {}
        self.outputfiles = []
        self.radiance = pd.Series(np.zeros(len(data)), index=data.index)
        self.totalsource = 0.
        for res in results:
            self.radiance += res.radiance
            self.totalsource += res.totalsource
            self.modelfiles[res.outputfile] = res.modelfile
            self.outputfiles.append(res.outputfile)
            self._iterations[res.outputfile] = res
        model_rate = self.totalsource / float(value_of(self.inputs.options.endtime))
        self.atoms_per_packet = 1e23 / model_rate
        self.radiance *= self.atoms_per_packet / 1e3            # kR
        self.determine_source_rate(scdata, use_weight=False)
        self.unfit_outputfiles = list(self.modelfiles.keys())
        print(self.totalsource, self.atoms_per_packet)
