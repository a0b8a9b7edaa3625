"""Oracle: the packet re-weighting of LOSResultFitted.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

Loop-for-loop restatement of reference nexoclom/data_simulation/LOSResultFitted.py:123-203
(dictionaries / per-spectrum Python loops, pandas replaced by plain arrays).

Pinned: tools/make_golden_products.py (`losfit`) EXECUTES the unmodified reference method
``LOSResultFitted.determine_source_from_data`` (its PostgreSQL search answered "nothing saved",
Output.restore / the unfitted-iteration pickle / IterationResultFitted replaced by in-memory
stand-ins) for use_weight in (None, 'dist', 'dist2', 'sigma') on the packets, lines of sight and
`used` sets of tests/golden/los.npz, and once with use_selected=True on a constant-step-like
output -> tests/golden/losfit.npz; this restatement and the
vectorised product code reproduce the re-weighted packets and the fitted radiances to 1e-12
(tests/test_losfit.py::test_reweighting_vs_reference_golden).
"""
import numpy as np


def fit(used_packets, packet_index, packet_xyz, packet_frac, packet_gsum, n0, spectra_xyz,
        data_radiance, model_radiance, mask, sigma, use_weight, dphi, rp_cm):
    """used_packets: list (per spectrum) of lists of packet rows.  Returns (weighting (n0,),
    new packet frac, fitted radiance per spectrum)."""
    nspec = len(used_packets)
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.asarray(data_radiance, dtype=float) / np.asarray(model_radiance, dtype=float)
    ratio[np.isnan(ratio)] = 0                                   # ratio.fillna(0) (:132)
    ratio_x_sigma = np.zeros(n0)
    sig = np.zeros(n0)
    for spnum in range(nspec):
        if not mask[spnum]:
            continue
        to_use = list(used_packets[spnum])
        if use_weight in ('dist2', 'dist'):
            sc_dist = np.sqrt(((packet_xyz[to_use] - spectra_xyz[spnum])**2).sum(axis=1))
            weight = 1 / sc_dist**2 if use_weight == 'dist2' else 1 / sc_dist
        elif use_weight == 'sigma':
            weight = np.ones(len(to_use)) / sigma[spnum] * 2
        else:
            weight = np.ones(len(to_use))
        for k, tu in enumerate(to_use):
            ind0 = packet_index[tu]
            ratio_x_sigma[ind0] += ratio[spnum] * weight[k]
            sig[ind0] += weight[k]
    used = sig > 0
    ratio_x_sigma[used] = ratio_x_sigma[used] / sig[used]
    weighting = ratio_x_sigma / ratio_x_sigma[used].mean()
    frac = packet_frac * weighting[packet_index]
    w = frac * packet_gsum / 1e6
    radiance = np.zeros(nspec)
    for spnum in range(nspec):
        to_use = list(used_packets[spnum])
        if len(to_use) > 0:
            d = np.linalg.norm(packet_xyz[to_use] - spectra_xyz[spnum][np.newaxis, :], axis=1)
            apix = np.pi * (d * np.sin(dphi))**2 * rp_cm**2
            radiance[spnum] = (w[to_use] / apix).sum()
    return weighting, frac, radiance
