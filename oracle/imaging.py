"""Oracle: image and line-of-sight accumulation.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

NumPy restatement of
  reference nexoclom/data_simulation/ModelImage.py:229-274, 367-384 -> create_image(), image_rotation()
  reference nexoclom/math/rotation_matrix.py:5-14                   -> rotation_matrix()
  reference nexoclom/data_simulation/ModelResult.py:140-170         -> packet_weighting()
  reference nexoclom/math/histogram.py:28-39                        -> np.histogram2d (same call)
  reference nexoclom/data_simulation/compute_iteration.py:98-222    -> los_iteration()
using the same third-party routines the reference calls (np.matmul,
np.histogram2d, np.interp, sklearn KDTree.query_radius).

Pinned by: the reference holds no golden images / LOS vectors of its own
(tests/unit_tests/math/test_histogram.py and test_rotation_matrix.py are empty), so
tools/make_golden_products.py EXECUTES the unmodified reference
ModelImage.create_image() and compute_iteration() (with their packet_weighting,
interpu, Histogram2d, KD-tree ladder) on committed packet sets; create_image() and
los_iteration() below reproduce those outputs (tests/golden/image.npz, los.npz) with
bit-identical pixel / hit counts, included masks and used sets and radiance within
1e-12 (tests/test_oracle_products_golden.py).
"""
import numpy as np


def rotation_matrix(theta, axis):
    """Rodrigues rotation about ``axis`` (reference math/rotation_matrix.py:5-14)."""
    u = axis / np.linalg.norm(axis)
    lx, ly, lz = u[0], u[1], u[2]
    c, s = np.cos(theta), np.sin(theta)
    return np.asmatrix(
        [[lx**2 + (1 - lx**2) * c, lx * ly * (1 - c) + lz * s, lx * lz * (1 - c) - ly * s],
         [lx * ly * (1 - c) - lz * s, ly**2 + (1 - ly**2) * c, ly * lz * (1 - c) + lx * s],
         [lx * lz * (1 - c) + ly * s, ly * lz * (1 - c) - lx * s, lz**2 + (1 - lz**2) * c]])


def image_rotation(subobslongitude, subobslatitude):
    """Rotation taking the Sun direction (0,-1,0) to the observer direction
    (reference ModelImage.py:367-384)."""
    slong, slat = subobslongitude, subobslatitude
    p_sun = np.array([0., -1., 0.])
    p_obs = np.array([np.sin(slong) * np.cos(slat), -np.cos(slong) * np.cos(slat),
                      np.sin(slat)])
    if np.array_equal(p_sun, p_obs):
        return np.eye(3)
    costh = np.dot(p_sun, p_obs) / np.linalg.norm(p_sun) / np.linalg.norm(p_obs)
    theta = np.arccos(np.clip(costh, -1, 1))
    return rotation_matrix(theta, np.cross(p_sun, p_obs))


def packet_weighting(quantity, frac, radvel_sun, gtables, out_of_shadow=1.):
    """reference ModelResult.py:140-170.  gtables: list of (v [R_p/s], g [1/s])."""
    if quantity in ('column', 'density'):
        return frac.copy()
    gg = np.zeros(len(frac))
    for v, g in gtables:
        gg += np.interp(radvel_sun, v, g)
    return frac * out_of_shadow * gg / 1e6


def create_image(x, y, z, vy, frac, *, vrplanet, M, dims, xrange, zrange, apix, quantity,
                 gtables=()):
    """reference ModelImage.py:229-274.  Returns (image, packet_image) histograms
    BEFORE the atoms_per_packet scaling (that is ModelImage.__init__ :102-105)."""
    radvel_sun = vy + vrplanet
    pts_sun = np.stack([x, y, z], axis=1)
    pts_obs = np.array(np.matmul(M, pts_sun.transpose()).transpose())
    rho_obs = np.linalg.norm(pts_obs[:, [0, 2]], axis=1)
    inview = (rho_obs > 1) | (pts_obs[:, 1] < 0)
    frac = frac * inview
    rho_sun = np.linalg.norm(pts_sun[:, [0, 2]], axis=1)
    out_of_shadow = (rho_sun > 1) | (pts_sun[:, 1] < 0)
    weight = packet_weighting(quantity, frac, radvel_sun, gtables, out_of_shadow)
    weight = weight / apix
    pts_obs = pts_obs.transpose()
    rng = [list(xrange), list(zrange)]
    image, xe, ze = np.histogram2d(pts_obs[0, :], pts_obs[2, :], weights=weight, bins=dims,
                                   range=rng)
    packim, _, _ = np.histogram2d(pts_obs[0, :], pts_obs[2, :], bins=dims, range=rng)
    return image, packim, xe, ze


def los_iteration(x, y, z, vy, frac, los, *, vrplanet, dphi, outeredge, rp_cm, gtables,
                  used=None):
    """reference compute_iteration.py:98-222 for quantity == 'radiance'.

    los: (nlos, 6) rows x,y,z,xbore,ybore,zbore.  Returns radiance (nlos,),
    npackets (nlos,) int64, included (N,) bool, dist_from_plan (nlos,).  If ``used`` is a
    list it receives, per line of sight, the set of packet indices with weight > 0
    (reference compute_iteration.py:210-211)."""
    from sklearn.neighbors import KDTree
    los = np.asarray(los, dtype=np.float64)
    sx, sy, sz, bx, by, bz = (los[:, k] for k in range(6))
    dist_from_plan = np.sqrt(sx**2 + sy**2 + sz**2)
    ang = np.arccos((-sx * bx - sy * by - sz * bz) / dist_from_plan)
    asize_plan = np.arcsin(1. / dist_from_plan)
    dist_from_plan = dist_from_plan.copy()
    dist_from_plan[ang > asize_plan] = 1e30

    pts = np.stack([x, y, z], axis=1)
    radvel_sun = vy + vrplanet
    tree = KDTree(pts)
    nlos = los.shape[0]
    rad = np.zeros(nlos)
    npack = np.zeros(nlos, dtype=np.int64)
    included = np.zeros(len(x), dtype=bool)
    for i in range(nlos):
        x_sc = los[i, 0:3].astype(float)
        bore = los[i, 3:6].astype(float)
        b = 2 * np.sum(x_sc * bore)
        c = np.linalg.norm(x_sc)**2 - outeredge**2
        dd = (-b + np.sqrt(b**2 - 4 * 1 * c)) / 2
        t = [np.sin(dphi)]
        while t[-1] < dd:
            t.append(t[-1] + t[-1] * np.sin(dphi))
        t = np.array(t)
        xbore = x_sc[np.newaxis, :] + bore[np.newaxis, :] * t[:, np.newaxis]
        wid = t * np.sin(dphi * 2)
        ind = np.concatenate(tree.query_radius(xbore, wid))
        ilocs = np.unique(ind).astype(int)

        rel = pts[ilocs] - x_sc[np.newaxis, :]
        dist_sc = np.linalg.norm(rel, axis=1)
        losrad = np.sum(rel * bore[np.newaxis, :], axis=1)
        cosang = np.sum(rel * bore[np.newaxis, :], axis=1) / dist_sc
        cosang[cosang > 1] = 1
        angp = np.arccos(cosang)
        inview = (losrad < dist_from_plan[i]) & (angp <= dphi)
        if np.any(inview):
            sel = ilocs[inview]
            d_in = dist_sc[inview]
            l_in = losrad[inview]
            included[sel] = True
            w = packet_weighting('radiance', frac[sel], radvel_sun[sel], gtables)
            apix = np.pi * (d_in * np.sin(dphi))**2 * rp_cm**2
            wtemp = w / apix
            hit = x_sc[np.newaxis, :] + bore[np.newaxis, :] * l_in[:, np.newaxis]
            rhohit = np.linalg.norm(hit[:, [0, 2]], axis=1)
            wtemp = wtemp * ((rhohit > 1) | (hit[:, 1] < 0))
            rad[i] = wtemp.sum()
            npack[i] = np.sum(inview)
            if used is not None:
                used.append(set(int(k) for k in sel[wtemp > 0]))
        elif used is not None:
            used.append(set())
    return rad, npack, included, dist_from_plan
