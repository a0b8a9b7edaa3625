"""Oracle: initial packet state.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

NumPy restatement of
  reference nexoclom/initial_state/source_distribution.py:12-34   -> xyz_from_lonlat()
  reference nexoclom/initial_state/source_distribution.py:37-134  -> surface position
  reference nexoclom/initial_state/source_distribution.py:137-189 -> speed
  reference nexoclom/initial_state/source_distribution.py:192-252 -> direction
  reference nexoclom/math/randomdeviates.py:8-83                  -> inverse-CDF / rejection
  reference nexoclom/particle_tracking/Output.py:136-147          -> time, frac
as a pure TRANSFORM of uniform deviates, so it can be driven either by NumPy's
generator (statistical checks, CPU baseline inputs) or by the same
Philox4x32-10 counters the CUDA kernel uses (exact parity of K1).

Pinned by: tests/golden/source_distribution.npz holds the outputs of the unmodified
reference surface_distribution / speed_distribution / angular_distribution together
with every random deviate they drew (tools/make_golden_products.py); transform()
replayed on those deviates reproduces them to < 5e-15 for five source configurations
(tests/test_oracle_products_golden.py), and so does the kernel's own transform
(csrc/nx_init.cuh, host build).

Philox4x32-10 is the published algorithm of Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3" (SC'11); counter = (id_lo, id_hi, draw, stream),
key = (seed_lo, seed_hi).  Known-answer vectors from the Random123 distribution
are checked in tests/test_oracle_golden.py.
"""
import numpy as np

from .tracking import local_frame_direction

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments uint32 arrays (or scalars)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & _MASK).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(a, b):
    """53-bit uniform in [0,1) from two uint32 words (NumPy Generator.random layout)."""
    return ((a.astype(np.uint64) >> np.uint64(5)) * np.uint64(67108864) +
            (b.astype(np.uint64) >> np.uint64(6))).astype(np.float64) / 9007199254740992.0


STREAM_INIT, STREAM_BOUNCE = 0, 1


def uniform_pair(seed, ids, stream, draw):
    ids = np.asarray(ids, dtype=np.uint64)
    w = philox4x32_10((ids & _MASK).astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                      np.asarray(draw, dtype=np.uint32), np.uint32(stream),
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return u53(w[0], w[1]), u53(w[2], w[3])


def bounce_uniforms(seed, first_id=0):
    """``uniforms(step_index, packet_indices)`` callback for
    oracle.tracking.integrate_constant reproducing the kernel's bounce stream."""
    def fn(ct, idx):
        ids = np.asarray(idx, dtype=np.uint64) + np.uint64(first_id)
        u_alt, u_az = uniform_pair(seed, ids, STREAM_BOUNCE, 2 * ct)
        u_prob, _ = uniform_pair(seed, ids, STREAM_BOUNCE, 2 * ct + 1)
        return u_alt, u_az, u_prob
    return fn


def xyz_from_lonlat(lon, lat, isplan, exobase):
    """reference source_distribution.py:12-34."""
    if isplan:
        x0 = exobase * np.sin(lon) * np.cos(lat)
    else:
        x0 = -exobase * np.sin(lon) * np.cos(lat)
    y0 = -exobase * np.cos(lon) * np.cos(lat)
    z0 = exobase * np.sin(lat)
    X0 = np.array([x0, y0, z0])
    assert np.all(np.isfinite(X0)), 'Non-Finite values of X0'
    return X0


def _bilinear(fmap, xa, ya, x, y):
    from scipy import interpolate
    return interpolate.interpn((xa, ya), fmap, (x, y))


def pooled_rejection(fmap, xa, ya, fmax, rounds, n):
    """reference math/randomdeviates.py:61-72 literally: every round draws n candidate
    (ux, uy, uf) triples, the accepted ones are POOLED in order and the first n kept.
    (transform() below gives every packet its own candidate sequence instead -- the same
    distribution, needed for a counter-based generator; this function exists so that the
    acceptance rule and the map interpolation can be replayed against the reference's
    recorded draws.)  rounds: iterable of (ux, uy, uf) uniform arrays."""
    xa_ = np.linspace(xa.min(), xa.max(), fmap.shape[0])
    ya_ = np.linspace(ya.min(), ya.max(), fmap.shape[1])
    xs, ys = [], []
    for ux, uy, uf in rounds:
        x = ux * (xa.max() - xa.min()) + xa.min()
        y = uy * (ya.max() - ya.min()) + ya.min()
        ok = uf * fmax < _bilinear(fmap, xa_, ya_, x, y)
        xs.extend(x[ok])
        ys.extend(y[ok])
        if len(xs) >= n:
            break
    return np.array(xs[:n]), np.array(ys[:n])


def transform(sp, u, sourcemap=None, speed_table=None, map_uniforms=None, lonlat=None,
              lon_table=None):
    """Uniform deviates -> X0 (N,14).

    sp: the SourceParams numbers (nexoclom_b200._lib.SourceParams or any object
    with the same attributes); u: dict of (N,) arrays 'time','sinlat','lon',
    'speed','alt','az','g0','g1'.  For map sampling ``map_uniforms(k)`` returns
    the k-th (ux, uy, uf) triple per packet.
    columns: time,x,y,z,vx,vy,vz,frac,v,longitude,latitude,local_time,altitude,azimuth"""
    n = len(u['time'])
    time = u['time'] * sp.endtime if sp.random_time else np.zeros(n) + sp.endtime

    if lonlat is not None:
        lon, lat = lonlat              # positions sampled elsewhere (pooled_rejection)
    elif sp.spatial_type == 0:
        sinlat = sp.sinlat0 + (sp.sinlat1 - sp.sinlat0) * u['sinlat']
        lat = np.arcsin(sinlat)
        lon = (sp.lon0 + (sp.lon1 - sp.lon0) * u['lon']) % (2 * np.pi)
    elif sp.spatial_type == 2:
        # longitude-only source map (source_distribution.py:72-76): lat = 0,
        # lon = random_deviates_1d = np.interp(U, cumsum, linspace) (randomdeviates.py:29-33)
        cdf, xl = lon_table
        lon = np.interp(u['lon'], cdf, xl)
        lat = np.zeros(n)
    else:
        fmap, xa, ya = sourcemap
        xa_ = np.linspace(xa.min(), xa.max(), fmap.shape[0])
        ya_ = np.linspace(ya.min(), ya.max(), fmap.shape[1])
        lon = np.full(n, np.nan)
        yv = np.full(n, np.nan)
        todo = np.ones(n, dtype=bool)
        k = 0
        while todo.any():
            ux, uy, uf = map_uniforms(k)
            x = ux * (xa.max() - xa.min()) + xa.min()
            y = uy * (ya.max() - ya.min()) + ya.min()
            val = _bilinear(fmap, xa_, ya_, x, y)
            ok = todo & (uf * sp.map_fmax < val)
            lon[ok], yv[ok] = x[ok], y[ok]
            todo &= ~ok
            k += 1
        lat = np.arcsin(yv) if sp.map_lat_is_sin else yv

    X_ = xyz_from_lonlat(lon, lat, bool(sp.is_planet), sp.exobase)
    local_time = (lon * 12 / np.pi + 12) % 24

    if sp.speed_type == 0:
        v0 = u['speed'] * 2 * sp.delv + sp.vprob - sp.delv
    elif sp.speed_type == 1:
        if sp.vsigma == 0.:
            v0 = np.zeros(n) + sp.vprob
        else:
            if 'normal' in u:          # standard normal deviates supplied directly
                z = u['normal']
            else:
                z = np.sqrt(-2.0 * np.log(1.0 - u['g0'])) * np.cos(2 * np.pi * u['g1'])
            v0 = z * sp.vsigma + sp.vprob
    else:
        cdf, xv = speed_table
        v0 = np.interp(u['speed'], cdf, xv)
    v = v0 * sp.v_scale

    pos = X_.T.copy()
    if sp.angular_type == 2:
        # '2d' (source_distribution.py:213-222, 253-283)
        cosalt = u['alt'] * (sp.sinalt1 - sp.sinalt0) + sp.sinalt0
        alt = np.arccos(cosalt)
        az = np.zeros(n)
        v_rad, v_tan = np.sin(alt), np.cos(alt)
        rad = pos[:, :2] / np.sqrt(pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1])[:, None]
        tan = np.stack([pos[:, 1], -pos[:, 0]], axis=1)
        tan = tan / np.sqrt(tan[:, 0]**2 + tan[:, 1]**2)[:, None]
        d2 = v_tan[:, None] * tan + v_rad[:, None] * rad
        d = np.concatenate([d2, np.zeros((n, 1))], axis=1)
    elif sp.angular_type == 0:
        # alt = pi/2 exactly: sin/cos evaluated on alt itself as the reference does
        alt = np.zeros(n) + np.pi / 2.
        az = np.zeros(n)
        k = n
        v_rad = np.sin(alt)
        v_tan0 = np.cos(alt) * np.cos(az)
        v_tan1 = np.cos(alt) * np.sin(az)
        x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
        rad = pos / np.sqrt((x * x + y * y) + z * z)[:, None]
        east = np.stack([y, -x, np.zeros(k)], axis=1)
        east = east / np.sqrt((east[:, 0]**2 + east[:, 1]**2) + east[:, 2]**2)[:, None]
        north = np.stack([-z * x, -z * y, x**2 + y**2], axis=1)
        north = north / np.sqrt((north[:, 0]**2 + north[:, 1]**2) + north[:, 2]**2)[:, None]
        d = v_tan0[:, None] * north + v_tan1[:, None] * east + v_rad[:, None] * rad
    else:
        sinalt = u['alt'] * (sp.sinalt1 - sp.sinalt0) + sp.sinalt0
        alt = np.arcsin(sinalt)
        az = sp.az0 + (sp.az1 - sp.az0) * u['az']
        d = local_frame_direction(pos, sinalt, az)

    out = np.empty((n, 14))
    out[:, 0] = time
    out[:, 1:4] = pos
    out[:, 4:7] = d * v[:, None]
    if getattr(sp, 'start_is_moon', 0):
        # extension (no reference code): satellite-local frame turned by the moon's phase at
        # the ejection time, moved to the moon, orbital velocity added
        phi = sp.moon_phi - sp.moon_omega * time
        c, s_ = np.cos(phi), np.sin(phi)
        lx, ly = out[:, 1] * sp.moon_radius, out[:, 2] * sp.moon_radius
        lvx, lvy = out[:, 4].copy(), out[:, 5].copy()
        vorb = sp.moon_a * sp.moon_omega
        out[:, 1] = (lx * c - ly * s_) - sp.moon_a * s_
        out[:, 2] = (lx * s_ + ly * c) + sp.moon_a * c
        out[:, 3] = out[:, 3] * sp.moon_radius
        out[:, 4] = (lvx * c - lvy * s_) - vorb * c
        out[:, 5] = (lvx * s_ + lvy * c) - vorb * s_
    out[:, 7] = 1.0
    out[:, 8] = v
    out[:, 9] = lon
    out[:, 10] = lat
    out[:, 11] = local_time
    out[:, 12] = alt
    out[:, 13] = az
    return out


def draw_x0(setup, n, seed, first_id=0, rng='philox'):
    """X0 (N,14) for a nexoclom_b200 RunSetup.  rng='philox' reproduces K1's
    counters; rng='numpy' draws in the reference's order from default_rng(seed)."""
    sp = setup.source_params(None)
    if rng == 'philox':
        ids = np.arange(first_id, first_id + n, dtype=np.uint64)
        u = {}
        u['time'], u['sinlat'] = uniform_pair(seed, ids, STREAM_INIT, 0)
        u['lon'], u['speed'] = uniform_pair(seed, ids, STREAM_INIT, 1)
        u['alt'], u['az'] = uniform_pair(seed, ids, STREAM_INIT, 2)
        u['g0'], u['g1'] = uniform_pair(seed, ids, STREAM_INIT, 3)

        def map_uniforms(k):
            ux, uy = uniform_pair(seed, ids, STREAM_INIT, 4 + 2 * k)
            uf, _ = uniform_pair(seed, ids, STREAM_INIT, 5 + 2 * k)
            return ux, uy, uf
    else:
        g = np.random.default_rng(seed)
        # reference draw order: time, sinlat, lon, speed, sinalt, az
        u = {k: g.random(n) for k in ('time', 'sinlat', 'lon', 'speed', 'alt', 'az',
                                       'g0', 'g1')}

        def map_uniforms(k):
            return g.random(n), g.random(n), g.random(n)
    return transform(sp, u, getattr(setup, 'sourcemap', None),
                     getattr(setup, 'speed_table', None), map_uniforms,
                     lon_table=getattr(setup, 'lon_table', None))
