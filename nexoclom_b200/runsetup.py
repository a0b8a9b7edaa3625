"""Per-run constants and tables: what reference ``Output.__init__``
(``particle_tracking/Output.py:102-133``) and the initial-state functions
(``initial_state/source_distribution.py``) derive from an ``Input`` before any
packet moves.  Pure host code producing the POD blocks of the C ABI.
"""
import os

import numpy as np

from ._lib import RunParams, SourceParams
from .atomicdata import RadPresConst, LossInfo, gValue, atomicmass, K_BOLTZMANN, AMU
from .solarsystem import planet_dist
from .surfaceinteraction import (SurfaceInteraction, MaxwellianDist, sputdist,
                                 thermal_speed_kms)
from .units import Quantity, value_of


class RunSetup:
    """Everything the kernels need for one Input: ``params`` (RunParams),
    radiation-pressure table, loss info, surface-interaction spline."""

    def __init__(self, inputs, strict_math=False):
        self.inputs = inputs
        planet = inputs.geometry.planet
        self.planet = planet
        self.radius_km = float(planet.radius.value)
        # GM in R_p^3/s^2 (negative)  -- Output.py:102-105
        self.GM = float(planet.GM.value) / (self.radius_km * 1e3)**3
        r, v_r = planet_dist(planet, inputs.geometry.taa)          # Output.py:108-110
        self.aplanet = float(r.value)
        self.vrplanet_kms = float(v_r.value)
        self.vrplanet = self.vrplanet_kms / self.radius_km

        opts = inputs.options
        lifetime = float(value_of(opts.lifetime))
        if lifetime <= 0:                                          # Output.py:113-118
            self.loss_info = LossInfo(opts.species, lifetime, self.aplanet)
        else:
            self.loss_info = None

        if inputs.forces.radpres:                                  # Output.py:121-128
            rp = RadPresConst(opts.species, self.aplanet)
            self.radpres_v = np.asarray(rp.velocity.value) / self.radius_km
            self.radpres_a = np.asarray(rp.accel.value) / self.radius_km
        else:
            self.radpres_v = self.radpres_a = None

        sint = inputs.surfaceinteraction
        self.surfaceint = None
        if ('stickcoef' not in sint.__dict__) or (sint.stickcoef != 1):   # Output.py:131-133
            self.surfaceint = SurfaceInteraction(inputs, nt=201, nv=101, nprob=101)

        p = RunParams()
        p.GM = self.GM
        p.vrplanet = self.vrplanet
        if lifetime > 0:
            p.loss_mode, p.loss_rate = 1, 1.0 / lifetime            # state.py:44-46
        elif self.loss_info is not None and self.loss_info.photo is not None:
            p.loss_mode, p.loss_rate = 2, float(self.loss_info.photo)    # state.py:48-52
        else:
            p.loss_mode, p.loss_rate = 0, 0.0
        p.outeredge = float(opts.outeredge)
        p.step_size = float(opts.step_size)
        p.resolution = float(opts.resolution) if opts.step_size == 0 else 0.0
        p.endtime = float(value_of(opts.endtime))
        p.gravity = int(bool(inputs.forces.gravity))
        p.radpres = int(bool(inputs.forces.radpres))
        p.sticktype = 1 if sint.sticktype == 'temperature dependent' else 0
        if sint.sticktype == 'surface map':
            assert 0                                                # SurfaceInteraction.py:23
        p.stickcoef = float(getattr(sint, 'stickcoef', 0.0))
        p.accomfactor = float(sint.accomfactor) if sint.accomfactor is not None else 0.0
        A = getattr(sint, 'A', (0., 0., 0.))
        p.stick_A[0], p.stick_A[1], p.stick_A[2] = A
        taa = float(value_of(inputs.geometry.taa))
        p.surf_t1 = 600. + 125 * (np.cos(taa) - 1) / 2.
        p.planet_radius_km = self.radius_km
        p.radpres_amax = (float(np.max(np.abs(self.radpres_a)))
                          if self.radpres_a is not None else 0.0)
        p.strict_math = int(bool(strict_math))
        self.moons = self._moons(inputs)
        p.nmoons = len(self.moons)
        for k, m in enumerate(self.moons):
            p.moon_GM[k], p.moon_a[k], p.moon_omega[k] = m['GM'], m['a'], m['omega']
            p.moon_phi[k], p.moon_r2[k] = m['phi'], m['radius']**2
        self.params = p

    def _moons(self, inputs):
        """Moons whose gravity is included (``geometry.objects``), inner first, with the
        orbital phases ``geometry.phi`` in that order (docs/nexoclom/inputfiles.rst:62-77).
        Extension beyond the reference, which asserts for planets with moons
        (Output.py:153-155): circular, prograde, equatorial orbits."""
        geo = inputs.geometry
        objs = getattr(geo, 'objects', None) or set()
        moons = sorted((o for o in objs if o.object != self.planet.object),
                       key=lambda o: float(o.a.value))
        if not moons:
            return []
        if len(moons) > 4:
            raise ValueError('at most 4 moons are supported')
        phi = getattr(geo, 'phi', None)
        if phi is None or len(phi) != len(moons):
            raise ValueError('geometry.phi must give one orbital phase per included moon')
        rp_m = self.radius_km * 1e3
        out = []
        for m, ph in zip(moons, phi):
            out.append(dict(name=m.object,
                            GM=float(m.GM.value) / rp_m**3,
                            a=float(m.a.value) / self.radius_km,
                            omega=2 * np.pi / float(m.orbperiod.to('s').value),
                            phi=float(value_of(ph)),
                            radius=float(m.radius.value) / self.radius_km))
        return out

    @property
    def spline_tck(self):
        return self.surfaceint.tck if self.surfaceint is not None else None

    def upload(self, engine):
        """Upload the per-run tables unless this very setup is what the engine holds."""
        if getattr(engine, '_uploaded_setup', None) is self:
            return
        engine.upload_tables(self.params, self.radpres_v, self.radpres_a, self.spline_tck)
        engine._uploaded_setup = self
        engine._uploaded_gtables = None

    # ---- g-value tables for radiance weighting (ModelResult.py:152-157) ----
    def gtables(self, wavelengths):
        tabs = []
        for w in wavelengths:
            gval = gValue(self.inputs.options.species, float(value_of(w)), self.aplanet)
            tabs.append((np.asarray(gval.velocity.value) / self.radius_km,
                         np.asarray(gval.g.value)))
        return tabs

    # ---- initial-state distributions -> SourceParams (+ tables) ----
    def source_params(self, engine=None):
        inputs = self.inputs
        sp = SourceParams()
        sd, vd, ad = inputs.spatialdist, inputs.speeddist, inputs.angulardist
        start = inputs.geometry.startpoint
        sp.is_planet = int(start == inputs.geometry.planet.object)   # xyz_from_lonlat(isplan)
        sp.start_is_moon = 0
        if not sp.is_planet:
            moon = [m for m in self.moons if m['name'] == start]
            if not moon:
                raise ValueError(f'StartPoint {start} must be one of geometry.objects')
            sp.start_is_moon = 1
            sp.moon_a, sp.moon_omega = moon[0]['a'], moon[0]['omega']
            sp.moon_phi, sp.moon_radius = moon[0]['phi'], moon[0]['radius']
        sp.endtime = float(value_of(inputs.options.endtime))
        sp.random_time = int(inputs.options.step_size == 0)          # Output.py:136-139
        sp.v_scale = 1.0 / self.radius_km
        sp.exobase = float(getattr(sd, 'exobase', 1.0))

        if sd.type == 'uniform':                                    # source_distribution.py:43-62
            sp.spatial_type = 0
            lat = [float(value_of(v)) for v in sd.latitude]
            sp.sinlat0, sp.sinlat1 = np.sin(lat[0]), np.sin(lat[1])
            lon = [float(value_of(v)) for v in sd.longitude]
            if lon[0] > lon[1]:
                lon = [lon[0], lon[1] + 2 * np.pi]
            sp.lon0, sp.lon1 = lon
        elif sd.type == 'surface spot':                             # source_distribution.py:96-121
            sp.spatial_type = 1
            lon0, lat0 = float(value_of(sd.longitude)), float(value_of(sd.latitude))
            sigma0 = float(value_of(sd.sigma))
            spot0 = (np.sin(lon0) * np.cos(lat0), -np.cos(lon0) * np.cos(lat0), np.sin(lat0))
            longitude = np.linspace(0, 2 * np.pi, 361)
            latitude = np.linspace(-np.pi / 2, np.pi / 2, 181)
            ptsx = np.outer(np.sin(longitude), np.cos(latitude))
            ptsy = -np.outer(np.cos(longitude), np.cos(latitude))
            ptsz = -np.outer(np.ones_like(longitude), np.sin(latitude))    # sign flip: Q19
            cosphi = ptsx * spot0[0] + ptsy * spot0[1] + ptsz * spot0[2]
            cosphi[cosphi > 1] = 1
            cosphi[cosphi < -1] = -1
            sourcemap = np.exp(-np.arccos(cosphi) / sigma0)
            sp.map_nx, sp.map_ny = sourcemap.shape
            sp.map_lat_is_sin = 0
            sp.map_fmax = float(sourcemap.max())
            if engine is not None:
                engine.upload_sourcemap(sourcemap, longitude, latitude)
            self.sourcemap = (sourcemap, longitude, latitude)
        elif sd.type == 'surface map':                              # source_distribution.py:63-95
            from .sourcemap import SourceMap
            if sd.mapfile == 'default':
                mapfile = os.path.join(os.path.dirname(__file__), 'data',
                                       f'{inputs.options.species}_surface_composition.pkl')
            else:
                mapfile = sd.mapfile
            smap = SourceMap(mapfile)
            sd.coordinate_system = smap.coordinate_system
            if 'planet' in smap.coordinate_system:
                if sd.subsolarlon is not None:
                    assert False, 'Need to verify this works'        # source_distribution.py:90
                raise ValueError('inputs.spatialdist.subsolarlon is None')
            if smap.latitude is None:
                # longitude-only map: lat = 0, lon by inverse CDF (source_distribution.py:72-76,
                # random_deviates_1d: randomdeviates.py:29-33)
                x = np.asarray(smap.longitude, dtype=float)
                f_x = np.asarray(smap.abundance, dtype=float)
                x_ = np.linspace(x.min(), x.max(), f_x.shape[0])
                cumsum = f_x.cumsum()
                cumsum -= cumsum.min()
                cumsum /= cumsum.max()
                sp.spatial_type = 2
                self.lon_table = (cumsum, x_)
                if engine is not None:
                    engine.upload_lontable(cumsum, x_)
                return self._speed_and_direction(sp, engine)
            sp.spatial_type = 1
            fmap = np.asarray(smap.abundance, dtype=float)
            xa = np.asarray(smap.longitude, dtype=float)
            ya = np.sin(np.asarray(smap.latitude, dtype=float))
            sp.map_nx, sp.map_ny = fmap.shape
            sp.map_lat_is_sin = 1
            sp.map_fmax = float(fmap.max())
            self.sourcemap = (fmap, xa, ya)
            if engine is not None:
                engine.upload_sourcemap(fmap, np.linspace(xa.min(), xa.max(), fmap.shape[0]),
                                        np.linspace(ya.min(), ya.max(), fmap.shape[1]))
        else:
            assert 0, 'Not a valid spatial distribution type'        # Output.py:164

        return self._speed_and_direction(sp, engine)

    def _speed_and_direction(self, sp, engine):
        inputs = self.inputs
        vd, ad = inputs.speeddist, inputs.angulardist
        species = inputs.options.species
        vtype = vd.type.lower()
        table = None
        if vtype == 'gaussian':                                     # source_distribution.py:141-147
            sp.speed_type = 1
            sp.vprob, sp.vsigma = float(value_of(vd.vprob)), float(value_of(vd.sigma))
        elif vtype == 'flat':                                       # :169-171
            sp.speed_type = 0
            sp.vprob, sp.delv = float(value_of(vd.vprob)), float(value_of(vd.delv))
        elif vtype == 'sputtering':                                 # :148-153
            velocity = np.linspace(.1, 50, 5000)
            table = (velocity, sputdist(velocity, float(value_of(vd.U)), vd.alpha, vd.beta,
                                        species))
        elif vtype == 'maxwellian':                                 # :154-164
            temperature = float(value_of(vd.temperature))
            if temperature != 0:
                v_th = thermal_speed_kms(temperature, species)
                velocity = np.linspace(0.1, v_th * 5, 5000)
                table = (velocity, MaxwellianDist(velocity, temperature, species))
            else:
                assert 0, 'Not implemented yet'
        elif vtype == 'user defined':                               # :172-179
            from .sourcemap import SourceMap
            from .input_classes import InputError
            if not os.path.exists(vd.vdistfile):
                raise InputError('speed_distribution', f'{vd.vdistfile} not found.')
            vdist = SourceMap(vd.vdistfile)
            table = (np.asarray(vdist.speed, dtype=float),
                     np.asarray(vdist.speed_dist, dtype=float))
        else:
            assert 0, 'Distribtuion does not exist'
        if table is not None:
            # random_deviates_1d (math/randomdeviates.py:29-33): inverse CDF by np.interp
            sp.speed_type = 2
            x, f_x = table
            x_ = np.linspace(x.min(), x.max(), f_x.shape[0])
            cumsum = f_x.cumsum()
            cumsum -= cumsum.min()
            cumsum /= cumsum.max()
            self.speed_table = (cumsum, x_)
            if engine is not None:
                engine.upload_speedtable(cumsum, x_)

        if ad.type == 'radial':                                     # :198-201
            sp.angular_type = 0
        elif ad.type == 'isotropic':                                # :202-212
            sp.angular_type = 1
            alt = [float(value_of(v)) for v in ad.altitude]
            sp.sinalt0, sp.sinalt1 = np.sin(alt[0]), np.sin(alt[1])
            az0, az1 = (float(value_of(v)) for v in ad.azimuth)
            m = (az0, az1) if az0 <= az1 else (az1, az0 + 2 * np.pi)
            sp.az0, sp.az1 = m
        elif ad.type == '2d':                                       # :213-222
            sp.angular_type = 2
            alt = [float(value_of(v)) for v in ad.altitude]
            sp.sinalt0, sp.sinalt1 = np.cos(alt[0]), np.cos(alt[1])
        else:
            assert 0, 'Angular Distribution not defined.'
        return sp


_setups = {}


def get_setup(inputs, strict_math=False):
    """RunSetup of `inputs`, cached on the CONTENT of the inputs (they are mutable): building
    the radiation-pressure / accommodation tables costs tens of milliseconds of host time,
    more than integrating a million packets."""
    from .catalogue import input_key
    key = (input_key(inputs), bool(strict_math))
    setup = _setups.get(key)
    if setup is None:
        if len(_setups) > 32:
            _setups.pop(next(iter(_setups)))
        setup = _setups[key] = RunSetup(inputs, strict_math=strict_math)
    else:
        setup.inputs = inputs
    return setup
