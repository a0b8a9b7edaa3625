"""The seven parameter groups of a nexoclom inputfile.

Same attribute names, defaults, clamping and error behaviour as the reference
``initial_state/input_classes.py`` (Geometry :19, SurfaceInteraction :250,
Forces :419, SpatialDist :490, SpeedDist :702, AngularDist :905, Options :1055),
including its quirks (Q10: ``Options.resolution`` read from a file stays a
string; the ``stepsize`` alias raises KeyError).  The PostgreSQL ``insert()`` /
``search()`` methods of the reference are out of scope;
``as_dict()`` provides the key a local catalogue can hash instead.
"""
import os

import numpy as np

from .solarsystem import SSObject
from .units import Quantity


class InputError(Exception):
    """Raised when a required parameter is not included in the inputfile
    (reference ``utilities/exceptions.py:2-6``)."""

    def __init__(self, expression, message):
        super().__init__(expression, message)
        self.expression = expression
        self.message = message


def _rad(x):
    return Quantity(x, 'rad')


class _Group:
    _prefix = ''

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return False
        keys_self, keys_other = set(self.__dict__.keys()), set(other.__dict__.keys())
        if keys_self != keys_other:
            return False
        return all(_same(self.__dict__[k], other.__dict__[k]) for k in keys_self)

    def __hash__(self):
        return hash(str(self))

    def __str__(self):
        return '\n'.join(f'{self._prefix}.{k} = {v}' for k, v in self.__dict__.items())

    def as_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, SSObject):
                v = v.object
            elif isinstance(v, (set, frozenset)):
                v = sorted(o.object if isinstance(o, SSObject) else o for o in v)
            elif isinstance(v, tuple):
                v = [float(np.asarray(x)) if isinstance(x, (Quantity, float, int)) else x
                     for x in v]
            elif isinstance(v, Quantity):
                v = float(np.asarray(v))
            out[k] = v
        return out


def _same(a, b):
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    r = a == b
    return bool(np.all(r))


class Geometry(_Group):
    _prefix = 'geometry'

    def __init__(self, gparam):
        planet = gparam.get('planet', None)
        if planet is None:
            raise InputError('Geometry.__init__', 'Planet not defined in inputfile.')
        self.planet = SSObject(planet.title())

        objlist = [self.planet.object]
        if self.planet.moons is not None:
            objlist.extend([m.object for m in self.planet.moons])

        self.startpoint = gparam.get('startpoint', self.planet.object).title()
        if self.startpoint not in objlist:
            print(f'{self.startpoint} is not a valid starting point.')
            olist = '\n\t'.join(objlist)
            print(f'Valid choices are:\n\t{olist}')
            raise ValueError

        if 'objects' in gparam:
            inc = set(i.strip().title() for i in gparam['objects'].split(','))
        else:
            inc = {self.planet.object, self.startpoint}
        for i in inc:
            if i not in objlist:
                raise InputError('Geometry.__init__',
                                 f'Invalid object {i} in geometry.include')
        self.objects = set(SSObject(o) for o in inc)
        if len(self.objects) == 0:
            self.objects = None

        if 'starttime' in gparam:
            self.type = 'geometry with starttime'
            self.time = gparam['starttime'].upper()
        else:
            self.type = 'geometry without starttime'
            if len(self.planet) == 1:
                self.phi = None
            elif 'phi' in gparam:
                phi = tuple(_rad(float(p)) for p in gparam['phi'].split(','))
                nmoons = len(self.objects - {self.planet})
                if len(phi) == nmoons:
                    self.phi = phi
                else:
                    raise InputError('Geometry.__init__',
                                     'The wrong number of orbital positions was given.')
            else:
                raise InputError('Geometry.__init__', 'geometry.phi was not specified.')

            if 'subsolarpoint' in gparam:
                subs = gparam['subsolarpoint'].split(',')
                try:
                    self.subsolarpoint = (_rad(float(subs[0])), _rad(float(subs[1])))
                except Exception:
                    raise InputError('Geometry.__init__',
                                     'The format for geometry.subsolarpoint is wrong.')
            else:
                self.subsolarpoint = (_rad(0), _rad(0))

            self.taa = _rad(float(gparam.get('taa', 0.)))


class SurfaceInteraction(_Group):
    _prefix = 'surfaceinteraction'

    def __init__(self, sparam):
        sticktype = sparam['sticktype'].lower() if 'sticktype' in sparam else None
        if sticktype == 'temperature dependent':
            self.sticktype = sticktype
            if 'accomfactor' in sparam:
                self.accomfactor = float(sparam['accomfactor'])
            else:
                raise InputError('SurfaceInteraction.__init__',
                                 'surfaceinteraction.accomfactor not given.')
            if 'a' in sparam:
                A = tuple(float(a) for a in sparam['a'].split(','))
                if len(A) == 3:
                    self.A = A
                else:
                    raise InputError('SurfaceInteraction.__init__',
                                     'surfaceinteraction.A must have 3 values')
            else:
                self.A = (1.57014, -0.006262, 0.1614157)
        elif sticktype == 'surface map':
            self.sticktype = sticktype
            self.stick_mapfile = sparam.get('stick_mapfile', 'default')
            if not os.path.exists(self.stick_mapfile):
                print('Warning: stick_mapfile does not exist')
            self.stick_map = None
            self.subsolarlon = sparam.get('subsolarlon', None)
            if self.subsolarlon is not None:
                self.subsolarlon = _rad(float(self.subsolarlon))
            if 'accomfactor' in sparam:
                self.accomfactor = float(sparam['accomfactor'])
            else:
                raise InputError('SurfaceInteraction.__init__',
                                 'surfaceinteraction.accomfactor not given.')
        elif 'stickcoef' in sparam:
            self.sticktype = 'constant'
            self.stickcoef = float(sparam['stickcoef'])
            if self.stickcoef < 0:
                self.stickcoef = 0
            elif self.stickcoef > 1:
                self.stickcoef = 1
            if 'accomfactor' in sparam:
                self.accomfactor = float(sparam['accomfactor'])
            elif self.stickcoef == 1:
                self.accomfactor = None
            else:
                raise InputError('SurfaceInteraction.__init__',
                                 'surfaceinteraction.accomfactor not given.')
        else:
            self.sticktype = 'constant'
            self.stickcoef = 1.
            self.accomfactor = None


class Forces(_Group):
    _prefix = 'forces'

    def __init__(self, fparam):
        self.gravity = (_parse_bool(fparam['gravity']) if 'gravity' in fparam else True)
        self.radpres = (_parse_bool(fparam['radpres']) if 'radpres' in fparam else True)


def _parse_bool(text):
    # the reference evaluates ``bool(eval(text.title()))``
    t = text.title()
    if t in ('True', '1'):
        return True
    if t in ('False', '0', 'None'):
        return False
    raise NameError(f"name '{t}' is not defined")


class SpatialDist(_Group):
    _prefix = 'SpatialDist'

    def __init__(self, sparam):
        if 'type' in sparam:
            self.type = sparam['type']
        else:
            raise InputError('SpatialDist.__init__', 'SpatialDist.type not given')

        if self.type == 'uniform':
            self.exobase = float(sparam['exobase']) if 'exobase' in sparam else 1.
            if 'longitude' in sparam:
                lon0, lon1 = (float(l.strip()) for l in sparam['longitude'].split(','))
                lon0 = min(max(lon0, 0.), 2 * np.pi)
                lon1 = min(max(lon1, 0.), 2 * np.pi)
                self.longitude = (_rad(lon0), _rad(lon1))
            else:
                self.longitude = (_rad(0.), _rad(2 * np.pi))
            if 'latitude' in sparam:
                lat0, lat1 = (float(l.strip()) for l in sparam['latitude'].split(','))
                lat0 = min(max(lat0, -np.pi / 2), np.pi / 2)
                lat1 = min(max(lat1, -np.pi / 2), np.pi / 2)
                if lat0 > lat1:
                    raise InputError('SpatialDist.__init__',
                                     'SpatialDist.latitude[0] > SpatialDist.latitude[1]')
                self.latitude = (_rad(lat0), _rad(lat1))
            else:
                self.latitude = (_rad(-np.pi / 2), _rad(np.pi / 2))
        elif self.type == 'surface map':
            self.exobase = float(sparam['exobase']) if 'exobase' in sparam else 1.
            self.mapfile = sparam.get('mapfile', 'default')
            self.subsolarlon = sparam.get('subsolarlon', None)
            if self.subsolarlon is not None:
                self.subsolarlon = _rad(float(self.subsolarlon))
            self.coordinate_system = sparam.get('coordinate_system', 'solar-fixed')
        elif self.type == 'surface spot':
            self.exobase = float(sparam['exobase']) if 'exobase' in sparam else 1.
            if 'longitude' in sparam:
                self.longitude = _rad(float(sparam['longitude']))
            else:
                raise InputError('SpatialDist.__init__', 'SpatialDist.longitude not given.')
            if 'latitude' in sparam:
                self.latitude = _rad(float(sparam['latitude']))
            else:
                raise InputError('SpatialDist.__init__', 'SpatialDist.latitude not given.')
            if 'sigma' in sparam:
                self.sigma = _rad(float(sparam['sigma']))
            else:
                raise InputError('SpatialDist.__init__', 'SpatialDist.sigma not given.')
        elif self.type == 'fitted output':
            self.unfit_outid = -1
            self.query = None
        else:
            raise InputError('SpatialDist.__init__',
                             f'SpatialDist.type = {self.type} not defined.')


class SpeedDist(_Group):
    _prefix = 'SpeedDist'

    def __init__(self, sparam):
        self.type = sparam['type']
        kms = lambda v: Quantity(float(v), 'km/s')   # noqa: E731
        if self.type == 'gaussian':
            if 'vprob' in sparam:
                self.vprob = kms(sparam['vprob'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.vprob not given.')
            if 'sigma' in sparam:
                self.sigma = kms(sparam['sigma'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.sigma not given.')
        elif self.type == 'sputtering':
            if 'alpha' in sparam:
                self.alpha = float(sparam['alpha'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.alpha not given.')
            if 'beta' in sparam:
                self.beta = float(sparam['beta'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.beta not given.')
            if 'u' in sparam:
                self.U = Quantity(float(sparam['u']), 'eV')
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.U not given.')
        elif self.type == 'maxwellian':
            if 'temperature' in sparam:
                self.temperature = Quantity(float(sparam['temperature']), 'K')
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.temperature not given.')
        elif self.type == 'flat':
            if 'vprob' in sparam:
                self.vprob = kms(sparam['vprob'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.vprob not given.')
            if 'delv' in sparam:
                self.delv = kms(sparam['delv'])
            else:
                raise InputError('SpatialDist.__init__', 'SpeedDist.delv not given.')
        elif self.type == 'user defined':
            self.vdistfile = sparam.get('vdistfile', 'default')
        elif self.type == 'fitted output':
            self.unfit_outid = -1
            self.query = None
        else:
            assert 0, f'SpeedDist.type = {self.type} not available'


class AngularDist(_Group):
    _prefix = 'AngularDist'

    def __init__(self, aparam):
        if 'type' in aparam:
            self.type = aparam['type'].lower()
            if self.type == 'radial':
                pass
            elif self.type == 'isotropic':
                if 'azimuth' in aparam:
                    az0, az1 = (float(l.strip()) for l in aparam['azimuth'].split(','))
                    az0 = min(max(az0, 0.), 2 * np.pi)
                    az1 = min(max(az1, 0.), 2 * np.pi)
                    self.azimuth = (_rad(az0), _rad(az1))
                else:
                    self.azimuth = (_rad(0), _rad(2 * np.pi))
                if 'altitude' in aparam:
                    alt0, alt1 = (float(l.strip()) for l in aparam['altitude'].split(','))
                    alt0 = min(max(alt0, 0), np.pi / 2)
                    alt1 = min(max(alt1, 0), np.pi / 2)
                    if alt0 > alt1:
                        raise InputError('AngularDist.__init__',
                                         'AngularDist.altitude[0] > AngularDist.altitude[1]')
                    self.altitude = (_rad(alt0), _rad(alt1))
                else:
                    self.altitude = (_rad(0), _rad(np.pi / 2))
            elif self.type == '2d':
                if 'altitude' in aparam:
                    alt0, alt1 = (float(l.strip()) for l in aparam['altitude'].split(','))
                    alt0 = min(max(alt0, 0), np.pi)
                    alt1 = min(max(alt1, 0), np.pi)
                    if alt0 > alt1:
                        raise InputError('AngularDist.__init__',
                                         'AngularDist.altitude[0] > AngularDist.altitude[1]')
                    self.altitude = (_rad(alt0), _rad(alt1))
                else:
                    self.altitude = (_rad(0), _rad(np.pi))
            else:
                raise InputError('AngularDist.__init__',
                                 f'AngularDist.type = {self.type} not defined.')
        else:
            self.type = 'isotropic'
            self.azimuth = (_rad(0), _rad(2 * np.pi))
            self.altitude = (_rad(0), _rad(np.pi / 2))


class Options(_Group):
    _prefix = 'options'

    def __init__(self, oparam):
        if 'endtime' in oparam:
            self.endtime = Quantity(float(oparam['endtime']), 's')
        else:
            raise InputError('Options.__init__', 'options.endtime not specified.')

        if 'species' in oparam:
            self.species = oparam['species'].capitalize()
        elif 'atom' in oparam:
            self.species = oparam['atom'].capitalize()
        else:
            raise InputError('Options.__init__', 'options.species not specified.')

        self.lifetime = Quantity(float(oparam.get('lifetime', 0)), 's')

        if 'outeredge' in oparam:
            self.outeredge = float(oparam['outeredge'])
        elif 'outer_edge' in oparam:
            self.outeredge = float(oparam['outer_edge'])
        else:
            self.outeredge = 1e30

        if 'step_size' in oparam:
            self.step_size = float(oparam['step_size'])
        elif 'stepsize' in oparam:
            self.step_size = float(oparam['step_size'])     # KeyError, as in the reference (Q10)
        else:
            self.step_size = 0.

        if self.step_size == 0:
            self.resolution = oparam.get('resolution', 1e-4)  # stays a str if given (Q10)
        else:
            self.resolution = None

        if 'fitted' in oparam:
            self.fitted = oparam['fitted'].casefold() == 'True'.casefold()
        else:
            self.fitted = False
