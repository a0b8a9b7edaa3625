"""``Output``: initial state + packet tracking on the GPU.

Drop-in for the reference ``particle_tracking/Output.py:23-202`` -- same
constructor, same attributes (``X0, X, npackets, totalsource, unit, GM, aplanet,
vrplanet, radpres, loss_info, compress, filename, idnum``) -- with the bodies of
``surface/speed/angular_distribution`` (:160-170), ``variable_step_size_driver``
(:221-366) and ``constant_step_size_driver`` (:368-455) replaced by calls through
the C ABI (``nx_init_state``, ``nx_integrate_adaptive``, ``nx_integrate_constant``).

Extension (import mode, north star): ``X0=`` accepts a reference-generated
initial state (DataFrame / dict / (N,8) array with time,x,y,z,vx,vy,vz,frac) that
is uploaded instead of being drawn on the device.
"""
import numpy as np
import pandas as pd

from . import catalogue
from .engine import get_engine, STATE_COLS, X0_COLS
from .runsetup import RunSetup
from .units import Quantity, def_unit


class RadPres:
    """Radiation-pressure table the reference attaches as ``output.radpres``
    (velocity [R_p/s], accel [R_p/s^2]; reference Output.py:113-121).  Module level: saved
    Outputs are pickles."""
    velocity = None
    accel = None


class Output:
    def __init__(self, inputs, npackets, compress=True, run_model=True, seed=None,
                 X0=None, device=0, strict_math=False, keep_trajectory=None):
        self.inputs = inputs
        self.planet = inputs.geometry.planet
        npackets = int(npackets)
        if run_model:
            # the reference seeds numpy's PCG64 (Output.py:92); here the seed keys the
            # per-packet Philox streams.  seed=None -> fresh entropy, like default_rng(None)
            if seed is None:
                seed = int(np.random.SeedSequence().entropy & ((1 << 63) - 1))
            self.seed = int(seed)

            assert self.inputs.geometry.type != 'geometry with time', (
                'Initialization with time stamp not implemented yet.')
            self.compress = compress
            self.unit = def_unit('R_' + self.planet.object, 'length',
                                 float(self.planet.radius.value) * 1e3)

            setup = RunSetup(inputs, strict_math=strict_math)
            self._setup = setup
            self.GM = setup.GM
            self.aplanet = setup.aplanet
            self.vrplanet = setup.vrplanet
            self.loss_info = setup.loss_info
            self.radpres = None
            if inputs.forces.radpres:
                self.radpres = RadPres()
                self.radpres.velocity = setup.radpres_v
                self.radpres.accel = setup.radpres_a
            if setup.surfaceint is not None:
                self.surfaceint = setup.surfaceint

            self.npackets = npackets
            self.totalsource = float(npackets)        # sum of frac == 1 (Output.py:150)

            # The reference stops here for planets with moons (`assert False, 'Not set up'`,
            # Output.py:153-155) and for StartPoint != planet (:173-175).  Extension: moons
            # listed in geometry.objects act on the packets from circular orbits at the
            # phases geometry.phi, and a moon can be the start point.
            if setup.moons:
                print('Including the gravity of ' + ', '.join(m['name'] for m in setup.moons))

            eng = get_engine(device)
            self._engine = eng
            setup.upload(eng)

            if X0 is not None:
                cols = self._coerce_x0(X0, npackets)
                self._host_cols = None
                if self.inputs.options.step_size == 0:
                    # adaptive run on imported packets: the copy is streamed behind the
                    # integrator (nx_integrate_adaptive_host), see variable_step_size_driver
                    self._host_cols = cols
                else:
                    eng.import_state(cols)
                    self._imported = True
                self.X0 = pd.DataFrame({c: cols[k] for k, c in enumerate(STATE_COLS)})
            else:
                if self.inputs.spatialdist.type not in ('uniform', 'surface map',
                                                        'surface spot'):
                    assert 0, 'Not a valid spatial distribution type'
                sp = setup.source_params(eng)
                eng.init_state(sp, self.seed, 0, npackets)
                x0 = eng.export_x0()
                self.X0 = pd.DataFrame({c: x0[k] for k, c in enumerate(X0_COLS)})

            if self.inputs.options.step_size == 0:
                print('Running variable step size integrator.')
                self.variable_step_size_driver()
            else:
                print('Running constant step size integrator.')
                self.constant_step_size_driver(keep_trajectory)
        else:
            print('Not running anything')
            self.compress = False
            self.X0 = pd.DataFrame()
            self.X = pd.DataFrame()
            self.npackets = npackets
            self.totalsource = npackets

        self.save()

    @staticmethod
    def _coerce_x0(X0, npackets):
        if isinstance(X0, pd.DataFrame):
            cols = [np.ascontiguousarray(X0[c].values, dtype=np.float64) for c in STATE_COLS]
        elif isinstance(X0, dict):
            cols = [np.ascontiguousarray(X0[c], dtype=np.float64) for c in STATE_COLS]
        else:
            a = np.asarray(X0, dtype=np.float64)
            assert a.ndim == 2 and a.shape[1] == 8, 'X0 must be (N, 8)'
            cols = [np.ascontiguousarray(a[:, k]) for k in range(8)]
        assert len(cols[0]) == npackets, 'X0 length != npackets'
        return cols

    def __len__(self):
        return self.npackets

    def __getitem__(self, keys):
        self.X = self.X.iloc[keys]

    # ------------------------------------------------------------------
    def variable_step_size_driver(self):
        """Adaptive driver on the GPU (K2).  Semantics of reference
        Output.py:221-366, including quirks Q1-Q9."""
        eng = self._engine
        host_cols = getattr(self, '_host_cols', None)
        if host_cols is not None:
            self.attempted_steps, self.accepted_steps = eng.integrate_adaptive_host(
                host_cols, nchunks=16)
            self._host_cols = None
        else:
            self.attempted_steps, self.accepted_steps = eng.integrate_adaptive(self.npackets)
        self.kernel_ms = eng.last_kernel_ms()
        x = eng.export_state()
        X = pd.DataFrame({c: x[k] for k, c in enumerate(STATE_COLS)})
        for c in ('v', 'altitude', 'azimuth'):
            if c in self.X0:
                X[c] = self.X0[c].values
        X['lossfrac'] = np.zeros(self.npackets)
        X['step_size'] = eng.export_step()
        X['Index'] = X.index
        self.X = X
        self._finish_units()

    def constant_step_size_driver(self, keep_trajectory=None):
        """Constant-step driver + bounce on the GPU (K3).  Semantics of reference
        Output.py:368-455: every step of every packet becomes a row of ``X``
        (``totalsource *= nsteps``).  The dense (N, 8, nsteps) tensor is only
        materialised on request / for small runs.  For large runs only the final state
        is kept (``trajectory_kept = False``) and ``ModelImage`` regenerates the rows
        inside the integrator: packets and bounce deviates are pure functions of
        (seed, packet id, step), so K1 + K3 with the image fused reproduce the same rows."""
        eng = self._engine
        p = self._setup.params
        self.nsteps = int(np.ceil(p.endtime / p.step_size + 1))
        if keep_trajectory is None:
            keep_trajectory = self.npackets * self.nsteps <= 50_000_000
        traj, nsteps, self.attempted_steps = eng.integrate_constant(
            seed=self.seed, first_id=0, trajectory=keep_trajectory, n=self.npackets)
        self.kernel_ms = eng.last_kernel_ms()
        self.totalsource *= self.nsteps
        self.trajectory_kept = bool(keep_trajectory)
        self.imported_x0 = getattr(self, '_imported', False)
        if keep_trajectory:
            n = self.npackets * self.nsteps
            X = pd.DataFrame()
            X['Index'] = np.repeat(np.arange(self.npackets), self.nsteps)
            for k, c in enumerate(STATE_COLS):
                X[c] = traj[:, k, :].reshape(n)
            frac = traj[:, 7, :]
            lossfrac = np.zeros_like(frac)
            lossfrac[:, 1:] = np.cumsum(frac[:, :-1] - frac[:, 1:], axis=1)   # Q13: base = 0
            X['lossfrac'] = lossfrac.reshape(n)
            self.X = X
        else:
            x = eng.export_state()
            self.X = pd.DataFrame({c: x[k] for k, c in enumerate(STATE_COLS)})
            self.X['Index'] = self.X.index
        self._finish_units()

    def _finish_units(self):
        # "Add units back in" (Output.py:361-366, 451-455)
        self.aplanet = Quantity(self.aplanet, 'au')
        self.vrplanet = Quantity(self.vrplanet * float(self.planet.radius.value), 'km/s')
        self.GM = Quantity(self.GM, '')

    # ------------------------------------------------------------------
    def save(self):
        """Register in the local catalogue; drop frac == 0 rows if ``compress``;
        down-cast every float64 column to float32 exactly like the reference
        (Output.py:522-543, quirk Q14)."""
        if len(self.X) > 0 and self.compress:
            self.X = self.X[self.X.frac > 0]
        for frame in (self.X0, self.X):
            for column in frame:
                if frame[column].dtype == np.int64:
                    frame[column] = frame[column].astype(np.int32)
                elif frame[column].dtype == np.float64:
                    frame[column] = frame[column].astype(np.float32)
        eng = self.__dict__.pop('_engine', None)
        setup = self.__dict__.pop('_setup', None)
        catalogue.register(self.inputs, self)
        self._engine, self._setup = eng, setup

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop('_engine', None)
        d.pop('_setup', None)
        return d

    @classmethod
    def restore(cls, filename):
        """Fetch a saved Output and up-cast to 64 bit (Output.py:550-572)."""
        import copy
        output = copy.copy(catalogue.fetch(filename))
        output.X0 = output.X0.copy()
        output.X = output.X.copy()
        for frame in (output.X0, output.X):
            for column in frame:
                if frame[column].dtype == np.int32:
                    frame[column] = frame[column].astype(np.int64)
                elif frame[column].dtype == np.float32:
                    frame[column] = frame[column].astype(np.float64)
        return output
