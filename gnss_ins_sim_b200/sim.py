"""Sim -- Monte-Carlo simulation facade with the interface of gnss_ins_sim.sim.ins_sim.Sim
(ins_sim.py:27-337: same constructor arguments, run(num_times), results(...),
get_data(names), the same data names / run keys / units), whose Monte-Carlo loops
(loop A ins_sim.py:490-506, loop B ins_algo_manager.py:73-95) run as CUDA kernels.

What stays on the CPU, by design (BASELINE north_star: "pathgen.path_gen stays CPU-side
and its reference trajectory is broadcast once"): the true trajectory.  `motion_def` is
  * a trajectory: dict / .npz path with time, ref_pos, ref_vel, ref_att, ref_accel,
    ref_gyro (what pathgen.path_gen returns, e.g. tests/golden/traj_*.npz), or
  * a motion-definition .csv / string exactly as the reference takes it; it is turned into a
    trajectory on the host by gnss_ins_sim_b200.pathgen.path_gen (C++ restatement of the
    reference's path generator, identical output, ~200x faster).

Dispatch on `algorithm`:
  * gnss_ins_sim_b200 FreeIntegration  -> K12, the fused noise+integration+error kernel;
    only per-run end-point errors, their ensemble statistics (and, on request, per-run
    process-error statistics) leave the device.  Per-run histories are materialised
    lazily: counter-based Philox makes any run reproducible in isolation, so
    get_data(['pos'])[0]['algo0_7'] re-runs just run 7 with history output on.
  * gnss_ins_sim_b200 Allan            -> K1 (noise) + K4 (Allan variance) per run block.
  * any other reference-style plugin   -> K1 generates gyro/accel on the device, the
    plugin's own .run() is called per run on the host (compatibility path).
  * None                               -> sensor data only.
Multi-GPU: runs are sharded by rank when torch.distributed is initialised (dist.py).
"""
import math
import os
import weakref
from collections.abc import Mapping

import numpy as np
import torch

from . import engine, dist, logged
from .free_integration import FreeIntegration
from .free_integration_odo import FreeIntegration as FreeIntegrationOdo
from .allan_analysis import Allan
from .ins_loose import InsLoose

D2R = math.pi / 180
R2D = 180 / math.pi
_RE = 6378137.0
_E_SQR = 0.0818191908426215 ** 2


# ------------------------------------------------------------------ helpers ---
def parse_env(env, fs):
    """Vibration DSL -> dict, as Sim.__parse_env (ins_sim.py:642-701):
    '[x y z](g|d)-random', '[x y z](g|d)-<f>Hz-sinusoidal', or (n,4) PSD array."""
    if env is None:
        return None
    if isinstance(env, np.ndarray):
        if env.ndim == 2 and env.shape[1] == 4:
            m = env.shape[0]
            if env[-1, 0] > 0.5 * fs:
                m = np.where(env[:, 0] > 0.5 * fs)[0][0]
            return {'type': 'psd', 'freq': env[:m, 0], 'x': env[:m, 1], 'y': env[:m, 2],
                    'z': env[:m, 3]}
        raise TypeError('env should be of size (n,2)')
    if not isinstance(env, str):
        raise TypeError('env should be a string or a numpy array of size (n,2)')
    text = env.lower()
    out = {}
    if 'random' in text:
        out['type'] = 'random'
        text = text.replace('-random', '')
    elif 'sinusoidal' in text:
        out['type'] = 'sinusoidal'
        text = text.replace('-sinusoidal', '')
        if text[-2:] != 'hz':
            raise ValueError('env = \'%s\' is not valid (No vib freq).' % env)
        cut = text.find('-')
        try:
            out['freq'] = math.fabs(float(text[cut + 1:-2]))
        except ValueError:
            raise ValueError('env = \'%s\' is not valid (invalid vib freq).' % env)
        text = text[:cut]
    else:
        raise ValueError('env = \'%s\' is not valid.' % env)
    scale = 1.0
    if text[-1] == 'g':
        scale, text = 9.8, text[:-1]
    elif text[-1] == 'd':
        scale, text = D2R, text[:-1]
    try:
        amp = scale * np.array(text[1:-1].split(' '), dtype='float64')
        out['x'], out['y'], out['z'] = amp[0], amp[1], amp[2]
    except (ValueError, IndexError):
        raise ValueError('Cannot convert \'%s\' to float' % env)
    return out


def lla2ecef(lla):
    """geoparams.lla2ecef_batch (geoparams.py:89-113), host side, for 'ned' error option."""
    lla = np.atleast_2d(np.asarray(lla, dtype=np.float64))
    sl, cl = np.sin(lla[:, 0]), np.cos(lla[:, 0])
    r = _RE / np.sqrt(1.0 - _E_SQR * sl * sl)
    rho = (r + lla[:, 2]) * cl
    return np.stack([rho * np.cos(lla[:, 1]), rho * np.sin(lla[:, 1]),
                     (r * (1.0 - _E_SQR) + lla[:, 2]) * sl], axis=1)


def ecef_to_ned(lat, lon):
    """attitude.ecef_to_ned (attitude.py: c_ne), rotation ECEF -> local NED."""
    sl, cl, so, co = math.sin(lat), math.cos(lat), math.sin(lon), math.cos(lon)
    return np.array([[-sl * co, -sl * so, cl], [-so, co, 0.0], [-cl * co, -cl * so, -sl]])


def euler2quat_zyx(att):
    """attitude.euler2quat 'zyx' (attitude.py:188-205), vectorised over rows: (n,3) -> (n,4)
    scalar-first.  The reference associates att_quat with every att_euler it holds
    (ins_sim.py:729-794, a per-sample Python loop that is 22 % of its run time)."""
    att = np.asarray(att, dtype=np.float64)
    c, s = np.cos(0.5 * att), np.sin(0.5 * att)
    return np.stack([c[:, 0] * c[:, 1] * c[:, 2] + s[:, 0] * s[:, 1] * s[:, 2],
                     c[:, 0] * c[:, 1] * s[:, 2] - s[:, 0] * s[:, 1] * c[:, 2],
                     c[:, 0] * s[:, 1] * c[:, 2] + s[:, 0] * c[:, 1] * s[:, 2],
                     s[:, 0] * c[:, 1] * c[:, 2] - c[:, 0] * s[:, 1] * s[:, 2]], axis=1)


class DerivedRuns(Mapping):
    """A per-run view computed from another per-run mapping on access (att_quat from att_euler)."""

    def __init__(self, src, fn):
        self._src, self._fn = src, fn

    def __iter__(self):
        return iter(self._src)

    def __len__(self):
        return len(self._src)

    def __contains__(self, key):
        return key in self._src

    def __getitem__(self, key):
        return self._fn(self._src[key])


def load_trajectory(src):
    """dict / npz path -> dict of float64 arrays with the pathgen names."""
    if isinstance(src, str):
        src = dict(np.load(src, allow_pickle=False))
    need = ('ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')
    alias = {'ref_att': 'ref_att_euler'}
    out = {}
    for k in need:
        key = k if k in src else alias.get(k, k)
        if key not in src:
            raise ValueError('trajectory is missing %r' % k)
        out[k] = np.ascontiguousarray(src[key], dtype=np.float64)
    n = out['ref_gyro'].shape[0]
    for k in need:
        if out[k].shape != (n, 3):
            raise ValueError('trajectory %s must be (n,3)' % k)
    if 'ref_odo' in src:
        out['ref_odo'] = np.ascontiguousarray(src['ref_odo'], dtype=np.float64).reshape(-1)
    if 'time' in src:
        out['time'] = np.asarray(src['time'], dtype=np.float64)
    if 'ini' in src:
        out['ini'] = np.asarray(src['ini'], dtype=np.float64)
    for k in ('gps_time', 'ref_gps', 'gps_visibility'):     # pathgen 'gps' rows, if the caller has them
        if k in src:
            out[k] = np.ascontiguousarray(src[k], dtype=np.float64)
    return out


def trajectory_from_motion_def(fs, motion_def, ref_frame, mode=None, magnetometer=False, odo=False,
                               gps=False, fs_gps=0.0):
    """Motion-definition csv/string -> trajectory dict, driven exactly as
    Sim.__gen_data_from_pathgen does (ins_sim.py:444-472): parse (ins_sim.py:578-640), then
    path_gen -- here the host-side restatement in csrc/pathgen_host.h (pathgen.py), ~200x faster
    than the reference's Python loop and identical to it to the last bits."""
    from . import pathgen
    ini_pva, cmd = pathgen.parse_motion(motion_def)
    mobility = pathgen.parse_mode(mode)
    output_def = np.array([[1.0, fs], [1.0 if gps else -1.0, fs_gps if gps else fs],
                           [1.0 if odo else -1.0, fs]])
    rtn = pathgen.path_gen(ini_pva, cmd, output_def, mobility, ref_frame, magnetometer)
    out = {'time': rtn['nav'][:, 0] / fs, 'ref_pos': np.ascontiguousarray(rtn['nav'][:, 1:4]),
           'ref_vel': np.ascontiguousarray(rtn['nav'][:, 4:7]),
           'ref_att': np.ascontiguousarray(rtn['nav'][:, 7:10]),
           'ref_accel': np.ascontiguousarray(rtn['imu'][:, 1:4]),
           'ref_gyro': np.ascontiguousarray(rtn['imu'][:, 4:7]), 'ini': ini_pva}
    if odo:
        out['ref_odo'] = np.ascontiguousarray(rtn['odo'][:, 2])
    if gps:
        out['gps_time'] = rtn['gps'][:, 0] / fs
        out['ref_gps'] = np.ascontiguousarray(rtn['gps'][:, 1:7])
        out['gps_visibility'] = np.ascontiguousarray(rtn['gps'][:, 7])
    return out


class _DataDict(dict):
    """Sim.data: a dict whose values may be registered as thunks and are built on first read."""

    class _Thunk:
        def __init__(self, fn):
            self.fn = fn

    def defer(self, key, fn):
        dict.__setitem__(self, key, _DataDict._Thunk(fn))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _DataDict._Thunk):
            v = v.fn()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


class _LazyDevice(dict):
    """{'ref_gyro', 'ref_accel', 'ref_nav'} -> CUDA tensors, uploaded when first asked for."""

    def __init__(self, sim):
        super().__init__()
        self._sim = weakref.ref(sim)     # no reference cycle: a Sim is freed (with its device arrays) when dropped

    def __missing__(self, key):
        sim = self._sim()
        src = sim._nav if key == 'ref_nav' else sim._traj[key]
        self[key] = engine.to_device(src)
        return self[key]


class LazyRuns(Mapping):
    """dict-like {key: (n,3) array} of per-run histories, materialised on first access by
    re-running the requested runs with history output (deterministic Philox streams).
    Keys are the reference's: the run index for sensor data (prefix None), '<algo>_<run>' for
    algorithm outputs.  Nothing is built per run until somebody asks (a Monte-Carlo
    experiment with 10^5 runs must not spend its time making 10^5 Python strings)."""

    def __init__(self, sim, name, count, prefix=None):
        # a weak reference: Sim.data -> LazyRuns -> Sim would be a cycle that only the cyclic collector frees,
        # and a Sim owns device and pinned buffers
        self._sim_ref, self._name, self._count, self._prefix = weakref.ref(sim), name, int(count), prefix

    @property
    def _sim(self):
        sim = self._sim_ref()
        if sim is None:
            raise ReferenceError('the Sim these run histories belong to no longer exists')
        return sim

    def _key(self, r):
        return r if self._prefix is None else '%s_%d' % (self._prefix, r)

    def _run_of(self, key):
        if self._prefix is None:
            r = key
        else:
            if not isinstance(key, str) or not key.startswith(self._prefix + '_'):
                raise KeyError(key)
            try:
                r = int(key[len(self._prefix) + 1:])
            except ValueError:
                raise KeyError(key)
        if not isinstance(r, (int, np.integer)) or not 0 <= r < self._count:
            raise KeyError(key)
        return int(r)

    def __contains__(self, key):
        try:
            self._run_of(key)
            return True
        except KeyError:
            return False

    def __iter__(self):
        return (self._key(r) for r in range(self._count))

    def __len__(self):
        return self._count

    def __getitem__(self, key):
        return self._sim._history(self._name, self._run_of(key))


# ------------------------------------------------------------------ the facade --
_UNITS = {  # name -> (description, units, output units)   ins_data_manager.py:85-216
    'att_euler': ('simulation attitude (Euler, ZYX)  from algo', ['rad'] * 3, ['deg'] * 3),
    'pos': ('simulation position from algo', ['rad', 'rad', 'm'], ['deg', 'deg', 'm']),
    'vel': ('simulation velocity from algo', ['m/s'] * 3, ['m/s'] * 3),
}


class Sim(object):
    '''
    INS Monte-Carlo simulation engine (device-backed).
    '''

    def __init__(self, fs, motion_def, ref_frame=0, imu=None, mode=None, env=None,
                 algorithm=None, seed=0, lanes_per_run=0, history_block=32, run_base=0):
        '''
        Args: as gnss_ins_sim.sim.ins_sim.Sim (ins_sim.py:31-124), plus
            seed: Philox key of the experiment (the reference is unseeded; here every
                (seed, run) pair names one reproducible noise realisation).
            lanes_per_run: CUDA lane-group width (0 = automatic).
            history_block: runs materialised together on a lazy history access.
            run_base: Philox stream id of run 0 (run r draws stream run_base + r), so that
                separate experiments can extend one ensemble without reusing streams.
        '''
        self.fs = list(fs) if isinstance(fs, (list, tuple, np.ndarray)) else [float(fs), 0.0, 0.0]
        self.imu = imu
        self.mode = mode
        self.env = env
        self.ref_frame = ref_frame if ref_frame in (0, 1) else 0
        self.seed = int(seed)
        self.lanes_per_run = int(lanes_per_run)
        self.history_block = int(history_block)
        self.run_base = int(run_base)
        self.data_src = motion_def
        self.sim_count = 1
        self.sim_complete = False
        self.sim_results = False
        self.sum = ''
        self.algo = algorithm
        if algorithm is not None and not isinstance(algorithm, (list, tuple)):
            self.algo = [algorithm]
        if self.algo is not None:
            for a in self.algo:   # InsAlgoMgr.__check_algo, ins_algo_manager.py:116-127
                try:
                    ok = len(a.input) >= 1 and len(a.output) >= 1
                except Exception:
                    ok = False
                if not ok:
                    raise ValueError('algorithm input or output is not a valid list or tuple.')
        self.data = _DataDict()  # name -> ndarray | dict-of-runs | LazyRuns (| thunk, until first read)
        self.err_stats = {}     # end-point ensemble statistics of the last run()
        self._traj = None
        self._logged = None      # data read from a logged-data directory (no sensor model)
        self._dev_cache = None
        self._cache = {}

    # ---- names ------------------------------------------------------------
    def algo_name(self, i):
        """InsAlgoMgr.get_algo_name, ins_algo_manager.py:98-114"""
        a = self.algo[i]
        return a.name if hasattr(a, 'name') else 'algo' + str(i)

    # ---- trajectory -------------------------------------------------------
    def _load_trajectory(self):
        src = self.data_src
        if isinstance(src, dict) or (isinstance(src, str) and src.endswith('.npz')):
            traj = load_trajectory(src)
        elif isinstance(src, str):
            if os.path.isdir(src):
                self._load_logged(src)
                return
            traj = trajectory_from_motion_def(self.fs[0], src, self.ref_frame, self.mode,
                                              bool(self.imu and self.imu.magnetometer),
                                              bool(self.imu and self.imu.odo),
                                              bool(self.imu and self.imu.gps) and self.fs[1] > 0,
                                              self.fs[1])
        else:
            raise TypeError('motion_def must be a trajectory dict, an .npz path or a motion '
                            'definition csv/string')
        n = traj['ref_gyro'].shape[0]
        if 'time' not in traj:
            traj['time'] = np.arange(n) / self.fs[0]
        self._traj = traj
        d = self.data
        d['fs'], d['ref_frame'], d['time'] = self.fs[0], self.ref_frame, traj['time']
        d['ref_pos'], d['ref_vel'], d['ref_att_euler'] = traj['ref_pos'], traj['ref_vel'], traj['ref_att']
        d['ref_accel'], d['ref_gyro'] = traj['ref_accel'], traj['ref_gyro']
        d.defer('ref_att_quat', lambda: euler2quat_zyx(traj['ref_att']))   # built when first read
        for k in ('ref_odo', 'gps_time', 'ref_gps', 'gps_visibility'):
            if k in traj:
                d[k] = traj[k]
        self._nav_end = np.concatenate([traj['ref_att'][-1], traj['ref_pos'][-1], traj['ref_vel'][-1]])
        self._nav_cache = None
        self._dev_cache = None
        self._ref_gps_dev = None

    @property
    def _nav(self):
        """[n][9] att,pos,vel of the true trajectory (built on first use)."""
        if self._nav_cache is None:
            t = self._traj
            self._nav_cache = np.ascontiguousarray(
                np.concatenate([t['ref_att'], t['ref_pos'], t['ref_vel']], axis=1))
        return self._nav_cache

    @property
    def _dev(self):
        """The trajectory on the device, each array uploaded on first use (the single-GPU run()
        goes through a plan that stages the host arrays itself; the Allan path never needs the
        [n][9] navigation rows)."""
        if self._dev_cache is None:
            self._dev_cache = _LazyDevice(self)
        return self._dev_cache

    # ---- run ----------------------------------------------------------------
    def run(self, num_times=1):
        '''
        run simulation.
        Args:
            num_times: run the simulation for num_times times with given IMU error model.
        '''
        self.sim_count = max(int(num_times), 1)
        if self._traj is None and self._logged is None:
            self._load_trajectory()
        if self._logged is not None:
            return self._run_logged()
        if self.imu is None:
            raise ValueError('imu must be an IMU model when data are generated from a trajectory')
        self._cache = {}
        self._all_hist = None
        self.err_stats = {}
        self._mc = {}
        self._vib_acc = parse_env(self.env['acc'], self.fs[0]) if self.env and 'acc' in self.env else None
        self._vib_gyro = parse_env(self.env['gyro'], self.fs[0]) if self.env and 'gyro' in self.env else None
        self._psd_cache = {}
        R = self.sim_count
        self._shard = dist.shard(R)
        self.data['accel'] = LazyRuns(self, 'accel', R)
        self.data['gyro'] = LazyRuns(self, 'gyro', R)
        if getattr(self.imu, 'odo', False):
            if 'ref_odo' not in self._traj:
                raise ValueError('imu.odo is on but the trajectory has no ref_odo')
            self.data['odo'] = LazyRuns(self, 'odo', R)
        if getattr(self.imu, 'gps', False) and 'ref_gps' in self._traj:
            self.data['gps'] = LazyRuns(self, 'gps', R)     # pathgen.gps_gen per run (ins_sim.py:497-500)
        if self.algo is not None:
            for i, a in enumerate(self.algo):
                if isinstance(a, FreeIntegration):     # incl. the odometer variant
                    self._run_free_integration(i, a)
                elif isinstance(a, Allan):
                    self._run_allan(i, a)
                elif isinstance(a, InsLoose):
                    self._run_ins_loose(i, a)
                else:
                    self._run_foreign(i, a)
        self.sim_complete = True

    # ---- logged data (ins_sim.py:415-451: data from files instead of pathgen) -------------------
    def _load_logged(self, path):
        """Every supported .csv of the directory in internal units; no sensor model is applied."""
        d = logged.read_data_dir(os.path.abspath(path), self.ref_frame)
        if 'time' not in d:
            for v in d.values():
                a = next(iter(v.values())) if isinstance(v, dict) else v
                d['time'] = np.arange(a.shape[0]) / self.fs[0]
                break
        self._logged = d
        self.data.update(d)
        self.data['fs'], self.data['ref_frame'] = self.fs[0], self.ref_frame
        if 'ref_att_euler' in d and 'ref_att_quat' not in d:
            self.data.defer('ref_att_quat', lambda: euler2quat_zyx(d['ref_att_euler']))
        # what the error statistics need of a trajectory
        self._traj = {k2: d[k1] for k1, k2 in (('ref_pos', 'ref_pos'), ('ref_vel', 'ref_vel'),
                                               ('ref_att_euler', 'ref_att'), ('time', 'time')) if k1 in d}

    def _logged_sets(self, name):
        """[R, n, ...] stack of the first sim_count sets (keys 0 .. R-1) of a per-run quantity."""
        sets = self._logged.get(name)
        if not isinstance(sets, dict):
            raise ValueError('the data directory holds no %s-<key>.csv' % name)
        try:
            return np.stack([sets[r] for r in range(self.sim_count)])
        except KeyError as e:
            raise ValueError('the data directory holds no %s-%s.csv' % (name, e.args[0]))

    def _run_logged(self):
        """Algorithms on the logged sets 0 .. sim_count-1 (InsAlgoMgr.run_algo over the keys,
        ins_algo_manager.py:73-95): one batched launch per algorithm, outputs keyed
        '<algo>_<key>'; end-point error statistics if the directory has the reference files."""
        import copy
        self._cache, self.err_stats, self._mc = {}, {}, {}
        R = self.sim_count
        self._shard = (0, R)
        d = self._logged
        for i, a in enumerate(self.algo or []):
            name = self.algo_name(i)
            if isinstance(a, FreeIntegration):
                self._mc[i] = {'base': a.run_times, 'end_err': None}
                if isinstance(a, FreeIntegrationOdo):
                    att, pos, vel = a.run_batch(self.ref_frame, self.fs[0], self._logged_sets('gyro'),
                                                self._logged_sets('odo'))
                else:
                    att, pos, vel = a.run_batch(self.ref_frame, self.fs[0], self._logged_sets('gyro'),
                                                self._logged_sets('accel'))
                for out, arr in (('att_euler', att), ('pos', pos), ('vel', vel)):
                    self.data[out] = dict(self.data[out]) if isinstance(self.data.get(out), dict) else {}
                    self.data[out].update({'%s_%d' % (name, r): arr[r] for r in range(R)})
                self.data['att_quat'] = DerivedRuns(self.data['att_euler'], euler2quat_zyx)
                if all(k in d for k in ('ref_att_euler', 'ref_pos', 'ref_vel')):
                    err = np.concatenate([
                        (att[:, -1] - d['ref_att_euler'][-1] + math.pi) % (2.0 * math.pi) - math.pi,
                        pos[:, -1] - d['ref_pos'][-1], vel[:, -1] - d['ref_vel'][-1]], axis=1)
                    self._mc[i]['end_err'] = err
                    self.err_stats[name] = engine.error_stats(engine.to_device(err)).cpu().numpy()
            elif isinstance(a, Allan):
                tau, ad_a, ad_g = a.run_batch(self.fs[0], self._logged_sets('accel'), self._logged_sets('gyro'))
                self.data['algo_time'] = {'%s_%d' % (name, r): tau for r in range(R)}
                self.data['ad_accel'] = {'%s_%d' % (name, r): ad_a[r] for r in range(R)}
                self.data['ad_gyro'] = {'%s_%d' % (name, r): ad_g[r] for r in range(R)}
            else:
                outs = {o: {} for o in a.output}
                for r in range(R):
                    args = []
                    for nm in a.input:
                        v = self.data[nm] if nm in self.data else None
                        if isinstance(v, dict):
                            v = v.get(r)
                        if v is None:
                            raise ValueError('algorithm input %r is not in the data directory' % nm)
                        args.append(v)
                    a.reset()
                    a.run(copy.deepcopy(args))
                    for o, v in zip(a.output, a.get_results()):
                        outs[o]['%s_%d' % (name, r)] = v
                for o in a.output:
                    self.data[o] = outs[o]
        self.sim_complete = True

    def _mc_config(self, ai, runs, r0, stats_start=-1, dump_runs=0):
        """Config for experiment runs [r0, r0+runs) of algorithm ai.  Philox streams are
        named by the experiment run index (all algorithms see the same sensor data, as in
        the reference where loop A runs once); the initial-state rule counts the plugin's
        own run_times."""
        algo = self.algo[ai]
        n = self._traj['ref_gyro'].shape[0]
        vib_gyro, vib_acc = self._vib_pair(runs, r0)
        return engine.make_mc_config(
            self.ref_frame, self.fs[0], n, runs, self.seed, self.imu.gyro_err, self.imu.accel_err,
            algo.ini_sets.shape[0], algo.ini_sets.shape[1], earth_rot=algo.earth_rot,
            run_offset=self.run_base + r0, ini_offset=self._mc[ai]['base'] + r0,
            vib_gyro=vib_gyro, vib_accel=vib_acc,
            lanes_per_run=self.lanes_per_run or algo.lanes_per_run, stats_start=stats_start,
            dump_runs=dump_runs, **self._odo_args(algo))

    def _odo_args(self, algo):
        """Odometer variant: noise model (imu.odo_err) and the true forward speed on the device."""
        if not isinstance(algo, FreeIntegrationOdo):
            return {}
        if not getattr(self.imu, 'odo', False) or 'ref_odo' not in self._traj:
            raise ValueError('free_integration_odo needs IMU(odo=True) and a trajectory with ref_odo')
        if getattr(self, '_ref_odo_dev', None) is None:
            self._ref_odo_dev = engine.to_device(self._traj['ref_odo'])
        return {'odo_err': self.imu.odo_err, 'ref_odo': self._ref_odo_dev}

    def _vib_pair(self, runs, r0):
        """(vib_gyro, vib_accel) arguments for experiment runs [r0, r0+runs): parsed dicts, or
        for the PSD model a VIB_SERIES over device series made by K5 for exactly those runs
        (time_series_from_psd is called per run and axis, pathgen.py:478-485, :541-548)."""
        out = []
        for sensor, v in ((1, self._vib_gyro), (0, self._vib_acc)):
            if v is not None and v['type'] == 'psd':
                key = (sensor, runs, r0)
                if key not in self._psd_cache:
                    n = self._traj['ref_gyro'].shape[0]
                    if len(self._psd_cache) > 8:    # evict, but never a series of THIS (runs, r0) block:
                        for k in [k for k in self._psd_cache if k[1:] != (runs, r0)]:   # its Vib is in use
                            del self._psd_cache[k]
                    self._psd_cache[key] = engine.psd_series(self.fs[0], n, runs, sensor, v, self.seed,
                                                             self.run_base + r0)
                series, N = self._psd_cache[key]
                out.append(engine.vib_series(series, N))   # the Vib keeps its series tensor alive
            else:
                out.append(v)
        return out[0], out[1]

    def _uses_psd(self):
        return any(v is not None and v['type'] == 'psd' for v in (self._vib_acc, self._vib_gyro))

    def _run_free_integration(self, i, algo):
        name = self.algo_name(i)
        lo, hi = self._shard
        self._mc[i] = {'base': algo.run_times, 'end_err': None}   # plugin's run counter at run 0
        if not isinstance(algo, FreeIntegrationOdo):
            # the plan path on this rank's shard (pinned staging, one H2D, K12, K3, one D2H);
            # with several ranks the [3][9] shard statistics are merged by one all_gather
            err, stats = np.zeros((0, 9)), np.zeros((3, 9))
            # the exchange path is chosen from the LARGEST shard, a rank-independent quantity (shards
            # differ by one run; every rank must take the same collective)
            w = dist.world()
            p2p = dist.fused_exchange(9) if w > 1 and -(-self.sim_count // w) * 9 <= (1 << 17) else None
            plan = None
            if hi > lo:
                cfg = self._mc_config(i, hi - lo, lo)
                t = self._traj
                plan = engine.get_plan(cfg.n, cfg.runs, cfg.ini_sets, cfg.ini_rows)
                if self._uses_psd():
                    # K5 wrote the vibration series on torch's current stream; the plan runs K12 on
                    # its own non-blocking stream: order the two
                    torch.cuda.current_stream().synchronize()
                err, stats = plan.run(cfg, t['ref_gyro'], t['ref_accel'], self._nav_end, algo.ini_sets,
                                      want_stats=p2p is None)
            self._mc[i]['end_err'] = err
            if p2p is not None:
                # K3x on the plan's device buffer: statistics + NVLink exchange + merge, one kernel
                # (plan.run has synchronised: the errors are final; torch's stream orders the copy)
                merged = p2p(plan.err_device_ptr() if plan else None, hi - lo).cpu().numpy().copy()
                if p2p.timed_out():      # a peer never arrived: the merge is incomplete, never use it
                    p2p.reset_timeout()
                    raise RuntimeError('statistics exchange (K3x) timed out waiting for a peer rank; '
                                       'the ensemble statistics of this run() are not available')
                self.err_stats[name] = merged
            else:
                self.err_stats[name] = dist.combine_local_stats(stats, hi - lo)
        else:
            d = self._dev
            err, stats = np.zeros((0, 9)), np.zeros((3, 9))
            if hi > lo:
                cfg = self._mc_config(i, hi - lo, lo)
                res = engine.mc_free_integration(cfg, d['ref_gyro'], d['ref_accel'], d['ref_nav'],
                                                 algo.ini_device())
                stats = engine.error_stats(res.end_err).cpu().numpy()
                err = res.end_err.cpu().numpy()
            self._mc[i]['end_err'] = err
            self.err_stats[name] = dist.combine_local_stats(stats, hi - lo)
        algo.run_times += self.sim_count
        for out in ('att_euler', 'pos', 'vel'):
            prev = self.data.get(out)
            lazy = LazyRuns(self, (i, out), self.sim_count, prefix=name)
            if isinstance(prev, _Merged):
                prev.add(lazy)
            elif isinstance(prev, LazyRuns) and prev._name[0] != i:
                self.data[out] = _Merged([prev, lazy])
            else:
                self.data[out] = lazy
        self.data['att_quat'] = DerivedRuns(self.data['att_euler'], euler2quat_zyx)

    def _noise_block(self, r0, r1, layout=engine.LAYOUT_RUN_MAJOR):
        """K1 for global-in-experiment runs [r0, r1): CUDA gyro, accel [r1-r0, n, 3]
        ([r1-r0, 3, n] with LAYOUT_CHANNEL_MAJOR)."""
        d = self._dev
        vib_gyro, vib_acc = self._vib_pair(r1 - r0, r0)
        return engine.imu_noise(self.fs[0], r1 - r0, d['ref_gyro'], d['ref_accel'],
                                self.imu.gyro_err, self.imu.accel_err, self.seed,
                                run_offset=self.run_base + r0, vib_gyro=vib_gyro,
                                vib_accel=vib_acc, layout=layout)

    def _run_allan(self, i, algo):
        """K1 (channel-major) + K4 for this rank's shard of the runs, in blocks sized to the free
        device memory; the [R, ntau, 3] deviations of all ranks are gathered (a few hundred KB)."""
        name = self.algo_name(i)
        R = self.sim_count
        lo, hi = self._shard
        n = self._traj['ref_gyro'].shape[0]
        if self._vib_acc is None and self._vib_gyro is None and n > 5040 and os.environ.get('B2INS_ALLAN_FUSED', '1') != '0':
            return self._run_allan_fused(i, name, R, lo, hi, n)
        # runs per block: K1 materialises 48 B and K4 needs ~2 B of workspace per run-sample;
        # use up to a third of the free device memory
        if torch.cuda.is_available():    # free on the device + cached by torch's allocator but unused
            free_b = (torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved()
                      - torch.cuda.memory_allocated())
        else:
            free_b = 2 ** 31
        block = max(1, min(max(hi - lo, 1), int(free_b / 3 // (n * 64)) or 1))
        tau_all, acc_blocks, gyr_blocks = None, [], []
        for r0 in range(lo, hi, block):
            r1 = min(hi, r0 + block)
            # every channel a contiguous series: K4 then streams them with bulk copies
            gyro, accel = self._noise_block(r0, r1, engine.LAYOUT_CHANNEL_MAJOR)
            tau, a, g = algo.run_batch(self.fs[0], accel, gyro, channel_major=True)
            acc_blocks.append(a)
            gyr_blocks.append(g)
            tau_all = tau
        if tau_all is None:      # a rank without runs still needs tau for the gather below
            tau_all = engine.allan_taus(n, self.fs[0])
        ntau = len(tau_all)
        a = np.concatenate(acc_blocks) if acc_blocks else np.zeros((0, ntau, 3))
        g = np.concatenate(gyr_blocks) if gyr_blocks else np.zeros((0, ntau, 3))
        if dist.world() > 1:
            both = np.concatenate([a.reshape(hi - lo, -1), g.reshape(hi - lo, -1)], axis=1)
            both = dist.gather_rows(torch.from_numpy(np.ascontiguousarray(both)), R)
            a = both[:, :ntau * 3].reshape(R, ntau, 3)
            g = both[:, ntau * 3:].reshape(R, ntau, 3)
        self.data['algo_time'] = {'%s_%d' % (name, r): tau_all for r in range(R)}
        self.data['ad_accel'] = {'%s_%d' % (name, r): a[r] for r in range(R)}
        self.data['ad_gyro'] = {'%s_%d' % (name, r): g[r] for r in range(R)}

    # ---- loosely-coupled GNSS/INS filter (K7) -------------------------------------------------
    def _ekf_inputs(self):
        """Device copies of what K7 reads beside the IMU truth: GPS truth rows, their IMU sample
        indices, visibility."""
        if getattr(self, '_ekf_dev', None) is None:
            t = self._traj
            if self.ref_frame != 0:
                raise ValueError('ins_loose works in ref_frame 0 (LLA positions, NED velocities)')
            if not (getattr(self.imu, 'gps', False) and 'ref_gps' in t and self.fs[1] > 0):
                raise ValueError('ins_loose needs IMU(gps=True), fs = [fs_imu, fs_gps, ...] and a trajectory '
                                 'with ref_gps / gps_time / gps_visibility')
            idx = np.rint(np.asarray(t['gps_time']) * self.fs[0]).astype(np.int64)
            self._ekf_dev = {'ref_gps': engine.to_device(t['ref_gps']),
                             'gps_idx': torch.from_numpy(np.ascontiguousarray(idx)).cuda(),
                             'gps_vis': engine.to_device(np.asarray(t['gps_visibility'], dtype=np.float64))}
        return self._ekf_dev

    def _ekf_launch(self, algo, r0, runs, stats_start=0, dump_runs=0, dump_stride=1):
        d, e = self._dev, self._ekf_inputs()
        ini = algo.ini if algo.ini is not None else self._traj.get('ini')
        if ini is None:
            raise ValueError('InsLoose needs ini_pos_vel_att (the trajectory carries no initial state)')
        return engine.ins_loose(self.fs[0], runs, self.seed, self.imu.gyro_err, self.imu.accel_err,
                                self.imu.gps_err, ini, d['ref_gyro'], d['ref_accel'], d['ref_nav'],
                                e['ref_gps'], e['gps_idx'], e['gps_vis'], run_offset=self.run_base + r0,
                                ini_att_std=algo.ini_att_std, earth_rot=algo.earth_rot,
                                stats_start=stats_start, dump_runs=dump_runs, dump_stride=dump_stride,
                                vel_rw=algo.vel_model_std, att_rw=algo.att_model_std)

    def _run_ins_loose(self, i, algo):
        """demo_ins_loose.py semantics, all runs of this rank in one K7 launch: end-point errors and their
        ensemble statistics (as for free integration), bias estimates, the consistency record."""
        name = self.algo_name(i)
        lo, hi = self._shard
        self._mc[i] = {'base': 0, 'end_err': None}
        err, stats, con, bias = np.zeros((0, 9)), np.zeros((3, 9)), np.zeros((0, 19)), np.zeros((0, 6))
        if hi > lo:
            start = int(round(min(30.0, 0.1 * len(self.data['time']) / self.fs[0]) * self.fs[0]))
            res = self._ekf_launch(algo, lo, hi - lo, stats_start=start)
            stats = engine.error_stats(res.end_err).cpu().numpy()
            err, con, bias = res.end_err.cpu().numpy(), res.consist.cpu().numpy(), res.end_bias.cpu().numpy()
        self._mc[i].update({'end_err': err, 'consist': con, 'end_bias': bias})
        self.err_stats[name] = dist.combine_local_stats(stats, hi - lo)
        algo.run_times += self.sim_count
        for out in ('att_euler', 'pos', 'vel', 'wb', 'ab'):
            self.data[out] = LazyRuns(self, (i, out), self.sim_count, prefix=name)

    def ekf_consistency(self, algo_index=0):
        '''
        The filter's consistency record over this rank's runs (GPS epochs after the settling time,
        after the update): {'nees': [R,3] mean NEES of the position / velocity / attitude blocks
        (expected 3 each), 'inside3': [R,15] fraction of epochs with |error| <= 3 sigma per state
        (p, v, phi, bg, ba), 'epochs': count, 'end_bias': [R,6] bias estimates at the last sample}.
        '''
        c = self._mc[algo_index]['consist']
        ep = np.maximum(c[:, 18:19], 1.0)
        return {'nees': c[:, 0:3] / ep, 'inside3': c[:, 3:18] / ep, 'epochs': int(c[0, 18]) if len(c) else 0,
                'end_bias': self._mc[algo_index]['end_bias']}

    def _run_allan_fused(self, i, name, R, lo, hi, n):
        """The Allan experiment without the series: K1 fused into K4's first level (engine.allan_mc).
        Used when no vibration model is set and the series is longer than one chunk; the only device
        memory is the decade-sum workspace (about 2 B per run-sample), so run blocks are rarely needed."""
        d = self._dev
        tau_all = engine.allan_taus(n, self.fs[0])
        ntau = len(tau_all)
        if torch.cuda.is_available():
            free_b = (torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved() - torch.cuda.memory_allocated())
        else:
            free_b = 2 ** 31
        block = max(1, min(max(hi - lo, 1), int(free_b / 2 // (n * 6 * 2)) or 1))
        parts = []
        for r0 in range(lo, hi, block):
            r1 = min(hi, r0 + block)
            avar, _ = engine.allan_mc(self.fs[0], r1 - r0, d['ref_gyro'], d['ref_accel'], self.imu.gyro_err,
                                      self.imu.accel_err, self.seed, run_offset=self.run_base + r0)
            parts.append(torch.sqrt(avar).permute(0, 2, 1).contiguous().cpu().numpy())     # [r, ntau, 6]
        both = np.concatenate(parts) if parts else np.zeros((0, ntau, 6))
        if dist.world() > 1:
            both = dist.gather_rows(torch.from_numpy(np.ascontiguousarray(both.reshape(hi - lo, -1))), R)
            both = both.reshape(R, ntau, 6)
        self.data['algo_time'] = {'%s_%d' % (name, r): tau_all for r in range(R)}
        self.data['ad_accel'] = {'%s_%d' % (name, r): both[r, :, 0:3] for r in range(R)}
        self.data['ad_gyro'] = {'%s_%d' % (name, r): both[r, :, 3:6] for r in range(R)}

    def _run_foreign(self, i, algo):
        """Reference-style plugin run on the host, sensor data from K1
        (the per-run protocol of InsAlgoMgr.run_algo, ins_algo_manager.py:73-95)."""
        import copy
        name = self.algo_name(i)
        outs = {o: {} for o in algo.output}
        static = {'fs': self.fs[0], 'ref_frame': self.ref_frame, 'time': self.data['time']}
        for r in range(self.sim_count):
            gyro, accel = self._noise_block(r, r + 1)
            per_run = {'gyro': gyro[0].cpu().numpy(), 'accel': accel[0].cpu().numpy()}
            args = []
            for nm in algo.input:
                if nm in per_run:
                    args.append(per_run[nm])
                elif nm in static:
                    args.append(static[nm])
                elif nm in self.data and not isinstance(self.data[nm], (dict, Mapping)):
                    args.append(self.data[nm])
                else:
                    raise ValueError('algorithm input %r is not generated by this engine' % nm)
            algo.reset()
            algo.run(copy.deepcopy(args))
            res = algo.get_results()
            for o, v in zip(algo.output, res):
                outs[o]['%s_%d' % (name, r)] = v
        for o in algo.output:
            self.data[o] = outs[o]

    # ---- lazy histories -------------------------------------------------------
    def _history(self, name, run):
        """(n,3) history of one run; materialises a block of neighbouring runs at once."""
        blk = run // self.history_block
        whole = getattr(self, '_all_hist', None)      # histories() has pulled every run already
        if whole is not None:
            ai_w, lo_w, arrs = whole
            key_w = name[1] if isinstance(name, tuple) and name[0] == ai_w else name
            if isinstance(key_w, str) and key_w in arrs and 0 <= run - lo_w < arrs[key_w].shape[0]:
                return arrs[key_w][run - lo_w]
        if isinstance(name, tuple):      # algorithm output (algo index, data name)
            ai, out = name
            key = ('nav', ai, blk)
            if key not in self._cache and isinstance(self.algo[ai], InsLoose):
                r0 = blk * self.history_block
                r1 = min(self.sim_count, r0 + self.history_block)
                res = self._ekf_launch(self.algo[ai], r0, r1 - r0, dump_runs=r1 - r0)
                self._cache[key] = {'att_euler': res.att.cpu().numpy(), 'pos': res.pos.cpu().numpy(),
                                    'vel': res.vel.cpu().numpy(), 'wb': res.wb.cpu().numpy(),
                                    'ab': res.ab.cpu().numpy()}
            if key not in self._cache:
                algo = self.algo[ai]
                r0 = blk * self.history_block
                r1 = min(self.sim_count, r0 + self.history_block)
                cfg = self._mc_config(ai, r1 - r0, r0, dump_runs=r1 - r0)
                d = self._dev
                res = engine.mc_free_integration(cfg, d['ref_gyro'], d['ref_accel'], d['ref_nav'],
                                                 algo.ini_device(), dump_nav=True, dump_imu=True)
                self._cache[key] = {'att_euler': res.att.cpu().numpy(), 'pos': res.pos.cpu().numpy(),
                                    'vel': res.vel.cpu().numpy()}
                imu_hist = {'gyro': res.gyro.cpu().numpy(), 'accel': res.accel.cpu().numpy()}
                if res.odo is not None:
                    imu_hist['odo'] = res.odo.cpu().numpy()
                self._cache.setdefault(('imu', blk), {}).update(imu_hist)
            return self._cache[key][out][run - blk * self.history_block]
        key = ('imu', blk)
        if key not in self._cache or name not in self._cache[key]:
            r0 = blk * self.history_block
            r1 = min(self.sim_count, r0 + self.history_block)
            hist = self._cache.setdefault(key, {})
            if name == 'odo':      # pathgen.odo_gen stream: a zero-length odometer experiment
                ai = [i for i, a in enumerate(self.algo or []) if isinstance(a, FreeIntegrationOdo)]
                if not ai:
                    raise KeyError('odo histories are produced with the free_integration_odo plugin')
                self._history((ai[0], 'pos'), run)
            elif name == 'gps':
                if getattr(self, '_ref_gps_dev', None) is None:
                    self._ref_gps_dev = engine.to_device(self._traj['ref_gps'])
                hist['gps'] = engine.gps_noise(r1 - r0, self._ref_gps_dev, self.imu.gps_err, self.ref_frame,
                                               self.seed, run_offset=self.run_base + r0).cpu().numpy()
            else:
                gyro, accel = self._noise_block(r0, r1)
                hist.update({'gyro': gyro.cpu().numpy(), 'accel': accel.cpu().numpy()})
        return self._cache[key][name][run - blk * self.history_block]

    def histories(self, algo_index=0, imu=False, stride=1, quat=False):
        '''
        Every run of this rank at once: {'att_euler', 'pos', 'vel'} -> [R_local, rows, 3] host arrays
        (plus 'gyro', 'accel' with imu=True, 'att_quat' [R_local, rows, 4] with quat=True) -- what
        the reference's Sim.run leaves in its data manager (ins_sim.py:184-187, att_quat associated
        :729-794), here by ONE launch with history output for all runs and one device-to-host copy
        per array into pinned memory.  stride > 1 keeps samples 0, stride, 2 stride, ... (rows =
        ceil(n / stride), 'time' decimated alike): error histories of many runs for plotting without
        72 bytes per run-step.  With stride 1 get_data() then serves single runs from these arrays.
        '''
        lo, hi = self._shard
        algo = self.algo[algo_index]
        if not isinstance(algo, FreeIntegration) or self._logged is not None:
            raise ValueError('histories() is for the fused free-integration experiment')
        runs = hi - lo
        n = self._traj['ref_gyro'].shape[0]
        out = {}
        if runs == 0:
            return {k: np.zeros((0, n, 3)) for k in ('att_euler', 'pos', 'vel')}
        cfg = self._mc_config(algo_index, runs, lo, dump_runs=runs)
        cfg.dump_stride = max(1, int(stride))
        d = self._dev
        pool = _hist_pool(self)      # device + pinned buffers: this Sim's, or a dead Sim's (never a live one's)
        res = engine.mc_free_integration(cfg, d['ref_gyro'], d['ref_accel'], d['ref_nav'], algo.ini_device(),
                                         dump_nav=True, dump_imu=imu, out=pool.get('res'), dump_quat=quat)
        pool['res'] = res
        pinned = pool.setdefault('pinned', {})
        names = [('att_euler', res.att), ('pos', res.pos), ('vel', res.vel)]
        if imu:
            names += [('gyro', res.gyro), ('accel', res.accel)]
        if quat:
            names += [('att_quat', res.quat)]
        for k, t in names:     # pinned staging is kept per process: pinning tens of MB costs milliseconds
            if k not in pinned or pinned[k].shape != t.shape:
                pinned[k] = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
        for k, t in names:
            pinned[k].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = {k: pinned[k].numpy() for k, _ in names}
        if cfg.dump_stride == 1:
            self._all_hist = (algo_index, lo, out)
        else:
            out['time'] = self.data['time'][::cfg.dump_stride]
        return out

    # ---- results --------------------------------------------------------------
    def get_names_of_available_data(self):
        return list(self.data.keys())

    def get_data(self, data_names):
        '''
        Get data by names (ins_sim.py:317-327): list of arrays / dicts of runs.
        '''
        return [self.data[n] if n in self.data else None for n in data_names]

    def save_data(self, data_dir, names=None, runs=None):
        """Write data to `<data_dir>/<name>[-<key>].csv` in the reference's file format
        (Sim_data.save_to_file, sim_data.py:117-165: output units, `legend (unit)` header), which
        both Sims read back as a logged-data directory.  names: data names (default: everything
        that has a file format); runs: keys of per-run data to write (default: all -- histories
        are materialised by re-running blocks of runs, mind the count)."""
        names = [n for n in self.data.keys() if n in logged.OUTPUT_FORMAT] if names is None else list(names)
        written = []
        for n in names:
            if n not in self.data or n not in logged.OUTPUT_FORMAT:
                raise KeyError('no file format for %r' % n)
            v = self.data[n]
            if isinstance(v, (dict, Mapping)):
                keys = list(v.keys()) if runs is None else [k for k in v.keys()
                                                          if k in runs or (isinstance(k, str) and
                                                                           k.rsplit('_', 1)[-1].isdigit() and
                                                                           int(k.rsplit('_', 1)[-1]) in runs)]
                v = {k: v[k] for k in keys}
            written += logged.write_data(data_dir, n, v, self.ref_frame)
        return written

    def end_point_errors(self, algo_index=0):
        '''
        [R_local, 9] end-point errors (att wrapped [rad], pos, vel) of this rank's runs
        (numpy) -- 72 bytes per run, the raw material of the ensemble statistics.
        '''
        return self._mc[algo_index]['end_err']

    def get_error_stats(self, data_name, err_stats_start=-1, angle=False, use_output_units=False,
                        extra_opt='', algo_index=0):
        '''
        InsDataMgr.get_error_stats (ins_data_manager.py:385-452) for att_euler / pos / vel.
        err_stats_start == -1: end-point statistics over runs {'max','avg','std'} (3,).
        otherwise: per-run process statistics from that time [s]: dicts keyed by run key.
        '''
        if data_name not in _UNITS:
            raise ValueError('error statistics exist for att_euler, pos and vel')
        c0 = {'att_euler': 0, 'pos': 3, 'vel': 6}[data_name]
        desc, units, out_units = _UNITS[data_name]
        if data_name == 'pos' and self.ref_frame == 1:
            units, out_units = ['m'] * 3, ['m'] * 3
        name = self.algo_name(algo_index)
        if err_stats_start == -1:
            if data_name == 'pos' and self.ref_frame == 0 and extra_opt in ('ned', 'ecef'):
                st = self._end_point_pos_stats(extra_opt, algo_index)
                units = out_units = ['m'] * 3
            else:
                s = self.err_stats[name]
                st = {'max': s[0, c0:c0 + 3].copy(), 'avg': s[1, c0:c0 + 3].copy(),
                      'std': s[2, c0:c0 + 3].copy()}
        else:
            st = self._process_stats(algo_index, err_stats_start, c0)
        if use_output_units:
            scale = np.array([R2D if (u == 'rad' and o == 'deg') else 1.0
                              for u, o in zip(units, out_units)])
            for k in ('max', 'avg', 'std'):
                if isinstance(st[k], dict):
                    st[k] = {r: v * scale for r, v in st[k].items()}
                else:
                    st[k] = st[k] * scale
            st['units'] = str(out_units)
        else:
            st['units'] = str(units)
        return st

    def _end_point_pos_stats(self, opt, algo_index):
        """'ned' / 'ecef' position error of LLA results, ins_data_manager.py:543-552."""
        err = self._mc[algo_index]['end_err']
        if dist.world() > 1:
            import torch as _t
            err = dist.gather_rows(_t.from_numpy(err), self.sim_count)
        r = self._traj['ref_pos'][-1]
        x = err[:, 3:6] + r            # end position = error + truth
        err = lla2ecef(x) - lla2ecef(r)[0]
        if opt == 'ned':
            err = err.dot(ecef_to_ned(r[0], r[1]).T)
        return {'max': np.max(np.abs(err), 0), 'avg': np.average(err, 0), 'std': np.std(err, 0)}

    def _process_stats(self, algo_index, start_s, c0):
        key = ('proc', algo_index, float(start_s))
        if key not in self._cache:
            t = self.data['time']
            idx = np.where(t >= start_s)[0]
            if idx.shape[0] == 0:
                print('err_stats_start exceeds max data points.')
                start = 0
            else:
                start = int(idx[0])
            algo = self.algo[algo_index]
            lo, hi = self._shard
            if self._logged is not None:
                # the histories are on the host already (array_error + __array_stats,
                # ins_data_manager.py:512-541, :797-808)
                nm = self.algo_name(algo_index)
                ps = np.zeros((self.sim_count, 3, 9))
                for r in range(self.sim_count):
                    k = '%s_%d' % (nm, r)
                    e = np.concatenate([
                        (self.data['att_euler'][k] - self._logged['ref_att_euler'] + math.pi) % (2.0 * math.pi)
                        - math.pi,
                        self.data['pos'][k] - self._logged['ref_pos'],
                        self.data['vel'][k] - self._logged['ref_vel']], axis=1)[start:]
                    ps[r] = np.stack([np.max(np.abs(e), 0), np.average(e, 0), np.std(e, 0)])
                self._cache[key] = ps
                return self._process_stats(algo_index, start_s, c0)
            d = self._dev
            ps = None
            if hi > lo:
                cfg = self._mc_config(algo_index, hi - lo, lo, stats_start=start)
                ps = engine.mc_free_integration(cfg, d['ref_gyro'], d['ref_accel'], d['ref_nav'],
                                                algo.ini_device()).proc_stats.reshape(hi - lo, 27)
            self._cache[key] = dist.gather_rows(ps, self.sim_count).reshape(-1, 3, 9)
        ps = self._cache[key]
        name = self.algo_name(algo_index)
        out = {'max': {}, 'avg': {}, 'std': {}}
        for r in range(self.sim_count):
            k = '%s_%d' % (name, r)
            out['max'][k] = ps[r, 0, c0:c0 + 3].copy()
            out['avg'][k] = ps[r, 1, c0:c0 + 3].copy()
            out['std'][k] = ps[r, 2, c0:c0 + 3].copy()
        return out

    def results(self, data_dir=None, err_stats_start=0, gen_kml=False, extra_opt=''):
        '''
        Simulation summary (ins_sim.py:194-251, :339-413): the configuration and the error
        statistics of att_euler / pos / vel in output units.  Returns the available data
        names.  data_dir saves summary.txt; per-run data go to .csv files on request
        (save_data: a Monte-Carlo experiment has thousands of runs); .kml export is out of scope.
        '''
        if not self.sim_complete:
            print("Sim.run() has not been called yet: nothing to summarise.")
            return None
        if gen_kml:
            raise NotImplementedError('kml export is the reference\'s kml_gen (out of scope)')
        s = '\n------------------------------------------------------------\n'
        s += 'Sample frequency of IMU: [fs] = %s Hz\n' % str(self.fs[0])
        s += 'Reference frame: %s\n' % str(self.ref_frame)
        s += 'Simulation time duration: %s s\n' % str(len(self.data['time']) / self.fs[0])
        s += 'Simulation runs: %s\n' % str(self.sim_count)
        has_mc = bool(getattr(self, '_mc', None)) and bool(self.err_stats)   # logged data may lack references
        if has_mc:
            s += '\n------------------------------------------------------------\n'
            s += 'The following are error statistics.'
            ai = sorted(self._mc.keys())[0]
            for dn in ('att_euler', 'pos', 'vel'):
                st = self.get_error_stats(dn, err_stats_start=err_stats_start,
                                          angle=(dn == 'att_euler'), use_output_units=True,
                                          extra_opt=extra_opt, algo_index=ai)
                s += '\n-----------statistics for %s (in units of %s)\n' % (_UNITS[dn][0], st['units'])
                if isinstance(st['max'], dict):
                    for run in sorted(st['max'].keys()):
                        s += '\tSimulation run %s:\n' % str(run)
                        s += '\t\t--Max error: %s\n' % str(st['max'][run])
                        s += '\t\t--Avg error: %s\n' % str(st['avg'][run])
                        s += '\t\t--Std of error: %s\n' % str(st['std'][run])
                else:
                    s += '\t--Max error: %s\n' % str(st['max'])
                    s += '\t--Avg error: %s\n' % str(st['avg'])
                    s += '\t--Std of error: %s\n' % str(st['std'])
        self.sum += s
        if dist.rank() == 0:
            print(self.sum)
            if data_dir is not None:
                os.makedirs(data_dir, exist_ok=True)
                with open(os.path.join(data_dir, 'summary.txt'), 'w') as f:
                    f.write(self.sum + '\n')
        self.sim_results = True
        return self.get_names_of_available_data()

    def plot(self, what_to_plot, sim_idx=None, opt=None, extra_opt=''):
        raise NotImplementedError('plotting is the reference\'s sim_data_plot (matplotlib), out '
                                  'of scope: use get_data() and plot the arrays')


# History staging (device buffers + pinned host tensors: tens of MB, milliseconds to pin) is handed from
# a Sim that no longer exists to the next one that asks; a live Sim keeps its own, so the arrays
# histories() returned stay valid as long as their Sim does.
_HIST_POOLS = []      # [weakref to the owning Sim or None, dict]


def _hist_pool(sim):
    import weakref
    for entry in _HIST_POOLS:
        if entry[0] is not None and entry[0]() is sim:
            return entry[1]
    for entry in _HIST_POOLS:
        if entry[0] is None or entry[0]() is None:
            entry[0] = weakref.ref(sim)
            return entry[1]
    _HIST_POOLS.append([weakref.ref(sim), {}])
    return _HIST_POOLS[-1][1]


class _Merged(Mapping):
    """Several algorithms writing the same output name: one dict view over all of them."""

    def __init__(self, parts):
        self._parts = list(parts)

    def add(self, p):
        self._parts.append(p)

    def __iter__(self):
        for p in self._parts:
            yield from p

    def __len__(self):
        return sum(len(p) for p in self._parts)

    def __getitem__(self, key):
        for p in self._parts:
            if key in p:
                return p[key]
        raise KeyError(key)
