"""Device-level Python API over the C ABI (include/b2ins.h).

PyTorch is plumbing here: it owns device memory (float64 CUDA tensors) and the current
stream; every kernel is in csrc/libb2ins.so.  All functions are asynchronous on the
current torch stream unless they return host data.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import LAYOUT_RUN_MAJOR, LAYOUT_TIME_MAJOR, LAYOUT_CHANNEL_MAJOR  # noqa: F401


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.B2insError('no CUDA device: the b2ins engine has no CPU fallback')


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), 'need contiguous cuda f64'
    return ctypes.c_void_p(t.data_ptr())


def to_device(a, device=None):
    """numpy / tensor -> contiguous float64 CUDA tensor (H2D copy if needed)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device or 'cuda', dtype=torch.float64).contiguous()
    a = np.ascontiguousarray(a, dtype=np.float64)
    t = torch.from_numpy(a)
    return t.to(device or 'cuda', non_blocking=t.is_pinned())


def ini_sets_from_plugin(ini_pos_vel_att):
    """FreeIntegration's ini_pos_vel_att ((9|10,) or (9|10, S)) -> [S][rows] row-major."""
    a = np.asarray(ini_pos_vel_att, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    elif a.ndim != 2:
        raise ValueError('Initial states should be a 1D or 2D numpy array, '
                         'but the dimension is %s.' % a.ndim)
    if a.shape[0] not in (9, 10):
        raise ValueError('Initial states need 9 (or 10 with gravity) rows, got %d' % a.shape[0])
    return np.ascontiguousarray(a.T)


def free_integration(ref_frame, fs, gyro, accel, ini, earth_rot=True, layout=LAYOUT_RUN_MAJOR,
                     run_offset=0, lanes_per_run=0):
    """K2.  gyro, accel: CUDA f64 [R,n,3] (RUN_MAJOR) or [n,3,R] (TIME_MAJOR);
    ini: CUDA f64 [S,9|10].  Returns att, pos, vel in the same layout."""
    _require_cuda()
    lib = _lib.load()
    if layout == LAYOUT_RUN_MAJOR:
        R, n, three = gyro.shape
    else:
        n, three, R = gyro.shape
    assert three == 3 and gyro.shape == accel.shape
    att = torch.empty_like(gyro)
    pos = torch.empty_like(gyro)
    vel = torch.empty_like(gyro)
    _lib.check(lib.b2ins_free_integration_f64(
        int(ref_frame), float(fs), R, n, _ptr(gyro), _ptr(accel), layout, _ptr(ini),
        ini.shape[0], ini.shape[1], int(run_offset), int(bool(earth_rot)),
        _ptr(att), _ptr(pos), _ptr(vel), int(lanes_per_run), _stream()))
    return att, pos, vel


def free_integration_odo(ref_frame, fs, gyro, odo, ini, earth_rot=True, layout=LAYOUT_RUN_MAJOR,
                         run_offset=0, lanes_per_run=0):
    """K2, odometer variant.  gyro [R,n,3] / odo [R,n] (RUN_MAJOR) or [n,3,R] / [n,R]."""
    _require_cuda()
    lib = _lib.load()
    if layout == LAYOUT_RUN_MAJOR:
        R, n, _ = gyro.shape
        assert tuple(odo.shape) == (R, n)
    else:
        n, _, R = gyro.shape
        assert tuple(odo.shape) == (n, R)
    att = torch.empty_like(gyro)
    pos = torch.empty_like(gyro)
    vel = torch.empty_like(gyro)
    _lib.check(lib.b2ins_free_integration_odo_f64(
        int(ref_frame), float(fs), R, n, _ptr(gyro), _ptr(odo), layout, _ptr(ini),
        ini.shape[0], ini.shape[1], int(run_offset), int(bool(earth_rot)),
        _ptr(att), _ptr(pos), _ptr(vel), int(lanes_per_run), _stream()))
    return att, pos, vel


def imu_noise(fs, runs, ref_gyro, ref_accel, gyro_err, accel_err, seed, run_offset=0,
              vib_gyro=None, vib_accel=None, layout=LAYOUT_RUN_MAJOR, dump_z=False):
    """K1.  ref_gyro/ref_accel: CUDA f64 [n,3]; *_err: imu_model dicts.
    Returns gyro, accel ([R,n,3], [n,3,R] or, LAYOUT_CHANNEL_MAJOR, [R,3,n]) and, if dump_z,
    z [R,n,12]."""
    _require_cuda()
    lib = _lib.load()
    n = ref_gyro.shape[0]
    shape = {LAYOUT_RUN_MAJOR: (runs, n, 3), LAYOUT_TIME_MAJOR: (n, 3, runs),
             LAYOUT_CHANNEL_MAJOR: (runs, 3, n)}[layout]
    gyro = torch.empty(shape, dtype=torch.float64, device=ref_gyro.device)
    accel = torch.empty_like(gyro)
    z = torch.empty((runs, n, 12), dtype=torch.float64, device=ref_gyro.device) if dump_z else None
    ge, ae = _lib.sensor_err(gyro_err, 'arw'), _lib.sensor_err(accel_err, 'vrw')
    vg = vib_gyro if isinstance(vib_gyro, _lib.Vib) else _lib.vib(vib_gyro)
    va = vib_accel if isinstance(vib_accel, _lib.Vib) else _lib.vib(vib_accel)
    _lib.check(lib.b2ins_imu_noise_f64(
        float(fs), runs, n, _ptr(ref_gyro), _ptr(ref_accel), ctypes.byref(ge), ctypes.byref(ae),
        ctypes.byref(vg), ctypes.byref(va), int(seed), int(run_offset), layout,
        _ptr(gyro), _ptr(accel), _ptr(z), _stream()))
    return (gyro, accel, z) if dump_z else (gyro, accel)


def gps_noise(runs, ref_gps, gps_err, gps_type, seed, run_offset=0):
    """K6: pathgen.gps_gen for `runs` runs.  ref_gps: CUDA f64 [m,6]; gps_err {'stdp','stdv'} [3];
    gps_type 0 (LLA, ref_frame 0) or 1 (xyz).  Returns [R,m,6]."""
    _require_cuda()
    lib = _lib.load()
    m = ref_gps.shape[0]
    out = torch.empty((runs, m, 6), dtype=torch.float64, device=ref_gps.device)
    stdp = np.ascontiguousarray(np.broadcast_to(np.asarray(gps_err['stdp'], dtype=np.float64), (3,)))
    stdv = np.ascontiguousarray(np.broadcast_to(np.asarray(gps_err['stdv'], dtype=np.float64), (3,)))
    _lib.check(lib.b2ins_gps_noise_f64(runs, m, _ptr(ref_gps), _lib.host_ptr(stdp), _lib.host_ptr(stdv),
                                       int(gps_type), int(seed), int(run_offset), _ptr(out), _stream()))
    return out


class McResult:
    """Device-side results of one fused Monte-Carlo launch."""

    def __init__(self):
        self.end_err = None      # [R,9] att(wrapped),pos,vel error at the last sample
        self.end_state = None    # [R,9]
        self.proc_stats = None   # [R,3,9] max|e|, mean, std per run (stats_start >= 0)
        self.att = self.pos = self.vel = None   # [dump_runs,n,3]
        self.gyro = self.accel = None           # [dump_runs,n,3]
        self.odo = None                         # [dump_runs,n] (odometer variant)
        self.quat = None                        # [dump_runs,rows,4] att_quat of the kept samples
        self.lanes_per_run = 0


def make_mc_config(ref_frame, fs, n, runs, seed, gyro_err, accel_err, ini_sets, ini_rows,
                   earth_rot=True, run_offset=0, vib_gyro=None, vib_accel=None,
                   lanes_per_run=0, stats_start=-1, dump_runs=0, ini_offset=None,
                   odo_err=None, ref_odo=None, dump_stride=1):
    """odo_err {'scale','stdv'} + ref_odo (CUDA f64 [n]) select the odometer variant."""
    cfg = _lib.McConfig()
    cfg.ref_frame = int(ref_frame)
    cfg.earth_rot = int(bool(earth_rot))
    cfg.fs = float(fs)
    cfg.n = int(n)
    cfg.runs = int(runs)
    cfg.run_offset = int(run_offset)
    cfg.ini_offset = int(run_offset if ini_offset is None else ini_offset)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cfg.gyro_err = _lib.sensor_err(gyro_err, 'arw')
    cfg.accel_err = _lib.sensor_err(accel_err, 'vrw')
    cfg.vib_gyro = vib_gyro if isinstance(vib_gyro, _lib.Vib) else _lib.vib(vib_gyro)
    cfg.vib_accel = vib_accel if isinstance(vib_accel, _lib.Vib) else _lib.vib(vib_accel)
    cfg._keep_vib = (vib_gyro, vib_accel)   # assignment copies the structs: keep the series tensors alive
    cfg.ini_sets = int(ini_sets)
    cfg.ini_rows = int(ini_rows)
    cfg.lanes_per_run = int(lanes_per_run)
    cfg.stats_start = int(stats_start)
    cfg.dump_runs = int(dump_runs)
    cfg.dump_stride = int(dump_stride)
    cfg.algo = 0
    if odo_err is not None:
        assert ref_odo is not None and ref_odo.is_cuda and ref_odo.dtype == torch.float64
        cfg.algo = 1
        cfg.odo_scale = float(odo_err['scale'])
        cfg.odo_stdv = float(odo_err['stdv'])
        cfg.ref_odo = ref_odo.data_ptr()
        cfg._keep = ref_odo            # keep the tensor alive with the config
    return cfg


def mc_free_integration(cfg, ref_gyro, ref_accel, ref_nav, ini, want_state=False,
                        dump_nav=False, dump_imu=False, out=None, dump_quat=False):
    """K12: fused noise generation + free integration + per-run errors.
    ref_gyro, ref_accel [n,3]; ref_nav [n,9] (att,pos,vel); ini [S,rows]: CUDA f64.
    `out` may carry a preallocated McResult to reuse buffers."""
    _require_cuda()
    lib = _lib.load()
    dev = ref_gyro.device
    R, n, D = cfg.runs, cfg.n, cfg.dump_runs
    rows = -(-n // max(1, cfg.dump_stride))          # histories keep every dump_stride-th sample
    res = out or McResult()

    def buf(cur, shape):
        if cur is not None and tuple(cur.shape) == tuple(shape):
            return cur
        return torch.empty(shape, dtype=torch.float64, device=dev)

    res.end_err = buf(res.end_err, (R, 9))
    res.end_state = buf(res.end_state, (R, 9)) if want_state else None
    res.proc_stats = buf(res.proc_stats, (R, 3, 9)) if cfg.stats_start >= 0 else None
    if dump_nav and D > 0:
        res.att, res.pos, res.vel = (buf(res.att, (D, rows, 3)), buf(res.pos, (D, rows, 3)),
                                     buf(res.vel, (D, rows, 3)))
        res.quat = buf(res.quat, (D, rows, 4)) if dump_quat else None
    else:
        res.att = res.pos = res.vel = res.quat = None
    if dump_imu and D > 0:
        res.gyro, res.accel = buf(res.gyro, (D, rows, 3)), buf(res.accel, (D, rows, 3))
        res.odo = buf(res.odo, (D, rows)) if cfg.algo == 1 else None
    else:
        res.gyro = res.accel = res.odo = None
    cfg.dump_odo = res.odo.data_ptr() if res.odo is not None else None
    cfg.dump_quat = res.quat.data_ptr() if res.quat is not None else None
    _lib.check(lib.b2ins_mc_free_integration_f64(
        ctypes.byref(cfg), _ptr(ref_gyro), _ptr(ref_accel), _ptr(ref_nav), _ptr(ini),
        _ptr(res.end_err), _ptr(res.end_state), _ptr(res.proc_stats),
        _ptr(res.att), _ptr(res.pos), _ptr(res.vel), _ptr(res.gyro), _ptr(res.accel), _stream()))
    return res


class McPlan:
    """b2ins_mc_plan: persistent device + pinned buffers and a stream for one experiment
    shape; run() = stage, H2D, K12, K3, D2H, sync (the low-latency host path of Sim.run)."""

    def __init__(self, n, max_runs, ini_sets, ini_rows):
        _require_cuda()
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.n, self.max_runs, self.ini_sets, self.ini_rows = int(n), int(max_runs), int(ini_sets), int(ini_rows)
        self.device = torch.cuda.current_device()
        _lib.check(self._lib.b2ins_mc_plan_create(self.n, self.max_runs, self.ini_sets,
                                                  self.ini_rows, ctypes.byref(self._h)))

    def run(self, cfg, ref_gyro, ref_accel, ref_nav_end, ini, want_err=True, want_stats=True):
        """Host float64 C-contiguous arrays in (ref_nav_end: the 9 values att,pos,vel of the true
        trajectory at its last sample); (end_err [runs,9] or None, stats [3,9] or None) out."""
        hp = _lib.host_ptr
        stats = np.empty((3, 9)) if want_stats else None
        err = np.empty((cfg.runs, 9)) if want_err else None
        _lib.check(self._lib.b2ins_mc_plan_run(self._h, ctypes.byref(cfg), hp(ref_gyro), hp(ref_accel),
                                               hp(ref_nav_end), hp(ini), hp(err), hp(stats)))
        return err, stats

    def err_device_ptr(self):
        """Device address of the plan's end_err buffer (multi-GPU statistics exchange)."""
        return self._lib.b2ins_mc_plan_err_device(self._h)

    def stream_ptr(self):
        return self._lib.b2ins_mc_plan_stream(self._h)

    def close(self):
        if self._h:
            self._lib.b2ins_mc_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plan_cache = {}


def get_plan(n, runs, ini_sets, ini_rows):
    """A cached plan that fits (n, runs, ini layout) on the current device."""
    _require_cuda()
    key = (torch.cuda.current_device(), int(n), int(ini_sets), int(ini_rows))
    plan = _plan_cache.get(key)
    if plan is None or plan.max_runs < runs:
        if plan is not None:
            plan.close()
        plan = McPlan(n, runs, ini_sets, ini_rows)
        _plan_cache[key] = plan
    return plan


def f64c(a):
    """C-contiguous float64 numpy view/copy of a."""
    return np.ascontiguousarray(a, dtype=np.float64)


_ws_cache = {}


def _stats_ws(ncomp, dev):
    key = (ncomp, str(dev))
    if key not in _ws_cache:
        nbytes = _lib.load().b2ins_error_stats_workspace_bytes(ncomp)
        _ws_cache[key] = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=dev)
    return _ws_cache[key]


def error_stats(err):
    """K3 on one shard.  err: CUDA f64 [R, ncomp] -> stats [3, ncomp] = max|e|, mean, std."""
    _require_cuda()
    lib = _lib.load()
    R, nc = err.shape
    stats = torch.empty((3, nc), dtype=torch.float64, device=err.device)
    _lib.check(lib.b2ins_error_stats_f64(R, nc, _ptr(err), _ptr(stats),
                                         _ptr(_stats_ws(nc, err.device)), _stream()))
    return stats


def error_partial(err):
    """K3 phase 1: [2*ncomp] = (sum e, max|e|) of this shard."""
    lib = _lib.load()
    R, nc = err.shape
    out = torch.empty(2 * nc, dtype=torch.float64, device=err.device)
    _lib.check(lib.b2ins_error_partial_f64(R, nc, _ptr(err), _ptr(out),
                                           _ptr(_stats_ws(nc, err.device)), _stream()))
    return out


def error_partial2(err, mean):
    """K3 phase 2: [ncomp] = sum (e - mean)^2 of this shard."""
    lib = _lib.load()
    R, nc = err.shape
    out = torch.empty(nc, dtype=torch.float64, device=err.device)
    _lib.check(lib.b2ins_error_partial2_f64(R, nc, _ptr(err), _ptr(mean.contiguous()), _ptr(out),
                                            _ptr(_stats_ws(nc, err.device)), _stream()))
    return out


def psd_series(fs, n, runs, sensor, vib_def, seed, run_offset=0):
    """K5.  vib_def: {'type': 'psd', 'freq': (L0,), 'x','y','z': (L0,)} (Sim.__parse_env output).
    Returns the device series [runs, 3, N] and N (hand both to a Vib of type VIB_SERIES)."""
    _require_cuda()
    lib = _lib.load()
    freq = to_device(vib_def['freq'])
    if fs < 2.0 * float(vib_def['freq'][-1]) or fs < 0.0:
        raise ValueError('PSD table exceeds fs/2 (time_series_from_psd.py:33-34)')
    sxx = to_device(np.stack([vib_def['x'], vib_def['y'], vib_def['z']]))
    N = lib.b2ins_psd_series_len(int(n))
    out = []
    for r0 in range(0, runs, 16384):                       # grid.y limit per launch
        r1 = min(runs, r0 + 16384)
        series = torch.empty((r1 - r0, 3, N), dtype=torch.float64, device=freq.device)
        ws = torch.empty(lib.b2ins_psd_workspace_bytes(int(n), r1 - r0) // 8 + 1, dtype=torch.float64,
                         device=freq.device)
        _lib.check(lib.b2ins_psd_series_f64(float(fs), int(n), r1 - r0, int(sensor), freq.numel(),
                                            _ptr(freq), _ptr(sxx), int(seed), int(run_offset) + r0,
                                            _ptr(series), _ptr(ws), _stream()))
        out.append(series)
    return (out[0] if len(out) == 1 else torch.cat(out)), N


def vib_series(series, N):
    """A Vib of type VIB_SERIES over a device series [runs, 3, N]."""
    v = _lib.Vib()
    v.type = _lib.VIB_SERIES
    v.series = series.data_ptr()
    v.series_len = int(N)
    v._keep = series               # the device series lives as long as the Vib does
    return v


def allan_num_tau(n, fs):
    lib = _lib.load()
    m = (ctypes.c_int64 * 128)()
    k = lib.b2ins_allan_num_tau(int(n), float(fs), m, 128)
    return list(m[:k])


def allan_taus(n, fs):
    """tau [s] of the cluster sizes allan_var uses for n samples at fs (allan.py:37-43, :58)."""
    return np.asarray(allan_num_tau(n, fs), dtype=np.float64) / float(fs)


def allan(fs, x, n, nseries, inner=1, outer_stride=None, sample_stride=1):
    """K4.  x: CUDA f64 buffer holding `nseries` series of n samples; series s, sample t at
    x.flat[(s // inner) * outer_stride + (s % inner) + t * sample_stride].
    Returns avar [nseries, ntau], tau [ntau] (CUDA)."""
    _require_cuda()
    lib = _lib.load()
    if outer_stride is None:
        outer_stride = n * sample_stride if inner == 1 else n * inner
    ntau = len(allan_num_tau(n, fs))
    avar = torch.zeros((nseries, ntau), dtype=torch.float64, device=x.device)
    tau = torch.zeros((ntau,), dtype=torch.float64, device=x.device)
    if ntau == 0 or nseries == 0:
        return avar, tau
    ws = torch.empty(lib.b2ins_allan_workspace_bytes(n, nseries) // 8 + 1, dtype=torch.float64,
                     device=x.device)
    _lib.check(lib.b2ins_allan_f64(float(fs), int(n), int(nseries), _ptr(x), int(inner),
                                   int(outer_stride), int(sample_stride), _ptr(avar), _ptr(tau),
                                   _ptr(ws), _stream()))
    return avar, tau


def allan_mc(fs, runs, ref_gyro, ref_accel, gyro_err, accel_err, seed, run_offset=0):
    """K1 fused into K4: Allan variance of `runs` Monte-Carlo runs x 6 channels whose series are
    generated inside the tau-binning kernel (never written).  ref_gyro, ref_accel: CUDA f64 [n,3].
    Returns avar [runs, 6, ntau] (channels: accel x y z, gyro x y z) and tau [ntau] (CUDA)."""
    _require_cuda()
    lib = _lib.load()
    n = ref_gyro.shape[0]
    ntau = len(allan_num_tau(n, fs))
    avar = torch.zeros((runs, 6, ntau), dtype=torch.float64, device=ref_gyro.device)
    tau = torch.zeros((ntau,), dtype=torch.float64, device=ref_gyro.device)
    if ntau == 0 or runs == 0:
        return avar, tau
    ws = torch.empty(lib.b2ins_allan_workspace_bytes(n, runs * 6) // 8 + 1, dtype=torch.float64,
                     device=ref_gyro.device)
    ge, ae = _lib.sensor_err(gyro_err, 'arw'), _lib.sensor_err(accel_err, 'vrw')
    _lib.check(lib.b2ins_allan_mc_f64(float(fs), int(n), int(runs), _ptr(ref_gyro), _ptr(ref_accel),
                                      ctypes.byref(ge), ctypes.byref(ae), int(seed), int(run_offset),
                                      _ptr(avar), _ptr(tau), _ptr(ws), _stream()))
    return avar, tau


class EkfResult:
    """Device-side results of one loosely-coupled-filter launch (K7)."""

    def __init__(self):
        self.end_err = None      # [R,9] att (wrapped), pos (LLA), vel error at the last sample
        self.end_bias = None     # [R,6] gyro, accel bias estimates at the last sample
        self.consist = None      # [R,19] NEES sums (pos, vel, att), inside-3-sigma counts [15], epochs
        self.att = self.pos = self.vel = self.wb = self.ab = None   # [dump_runs,rows,3]


def ins_loose(fs, runs, seed, gyro_err, accel_err, gps_err, ini, ref_gyro, ref_accel, ref_nav, ref_gps,
              gps_idx, gps_vis, run_offset=0, ini_att_std=(0.02, 0.005, 0.005), earth_rot=True,
              stats_start=0, dump_runs=0, dump_stride=1, out=None, vel_rw=0.02, att_rw=0.0):
    """K7: Monte-Carlo loosely-coupled GNSS/INS filter (the spec: DESIGN.md section 11; csrc/ekf_kernel.cuh).
    ref_gyro, ref_accel [n,3], ref_nav [n,9], ref_gps [m,6], gps_vis [m]: CUDA f64; gps_idx [m]: CUDA
    int64 (IMU sample index of every GPS row).  ini: the 9 true initial values (LLA, body velocity, Euler
    angles).  Asynchronous on the current stream."""
    _require_cuda()
    lib = _lib.load()
    n, m = ref_gyro.shape[0], ref_gps.shape[0]
    dev = ref_gyro.device
    assert gps_idx.dtype == torch.int64 and gps_idx.is_cuda and gps_idx.is_contiguous()
    cfg = _lib.EkfConfig()
    cfg.fs, cfg.n, cfg.runs, cfg.run_offset, cfg.m = float(fs), int(n), int(runs), int(run_offset), int(m)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cfg.gyro_err = _lib.sensor_err(gyro_err, 'arw')
    cfg.accel_err = _lib.sensor_err(accel_err, 'vrw')
    stdp = np.broadcast_to(np.asarray(gps_err['stdp'], dtype=np.float64), (3,))
    stdv = np.broadcast_to(np.asarray(gps_err['stdv'], dtype=np.float64), (3,))
    ini = np.asarray(ini, dtype=np.float64).reshape(-1)
    for c in range(3):
        cfg.gps_stdp[c], cfg.gps_stdv[c] = float(stdp[c]), float(stdv[c])
        cfg.ini_att_std[c] = float(ini_att_std[c])
    for c in range(9):
        cfg.ini[c] = float(ini[c])
    cfg.stats_start, cfg.dump_runs, cfg.dump_stride = int(stats_start), int(dump_runs), int(dump_stride)
    cfg.earth_rot = int(bool(earth_rot))
    cfg.vel_rw, cfg.att_rw = float(vel_rw), float(att_rw)
    res = out or EkfResult()

    def buf(cur, shape):
        if cur is not None and tuple(cur.shape) == tuple(shape):
            return cur
        return torch.empty(shape, dtype=torch.float64, device=dev)
    res.end_err = buf(res.end_err, (runs, 9))
    res.end_bias = buf(res.end_bias, (runs, 6))
    res.consist = buf(res.consist, (runs, 19))
    if dump_runs > 0:
        rows = -(-n // max(1, int(dump_stride)))
        res.att, res.pos, res.vel, res.wb, res.ab = (buf(getattr(res, k), (dump_runs, rows, 3))
                                                     for k in ('att', 'pos', 'vel', 'wb', 'ab'))
    else:
        res.att = res.pos = res.vel = res.wb = res.ab = None
    _lib.check(lib.b2ins_ins_loose_f64(
        ctypes.byref(cfg), _ptr(ref_gyro), _ptr(ref_accel), _ptr(ref_nav), _ptr(ref_gps),
        ctypes.c_void_p(gps_idx.data_ptr()), _ptr(gps_vis), _ptr(res.end_err), _ptr(res.end_bias),
        _ptr(res.consist), _ptr(res.att), _ptr(res.pos), _ptr(res.vel), _ptr(res.wb), _ptr(res.ab), _stream()))
    return res
