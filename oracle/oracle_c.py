"""ORACLE (test infrastructure): ctypes binding of oracle/_build/liboracle.so, the plain-C
restatement in oracle.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'liboracle.so')


class SensorErr(ctypes.Structure):
    _fields_ = [('b', ctypes.c_double * 3), ('b_drift', ctypes.c_double * 3),
                ('b_corr', ctypes.c_double * 3), ('rw', ctypes.c_double * 3)]


def build(force=False):
    src = os.path.join(HERE, 'oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', HERE, '-s'] + (['-B'] if force else []), check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_mc_free_integration.restype = ctypes.c_int
        _lib.orc_allan_var.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _err(d, key):
    s = SensorErr()
    for c in range(3):
        s.b[c], s.b_drift[c] = float(d['b'][c]), float(d['b_drift'][c])
        s.b_corr[c], s.rw[c] = float(d['b_corr'][c]), float(d[key][c])
    return s


def free_integration(ref_frame, fs, gyro, accel, ini, earth_rot=True):
    """gyro, accel [R,n,3]; ini [R, 9|10] -> att, pos, vel [R,n,3]."""
    lib = load()
    gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    accel = np.ascontiguousarray(accel, dtype=np.float64)
    ini = np.ascontiguousarray(ini, dtype=np.float64)
    R, n, _ = gyro.shape
    att, pos, vel = np.empty_like(gyro), np.empty_like(gyro), np.empty_like(gyro)
    for r in range(R):
        lib.orc_free_integration_run(
            ctypes.c_int(ref_frame), ctypes.c_double(fs), ctypes.c_int64(n), _p(gyro[r]),
            _p(accel[r]), _p(ini[r]), ctypes.c_int(ini.shape[1]), ctypes.c_int(int(earth_rot)),
            _p(att[r]), _p(pos[r]), _p(vel[r]), None)
    return att, pos, vel


def imu_noise(fs, ref_gyro, ref_accel, gyro_err, accel_err, seed, run_ids):
    lib = load()
    rg = np.ascontiguousarray(ref_gyro, dtype=np.float64)
    ra = np.ascontiguousarray(ref_accel, dtype=np.float64)
    n = rg.shape[0]
    ge, ae = _err(gyro_err, 'arw'), _err(accel_err, 'vrw')
    gyro = np.empty((len(run_ids), n, 3))
    accel = np.empty_like(gyro)
    for i, r in enumerate(run_ids):
        lib.orc_imu_noise_run(ctypes.c_double(fs), ctypes.c_int64(n), _p(rg), _p(ra),
                              ctypes.byref(ge), ctypes.byref(ae), ctypes.c_uint64(int(seed)),
                              ctypes.c_uint64(int(r)), _p(gyro[i]), _p(accel[i]))
    return gyro, accel


def mc_free_integration(ref_frame, fs, runs, run0, ref_gyro, ref_accel, ref_nav_end, gyro_err,
                        accel_err, seed, ini_sets, earth_rot=True, ini_offset=None, threads=0):
    """Fused Monte-Carlo loop on the host (OpenMP over runs).  ini_sets [S, 9|10].
    Returns end_err [runs, 9] and the number of threads used."""
    lib = load()
    rg = np.ascontiguousarray(ref_gyro, dtype=np.float64)
    ra = np.ascontiguousarray(ref_accel, dtype=np.float64)
    end = np.ascontiguousarray(ref_nav_end, dtype=np.float64)
    ini = np.ascontiguousarray(ini_sets, dtype=np.float64)
    n = rg.shape[0]
    ge, ae = _err(gyro_err, 'arw'), _err(accel_err, 'vrw')
    out = np.empty((runs, 9))
    used = lib.orc_mc_free_integration(
        ctypes.c_int(ref_frame), ctypes.c_double(fs), ctypes.c_int64(n), ctypes.c_int64(runs),
        ctypes.c_int64(run0), ctypes.c_int64(run0 if ini_offset is None else ini_offset), _p(rg),
        _p(ra), _p(end), ctypes.byref(ge), ctypes.byref(ae), ctypes.c_uint64(int(seed)), _p(ini),
        ctypes.c_int(ini.shape[0]), ctypes.c_int(ini.shape[1]), ctypes.c_int(int(earth_rot)),
        _p(out), ctypes.c_int(threads))
    return out, used


def array_stats(err):
    lib = load()
    err = np.ascontiguousarray(err, dtype=np.float64)
    st = np.empty((3, err.shape[1]))
    lib.orc_array_stats(ctypes.c_int64(err.shape[0]), ctypes.c_int(err.shape[1]), _p(err), _p(st))
    return st


def allan_var(x, fs):
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    avar, tau = np.zeros(128), np.zeros(128)
    k = lib.orc_allan_var(_p(x), ctypes.c_int64(x.size), ctypes.c_int64(1), ctypes.c_double(fs),
                          _p(avar), _p(tau))
    return avar[:k].copy(), tau[:k].copy()
