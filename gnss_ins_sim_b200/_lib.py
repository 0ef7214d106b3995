"""ctypes binding of csrc/libb2ins.so (include/b2ins.h).  No CPU fallback: if the
library is missing or a call fails, the product path raises."""
import ctypes
import os

import numpy as np

from . import build as _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

OK, ERR_ARG, ERR_CUDA, ERR_NODEV = 0, 1, 2, 3
LAYOUT_RUN_MAJOR, LAYOUT_TIME_MAJOR, LAYOUT_CHANNEL_MAJOR = 0, 1, 2
VIB_NONE, VIB_RANDOM, VIB_SINUSOIDAL, VIB_SERIES = 0, 1, 2, 3


class SensorErr(ctypes.Structure):
    _fields_ = [('b', ctypes.c_double * 3), ('b_drift', ctypes.c_double * 3),
                ('b_corr', ctypes.c_double * 3), ('rw', ctypes.c_double * 3)]


class Vib(ctypes.Structure):
    _fields_ = [('type', ctypes.c_int32), ('series_len', ctypes.c_int32),
                ('amp', ctypes.c_double * 3), ('freq', ctypes.c_double),
                ('series', ctypes.c_void_p)]


class McConfig(ctypes.Structure):
    _fields_ = [('ref_frame', ctypes.c_int32), ('earth_rot', ctypes.c_int32),
                ('fs', ctypes.c_double), ('n', ctypes.c_int64), ('runs', ctypes.c_int64),
                ('run_offset', ctypes.c_int64), ('ini_offset', ctypes.c_int64),
                ('seed', ctypes.c_uint64),
                ('gyro_err', SensorErr), ('accel_err', SensorErr),
                ('vib_gyro', Vib), ('vib_accel', Vib),
                ('ini_sets', ctypes.c_int32), ('ini_rows', ctypes.c_int32),
                ('lanes_per_run', ctypes.c_int32), ('stats_start', ctypes.c_int32),
                ('dump_runs', ctypes.c_int64),
                ('algo', ctypes.c_int32), ('dump_stride', ctypes.c_int32),
                ('odo_scale', ctypes.c_double), ('odo_stdv', ctypes.c_double),
                ('ref_odo', ctypes.c_void_p), ('dump_odo', ctypes.c_void_p),
                ('dump_quat', ctypes.c_void_p)]


class EkfConfig(ctypes.Structure):
    _fields_ = [('fs', ctypes.c_double), ('n', ctypes.c_int64), ('runs', ctypes.c_int64),
                ('run_offset', ctypes.c_int64), ('m', ctypes.c_int64), ('seed', ctypes.c_uint64),
                ('gyro_err', SensorErr), ('accel_err', SensorErr),
                ('gps_stdp', ctypes.c_double * 3), ('gps_stdv', ctypes.c_double * 3),
                ('ini', ctypes.c_double * 9), ('ini_att_std', ctypes.c_double * 3),
                ('stats_start', ctypes.c_int64), ('dump_runs', ctypes.c_int64),
                ('dump_stride', ctypes.c_int32), ('earth_rot', ctypes.c_int32),
                ('vel_rw', ctypes.c_double), ('att_rw', ctypes.c_double)]


class B2insError(RuntimeError):
    pass


_I, _L, _D, _P, _U64 = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_uint64
_SE, _VB, _MC = ctypes.POINTER(SensorErr), ctypes.POINTER(Vib), ctypes.POINTER(McConfig)

# name -> (restype, argtypes); every symbol include/b2ins.h declares
SIGNATURES = {
    'b2ins_version': (_I, []),
    'b2ins_last_error': (ctypes.c_char_p, []),
    'b2ins_device_count': (_I, []),
    'b2ins_allan_num_tau': (_I, [_L, _D, c_int64_p, _I]),
    'b2ins_free_integration_f64': (_I, [_I, _D, _L, _L, _P, _P, _I, _P, _I, _I, _L, _I, _P, _P, _P, _I, _P]),
    'b2ins_free_integration_odo_f64': (_I, [_I, _D, _L, _L, _P, _P, _I, _P, _I, _I, _L, _I, _P, _P, _P, _I, _P]),
    'b2ins_free_integration_f64_host': (_I, [_I, _D, _L, _L, _P, _P, _I, _P, _I, _I, _L, _I, _P, _P, _P, _I]),
    'b2ins_imu_noise_f64': (_I, [_D, _L, _L, _P, _P, _SE, _SE, _VB, _VB, _U64, _L, _I, _P, _P, _P, _P]),
    'b2ins_imu_noise_f64_host': (_I, [_D, _L, _L, _P, _P, _SE, _SE, _VB, _VB, _U64, _L, _I, _P, _P, _P]),
    'b2ins_gps_noise_f64': (_I, [_L, _L, _P, _P, _P, _I, _U64, _L, _P, _P]),
    'b2ins_mc_free_integration_f64': (_I, [_MC, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'b2ins_mc_free_integration_f64_host': (_I, [_MC, _P, _P, _P, _P, _P, _P]),
    'b2ins_mc_plan_create': (_I, [_L, _L, _I, _I, ctypes.POINTER(ctypes.c_void_p)]),
    'b2ins_mc_plan_run': (_I, [_P, _MC, _P, _P, _P, _P, _P, _P]),
    'b2ins_mc_plan_destroy': (_I, [_P]),
    'b2ins_mc_plan_err_device': (_P, [_P]),
    'b2ins_mc_plan_stream': (_P, [_P]),
    'b2ins_error_stats_workspace_bytes': (_L, [_I]),
    'b2ins_error_partial_f64': (_I, [_L, _I, _P, _P, _P, _P]),
    'b2ins_error_partial2_f64': (_I, [_L, _I, _P, _P, _P, _P, _P]),
    'b2ins_error_stats_f64': (_I, [_L, _I, _P, _P, _P, _P]),
    'b2ins_error_stats_exchange_f64': (_I, [_L, _I, _P, _I, _I, ctypes.POINTER(ctypes.c_uint64), _U64, _P, _P, _P]),
    'b2ins_allan_workspace_bytes': (_L, [_L, _L]),
    'b2ins_allan_f64': (_I, [_D, _L, _L, _P, _L, _L, _L, _P, _P, _P, _P]),
    'b2ins_allan_mc_f64': (_I, [_D, _L, _L, _P, _P, _SE, _SE, _U64, _L, _P, _P, _P, _P]),
    'b2ins_allan_f64_host': (_I, [_D, _L, _L, _P, _L, _L, _L, _P, _P]),
    'b2ins_psd_series_len': (_I, [_L]),
    'b2ins_psd_workspace_bytes': (_L, [_L, _L]),
    'b2ins_psd_series_f64': (_I, [_D, _L, _L, _I, _I, _P, _P, _U64, _L, _P, _P, _P]),
    'b2ins_path_rows': (_L, [_P, _L, _D]),
    'b2ins_path_gen_host': (_L, [_P, _P, _L, _D, _D, _D, _D, _P, _I, _L, _P, _P, _P, c_int64_p, _P]),
    'b2ins_ins_loose_f64': (_I, [ctypes.POINTER(EkfConfig)] + [_P] * 15),
    'b2ins_diag_dfma_rate': (_I, [c_double_p]),
    'b2ins_diag_auto_lanes': (_I, [_L, _I, _I]),
    'b2ins_diag_mc_shape': (_I, [_I, _I, ctypes.POINTER(ctypes.c_int)]),
}

_lib = None
_load_error = None       # a failed build is not retried in the same process (every caller gets the same error)


def lib_path():
    return _build.LIB


def load():
    """Load libb2ins.so (building it first if nvcc is here and it is stale)."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if _load_error is not None:
        raise B2insError(_load_error)
    path = lib_path()
    if _build.stale():          # missing, or built from other sources than the ones in this tree
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box and no usable prebuilt library
            if not os.path.exists(path):
                _load_error = 'libb2ins.so is missing and could not be built: %s' % e
            else:
                _load_error = 'libb2ins.so was built from other sources and could not be rebuilt: %s' % e
            raise B2insError(_load_error)
    elif _build.LAST_BUILD == 'not checked':
        _build.LAST_BUILD = 'reused'
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def mc_shape(lanes, ref_frame=1):
    """'P,WI,split' of the fused Monte-Carlo launch for a lane-group width and a reference frame
    ('0' = single-warp form)."""
    out = (ctypes.c_int * 3)()
    check(load().b2ins_diag_mc_shape(int(lanes), int(ref_frame), out))
    return '%d,%d,%d' % (out[0], out[1], out[2]) if out[0] else '0'


def check(rc):
    if rc != OK:
        msg = load().b2ins_last_error().decode('utf-8', 'replace')
        if rc == ERR_ARG:
            raise ValueError('b2ins: ' + msg)
        raise B2insError('b2ins error %d: %s' % (rc, msg))


def host_ptr(a):
    """void* of a C-contiguous float64 numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(ctypes.c_void_p)


def sensor_err(err, white_key):
    """imu_model-style dict {'b','b_drift','b_corr',white_key} -> SensorErr."""
    a = np.array([err['b'], err['b_drift'], err['b_corr'], err[white_key]], dtype=np.float64)   # [4][3] = the struct
    assert a.shape == (4, 3)
    return SensorErr.from_buffer_copy(a)


def vib(vib_def, series_ptr=None, series_len=0):
    """Sim.__parse_env-style dict (or None) -> Vib."""
    v = Vib()
    v.type = VIB_NONE
    if vib_def is None:
        return v
    kind = vib_def['type'].lower()
    if kind == 'random':
        v.type = VIB_RANDOM
    elif kind == 'sinusoidal':
        v.type = VIB_SINUSOIDAL
        v.freq = float(vib_def['freq'])
    elif kind == 'psd':
        v.type = VIB_SERIES
        v.series = series_ptr
        v.series_len = int(series_len)
        return v
    else:
        raise ValueError('unknown vibration type %r' % vib_def['type'])
    v.amp[0], v.amp[1], v.amp[2] = float(vib_def['x']), float(vib_def['y']), float(vib_def['z'])
    return v
