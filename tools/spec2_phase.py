"""Where do the cycles of mc_spec_kernel go?  Uses tools/libb2ins_prof.so (built with
-DB2INS_PHASE_CLOCKS): per warp and per step, the cycles an integrator warp spends waiting at the
hand-over barrier / stepping, and a producer warp waiting for tiles / producing / waiting at the
barrier.  GPU box only."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import build as b  # noqa: E402
b.LIB = os.path.join(ROOT, 'tools', 'libb2ins_prof.so')      # load the instrumented build
b.stale = lambda: False
from gnss_ins_sim_b200 import engine, _lib  # noqa: E402

MID_G = {'b': np.zeros(3), 'b_drift': np.full(3, 3.5 * np.pi / 180 / 3600),
         'b_corr': np.full(3, 100.0), 'arw': np.full(3, 0.25 * np.pi / 180 / 60)}
MID_A = {'b': np.zeros(3), 'b_drift': np.full(3, 5e-5), 'b_corr': np.full(3, 100.0),
         'vrw': np.full(3, 0.03 / 60)}


def main():
    _lib.load()
    diag = ctypes.CDLL(b.LIB).b2ins_diag_phase_clocks
    cases = [(1000, 4, '3,1,0'), (1000, 4, '3,1,1'), (1000, 4, '6,1,0'), (1000, 16, '1,4,1'), (4000, 1, '3,1,0'),
             (4000, 1, '6,1,0'), (12500, 1, '3,1,0'), (12500, 1, '6,1,0'), (12500, 2, '6,1,0'), (100000, 1, '6,1,0')]
    for rf in (1, 0):
        g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'traj_90deg_turn_100hz_rf%d.npz' % rf)))
        nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
        n = nav.shape[0]
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda()
               for a in (g['ref_gyro'], g['ref_accel'], nav, g['ini'][None])]
        for runs, lanes, shape, dbg in [c + (d,) for c in cases for d in (0, 1, 2)]:
            os.environ['B2INS_MC_SHAPE'] = shape
            os.environ['B2INS_MC_DEBUG'] = str(dbg)
            P = int(shape.split(',')[0])
            cfg = engine.make_mc_config(rf, 100.0, n, runs, 1, MID_G, MID_A, 1, 9, lanes_per_run=lanes)
            res = engine.mc_free_integration(cfg, *dev)
            torch.cuda.synchronize()
            out = (ctypes.c_ulonglong * 8)()
            diag(None, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.mc_free_integration(cfg, *dev, out=res)
            e1.record()
            torch.cuda.synchronize()
            diag(out, 0)
            iw = -(-runs * lanes // 32)
            pw = iw * P
            print(json.dumps({'rf': rf, 'runs': runs, 'lanes': lanes, 'shape_P_WI_split': shape,
                              'idle': {0: 'none', 1: 'producers', 2: 'integrators'}[dbg],
                              'ms': round(e0.elapsed_time(e1), 4),
                              'integrator_cycles_per_step': {'barrier_wait': round(out[4] / iw / n, 1),
                                                             'stepping': round(out[5] / iw / n, 1)},
                              'producer_cycles_per_step': {'tile_wait': round(out[0] / pw / n, 1),
                                                           'producing': round(out[7] / pw / n, 1),
                                                           'barrier_wait': round(out[6] / pw / n, 1)}}),
                  flush=True)
    os.environ.pop('B2INS_MC_SHAPE', None)
    os.environ.pop('B2INS_MC_DEBUG', None)


if __name__ == '__main__':
    main()
