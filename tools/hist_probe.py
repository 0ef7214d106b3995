"""Where the time of Sim.histories() goes at config 2 (1000 runs x 1000 samples): the dump launch, the
device-to-host copies, the rest.  GPU box only."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import engine, imu_model  # noqa: E402
from gnss_ins_sim_b200.sim import Sim  # noqa: E402
from gnss_ins_sim_b200.free_integration import FreeIntegration  # noqa: E402


def main():
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'traj_90deg_turn_100hz_rf1.npz')))
    imu = imu_model.IMU('mid-accuracy', axis=6, gps=False)
    traj = {k: g[k] for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
    dev = [engine.to_device(a) for a in (g['ref_gyro'], g['ref_accel'], nav, g['ini'][None])]
    out = {}
    for stride in (1, 10):
        for lanes in (0, 1, 4, 32):
            cfg = engine.make_mc_config(1, 100.0, 1000, 1000, 1, imu.gyro_err, imu.accel_err, 1, 9, dump_runs=1000,
                                        lanes_per_run=lanes, dump_stride=stride)
            res = engine.mc_free_integration(cfg, *dev, dump_nav=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                engine.mc_free_integration(cfg, *dev, dump_nav=True, out=res)
            e1.record()
            torch.cuda.synchronize()
            out['dump_kernel_ms_stride%d_lanes%d' % (stride, lanes)] = e0.elapsed_time(e1) / 5
    pin = torch.empty(res.att.shape, dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pin.copy_(res.att, non_blocking=True)
    torch.cuda.synchronize()
    out['d2h_ms_per_array_%dMB' % (res.att.numel() * 8 // 1000000)] = (time.perf_counter() - t0) / 5 * 1e3
    sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=FreeIntegration(g['ini']), seed=1)
    sim.run(1000)
    sim.histories()
    t0 = time.perf_counter()
    for _ in range(5):
        sim.histories()
    out['sim_histories_ms'] = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps(out))


if __name__ == '__main__':
    main()
