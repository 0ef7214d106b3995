"""cProfile of bench.py's e2e step (Sim construction + run + get_error_stats from pinned host arrays):
where the host time of the public API goes.  GPU box only.
    python tools/e2e_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import imu_model  # noqa: E402
from gnss_ins_sim_b200.sim import Sim  # noqa: E402
from gnss_ins_sim_b200.free_integration import FreeIntegration  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'traj_90deg_turn_100hz_rf1.npz')))
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(g[k])).pin_memory()
              for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    traj = {k: v.numpy() for k, v in pinned.items()}

    def step():
        algo = FreeIntegration(g['ini'])
        sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=algo, seed=12345)
        sim.run(1000)
        return sim.get_error_stats('pos', err_stats_start=-1)

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    print('e2e step %.1f us' % (dt * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(28)


if __name__ == '__main__':
    main()
