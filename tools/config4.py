"""BASELINE config 4 on one GPU: static 10 h @400 Hz (n = 14.4 M), 'low-accuracy' IMU, 256 runs,
Allan deviation of the 6 channels through Sim (K1 noise + K4 Allan, run blocks sized to memory).
Prints one JSON line with the wall time."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import imu_model  # noqa: E402
from gnss_ins_sim_b200.sim import Sim  # noqa: E402
from gnss_ins_sim_b200.allan_analysis import Allan  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n, fs = 14400000, 400.0
    ref_accel = np.tile(np.array([4.9, 0.0, -8.487]), (n, 1))
    traj = {'ref_pos': np.zeros((n, 3)), 'ref_vel': np.zeros((n, 3)), 'ref_att': np.zeros((n, 3)),
            'ref_accel': ref_accel, 'ref_gyro': np.zeros((n, 3))}
    imu = imu_model.IMU(accuracy='low-accuracy', axis=6, gps=False)
    sim = Sim([fs, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=Allan(), seed=1)
    sim.run(2)
    torch.cuda.synchronize()
    times = []
    for _ in range(2):      # the first pass also pays for the device allocations of the run blocks
        t0 = time.perf_counter()
        sim.run(runs)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = times[-1]
    ad = sim.get_data(['ad_gyro'])[0]['algo0_0']
    print(json.dumps({'config': 'BASELINE config 4', 'runs': runs, 'samples': n, 'channels': 6,
                      'seconds': dt, 'seconds_first_pass': times[0], 'sample_channels_per_s': runs * n * 6 / dt, 'ntau': int(ad.shape[0]),
                      'reference_estimate_core_hours': 3.7}))


if __name__ == '__main__':
    main()
