"""BASELINE config 3, one GPU's share: motion_def-long_drive.csv @200 Hz (n = 193 036, host
path generator), 'low-accuracy' IMU, ref_frame 0, 12 500 runs (100 000 runs / 8 GPUs) through Sim.
Prints one JSON line with the path-generation time, the Monte-Carlo time and run-steps/s."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import imu_model, pathgen  # noqa: E402
from gnss_ins_sim_b200.sim import Sim  # noqa: E402
from gnss_ins_sim_b200.free_integration import FreeIntegration  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
    csv = os.path.join(ROOT, 'tests', 'golden', 'motion_def-long_drive.csv')
    ini, _ = pathgen.parse_motion(csv)
    imu = imu_model.IMU(accuracy='low-accuracy', axis=6, gps=False)
    sim = Sim([200.0, 0.0, 0.0], csv, ref_frame=0, imu=imu, algorithm=FreeIntegration(ini), seed=1)
    t0 = time.perf_counter()
    sim._load_trajectory()
    t_path = time.perf_counter() - t0
    sim.run(8)                               # warm-up (plan allocation, kernel load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run(runs)
    st = sim.get_error_stats('pos', -1, extra_opt='ned')
    t_mc = time.perf_counter() - t0
    n = sim._traj['ref_gyro'].shape[0]
    print(json.dumps({'config': 'BASELINE config 3, one of 8 shards', 'runs': runs, 'samples': n,
                      'path_gen_s': t_path, 'mc_s': t_mc, 'run_steps_per_s': runs * n / t_mc,
                      'pos_ned_std_m': st['std'].tolist()}))


if __name__ == '__main__':
    main()
