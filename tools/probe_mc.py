"""Throughput probe of the fused Monte-Carlo kernel: run-steps/s for a sweep of
(runs, lanes_per_run, ref_frame).  GPU box only; prints one JSON line per point."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnss_ins_sim_b200 import engine  # noqa: E402


def one(runs, lanes, rf, reps):
    """A single configuration, for runs under ncu."""
    g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                                  'traj_90deg_turn_100hz_rf%d.npz' % rf)))
    mid_g = {'b': np.zeros(3), 'b_drift': np.full(3, 3.5 * np.pi / 180 / 3600),
             'b_corr': np.full(3, 100.0), 'arw': np.full(3, 0.25 * np.pi / 180 / 60)}
    mid_a = {'b': np.zeros(3), 'b_drift': np.full(3, 5e-5), 'b_corr': np.full(3, 100.0),
             'vrw': np.full(3, 0.03 / 60)}
    nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
    n = nav.shape[0]
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda()
           for a in (g['ref_gyro'], g['ref_accel'], nav, g['ini'][None])]
    cfg = engine.make_mc_config(rf, 100.0, n, runs, 1, mid_g, mid_a, 1, 9, lanes_per_run=lanes)
    res = engine.mc_free_integration(cfg, *dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.mc_free_integration(cfg, *dev, out=res)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({'rf': rf, 'runs': runs, 'n': n, 'lanes': lanes, 'ms': round(ms, 4),
                      'run_steps_per_s': runs * n / (ms * 1e-3)}), flush=True)


def main():
    if len(sys.argv) > 1:
        runs, lanes, rf = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
        one(runs, lanes, rf, int(sys.argv[4]) if len(sys.argv) > 4 else 3)
        return
    g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                                  'traj_90deg_turn_100hz_rf1.npz')))
    g0 = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                                   'traj_90deg_turn_100hz_rf0.npz')))
    mid_g = {'b': np.zeros(3), 'b_drift': np.full(3, 3.5 * np.pi / 180 / 3600),
             'b_corr': np.full(3, 100.0), 'arw': np.full(3, 0.25 * np.pi / 180 / 60)}
    mid_a = {'b': np.zeros(3), 'b_drift': np.full(3, 5e-5), 'b_corr': np.full(3, 100.0),
             'vrw': np.full(3, 0.03 / 60)}
    sweeps = [(1000, [1, 4, 8, 16, 32]), (10000, [1, 2, 4, 8, 32]), (100000, [1, 2, 4]),
              (1000000, [1])]
    for rf, gg in ((1, g), (0, g0)):
        nav = np.concatenate([gg['ref_att'], gg['ref_pos'], gg['ref_vel']], axis=1)
        n = nav.shape[0]
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda()
               for a in (gg['ref_gyro'], gg['ref_accel'], nav, gg['ini'][None])]
        for R, lanes_list in sweeps:
            for lanes in lanes_list:
                cfg = engine.make_mc_config(rf, 100.0, n, R, 1, mid_g, mid_a, 1, 9,
                                            lanes_per_run=lanes)
                res = engine.mc_free_integration(cfg, *dev)
                torch.cuda.synchronize()
                reps = 5 if R <= 100000 else 2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    engine.mc_free_integration(cfg, *dev, out=res)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(json.dumps({'rf': rf, 'runs': R, 'n': n, 'lanes': lanes, 'ms': round(ms, 4),
                                  'run_steps_per_s': R * n / (ms * 1e-3)}), flush=True)


if __name__ == '__main__':
    main()
