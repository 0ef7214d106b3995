"""One launch of each kernel family at a representative size, for ncu captures
(profiles/ncu_*_r02.json).  GPU box only.
    ncu --set full --import-source on --clock-control none -k regex:<kernel> -c 1 -o out python tools/evidence_r02.py <mode>
modes: k1 (imu_noise_kernel), k3 (stats: staged path), k5 (psd), k6 (gps), k7 (ekf), allan_gen, allan_stream,
       mc_c3 (mc_spec_kernel at the config-3 shape: G = 1, ref_frame 0), mc_plain (mc_kernel, 10^6 runs)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnss_ins_sim_b200 import engine, imu_model  # noqa: E402


def main():
    mode = sys.argv[1]
    imu = imu_model.IMU('low-accuracy', axis=6, gps=True)
    g = {rf: dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'traj_90deg_turn_100hz_rf%d.npz' % rf))) for rf in (0, 1)}
    dev = {rf: [engine.to_device(a) for a in (g[rf]['ref_gyro'], g[rf]['ref_accel'],
                                              np.concatenate([g[rf]['ref_att'], g[rf]['ref_pos'], g[rf]['ref_vel']], 1),
                                              g[rf]['ini'][None])] for rf in (0, 1)}
    if mode == 'k1':          # 65536 runs x 1000 samples, [R][n][3]
        engine.imu_noise(100.0, 65536, dev[1][0], dev[1][1], imu.gyro_err, imu.accel_err, 1)
    elif mode == 'k3':        # 10^6 runs x 9 components: the staged statistics path
        engine.error_stats(torch.randn(1000000, 9, dtype=torch.float64, device='cuda'))
    elif mode == 'k5':        # 512 series of N = 16384
        tab = np.stack([np.linspace(0, 50, 200), np.ones(200), np.ones(200), np.ones(200)], 1)
        vib = {'type': 'psd', 'freq': tab[:, 0], 'x': tab[:, 1], 'y': tab[:, 2], 'z': tab[:, 3]}
        engine.psd_series(100.0, 20000, 171, 0, vib, 1)
    elif mode == 'k6':        # 100000 runs x 100 GPS samples
        gg = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'gps_90deg_rf1.npz')))
        engine.gps_noise(100000, engine.to_device(gg['ref_gps']), imu.gps_err, 1, 1)
    elif mode == 'k7':        # 4096 runs x 1000 samples with GPS at 10 Hz
        gg = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'gps_90deg_rf0.npz')))
        idx = torch.from_numpy(np.rint(gg['gps_time'] * 100.0).astype(np.int64)).cuda()
        engine.ins_loose(100.0, 4096, 1, imu.gyro_err, imu.accel_err, imu.gps_err, g[0]['ini'], dev[0][0], dev[0][1],
                         dev[0][2], engine.to_device(gg['ref_gps']), idx,
                         engine.to_device(np.ones(len(idx))))
    elif mode in ('allan_gen', 'allan_stream'):   # 64 runs x 6 channels x 2 M samples
        n = 2000000
        rg, ra = torch.zeros((n, 3), dtype=torch.float64, device='cuda'), torch.zeros((n, 3), dtype=torch.float64, device='cuda')
        if mode == 'allan_gen':
            engine.allan_mc(400.0, 64, rg, ra, imu.gyro_err, imu.accel_err, 1)
        else:
            x = torch.randn(96, n, dtype=torch.float64, device='cuda')
            engine.allan(400.0, x, n, 96)
    elif mode == 'mc_c3':     # 12500 runs x 1000 samples, ref_frame 0: the config-3 shard shape per 1000 steps
        cfg = engine.make_mc_config(0, 100.0, 1000, 12500, 1, imu.gyro_err, imu.accel_err, 1, 9)
        engine.mc_free_integration(cfg, *dev[0])
    elif mode == 'mc_plain':  # 10^6 runs: the single-warp throughput form
        cfg = engine.make_mc_config(1, 100.0, 1000, 1000000, 1, imu.gyro_err, imu.accel_err, 1, 9)
        engine.mc_free_integration(cfg, *dev[1])
    else:
        raise SystemExit('unknown mode ' + mode)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
