"""Multi-GPU path (needs >= 2 GPUs, otherwise skipped): runs sharded over NCCL ranks give the
same ensemble statistics and per-run errors as one GPU, for any world size."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, assert_close

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _worker(rank, world, port, tmp):
    import torch.distributed as td
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    td.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank,
                          world_size=world, device_id=torch.device('cuda', rank))
    from gnss_ins_sim_b200 import imu_model, dist
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'philox_90deg_mid_rf1.npz')))
    traj = {k: g[k] for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    traj = dist.broadcast_trajectory(traj if rank == 0 else None)
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=FreeIntegration(g['ini']),
              seed=int(g['seed']))
    sim.run(1003)                                   # uneven shards
    st = {k: sim.get_error_stats('pos', -1)[k] for k in ('max', 'avg', 'std')}
    mine = sim.end_point_errors()
    allrows = dist.gather_rows(torch.from_numpy(mine), 1003)
    ps = sim.get_error_stats('vel', err_stats_start=5.0)
    # K3x: the fused statistics + peer-memory exchange kernel, twice (double-buffered windows)
    fused = np.zeros((2, 3, 9))
    try:
        p2p = dist.P2PStats(9)
        dev_err = torch.from_numpy(mine).cuda()
        fused[0] = p2p(dev_err, mine.shape[0]).cpu().numpy()
        fused[1] = p2p(dev_err * 2.0, mine.shape[0]).cpu().numpy()
        assert not p2p.timed_out()
        fused_ok = 1
    except Exception as e:          # symmetric memory unavailable on this box
        sys.stderr.write('P2PStats skipped: %s\n' % e)
        fused_ok = 0
    # the Allan experiment shards its runs too (5 runs on 2 ranks) and gathers the deviations
    from gnss_ins_sim_b200.allan_analysis import Allan
    sa = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=Allan(), seed=7)
    sa.run(5)
    ad = np.stack([sa.get_data(['ad_gyro'])[0]['algo0_%d' % r] for r in range(5)])
    np.savez(os.path.join(tmp, 'r%d.npz' % rank), rows=allrows, local=mine.shape[0], fused=fused,
             fused_ok=fused_ok, ad_gyro=ad,
             proc=np.stack([ps['std']['algo0_%d' % r] for r in (0, 501, 1002)]), **st)
    td.destroy_process_group()


def test_sharded_runs_match_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single-GPU reference of the same experiment
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf1.npz')
    traj = {k: g[k] for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=FreeIntegration(g['ini']),
              seed=int(g['seed']))
    sim.run(1003)
    one = sim.get_error_stats('pos', -1)
    rows = sim.end_point_errors()
    ps = sim.get_error_stats('vel', err_stats_start=5.0)
    from gnss_ins_sim_b200.allan_analysis import Allan
    sa = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu, algorithm=Allan(), seed=7)
    sa.run(5)
    ad_one = np.stack([sa.get_data(['ad_gyro'])[0]['algo0_%d' % r] for r in range(5)])
    locals_ = []
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), 'r%d.npz' % r))
        for k in ('max', 'avg', 'std'):
            assert_close(z[k], one[k], 1e-12, 1e-12, k)
        assert_close(z['rows'], rows, 1e-9, 1e-6, 'per-run errors')
        assert_close(z['proc'], np.stack([ps['std']['algo0_%d' % i] for i in (0, 501, 1002)]),
                     1e-9, 1e-9, 'process std')
        assert_close(z['ad_gyro'], ad_one, 1e-12, 0.0, 'sharded Allan deviation')
        locals_.append(int(z['local']))
        if int(z['fused_ok']):
            full = np.stack([np.abs(rows).max(0), rows.mean(0), rows.std(0)])
            assert_close(z['fused'][0], full, 1e-11, 1e-12, 'K3x fused exchange')
            assert_close(z['fused'][1], np.stack([full[0] * 2, full[1] * 2, full[2] * 2]), 1e-11, 1e-12,
                         'K3x second call')
    assert sorted(locals_) == [501, 502]
    # the first 8 runs are the golden ones
    assert_close(rows[:8, 3:6], g['pos'][:, -1] - g['ref_pos'][-1], 1e-6, 1e-2, 'golden')
