"""BASELINE config 5: the loosely-coupled GNSS/INS filter (K7).

PARITY WITH THE REFERENCE IS UNPINNED: demo_algorithms/ins_loose.py is a stub (prediction / correction
are `pass`).  What is tested instead: (1) the first-principles spec (oracle/ekf_np.py) is a CONSISTENT
filter -- NEES of the position / velocity / attitude blocks near 3, errors inside 3 sigma, errors far
below free integration; (2) the CUDA kernel equals the spec on identical Philox draws; (3) the kernel is
consistent at scale on the config-5 trajectory (motion_def-ins.csv, 73 250 samples, thousands of runs)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, assert_close, wrap_pi
import ekf_np
import oracle_np as onp

DEMO_IMU = {'gyro_b': np.zeros(3), 'gyro_arw': np.array([0.25, 0.25, 0.25]),
            'gyro_b_stability': np.array([3.5, 3.5, 3.5]), 'gyro_b_corr': np.array([100.0, 100.0, 100.0]),
            'accel_b': np.zeros(3), 'accel_vrw': np.array([0.03119, 0.03009, 0.04779]),
            'accel_b_stability': np.array([4.29e-5, 5.72e-5, 8.02e-5]),
            'accel_b_corr': np.array([200.0, 200.0, 200.0])}       # demo_ins_loose.py:28-37


def _imu():
    from gnss_ins_sim_b200 import imu_model
    return imu_model.IMU(accuracy=DEMO_IMU, axis=6, gps=True)


def _turn_case():
    """The 90-degree-turn trajectory in ref_frame 0 with its 10 Hz GPS truth (reference fixtures)."""
    t = load_golden('traj_90deg_turn_100hz_rf0.npz')
    g = dict(load_golden('gps_90deg_rf0.npz'))
    g['gps_visibility'] = np.ones_like(g['gps_visibility'])     # the motion definition says 0 throughout
    nav = np.concatenate([t['ref_att'], t['ref_pos'], t['ref_vel']], axis=1)
    idx = np.rint(g['gps_time'] * 100.0).astype(np.int64)
    return t, g, nav, idx


def test_spec_is_a_consistent_filter():
    t, g, nav, idx = _turn_case()
    imu = _imu()
    out = ekf_np.ins_loose(100.0, t['ref_gyro'], t['ref_accel'], nav, g['ref_gps'], idx, g['gps_visibility'],
                           imu.gyro_err, imu.accel_err, imu.gps_err, 11, np.arange(64), t['ini'], stats_start=100)
    assert out['epochs'] == 90
    nees = out['nees'].mean(0)
    assert np.all(nees > 1.5) and np.all(nees < 5.0), nees          # expected 3 per block
    assert out['inside3'].mean(0).min() > 0.97
    # the position error stays at the GPS level / sqrt(updates) while free integration of the same
    # measurements would keep the initial 5 m error: compare with the filter's own sigma
    e = out['end_err']
    sig = np.sqrt(out['P_diag_end'].mean(0))
    north_m = e[:, 3] * 6.37e6
    assert north_m.std() < 2.0 * sig[0] and sig[0] < 2.5


@pytest.fixture(scope='module')
def gpu():
    torch = pytest.importorskip('torch')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch


def _launch(torch, t, g, nav, idx, imu, seed, runs, run_offset=0, **kw):
    from gnss_ins_sim_b200 import engine
    dev = [engine.to_device(a) for a in (t['ref_gyro'], t['ref_accel'], nav, g['ref_gps'])]
    return engine.ins_loose(100.0, runs, seed, imu.gyro_err, imu.accel_err, imu.gps_err, t['ini'], dev[0], dev[1],
                            dev[2], dev[3], torch.from_numpy(idx).cuda(),
                            engine.to_device(np.asarray(g['gps_visibility'], dtype=np.float64)),
                            run_offset=run_offset, vel_rw=kw.pop('vel_rw', 0.0), **kw)


@pytest.mark.gpu
def test_kernel_equals_the_spec(gpu):
    """Same Philox draws (IMU noise, GPS noise, initial errors): histories, bias estimates, end-point
    errors and the consistency record of the kernel against the NumPy spec."""
    t, g, nav, idx = _turn_case()
    imu = _imu()
    R, r0, seed = 12, 5, 2025
    o = ekf_np.ins_loose(100.0, t['ref_gyro'], t['ref_accel'], nav, g['ref_gps'], idx, g['gps_visibility'],
                         imu.gyro_err, imu.accel_err, imu.gps_err, seed, np.arange(r0, r0 + R), t['ini'],
                         stats_start=100, want_hist=True)
    res = _launch(gpu, t, g, nav, idx, imu, seed, R, run_offset=r0, stats_start=100, dump_runs=R)
    att, pos, vel = res.att.cpu().numpy(), res.pos.cpu().numpy(), res.vel.cpu().numpy()
    assert np.abs(wrap_pi(att - o['att'])).max() < 1e-9
    assert_close(pos[:, :, :2], o['pos'][:, :, :2], 1e-9, 1e-4, 'lat/lon')        # 1e-13 rad ~ 1e-6 m
    assert_close(pos[:, :, 2], o['pos'][:, :, 2], 1e-9, 1e-2, 'alt')
    assert_close(vel, o['vel'], 1e-9, 1e-2, 'vel')
    assert_close(res.wb.cpu().numpy(), o['wb'], 1e-7, 1e-6, 'gyro bias estimate')
    assert_close(res.ab.cpu().numpy(), o['ab'], 1e-7, 1e-5, 'accel bias estimate')
    assert_close(res.end_err.cpu().numpy(), o['end_err'], 1e-7, 1e-6, 'end-point error')
    assert_close(res.end_bias.cpu().numpy(), o['end_bias'], 1e-7, 1e-6, 'end biases')
    con = res.consist.cpu().numpy()
    assert np.all(con[:, 18] == o['epochs'])
    assert_close(con[:, 0:3] / con[:, 18:19], o['nees'], 1e-6, 1e-3, 'NEES')
    assert np.abs(con[:, 3:18] / con[:, 18:19] - o['inside3']).max() < 1.5 / o['epochs']
    # decimated histories are rows of the full ones; invisible GPS = pure free integration + decay
    dec = _launch(gpu, t, g, nav, idx, imu, seed, R, run_offset=r0, dump_runs=4, dump_stride=7)
    assert np.array_equal(dec.pos.cpu().numpy(), pos[:4, ::7])
    g2 = dict(g)
    g2['gps_visibility'] = np.zeros_like(g['gps_visibility'])
    blind = _launch(gpu, t, g2, nav, idx, imu, seed, R, run_offset=r0)
    assert np.abs(blind.end_err.cpu().numpy()[:, 3:5]).max() > np.abs(res.end_err.cpu().numpy()[:, 3:5]).max()
    assert np.all(blind.end_bias.cpu().numpy() == 0.0)


@pytest.mark.gpu
def test_config5_filter_is_consistent_at_scale(gpu):
    """motion_def-ins.csv @100 Hz (n = 73 250, 7 325 GPS samples), demo_ins_loose.py's IMU, 2048 runs
    through Sim: NEES near 3, >= 98.5 % inside 3 sigma for every state, metre-level end-point errors."""
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.ins_loose import InsLoose
    sim = Sim([100.0, 10.0, 0.0], os.path.join(GOLDEN, 'motion_def-ins.csv'), ref_frame=0, imu=_imu(),
              algorithm=InsLoose(), seed=5)
    sim.run(2048)
    assert sim.data['time'].shape[0] == 73250
    c = sim.ekf_consistency()
    nees = c['nees'].mean(0)
    assert np.all(nees > 1.3) and np.all(nees < 3.8), nees     # 3 per block; the velocity block is conservative
                                                                # (the model-mismatch random walk, InsLoose docstring)
    assert c['inside3'].mean(0).min() > 0.985, c['inside3'].mean(0)
    st = sim.get_error_stats('pos', -1, extra_opt='ned')
    assert np.all(st['std'] < 1.0) and np.all(st['max'] < 4.0), st          # metres, from 5 / 7 m GPS noise
    sv = sim.get_error_stats('vel', -1)
    assert np.all(sv['std'] < 0.05)
    sa = sim.get_error_stats('att_euler', -1, angle=True)
    assert np.all(sa['std'][1:] < 1e-3) and sa['std'][0] < 1e-2
    # the bias estimates track the simulated drift: error well below the drift's own sigma for the accel
    wb = sim.get_data(['wb'])[0]['algo0_3']
    assert wb.shape == (73250, 3) and np.isfinite(wb).all()


@pytest.mark.gpu
def test_ins_loose_needs_its_inputs(gpu):
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.ins_loose import InsLoose
    t = load_golden('traj_90deg_turn_100hz_rf1.npz')
    traj = {k: t[k] for k in ('ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    with pytest.raises(ValueError):
        Sim([100.0, 10.0, 0.0], traj, ref_frame=1, imu=_imu(), algorithm=InsLoose(t['ini'])).run(2)
    with pytest.raises(NotImplementedError):
        InsLoose().run([100.0, None, None, None, None, None])
