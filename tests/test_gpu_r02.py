"""GPU tests of what round 2 added: the warp-specialised launch shapes, parity on the reference's
white-bias-drift, PSD-vibration and 'ned' goldens, bulk / decimated histories with kernel-written
quaternions, and the fixes of the stream / cache / run_base findings."""
import os

import numpy as np
import pytest

from conftest import load_golden, assert_close, wrap_pi

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return True


def _traj(g):
    return {k: g[k] for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}


def _mid():
    from gnss_ins_sim_b200 import imu_model
    return imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)


SHAPES = [(1, '3,1,0'), (1, '6,1,0'), (1, '0'), (2, '3,1,0'), (2, '6,1,0'), (4, '3,1,0'), (4, '3,1,1'),
          (4, '6,1,0'), (4, '6,1,1'), (4, '6,2,0'), (8, '6,1,0'), (8, '6,2,0'), (8, '1,2,0'), (16, '1,4,0'), (16, '1,4,1'), (32, '1,4,1'), (32, '0')]


@pytest.mark.parametrize('rf', [1, 0])
def test_every_launch_shape_matches_the_reference(gpu, rf, monkeypatch):
    """All instantiated (G, P, WI, split) shapes of mc_spec_kernel and the single-warp form, on the
    Philox-injected reference golden: histories at 1e-9, equal end-point errors across shapes."""
    from gnss_ins_sim_b200 import engine
    g = load_golden('philox_90deg_mid_rf%d.npz' % rf)
    imu = _mid()
    nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
    dev = [engine.to_device(a) for a in (g['ref_gyro'], g['ref_accel'], nav, g['ini'][None])]
    R, n = 8, nav.shape[0]
    first = None
    for lanes, shape in SHAPES:
        monkeypatch.setenv('B2INS_MC_SHAPE', shape)
        cfg = engine.make_mc_config(rf, 100.0, n, R, int(g['seed']), imu.gyro_err, imu.accel_err, 1, 9,
                                    lanes_per_run=lanes, dump_runs=R)
        try:
            res = engine.mc_free_integration(cfg, *dev, dump_nav=True, dump_imu=True)
        except ValueError as e:       # a shape that is not instantiated says so
            assert 'no specialised kernel' in str(e), (lanes, shape, e)
            continue
        what = 'lanes %d shape %s' % (lanes, shape)
        assert_close(res.gyro.cpu().numpy(), g['gyro'], 1e-12, 1.0, what + ' gyro')
        assert_close(res.accel.cpu().numpy(), g['accel'], 1e-12, 1.0, what + ' accel')
        assert np.abs(wrap_pi(res.att.cpu().numpy() - g['att'])).max() < 1e-9, what
        assert_close(res.vel.cpu().numpy(), g['vel'], 1e-9, 1.0, what + ' vel')
        pos = res.pos.cpu().numpy()
        if rf == 1:
            assert_close(pos - pos[:, :1], g['pos'] - g['pos'][:, :1], 1e-9, 1.0, what + ' pos')
        else:
            assert_close(pos, g['pos'], 1e-9, 1e-3, what + ' pos')
        err = res.end_err.cpu().numpy()
        if first is None:
            first = err
        assert np.abs(err - first).max() < 1e-10, what
    monkeypatch.delenv('B2INS_MC_SHAPE')


@pytest.mark.parametrize('rf', [1, 0])
def test_ragged_run_counts_and_lengths_in_the_specialised_form(gpu, rf):
    """Run counts that do not fill the last CTA / lane group and series lengths that end inside a tile,
    a round, a pass and a speculative block of four steps: same end-point errors as the single-warp form
    of the same experiment ('' = the default shape: the attitude / velocity split for groups of 4 and 8 in
    ref_frame 1, the fused warp-specialised form otherwise)."""
    from gnss_ins_sim_b200 import engine
    g = load_golden('philox_90deg_mid_rf%d.npz' % rf)
    imu = _mid()
    nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
    cases = [(1, 4, ''), (9, 4, ''), (33, 1, ''), (5, 8, ''), (3, 16, ''), (37, 2, ''),
             (9, 4, '6,1,0'), (5, 8, '6,1,0'), (9, 4, '6,1,1')]
    if rf == 1:
        cases += [(9, 4, '6,2,0'), (5, 8, '6,2,0'), (1, 4, '6,2,0')]
    for n in (1, 2, 4, 5, 7, 9, 129, 131, 777):
        dev = [engine.to_device(a) for a in (g['ref_gyro'][:n], g['ref_accel'][:n], nav[:n], g['ini'][None])]
        for R, lanes, spec in cases:
            out = {}
            for shape in (spec, '0'):
                if shape:
                    os.environ['B2INS_MC_SHAPE'] = shape
                try:
                    cfg = engine.make_mc_config(rf, 100.0, n, R, 99, imu.gyro_err, imu.accel_err, 1, 9,
                                                lanes_per_run=lanes, run_offset=1000)
                    out[shape] = engine.mc_free_integration(cfg, *dev, want_state=True)
                    out[shape] = (out[shape].end_err.cpu().numpy(), out[shape].end_state.cpu().numpy())
                finally:
                    os.environ.pop('B2INS_MC_SHAPE', None)
            assert np.abs(out[spec][0] - out['0'][0]).max() < 1e-11, (rf, n, R, lanes, spec)
            assert np.abs(out[spec][1] - out['0'][1]).max() < 1e-9 * 5e6, (rf, n, R, lanes, spec)


@pytest.mark.parametrize('rf', [1, 0])
def test_white_bias_drift_dict_imu_matches_reference(gpu, rf):
    """A dict-`accuracy` IMU without *_b_corr (white bias drift, pathgen.py:591-593) through Sim."""
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_whitedrift_rf%d.npz' % rf)
    acc = {k[4:]: g[k] for k in g if k.startswith('acc_')}
    imu = imu_model.IMU(accuracy=acc, axis=6, gps=False)
    assert_close(imu.gyro_err['b_drift'], g['gyro_b_drift'], 1e-15, 0.0, 'the same model as the reference')
    R, r0 = g['gyro'].shape[0], int(g['run_ids'][0])
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=rf, imu=imu, algorithm=FreeIntegration(g['ini']),
              seed=int(g['seed']), run_base=r0)
    sim.run(R)
    for dn, key in (('att_euler', 'att_euler'), ('pos', 'pos'), ('vel', 'vel')):
        st = sim.get_error_stats(dn, err_stats_start=-1, angle=(dn == 'att_euler'))
        for k in ('max', 'avg', 'std'):
            assert_close(st[k], g['stat_%s_%s' % (key, k)], 1e-6, 1e-3, '%s %s' % (dn, k))
    h = sim.histories(imu=True)
    assert_close(h['gyro'], g['gyro'], 1e-12, 1.0, 'gyro')
    assert_close(h['accel'], g['accel'], 1e-12, 1.0, 'accel')
    assert np.abs(wrap_pi(h['att_euler'] - g['att'])).max() < 1e-9
    assert_close(h['vel'], g['vel'], 1e-9, 1.0, 'vel')


def test_psd_vibration_through_sim_matches_reference_twice(gpu):
    """env = PSD tables through Sim.run (K5 on torch's stream, K12 on the plan's stream: ordered), two
    runs of the same experiment: both equal the reference Sim fed the same phases."""
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf1_psd.npz')
    R, r0 = g['gyro'].shape[0], int(g['run_ids'][0])
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=1, imu=_mid(),
              env={'acc': g['env_acc'].copy(), 'gyro': g['env_gyro'].copy()},
              algorithm=FreeIntegration(g['ini']), seed=int(g['seed']), run_base=r0)
    for attempt in range(2):
        sim.run(R)
        for dn in ('att_euler', 'pos', 'vel'):
            st = sim.get_error_stats(dn, err_stats_start=-1, angle=(dn == 'att_euler'))
            for k in ('max', 'avg', 'std'):
                assert_close(st[k], g['stat_%s_%s' % (dn, k)], 1e-6, 1e-3, '%s %s (run %d)' % (dn, k, attempt))
    h = sim.histories(imu=True)
    assert_close(h['gyro'], g['gyro'], 1e-10, 1.0, 'gyro with PSD vibration')
    assert_close(h['accel'], g['accel'], 1e-10, 1.0, 'accel with PSD vibration')
    assert_close(h['vel'], g['vel'], 1e-9, 1.0, 'vel')
    # the vibration is really there
    assert np.abs(g['accel'] - g['ref_accel'][None]).std() > 0.05


def test_psd_series_survive_many_history_blocks(gpu):
    """Six one-run history blocks with PSD vibration on both sensors (the cache holds more than eight
    series): every block still reads ITS series (no eviction of a series that is in use)."""
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf1_psd.npz')
    env = {'acc': g['env_acc'].copy(), 'gyro': g['env_gyro'].copy()}
    mk = lambda hb: Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=1, imu=_mid(), env=env,    # noqa: E731
                        algorithm=FreeIntegration(g['ini']), seed=5, history_block=hb)
    a, b = mk(1), mk(32)
    a.run(6)
    b.run(6)
    ga, gb = a.get_data(['gyro'])[0], b.get_data(['gyro'])[0]
    for r in range(6):
        assert np.array_equal(ga[r], gb[r]), r
    assert not np.array_equal(ga[4], ga[5])


def test_ned_and_ecef_position_error_options_match_reference(gpu):
    """get_error_stats('pos', extra_opt='ned'|'ecef') against the reference's own output
    (ins_data_manager.py:543-552) for the philox_90deg_mid_rf0 experiment."""
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g, s = load_golden('philox_90deg_mid_rf0.npz'), load_golden('ned_stats_90deg_mid_rf0.npz')
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=0, imu=_mid(), algorithm=FreeIntegration(g['ini']),
              seed=int(s['seed']))
    sim.run(len(s['run_ids']))
    for opt in ('ned', 'ecef'):
        st = sim.get_error_stats('pos', -1, extra_opt=opt)
        for k in ('max', 'avg', 'std'):
            assert_close(st[k], s['stat_pos_%s_%s' % (opt, k)], 1e-6, 1e-3, '%s %s' % (opt, k))
        assert st['units'] == "['m', 'm', 'm']"


@pytest.mark.parametrize('rf', [1, 0])
def test_bulk_and_decimated_histories_with_quaternions(gpu, rf):
    from gnss_ins_sim_b200.sim import Sim, euler2quat_zyx
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf%d.npz' % rf)
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=rf, imu=_mid(), algorithm=FreeIntegration(g['ini']),
              seed=int(g['seed']))
    sim.run(8)
    full = sim.histories(quat=True)
    assert full['att_euler'].shape == (8, 1000, 3) and full['att_quat'].shape == (8, 1000, 4)
    assert np.abs(wrap_pi(full['att_euler'] - g['att'])).max() < 1e-9
    assert_close(full['vel'], g['vel'], 1e-9, 1.0, 'vel')
    # att_quat: attitude.euler2quat of every att_euler sample (ins_sim.py:729-794), written by the kernel
    q = np.stack([euler2quat_zyx(g['att'][r]) for r in range(8)])
    assert np.abs(full['att_quat'] - q).max() < 1e-9
    assert np.abs(np.linalg.norm(full['att_quat'], axis=2) - 1.0).max() < 1e-14
    # get_data serves single runs from the bulk arrays
    assert np.array_equal(sim.get_data(['pos'])[0]['algo0_5'], full['pos'][5])
    keep = {k: v.copy() for k, v in full.items()}
    for stride in (7, 10, 999, 1000, 5000):
        dec = sim.histories(stride=stride, quat=True, imu=True)
        rows = -(-1000 // stride)
        assert dec['pos'].shape == (8, rows, 3) and dec['att_quat'].shape == (8, rows, 4)
        assert dec['time'].shape == (rows,) and dec['gyro'].shape == (8, rows, 3)
        for k in ('att_euler', 'pos', 'vel', 'att_quat'):
            assert np.array_equal(dec[k], keep[k][:, ::stride]), (k, stride)
        assert_close(dec['gyro'], g['gyro'][:, ::stride], 1e-12, 1.0, 'gyro')


def test_gps_streams_follow_run_base(gpu):
    """Run r of an experiment with run_base b draws GPS stream b + r, like its IMU data."""
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    g = load_golden('philox_90deg_mid_rf1.npz')
    gg = load_golden('gps_90deg_rf1.npz')
    traj = dict(_traj(g), ref_gps=gg['ref_gps'], gps_time=gg['gps_time'], gps_visibility=gg['gps_visibility'])
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=True,
                        gps_opt={'stdp': gg['stdp'], 'stdv': gg['stdv']})
    a = Sim([100.0, 10.0, 0.0], traj, ref_frame=1, imu=imu, seed=3)
    b = Sim([100.0, 10.0, 0.0], traj, ref_frame=1, imu=imu, seed=3, run_base=2)
    a.run(4)
    b.run(2)
    ga, gb = a.get_data(['gps'])[0], b.get_data(['gps'])[0]
    assert np.array_equal(ga[2], gb[0]) and np.array_equal(ga[3], gb[1])
    assert not np.array_equal(ga[0], gb[0])


@pytest.mark.parametrize('n', [5041, 20000, 50400, 123457])
def test_fused_allan_equals_materialised_series(gpu, n):
    """K1 fused into K4 (the series generated inside the tau-binning kernel, never written) against
    K1 -> K4 on the materialised series of the same runs: same Philox draws, same Allan variances
    (the Gauss-Markov scan is associated differently: last-bit differences only)."""
    from gnss_ins_sim_b200 import engine, imu_model
    rng = np.random.default_rng(n)
    ref_gyro = engine.to_device(0.01 * rng.standard_normal((n, 3)))
    ref_accel = engine.to_device(np.array([0.3, -0.2, -9.8]) + 0.01 * rng.standard_normal((n, 3)))
    for accuracy in ('low-accuracy', {'gyro_b': np.array([10.0, -5.0, 2.0]), 'gyro_arw': np.array([0.3, 0.2, 0.1]),
                                      'gyro_b_stability': np.array([8.0, 6.0, 4.0]),
                                      'accel_b': np.array([1e-3, 0.0, -2e-3]), 'accel_vrw': np.array([0.04, 0.03, 0.05]),
                                      'accel_b_stability': np.array([1e-4, 2e-4, 5e-5]),
                                      'gyro_b_corr': np.array([50.0, np.inf, 300.0])}):
        imu = imu_model.IMU(accuracy=accuracy, axis=6, gps=False)
        R, r0, seed = 7, 40, 99
        avar, tau = engine.allan_mc(200.0, R, ref_gyro, ref_accel, imu.gyro_err, imu.accel_err, seed, run_offset=r0)
        gyro, accel = engine.imu_noise(200.0, R, ref_gyro, ref_accel, imu.gyro_err, imu.accel_err, seed,
                                       run_offset=r0, layout=engine.LAYOUT_CHANNEL_MAJOR)
        av_a, tau_a = engine.allan(200.0, accel, n, R * 3)
        av_g, _ = engine.allan(200.0, gyro, n, R * 3)
        assert np.array_equal(tau.cpu().numpy(), tau_a.cpu().numpy())
        ref = np.concatenate([av_a.cpu().numpy().reshape(R, 3, -1), av_g.cpu().numpy().reshape(R, 3, -1)], axis=1)
        got = avar.cpu().numpy()
        assert got.shape == ref.shape and np.all(ref > 0)
        assert np.abs(got / ref - 1.0).max() < 1e-9, (n, np.abs(got / ref - 1.0).max())


def test_allan_experiment_takes_the_fused_path_and_agrees(gpu, monkeypatch):
    """Sim + Allan: the default (fused) experiment equals the materialising one (B2INS_ALLAN_FUSED=0)."""
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.allan_analysis import Allan
    n = 30000
    z = np.zeros((n, 3))
    traj = {'ref_pos': z, 'ref_vel': z, 'ref_att': z, 'ref_accel': np.tile([0.0, 0.0, -9.8], (n, 1)), 'ref_gyro': z}
    out = {}
    for fused in ('1', '0'):
        monkeypatch.setenv('B2INS_ALLAN_FUSED', fused)
        sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=1, imu=imu_model.IMU('low-accuracy', axis=6, gps=False),
                  algorithm=Allan(), seed=21)
        sim.run(5)
        out[fused] = (np.stack([sim.get_data(['ad_gyro'])[0]['algo0_%d' % r] for r in range(5)]),
                      np.stack([sim.get_data(['ad_accel'])[0]['algo0_%d' % r] for r in range(5)]),
                      sim.get_data(['algo_time'])[0]['algo0_0'])
    for a, b in zip(out['1'], out['0']):
        assert a.shape == b.shape and np.abs(a / b - 1.0).max() < 1e-9
