"""GPU tests of the host-side mirror of the reference interface: the FreeIntegration / Allan
plugins through the reference's plugin protocol and the Sim facade, against the reference's
golden vectors."""
import os

import numpy as np
import pytest

import oracle_np as onp
from conftest import GOLDEN, load_golden, assert_close, wrap_pi, write_logged_dir

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def gpu():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return True


def _traj(g):
    return {k: g[k] for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}


def test_plugin_protocol_like_insalgomgr(gpu):
    """reset() -> run(set_of_input) -> get_results() per run, ins_algo_manager.py:77-95."""
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    for rf in (1, 0):
        g = load_golden('seeded_90deg_rf%d.npz' % rf)
        algo = FreeIntegration(g['ini'])
        for r in range(2):
            algo.reset()
            algo.run([rf, 100.0, g['gyro'][r].copy(), g['accel'][r].copy()])
            att, pos, vel = algo.get_results()
            assert att.shape == (1000, 3)
            assert np.abs(wrap_pi(att - g['att'][r])).max() < 1e-9
            assert_close(pos, g['pos'][r], 1e-9, 1.0, 'pos')
            assert_close(vel, g['vel'][r], 1e-9, 1.0, 'vel')
        assert algo.run_times == 2
    # run k uses initial-state set k-1 while k <= sets, then set 0 (free_integration.py:85-87)
    g = load_golden('seeded_90deg_rf1.npz')
    inis = np.tile(g['ini'][:, None], (1, 2))
    inis[3, 1] += 0.5
    algo = FreeIntegration(inis)
    outs = []
    for r in range(3):
        algo.run([1, 100.0, g['gyro'][0], g['accel'][0]])
        outs.append(algo.get_results()[2][0].copy())
    assert not np.allclose(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize('rf', [1, 0])
def test_sim_facade_matches_reference_summary(gpu, rf, capsys):
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf%d.npz' % rf)
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=rf, imu=imu,
              algorithm=FreeIntegration(g['ini']), seed=int(g['seed']))
    sim.run(8)
    names = sim.results(err_stats_start=-1)
    out = capsys.readouterr().out
    assert 'Simulation runs: 8' in out and 'Reference frame: %d' % rf in out
    assert 'statistics for simulation position from algo' in out
    for n in ('att_euler', 'pos', 'vel', 'accel', 'gyro', 'ref_pos', 'time'):
        assert n in names
    for dn, key in (('att_euler', 'att_euler'), ('pos', 'pos'), ('vel', 'vel')):
        st = sim.get_error_stats(dn, err_stats_start=-1, angle=(dn == 'att_euler'))
        for k in ('max', 'avg', 'std'):
            assert_close(st[k], g['stat_%s_%s' % (key, k)], 1e-6, 1e-3, '%s %s' % (dn, k))
    # output units: deg for angles (and for lat/lon when ref_frame == 0)
    st = sim.get_error_stats('att_euler', -1, angle=True, use_output_units=True)
    assert_close(st['std'], g['stat_att_euler_std'] * 180 / np.pi, 1e-6, 1e-3, 'deg')
    assert st['units'] == "['deg', 'deg', 'deg']"
    # lazily materialised histories use the reference's keys
    pos = sim.get_data(['pos'])[0]
    assert list(pos.keys())[:2] == ['algo0_0', 'algo0_1'] and len(pos) == 8
    assert_close(pos['algo0_3'], g['pos'][3], 1e-9, 1.0, 'pos history')
    att = sim.get_data(['att_euler'])[0]['algo0_7']
    assert np.abs(wrap_pi(att - g['att'][7])).max() < 1e-8
    gyro = sim.get_data(['gyro'])[0]
    assert_close(gyro[5], g['gyro'][5], 1e-12, 1.0, 'gyro history')
    assert_close(sim.get_data(['accel'])[0][0], g['accel'][0], 1e-12, 1.0, 'accel history')
    # per-run process statistics from t >= 2.5 s (ins_data_manager.py:761-795)
    ps = sim.get_error_stats('vel', err_stats_start=2.5)
    o = onp.process_error_stats(g['vel'], g['ref_vel'], 250)
    for k in ('max', 'avg', 'std'):
        got = np.stack([ps[k]['algo0_%d' % r] for r in range(8)])
        assert_close(got, o[k], 1e-6, 1e-4, 'proc ' + k)
    if rf == 0:   # 'ned' option: LLA error -> metres in the local NED frame (:543-552), against the
        # reference's own get_error_stats(..., extra_opt='ned') for this experiment
        s = load_golden('ned_stats_90deg_mid_rf0.npz')
        st = sim.get_error_stats('pos', -1, extra_opt='ned')
        for k in ('max', 'avg', 'std'):
            assert_close(st[k], s['stat_pos_ned_' + k], 1e-6, 1e-3, 'ned ' + k)


def test_sim_vibration_env(gpu):
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf1_vibrand.npz')
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=1, imu=imu,
              env={'acc': '[0.03 0.001 0.01]-random', 'gyro': '[6 5 4]d-random'},
              algorithm=FreeIntegration(g['ini']), seed=int(g['seed']))
    sim.run(3)
    st = sim.get_error_stats('pos', -1)
    assert_close(st['avg'], g['stat_pos_avg'], 1e-6, 1e-3, 'avg')
    assert_close(sim.get_data(['gyro'])[0][2], g['gyro'][2], 1e-12, 1.0, 'gyro')


def test_sim_allan_and_foreign_plugins(gpu):
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.allan_analysis import Allan
    g = load_golden('philox_90deg_low_rf1_run1000.npz')
    imu = imu_model.IMU(accuracy='low-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=1, imu=imu, algorithm=Allan(),
              seed=int(g['seed']), run_base=1000)
    sim.run(4)
    tau, ada, adg = sim.get_data(['algo_time', 'ad_accel', 'ad_gyro'])
    assert sorted(ada.keys()) == ['algo0_0', 'algo0_1', 'algo0_2', 'algo0_3']
    for r in range(4):
        for c in range(3):
            av, t = onp.allan_var(g['accel'][r][:, c], 100.0)
            assert_close(ada['algo0_%d' % r][:, c], np.sqrt(av), 1e-9, 0.0, 'ad_accel')
            av, t = onp.allan_var(g['gyro'][r][:, c], 100.0)
            assert_close(adg['algo0_%d' % r][:, c], np.sqrt(av), 1e-9, 0.0, 'ad_gyro')
        assert_close(tau['algo0_%d' % r], t, 1e-15, 0.0, 'tau')

    class MeanGyro(object):          # a reference-style plugin this engine knows nothing about
        name = 'mean'

        def __init__(self):
            self.input = ['fs', 'gyro']
            self.output = ['wb']
            self.batch = True
            self.results = None

        def run(self, s):
            self.results = [s[1].mean(0) * s[0]]

        def get_results(self):
            return self.results

        def reset(self):
            self.results = None
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=1, imu=imu, algorithm=MeanGyro(),
              seed=int(g['seed']), run_base=1000)
    sim.run(2)
    wb = sim.get_data(['wb'])[0]
    assert sorted(wb.keys()) == ['mean_0', 'mean_1']
    assert_close(wb['mean_1'], g['gyro'][1].mean(0) * 100.0, 1e-12, 1e-6, 'foreign plugin')


@pytest.mark.parametrize('rf', [1, 0])
def test_odometer_variant_matches_reference(gpu, rf):
    """free_integration_odo (demo_free_integration.py's algo1): K2 with supplied odometer data,
    and the fused path with pathgen.odo_gen noise through Sim, against the reference."""
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration_odo import FreeIntegration as FreeIntegrationOdo
    g = load_golden('philox_90deg_mid_rf%d_odo.npz' % rf)
    R = g['odo'].shape[0]
    for lanes in (1, 8, 32):
        algo = FreeIntegrationOdo(g['ini'], lanes_per_run=lanes)
        assert algo.input == ['ref_frame', 'fs', 'gyro', 'odo']
        att, pos, vel = algo.run_batch(rf, 100.0, g['gyro'], g['odo'])
        assert np.abs(wrap_pi(att - g['att'])).max() < 1e-9
        assert_close(pos, g['pos'], 1e-9, 1.0 if rf == 1 else 1e-7, 'pos')
        assert_close(vel, g['vel'], 1e-9, 1.0, 'vel')
    algo = FreeIntegrationOdo(g['ini'])
    algo.run([rf, 100.0, g['gyro'][2], g['odo'][2]])          # plugin protocol, one run
    assert_close(algo.get_results()[2], g['vel'][2], 1e-9, 1.0, 'vel single')
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False, odo=True,
                        odo_opt={'scale': float(g['odo_scale']), 'stdv': float(g['odo_stdv'])})
    traj = dict(_traj(g), ref_odo=g['ref_odo'])
    sim = Sim([100.0, 0.0, 0.0], traj, ref_frame=rf, imu=imu, algorithm=FreeIntegrationOdo(g['ini']),
              seed=int(g['seed']))
    sim.run(R)
    for dn in ('att_euler', 'pos', 'vel'):
        st = sim.get_error_stats(dn, -1, angle=(dn == 'att_euler'))
        for k in ('max', 'avg', 'std'):
            assert_close(st[k], g['stat_%s_%s' % (dn, k)], 1e-6, 1e-3, '%s %s' % (dn, k))
    assert_close(sim.get_data(['odo'])[0][3], g['odo'][3], 1e-12, 1.0, 'odo history')
    assert_close(sim.get_data(['pos'])[0]['algo0_4'], g['pos'][4], 1e-9, 1.0 if rf == 1 else 1e-7, 'pos hist')
    assert_close(sim.get_data(['accel'])[0][1], g['accel'][1], 1e-12, 1.0, 'accel history')


@pytest.mark.parametrize('rf', [0, 1])
def test_gps_measurements_match_reference(gpu, rf):
    """K6 (pathgen.gps_gen for all runs) against the reference fed the same normals, directly and
    as Sim's per-run 'gps' data with the trajectory generated from the motion definition."""
    from gnss_ins_sim_b200 import engine, imu_model
    from gnss_ins_sim_b200.sim import Sim
    g = load_golden('gps_90deg_rf%d.npz' % rf)
    R, m, _ = g['gps'].shape
    err = {'stdp': g['stdp'], 'stdv': g['stdv']}
    out = engine.gps_noise(R, engine.to_device(g['ref_gps']), err, rf, int(g['seed']),
                           run_offset=int(g['run_ids'][0])).cpu().numpy()
    scale = np.array([1e-6, 1e-6, 1.0, 1.0, 1.0, 1.0]) if rf == 0 else 1.0
    assert_close(out, g['gps'], 1e-12, scale, 'gps')
    assert engine.gps_noise(0, engine.to_device(g['ref_gps']), err, rf, 1).shape == (0, m, 6)
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=True, gps_opt=err)
    sim = Sim([100.0, 10.0, 0.0], os.path.join(GOLDEN, 'motion_def-90deg_turn.csv'), ref_frame=rf, imu=imu,
              algorithm=None, seed=int(g['seed']))
    sim.run(int(g['run_ids'][-1]) + 1)
    assert_close(sim.get_data(['ref_gps'])[0], g['ref_gps'], 1e-12, scale, 'ref_gps')
    assert_close(sim.get_data(['gps_time'])[0], g['gps_time'], 1e-12, 1.0, 'gps_time')
    r = int(g['run_ids'][2])
    assert_close(sim.get_data(['gps'])[0][r], g['gps'][2], 1e-12, scale, 'gps history')


@pytest.mark.parametrize('name', ['bosch', 'nxp'])
def test_sim_on_a_logged_data_directory(gpu, tmp_path, name):
    """demo_free_integration_openimu.py: Sim on a directory of logged .csv files (no IMU model),
    FreeIntegration with the gravity override and earth_rot=False, against the reference's result
    on the same data; error statistics against the (all-zero) reference files."""
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    from gnss_ins_sim_b200.allan_analysis import Allan
    g = load_golden('logged_%s.npz' % name)
    d = write_logged_dir(str(tmp_path / name), g)
    algo = FreeIntegration(g['ini'], earth_rot=False)
    sim = Sim([100.0, 0.0, 0.0], d, ref_frame=0, imu=None, algorithm=[algo, Allan()])
    sim.run(1)
    att, pos, vel = (sim.get_data([k])[0]['algo0_0'] for k in ('att_euler', 'pos', 'vel'))
    assert np.abs(wrap_pi(att - g['att'])).max() < 1e-9
    assert_close(pos, g['pos'], 1e-9, 1e-7, 'pos')
    assert_close(vel, g['vel'], 1e-9, 1.0, 'vel')
    st = sim.get_error_stats('vel', -1)
    assert_close(st['max'], np.abs(g['vel'][-1]), 1e-9, 1e-9, 'end-point |error| vs a zero reference')
    assert_close(st['std'], np.zeros(3), 0.0, 1e-12, 'one run: no spread')
    ps = sim.get_error_stats('att_euler', err_stats_start=2.0)
    assert_close(ps['max']['algo0_0'], np.abs(wrap_pi(g['att'][200:])).max(0), 1e-9, 1e-9, 'process max')
    ned = sim.get_error_stats('pos', -1, extra_opt='ned')
    assert ned['units'] == "['m', 'm', 'm']" and np.isfinite(ned['max']).all()
    ad = sim.get_data(['ad_gyro'])[0]['algo1_0']
    o, _ = onp.allan_var(g['gyro'][:, 2], 100.0)
    assert_close(ad[:, 2], np.sqrt(o), 1e-9, 0.0, 'Allan deviation of the logged gyro z')
    assert 'vel' in sim.results(err_stats_start=-1, extra_opt='ned')
    with pytest.raises(ValueError):
        sim.run(2)          # there is no gyro-1.csv


@pytest.mark.parametrize('rf', [0, 1])
def test_saved_run_reads_back_as_logged_data(gpu, tmp_path, rf):
    """Monte-Carlo run -> save_data (the reference's csv format) -> a second Sim on that directory:
    the same algorithm on the read-back sensor data reproduces the histories and the statistics."""
    from gnss_ins_sim_b200 import imu_model
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration
    g = load_golden('philox_90deg_mid_rf%d.npz' % rf)
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    sim = Sim([100.0, 0.0, 0.0], _traj(g), ref_frame=rf, imu=imu, algorithm=FreeIntegration(g['ini']),
              seed=int(g['seed']))
    sim.run(3)
    d = str(tmp_path / 'saved')
    files = sim.save_data(d, names=['time', 'ref_pos', 'ref_vel', 'ref_att_euler', 'gyro', 'accel'])
    assert len(files) == 4 + 2 * 3
    again = Sim([100.0, 0.0, 0.0], d, ref_frame=rf, imu=None, algorithm=FreeIntegration(g['ini']))
    again.run(3)
    for r in range(3):
        k = 'algo0_%d' % r
        assert np.abs(wrap_pi(again.get_data(['att_euler'])[0][k] - sim.get_data(['att_euler'])[0][k])).max() < 1e-9
        assert_close(again.get_data(['pos'])[0][k], sim.get_data(['pos'])[0][k], 1e-9, 1.0 if rf == 1 else 1e-3, 'pos')   # LLA: 1e-12 rad = 6 um (the csv keeps 17 digits; the two launches differ in the last bit)
        assert_close(again.get_data(['vel'])[0][k], sim.get_data(['vel'])[0][k], 1e-9, 1.0, 'vel')
    for dn in ('att_euler', 'pos', 'vel'):
        a, b = again.get_error_stats(dn, -1), sim.get_error_stats(dn, -1)
        for kk in ('max', 'avg', 'std'):
            assert_close(a[kk], b[kk], 1e-6, 1e-6 if dn != 'pos' or rf == 1 else 1e-9, '%s %s' % (dn, kk))
