"""ORACLE (test infrastructure) -- the loosely-coupled GNSS/INS filter of BASELINE config 5.

PARITY UNPINNED BY THE REFERENCE: demo_algorithms/ins_loose.py is a stub (prediction() and
correction() are `pass`, ins_loose.py:124-134; the demo prints "Still under development",
demo_ins_loose.py:58-60), so there is nothing to freeze golden vectors from.  This file is the
first-principles SPEC of the filter the CUDA kernel (csrc/ekf_kernel.cuh) implements, vectorised over
runs; the kernel is held to it on identical Philox draws, and the filter itself is validated
statistically (NEES, 3-sigma containment: tests/test_ekf_consistency.py).  What IS taken from the
reference: the plugin interface (input / output names, ins_loose.py:31-33), the sensor models that
feed it (pathgen.acc_gen / gyro_gen / gps_gen, restated in oracle_np), the LLA / NED conventions and
the strapdown mechanization (free_integration.py:133-172, ref_frame 0).

The filter: 15 error states, closed loop.
  nominal state   att (ZYX Euler), vel (NED), pos (lat, lon, alt), gyro bias bg, accel bias ba;
                  propagated with the free-integration step on bias-corrected measurements.
  error state     x = [dp (N,E,D metres), dv (NED), phi (NED misalignment), dbg, dba], estimate minus
                  truth, with  C_hat(b->n) = (I - [phi x]) C(b->n):
                      dp'  = dv
                      dv'  = [f_n x] phi - C dba + C w_a           f_n = C (accel - ba)
                      phi' = C dbg - C w_g
                      dbg' = -dbg / tau_g - q_g,   dba' = -dba / tau_a - q_a    (first-order Gauss-Markov)
                  Phi = I + F dt;  Q: C diag(vrw^2) C^T dt, C diag(arw^2) C^T dt, and the bias drives
                  exactly as the generator draws them (b^2 = drift^2 (1 - exp(-2 dt / tau)),
                  a = 1 - dt / tau, pathgen.py:583-586); plus vel_rw^2 dt I and att_rw^2 dt I, the
                  model-mismatch random walks (the reference's truth generator and its first-order
                  mechanization disagree: noise-free free integration of motion_def-ins.csv drifts by
                  0.37 m/s; the product's default vel_rw = 0.02 m/s/sqrt(s) keeps the filter consistent).
  measurement     every GPS sample (visibility 1): z = [p_hat - p_gps (metres NED); v_hat - v_gps],
                  H = [I 0 0 0 0; 0 I 0 0 0], R = diag(stdp^2, stdv^2): six scalar updates, then the
                  nominal state is corrected (p, v, biases subtract; C <- (I + [phi x]) C_hat; Euler
                  angles re-extracted) and the error state is zero again.
"""
import math

import numpy as np

import oracle_np as onp


def skew(v):
    z = np.zeros(v.shape[0])
    return np.stack([np.stack([z, -v[:, 2], v[:, 1]], 1), np.stack([v[:, 2], z, -v[:, 0]], 1),
                     np.stack([-v[:, 1], v[:, 0], z], 1)], 1)


def dcm2euler_zyx(c):
    """n->b DCM (attitude.euler2dcm 'zyx' layout) -> [yaw, pitch, roll]."""
    return np.stack([np.arctan2(c[:, 0, 1], c[:, 0, 0]), -np.arcsin(np.clip(c[:, 0, 2], -1.0, 1.0)),
                     np.arctan2(c[:, 1, 2], c[:, 2, 2])], 1)


def nav_step_rf0(att, pos, vel, w, f, dt, earth_rot=True):
    """One step of FreeIntegration.run in ref_frame 0 (free_integration.py:133-172), batched."""
    c_bn = onp.euler2dcm_zyx(att)
    rm, rn, g, sl, cl = onp.geo_param(pos[:, 0], pos[:, 2])
    rm_e, rn_e = rm + pos[:, 2], rn + pos[:, 2]
    w_en = np.stack([vel[:, 1] / rn_e, -vel[:, 0] / rm_e, -vel[:, 1] * sl / cl / rn_e], 1)
    w_ie = np.zeros_like(w_en)
    if earth_rot:
        w_ie[:, 0] = onp.W_IE * cl
        w_ie[:, 2] = -onp.W_IE * sl
    w_nb = w - onp._mv(c_bn, w_en + w_ie)
    att_new = onp.euler_update_zyx(att, w_nb, dt)
    g_n = np.zeros_like(vel)
    g_n[:, 2] = g
    vel_dot = onp._mtv(c_bn, f) + g_n - onp.cross3(2 * w_ie + w_en, vel)
    vel_new = vel + vel_dot * dt
    pos_new = pos.copy()
    pos_new[:, 0] += vel[:, 0] / rm_e * dt
    pos_new[:, 1] += vel[:, 1] / rn_e / cl * dt
    pos_new[:, 2] += -vel[:, 2] * dt
    return att_new, pos_new, vel_new


def default_p0(gyro_err, accel_err, gps_err, ini_att_std):
    """Initial covariance: GPS-grade position / velocity, the given attitude uncertainty, biases at
    their steady-state drift plus the constant bias the filter does not know."""
    bg = np.asarray(gyro_err['b_drift']) ** 2 + np.asarray(gyro_err['b']) ** 2
    ba = np.asarray(accel_err['b_drift']) ** 2 + np.asarray(accel_err['b']) ** 2
    return np.concatenate([np.asarray(gps_err['stdp'], dtype=np.float64) ** 2,
                           np.asarray(gps_err['stdv'], dtype=np.float64) ** 2,
                           np.asarray(ini_att_std, dtype=np.float64) ** 2, bg, ba])


INI_DRAW = 27          # Philox draw ids 27..31 (t = 0xFFFFFFFE): the initial-state errors


def initial_errors(run_ids, seed, p0):
    """[R, 9] initial position (m, NED), velocity, misalignment errors ~ N(0, P0): pairs
    (t = 0xFFFFFFFE, draw 27 + j), j = 0..4 -> z0, z1 flattened, the first nine."""
    run_ids = np.asarray(run_ids, dtype=np.uint64)
    t = np.full(1, 0xFFFFFFFE, dtype=np.uint64)[None, :]
    z = np.empty((run_ids.size, 10))
    for j in range(5):
        z0, z1 = onp.normal_pair(t, INI_DRAW + j, run_ids[:, None], seed)
        z[:, 2 * j], z[:, 2 * j + 1] = z0[:, 0], z1[:, 0]
    return z[:, :9] * np.sqrt(p0[:9])[None, :]


def ins_loose(fs, ref_gyro, ref_accel, ref_nav, ref_gps, gps_idx, gps_vis, gyro_err, accel_err, gps_err,
              seed, run_ids, ini, ini_att_std=(0.02, 0.005, 0.005), earth_rot=True, stats_start=0,
              want_hist=False, vel_rw=0.0, att_rw=0.0):
    """Monte-Carlo loosely-coupled filter, runs vectorised.
    ref_nav [n, 9] = true att, pos (LLA), vel (NED); ref_gps [m, 6]; gps_idx [m] IMU sample index of
    every GPS row; ini (9,) true initial LLA, body velocity, Euler angles.
    Returns dict: end_err [R, 9] (att, pos LLA, vel: estimate - truth at n-1), end_bias [R, 6],
    nees [R, 3] mean NEES of position / velocity / attitude blocks over the GPS epochs >= stats_start
    (after the update), inside3 [R, 15] fraction of those epochs with |error| <= 3 sigma, and with
    want_hist the histories att/pos/vel/wb/ab [R, n, 3]."""
    run_ids = np.asarray(run_ids)
    R, n = run_ids.size, ref_gyro.shape[0]
    dt = 1.0 / fs
    # ---- truth generators: exactly the Monte-Carlo sensor data of the free-integration path ----
    z = onp.noise_normals(n, run_ids, seed)
    accel = onp.sensor_gen(fs, ref_accel, accel_err, 'vrw', z['acc_gm'], z['acc_w'])
    gyro = onp.sensor_gen(fs, ref_gyro, gyro_err, 'arw', z['gyr_gm'], z['gyr_w'])
    bias_g = np.asarray(gyro_err['b'])[None, None] + onp.bias_drift(gyro_err['b_corr'], gyro_err['b_drift'], n, fs, z['gyr_gm'])
    bias_a = np.asarray(accel_err['b'])[None, None] + onp.bias_drift(accel_err['b_corr'], accel_err['b_drift'], n, fs, z['acc_gm'])
    m = ref_gps.shape[0]
    gps = onp.gps_gen(ref_gps, gps_err, 0, onp.gps_normals(m, run_ids, seed))
    # ---- filter constants --------------------------------------------------------------------
    a_g, b_g = onp.gm_coeffs(gyro_err['b_corr'], gyro_err['b_drift'], fs)
    a_a, b_a = onp.gm_coeffs(accel_err['b_corr'], accel_err['b_drift'], fs)
    white_g = np.isinf(np.asarray(gyro_err['b_corr'], dtype=np.float64))
    white_a = np.isinf(np.asarray(accel_err['b_corr'], dtype=np.float64))
    a_g, b_g = np.where(white_g, 0.0, a_g), np.where(white_g, np.asarray(gyro_err['b_drift']), b_g)
    a_a, b_a = np.where(white_a, 0.0, a_a), np.where(white_a, np.asarray(accel_err['b_drift']), b_a)
    arw2 = np.asarray(gyro_err['arw'], dtype=np.float64) ** 2
    vrw2 = np.asarray(accel_err['vrw'], dtype=np.float64) ** 2
    r_diag = np.concatenate([np.asarray(gps_err['stdp'], dtype=np.float64) ** 2,
                             np.asarray(gps_err['stdv'], dtype=np.float64) ** 2])
    p0 = default_p0(gyro_err, accel_err, gps_err, ini_att_std)
    # ---- initial nominal state: truth + a draw from P0 ------------------------------------------
    e0 = initial_errors(run_ids, seed, p0)
    ini = np.asarray(ini, dtype=np.float64)
    att_t = np.tile(ini[6:9], (R, 1))
    c_t = onp.euler2dcm_zyx(att_t)                                   # n -> b
    vel_t = onp._mtv(c_t, np.tile(ini[3:6], (R, 1)))                 # NED
    rm, rn, _, _, cl = onp.geo_param(ini[0], ini[2])
    pos = np.tile(ini[0:3], (R, 1))
    pos[:, 0] += e0[:, 0] / (rm + ini[2])
    pos[:, 1] += e0[:, 1] / ((rn + ini[2]) * cl)
    pos[:, 2] -= e0[:, 2]
    vel = vel_t + e0[:, 3:6]
    # C_hat(b->n) = (I - [phi x]) C(b->n)  <=>  C_hat(n->b) = C(n->b) (I + [phi x])
    c_hat = np.einsum('rij,rjk->rik', c_t, np.eye(3)[None] + skew(e0[:, 6:9]))
    att = dcm2euler_zyx(c_hat)
    bg, ba = np.zeros((R, 3)), np.zeros((R, 3))
    P = np.tile(np.diag(p0), (R, 1, 1))
    j = 0                                    # next GPS row
    acc = {'nees': np.zeros((R, 3)), 'inside': np.zeros((R, 15)), 'cnt': 0}
    hist = {k: np.zeros((R, n, 3)) for k in ('att', 'pos', 'vel', 'wb', 'ab')} if want_hist else None
    I15 = np.eye(15)
    for i in range(n):
        # ---- GPS update for sample i ------------------------------------------------------------
        if j < m and gps_idx[j] == i:
            if gps_vis[j] > 0:
                rm, rn, _, _, cl = onp.geo_param(pos[:, 0], pos[:, 2])
                zm = np.empty((R, 6))
                zm[:, 0] = (pos[:, 0] - gps[:, j, 0]) * (rm + pos[:, 2])
                zm[:, 1] = (pos[:, 1] - gps[:, j, 1]) * (rn + pos[:, 2]) * cl
                zm[:, 2] = -(pos[:, 2] - gps[:, j, 2])
                zm[:, 3:6] = vel - gps[:, j, 3:6]
                x = np.zeros((R, 15))
                for k in range(6):                     # scalar updates, R is diagonal
                    s = P[:, k, k] + r_diag[k]
                    K = P[:, :, k] / s[:, None]
                    x = x + K * (zm[:, k] - x[:, k])[:, None]
                    P = P - K[:, :, None] * P[:, k, None, :]
                    P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
                # close the loop
                pos[:, 0] -= x[:, 0] / (rm + pos[:, 2])
                pos[:, 1] -= x[:, 1] / ((rn + pos[:, 2]) * cl)
                pos[:, 2] += x[:, 2]
                vel = vel - x[:, 3:6]
                c_nb = onp.euler2dcm_zyx(att)          # estimated n -> b
                # C(b->n) = (I + [phi x]) C_hat(b->n)  <=>  C(n->b) = C_hat(n->b) (I - [phi x])
                c_nb = np.einsum('rij,rjk->rik', c_nb, np.eye(3)[None] - skew(x[:, 6:9]))
                att = dcm2euler_zyx(c_nb)
                bg = bg - x[:, 9:12]
                ba = ba - x[:, 12:15]
            # ---- consistency of the (posterior) estimate at this epoch --------------------------
            if i >= stats_start:
                rm, rn, _, _, cl = onp.geo_param(ref_nav[i, 3], ref_nav[i, 5])
                e = np.zeros((R, 15))
                e[:, 0] = (pos[:, 0] - ref_nav[i, 3]) * (rm + ref_nav[i, 5])
                e[:, 1] = (pos[:, 1] - ref_nav[i, 4]) * (rn + ref_nav[i, 5]) * cl
                e[:, 2] = -(pos[:, 2] - ref_nav[i, 5])
                e[:, 3:6] = vel - ref_nav[i, 6:9]
                c_true = onp.euler2dcm_zyx(np.tile(ref_nav[i, 0:3], (R, 1)))     # n -> b
                c_est = onp.euler2dcm_zyx(att)
                # C_hat(b->n) C(n->b) = I - [phi x]
                mm = np.einsum('rji,rjk->rik', c_est, c_true)
                e[:, 6] = -0.5 * (mm[:, 2, 1] - mm[:, 1, 2])
                e[:, 7] = -0.5 * (mm[:, 0, 2] - mm[:, 2, 0])
                e[:, 8] = -0.5 * (mm[:, 1, 0] - mm[:, 0, 1])
                e[:, 9:12] = bg - bias_g[:, i]
                e[:, 12:15] = ba - bias_a[:, i]
                for b in range(3):
                    blk = slice(3 * b, 3 * b + 3)
                    acc['nees'][:, b] += np.einsum('ri,rij,rj->r', e[:, blk], np.linalg.inv(P[:, blk, blk]), e[:, blk])
                sig = np.sqrt(np.einsum('rii->ri', P))
                acc['inside'] += (np.abs(e) <= 3.0 * sig)
                acc['cnt'] += 1
            j += 1
        if want_hist:
            hist['att'][:, i], hist['pos'][:, i], hist['vel'][:, i] = att, pos, vel
            hist['wb'][:, i], hist['ab'][:, i] = bg, ba
        if i == n - 1:
            break
        # ---- prediction with the measurements of sample i -----------------------------------------
        w = gyro[:, i] - bg
        f = accel[:, i] - ba
        c_nb = onp.euler2dcm_zyx(att)                  # n -> b, of sample i
        f_n = onp._mtv(c_nb, f)
        c_bn = np.transpose(c_nb, (0, 2, 1))           # b -> n
        Phi = np.tile(I15, (R, 1, 1))
        Phi[:, 0:3, 3:6] = np.eye(3) * dt
        Phi[:, 3:6, 6:9] = skew(f_n) * dt
        Phi[:, 3:6, 12:15] = -c_bn * dt
        Phi[:, 6:9, 9:12] = c_bn * dt
        Phi[:, 9:12, 9:12] = np.diag(a_g)
        Phi[:, 12:15, 12:15] = np.diag(a_a)
        Q = np.zeros((R, 15, 15))
        Q[:, 3:6, 3:6] = np.einsum('rij,j,rkj->rik', c_bn, vrw2, c_bn) * dt + np.eye(3) * (vel_rw * vel_rw * dt)
        Q[:, 6:9, 6:9] = np.einsum('rij,j,rkj->rik', c_bn, arw2, c_bn) * dt + np.eye(3) * (att_rw * att_rw * dt)
        Q[:, 9:12, 9:12] = np.diag(b_g ** 2)
        Q[:, 12:15, 12:15] = np.diag(b_a ** 2)
        P = np.einsum('rij,rjk,rlk->ril', Phi, P, Phi) + Q
        att, pos, vel = nav_step_rf0(att, pos, vel, w, f, dt, earth_rot)
        bg = bg * a_g
        ba = ba * a_a
    end = ref_nav[n - 1]
    end_err = np.concatenate([onp.angle_range_pi(att - end[0:3]), pos - end[3:6], vel - end[6:9]], 1)
    cnt = max(acc['cnt'], 1)
    out = {'end_err': end_err, 'end_bias': np.concatenate([bg, ba], 1), 'nees': acc['nees'] / cnt,
           'inside3': acc['inside'] / cnt, 'epochs': acc['cnt'], 'P_diag_end': np.einsum('rii->ri', P)}
    if want_hist:
        out.update(hist)
    return out
