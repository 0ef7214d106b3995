/*
 * ORACLE (test infrastructure, not product): plain-C restatement of the Monte-Carlo
 * free-integration hot path of gnss-ins-sim, one run at a time, scalar double
 * arithmetic, libm.  Built by oracle/Makefile into oracle/_build/liboracle.so and used
 *   - by tests/ as the checker at sizes where the NumPy oracle is too slow,
 *   - by bench.py as the CPU baseline (`cpu_baseline.kind = "port"`, pthreads over runs).
 * The product (gnss_ins_sim_b200) never links or loads it.
 *
 * Pinned against the golden vectors of the unmodified reference in
 * tests/test_oracle_golden.py (via oracle/oracle_c.py).
 *
 * Each function cites the reference file:line it follows.
 * Compiled with -ffp-contract=off so that a*b+c rounds twice like NumPy does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* geoparams.py:18-23, :40-43 */
#define RE 6378137.0
#define FLATTENING (1.0 / 298.257223563)
#define ECC 0.0818191908426215
#define E_SQR (ECC * ECC)
#define W_IE 7292115e-11
#define NORMAL_GRAVITY 9.7803253359
#define K_GRAV 0.00193185265241
#define M_GRAV 0.00344978650684
#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)
#define HALF_PI (0.5 * PI)

typedef struct {
  double b[3], b_drift[3], b_corr[3], rw[3];
} orc_sensor_err;

/* geoparams.geo_param, geoparams.py:25-53 */
static void geo_param(double lat, double h, double* rm, double* rn, double* g, double* sl,
                      double* cl) {
  *sl = sin(lat);
  *cl = cos(lat);
  double sl_sqr = (*sl) * (*sl);
  *rm = (RE * (1 - E_SQR)) / (sqrt(1.0 - E_SQR * sl_sqr) * (1.0 - E_SQR * sl_sqr));
  *rn = RE / (sqrt(1.0 - E_SQR * sl_sqr));
  double g1 = NORMAL_GRAVITY * (1 + K_GRAV * sl_sqr) / sqrt(1.0 - E_SQR * sl_sqr);
  *g = g1 * (1.0 - (2.0 / RE) * (1.0 + FLATTENING + M_GRAV - 2.0 * FLATTENING * sl_sqr) * h +
             3.0 * h * h / RE / RE);
}

/* geoparams.lla2ecef, geoparams.py:70-87 */
static void lla2ecef(const double* lla, double* xyz) {
  double sl = sin(lla[0]), cl = cos(lla[0]);
  double r = RE / sqrt(1.0 - E_SQR * sl * sl);
  double rho = (r + lla[2]) * cl;
  xyz[0] = rho * cos(lla[1]);
  xyz[1] = rho * sin(lla[1]);
  xyz[2] = (r * (1.0 - E_SQR) + lla[2]) * sl;
}

/* attitude.euler2dcm 'zyx', attitude.py:361-371 */
static void euler2dcm(const double* a, double c[3][3]) {
  double c0 = cos(a[0]), c1 = cos(a[1]), c2 = cos(a[2]);
  double s0 = sin(a[0]), s1 = sin(a[1]), s2 = sin(a[2]);
  c[0][0] = c1 * c0;
  c[0][1] = c1 * s0;
  c[0][2] = -s1;
  c[1][0] = s2 * s1 * c0 - c2 * s0;
  c[1][1] = s2 * s1 * s0 + c2 * c0;
  c[1][2] = c1 * s2;
  c[2][0] = s1 * c2 * c0 + s0 * s2;
  c[2][1] = s1 * c2 * s0 - c0 * s2;
  c[2][2] = c1 * c2;
}

/* attitude.euler_update_zyx, attitude.py:679-721 */
static void euler_update_zyx(const double* x, const double* w, double dt, double* y) {
  double c_psi = cos(x[2]), s_psi = sin(x[2]);
  double phi_dot = (w[2] * c_psi + w[1] * s_psi) / cos(x[1]);
  double theta_dot = w[1] * c_psi - w[2] * s_psi;
  double psi_dot = w[0] + (w[2] * c_psi + w[1] * s_psi) * tan(x[1]);
  y[0] = x[0] + phi_dot * dt;
  y[1] = x[1] + theta_dot * dt;
  y[2] = x[2] + psi_dot * dt;
  if (y[1] > HALF_PI) {
    y[1] = PI - y[1];
    y[0] = y[0] + PI;
    y[2] = y[2] + PI;
  } else if (y[1] < -HALF_PI) {
    y[1] = -PI - y[1];
    y[0] = y[0] + PI;
    y[2] = y[2] + PI;
  }
  if (y[0] > PI)
    y[0] = y[0] - TWO_PI;
  else if (y[0] < -PI)
    y[0] = y[0] + TWO_PI;
  if (y[2] > PI)
    y[2] = y[2] - TWO_PI;
  else if (y[2] < -PI)
    y[2] = y[2] + TWO_PI;
}

/* attitude.cross3, attitude.py:758-770 */
static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

static void mv(double c[3][3], const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = c[i][0] * v[0] + c[i][1] * v[1] + c[i][2] * v[2];
}
static void mtv(double c[3][3], const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = c[0][i] * v[0] + c[1][i] * v[1] + c[2][i] * v[2];
}

/* attitude.angle_range_pi, attitude.py:799-812 (python float %) */
static double angle_range_pi(double x) {
  double m = fmod(x, TWO_PI);
  if (m < 0) m += TWO_PI;
  if (m > PI) m -= TWO_PI;
  return m;
}

/*
 * FreeIntegration.run, demo_algorithms/free_integration.py:63-174, one run.
 * gyro/accel: sample t at gyro[t*3 + c].  ini: 9 or 10 values.  att/pos/vel (nullable):
 * (n,3) histories.  end (nullable): 9 values att,pos,vel at n-1.
 */
void orc_free_integration_run(int ref_frame, double fs, int64_t n, const double* gyro,
                              const double* accel, const double* ini, int ini_rows, int earth_rot,
                              double* att, double* pos, double* vel, double* end) {
  const double dt = 1.0 / fs;
  double a[3] = {ini[6], ini[7], ini[8]};
  double p[3], v[3], vb[3] = {ini[3], ini[4], ini[5]};
  double c_bn[3][3];
  double g_n[3] = {0, 0, 0};
  euler2dcm(a, c_bn);
  mtv(c_bn, vb, v);
  if (ref_frame == 1) {
    double rm, rn, g, sl, cl;
    if (ini_rows > 9) {
      g_n[2] = ini[9];
    } else {
      geo_param(ini[0], ini[2], &rm, &rn, &g, &sl, &cl);
      g_n[2] = g;
    }
    lla2ecef(ini, p);
  } else {
    p[0] = ini[0];
    p[1] = ini[1];
    p[2] = ini[2];
  }
  double w_en_n[3] = {0, 0, 0}, w_ie_n[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    if (i > 0) {
      const double* w = gyro + (i - 1) * 3;
      const double* f = accel + (i - 1) * 3;
      double a_new[3];
      if (ref_frame == 1) {
        /* :104-116 */
        double cg[3], wxv[3], v_old[3] = {v[0], v[1], v[2]};
        euler_update_zyx(a, w, dt, a_new);
        mv(c_bn, g_n, cg);
        cross3(w, vb, wxv);
        for (int k = 0; k < 3; ++k) vb[k] = vb[k] + (f[k] + cg[k]) * dt - wxv[k] * dt;
        euler2dcm(a_new, c_bn);
        mtv(c_bn, vb, v);
        for (int k = 0; k < 3; ++k) p[k] = p[k] + v_old[k] * dt;
      } else {
        /* :133-172 */
        double rm, rn, g, sl, cl;
        geo_param(p[0], p[2], &rm, &rn, &g, &sl, &cl);
        double rm_e = rm + p[2], rn_e = rn + p[2];
        g_n[2] = (ini_rows > 9) ? ini[9] : g;
        w_en_n[0] = v[1] / rn_e;
        w_en_n[1] = -v[0] / rm_e;
        w_en_n[2] = -v[1] * sl / cl / rn_e;
        if (earth_rot) {
          w_ie_n[0] = W_IE * cl;
          w_ie_n[2] = -W_IE * sl;
        }
        double wsum[3] = {w_en_n[0] + w_ie_n[0], w_en_n[1] + w_ie_n[1], w_en_n[2] + w_ie_n[2]};
        double cw[3], w_nb_b[3], fa[3], w2[3], cor[3];
        mv(c_bn, wsum, cw);
        for (int k = 0; k < 3; ++k) w_nb_b[k] = w[k] - cw[k];
        euler_update_zyx(a, w_nb_b, dt, a_new);
        mtv(c_bn, f, fa);
        for (int k = 0; k < 3; ++k) w2[k] = 2 * w_ie_n[k] + w_en_n[k];
        cross3(w2, v, cor);
        double v_old[3] = {v[0], v[1], v[2]};
        for (int k = 0; k < 3; ++k) v[k] = v_old[k] + (fa[k] + g_n[k] - cor[k]) * dt;
        p[0] = p[0] + v_old[0] / rm_e * dt;
        p[1] = p[1] + v_old[1] / rn_e / cl * dt;
        p[2] = p[2] + (-v_old[2]) * dt;
        euler2dcm(a_new, c_bn);
      }
      a[0] = a_new[0];
      a[1] = a_new[1];
      a[2] = a_new[2];
    }
    if (att) {
      for (int k = 0; k < 3; ++k) {
        att[i * 3 + k] = a[k];
        pos[i * 3 + k] = p[k];
        vel[i * 3 + k] = v[k];
      }
    }
  }
  if (end) {
    for (int k = 0; k < 3; ++k) {
      end[k] = a[k];
      end[3 + k] = p[k];
      end[6 + k] = v[k];
    }
  }
}

/* ---- b2ins noise spec: Philox4x32-10 + Box-Muller (see oracle_np.py) ---- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = (uint32_t)p1;
    c[2] = n2;
    c[3] = (uint32_t)p0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void orc_philox(uint32_t* c, uint32_t k0, uint32_t k1) { philox4x32_10(c, k0, k1); }

static void normal_pair(uint32_t t, uint32_t draw, uint64_t run, uint64_t seed, double* z0,
                        double* z1) {
  uint32_t c[4] = {t, draw, (uint32_t)run, (uint32_t)(run >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t a = ((uint64_t)c[1] << 32) | c[0];
  uint64_t b = ((uint64_t)c[3] << 32) | c[2];
  double u1 = 1.0 - (double)(a >> 12) * 0x1p-52;
  double u2 = (double)(b >> 12) * 0x1p-52;
  double r = sqrt(-2.0 * log(u1));
  double th = TWO_PI * u2;
  *z0 = r * cos(th);
  *z1 = r * sin(th);
}

/*
 * pathgen.acc_gen + gyro_gen + bias_drift (pathgen.py:441-594) for one run, no vibration:
 * meas = ref + b + drift + white, summed in that order (pathgen.py:500, :562).
 */
void orc_imu_noise_run(double fs, int64_t n, const double* ref_gyro, const double* ref_accel,
                       const orc_sensor_err* ge, const orc_sensor_err* ae, uint64_t seed,
                       uint64_t run, double* gyro, double* accel) {
  const double dt = 1.0 / fs;
  const orc_sensor_err* errs[2] = {ae, ge};
  const double* refs[2] = {ref_accel, ref_gyro};
  double* outs[2] = {accel, gyro};
  for (int s = 0; s < 2; ++s) {
    const orc_sensor_err* e = errs[s];
    double a[3], b[3], d[3] = {0, 0, 0}, w[3];
    int white[3];
    for (int c = 0; c < 3; ++c) {
      white[c] = isinf(e->b_corr[c]);
      a[c] = 1 - 1 / fs / e->b_corr[c];                                  /* :583 */
      b[c] = e->b_drift[c] * sqrt(1.0 - exp(-2 / (fs * e->b_corr[c])));  /* :586 */
      w[c] = e->rw[c] / sqrt(dt);                                        /* :496 */
    }
    for (int64_t t = 0; t < n; ++t) {
      for (int c = 0; c < 3; ++c) {
        double zg, zw;
        normal_pair((uint32_t)t, (uint32_t)(3 * s + c), run, seed, &zg, &zw);
        double drift = white[c] ? e->b_drift[c] * zg : d[c];
        double white_noise = w[c] * zw;
        outs[s][t * 3 + c] = refs[s][t * 3 + c] + e->b[c] + drift + white_noise;
        if (!white[c]) d[c] = a[c] * d[c] + b[c] * zg;                   /* :589-590 */
      }
    }
  }
}

/*
 * Loops A + B of Sim.run (ins_sim.py:490-506, ins_algo_manager.py:73-95) + per-run end-point
 * error (ins_data_manager.py:536-541) for runs [run0, run0+runs): worker threads pull runs
 * from an atomic counter.  ref_nav_end: att,pos,vel of the true trajectory at sample n-1.
 * ini: [ini_sets][ini_rows].  end_err [runs][9].  threads <= 0: one per online core.
 * Returns the number of threads used.
 */
typedef struct {
  int ref_frame, ini_sets, ini_rows, earth_rot;
  double fs;
  int64_t n, runs, run0, ini_offset;
  const double *ref_gyro, *ref_accel, *ref_nav_end, *ini;
  const orc_sensor_err *ge, *ae;
  uint64_t seed;
  double* end_err;
  atomic_llong next;
} mc_job;

static void* mc_worker(void* arg) {
  mc_job* j = (mc_job*)arg;
  double* gyro = (double*)malloc(sizeof(double) * j->n * 3);
  double* accel = (double*)malloc(sizeof(double) * j->n * 3);
  for (;;) {
    int64_t r = (int64_t)atomic_fetch_add(&j->next, 1);
    if (r >= j->runs) break;
    double end[9];
    int64_t irun = j->ini_offset + r;
    int64_t set = irun < j->ini_sets ? irun : 0; /* free_integration.py:85-87 */
    orc_imu_noise_run(j->fs, j->n, j->ref_gyro, j->ref_accel, j->ge, j->ae, j->seed,
                      (uint64_t)(j->run0 + r), gyro, accel);
    orc_free_integration_run(j->ref_frame, j->fs, j->n, gyro, accel, j->ini + set * j->ini_rows,
                             j->ini_rows, j->earth_rot, NULL, NULL, NULL, end);
    for (int k = 0; k < 3; ++k) {
      j->end_err[r * 9 + k] = angle_range_pi(end[k] - j->ref_nav_end[k]);
      j->end_err[r * 9 + 3 + k] = end[3 + k] - j->ref_nav_end[3 + k];
      j->end_err[r * 9 + 6 + k] = end[6 + k] - j->ref_nav_end[6 + k];
    }
  }
  free(gyro);
  free(accel);
  return NULL;
}

int orc_mc_free_integration(int ref_frame, double fs, int64_t n, int64_t runs, int64_t run0,
                            int64_t ini_offset, const double* ref_gyro, const double* ref_accel,
                            const double* ref_nav_end, const orc_sensor_err* ge,
                            const orc_sensor_err* ae, uint64_t seed, const double* ini,
                            int ini_sets, int ini_rows, int earth_rot, double* end_err,
                            int threads) {
  if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (threads < 1) threads = 1;
  if (threads > 1024) threads = 1024;
  if ((int64_t)threads > runs) threads = runs > 0 ? (int)runs : 1;
  mc_job job = {ref_frame, ini_sets, ini_rows, earth_rot, fs, n, runs, run0, ini_offset,
                ref_gyro, ref_accel, ref_nav_end, ini, ge, ae, seed, end_err, 0};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  int started = 0;
  for (int i = 1; i < threads; ++i)
    if (pthread_create(&th[started], NULL, mc_worker, &job) == 0) ++started;
  mc_worker(&job);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  free(th);
  return started + 1;
}

/* InsDataMgr.__array_stats, ins_data_manager.py:797-808: stats[3][nc] = max|e|, mean, std(ddof 0) */
void orc_array_stats(int64_t runs, int nc, const double* err, double* stats) {
  for (int c = 0; c < nc; ++c) {
    double mx = 0, s = 0;
    for (int64_t r = 0; r < runs; ++r) {
      double e = err[r * nc + c];
      if (fabs(e) > mx) mx = fabs(e);
      s += e;
    }
    double mean = s / (double)runs, q = 0;
    for (int64_t r = 0; r < runs; ++r) {
      double d = err[r * nc + c] - mean;
      q += d * d;
    }
    stats[c] = mx;
    stats[nc + c] = mean;
    stats[2 * nc + c] = sqrt(q / (double)runs);
  }
}

/* allan.allan_var, allan.py:18-59, for one strided series; returns ntau (avar/tau sized >= 128) */
int orc_allan_var(const double* x, int64_t n, int64_t stride, double fs, double* avar,
                  double* tau) {
  double ts = 1.0 / fs;
  int64_t max_bin = (int64_t)floor(n / 9.0);
  if (max_bin * ts < 1) return 0;
  int nextpow10 = (int)ceil(log10((double)max_bin));
  int ntau = 0;
  double scale = 0.1;
  for (int i = 0; i < nextpow10; ++i) {
    scale *= 10;
    for (int j = 1; j < 10; ++j) {
      int64_t m = (int64_t)(j * scale);
      if (m > max_bin) break;
      int64_t nb = n / m;
      if (nb < 9) break;
      double prev = 0, acc = 0;
      for (int64_t b = 0; b < nb; ++b) {
        double s = 0;
        for (int64_t q = 0; q < m; ++q) s += x[(b * m + q) * stride];
        double mean = s / (double)m;
        if (b > 0) acc += (mean - prev) * (mean - prev);
        prev = mean;
      }
      avar[ntau] = 0.5 / (double)(nb - 1) * acc;
      tau[ntau] = (double)m * ts;
      ++ntau;
    }
  }
  return ntau;
}
