// b2ins C ABI (include/b2ins.h): argument checking, host-side pre-digestion of the error
// models, kernel dispatch, and the *_host convenience wrappers.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/b2ins.h"
#include "allan_kernel.cuh"
#include "internal.h"
#include "noise_kernel.cuh"
#include "pathgen_host.h"
#include "psd_kernel.cuh"
#include "gps_kernel.cuh"
#include "ekf_kernel.cuh"
#include "stats_kernel.cuh"

#ifdef B2INS_SINGLE_TU
#define B2_RF 0
#define B2_PLAIN_NAME launch_mc_plain_rf0
#define B2_SPEC_NAME launch_mc_spec_rf0
#include "mc_plain_launch.cuh"
#include "mc_spec_launch.cuh"
#undef B2_RF
#undef B2_PLAIN_NAME
#undef B2_SPEC_NAME
#define B2_RF 1
#define B2_PLAIN_NAME launch_mc_plain_rf1
#define B2_SPEC_NAME launch_mc_spec_rf1
#include "mc_plain_launch.cuh"
#include "mc_spec_launch.cuh"
#endif

using namespace b2ins;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU_CHECK(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(B2INS_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                              \
  } while (0)

#define ARG_CHECK(cond, ...) \
  do {                       \
    if (!(cond)) return fail(B2INS_ERR_ARG, __VA_ARGS__); \
  } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// bias_drift coefficients exactly as written at pathgen.py:583-586 (a: first-order
// approximation, b: exact exponential) and the white-noise scale rw/sqrt(dt) (:496-498).
int digest_triad(const b2ins_sensor_err* e, const b2ins_vib* v, double fs, TriadNoise* out) {
  const double dt = 1.0 / fs;
  for (int c = 0; c < 3; ++c) {
    out->b[c] = e->b[c];
    if (std::isinf(e->b_corr[c])) {
      out->gm_a[c] = 0.0;
      out->gm_b[c] = 0.0;
      out->wd[c] = e->b_drift[c];
    } else {
      out->gm_a[c] = 1 - 1 / fs / e->b_corr[c];
      out->gm_b[c] = e->b_drift[c] * std::sqrt(1.0 - std::exp(-2 / (fs * e->b_corr[c])));
      out->wd[c] = 0.0;
    }
    out->w[c] = e->rw[c] / std::sqrt(dt);
  }
  out->vib_type = B2INS_VIB_NONE;
  out->series_len = 0;
  out->series = nullptr;
  out->vib_w = 0.0;
  for (int c = 0; c < 3; ++c) out->vib_amp[c] = 0.0;
  if (v && v->type != B2INS_VIB_NONE) {
    ARG_CHECK(v->type >= 1 && v->type <= 3, "vib type %d is not one of B2INS_VIB_*", v->type);
    out->vib_type = v->type;
    for (int c = 0; c < 3; ++c) out->vib_amp[c] = v->amp[c];
    out->vib_w = 2.0 * M_PI * v->freq * dt;  // np.sin(2.0*math.pi*freq*dt*k), pathgen.py:491
    if (v->type == B2INS_VIB_SERIES) {
      ARG_CHECK(v->series && v->series_len > 0, "VIB_SERIES needs series and series_len > 0");
      out->series = v->series;
      out->series_len = v->series_len;
    }
  }
  return B2INS_OK;
}

void layout_strides(int layout, int64_t runs, int64_t n, int64_t* sr, int64_t* st, int64_t* sc) {
  if (layout == B2INS_LAYOUT_RUN_MAJOR) {
    *sr = n * 3;
    *st = 3;
    *sc = 1;
  } else if (layout == B2INS_LAYOUT_CHANNEL_MAJOR) {
    *sr = n * 3;
    *st = 1;
    *sc = n;
  } else {
    *sr = 1;
    *st = 3 * runs;
    *sc = runs;
  }
}

int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      cached = 148;
  }
  return cached;
}

// tools: B2INS_MC_SHAPE="P,WI,split" overrides the specialised shape of the chosen G ("0" = the
// single-warp form); read at every launch so that one process can sweep
bool shape_override(McShape* sh) {
  const char* e = std::getenv("B2INS_MC_SHAPE");
  if (!e || !*e) return false;
  int P = 0, WI = 0, split = 0;
  const int got = std::sscanf(e, "%d,%d,%d", &P, &WI, &split);
  if (got >= 1 && P == 0) {
    sh->spec = false;
    return true;
  }
  if (got == 3) {
    sh->P = P;
    sh->WI = WI;
    sh->split = split != 0;
    sh->spec = true;
    return true;
  }
  return false;
}

// The specialised shapes that are instantiated, by group width: four-warp CTAs (one SM sub-partition
// per warp) where the group is narrow, the paired layout (producer and integrator of a group on the
// same sub-partition) for the widest groups.
// (ref_frame 1, groups of 4 and 8 lanes: the step itself split over an attitude and a velocity warp,
// mc_av_kernel.cuh, key "6,2,0" -- measured 0.167 against 0.183 ms at 1000 runs, 0.146 against 0.177 at 500)
McShape default_shape(int G, int rf) {
  if (rf == 1 && (G == 4 || G == 8)) return McShape{G, 6, 2, false, true};
  switch (G) {
    case 1: return McShape{1, 6, 1, false, true};
    case 2: return McShape{2, 6, 1, false, true};
    case 4: return McShape{4, 6, 1, false, true};
    case 8: return McShape{8, 6, 1, false, true};
    case 16: return McShape{16, 1, 4, true, true};
    default: return McShape{32, 1, 4, true, true};
  }
}

// Lanes per run.  The recurrence is serial in time, so the only parallelism is across runs (and,
// for the noise, across the samples and channels the producers take).  One integrator warp needs
// ~350 cycles per step (ref_frame 1; its attitude half alone ~273) whatever the number of runs it carries
// -- the dependent-issue latency of a single warp -- so with few runs the narrowest group that still gives
// every SM a CTA wins: fewer replicated lanes, fewer producer jobs per step.  With many runs G = 1 is the
// throughput form, and beyond ~2.6e5 runs the single-warp kernel (every warp generates and integrates, five
// CTAs per SM) overtakes the specialised one.  Measured on B200 (profiles/spec2_probe_r02.jsonl and
// kernel_bench_r02.jsonl, run-steps/s at n = 1000, ref_frame 1): 500 runs G = 8 3.4e9; 1000 runs G = 4
// 6.0e9; 2000 runs G = 2 6.9e9; 100 000 runs G = 1 1.37e10; 10^6 runs single-warp form 1.95e10.
// spec_ok: the launch can take the warp-specialised form (fused noise, end-point statistics only)
constexpr int64_t kSpecMaxRunsG1 = int64_t(1) << 18;

int auto_lanes(int64_t runs, bool spec_ok, int sm_override = 0) {
  const int64_t sms = sm_override > 0 ? sm_override : sm_count();
  if (spec_ok) {
    // runs per CTA of the specialised shapes: 32 / G
    if (runs <= sms * 4) return 8;
    if (runs <= sms * 8) return 4;
    if (runs <= sms * 32) return 2;
    return 1;
  }
  // supplied data / process statistics (single-warp form): about one warp per SM sub-partition
  const int64_t smsp = sms * 4;
  int g = 32;
  while (g > 1 && runs * g * 2 > smsp * 32 * 3) g >>= 1;   // warps <= 1.5 per sub-partition
  return g;
}

int launch_mc(const McParams& p, int lanes, int rf, bool fed, bool proc, cudaStream_t s) {
  if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
    return fail(B2INS_ERR_ARG, "lanes_per_run must be 0,1,2,4,8,16 or 32, got %d", lanes);
  if (!fed && !proc && p.algo == 0 && !(lanes == 1 && p.runs >= kSpecMaxRunsG1 && !std::getenv("B2INS_MC_SHAPE"))) {
    // fused noise, end-point results (and histories): the warp-specialised form
#ifdef B2INS_PHASE_CLOCKS
    if (const char* e = std::getenv("B2INS_MC_DEBUG")) const_cast<McParams&>(p).debug = std::atoi(e);
#endif
    McShape sh = default_shape(lanes, rf);
    shape_override(&sh);
    if (sh.spec) {
      sh.G = lanes;
      const bool ok = (rf == 1) ? launch_mc_spec_rf1(p, sh, s) : launch_mc_spec_rf0(p, sh, s);
      if (!ok)
        return fail(B2INS_ERR_ARG, "no specialised kernel for lanes=%d P=%d WI=%d split=%d", sh.G, sh.P,
                    sh.WI, static_cast<int>(sh.split));
      CU_CHECK(cudaGetLastError());
      return B2INS_OK;
    }
  }
  if (rf == 1)
    launch_mc_plain_rf1(p, lanes, fed, proc, s);
  else
    launch_mc_plain_rf0(p, lanes, fed, proc, s);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  double* d() const { return static_cast<double*>(p); }
};

struct Stream {
  cudaStream_t s = nullptr;
  ~Stream() {
    if (s) cudaStreamDestroy(s);
  }
  cudaError_t create() { return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
};

}  // namespace

extern "C" {

int b2ins_version(void) { return B2INS_VERSION; }

const char* b2ins_last_error(void) { return g_err; }

int b2ins_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// allan.py:29-44
int b2ins_allan_num_tau(int64_t n, double fs, int64_t* m, int m_cap) {
  const double ts = 1.0 / fs;
  const int64_t max_bin = static_cast<int64_t>(std::floor(n / 9.0));
  if (max_bin * ts < 1) return 0;
  const int nextpow10 = static_cast<int>(std::ceil(std::log10(static_cast<double>(max_bin))));
  int count = 0;
  double scale = 0.1;
  for (int i = 0; i < nextpow10; ++i) {
    scale *= 10;
    for (int j = 1; j < 10; ++j) {
      const int64_t tmp = static_cast<int64_t>(j * scale);
      if (tmp <= max_bin) {
        if (m && count < m_cap) m[count] = tmp;
        ++count;
      } else {
        break;
      }
    }
  }
  return count;
}

// ---------------------------------------------------------------- K2 --------
static int free_integration_fed(int algo, int ref_frame, double fs, int64_t runs, int64_t n,
                                const double* gyro, const double* accel, int layout,
                                const double* ini, int ini_sets, int ini_rows, int64_t run_offset,
                                int earth_rot, double* att, double* pos, double* vel,
                                int lanes_per_run, void* stream) {
  ARG_CHECK(ref_frame == 0 || ref_frame == 1, "ref_frame must be 0 or 1, got %d", ref_frame);
  ARG_CHECK(fs > 0.0, "fs must be positive");
  ARG_CHECK(runs >= 0 && n >= 0, "runs and n must be non-negative");
  ARG_CHECK(layout == 0 || layout == 1, "layout must be B2INS_LAYOUT_*");
  ARG_CHECK(ini_sets >= 1 && (ini_rows == 9 || ini_rows == 10), "ini must be [sets>=1][9|10]");
  if (runs == 0 || n == 0) return B2INS_OK;
  ARG_CHECK(gyro && accel && ini && att && pos && vel, "null buffer");
  ARG_CHECK(n < (int64_t(1) << 32), "n must be < 2^32");
  McParams p;
  std::memset(&p, 0, sizeof(p));
  p.n = n;
  p.runs = runs;
  p.run_offset = run_offset;
  p.ini_offset = run_offset;
  p.dt = 1.0 / fs;
  p.earth_rot = earth_rot;
  p.ini = ini;
  p.ini_sets = ini_sets;
  p.ini_rows = ini_rows;
  p.fed_gyro = gyro;
  p.fed_accel = accel;
  layout_strides(layout, runs, n, &p.sr, &p.st, &p.sc);
  p.algo = algo;
  if (algo == 1) {   // `accel` carries the odometer series [R][n] or [n][R]
    p.fed_accel = nullptr;
    p.fed_odo = accel;
    p.so_r = (layout == B2INS_LAYOUT_RUN_MAJOR) ? n : 1;
    p.so_t = (layout == B2INS_LAYOUT_RUN_MAJOR) ? 1 : runs;
  }
  p.out_att = att;
  p.out_pos = pos;
  p.out_vel = vel;
  p.osr = p.sr;
  p.ost = p.st;
  p.osc = p.sc;
  p.dump_runs = runs;
  p.dump_stride = 1;
  p.dump_rows = n;
  p.stats_start = -1;
  int lanes = lanes_per_run;
  if (lanes == 0) {
    lanes = auto_lanes(runs, false);
    // run-major rows are 24-byte strided per lane when G = 1; a wider group reads whole
    // contiguous stretches of one run
    if (layout == B2INS_LAYOUT_RUN_MAJOR && lanes < 8) lanes = 8;
  }
  return launch_mc(p, lanes, ref_frame, true, false, static_cast<cudaStream_t>(stream));
}

int b2ins_free_integration_f64(int ref_frame, double fs, int64_t runs, int64_t n,
                               const double* gyro, const double* accel, int layout,
                               const double* ini, int ini_sets, int ini_rows, int64_t run_offset,
                               int earth_rot, double* att, double* pos, double* vel,
                               int lanes_per_run, void* stream) {
  return free_integration_fed(0, ref_frame, fs, runs, n, gyro, accel, layout, ini, ini_sets,
                              ini_rows, run_offset, earth_rot, att, pos, vel, lanes_per_run, stream);
}

int b2ins_free_integration_odo_f64(int ref_frame, double fs, int64_t runs, int64_t n,
                                   const double* gyro, const double* odo, int layout,
                                   const double* ini, int ini_sets, int ini_rows,
                                   int64_t run_offset, int earth_rot, double* att, double* pos,
                                   double* vel, int lanes_per_run, void* stream) {
  return free_integration_fed(1, ref_frame, fs, runs, n, gyro, odo, layout, ini, ini_sets,
                              ini_rows, run_offset, earth_rot, att, pos, vel, lanes_per_run, stream);
}

int b2ins_free_integration_f64_host(int ref_frame, double fs, int64_t runs, int64_t n,
                                    const double* gyro, const double* accel, int layout,
                                    const double* ini, int ini_sets, int ini_rows,
                                    int64_t run_offset, int earth_rot, double* att, double* pos,
                                    double* vel, int lanes_per_run) {
  ARG_CHECK(runs >= 0 && n >= 0, "runs and n must be non-negative");
  if (runs == 0 || n == 0) return B2INS_OK;
  ARG_CHECK(gyro && accel && ini && att && pos && vel, "null buffer");
  ARG_CHECK(ini_sets >= 1 && (ini_rows == 9 || ini_rows == 10), "ini must be [sets>=1][9|10]");
  const size_t bytes = static_cast<size_t>(runs) * n * 3 * sizeof(double);
  const size_t ini_bytes = static_cast<size_t>(ini_sets) * ini_rows * sizeof(double);
  DevBuf dg, da, di, oa, op, ov;
  Stream st;
  CU_CHECK(st.create());
  CU_CHECK(dg.alloc(bytes));
  CU_CHECK(da.alloc(bytes));
  CU_CHECK(di.alloc(ini_bytes));
  CU_CHECK(oa.alloc(bytes));
  CU_CHECK(op.alloc(bytes));
  CU_CHECK(ov.alloc(bytes));
  CU_CHECK(cudaMemcpyAsync(dg.p, gyro, bytes, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(da.p, accel, bytes, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(di.p, ini, ini_bytes, cudaMemcpyHostToDevice, st.s));
  const int rc = b2ins_free_integration_f64(ref_frame, fs, runs, n, dg.d(), da.d(), layout, di.d(),
                                            ini_sets, ini_rows, run_offset, earth_rot, oa.d(),
                                            op.d(), ov.d(), lanes_per_run, st.s);
  if (rc != B2INS_OK) return rc;
  CU_CHECK(cudaMemcpyAsync(att, oa.p, bytes, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaMemcpyAsync(pos, op.p, bytes, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaMemcpyAsync(vel, ov.p, bytes, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaStreamSynchronize(st.s));
  return B2INS_OK;
}

// ---------------------------------------------------------------- K1 --------
int b2ins_imu_noise_f64(double fs, int64_t runs, int64_t n, const double* ref_gyro,
                        const double* ref_accel, const b2ins_sensor_err* gyro_err,
                        const b2ins_sensor_err* accel_err, const b2ins_vib* vib_gyro,
                        const b2ins_vib* vib_accel, uint64_t seed, int64_t run_offset, int layout,
                        double* gyro, double* accel, double* z_dump, void* stream) {
  ARG_CHECK(fs > 0.0, "fs must be positive");
  ARG_CHECK(runs >= 0 && n >= 0, "runs and n must be non-negative");
  ARG_CHECK(layout >= 0 && layout <= 2, "layout must be B2INS_LAYOUT_*");
  if (runs == 0 || n == 0) return B2INS_OK;
  ARG_CHECK(ref_gyro && ref_accel && gyro_err && accel_err && gyro && accel, "null buffer");
  ARG_CHECK(n < (int64_t(1) << 32), "n must be < 2^32");
  NoiseParams p;
  std::memset(&p, 0, sizeof(p));
  p.n = n;
  p.runs = runs;
  p.run_offset = run_offset;
  p.dt = 1.0 / fs;
  p.k0 = static_cast<uint32_t>(seed);
  p.k1 = static_cast<uint32_t>(seed >> 32);
  int rc = digest_triad(gyro_err, vib_gyro, fs, &p.gyro);
  if (rc != B2INS_OK) return rc;
  rc = digest_triad(accel_err, vib_accel, fs, &p.accel);
  if (rc != B2INS_OK) return rc;
  p.ref_gyro = ref_gyro;
  p.ref_accel = ref_accel;
  p.out_gyro = gyro;
  p.out_accel = accel;
  layout_strides(layout, runs, n, &p.osr, &p.ost, &p.osc);
  p.z_dump = z_dump;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Few runs and a long series (the Allan configuration): split the time axis into segments so
  // that every SM has work.  The Gauss-Markov state at a segment start needs the draws before it:
  // pass 1 reduces every segment to its zero-state end value, a tiny serial kernel chains them,
  // pass 0 regenerates (counter-based Philox: nothing is stored) and writes.  Costs the noise
  // twice, so it is only used when one CTA per run would leave most of the GPU idle.
  int nseg = 1;
  const int64_t want_ctas = static_cast<int64_t>(sm_count()) * 2;
  if (runs < want_ctas && n >= (int64_t(1) << 18)) {
    nseg = static_cast<int>((want_ctas + runs - 1) / runs);
    const int64_t max_seg = n / (int64_t(1) << 16);
    if (nseg > max_seg) nseg = static_cast<int>(max_seg);
    if (nseg < 1) nseg = 1;
  }
  p.nseg = nseg;
  p.seg_len = n;
  p.pass = 0;
  p.seg_carry = nullptr;
  p.seg_end = nullptr;
  double* scratch = nullptr;
  if (nseg > 1) {
    int64_t len = (n + nseg - 1) / nseg;
    len = (len + kNoiseTile - 1) / kNoiseTile * kNoiseTile;   // whole tiles per segment
    p.seg_len = len;
    p.nseg = static_cast<int>((n + len - 1) / len);
    CU_CHECK(cudaMallocAsync(&scratch, sizeof(double) * runs * p.nseg * 12, s));
    p.seg_end = scratch;
    p.seg_carry = scratch + runs * p.nseg * 6;
    // pass 1 only needs the drives that still matter at the segment end: a^L < 1e-20
    int64_t keep = 1;
    for (int c = 0; c < 6; ++c) {
      const double a = (c < 3) ? p.accel.gm_a[c] : p.gyro.gm_a[c - 3];
      if (a >= 1.0) {
        keep = len;
      } else if (a > 0.0) {
        const double need = std::ceil(std::log(1e-20) / std::log(a));
        if (need > static_cast<double>(keep)) keep = need >= static_cast<double>(len) ? len : static_cast<int64_t>(need);
      }
    }
    keep = (keep + kNoiseTile - 1) / kNoiseTile * kNoiseTile;
    p.pass1_len = keep < len ? keep : len;
    p.pass = 1;
    if (p.nseg > 1)
      imu_noise_kernel<<<static_cast<unsigned>(runs * (p.nseg - 1)), kNoiseThreads, 0, s>>>(p);
    noise_carry_kernel<<<static_cast<unsigned>((runs * 6 + 127) / 128), 128, 0, s>>>(p);
    p.pass = 0;
  }
  imu_noise_kernel<<<static_cast<unsigned>(runs * p.nseg), kNoiseThreads, 0, s>>>(p);
  if (scratch) CU_CHECK(cudaFreeAsync(scratch, s));
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_gps_noise_f64(int64_t runs, int64_t m, const double* ref_gps, const double* stdp,
                        const double* stdv, int gps_type, uint64_t seed, int64_t run_offset,
                        double* gps, void* stream) {
  ARG_CHECK(runs >= 0 && m >= 0, "runs and m must be non-negative");
  ARG_CHECK(gps_type == 0 || gps_type == 1, "gps_type must be 0 (LLA) or 1 (xyz)");
  if (runs == 0 || m == 0) return B2INS_OK;
  ARG_CHECK(ref_gps && stdp && stdv && gps, "null buffer");
  ARG_CHECK(m < (int64_t(1) << 32), "m must be < 2^32");
  GpsParams p;
  p.m = m;
  p.runs = runs;
  p.run_offset = run_offset;
  p.ref = ref_gps;
  p.out = gps;
  for (int i = 0; i < 3; ++i) {
    p.stdp[i] = stdp[i];
    p.stdv[i] = stdv[i];
  }
  p.k0 = static_cast<uint32_t>(seed);
  p.k1 = static_cast<uint32_t>(seed >> 32);
  p.gps_type = gps_type;
  const int64_t total = runs * m;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  gps_noise_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_imu_noise_f64_host(double fs, int64_t runs, int64_t n, const double* ref_gyro,
                             const double* ref_accel, const b2ins_sensor_err* gyro_err,
                             const b2ins_sensor_err* accel_err, const b2ins_vib* vib_gyro,
                             const b2ins_vib* vib_accel, uint64_t seed, int64_t run_offset,
                             int layout, double* gyro, double* accel, double* z_dump) {
  ARG_CHECK(runs >= 0 && n >= 0, "runs and n must be non-negative");
  if (runs == 0 || n == 0) return B2INS_OK;
  ARG_CHECK(ref_gyro && ref_accel && gyro && accel, "null buffer");
  ARG_CHECK(!(vib_gyro && vib_gyro->type == B2INS_VIB_SERIES) &&
                !(vib_accel && vib_accel->type == B2INS_VIB_SERIES),
            "VIB_SERIES takes a device pointer: use the device entry point");
  const size_t ref_bytes = static_cast<size_t>(n) * 3 * sizeof(double);
  const size_t bytes = static_cast<size_t>(runs) * ref_bytes;
  DevBuf rg, ra, og, oa, zd;
  Stream st;
  CU_CHECK(st.create());
  CU_CHECK(rg.alloc(ref_bytes));
  CU_CHECK(ra.alloc(ref_bytes));
  CU_CHECK(og.alloc(bytes));
  CU_CHECK(oa.alloc(bytes));
  if (z_dump) CU_CHECK(zd.alloc(bytes * 4));
  CU_CHECK(cudaMemcpyAsync(rg.p, ref_gyro, ref_bytes, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(ra.p, ref_accel, ref_bytes, cudaMemcpyHostToDevice, st.s));
  const int rc = b2ins_imu_noise_f64(fs, runs, n, rg.d(), ra.d(), gyro_err, accel_err, vib_gyro,
                                     vib_accel, seed, run_offset, layout, og.d(), oa.d(),
                                     z_dump ? zd.d() : nullptr, st.s);
  if (rc != B2INS_OK) return rc;
  CU_CHECK(cudaMemcpyAsync(gyro, og.p, bytes, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaMemcpyAsync(accel, oa.p, bytes, cudaMemcpyDeviceToHost, st.s));
  if (z_dump) CU_CHECK(cudaMemcpyAsync(z_dump, zd.p, bytes * 4, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaStreamSynchronize(st.s));
  return B2INS_OK;
}

// ---------------------------------------------------------------- K12 -------
int b2ins_mc_free_integration_f64(const b2ins_mc_config* cfg, const double* ref_gyro,
                                  const double* ref_accel, const double* ref_nav,
                                  const double* ini, double* end_err, double* end_state,
                                  double* proc_stats, double* dump_att, double* dump_pos,
                                  double* dump_vel, double* dump_gyro, double* dump_accel,
                                  void* stream) {
  ARG_CHECK(cfg, "cfg is null");
  ARG_CHECK(cfg->ref_frame == 0 || cfg->ref_frame == 1, "ref_frame must be 0 or 1");
  ARG_CHECK(cfg->fs > 0.0, "fs must be positive");
  ARG_CHECK(cfg->runs >= 0 && cfg->n >= 0, "runs and n must be non-negative");
  ARG_CHECK(cfg->ini_sets >= 1 && (cfg->ini_rows == 9 || cfg->ini_rows == 10),
            "ini must be [sets>=1][9|10]");
  if (cfg->runs == 0 || cfg->n == 0) return B2INS_OK;
  ARG_CHECK(ref_gyro && ref_accel && ini, "null input buffer");
  ARG_CHECK(aligned16(ref_gyro) && aligned16(ref_accel) && aligned16(ref_nav),
            "ref_gyro / ref_accel / ref_nav must be 16-byte aligned (bulk async copies)");
  ARG_CHECK(cfg->n < (int64_t(1) << 32), "n must be < 2^32");
  ARG_CHECK(!(end_err || cfg->stats_start >= 0) || ref_nav, "ref_nav is needed for errors");
  ARG_CHECK(cfg->stats_start < 0 || proc_stats, "stats_start >= 0 needs proc_stats");
  ARG_CHECK(cfg->dump_runs >= 0 && cfg->dump_runs <= cfg->runs, "dump_runs out of range");
  ARG_CHECK((!dump_att && !dump_pos && !dump_vel) || (dump_att && dump_pos && dump_vel),
            "dump_att/pos/vel must be given together");
  ARG_CHECK((!dump_gyro) == (!dump_accel), "dump_gyro/accel must be given together");
  McParams p;
  std::memset(&p, 0, sizeof(p));
  p.n = cfg->n;
  p.runs = cfg->runs;
  p.run_offset = cfg->run_offset;
  p.ini_offset = cfg->ini_offset;
  p.dt = 1.0 / cfg->fs;
  p.earth_rot = cfg->earth_rot;
  p.k0 = static_cast<uint32_t>(cfg->seed);
  p.k1 = static_cast<uint32_t>(cfg->seed >> 32);
  int rc = digest_triad(&cfg->gyro_err, &cfg->vib_gyro, cfg->fs, &p.gyro);
  if (rc != B2INS_OK) return rc;
  rc = digest_triad(&cfg->accel_err, &cfg->vib_accel, cfg->fs, &p.accel);
  if (rc != B2INS_OK) return rc;
  p.ref_gyro = ref_gyro;
  p.ref_accel = ref_accel;
  p.ref_nav = ref_nav;
  p.ini = ini;
  p.ini_sets = cfg->ini_sets;
  p.ini_rows = cfg->ini_rows;
  p.out_att = dump_att;
  p.out_pos = dump_pos;
  p.out_vel = dump_vel;
  p.out_gyro = dump_gyro;
  p.out_accel = dump_accel;
  ARG_CHECK(cfg->dump_stride >= 0, "dump_stride must be >= 0");
  p.dump_stride = cfg->dump_stride > 1 ? cfg->dump_stride : 1;
  p.dump_rows = (cfg->n + p.dump_stride - 1) / p.dump_stride;
  layout_strides(B2INS_LAYOUT_RUN_MAJOR, cfg->dump_runs, p.dump_rows, &p.osr, &p.ost, &p.osc);
  p.dump_runs = (dump_att || dump_gyro) ? cfg->dump_runs : 0;
  ARG_CHECK(!cfg->dump_quat || dump_att, "dump_quat needs the attitude histories (dump_att)");
  p.out_quat = cfg->dump_quat;
  p.end_err = end_err;
  p.end_state = end_state;
  p.proc_stats = proc_stats;
  p.stats_start = cfg->stats_start;
  ARG_CHECK(cfg->algo == 0 || cfg->algo == 1, "algo must be 0 (free integration) or 1 (odometer)");
  p.algo = cfg->algo;
  if (cfg->algo == 1) {
    ARG_CHECK(cfg->ref_odo, "algo 1 needs ref_odo");
    p.ref_odo = cfg->ref_odo;
    p.odo_scale = cfg->odo_scale;
    p.odo_stdv = cfg->odo_stdv;
    p.out_odo = cfg->dump_odo;
    if (cfg->dump_odo && p.dump_runs == 0) p.dump_runs = cfg->dump_runs;
  }
  const int lanes = cfg->lanes_per_run ? cfg->lanes_per_run : auto_lanes(cfg->runs, cfg->stats_start < 0);
  return launch_mc(p, lanes, cfg->ref_frame, false, cfg->stats_start >= 0,
                   static_cast<cudaStream_t>(stream));
}

int b2ins_mc_free_integration_f64_host(const b2ins_mc_config* cfg, const double* ref_gyro,
                                       const double* ref_accel, const double* ref_nav,
                                       const double* ini, double* end_err, double* stats) {
  ARG_CHECK(cfg, "cfg is null");
  ARG_CHECK(cfg->runs > 0 && cfg->n > 0, "runs and n must be positive");
  ARG_CHECK(ref_gyro && ref_accel && ref_nav && ini && stats, "null buffer");
  ARG_CHECK(cfg->vib_gyro.type != B2INS_VIB_SERIES && cfg->vib_accel.type != B2INS_VIB_SERIES,
            "VIB_SERIES takes a device pointer: use the device entry point");
  ARG_CHECK(cfg->ini_sets >= 1 && (cfg->ini_rows == 9 || cfg->ini_rows == 10),
            "ini must be [sets>=1][9|10]");
  const size_t ref_bytes = static_cast<size_t>(cfg->n) * 3 * sizeof(double);
  const size_t ini_bytes = static_cast<size_t>(cfg->ini_sets) * cfg->ini_rows * sizeof(double);
  const size_t err_bytes = static_cast<size_t>(cfg->runs) * 9 * sizeof(double);
  DevBuf rg, ra, rn, di, de, ds, ws;
  Stream st;
  CU_CHECK(st.create());
  CU_CHECK(rg.alloc(ref_bytes));
  CU_CHECK(ra.alloc(ref_bytes));
  CU_CHECK(rn.alloc(ref_bytes * 3));
  CU_CHECK(di.alloc(ini_bytes));
  CU_CHECK(de.alloc(err_bytes));
  CU_CHECK(ds.alloc(27 * sizeof(double)));
  CU_CHECK(ws.alloc(static_cast<size_t>(b2ins_error_stats_workspace_bytes(9))));
  CU_CHECK(cudaMemcpyAsync(rg.p, ref_gyro, ref_bytes, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(ra.p, ref_accel, ref_bytes, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(rn.p, ref_nav, ref_bytes * 3, cudaMemcpyHostToDevice, st.s));
  CU_CHECK(cudaMemcpyAsync(di.p, ini, ini_bytes, cudaMemcpyHostToDevice, st.s));
  b2ins_mc_config c = *cfg;
  c.stats_start = -1;
  c.dump_runs = 0;
  int rc = b2ins_mc_free_integration_f64(&c, rg.d(), ra.d(), rn.d(), di.d(), de.d(), nullptr,
                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         st.s);
  if (rc != B2INS_OK) return rc;
  rc = b2ins_error_stats_f64(cfg->runs, 9, de.d(), ds.d(), ws.d(), st.s);
  if (rc != B2INS_OK) return rc;
  if (end_err) CU_CHECK(cudaMemcpyAsync(end_err, de.p, err_bytes, cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaMemcpyAsync(stats, ds.p, 27 * sizeof(double), cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaStreamSynchronize(st.s));
  return B2INS_OK;
}

// ---------------------------------------------------------------- plan ------
struct b2ins_mc_plan {
  int64_t n = 0, max_runs = 0;
  int ini_sets = 0, ini_rows = 0, device = 0;
  cudaStream_t stream = nullptr;
  double* d_in = nullptr;    // [n*3 gyro][n*3 accel][n*9 nav][sets*rows ini]
  double* d_out = nullptr;   // [27 stats][max_runs*9 end_err]
  double* d_ws = nullptr;
  double* h_in = nullptr;    // pinned: [n*3][n*3][9 last nav row][sets*rows]
  double* h_out = nullptr;   // pinned: [27][max_runs*9]
  // sub-buffers start on 16-byte boundaries (bulk async copies): offsets in doubles
  size_t n3p() const { return (static_cast<size_t>(n) * 3 + 1) & ~size_t(1); }
  size_t n9p() const { return (static_cast<size_t>(n) * 9 + 1) & ~size_t(1); }
  size_t in_doubles() const { return 2 * n3p() + n9p() + static_cast<size_t>(ini_sets) * ini_rows; }
};

int b2ins_mc_plan_destroy(b2ins_mc_plan* plan) {
  if (!plan) return B2INS_OK;
  if (plan->d_in) cudaFree(plan->d_in);
  if (plan->d_out) cudaFree(plan->d_out);
  if (plan->d_ws) cudaFree(plan->d_ws);
  if (plan->h_in) cudaFreeHost(plan->h_in);
  if (plan->h_out) cudaFreeHost(plan->h_out);
  if (plan->stream) cudaStreamDestroy(plan->stream);
  delete plan;
  return B2INS_OK;
}

int b2ins_mc_plan_create(int64_t n, int64_t max_runs, int ini_sets, int ini_rows,
                         b2ins_mc_plan** out) {
  ARG_CHECK(out, "null plan pointer");
  *out = nullptr;
  ARG_CHECK(n > 0 && max_runs > 0, "n and max_runs must be positive");
  ARG_CHECK(n < (int64_t(1) << 32), "n must be < 2^32");
  ARG_CHECK(ini_sets >= 1 && (ini_rows == 9 || ini_rows == 10), "ini must be [sets>=1][9|10]");
  b2ins_mc_plan* p = new b2ins_mc_plan();
  p->n = n;
  p->max_runs = max_runs;
  p->ini_sets = ini_sets;
  p->ini_rows = ini_rows;
  const size_t in_bytes = p->in_doubles() * sizeof(double);
  const size_t stage_bytes = (2 * p->n3p() + 10 + static_cast<size_t>(ini_sets) * ini_rows) * sizeof(double);
  const size_t out_bytes = (27 + static_cast<size_t>(max_runs) * 9) * sizeof(double);
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_in, in_bytes + 64);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_out, out_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_ws, static_cast<size_t>(b2ins_error_stats_workspace_bytes(9)));
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_in, stage_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_out, out_bytes);
  if (e != cudaSuccess) {
    b2ins_mc_plan_destroy(p);
    return fail(B2INS_ERR_CUDA, "plan allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return B2INS_OK;
}

int b2ins_mc_plan_run(b2ins_mc_plan* plan, const b2ins_mc_config* cfg, const double* ref_gyro,
                      const double* ref_accel, const double* ref_nav_end, const double* ini,
                      double* end_err, double* stats) {
  ARG_CHECK(plan && cfg, "null plan / cfg");
  ARG_CHECK(cfg->n == plan->n && cfg->runs >= 1 && cfg->runs <= plan->max_runs &&
                cfg->ini_sets == plan->ini_sets && cfg->ini_rows == plan->ini_rows,
            "cfg does not fit the plan (n=%lld runs<=%lld ini=[%d][%d])",
            static_cast<long long>(plan->n), static_cast<long long>(plan->max_runs), plan->ini_sets,
            plan->ini_rows);
  ARG_CHECK(ref_gyro && ref_accel && ref_nav_end && ini, "null buffer");
  ARG_CHECK(cfg->stats_start < 0 && cfg->dump_runs == 0,
            "a plan computes end-point errors and their statistics only");
  const int64_t n = plan->n;
  const size_t n3 = static_cast<size_t>(n) * 3;
  const size_t ini_d = static_cast<size_t>(plan->ini_sets) * plan->ini_rows;
  // stage: gyro | accel | last nav row | ini  (gyro/accel at the device offsets)
  const size_t n3p = plan->n3p();
  std::memcpy(plan->h_in, ref_gyro, n3 * sizeof(double));
  std::memcpy(plan->h_in + n3p, ref_accel, n3 * sizeof(double));
  // one slack double keeps the (virtual) base of the navigation rows 16-byte aligned
  const size_t nav_at = 2 * n3p + ((static_cast<size_t>(n - 1) * 9) & 1);
  std::memcpy(plan->h_in + nav_at, ref_nav_end, 9 * sizeof(double));
  std::memcpy(plan->h_in + nav_at + 9, ini, ini_d * sizeof(double));
  // the device block mirrors the staging block (gyro | accel | last nav row | ini): ONE copy.  The
  // kernel only reads row n-1 of the navigation rows (end-point errors), so their base pointer is
  // set n-1 rows below the staged row; the rest of the [n][9] block is never touched.
  double* d_gyro = plan->d_in;
  double* d_accel = plan->d_in + n3p;
  double* d_nav_end = plan->d_in + nav_at;
  double* d_nav = d_nav_end - (n - 1) * 9;
  double* d_ini = d_nav_end + 9;
  CU_CHECK(cudaMemcpyAsync(d_gyro, plan->h_in, (nav_at + 9 + ini_d) * sizeof(double),
                           cudaMemcpyHostToDevice, plan->stream));
  double* d_stats = plan->d_out;
  double* d_err = plan->d_out + 27;
  int rc = b2ins_mc_free_integration_f64(cfg, d_gyro, d_accel, d_nav, d_ini, d_err, nullptr,
                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         plan->stream);
  if (rc != B2INS_OK) return rc;
  if (stats) {
    rc = b2ins_error_stats_f64(cfg->runs, 9, d_err, d_stats, plan->d_ws, plan->stream);
    if (rc != B2INS_OK) return rc;
  }
  const size_t out_d = 27 + (end_err ? static_cast<size_t>(cfg->runs) * 9 : 0);
  CU_CHECK(cudaMemcpyAsync(plan->h_out, plan->d_out, out_d * sizeof(double), cudaMemcpyDeviceToHost,
                           plan->stream));
  CU_CHECK(cudaStreamSynchronize(plan->stream));
  if (stats) std::memcpy(stats, plan->h_out, 27 * sizeof(double));
  if (end_err) std::memcpy(end_err, plan->h_out + 27, static_cast<size_t>(cfg->runs) * 9 * sizeof(double));
  return B2INS_OK;
}

double* b2ins_mc_plan_err_device(b2ins_mc_plan* plan) { return plan ? plan->d_out + 27 : nullptr; }
void* b2ins_mc_plan_stream(b2ins_mc_plan* plan) { return plan ? plan->stream : nullptr; }

// ---------------------------------------------------------------- K7 --------
int b2ins_ins_loose_f64(const b2ins_ekf_config* cfg, const double* ref_gyro, const double* ref_accel,
                        const double* ref_nav, const double* ref_gps, const int64_t* gps_idx,
                        const double* gps_vis, double* end_err, double* end_bias, double* consist,
                        double* dump_att, double* dump_pos, double* dump_vel, double* dump_wb,
                        double* dump_ab, void* stream) {
  ARG_CHECK(cfg, "cfg is null");
  ARG_CHECK(cfg->fs > 0.0, "fs must be positive");
  ARG_CHECK(cfg->runs >= 0 && cfg->n >= 0 && cfg->m >= 0, "runs, n and m must be non-negative");
  if (cfg->runs == 0 || cfg->n == 0) return B2INS_OK;
  ARG_CHECK(cfg->n < (int64_t(1) << 32), "n must be < 2^32");
  ARG_CHECK(ref_gyro && ref_accel && ref_nav && end_err, "null buffer");
  ARG_CHECK(cfg->m == 0 || (ref_gps && gps_idx && gps_vis), "m > 0 needs ref_gps, gps_idx and gps_vis");
  ARG_CHECK(cfg->dump_runs >= 0 && cfg->dump_runs <= cfg->runs, "dump_runs out of range");
  const int ndump = (dump_att != nullptr) + (dump_pos != nullptr) + (dump_vel != nullptr) + (dump_wb != nullptr) +
                    (dump_ab != nullptr);
  ARG_CHECK(ndump == 0 || ndump == 5, "dump_att/pos/vel/wb/ab must be given together");
  ARG_CHECK(cfg->dump_stride >= 0, "dump_stride must be >= 0");
  EkfParams p;
  std::memset(&p, 0, sizeof(p));
  p.n = cfg->n;
  p.runs = cfg->runs;
  p.run_offset = cfg->run_offset;
  p.m = cfg->m;
  p.dt = 1.0 / cfg->fs;
  p.earth_rot = cfg->earth_rot;
  p.k0 = static_cast<uint32_t>(cfg->seed);
  p.k1 = static_cast<uint32_t>(cfg->seed >> 32);
  int rc = digest_triad(&cfg->gyro_err, nullptr, cfg->fs, &p.gyro);
  if (rc != B2INS_OK) return rc;
  rc = digest_triad(&cfg->accel_err, nullptr, cfg->fs, &p.accel);
  if (rc != B2INS_OK) return rc;
  p.ref_gyro = ref_gyro;
  p.ref_accel = ref_accel;
  p.ref_nav = ref_nav;
  p.ref_gps = ref_gps;
  p.gps_idx = gps_idx;
  p.gps_vis = gps_vis;
  for (int c = 0; c < 3; ++c) {
    p.stdp[c] = cfg->gps_stdp[c];
    p.stdv[c] = cfg->gps_stdv[c];
    p.p0[c] = cfg->gps_stdp[c] * cfg->gps_stdp[c];
    p.p0[3 + c] = cfg->gps_stdv[c] * cfg->gps_stdv[c];
    p.p0[6 + c] = cfg->ini_att_std[c] * cfg->ini_att_std[c];
    p.p0[9 + c] = cfg->gyro_err.b_drift[c] * cfg->gyro_err.b_drift[c] + cfg->gyro_err.b[c] * cfg->gyro_err.b[c];
    p.p0[12 + c] = cfg->accel_err.b_drift[c] * cfg->accel_err.b_drift[c] + cfg->accel_err.b[c] * cfg->accel_err.b[c];
    // the filter's bias model is the generator's: a = 1 - dt/tau, b^2 (white drift: a = 0, b = drift)
    const bool wg = std::isinf(cfg->gyro_err.b_corr[c]), wa = std::isinf(cfg->accel_err.b_corr[c]);
    p.ag[c] = wg ? 0.0 : p.gyro.gm_a[c];
    p.qg[c] = wg ? p.gyro.wd[c] * p.gyro.wd[c] : p.gyro.gm_b[c] * p.gyro.gm_b[c];
    p.aa[c] = wa ? 0.0 : p.accel.gm_a[c];
    p.qa[c] = wa ? p.accel.wd[c] * p.accel.wd[c] : p.accel.gm_b[c] * p.accel.gm_b[c];
    p.arw2dt[c] = cfg->gyro_err.rw[c] * cfg->gyro_err.rw[c] * p.dt;
    p.vrw2dt[c] = cfg->accel_err.rw[c] * cfg->accel_err.rw[c] * p.dt;
  }
  for (int c = 0; c < 9; ++c) p.ini[c] = cfg->ini[c];
  ARG_CHECK(cfg->vel_rw >= 0.0 && cfg->att_rw >= 0.0, "vel_rw and att_rw must be >= 0");
  p.qv_extra = cfg->vel_rw * cfg->vel_rw * p.dt;
  p.qphi_extra = cfg->att_rw * cfg->att_rw * p.dt;
  p.stats_start = cfg->stats_start;
  p.end_err = end_err;
  p.end_bias = end_bias;
  p.consist = consist;
  p.out_att = dump_att;
  p.out_pos = dump_pos;
  p.out_vel = dump_vel;
  p.out_wb = dump_wb;
  p.out_ab = dump_ab;
  p.dump_runs = ndump ? cfg->dump_runs : 0;
  p.dump_stride = cfg->dump_stride > 1 ? cfg->dump_stride : 1;
  p.dump_rows = (cfg->n + p.dump_stride - 1) / p.dump_stride;
  const unsigned grid = static_cast<unsigned>((cfg->runs + kEkfRuns - 1) / kEkfRuns);
  ekf_kernel<<<grid, kEkfThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

// ---------------------------------------------------------------- K3 --------
int64_t b2ins_error_stats_workspace_bytes(int ncomp) {
  if (ncomp < 1) return 0;
  // per-block partials + partial[2nc] + mean... : [kStatBlocks][2][nc] + 4*nc
  return static_cast<int64_t>(sizeof(double)) * (static_cast<int64_t>(kStatBlocks) * 2 + 4) * ncomp;
}

static int stage1_grid(int64_t runs, int ncomp, int threads) {
  const int64_t total = runs * ncomp;
  int64_t g = (total + threads - 1) / threads;
  if (g > kStatBlocks) g = kStatBlocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int b2ins_error_partial_f64(int64_t runs, int ncomp, const double* err, double* partial,
                            void* workspace, void* stream) {
  ARG_CHECK(runs > 0 && ncomp >= 1 && ncomp <= kStatMaxComp, "bad runs/ncomp");
  ARG_CHECK(err && partial && workspace, "null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = stat_threads(ncomp);
  const int grid = stage1_grid(runs, ncomp, threads);
  double* ws = static_cast<double*>(workspace);
  err_stage1_kernel<0><<<grid, threads, 2 * threads * sizeof(double), s>>>(runs, ncomp, err,
                                                                           nullptr, ws);
  err_stage2_kernel<0><<<1, 32, 0, s>>>(grid, ncomp, ws, partial);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_error_partial2_f64(int64_t runs, int ncomp, const double* err, const double* mean,
                             double* partial2, void* workspace, void* stream) {
  ARG_CHECK(runs > 0 && ncomp >= 1 && ncomp <= kStatMaxComp, "bad runs/ncomp");
  ARG_CHECK(err && mean && partial2 && workspace, "null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = stat_threads(ncomp);
  const int grid = stage1_grid(runs, ncomp, threads);
  double* ws = static_cast<double*>(workspace);
  err_stage1_kernel<1><<<grid, threads, 2 * threads * sizeof(double), s>>>(runs, ncomp, err, mean,
                                                                           ws);
  err_stage2_kernel<1><<<1, 32, 0, s>>>(grid, ncomp, ws, partial2);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_error_stats_f64(int64_t runs, int ncomp, const double* err, double* stats,
                          void* workspace, void* stream) {
  ARG_CHECK(runs > 0 && ncomp >= 1 && ncomp <= kStatMaxComp, "bad runs/ncomp");
  ARG_CHECK(err && stats && workspace, "null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (runs * ncomp <= kStatSmallMax) {
    stats_small_kernel<<<1, kStatSmallThreads, 0, s>>>(runs, ncomp, err, stats);
    CU_CHECK(cudaGetLastError());
    return B2INS_OK;
  }
  double* ws = static_cast<double*>(workspace);
  double* partial = ws + static_cast<int64_t>(kStatBlocks) * 2 * ncomp;  // [2nc]
  double* partial2 = partial + 2 * ncomp;                                // [nc]
  int rc = b2ins_error_partial_f64(runs, ncomp, err, partial, workspace, stream);
  if (rc != B2INS_OK) return rc;
  stats_mean_kernel<<<1, 32, 0, s>>>(runs, ncomp, partial, stats);
  rc = b2ins_error_partial2_f64(runs, ncomp, err, stats + ncomp, partial2, workspace, stream);
  if (rc != B2INS_OK) return rc;
  stats_std_kernel<<<1, 32, 0, s>>>(runs, ncomp, partial2, stats);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_error_stats_exchange_f64(int64_t runs, int ncomp, const double* err, int rank, int world,
                                   const uint64_t* windows, uint64_t seq, double* stats,
                                   int* timeout_flag, void* stream) {
  ARG_CHECK(runs >= 0 && ncomp >= 1 && 3 * ncomp + 2 <= kXchgSlot, "bad runs/ncomp");
  ARG_CHECK(runs * ncomp <= kStatSmallMax, "the fused exchange handles runs*ncomp <= 2^17 per rank");
  ARG_CHECK(world >= 1 && world <= kXchgMaxWorld && rank >= 0 && rank < world, "bad rank/world");
  ARG_CHECK(windows && stats && timeout_flag && seq >= 1, "null buffer / seq must start at 1");
  ARG_CHECK(runs == 0 || err, "null err");
  XchgParams p;
  std::memset(&p, 0, sizeof(p));
  p.runs = runs;
  p.ncomp = ncomp;
  p.rank = rank;
  p.world = world;
  p.seq = seq;
  p.err = err;
  for (int q = 0; q < world; ++q) p.peer[q] = reinterpret_cast<double*>(windows[q]);
  p.out = stats;
  p.timeout_flag = timeout_flag;
  stats_exchange_kernel<<<1, kStatSmallThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

// ---------------------------------------------------------------- K4 --------
int64_t b2ins_allan_workspace_bytes(int64_t n, int64_t nseries) {
  return allan_workspace_bytes(n, nseries);
}

int b2ins_allan_f64(double fs, int64_t n, int64_t nseries, const double* x, int64_t inner,
                    int64_t outer_stride, int64_t sample_stride, double* avar, double* tau,
                    void* workspace, void* stream) {
  ARG_CHECK(fs > 0.0 && n >= 0 && nseries >= 0, "bad fs/n/nseries");
  ARG_CHECK(inner >= 1 && sample_stride >= 1, "bad strides");
  int64_t mult[128];
  const int ntau = b2ins_allan_num_tau(n, fs, mult, 128);
  if (ntau == 0 || nseries == 0) return B2INS_OK;
  ARG_CHECK(ntau <= 128, "too many tau");
  ARG_CHECK(x && avar && tau && workspace, "null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rc = allan_launch(fs, n, nseries, x, inner, outer_stride, sample_stride, mult, ntau,
                              avar, tau, workspace, sm_count(), s);
  if (rc != 0) return fail(B2INS_ERR_CUDA, "allan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_allan_mc_f64(double fs, int64_t n, int64_t runs, const double* ref_gyro,
                       const double* ref_accel, const b2ins_sensor_err* gyro_err,
                       const b2ins_sensor_err* accel_err, uint64_t seed, int64_t run_offset,
                       double* avar, double* tau, void* workspace, void* stream) {
  ARG_CHECK(fs > 0.0 && n >= 0 && runs >= 0, "bad fs/n/runs");
  int64_t mult[128];
  const int ntau = b2ins_allan_num_tau(n, fs, mult, 128);
  if (ntau == 0 || runs == 0) return B2INS_OK;
  ARG_CHECK(ntau <= 128, "too many tau");
  ARG_CHECK(n > kAllanChunk, "the fused Allan path needs more than %d samples per series", kAllanChunk);
  ARG_CHECK(n < (int64_t(1) << 32), "n must be < 2^32");
  ARG_CHECK(ref_gyro && ref_accel && gyro_err && accel_err && avar && tau && workspace, "null buffer");
  AllanGenParams g;
  std::memset(&g, 0, sizeof(g));
  g.n = n;
  g.run_offset = run_offset;
  g.k0 = static_cast<uint32_t>(seed);
  g.k1 = static_cast<uint32_t>(seed >> 32);
  int rc = digest_triad(gyro_err, nullptr, fs, &g.gyro);
  if (rc != B2INS_OK) return rc;
  rc = digest_triad(accel_err, nullptr, fs, &g.accel);
  if (rc != B2INS_OK) return rc;
  g.ref_gyro = ref_gyro;
  g.ref_accel = ref_accel;
  rc = allan_launch(fs, n, runs * 6, nullptr, 1, n, 1, mult, ntau, avar, tau, workspace, sm_count(),
                    static_cast<cudaStream_t>(stream), &g);
  if (rc != 0) return fail(B2INS_ERR_CUDA, "allan launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

int b2ins_allan_f64_host(double fs, int64_t n, int64_t nseries, const double* x, int64_t inner,
                         int64_t outer_stride, int64_t sample_stride, double* avar, double* tau) {
  ARG_CHECK(fs > 0.0 && n >= 0 && nseries >= 0, "bad fs/n/nseries");
  ARG_CHECK(inner >= 1 && sample_stride >= 1, "bad strides");
  const int ntau = b2ins_allan_num_tau(n, fs, nullptr, 0);
  if (ntau == 0 || nseries == 0) return B2INS_OK;
  ARG_CHECK(x && avar && tau, "null buffer");
  // extent of x touched
  const int64_t outer = (nseries + inner - 1) / inner;
  const int64_t elems = (outer - 1) * outer_stride + (inner - 1) + (n - 1) * sample_stride + 1;
  DevBuf dx, dav, dtau, ws;
  Stream st;
  CU_CHECK(st.create());
  CU_CHECK(dx.alloc(static_cast<size_t>(elems) * sizeof(double)));
  CU_CHECK(dav.alloc(static_cast<size_t>(nseries) * ntau * sizeof(double)));
  CU_CHECK(dtau.alloc(static_cast<size_t>(ntau) * sizeof(double)));
  CU_CHECK(ws.alloc(static_cast<size_t>(allan_workspace_bytes(n, nseries))));
  CU_CHECK(cudaMemcpyAsync(dx.p, x, static_cast<size_t>(elems) * sizeof(double),
                           cudaMemcpyHostToDevice, st.s));
  const int rc = b2ins_allan_f64(fs, n, nseries, dx.d(), inner, outer_stride, sample_stride,
                                 dav.d(), dtau.d(), ws.p, st.s);
  if (rc != B2INS_OK) return rc;
  CU_CHECK(cudaMemcpyAsync(avar, dav.p, static_cast<size_t>(nseries) * ntau * sizeof(double),
                           cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaMemcpyAsync(tau, dtau.p, static_cast<size_t>(ntau) * sizeof(double),
                           cudaMemcpyDeviceToHost, st.s));
  CU_CHECK(cudaStreamSynchronize(st.s));
  return B2INS_OK;
}

// ---------------------------------------------------------------- K5 --------
int b2ins_psd_series_len(int64_t n) { return n > 0 ? psd_series_len(n) : 0; }

int64_t b2ins_psd_workspace_bytes(int64_t n, int64_t runs) {
  if (n <= 0 || runs <= 0) return 16;
  const int64_t L = psd_series_len(n) / 2 + 1;
  // (A, B) of every bin, then the transform of the chirp for the Bluestein lengths (<= 8192 complex)
  return runs * 3 * L * 2 * static_cast<int64_t>(sizeof(double)) + 8192 * 16 + 64;
}

int b2ins_psd_series_f64(double fs, int64_t n, int64_t runs, int sensor, int table_len,
                         const double* freq, const double* sxx3, uint64_t seed,
                         int64_t run_offset, double* series, void* workspace, void* stream) {
  ARG_CHECK(fs > 0.0 && n > 0 && runs >= 0, "bad fs/n/runs");
  ARG_CHECK(sensor == 0 || sensor == 1, "sensor must be 0 (accel) or 1 (gyro)");
  ARG_CHECK(table_len >= 2, "the PSD table needs at least two rows");
  if (runs == 0) return B2INS_OK;
  ARG_CHECK(freq && sxx3 && series && workspace, "null buffer");
  ARG_CHECK(runs * 3 <= 65535, "at most 21845 runs per call (grid.y)");
  PsdParams p;
  p.fs = fs;
  p.runs = runs;
  p.run_offset = run_offset;
  p.N = psd_series_len(n);
  p.L = p.N / 2 + 1;
  p.L0 = table_len;
  p.sensor = sensor;
  p.interp = (table_len != p.L) ? 1 : 0;  // time_series_from_psd.py:46
  p.k0 = static_cast<uint32_t>(seed);
  p.k1 = static_cast<uint32_t>(seed >> 32);
  p.freq = freq;
  p.sxx = sxx3;
  p.ab = static_cast<double*>(workspace);
  p.series = series;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 g1((p.L + kPsdThreads - 1) / kPsdThreads, static_cast<unsigned>(runs * 3));
  psd_phase_kernel<<<g1, kPsdThreads, 0, s>>>(p);
  int bluestein = 0;
  static const bool no_fft = std::getenv("B2INS_PSD_DIRECT") != nullptr;     // tools: A/B the two paths
  const int P = no_fft ? 0 : psd_fft_plan(p.N, &bluestein);
  if (P == 0) {      // lengths without an FFT path: the O(N L) cosine synthesis
    dim3 g2((p.N + kPsdThreads - 1) / kPsdThreads, static_cast<unsigned>(runs * 3));
    psd_synth_kernel<<<g2, kPsdThreads, 0, s>>>(p);
    CU_CHECK(cudaGetLastError());
    return B2INS_OK;
  }
  PsdFftParams f;
  f.nseries = runs * 3;
  f.N = p.N;
  f.L = p.L;
  f.M = p.N / 2;
  f.P = P;
  f.logP = 0;
  while ((1 << f.logP) < P) ++f.logP;
  f.bluestein = bluestein;
  f.ab = p.ab;
  // the chirp transform sits behind the (A, B) block, 16-byte aligned
  uintptr_t tail = reinterpret_cast<uintptr_t>(p.ab + runs * 3 * static_cast<int64_t>(p.L) * 2);
  tail = (tail + 15) & ~static_cast<uintptr_t>(15);
  f.bhat = reinterpret_cast<double2*>(tail);
  f.series = series;
  const size_t smem = static_cast<size_t>(P) * 16 + static_cast<size_t>(P / 2) * 16;
  static int attr_dev = -1;
  int dev = 0;
  CU_CHECK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    CU_CHECK(cudaFuncSetAttribute(psd_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24));
    CU_CHECK(cudaFuncSetAttribute(psd_chirp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24));
    attr_dev = dev;
  }
  if (bluestein) psd_chirp_kernel<<<1, kFftThreads, smem, s>>>(f);
  const int64_t grid = f.nseries < 2 * sm_count() ? f.nseries : 2 * sm_count();
  psd_fft_kernel<<<static_cast<unsigned>(grid), kFftThreads, smem, s>>>(f);
  CU_CHECK(cudaGetLastError());
  return B2INS_OK;
}

// ---------------------------------------------------------------- host ------
int64_t b2ins_path_rows(const double* motion_def, int64_t segs, double fs) {
  if (!motion_def || segs <= 0 || !(fs > 0.0)) return -1;
  return b2ins_host::path_rows(motion_def, segs, fs);
}

int64_t b2ins_path_gen_host(const double* ini, const double* motion_def, int64_t segs, double fs,
                            double osr, double fs_gps, double fs_odo, const double* mobility,
                            int ref_frame, int64_t cap, double* imu, double* nav, double* gps,
                            int64_t* gps_rows, double* odo) {
  if (!ini || !motion_def || !mobility || !imu || !nav || segs <= 0 || !(fs > 0.0) || !(osr >= 1.0) ||
      (ref_frame != 0 && ref_frame != 1) || (gps && !(fs_gps > 0.0))) {
    fail(B2INS_ERR_ARG, "bad argument to b2ins_path_gen_host");
    return -1;
  }
  return b2ins_host::path_gen(ini, motion_def, segs, fs, osr, fs_gps, fs_odo, mobility, ref_frame,
                              cap, imu, nav, gps, gps_rows, odo);
}

// ---------------------------------------------------------------- diag ------
__global__ void dfma_rate_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,
         a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int b2ins_diag_auto_lanes(int64_t runs, int fused, int sm_count_arg) {
  return auto_lanes(runs < 1 ? 1 : runs, fused != 0, sm_count_arg);
}

int b2ins_diag_mc_shape(int lanes_per_run, int ref_frame, int* shape3) {
  ARG_CHECK(shape3, "null output");
  ARG_CHECK(lanes_per_run == 1 || lanes_per_run == 2 || lanes_per_run == 4 || lanes_per_run == 8 ||
                lanes_per_run == 16 || lanes_per_run == 32, "lanes_per_run must be 1,2,4,8,16 or 32");
  McShape sh = default_shape(lanes_per_run, ref_frame);
  shape_override(&sh);   // (G = 1 launches of 2^18 runs and more take the single-warp form whatever this says)
  shape3[0] = sh.spec ? sh.P : 0;
  shape3[1] = sh.spec ? sh.WI : 0;
  shape3[2] = (sh.spec && sh.split) ? 1 : 0;
  return B2INS_OK;
}

int b2ins_diag_dfma_rate(double* dfma_per_s) {
  ARG_CHECK(dfma_per_s, "null output");
  const int blocks = sm_count() * 8, threads = 256, iters = 8000;
  DevBuf out;
  CU_CHECK(out.alloc(sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CU_CHECK(cudaEventCreate(&e0));
  CU_CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
    CU_CHECK(cudaEventRecord(e0, nullptr));
    dfma_rate_kernel<<<blocks, threads>>>(out.d(), iters);
    CU_CHECK(cudaEventRecord(e1, nullptr));
    CU_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *dfma_per_s = static_cast<double>(blocks) * threads * iters * 8.0 / (best * 1e-3);
  return B2INS_OK;
}

#ifdef B2INS_PHASE_CLOCKS
// tools only: cumulative warp-cycles in (tile wait, phase A noise, phase A incl. GM scan, phase B)
int b2ins_diag_phase_clocks(unsigned long long* out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, g_phase_clocks, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z));
  }
  return B2INS_OK;
}
#endif

}  // extern "C"
