// Host build of the strapdown step the kernels run (csrc/mech.cuh with B2INS_HOST_TEST): one run of
// free integration on supplied gyro/accel, for the CPU-side check against the oracle
// (tests/test_cpu_step.py).  Test tooling; not part of libb2ins.so.
//   nvcc -O2 -std=c++17 -shared -Xcompiler -fPIC -DB2INS_HOST_TEST -o tools/libstep_host.so tools/step_host.cu
#include "../gnss_ins_sim_b200/csrc/mech.cuh"

using namespace b2ins;

template <int RF>
static void run(int64_t n, double dt, int earth_rot, int odo, const double* gyro, const double* accel,
                const double* ini, int ini_rows, int resync_every, double* att, double* pos, double* vel) {
  NavState st;
  nav_init<RF>(st, ini, ini_rows, dt);
  for (int64_t i = 0; i < n; ++i) {
    att[i * 3 + 0] = i ? wrap_once(st.yaw) : st.yaw;
    att[i * 3 + 1] = st.pitch;
    att[i * 3 + 2] = i ? wrap_once(st.roll) : st.roll;
    pos[i * 3 + 0] = st.pos.x; pos[i * 3 + 1] = st.pos.y; pos[i * 3 + 2] = st.pos.z;
    vel[i * 3 + 0] = st.vel.x; vel[i * 3 + 1] = st.vel.y; vel[i * 3 + 2] = st.vel.z;
    if (i + 1 == n) break;
    const Vec3 w{gyro[i * 3], gyro[i * 3 + 1], gyro[i * 3 + 2]};
    const Vec3 f{accel[i * 3], accel[i * 3 + 1], accel[i * 3 + 2]};
    const bool resync = resync_every > 0 ? ((i + 1) % resync_every) == 0 : false;
    if (odo)
      nav_step<RF, false, 1>(st, w, f, dt, earth_rot != 0, 0, resync);
    else
      nav_step<RF, false, 0>(st, w, f, dt, earth_rot != 0, 0, resync);
  }
}

extern "C" int step_host_free_integration(int ref_frame, int64_t n, double fs, int earth_rot, int odo,
                                          const double* gyro, const double* accel, const double* ini,
                                          int ini_rows, int resync_every, double* att, double* pos,
                                          double* vel) {
  if (ref_frame == 1)
    run<1>(n, 1.0 / fs, earth_rot, odo, gyro, accel, ini, ini_rows, resync_every, att, pos, vel);
  else
    run<0>(n, 1.0 / fs, earth_rot, odo, gyro, accel, ini, ini_rows, resync_every, att, pos, vel);
  return 0;
}
extern "C" int step_host_resync_default(void) { return kResync; }
