"""bench.py -- Monte-Carlo free-integration throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, NCCL)

Workload = BASELINE.json configs[1]: free_integration on motion_def-90deg_turn.csv
(true trajectory from the reference's path_gen, frozen in tests/golden/traj_*.npz: n = 1000
samples at 100 Hz), 'mid-accuracy' IMU, ref_frame = 1, 1000 Monte-Carlo runs per GPU
(weak scaling: N GPUs -> 1000 N runs, global run ids rank-independent).

A "step" is one pass of the hot path over that batch: on-device IMU error generation
(Philox) -> strapdown free integration -> per-run end-point errors (K12), then the ensemble
statistics (K3; for N > 1 two all-reduces over NCCL).  `value` = runs x samples / device time
with inputs resident in HBM; `e2e` = the same through the public API (Sim.run + error
statistics) with HOST buffers, H2D of the trajectory and D2H of the statistics inside the
timed region.  L2 is flushed between timed steps.  See DESIGN.md section 7.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None


def emit(obj):
    line = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)

METRIC = 'MC-runs x timesteps/sec free_integration @100Hz'
UNIT = 'run-steps/s'
RUNS_PER_GPU = 1000
FS = 100.0
SEED = 12345
TRAJ = os.path.join(ROOT, 'tests', 'golden', 'traj_90deg_turn_100hz_rf1.npz')
WORKLOAD = ("free_integration, motion_def-90deg_turn.csv (n=1000 @100Hz), 'mid-accuracy' IMU, "
            "ref_frame=1, 1000 MC runs per GPU")
# FP64 thread-instructions (DFMA/DMUL/DADD/DSETP) one run-step costs in mc_kernel, by lane-group
# width, from the ncu source-level counts in profiles/ncu_mc_kernel_r01_v5_*.json (lanes of a
# group replicate the serial step, so wide groups spend more instructions per run-step)
FP64_INST_PER_RUN_STEP = {16: 1974.4, 1: 634.3}
# dram__bytes_read.sum + dram__bytes_write.sum of one mc_kernel launch at this workload
# (profiles/ncu_mc_kernel_r01_v8_cfg2_lanes16_spec.json): the trajectory; the 72 KB of results stay in L2
NCU_DRAM_BYTES_PER_LAUNCH = 135424


def load_workload():
    g = dict(np.load(TRAJ))
    nav = np.concatenate([g['ref_att'], g['ref_pos'], g['ref_vel']], axis=1)
    return g, np.ascontiguousarray(nav)


def host_info():
    model = ''
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    model = line.split(':', 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'samples': len(sm),
                'reasons': sorted(reasons)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its C port,
    oracle/oracle.c -- the Python original cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle_c
    from gnss_ins_sim_b200 import imu_model
    g, nav = load_workload()
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    n = nav.shape[0]
    runs = RUNS_PER_GPU * args.gpus
    ini = g['ini'][None]

    def step(k):
        return oracle_c.mc_free_integration(1, FS, runs, 0, g['ref_gyro'], g['ref_accel'], nav[-1],
                                            imu.gyro_err, imu.accel_err, SEED + k, ini, threads=0)
    used = 1
    for k in range(args.warmup):
        _, used = step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        err, used = step(args.warmup + k)
    dt = time.perf_counter() - t0
    value = runs * n * args.steps / dt
    model, ncpu = host_info()
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'runs': runs, 'samples': n, 'global_run_steps': runs * n},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port',
                         'sample': 'the full workload (%d runs x %d samples) per step, C port of '
                                   'the reference path, %d threads; host: %s (%s logical cpus)'
                                   % (runs, n, used, model, ncpu)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def cpu_baseline_sample(g, nav, imu, budget_s=4.0):
    """Rank 0, N = 1: the C port timed on the host cores on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle_c
    n = nav.shape[0]
    ini = g['ini'][None]

    def go(runs, threads):
        t0 = time.perf_counter()
        _, used = oracle_c.mc_free_integration(1, FS, runs, 0, g['ref_gyro'], g['ref_accel'],
                                               nav[-1], imu.gyro_err, imu.accel_err, SEED, ini,
                                               threads=threads)
        return time.perf_counter() - t0, used
    go(64, 0)
    t_probe, used = go(512, 0)
    rate = 512 * n / t_probe
    runs = int(max(512, min(2_000_000, rate * budget_s / n)))
    t, used = go(runs, 0)
    t1, _ = go(256, 1)
    model, ncpu = host_info()
    value, per_core = runs * n / t, 256 * n / t1
    # the GPU box's host is shared between the boxes of the pod: the threads rarely get a core each
    return {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port',
            'per_core_value': per_core, 'effective_cores': round(value / per_core, 1),
            'sample': '%d runs x %d samples of the same workload (%.1f s), C port of the reference '
                      'path (oracle/oracle.c), %d threads; host: %s (%s logical cpus); the Python '
                      'reference itself ran 5.9e4 run-steps/s/core in the survey container'
                      % (runs, n, t, used, model, ncpu)}


def run_b200(args):
    import torch
    import torch.distributed as td
    from gnss_ins_sim_b200 import engine, imu_model, dist, _lib
    from gnss_ins_sim_b200.sim import Sim
    from gnss_ins_sim_b200.free_integration import FreeIntegration

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d' % args.gpus

    g, nav_h = load_workload()
    imu = imu_model.IMU(accuracy='mid-accuracy', axis=6, gps=False)
    n = nav_h.shape[0]
    R = RUNS_PER_GPU
    total_runs = R * world
    ref_gyro, ref_accel = engine.to_device(g['ref_gyro']), engine.to_device(g['ref_accel'])
    nav, ini = engine.to_device(nav_h), engine.to_device(g['ini'][None])
    cfg = engine.make_mc_config(1, FS, n, R, SEED, imu.gyro_err, imu.accel_err, 1, 9,
                                run_offset=rank * R, ini_offset=rank * R, lanes_per_run=args.lanes)
    res = engine.McResult()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    merger = p2p = None
    exchange = 'none'
    if world > 1:
        try:        # K3x: statistics + exchange + merge fused in one kernel over NVLink peer memory
            p2p = dist.P2PStats(9)
            exchange = 'fused peer-memory kernel (K3x)'
        except Exception as e:     # no symmetric memory on this box: NCCL all_gather of 28 doubles
            sys.stderr.write('P2PStats unavailable (%s); using NCCL all_gather\n' % e)
            merger = dist.StatsMerger(9)
            exchange = 'NCCL all_gather + host merge'

    def step():
        engine.mc_free_integration(cfg, ref_gyro, ref_accel, nav, ini, out=res)
        if world == 1:
            return engine.error_stats(res.end_err)
        if p2p is not None:
            return p2p(res.end_err, R).cpu().numpy()
        return merger(engine.error_stats(res.end_err), R)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        stats = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stats = step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = total_runs * n * args.steps / (dev_ms * 1e-3)
    stats = stats.cpu().numpy() if hasattr(stats, 'cpu') else stats

    # ---- dominant kernel alone: mc_kernel launch duration -> roofline ----------------
    kev = []
    for _ in range(max(args.steps, 5)):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.mc_free_integration(cfg, ref_gyro, ref_accel, nav, ini, out=res)
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    # algorithmic HBM bytes of one launch: the shared trajectory once (n x 48 B), the last
    # ref_nav row, the initial state, and 72 B of end-point error per run
    alg_bytes = n * 48 + 72 + 72 + R * 72
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    dfma = ctypes.c_double(0.0)
    _lib.check(_lib.load().b2ins_diag_dfma_rate(ctypes.byref(dfma)))
    k_rate = R * n / (k_ms * 1e-3)
    fp64 = {'bound': 'fp64-issue', 'peak_dfma_per_s': dfma.value, 'peak_source': 'measured live '
            '(b2ins_diag_dfma_rate)', 'kernel_run_steps_per_s': k_rate,
            'dfma_slots_per_run_step': dfma.value / k_rate}
    lanes_used = args.lanes or 16      # auto_lanes(1000) = 16 on a 148-SM part
    if lanes_used in FP64_INST_PER_RUN_STEP:
        fp64['lanes_per_run'] = lanes_used
        fp64['fp64_inst_per_run_step'] = FP64_INST_PER_RUN_STEP[lanes_used]
        fp64['frac'] = FP64_INST_PER_RUN_STEP[lanes_used] * k_rate / dfma.value
        fp64['frac_note'] = ('FP64 instructions issued / measured FP64-FMA issue rate; 1000 runs leave '
                             'the serial recurrence latency-bound (one warp per SM sub-partition); the '
                             'same kernel reaches 0.59 at 10^6 runs (profiles/)')

    if args.quick:
        if rank == 0:
            emit({'metric': METRIC, 'value': value, 'unit': UNIT, 'quick': True,
                  'ms_per_step': dev_ms / args.steps, 'kernel_ms': k_ms})
        return
    # ---- e2e: public API, host buffers, copies inside the timed region -----------------
    # the step's inputs live in PINNED host memory (numpy views of pinned tensors)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(g[k])).pin_memory()
              for k in ('time', 'ref_pos', 'ref_vel', 'ref_att', 'ref_accel', 'ref_gyro')}
    traj_host = {k: v.numpy() for k, v in pinned.items()}
    lo = rank * R

    def e2e_step():
        algo = FreeIntegration(g['ini'], lanes_per_run=args.lanes)
        sim = Sim([FS, 0.0, 0.0], traj_host, ref_frame=1, imu=imu, algorithm=algo, seed=SEED,
                  lanes_per_run=args.lanes)
        sim.run(total_runs)
        return sim.get_error_stats('pos', err_stats_start=-1)
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_value = total_runs * n * args.steps / float(t.item())
    # plan path (N = 1): true IMU samples + last ref_nav row + initial state up,
    # statistics + per-run end-point errors down
    h2d = (n * 6 + 9 + 9) * 8
    d2h = (27 + R * 9) * 8

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'runs': total_runs, 'runs_per_gpu': R, 'samples': n,
                   'global_run_steps': total_runs * n, 'l2_flush_between_steps': True,
                   'lanes_per_run': args.lanes or 'auto', 'seed': SEED,
                   'parallelism': 'runs sharded x%d; statistics exchange: %s' % (world, exchange)},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'api': 'gnss_ins_sim_b200.sim.Sim.run + get_error_stats'},
        # per step: mc_kernel + stats_small_kernel (the NCCL all_gather is not ours)
        'gpu_launches': args.steps * 2,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': achieved / hbm_peak, 'traffic': NCU_DRAM_BYTES_PER_LAUNCH,
                     'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                     'kernel': 'mc_kernel (K12)', 'kernel_ms': k_ms,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'note': 'K12 reads the shared trajectory once and writes 72 B per run: it is '
                             'FP64-issue-bound, not HBM-bound; see roofline_fp64'},
        'roofline_fp64': fp64,
        'accuracy': {'end_point_rmse': {
            'att_rad': np.sqrt(stats[1, 0:3] ** 2 + stats[2, 0:3] ** 2).tolist(),
            'pos_m': np.sqrt(stats[1, 3:6] ** 2 + stats[2, 3:6] ** 2).tolist(),
            'vel_mps': np.sqrt(stats[1, 6:9] ** 2 + stats[2, 6:9] ** 2).tolist()},
            'parity': 'tests/test_gpu_parity.py: <= 1e-6 rel vs the reference on identical draws'},
    }
    if rank == 0 and world == 1:
        out['cpu_baseline'] = cpu_baseline_sample(g, nav_h, imu)
    if rank == 0:
        emit(out)
    if world > 1:
        td.destroy_process_group()


def main():
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON
    # line on stdout, so everything else is sent to stderr and the line is written to the saved fd.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--lanes', type=int, default=0, help='lanes per run (0 = auto)')
    ap.add_argument('--quick', action='store_true',
                    help='device-timed part only (for runs under a profiler): no e2e, no cpu baseline')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
