// K4: non-overlapped Allan variance on cluster sizes m = j*10^k (j = 1..9) --
// allan.allan_var, allan/allan.py:18-59.
//
// All clusters start at sample 0, so a cluster of 10m samples is the union of 10 clusters
// of m samples.  Level k works on the decade sums S_k (S_0 = x, S_{k+1}[i] = sum of
// S_k[10i..10i+9]) and yields the nine cluster sizes j*10^k plus S_{k+1}; the series is
// read from HBM once (level 0) instead of once per tau as the reference does.
//
// One CTA owns one (series, chunk) pair; a chunk is kAllanChunk = 2*2520 level-k elements
// (2520 = lcm(1..10): every cluster size starts a cluster at each chunk start) plus a
// 9-element halo for the cluster that ends where the chunk starts.  The tile is prefix-
// summed in shared memory (after subtracting its first element, which cancels exactly in
// every difference and keeps the prefix small), so a successive-difference term is
//   sum(bin b+1) - sum(bin b) = P[(b+2)j] - 2 P[(b+1)j] + P[bj].
#pragma once
#include <cstring>

#include "common.cuh"

namespace b2ins {

constexpr int kAllanChunk = 5040;
constexpr int kAllanHalo = 9;
constexpr int kAllanThreads = 256;
constexpr int kAllanMaxLevels = 10;

struct AllanLevelParams {
  int64_t len;           // N_k: elements of this level per series
  int64_t next_len;      // N_{k+1} = N_k / 10 (0: do not produce)
  int64_t nseries;
  const double* src;     // level 0: x ; else S_k [nseries][len]
  int64_t inner, outer_stride, sample_stride;  // level 0 addressing
  int level0;
  double* next;          // S_{k+1} [nseries][next_len]
  double* partial;       // [nseries][chunks][9]
  int64_t chunks;        // chunks per series at this level
  int jmax;              // cluster multipliers 1..jmax are wanted at this level
  int64_t chunk_first;   // this launch covers chunks [chunk_first, chunk_first + chunk_count)
  int64_t chunk_count;
};

__global__ void __launch_bounds__(kAllanThreads) allan_level_kernel(const __grid_constant__ AllanLevelParams p) {
  extern __shared__ double tile[];  // [kAllanHalo + kAllanChunk + 1] prefix, tile[0] = 0
  __shared__ double red[kAllanThreads / 32][9];
  __shared__ double sh_scan[kAllanThreads];
  const int64_t series = blockIdx.x / p.chunk_count;
  const int64_t chunk = p.chunk_first + blockIdx.x % p.chunk_count;
  const int64_t c0 = chunk * kAllanChunk;                 // first element of the chunk
  const int halo = (chunk == 0) ? 0 : kAllanHalo;          // elements before c0 in the tile
  const int64_t lo = c0 - halo;
  const int cnt = static_cast<int>(min64(kAllanChunk, p.len - c0)) + halo;  // tile elems
  const double* base;
  int64_t stride;
  if (p.level0) {
    base = p.src + (series / p.inner) * p.outer_stride + (series % p.inner);
    stride = p.sample_stride;
  } else {
    base = p.src + series * p.len;
    stride = 1;
  }
  const double off = base[lo * stride];
  // P[i] = sum_{q<i} (x[lo+q] - off), i = 0..cnt ; stored at tile[i]
  constexpr int kPer = (kAllanChunk + kAllanHalo + kAllanThreads - 1) / kAllanThreads;  // 20
  // coalesced load into the tile (raw values), then a per-thread serial scan of kPer
  // consecutive elements + block scan of the thread totals
  for (int i = threadIdx.x; i < cnt; i += kAllanThreads) tile[1 + i] = base[(lo + i) * stride] - off;
  if (threadIdx.x == 0) tile[0] = 0.0;
  __syncthreads();
  const int b0 = threadIdx.x * kPer;
  double run = 0.0;
  for (int q = 0; q < kPer; ++q) {
    const int i = b0 + q;
    if (i < cnt) {
      run += tile[1 + i];
      tile[1 + i] = run;
    }
  }
  sh_scan[threadIdx.x] = run;
  __syncthreads();
  // exclusive scan of thread totals (Hillis-Steele in shared memory)
  for (int o = 1; o < kAllanThreads; o <<= 1) {
    const double v = (threadIdx.x >= o) ? sh_scan[threadIdx.x - o] : 0.0;
    __syncthreads();
    sh_scan[threadIdx.x] += v;
    __syncthreads();
  }
  const double pre = (threadIdx.x == 0) ? 0.0 : sh_scan[threadIdx.x - 1];
  for (int q = 0; q < kPer; ++q) {
    const int i = b0 + q;
    if (i < cnt) tile[1 + i] += pre;
  }
  __syncthreads();

  // successive-difference terms whose SECOND bin starts inside this chunk
  double acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.0;
#pragma unroll
  for (int j = 1; j <= 9; ++j) {
    if (j <= p.jmax) {
      const int64_t nb = p.len / j;  // bins of this cluster size in the whole series
      // second bins b2 with c0 <= b2*j < c0 + kAllanChunk, 1 <= b2 <= nb-1
      int64_t b2_lo = (c0 + j - 1) / j;
      if (b2_lo < 1) b2_lo = 1;
      int64_t b2_hi = (c0 + kAllanChunk + j - 1) / j;  // exclusive
      if (b2_hi > nb) b2_hi = nb;
      for (int64_t b2 = b2_lo + threadIdx.x; b2 < b2_hi; b2 += kAllanThreads) {
        const int e1 = static_cast<int>(b2 * j - lo);  // tile index of the bin boundary
        const double d = tile[e1 + j] - 2.0 * tile[e1] + tile[e1 - j];
        acc[j - 1] += d * d;
      }
    }
  }
  // block reduction (warp shuffles, then one shared-memory hop), fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0.0;
    for (int w = 0; w < kAllanThreads / 32; ++w) v += red[w][threadIdx.x];
    p.partial[(series * p.chunks + chunk) * 9 + threadIdx.x] = v;
  }
  // decade sums for the next level
  if (p.next_len > 0) {
    const int64_t d_lo = c0 / 10;
    for (int i = threadIdx.x; i < kAllanChunk / 10; i += kAllanThreads) {
      const int64_t di = d_lo + i;
      if (di < p.next_len) {
        const int e = halo + i * 10;
        p.next[series * p.next_len + di] = (tile[e + 10] - tile[e]) + 10.0 * off;
      }
    }
  }
}

// ---- fast path: FULL chunks ---------------------------------------------------------------
// A full chunk holds an integer number of clusters of every size (5040 = 2 lcm(1..10)), so the
// cluster sums can be built hierarchically in registers, a few samples per work item, without
// any bound checks:
//   role A (8 samples / item, from a copy padded 8 -> 9 doubles: conflict-free):  j = 1, 2, 4, 8
//   role B (18 samples / item):  j = 3, 6, 9        role C (10 / item):  j = 5 and the decade sums
//   role D (7 / item):  j = 7
// An item also rebuilds the LAST clusters of the samples just before it (at most 9 samples, the
// halo for item 0) for the difference that straddles its left edge.  ~25 instructions per sample
// against ~107 of the prefix-sum kernel above, which remains the path for the ragged last chunk.
constexpr int kAllanPad8 = kAllanChunk / 8 * 9;
constexpr int kAllanFullThreads = 512;   // two 86 KB CTAs per SM: 32 warps to hide the load latency

__device__ __forceinline__ double sq_acc(double a, double b, double acc) {
  const double d = a - b;
  return fma(d, d, acc);
}

__global__ void __launch_bounds__(kAllanFullThreads) allan_full_kernel(const __grid_constant__ AllanLevelParams p) {
  extern __shared__ double smem[];
  double* raw = smem;                                // [kAllanHalo + kAllanChunk]: raw[h + e]
  double* pad8 = smem + kAllanHalo + kAllanChunk + 1;  // [kAllanPad8]: element e at e + e/8
  __shared__ double red[kAllanFullThreads / 32][9];
  const int64_t series = blockIdx.x / p.chunk_count;
  const int64_t chunk = p.chunk_first + blockIdx.x % p.chunk_count;
  const int64_t c0 = chunk * kAllanChunk;
  const int h = (chunk == 0) ? 0 : kAllanHalo;
  const bool has_prev = chunk != 0;
  const int64_t lo = c0 - h;
  const double* base;
  int64_t stride;
  if (p.level0) {
    base = p.src + (series / p.inner) * p.outer_stride + (series % p.inner);
    stride = p.sample_stride;
  } else {
    base = p.src + series * p.len;
    stride = 1;
  }
  const double off = base[lo * stride];
  {
    // all the loads of a thread are issued before the first use: 20 independent requests in flight
    // per thread instead of a load -> store chain that exposes the DRAM latency 20 times
    constexpr int kPer = (kAllanChunk + kAllanHalo + kAllanFullThreads - 1) / kAllanFullThreads;
    double v[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * kAllanFullThreads;
      v[q] = (i < kAllanChunk + h) ? base[(lo + i) * stride] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = threadIdx.x + q * kAllanFullThreads;
      if (i < kAllanChunk + h) {
        const double w = v[q] - off;
        raw[i] = w;
        const int e = i - h;
        if (e >= 0) pad8[e + (e >> 3)] = w;
      }
    }
  }
  __syncthreads();
  const double* x = raw + h;   // x[e], e in [-h, kAllanChunk)
  double a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0, a9 = 0;
  const int jm = p.jmax;

  // ---- role A: j = 1, 2, 4, 8 ------------------------------------------------------------
  for (int it = threadIdx.x; it < kAllanChunk / 8; it += kAllanFullThreads) {
    const double* c = pad8 + 9 * it;
    const double x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3], x4 = c[4], x5 = c[5], x6 = c[6], x7 = c[7];
    const double p0 = x0 + x1, p1 = x2 + x3, p2 = x4 + x5, p3 = x6 + x7;
    const double q0 = p0 + p1, q1 = p2 + p3;
    const double r = q0 + q1;
    a1 = sq_acc(x1, x0, a1); a1 = sq_acc(x2, x1, a1); a1 = sq_acc(x3, x2, a1); a1 = sq_acc(x4, x3, a1);
    a1 = sq_acc(x5, x4, a1); a1 = sq_acc(x6, x5, a1); a1 = sq_acc(x7, x6, a1);
    a2 = sq_acc(p1, p0, a2); a2 = sq_acc(p2, p1, a2); a2 = sq_acc(p3, p2, a2);
    a4 = sq_acc(q1, q0, a4);
    if (it > 0 || has_prev) {
      double y[8];
      if (it > 0) {
        const double* d = c - 9;
#pragma unroll
        for (int q = 0; q < 8; ++q) y[q] = d[q];
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) y[q] = x[q - 8];
      }
      const double pp2 = y[4] + y[5], pp3 = y[6] + y[7];
      const double qq1 = pp2 + pp3, qq0 = (y[0] + y[1]) + (y[2] + y[3]);
      a1 = sq_acc(x0, y[7], a1);
      a2 = sq_acc(p0, pp3, a2);
      a4 = sq_acc(q0, qq1, a4);
      a8 = sq_acc(r, qq0 + qq1, a8);
    }
  }
  // ---- role B: j = 3, 6, 9 -----------------------------------------------------------------
  if (jm >= 3) {
    for (int it = threadIdx.x; it < kAllanChunk / 18; it += kAllanFullThreads) {
      const double* c = x + 18 * it;
      double t[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) t[q] = (c[3 * q] + c[3 * q + 1]) + c[3 * q + 2];
      const double s0 = t[0] + t[1], s1 = t[2] + t[3], s2 = t[4] + t[5];
      const double n0 = s0 + t[2], n1 = t[3] + s2;
#pragma unroll
      for (int q = 1; q < 6; ++q) a3 = sq_acc(t[q], t[q - 1], a3);
      a6 = sq_acc(s1, s0, a6); a6 = sq_acc(s2, s1, a6);
      a9 = sq_acc(n1, n0, a9);
      if (it > 0 || has_prev) {
        // the clusters that end at the left edge: 3, 6 and 9 samples back
        const double u2 = (c[-3] + c[-2]) + c[-1], u1 = (c[-6] + c[-5]) + c[-4], u0 = (c[-9] + c[-8]) + c[-7];
        a3 = sq_acc(t[0], u2, a3);
        a6 = sq_acc(s0, u1 + u2, a6);
        a9 = sq_acc(n0, (u0 + u1) + u2, a9);
      }
    }
  }
  // ---- role C: j = 5 and the decade sums ------------------------------------------------------
  {
    const int64_t d_lo = c0 / 10;
    for (int it = threadIdx.x; it < kAllanChunk / 10; it += kAllanFullThreads) {
      const double* c = x + 10 * it;
      const double f0 = ((c[0] + c[1]) + (c[2] + c[3])) + c[4];
      const double f1 = ((c[5] + c[6]) + (c[7] + c[8])) + c[9];
      a5 = sq_acc(f1, f0, a5);
      if (it > 0 || has_prev) {
        const double g1 = ((c[-5] + c[-4]) + (c[-3] + c[-2])) + c[-1];
        a5 = sq_acc(f0, g1, a5);
      }
      if (p.next_len > 0) p.next[series * p.next_len + d_lo + it] = (f0 + f1) + 10.0 * off;
    }
  }
  // ---- role D: j = 7 -----------------------------------------------------------------------------
  if (jm >= 7) {
    for (int it = threadIdx.x; it < kAllanChunk / 7; it += kAllanFullThreads) {
      const double* c = x + 7 * it;
      const double g = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + c[6]);
      if (it > 0 || has_prev) {
        const double gp = ((c[-7] + c[-6]) + (c[-5] + c[-4])) + ((c[-3] + c[-2]) + c[-1]);
        a7 = sq_acc(g, gp, a7);
      }
    }
  }
  // ---- block reduction, fixed order -------------------------------------------------------------
  const double acc[9] = {a1, a2, a3, a4, a5, a6, a7, a8, a9};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0.0;
    for (int w = 0; w < kAllanFullThreads / 32; ++w) v += red[w][threadIdx.x];
    p.partial[(series * p.chunks + chunk) * 9 + threadIdx.x] = (threadIdx.x < jm) ? v : 0.0;
  }
}

struct AllanFinalParams {
  int64_t nseries;
  int ntau;
  double ts;
  double* avar;  // [nseries][ntau]
  double* tau;   // [ntau]
  const double* partial[kAllanMaxLevels];
  int64_t chunks[kAllanMaxLevels];
  int64_t m[128];       // cluster sizes
  int64_t nbins[128];   // floor(n / m)
  int level_of[128];
  int j_of[128];
};

__global__ void allan_final_kernel(const __grid_constant__ AllanFinalParams p) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= p.nseries * p.ntau) return;
  const int64_t series = idx / p.ntau;
  const int i = static_cast<int>(idx % p.ntau);
  const int k = p.level_of[i], j = p.j_of[i];
  const double* part = p.partial[k] + series * p.chunks[k] * 9 + (j - 1);
  double s = 0.0;
  for (int64_t c = 0; c < p.chunks[k]; ++c) s += part[c * 9];
  const double m = static_cast<double>(p.m[i]);
  // avar = 0.5/(nbins-1) * sum (mean[b+1]-mean[b])^2, allan.py:54-57
  p.avar[series * p.ntau + i] = 0.5 / static_cast<double>(p.nbins[i] - 1) * (s / (m * m));
  if (series == 0) p.tau[i] = m * p.ts;  // allan.py:58
}

inline int64_t allan_workspace_bytes(int64_t n, int64_t nseries) {
  if (n <= 0 || nseries <= 0) return 16;
  int64_t doubles = 0;
  const int64_t n1 = n / 10 + 1;
  doubles += 2 * n1 * nseries;  // ping-pong decade sums
  int64_t len = n;
  for (int k = 0; k < kAllanMaxLevels && len > 0; ++k) {
    doubles += ((len + kAllanChunk - 1) / kAllanChunk) * 9 * nseries;
    len /= 10;
  }
  return doubles * static_cast<int64_t>(sizeof(double)) + 256;
}

// returns 0 on success
inline int allan_launch(double fs, int64_t n, int64_t nseries, const double* x, int64_t inner,
                        int64_t outer_stride, int64_t sample_stride, const int64_t* mult, int ntau,
                        double* avar, double* tau, void* workspace, cudaStream_t s) {
  AllanFinalParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.nseries = nseries;
  fp.ntau = ntau;
  fp.ts = 1.0 / fs;
  fp.avar = avar;
  fp.tau = tau;
  int levels = 0;
  int jmax[kAllanMaxLevels] = {0};
  {
    int64_t scale = 1;
    int i = 0;
    for (int k = 0; k < kAllanMaxLevels && i < ntau; ++k, scale *= 10) {
      while (i < ntau && mult[i] / scale >= 1 && mult[i] / scale <= 9 && mult[i] % scale == 0) {
        fp.m[i] = mult[i];
        fp.nbins[i] = n / mult[i];
        fp.level_of[i] = k;
        fp.j_of[i] = static_cast<int>(mult[i] / scale);
        jmax[k] = fp.j_of[i];
        ++i;
      }
      levels = k + 1;
    }
    if (i != ntau) return 1;
  }
  double* ws = static_cast<double*>(workspace);
  const int64_t n1 = n / 10 + 1;
  double* buf[2] = {ws, ws + n1 * nseries};
  double* part = ws + 2 * n1 * nseries;
  int64_t len = n;
  const size_t smem = (kAllanChunk + kAllanHalo + 1 + 16) * sizeof(double);
  const size_t smem_full = (kAllanChunk + kAllanHalo + 1 + kAllanPad8 + 16) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(allan_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return 2;
    if (cudaFuncSetAttribute(allan_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_full)) != cudaSuccess)
      return 2;
    attr_set = true;
  }
  for (int k = 0; k < levels; ++k) {
    AllanLevelParams lp;
    lp.len = len;
    lp.next_len = (k + 1 < levels) ? len / 10 : 0;
    lp.nseries = nseries;
    lp.level0 = (k == 0);
    lp.src = (k == 0) ? x : buf[(k - 1) & 1];
    lp.inner = inner;
    lp.outer_stride = outer_stride;
    lp.sample_stride = sample_stride;
    lp.next = buf[k & 1];
    lp.chunks = (len + kAllanChunk - 1) / kAllanChunk;
    lp.partial = part;
    lp.jmax = jmax[k];
    fp.partial[k] = part;
    fp.chunks[k] = lp.chunks;
    part += lp.chunks * 9 * nseries;
    if (lp.chunks * nseries >= (int64_t(1) << 31)) return 4;
    // full chunks (every cluster complete, next-level decades complete) take the fast kernel
    const int64_t full = len / kAllanChunk;
    if (full > 0) {
      lp.chunk_first = 0;
      lp.chunk_count = full;
      allan_full_kernel<<<static_cast<unsigned>(full * nseries), kAllanFullThreads, smem_full, s>>>(lp);
    }
    if (lp.chunks > full) {   // the ragged last chunk
      lp.chunk_first = full;
      lp.chunk_count = lp.chunks - full;
      allan_level_kernel<<<static_cast<unsigned>(lp.chunk_count * nseries), kAllanThreads, smem, s>>>(lp);
    }
    len /= 10;
  }
  const int64_t total = nseries * ntau;
  allan_final_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, s>>>(fp);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

}  // namespace b2ins
