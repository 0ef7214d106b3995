// Internal interface between the translation units of libb2ins.so (not part of the C ABI).
// The fused Monte-Carlo kernels are compiled in four units (single-warp and warp-specialised form,
// one per reference frame) next to b2ins_api.cu, so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "mc_kernel.cuh"

namespace b2ins {

// launch shape of the warp-specialised form (mc_spec_kernel.cuh)
struct McShape {
  int G, P, WI;
  bool split, spec;
};

// mc_kernel<G, RF, FED, PROC>: the single-warp form (supplied data, process statistics, odometer)
void launch_mc_plain_rf0(const McParams& p, int lanes, bool fed, bool proc, cudaStream_t s);
void launch_mc_plain_rf1(const McParams& p, int lanes, bool fed, bool proc, cudaStream_t s);
// mc_spec_kernel<G, RF, P, WI, SPLIT>; false: the shape is not instantiated
bool launch_mc_spec_rf0(const McParams& p, const McShape& sh, cudaStream_t s);
bool launch_mc_spec_rf1(const McParams& p, const McShape& sh, cudaStream_t s);

}  // namespace b2ins
