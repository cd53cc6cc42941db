// Shared device helpers for the OA-ReactDiff B200 hot path (sm_100a).
#pragma once
#include <cstdlib>
#include <cstring>
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>

#define OARD_EPS 1e-6f  // reference model/leftnet.py:15
#define OARD_PI 3.14159265358979323846

namespace oard {

// cudaFuncSetAttribute acts on the CURRENT device: a process that drives several GPUs (one handle per device) must opt in
// to the large dynamic shared memory once per (kernel instantiation, device), not once per process.
struct PerDeviceOnce {
  unsigned long long done = 0;
  bool first_time() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// MUFU wrappers with flush-to-zero: the non-ftz intrinsics (__expf, __fdividef) wrap every MUFU in range checks and
// rescaling multiplies (3-4 extra instructions each), which dominated the activation-heavy kernels (ncu: r1o).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ex2.approx / rcp.approx based SiLU (relative error ~1e-6): used where thousands of activations per thread dominate.
// x -> -inf gives x * 0 = -0; x -> +inf gives x * 1.
__device__ __forceinline__ float silu_fast(float x) {
  return x * fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole block; every thread gets the total.  `sm` needs >= 33 floats.  Deterministic.
__device__ __forceinline__ float block_sum(float v, float* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect sm reuse across consecutive calls
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sm[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}

// LayerNorm statistics of one row held one-element-per-thread (threads >= H idle with v = 0, valid=false).
__device__ __forceinline__ void block_ln_stats(float v, bool valid, int H, float* sm, float& mean, float& rstd) {
  mean = block_sum(valid ? v : 0.f, sm) / (float)H;
  const float d = valid ? v - mean : 0.f;
  const float var = block_sum(d * d, sm) / (float)H;  // biased, like torch.nn.LayerNorm
  rstd = rsqrtf(var + 1e-5f);
}

}  // namespace oard
