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

// Four SiLUs with ONE reciprocal (MUFU budget 5 instead of 8): 1/d_i from r = 1/(d0 d1 d2 d3).  The exponent is clamped
// at 20 so the product stays below 6e34; for u < -20 the result is u * 2e-9 instead of ~0 (|error| < 1e-7 |u| / 50).
__device__ __forceinline__ void silu4_shared_rcp(float& u0, float& u1, float& u2, float& u3) {
  const float L = -1.4426950408889634f, C = 28.853900817779268f;  // -log2(e); 20 log2(e)
  const float d0 = 1.0f + fast_ex2(fminf(L * u0, C)), d1 = 1.0f + fast_ex2(fminf(L * u1, C));
  const float d2 = 1.0f + fast_ex2(fminf(L * u2, C)), d3 = 1.0f + fast_ex2(fminf(L * u3, C));
  const float d01 = d0 * d1, d23 = d2 * d3;
  const float r = fast_rcp(d01 * d23);
  const float r01 = r * d23, r23 = r * d01;
  u0 *= r01 * d1; u1 *= r01 * d0; u2 *= r23 * d3; u3 *= r23 * d2;
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issued instruction).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 ld2(const float2& w) { return *reinterpret_cast<const f32x2*>(&w); }
// silu4_shared_rcp on pairs: the same formulas element-wise (ex2 / rcp / min stay scalar: MUFU and FMNMX have no pair form)
__device__ __forceinline__ void silu4_shared_rcp2(f32x2& u0, f32x2& u1, f32x2& u2, f32x2& u3) {
  const f32x2 L = pk2(-1.4426950408889634f, -1.4426950408889634f), one = pk2(1.0f, 1.0f);
  const float C = 28.853900817779268f;
  auto den = [&](f32x2 u) {
    float a, b;
    upk2(mul2(L, u), a, b);
    return add2(one, pk2(fast_ex2(fminf(a, C)), fast_ex2(fminf(b, C))));
  };
  const f32x2 d0 = den(u0), d1 = den(u1), d2 = den(u2), d3 = den(u3);
  const f32x2 d01 = mul2(d0, d1), d23 = mul2(d2, d3);
  float pa, pb;
  upk2(mul2(d01, d23), pa, pb);
  const f32x2 r = pk2(fast_rcp(pa), fast_rcp(pb));
  const f32x2 r01 = mul2(r, d23), r23 = mul2(r, d01);
  u0 = mul2(u0, mul2(r01, d1)); u1 = mul2(u1, mul2(r01, d0));
  u2 = mul2(u2, mul2(r23, d3)); u3 = mul2(u3, mul2(r23, d2));
}

// SiLU of a packed pair with its own reciprocal: 3.5 issue slots per value (the shared-reciprocal form needs ~5 because of its
// extra multiplies and clamps) at 2 MUFU operations per value — for code that is issue-bound, not MUFU-bound (GEMM epilogues).
__device__ __forceinline__ f32x2 silu2(f32x2 u) {
  float a, b;
  upk2(mul2(pk2(-1.4426950408889634f, -1.4426950408889634f), u), a, b);
  const f32x2 d = add2(pk2(1.0f, 1.0f), pk2(fast_ex2(a), fast_ex2(b)));
  float da, db;
  upk2(d, da, db);
  return mul2(u, pk2(fast_rcp(da), fast_rcp(db)));
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
