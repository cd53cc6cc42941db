// Exact-fp32 SIMT GEMM with fused gather prologue and fused epilogue.
//   C[m, n] = epi( sum_k A[arow(m), k] * W[n, k] )        (both operands K-contiguous: x @ W^T like nn.Linear)
// Used for every dense contraction of the LEFTNet forward in the exact-fp32 path (node-level GEMMs always;
// edge-level GEMMs unless the tcgen05 path is enabled).  Tile 128x64x16, 256 threads, 8x4 outputs per thread.
#pragma once
#include "common.cuh"

namespace oard {

struct GemmArgs {
  // operands
  const float* A; int lda; const int* aidx;  // aidx: optional row gather (A row = aidx[m])
  const float* W; int ldw;
  float* C; int ldc;
  int M, N, K;
  const int* m_dev;  // optional: actual M lives in device memory (dynamic active-edge count); M is then the cap
  // epilogue, applied in this order
  const float* prescale;                             // * prescale[m]  (row scale of the raw product, before bias)
  const float* bias;                                 // + bias[n]
  const float* radd1; const int* ridx1; int ld1;     // + radd1[ridx1[m], n]
  const float* radd2; const int* ridx2; int ld2;     // + radd2[ridx2[m], n]
  int act;                                           // 1: SiLU
  const float* rowscale; const int* rsidx;           // * rowscale[rsidx ? rsidx[m] : m]
  const float* mul; int ldmul;                       // * mul[m, n]
  const float* resid; int ldres;                     // + resid[m, n]   (may alias C)
  // optional second, row-scattered copy of the result: C2[c2idx[m], n] = C[m, n] for rows with c2idx[m] >= 0
  float* C2; const int* c2idx; int ldc2;
  // L2 eviction-priority hints of the tensor-core kernels' TMA streams (gemm_tc.cuh l2_policy): A operand, aux (mul / resid)
  // blocks, C stores.  0 none, 1 evict_first, 2 evict_last.
  int hintA, hintX, hintC;
  // bring-up only (pair16 kernel, tools/pair_probe.py): bit 0 no A loads, bit 1 no epilogue global traffic, bit 2 no weight
  // loads, bit 3 no MMAs.  Results are meaningless with any bit set.
  int ablate;
  long long* ts;  // bring-up only: clock64 marks of CTA 0 around its 4th tile (pair16 kernel, tools/pair_probe.py timeline)
};

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(GTHREADS) gemm_simt_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Ws[GBK][GBN + 4];
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  if (m0 >= M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  // global->smem load assignment: A: 2 x (row, 4 k) per thread; W: 1 x (row, 4 k)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const float* arow[2];
  bool aok[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int m = m0 + lr + i * 64;
    aok[i] = m < M;
    const int src = aok[i] ? (g.aidx ? g.aidx[m] : m) : 0;
    arow[i] = g.A + (size_t)src * g.lda;
  }
  const bool wok = (n0 + lr) < g.N;
  const float* wrow = g.W + (size_t)(wok ? n0 + lr : 0) * g.ldw;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  float4 ra[2], rw;
  auto gload = [&](int k0) {
    const int k = k0 + lk;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (aok[i]) {
        if (VEC) {
          if (k < g.K) ra[i] = *reinterpret_cast<const float4*>(arow[i] + k);
        } else {
          if (k + 0 < g.K) ra[i].x = arow[i][k + 0];
          if (k + 1 < g.K) ra[i].y = arow[i][k + 1];
          if (k + 2 < g.K) ra[i].z = arow[i][k + 2];
          if (k + 3 < g.K) ra[i].w = arow[i][k + 3];
        }
      }
    }
    rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wok) {
      if (VEC) {
        if (k < g.K) rw = *reinterpret_cast<const float4*>(wrow + k);
      } else {
        if (k + 0 < g.K) rw.x = wrow[k + 0];
        if (k + 1 < g.K) rw.y = wrow[k + 1];
        if (k + 2 < g.K) rw.z = wrow[k + 2];
        if (k + 3 < g.K) rw.w = wrow[k + 3];
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      As[lk + 0][lr + i * 64] = ra[i].x;
      As[lk + 1][lr + i * 64] = ra[i].y;
      As[lk + 2][lr + i * 64] = ra[i].z;
      As[lk + 3][lr + i * 64] = ra[i].w;
    }
    Ws[lk + 0][lr] = rw.x;
    Ws[lk + 1][lr] = rw.y;
    Ws[lk + 2][lr] = rw.z;
    Ws[lk + 3][lr] = rw.w;
  };

  gload(0);
  for (int k0 = 0; k0 < g.K; k0 += GBK) {
    sstore();
    __syncthreads();
    if (k0 + GBK < g.K) gload(k0 + GBK);  // overlap next global load with this tile's math
#pragma unroll
    for (int k = 0; k < GBK; k++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* r1 = g.radd1 ? g.radd1 + (size_t)(g.ridx1 ? g.ridx1[m] : m) * g.ld1 : nullptr;
    const float* r2 = g.radd2 ? g.radd2 + (size_t)(g.ridx2 ? g.ridx2[m] : m) * g.ld2 : nullptr;
    const float rs = g.rowscale ? g.rowscale[g.rsidx ? g.rsidx[m] : m] : 1.f;
    const float ps = g.prescale ? g.prescale[m] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * ps;
      if (g.bias) v += g.bias[n];
      if (r1) v += r1[n];
      if (r2) v += r2[n];
      if (g.act == 1) v = silu(v);
      v *= rs;
      if (g.mul) v *= g.mul[(size_t)m * g.ldmul + n];
      if (g.resid) v += g.resid[(size_t)m * g.ldres + n];
      g.C[(size_t)m * g.ldc + n] = v;
      if (g.C2 && g.c2idx[m] >= 0) g.C2[(size_t)g.c2idx[m] * g.ldc2 + n] = v;
    }
  }
}

inline cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  dim3 grid((g.M + GBM - 1) / GBM, (g.N + GBN - 1) / GBN);
  const bool vec = (g.K % 4 == 0) && (g.lda % 4 == 0) && (g.ldw % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);
  if (vec)
    gemm_simt_kernel<true><<<grid, GTHREADS, 0, st>>>(g);
  else
    gemm_simt_kernel<false><<<grid, GTHREADS, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace oard
