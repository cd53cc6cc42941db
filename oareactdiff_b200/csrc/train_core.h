// Differentiable LEFTNet core for TRAINING (BASELINE config 5): forward with saved activations and the
// hand-derived backward w.r.t. every parameter and the node-feature input.  Reference: oa_reactdiff/model/leftnet.py
// :724-891 (forward), autograd of the same for the backward; the dense formulation of the reference is kept (every stage
// runs on all E edges; masked edges carry zero geometry / rbf exactly as in the reference).
//
// Geometry is an INPUT here (edge frame, rbounds, rbf, node frame, pos_prjt): no parameter lies upstream of the
// positions, so it carries no gradient; on the device it is produced by the (validated) graph-artefact kernels of the
// inference path, in the tests by the oracle.
//
// One source, two builds:
//   * CUDA (oard.cu): `par_for` launches a grid-stride kernel over an extended __host__ __device__ lambda, scatter-adds are
//     atomicAdd, contractions go through `gemm()` (tiled SIMT kernel with generic strides; exact fp32);
//   * host emulation (train_emu.cpp, -DOARD_HOST_EMU, plain g++): `par_for` is a loop, `gemm` three loops.  The whole
//     forward + backward is checked on the CPU against the oracle's values and against golden gradients of the
//     unmodified reference's autograd (tests/test_train_emu.py) — the orchestration and every formula are exactly the
//     code the device runs.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#ifdef OARD_HOST_EMU
#define OARD_HD
#define OARD_LAMBDA [=]
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#define OARD_HD __host__ __device__
#define OARD_LAMBDA [=] __host__ __device__
#endif

namespace oard_train {

OARD_HD inline float t_silu(float x) { return x / (1.0f + expf(-x)); }
OARD_HD inline float t_dsilu(float x) {  // d/dx [x sigmoid(x)]
  const float sg = 1.0f / (1.0f + expf(-x));
  return sg * (1.0f + x * (1.0f - sg));
}
OARD_HD inline void t_atomic_add(float* p, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
// Same, for the case where ALL currently converged threads of the warp add to the SAME address (reductions of parameter
// gradients over edges / nodes): one atomic per warp instead of 32 (cooperative-groups reduce over the coalesced threads,
// valid for any active set).
OARD_HD inline void t_atomic_add_uniform(float* p, float v) {
#if defined(__CUDA_ARCH__)
  auto grp = cooperative_groups::coalesced_threads();
  const float sum = cooperative_groups::reduce(grp, v, cooperative_groups::plus<float>());
  if (grp.thread_rank() == 0) atomicAdd(p, sum);
#else
  *p += v;
#endif
}

// ------------------------------------------------------------------------------------------------ backends
#ifdef OARD_HOST_EMU
template <class F>
inline void par_for(void*, size_t n, F f) {
  for (size_t i = 0; i < n; i++) f(i);
}
inline float* dev_alloc(size_t n) { return static_cast<float*>(calloc(n ? n : 1, sizeof(float))); }
inline void dev_free(float* p) { free(p); }
inline void dev_zero(void*, float* p, size_t n) { memset(p, 0, n * sizeof(float)); }
// C[m, n] = beta * C[m, n] + sum_k A[m a_rs + k a_cs] * B[k b_rs + n b_cs]
inline void gemm(void*, int M, int N, int K, const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs,
                 float* C, long ldc, float beta) {
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      double acc = 0.0;
      for (int k = 0; k < K; k++) acc += (double)A[m * a_rs + k * a_cs] * (double)B[k * b_rs + n * b_cs];
      C[(size_t)m * ldc + n] = (beta != 0.f ? beta * C[(size_t)m * ldc + n] : 0.f) + (float)acc;
    }
}
#else
template <class F>
__global__ void k_par_for(size_t n, F f) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f(i);
}
template <class F>
inline void par_for(void* stream, size_t n, F f) {
  if (!n) return;
  const size_t blocks = (n + 255) / 256;
  k_par_for<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, (cudaStream_t)stream>>>(n, f);
}
inline float* dev_alloc(size_t n) {
  float* p = nullptr;
  if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) return nullptr;
  // cleared on the NULL stream and waited for: the caller's stream may be a non-blocking one (PyTorch's side streams are),
  // which is not ordered against the NULL stream — the clear must not land after the first kernel that writes the buffer
  cudaMemset(p, 0, (n ? n : 1) * sizeof(float));
  cudaStreamSynchronize(0);
  return p;
}
inline void dev_free(float* p) { cudaFree(p); }
inline void dev_zero(void* stream, float* p, size_t n) { cudaMemsetAsync(p, 0, n * sizeof(float), (cudaStream_t)stream); }
// Tiled SIMT GEMM with generic strides (64 x 64 tile, 16-deep K slices, 4 x 4 outputs per thread); exact fp32.
// The tile loads walk the unit-stride dimension of each operand with consecutive threads (row- or column-major A / B).
// ACC (beta == 1 only): the K range is split over blockIdx.z and the partial tiles are added with atomics — the
// weight-gradient contractions reduce over up to 2e5 edge rows into a tile grid of ~100 CTAs.
template <bool ACC>
__global__ void __launch_bounds__(256) k_gemm_strided(int M, int N, int K, const float* __restrict__ A, long a_rs, long a_cs,
                                                       const float* __restrict__ B, long b_rs, long b_cs, float* __restrict__ C,
                                                       long ldc, float beta, int k_per_z) {
  __shared__ float As[16][64 + 1], Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kb0 = ACC ? blockIdx.z * k_per_z : 0, kb1 = ACC ? min(K, kb0 + k_per_z) : K;
  float acc[4][4] = {};
  for (int k0 = kb0; k0 < kb1; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      int kk, mm;
      if (a_cs == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < kb1) ? A[(long)m * a_rs + (long)k * a_cs] : 0.f;
      int nn, kb;
      if (b_cs == 1) { nn = i & 63; kb = i >> 6; } else { kb = i & 15; nn = i >> 4; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < N && k2 < kb1) ? B[(long)k2 * b_rs + (long)n * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        if (ACC) atomicAdd(&C[(size_t)m * ldc + n], acc[i][j]);
        else C[(size_t)m * ldc + n] = (beta != 0.f ? beta * C[(size_t)m * ldc + n] : 0.f) + acc[i][j];
      }
    }
}
inline void gemm(void* stream, int M, int N, int K, const float* A, long a_rs, long a_cs, const float* B, long b_rs,
                 long b_cs, float* C, long ldc, float beta) {
  if (M <= 0 || N <= 0) return;
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  if (beta == 1.f && K >= 8192 && tiles < 148 * 8) {  // long reduction into few tiles: split K
    const int k_per_z = 2048;
    dim3 grid((N + 63) / 64, (M + 63) / 64, (K + k_per_z - 1) / k_per_z);
    k_gemm_strided<true><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, beta, k_per_z);
    return;
  }
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  k_gemm_strided<false><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, beta, 0);
}
#endif

// ------------------------------------------------------------------------------------------------ context
struct Geometry {            // dense per-edge / per-node constants of one forward (device or host pointers)
  const int* ei;             // [E] edge_index[0] (source; PyG x_j)
  const int* ej;             // [E] edge_index[1] (target; PyG x_i, aggregation index of propagate)
  const float* frame;        // [E, 3(k), 3(xyz)]: rows k = 0 coord_diff, 1 coord_cross, 2 coord_vertical (all masked)
  const float* rb;           // [E] rbounds (1 on masked edges)
  const float* rbf;          // [E, R] (masked)
  const float* inv_deg_i;    // [N] 1 / max(1, #edges with ei == t)  (GCL mean aggregation at edge_index[0])
  const float* nodeframe;    // [N, 3(xyz), 3(k)]
  const float* pos_prjt;     // [N, 3]
  const int* act_idx;        // [n_act] edges with mask = 1, ascending: EquiMessage (dir_proj / rbf_proj / messages) runs on
  int n_act;                 //   these only — on masked edges rbf = 0, so its forward AND backward contributions vanish
};

struct Ctx {
  int N = 0, E = 0, H = 0, R = 0, C = 0, L = 0, reflect = 1, legacy = 1;
  void* stream = nullptr;
  std::map<std::string, float*> W;    // parameters by reference name (device / host pointers, not owned)
  std::map<std::string, float*> dW;   // gradients by the same names (owned, zeroed by zero_grads)
  std::map<std::string, size_t> wn;   // numel
  std::map<std::string, float*> act;  // saved activations / scratch (owned)
  std::map<std::string, size_t> actn;
  float* A(const std::string& name, size_t n) {
    auto it = act.find(name);
    if (it != act.end() && actn[name] >= n) return it->second;
    if (it != act.end()) dev_free(it->second);
    float* p = dev_alloc(n);
    act[name] = p; actn[name] = n;
    return p;
  }
  float* Z(const std::string& name, size_t n) {  // zero-filled
    float* p = A(name, n);
    dev_zero(stream, p, n);
    return p;
  }
  float* w(const std::string& n) const { return W.at(n); }
  float* g(const std::string& n) { return dW.at(n); }
  void release() {
    for (auto& kv : act) dev_free(kv.second);
    for (auto& kv : dW) dev_free(kv.second);
    act.clear(); actn.clear(); dW.clear();
  }
};

// ------------------------------------------------------------------------------------------------ building blocks
// Y[M, N] = X[M, K] W^T (+ b);  W = Wp[N rows, ldw] (a column block of a wider weight when ldw > K)
inline void lin_fwd(Ctx& c, int M, int N, int K, const float* X, long ldx, const float* Wp, long ldw, const float* b,
                    float* Y, long ldy, float beta = 0.f) {
  gemm(c.stream, M, N, K, X, ldx, 1, Wp, 1, ldw, Y, ldy, beta);
  if (b) par_for(c.stream, (size_t)M * N, OARD_LAMBDA(size_t i) { Y[(i / N) * ldy + (i % N)] += b[i % N]; });
}
// dX (+)= dY W ; dW += dY^T X ; db += colsum(dY)
inline void lin_bwd(Ctx& c, int M, int N, int K, const float* X, long ldx, const float* Wp, long ldw, const float* dY,
                    long ldy, float* dX, long lddx, float beta_dx, float* dWp, float* db) {
  if (dX) gemm(c.stream, M, K, N, dY, ldy, 1, Wp, ldw, 1, dX, lddx, beta_dx);
  if (dWp) gemm(c.stream, N, K, M, dY, 1, ldy, X, ldx, 1, dWp, ldw, 1.f);
  if (db) {  // column sums: 256-row chunks in parallel, one atomic per (chunk, column)
    const int chunks = (M + 255) / 256;
    par_for(c.stream, (size_t)chunks * N, OARD_LAMBDA(size_t i) {
      const int ch = (int)(i / N), n = (int)(i % N);
      const int m1 = (ch + 1) * 256 < M ? (ch + 1) * 256 : M;
      float s = 0.f;
      for (int m = ch * 256; m < m1; m++) s += dY[(size_t)m * ldy + n];
      t_atomic_add(&db[n], s);
    });
  }
}
inline void silu_fwd(Ctx& c, size_t n, const float* pre, float* out) {
  par_for(c.stream, n, OARD_LAMBDA(size_t i) { out[i] = t_silu(pre[i]); });
}
// gpre = gout * silu'(pre)   (in place on gout allowed)
inline void silu_bwd(Ctx& c, size_t n, const float* pre, const float* gout, float* gpre) {
  par_for(c.stream, n, OARD_LAMBDA(size_t i) { gpre[i] = gout[i] * t_dsilu(pre[i]); });
}
// y = LN(x (+ add)) * gamma + beta per row (biased variance, eps 1e-5); gamma == nullptr: no affine
inline void ln_fwd(Ctx& c, int rows, int H, const float* x, const float* add, const float* gamma, const float* beta,
                   float* y) {
  par_for(c.stream, (size_t)rows, OARD_LAMBDA(size_t r) {
    const float* xr = x + r * H;
    const float* ar = add ? add + r * H : nullptr;
    float mean = 0.f;
    for (int h = 0; h < H; h++) mean += xr[h] + (ar ? ar[h] : 0.f);
    mean /= (float)H;
    float var = 0.f;
    for (int h = 0; h < H; h++) { const float d = xr[h] + (ar ? ar[h] : 0.f) - mean; var += d * d; }
    const float rstd = 1.0f / sqrtf(var / (float)H + 1e-5f);
    for (int h = 0; h < H; h++) {
      const float n = (xr[h] + (ar ? ar[h] : 0.f) - mean) * rstd;
      y[r * H + h] = gamma ? n * gamma[h] + beta[h] : n;
    }
  });
}
// gx (+)= dLN ; ggamma += sum_r gy * n ; gbeta += sum_r gy      (x is the LN input, add already included by the caller)
inline void ln_bwd(Ctx& c, int rows, int H, const float* x, const float* add, const float* gamma, const float* gy, float* gx,
                   bool acc_gx, float* ggamma, float* gbeta) {
  par_for(c.stream, (size_t)rows, OARD_LAMBDA(size_t r) {
    const float* xr = x + r * H;
    const float* ar = add ? add + r * H : nullptr;
    const float* gr = gy + r * H;
    float mean = 0.f;
    for (int h = 0; h < H; h++) mean += xr[h] + (ar ? ar[h] : 0.f);
    mean /= (float)H;
    float var = 0.f;
    for (int h = 0; h < H; h++) { const float d = xr[h] + (ar ? ar[h] : 0.f) - mean; var += d * d; }
    const float rstd = 1.0f / sqrtf(var / (float)H + 1e-5f);
    float s1 = 0.f, s2 = 0.f;  // sum(gn), sum(gn * n)
    for (int h = 0; h < H; h++) {
      const float n = (xr[h] + (ar ? ar[h] : 0.f) - mean) * rstd;
      const float gn = gr[h] * (gamma ? gamma[h] : 1.f);
      s1 += gn; s2 += gn * n;
    }
    for (int h = 0; h < H; h++) {
      const float n = (xr[h] + (ar ? ar[h] : 0.f) - mean) * rstd;
      const float gn = gr[h] * (gamma ? gamma[h] : 1.f);
      const float v = rstd * (gn - s1 / (float)H - n * s2 / (float)H);
      gx[r * H + h] = acc_gx ? gx[r * H + h] + v : v;
      if (ggamma) { t_atomic_add_uniform(&ggamma[h], gr[h] * n); t_atomic_add_uniform(&gbeta[h], gr[h]); }
    }
  });
}

// small per-element MLPs of the scalarisation: 3 -> Hq -> 1 (edge lin3) and 3 -> 48 -> 8 -> 1 (update lin3), SiLU between
struct Lin3Small { const float *w0, *b0, *w2, *b2, *w4, *b4; float *gw0, *gb0, *gw2, *gb2, *gw4, *gb4; int h1, h2; };
OARD_HD inline float lin3_eval(const Lin3Small& p, float s0, float s1, float s2) {
  if (p.h2 == 0) {
    float out = p.b2[0];
    for (int m = 0; m < p.h1; m++) out += p.w2[m] * t_silu(p.w0[m * 3] * s0 + p.w0[m * 3 + 1] * s1 + p.w0[m * 3 + 2] * s2 + p.b0[m]);
    return out;
  }
  float a2[8];
  for (int q = 0; q < p.h2; q++) a2[q] = p.b2[q];
  for (int m = 0; m < p.h1; m++) {
    const float u = t_silu(p.w0[m * 3] * s0 + p.w0[m * 3 + 1] * s1 + p.w0[m * 3 + 2] * s2 + p.b0[m]);
    for (int q = 0; q < p.h2; q++) a2[q] += p.w2[q * p.h1 + m] * u;
  }
  float out = p.b4[0];
  for (int q = 0; q < p.h2; q++) out += p.w4[q] * t_silu(a2[q]);
  return out;
}
// backward of lin3_eval for upstream gradient g: accumulates the weight gradients (atomics) and returns d/d(s0, s1, s2)
OARD_HD inline void lin3_grad(const Lin3Small& p, float s0, float s1, float s2, float g, float& g0, float& g1, float& g2) {
  g0 = g1 = g2 = 0.f;  // (no early exit for g == 0: every thread of the warp walks the same weights, see t_atomic_add_uniform)
  if (p.h2 == 0) {
    t_atomic_add_uniform(&p.gb2[0], g);
    for (int m = 0; m < p.h1; m++) {
      const float z = p.w0[m * 3] * s0 + p.w0[m * 3 + 1] * s1 + p.w0[m * 3 + 2] * s2 + p.b0[m];
      t_atomic_add_uniform(&p.gw2[m], g * t_silu(z));
      const float gz = g * p.w2[m] * t_dsilu(z);
      t_atomic_add_uniform(&p.gw0[m * 3], gz * s0); t_atomic_add_uniform(&p.gw0[m * 3 + 1], gz * s1); t_atomic_add_uniform(&p.gw0[m * 3 + 2], gz * s2);
      t_atomic_add_uniform(&p.gb0[m], gz);
      g0 += gz * p.w0[m * 3]; g1 += gz * p.w0[m * 3 + 1]; g2 += gz * p.w0[m * 3 + 2];
    }
    return;
  }
  float a2[8], ga2[8];
  for (int q = 0; q < p.h2; q++) a2[q] = p.b2[q];
  for (int m = 0; m < p.h1; m++) {
    const float u = t_silu(p.w0[m * 3] * s0 + p.w0[m * 3 + 1] * s1 + p.w0[m * 3 + 2] * s2 + p.b0[m]);
    for (int q = 0; q < p.h2; q++) a2[q] += p.w2[q * p.h1 + m] * u;
  }
  t_atomic_add_uniform(&p.gb4[0], g);
  for (int q = 0; q < p.h2; q++) {
    t_atomic_add_uniform(&p.gw4[q], g * t_silu(a2[q]));
    ga2[q] = g * p.w4[q] * t_dsilu(a2[q]);
    t_atomic_add_uniform(&p.gb2[q], ga2[q]);
  }
  for (int m = 0; m < p.h1; m++) {
    const float z = p.w0[m * 3] * s0 + p.w0[m * 3 + 1] * s1 + p.w0[m * 3 + 2] * s2 + p.b0[m];
    const float u = t_silu(z);
    float gu = 0.f;
    for (int q = 0; q < p.h2; q++) { t_atomic_add_uniform(&p.gw2[q * p.h1 + m], ga2[q] * u); gu += ga2[q] * p.w2[q * p.h1 + m]; }
    const float gz = gu * t_dsilu(z);
    t_atomic_add_uniform(&p.gw0[m * 3], gz * s0); t_atomic_add_uniform(&p.gw0[m * 3 + 1], gz * s1); t_atomic_add_uniform(&p.gw0[m * 3 + 2], gz * s2);
    t_atomic_add_uniform(&p.gb0[m], gz);
    g0 += gz * p.w0[m * 3]; g1 += gz * p.w0[m * 3 + 1]; g2 += gz * p.w0[m * 3 + 2];
  }
}

inline std::string LS(const char* a, int l, const char* b) { return std::string(a) + std::to_string(l) + b; }

// ------------------------------------------------------------------------------------------------ forward
// h_in [N, C] -> h_out [N, C], dpos [N, 3]; every intermediate the backward needs stays in c.act
inline void forward(Ctx& c, const Geometry& G, const float* h_in, float* h_out, float* dpos) {
  const int N = c.N, E = c.E, H = c.H, R = c.R, C = c.C, L = c.L, D = 3 * H + R, Hq = H / 4;
  const int* ei = G.ei; const int* ej = G.ej;
  const float* frame = G.frame; const float* rb = G.rb; const float* rbf = G.rbf;
  const float inv_sqrt_2 = 0.70710678118654752f, inv_sqrt_3 = 0.57735026918962576f, inv_sqrt_h = 1.0f / sqrtf((float)H);
  const int reflect = c.reflect;

  float* z_emb = c.A("z_emb", (size_t)N * H);
  lin_fwd(c, N, H, C, h_in, C, c.w("embedding.weight"), C, c.w("embedding.bias"), z_emb, H);
  float* rl_pre = c.A("rl_pre", (size_t)E * H); float* rl_h = c.A("rl_h", (size_t)E * H);
  lin_fwd(c, E, H, R, rbf, R, c.w("radial_lin.0.weight"), R, c.w("radial_lin.0.bias"), rl_pre, H);
  silu_fwd(c, (size_t)E * H, rl_pre, rl_h);
  float* f_pre = c.A("f_pre", (size_t)E * H); float* f = c.A("f", (size_t)E * H);
  lin_fwd(c, E, H, H, rl_h, H, c.w("radial_lin.2.weight"), H, c.w("radial_lin.2.bias"), f_pre, H);
  par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) { f[i] = rb[i / H] * f_pre[i]; });
  float* ne_pre = c.A("ne_pre", (size_t)N * H); float* ne = c.A("ne", (size_t)N * H);
  lin_fwd(c, N, H, C, h_in, C, c.w("neighbor_emb.embedding.weight"), C, c.w("neighbor_emb.embedding.bias"), ne_pre, H);
  ln_fwd(c, N, H, ne_pre, nullptr, nullptr, nullptr, ne);
  float* s = c.A("s_0", (size_t)N * H);  // s entering layer 0
  par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { s[i] = z_emb[i]; });
  par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
    const size_t e = i / H, h = i % H;
    t_atomic_add(&s[(size_t)ej[e] * H + h], f[i] * ne[(size_t)ei[e] * H + h]);
  });
  float* q_pre = c.A("q_pre", (size_t)N * H); float* q_ln = c.A("q_ln", (size_t)N * H); float* q = c.A("q", (size_t)N * H);
  lin_fwd(c, N, H, H, s, H, c.w("s2v.lin1.0.weight"), H, c.w("s2v.lin1.0.bias"), q_pre, H);
  ln_fwd(c, N, H, q_pre, nullptr, nullptr, nullptr, q_ln);
  silu_fwd(c, (size_t)N * H, q_ln, q);
  float* NE1 = c.Z("NE1", (size_t)N * 3 * H);
  par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
    const size_t e = i / H, h = i % H;
    const float fq = f[i] * q[(size_t)ei[e] * H + h];
    for (int cc = 0; cc < 3; cc++) t_atomic_add(&NE1[((size_t)ej[e] * 3 + cc) * H + h], fq * frame[e * 9 + cc]);
  });
  // edge scalarisation + lin3 -> initial edge state e_0 = [sc3 rb | sc4 rb | f | rbf]
  float* S12 = c.A("S12", (size_t)E * 2 * 3 * H);  // [E][side][k][H] (after abs)
  float* S12sgn = c.A("S12sgn", (size_t)E * 2 * H); // sign of the k = 1 row before abs
  float* e0 = c.A("e_0", (size_t)E * D);
  Lin3Small l3{c.w("lin3.0.weight"), c.w("lin3.0.bias"), c.w("lin3.2.weight"), c.w("lin3.2.bias"), nullptr, nullptr,
               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, Hq, 0};
  par_for(c.stream, (size_t)E * 2 * H, OARD_LAMBDA(size_t i) {
    const size_t e = i / (2 * H), r = i % (2 * H), side = r / H, h = r % H;
    const size_t node = side ? ej[e] : ei[e];
    float sk[3];
    for (int k = 0; k < 3; k++) {
      float a = 0.f;
      for (int cc = 0; cc < 3; cc++) a += NE1[(node * 3 + cc) * H + h] * frame[e * 9 + k * 3 + cc];
      sk[k] = a;
    }
    S12sgn[i] = sk[1] < 0.f ? -1.f : 1.f;
    if (reflect) sk[1] = fabsf(sk[1]);
    for (int k = 0; k < 3; k++) S12[((e * 2 + side) * 3 + k) * H + h] = sk[k];
    e0[e * D + side * H + h] = (lin3_eval(l3, sk[0], sk[1], sk[2]) + sk[0]) * rb[e];
  });
  par_for(c.stream, (size_t)E * (H + R), OARD_LAMBDA(size_t i) {
    const size_t e = i / (H + R), r = i % (H + R);
    e0[e * D + 2 * H + r] = r < (size_t)H ? f[e * H + r] : rbf[e * R + (r - H)];
  });
  // pos_expansion (shared weights, layer-independent input)
  float* pe_pre = c.A("pe_pre", (size_t)N * (H / 2)); float* pe_t = c.A("pe_t", (size_t)N * (H / 2)); float* pe = c.A("pe", (size_t)N * H);
  lin_fwd(c, N, H / 2, 3, G.pos_prjt, 3, c.w("pos_expansion.mlp.0.linear.weight"), 3, nullptr, pe_pre, H / 2);
  silu_fwd(c, (size_t)N * (H / 2), pe_pre, pe_t);
  lin_fwd(c, N, H, H / 2, pe_t, H / 2, c.w("pos_expansion.mlp.1.linear.weight"), H / 2, nullptr, pe, H);
  float* vec = c.Z("vec_0", (size_t)N * 3 * H);
  float* e = e0;

  for (int l = 0; l < L; l++) {
    const std::string g = LS("gcl_layers.", l, "."), ml = LS("message_layers.", l, "."), u = LS("update_layers.", l, ".");
    const std::string sl = std::to_string(l);
    const int ldw0 = 2 * H + D;
    // ---- GCLMessage
    float* x = c.A("x_" + sl, (size_t)N * H);
    ln_fwd(c, N, H, s, pe, c.w(g + "x_layernorm.weight"), c.w(g + "x_layernorm.bias"), x);
    float* P = c.A("P_" + sl, (size_t)N * H); float* Q = c.A("Q_" + sl, (size_t)N * H);
    const float* Wa = c.w(g + "edge_mlp.mlp.0.linear.weight");
    lin_fwd(c, N, H, H, x, H, Wa, ldw0, c.w(g + "edge_mlp.mlp.0.linear.bias"), P, H);
    lin_fwd(c, N, H, H, x, H, Wa + H, ldw0, nullptr, Q, H);
    float* h1_pre = c.A("h1_pre_" + sl, (size_t)E * H); float* h1 = c.A("h1_" + sl, (size_t)E * H);
    lin_fwd(c, E, H, D, e, D, Wa + 2 * H, ldw0, nullptr, h1_pre, H);
    par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
      const size_t ee = i / H, h = i % H;
      h1_pre[i] += P[(size_t)ei[ee] * H + h] + Q[(size_t)ej[ee] * H + h];
      h1[i] = t_silu(h1_pre[i]);
    });
    float* m_pre = c.A("m_pre_" + sl, (size_t)E * H); float* m = c.A("m_" + sl, (size_t)E * H);
    lin_fwd(c, E, H, H, h1, H, c.w(g + "edge_mlp.mlp.1.linear.weight"), H, c.w(g + "edge_mlp.mlp.1.linear.bias"), m_pre, H);
    silu_fwd(c, (size_t)E * H, m_pre, m);
    float* a_pre = c.A("a_pre_" + sl, (size_t)E); float* mg = c.A("mg_" + sl, (size_t)E * H);
    {
      const float* wat = c.w(g + "att_mlp.mlp.0.linear.weight"); const float* bat = c.w(g + "att_mlp.mlp.0.linear.bias");
      par_for(c.stream, (size_t)E, OARD_LAMBDA(size_t ee) {
        float a = bat[0];
        for (int h = 0; h < H; h++) a += wat[h] * m[ee * H + h];
        a_pre[ee] = a;
        const float att = t_silu(a);
        for (int h = 0; h < H; h++) mg[ee * H + h] = m[ee * H + h] * att;
      });
    }
    float* xa = c.Z("xa_" + sl, (size_t)N * 2 * H);  // [x | agg]
    {
      const float* idg = G.inv_deg_i;
      par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { xa[(i / H) * 2 * H + (i % H)] = x[i]; });
      par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
        const size_t ee = i / H, h = i % H;
        t_atomic_add(&xa[(size_t)ei[ee] * 2 * H + H + h], mg[i] * idg[ei[ee]]);
      });
    }
    float* tn_pre = c.A("tn_pre_" + sl, (size_t)N * H); float* tn = c.A("tn_" + sl, (size_t)N * H);
    lin_fwd(c, N, H, 2 * H, xa, 2 * H, c.w(g + "node_mlp.mlp.0.linear.weight"), 2 * H, c.w(g + "node_mlp.mlp.0.linear.bias"), tn_pre, H);
    silu_fwd(c, (size_t)N * H, tn_pre, tn);
    float* s1 = c.A("s1_" + sl, (size_t)N * H);  // x + node_mlp(...)   (legacy: no activation on the last layer)
    lin_fwd(c, N, H, H, tn, H, c.w(g + "node_mlp.mlp.1.linear.weight"), H, c.w(g + "node_mlp.mlp.1.linear.bias"), s1, H);
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { s1[i] += x[i]; });
    float* eo_pre = c.A("eo_pre_" + sl, (size_t)E * D); float* e_new = c.A("e_" + std::to_string(l + 1), (size_t)E * D);
    lin_fwd(c, E, D, H, mg, H, c.w(g + "edge_out_trans.mlp.0.linear.weight"), H, c.w(g + "edge_out_trans.mlp.0.linear.bias"), eo_pre, D);
    par_for(c.stream, (size_t)E * D, OARD_LAMBDA(size_t i) { e_new[i] = e[i] + t_silu(eo_pre[i]); });
    // ---- EquiMessage
    float* y = c.A("y_" + sl, (size_t)N * H);
    ln_fwd(c, N, H, s1, nullptr, c.w(ml + "x_layernorm.weight"), c.w(ml + "x_layernorm.bias"), y);
    float* xh_pre = c.A("xh_pre_" + sl, (size_t)N * H); float* xh = c.A("xh_" + sl, (size_t)N * H); float* X = c.A("X_" + sl, (size_t)N * 3 * H);
    lin_fwd(c, N, H, H, y, H, c.w(ml + "x_proj.0.weight"), H, nullptr, xh_pre, H);
    silu_fwd(c, (size_t)N * H, xh_pre, xh);
    lin_fwd(c, N, 3 * H, H, xh, H, c.w(ml + "x_proj.2.weight"), H, nullptr, X, 3 * H);
    // active edges only (compact rows p = 0 .. n_act-1 of edge act_idx[p])
    const int nA = G.n_act; const int* aidx = G.act_idx;
    float* ea = c.A("ea_" + sl, (size_t)nA * D); float* rbfa = c.A("rbfa", (size_t)nA * R);
    par_for(c.stream, (size_t)nA * D, OARD_LAMBDA(size_t i) { ea[i] = e_new[(size_t)aidx[i / D] * D + (i % D)]; });
    if (l == 0) par_for(c.stream, (size_t)nA * R, OARD_LAMBDA(size_t i) { rbfa[i] = rbf[(size_t)aidx[i / R] * R + (i % R)]; });
    float* d_pre = c.A("d_pre_" + sl, (size_t)nA * 3 * H); float* d1 = c.A("d1_" + sl, (size_t)nA * 3 * H);
    float* D2 = c.A("D2_" + sl, (size_t)nA * 3 * H); float* RB = c.A("RB_" + sl, (size_t)nA * 3 * H);
    lin_fwd(c, nA, 3 * H, D, ea, D, c.w(ml + "dir_proj.0.weight"), D, c.w(ml + "dir_proj.0.bias"), d_pre, 3 * H);
    silu_fwd(c, (size_t)nA * 3 * H, d_pre, d1);
    lin_fwd(c, nA, 3 * H, 3 * H, d1, 3 * H, c.w(ml + "dir_proj.2.weight"), 3 * H, c.w(ml + "dir_proj.2.bias"), D2, 3 * H);
    lin_fwd(c, nA, 3 * H, R, rbfa, R, c.w(ml + "rbf_proj.weight"), R, nullptr, RB, 3 * H);
    float* s2 = c.A("s2_" + sl, (size_t)N * H); float* vec1 = c.A("vec1_" + sl, (size_t)N * 3 * H);
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { s2[i] = s1[i]; });  // accumulates dx, scaled below
    par_for(c.stream, (size_t)N * 3 * H, OARD_LAMBDA(size_t i) { vec1[i] = vec[i]; });
    par_for(c.stream, (size_t)nA * H, OARD_LAMBDA(size_t i) {
      const size_t p = i / H, h = i % H, ee = aidx[p];
      const size_t a = ei[ee], t = ej[ee];
      const float al = (X[a * 3 * H + h] + X[t * 3 * H + h]) * RB[p * 3 * H + h] * D2[p * 3 * H + h];
      const float be = (X[a * 3 * H + H + h] + X[t * 3 * H + H + h]) * RB[p * 3 * H + H + h] * D2[p * 3 * H + H + h] * inv_sqrt_3;
      const float ga = (X[a * 3 * H + 2 * H + h] + X[t * 3 * H + 2 * H + h]) * RB[p * 3 * H + 2 * H + h] * D2[p * 3 * H + 2 * H + h];
      t_atomic_add(&s2[t * H + h], al);
      for (int cc = 0; cc < 3; cc++) {
        float v = vec[(a * 3 + cc) * H + h] * be + ga * frame[ee * 9 + cc];
        if (!reflect) v += al * frame[ee * 9 + 3 + cc];
        t_atomic_add(&vec1[(t * 3 + cc) * H + h], v * inv_sqrt_h);
      }
    });
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { s2[i] *= inv_sqrt_2; });
    // ---- EquiUpdate
    float* VP = c.A("VP_" + sl, (size_t)N * 3 * 2 * H);
    lin_fwd(c, 3 * N, 2 * H, H, vec1, H, c.w(u + "vec_proj.weight"), H, nullptr, VP, 2 * H);
    float* sx = c.A("sx_" + sl, (size_t)N * 2 * H); float* Sc = c.A("Sc_" + sl, (size_t)N * 3 * H); float* Scs = c.A("Scs_" + sl, (size_t)N * H);
    float* vd = c.A("vd_" + sl, (size_t)N * H);
    Lin3Small lu{c.w(u + "lin3.0.weight"), c.w(u + "lin3.0.bias"), c.w(u + "lin3.2.weight"), c.w(u + "lin3.2.bias"),
                 c.w(u + "lin3.4.weight"), c.w(u + "lin3.4.bias"), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 48, 8};
    {
      const float* nfm = G.nodeframe;
      par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
        const size_t t = i / H, h = i % H;
        float v1[3], v2[3], sk[3];
        for (int cc = 0; cc < 3; cc++) { v1[cc] = VP[(t * 3 + cc) * 2 * H + h]; v2[cc] = VP[(t * 3 + cc) * 2 * H + H + h]; }
        for (int k = 0; k < 3; k++) sk[k] = v1[0] * nfm[t * 9 + k] + v1[1] * nfm[t * 9 + 3 + k] + v1[2] * nfm[t * 9 + 6 + k];
        Scs[i] = sk[1] < 0.f ? -1.f : 1.f;
        if (reflect) sk[1] = fabsf(sk[1]);
        for (int k = 0; k < 3; k++) Sc[(t * 3 + k) * H + h] = sk[k];
        sx[t * 2 * H + h] = s2[i];
        sx[t * 2 * H + H + h] = lin3_eval(lu, sk[0], sk[1], sk[2]);
        vd[i] = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) * inv_sqrt_h;
      });
    }
    float* t2_pre = c.A("t2_pre_" + sl, (size_t)N * H); float* t2 = c.A("t2_" + sl, (size_t)N * H); float* XV = c.A("XV_" + sl, (size_t)N * 3 * H);
    lin_fwd(c, N, H, 2 * H, sx, 2 * H, c.w(u + "xvec_proj.0.weight"), 2 * H, nullptr, t2_pre, H);
    silu_fwd(c, (size_t)N * H, t2_pre, t2);
    lin_fwd(c, N, 3 * H, H, t2, H, c.w(u + "xvec_proj.2.weight"), H, nullptr, XV, 3 * H);
    float* s_next = c.A("s_" + std::to_string(l + 1), (size_t)N * H); float* vec_next = c.A("vec_" + std::to_string(l + 1), (size_t)N * 3 * H);
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
      const size_t t = i / H, h = i % H;
      s_next[i] = s2[i] + (XV[t * 3 * H + h] + XV[t * 3 * H + H + h] + vd[i]) * inv_sqrt_2;
      const float x3 = XV[t * 3 * H + 2 * H + h];
      for (int cc = 0; cc < 3; cc++) vec_next[(t * 3 + cc) * H + h] = vec1[(t * 3 + cc) * H + h] + x3 * VP[(t * 3 + cc) * 2 * H + H + h];
    });
    s = s_next; vec = vec_next; e = e_new;
  }
  // ---- output head
  const std::string o = "out_pos.output_network.0.";
  float* O1 = c.A("O1", (size_t)N * 3 * H); float* sn = c.A("sn", (size_t)N * 2 * H);
  lin_fwd(c, 3 * N, H, H, vec, H, c.w(o + "vec1_proj.weight"), H, nullptr, O1, H);
  par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
    const size_t t = i / H, h = i % H;
    const float a = O1[(t * 3) * H + h], b = O1[(t * 3 + 1) * H + h], cc = O1[(t * 3 + 2) * H + h];
    sn[t * 2 * H + h] = s[i];
    sn[t * 2 * H + H + h] = sqrtf(a * a + b * b + cc * cc);
  });
  float* tu_pre = c.A("tu_pre", (size_t)N * H); float* tu = c.A("tu", (size_t)N * H);
  lin_fwd(c, N, H, 2 * H, sn, 2 * H, c.w(o + "update_net.0.weight"), 2 * H, c.w(o + "update_net.0.bias"), tu_pre, H);
  silu_fwd(c, (size_t)N * H, tu_pre, tu);
  float* gate = c.A("gate", (size_t)N); float* vdot2 = c.A("vdot2", (size_t)N * 3);
  {
    const float* wu2 = c.w(o + "update_net.2.weight"); const float* bu2 = c.w(o + "update_net.2.bias");
    const float* wo2 = c.w(o + "vec2_proj.weight");
    par_for(c.stream, (size_t)N, OARD_LAMBDA(size_t t) {
      float gt = bu2[1];
      for (int h = 0; h < H; h++) gt += wu2[H + h] * tu[t * H + h];
      gate[t] = gt;
      for (int cc = 0; cc < 3; cc++) {
        float d = 0.f;
        for (int h = 0; h < H; h++) d += vec[(t * 3 + cc) * H + h] * wo2[h];
        vdot2[t * 3 + cc] = d;
        dpos[t * 3 + cc] = gt * d;
      }
    });
  }
  lin_fwd(c, N, C, H, s, H, c.w("embedding_out.weight"), H, c.w("embedding_out.bias"), h_out, C);
}

// ------------------------------------------------------------------------------------------------ backward
// Upstream gradients g_hout [N, C], g_dpos [N, 3] -> g_hin [N, C]; parameter gradients are ACCUMULATED into c.dW.
// Must follow forward() on the same context (reads its saved activations).
inline void backward(Ctx& c, const Geometry& G, const float* h_in, const float* g_hout, const float* g_dpos, float* g_hin) {
  const int N = c.N, E = c.E, H = c.H, R = c.R, C = c.C, L = c.L, D = 3 * H + R, Hq = H / 4;
  const int* ei = G.ei; const int* ej = G.ej;
  const float* frame = G.frame; const float* rb = G.rb; const float* rbf = G.rbf;
  const float inv_sqrt_2 = 0.70710678118654752f, inv_sqrt_3 = 0.57735026918962576f, inv_sqrt_h = 1.0f / sqrtf((float)H);
  const int reflect = c.reflect;
  auto act = [&](const std::string& n) { return c.act.at(n); };
  const std::string sL = std::to_string(L);

  // ---- output head
  const std::string o = "out_pos.output_network.0.";
  float* s = act("s_" + sL); float* vec = act("vec_" + sL);
  float* gs = c.A("g_s", (size_t)N * H); float* gvec = c.A("g_vec", (size_t)N * 3 * H); float* ge = c.Z("g_e", (size_t)E * D);
  lin_bwd(c, N, C, H, s, H, c.w("embedding_out.weight"), H, g_hout, C, gs, H, 0.f, c.g("embedding_out.weight"), c.g("embedding_out.bias"));
  float* g_tu = c.A("g_tu", (size_t)N * H);
  {
    const float* gate = act("gate"); const float* vdot2 = act("vdot2"); const float* tu = act("tu");
    const float* wu2 = c.w(o + "update_net.2.weight"); const float* wo2 = c.w(o + "vec2_proj.weight");
    float* g_wu2 = c.g(o + "update_net.2.weight"); float* g_bu2 = c.g(o + "update_net.2.bias"); float* g_wo2 = c.g(o + "vec2_proj.weight");
    par_for(c.stream, (size_t)N, OARD_LAMBDA(size_t t) {
      float gg = 0.f;
      for (int cc = 0; cc < 3; cc++) gg += g_dpos[t * 3 + cc] * vdot2[t * 3 + cc];
      t_atomic_add_uniform(&g_bu2[1], gg);
      for (int h = 0; h < H; h++) {
        g_tu[t * H + h] = gg * wu2[H + h];
        t_atomic_add_uniform(&g_wu2[H + h], gg * tu[t * H + h]);
        float acc = 0.f;
        for (int cc = 0; cc < 3; cc++) {
          const float gv = g_dpos[t * 3 + cc] * gate[t];
          gvec[(t * 3 + cc) * H + h] = gv * wo2[h];
          acc += gv * vec[(t * 3 + cc) * H + h];
        }
        t_atomic_add_uniform(&g_wo2[h], acc);
      }
    });
  }
  silu_bwd(c, (size_t)N * H, act("tu_pre"), g_tu, g_tu);
  float* g_sn = c.A("g_sn", (size_t)N * 2 * H);
  lin_bwd(c, N, H, 2 * H, act("sn"), 2 * H, c.w(o + "update_net.0.weight"), 2 * H, g_tu, H, g_sn, 2 * H, 0.f,
          c.g(o + "update_net.0.weight"), c.g(o + "update_net.0.bias"));
  float* g_O1 = c.A("g_O1", (size_t)N * 3 * H);
  {
    const float* O1 = act("O1"); const float* sn = act("sn");
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
      const size_t t = i / H, h = i % H;
      gs[i] += g_sn[t * 2 * H + h];
      const float nrm = sn[t * 2 * H + H + h], gn = g_sn[t * 2 * H + H + h];
      for (int cc = 0; cc < 3; cc++) g_O1[(t * 3 + cc) * H + h] = nrm > 0.f ? gn * O1[(t * 3 + cc) * H + h] / nrm : 0.f;
    });
  }
  lin_bwd(c, 3 * N, H, H, vec, H, c.w(o + "vec1_proj.weight"), H, g_O1, H, gvec, H, 1.f, c.g(o + "vec1_proj.weight"), nullptr);

  float* g_pe = c.Z("g_pe", (size_t)N * H);
  for (int l = L - 1; l >= 0; l--) {
    const std::string g = LS("gcl_layers.", l, "."), ml = LS("message_layers.", l, "."), u = LS("update_layers.", l, ".");
    const std::string sl = std::to_string(l);
    const int ldw0 = 2 * H + D;
    float* s_l = act("s_" + sl); float* vec_l = act("vec_" + sl); float* e_l = act("e_" + sl); float* e_new = act("e_" + std::to_string(l + 1));
    float* x = act("x_" + sl); float* xa = act("xa_" + sl); float* s1 = act("s1_" + sl); float* s2 = act("s2_" + sl);
    float* vec1 = act("vec1_" + sl); float* VP = act("VP_" + sl); float* XV = act("XV_" + sl); float* X = act("X_" + sl);
    float* D2 = act("D2_" + sl); float* RB = act("RB_" + sl); float* m = act("m_" + sl); float* mg = act("mg_" + sl);
    (void)s2; (void)xa; (void)e_new;
    // ---- EquiUpdate (apply)
    float* g_XV = c.A("g_XV", (size_t)N * 3 * H); float* g_VP = c.Z("g_VP", (size_t)N * 3 * 2 * H); float* g_vd = c.A("g_vd", (size_t)N * H);
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
      const size_t t = i / H, h = i % H;
      const float gq = gs[i] * inv_sqrt_2;
      g_XV[t * 3 * H + h] = gq; g_XV[t * 3 * H + H + h] = gq; g_vd[i] = gq;
      const float x3 = XV[t * 3 * H + 2 * H + h];
      float acc = 0.f;
      for (int cc = 0; cc < 3; cc++) {
        const float gv = gvec[(t * 3 + cc) * H + h];
        acc += gv * VP[(t * 3 + cc) * 2 * H + H + h];
        g_VP[(t * 3 + cc) * 2 * H + H + h] = gv * x3;
      }
      g_XV[t * 3 * H + 2 * H + h] = acc;
    });
    // gs now plays g_s2, gvec plays g_vec1
    float* g_t2 = c.A("g_t2", (size_t)N * H);
    lin_bwd(c, N, 3 * H, H, act("t2_" + sl), H, c.w(u + "xvec_proj.2.weight"), H, g_XV, 3 * H, g_t2, H, 0.f, c.g(u + "xvec_proj.2.weight"), nullptr);
    silu_bwd(c, (size_t)N * H, act("t2_pre_" + sl), g_t2, g_t2);
    float* g_sx = c.A("g_sx", (size_t)N * 2 * H);
    lin_bwd(c, N, H, 2 * H, act("sx_" + sl), 2 * H, c.w(u + "xvec_proj.0.weight"), 2 * H, g_t2, H, g_sx, 2 * H, 0.f, c.g(u + "xvec_proj.0.weight"), nullptr);
    {
      Lin3Small lu{c.w(u + "lin3.0.weight"), c.w(u + "lin3.0.bias"), c.w(u + "lin3.2.weight"), c.w(u + "lin3.2.bias"),
                   c.w(u + "lin3.4.weight"), c.w(u + "lin3.4.bias"), c.g(u + "lin3.0.weight"), c.g(u + "lin3.0.bias"),
                   c.g(u + "lin3.2.weight"), c.g(u + "lin3.2.bias"), c.g(u + "lin3.4.weight"), c.g(u + "lin3.4.bias"), 48, 8};
      const float* Sc = act("Sc_" + sl); const float* Scs = act("Scs_" + sl); const float* nfm = G.nodeframe;
      par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) {
        const size_t t = i / H, h = i % H;
        gs[i] += g_sx[t * 2 * H + h];
        float gS[3];
        lin3_grad(lu, Sc[(t * 3) * H + h], Sc[(t * 3 + 1) * H + h], Sc[(t * 3 + 2) * H + h], g_sx[t * 2 * H + H + h], gS[0], gS[1], gS[2]);
        if (reflect) gS[1] *= Scs[i];
        const float gd = g_vd[i] * inv_sqrt_h;
        for (int cc = 0; cc < 3; cc++) {
          const float v1 = VP[(t * 3 + cc) * 2 * H + h], v2 = VP[(t * 3 + cc) * 2 * H + H + h];
          g_VP[(t * 3 + cc) * 2 * H + h] = gS[0] * nfm[t * 9 + cc * 3] + gS[1] * nfm[t * 9 + cc * 3 + 1] + gS[2] * nfm[t * 9 + cc * 3 + 2] + gd * v2;
          g_VP[(t * 3 + cc) * 2 * H + H + h] += gd * v1;
        }
      });
    }
    lin_bwd(c, 3 * N, 2 * H, H, vec1, H, c.w(u + "vec_proj.weight"), H, g_VP, 2 * H, gvec, H, 1.f, c.g(u + "vec_proj.weight"), nullptr);
    // ---- EquiMessage: s2 = (s1 + dx)/sqrt2, vec1 = vec_l + dvec
    const int nA = G.n_act; const int* aidx = G.act_idx;
    float* g_X = c.Z("g_X", (size_t)N * 3 * H); float* g_RB = c.A("g_RB", (size_t)nA * 3 * H); float* g_D2 = c.A("g_D2", (size_t)nA * 3 * H);
    float* g_vecl = c.A("g_vecl", (size_t)N * 3 * H);
    par_for(c.stream, (size_t)N * 3 * H, OARD_LAMBDA(size_t i) { g_vecl[i] = gvec[i]; });
    par_for(c.stream, (size_t)nA * H, OARD_LAMBDA(size_t i) {
      const size_t p = i / H, h = i % H, ee = aidx[p];
      const size_t a = ei[ee], t = ej[ee];
      float xs[3], gm[3];
      for (int k = 0; k < 3; k++) {
        xs[k] = X[a * 3 * H + k * H + h] + X[t * 3 * H + k * H + h];
        gm[k] = RB[p * 3 * H + k * H + h] * D2[p * 3 * H + k * H + h];
      }
      const float be = xs[1] * gm[1] * inv_sqrt_3;
      float g_al = gs[t * H + h] * inv_sqrt_2, g_be = 0.f, g_ga = 0.f;
      for (int cc = 0; cc < 3; cc++) {
        const float gd = gvec[(t * 3 + cc) * H + h] * inv_sqrt_h;
        g_be += gd * vec_l[(a * 3 + cc) * H + h];
        g_ga += gd * frame[ee * 9 + cc];
        if (!reflect) g_al += gd * frame[ee * 9 + 3 + cc];
        t_atomic_add(&g_vecl[(a * 3 + cc) * H + h], gd * be);
      }
      const float gxs[3] = {g_al * gm[0], g_be * gm[1] * inv_sqrt_3, g_ga * gm[2]};
      const float ggm[3] = {g_al * xs[0], g_be * xs[1] * inv_sqrt_3, g_ga * xs[2]};
      for (int k = 0; k < 3; k++) {
        t_atomic_add(&g_X[a * 3 * H + k * H + h], gxs[k]);
        t_atomic_add(&g_X[t * 3 * H + k * H + h], gxs[k]);
        g_RB[p * 3 * H + k * H + h] = ggm[k] * D2[p * 3 * H + k * H + h];
        g_D2[p * 3 * H + k * H + h] = ggm[k] * RB[p * 3 * H + k * H + h];
      }
    });
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { gs[i] *= inv_sqrt_2; });  // g_s1 (from s2)
    lin_bwd(c, nA, 3 * H, R, act("rbfa"), R, c.w(ml + "rbf_proj.weight"), R, g_RB, 3 * H, nullptr, 0, 0.f, c.g(ml + "rbf_proj.weight"), nullptr);
    float* g_d1 = c.A("g_d1", (size_t)nA * 3 * H);
    lin_bwd(c, nA, 3 * H, 3 * H, act("d1_" + sl), 3 * H, c.w(ml + "dir_proj.2.weight"), 3 * H, g_D2, 3 * H, g_d1, 3 * H, 0.f,
            c.g(ml + "dir_proj.2.weight"), c.g(ml + "dir_proj.2.bias"));
    silu_bwd(c, (size_t)nA * 3 * H, act("d_pre_" + sl), g_d1, g_d1);
    float* g_ea = c.A("g_ea", (size_t)nA * D);
    lin_bwd(c, nA, 3 * H, D, act("ea_" + sl), D, c.w(ml + "dir_proj.0.weight"), D, g_d1, 3 * H, g_ea, D, 0.f,
            c.g(ml + "dir_proj.0.weight"), c.g(ml + "dir_proj.0.bias"));
    par_for(c.stream, (size_t)nA * D, OARD_LAMBDA(size_t i) { ge[(size_t)aidx[i / D] * D + (i % D)] += g_ea[i]; });  // distinct rows
    float* g_xh = c.A("g_xh", (size_t)N * H); float* g_y = c.A("g_y", (size_t)N * H);
    lin_bwd(c, N, 3 * H, H, act("xh_" + sl), H, c.w(ml + "x_proj.2.weight"), H, g_X, 3 * H, g_xh, H, 0.f, c.g(ml + "x_proj.2.weight"), nullptr);
    silu_bwd(c, (size_t)N * H, act("xh_pre_" + sl), g_xh, g_xh);
    lin_bwd(c, N, H, H, act("y_" + sl), H, c.w(ml + "x_proj.0.weight"), H, g_xh, H, g_y, H, 0.f, c.g(ml + "x_proj.0.weight"), nullptr);
    ln_bwd(c, N, H, s1, nullptr, c.w(ml + "x_layernorm.weight"), g_y, gs, true, c.g(ml + "x_layernorm.weight"), c.g(ml + "x_layernorm.bias"));
    // ---- GCLMessage.  ge = grad w.r.t. e_{l+1} (complete now); gs = grad w.r.t. s1
    float* g_eo = c.A("g_eo", (size_t)E * D);
    silu_bwd(c, (size_t)E * D, act("eo_pre_" + sl), ge, g_eo);
    float* g_mg = c.A("g_mg", (size_t)E * H);
    lin_bwd(c, E, D, H, mg, H, c.w(g + "edge_out_trans.mlp.0.linear.weight"), H, g_eo, D, g_mg, H, 0.f,
            c.g(g + "edge_out_trans.mlp.0.linear.weight"), c.g(g + "edge_out_trans.mlp.0.linear.bias"));
    float* g_tn = c.A("g_tn", (size_t)N * H); float* g_xa = c.A("g_xa", (size_t)N * 2 * H);
    lin_bwd(c, N, H, H, act("tn_" + sl), H, c.w(g + "node_mlp.mlp.1.linear.weight"), H, gs, H, g_tn, H, 0.f,
            c.g(g + "node_mlp.mlp.1.linear.weight"), c.g(g + "node_mlp.mlp.1.linear.bias"));
    silu_bwd(c, (size_t)N * H, act("tn_pre_" + sl), g_tn, g_tn);
    lin_bwd(c, N, H, 2 * H, act("xa_" + sl), 2 * H, c.w(g + "node_mlp.mlp.0.linear.weight"), 2 * H, g_tn, H, g_xa, 2 * H, 0.f,
            c.g(g + "node_mlp.mlp.0.linear.weight"), c.g(g + "node_mlp.mlp.0.linear.bias"));
    float* g_x = c.A("g_x", (size_t)N * H);
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { g_x[i] = gs[i] + g_xa[(i / H) * 2 * H + (i % H)]; });
    float* g_m = c.A("g_m", (size_t)E * H);
    {
      const float* a_pre = act("a_pre_" + sl); const float* idg = G.inv_deg_i;
      const float* wat = c.w(g + "att_mlp.mlp.0.linear.weight");
      float* g_wat = c.g(g + "att_mlp.mlp.0.linear.weight"); float* g_bat = c.g(g + "att_mlp.mlp.0.linear.bias");
      par_for(c.stream, (size_t)E, OARD_LAMBDA(size_t ee) {
        const size_t a = ei[ee];
        const float att = t_silu(a_pre[ee]);
        float g_att = 0.f;
        for (int h = 0; h < H; h++) {
          const float gmg = g_mg[ee * H + h] + g_xa[a * 2 * H + H + h] * idg[a];
          g_att += gmg * m[ee * H + h];
          g_m[ee * H + h] = gmg * att;
        }
        const float ga = g_att * t_dsilu(a_pre[ee]);
        t_atomic_add_uniform(&g_bat[0], ga);
        for (int h = 0; h < H; h++) {
          g_m[ee * H + h] += ga * wat[h];
          t_atomic_add_uniform(&g_wat[h], ga * m[ee * H + h]);
        }
      });
    }
    silu_bwd(c, (size_t)E * H, act("m_pre_" + sl), g_m, g_m);
    float* g_h1 = c.A("g_h1", (size_t)E * H);
    lin_bwd(c, E, H, H, act("h1_" + sl), H, c.w(g + "edge_mlp.mlp.1.linear.weight"), H, g_m, H, g_h1, H, 0.f,
            c.g(g + "edge_mlp.mlp.1.linear.weight"), c.g(g + "edge_mlp.mlp.1.linear.bias"));
    silu_bwd(c, (size_t)E * H, act("h1_pre_" + sl), g_h1, g_h1);
    const float* Wa = c.w(g + "edge_mlp.mlp.0.linear.weight"); float* gWa = c.g(g + "edge_mlp.mlp.0.linear.weight");
    // grad w.r.t. e_l = residual path (ge) + through the edge MLP
    lin_bwd(c, E, H, D, e_l, D, Wa + 2 * H, ldw0, g_h1, H, ge, D, 1.f, gWa + 2 * H, nullptr);
    float* g_P = c.Z("g_P", (size_t)N * H); float* g_Q = c.Z("g_Q", (size_t)N * H);
    par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
      const size_t ee = i / H, h = i % H;
      t_atomic_add(&g_P[(size_t)ei[ee] * H + h], g_h1[i]);
      t_atomic_add(&g_Q[(size_t)ej[ee] * H + h], g_h1[i]);
    });
    lin_bwd(c, N, H, H, x, H, Wa, ldw0, g_P, H, g_x, H, 1.f, gWa, c.g(g + "edge_mlp.mlp.0.linear.bias"));
    lin_bwd(c, N, H, H, x, H, Wa + H, ldw0, g_Q, H, g_x, H, 1.f, gWa + H, nullptr);
    // x = LN(s_l + pe): gradient w.r.t. (s_l + pe) -> both s_l and pe
    ln_bwd(c, N, H, s_l, act("pe"), c.w(g + "x_layernorm.weight"), g_x, gs, false, c.g(g + "x_layernorm.weight"), c.g(g + "x_layernorm.bias"));
    par_for(c.stream, (size_t)N * H, OARD_LAMBDA(size_t i) { g_pe[i] += gs[i]; });
    par_for(c.stream, (size_t)N * 3 * H, OARD_LAMBDA(size_t i) { gvec[i] = g_vecl[i]; });
    (void)vec_l;
  }
  // ---- pos_expansion
  {
    float* g_pet = c.A("g_pet", (size_t)N * (H / 2));
    lin_bwd(c, N, H, H / 2, act("pe_t"), H / 2, c.w("pos_expansion.mlp.1.linear.weight"), H / 2, g_pe, H, g_pet, H / 2, 0.f,
            c.g("pos_expansion.mlp.1.linear.weight"), nullptr);
    silu_bwd(c, (size_t)N * (H / 2), act("pe_pre"), g_pet, g_pet);
    lin_bwd(c, N, H / 2, 3, G.pos_prjt, 3, c.w("pos_expansion.mlp.0.linear.weight"), 3, g_pet, H / 2, nullptr, 0, 0.f,
            c.g("pos_expansion.mlp.0.linear.weight"), nullptr);
  }
  // ---- initial edge state e_0 = [sc3 rb | sc4 rb | f | rbf]   (ge = grad w.r.t. e_0, gs = grad w.r.t. s_0)
  float* g_f = c.A("g_f", (size_t)E * H); float* g_NE1 = c.Z("g_NE1", (size_t)N * 3 * H);
  par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) { g_f[i] = ge[(i / H) * D + 2 * H + (i % H)]; });
  {
    Lin3Small l3{c.w("lin3.0.weight"), c.w("lin3.0.bias"), c.w("lin3.2.weight"), c.w("lin3.2.bias"), nullptr, nullptr,
                 c.g("lin3.0.weight"), c.g("lin3.0.bias"), c.g("lin3.2.weight"), c.g("lin3.2.bias"), nullptr, nullptr, Hq, 0};
    const float* S12 = act("S12"); const float* S12sgn = act("S12sgn");
    par_for(c.stream, (size_t)E * 2 * H, OARD_LAMBDA(size_t i) {
      const size_t e = i / (2 * H), r = i % (2 * H), side = r / H, h = r % H;
      const size_t node = side ? ej[e] : ei[e];
      const float gsc = ge[e * D + side * H + h] * rb[e];
      const float* S = S12 + ((e * 2 + side) * 3) * H + h;
      float gS[3];
      lin3_grad(l3, S[0], S[H], S[2 * H], gsc, gS[0], gS[1], gS[2]);
      gS[0] += gsc;
      if (reflect) gS[1] *= S12sgn[i];
      for (int cc = 0; cc < 3; cc++)
        t_atomic_add(&g_NE1[(node * 3 + cc) * H + h], gS[0] * frame[e * 9 + cc] + gS[1] * frame[e * 9 + 3 + cc] + gS[2] * frame[e * 9 + 6 + cc]);
    });
  }
  float* g_q = c.Z("g_q", (size_t)N * H);
  {
    const float* f = act("f"); const float* q = act("q");
    par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
      const size_t e = i / H, h = i % H;
      const size_t a = ei[e], t = ej[e];
      float gsum = 0.f;
      for (int cc = 0; cc < 3; cc++) gsum += g_NE1[(t * 3 + cc) * H + h] * frame[e * 9 + cc];
      g_f[i] += gsum * q[a * H + h];
      t_atomic_add(&g_q[a * H + h], gsum * f[i]);
    });
  }
  silu_bwd(c, (size_t)N * H, act("q_ln"), g_q, g_q);
  float* g_qpre = c.A("g_qpre", (size_t)N * H);
  ln_bwd(c, N, H, act("q_pre"), nullptr, nullptr, g_q, g_qpre, false, nullptr, nullptr);
  lin_bwd(c, N, H, H, act("s_0"), H, c.w("s2v.lin1.0.weight"), H, g_qpre, H, gs, H, 1.f, c.g("s2v.lin1.0.weight"), c.g("s2v.lin1.0.bias"));
  // ---- NeighborEmb: s_0 = z_emb + sum_{e: ej = t} f_e * ne[ei]
  float* g_ne = c.Z("g_ne", (size_t)N * H);
  {
    const float* f = act("f"); const float* ne = act("ne");
    par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) {
      const size_t e = i / H, h = i % H;
      const float gt = gs[(size_t)ej[e] * H + h];
      g_f[i] += gt * ne[(size_t)ei[e] * H + h];
      t_atomic_add(&g_ne[(size_t)ei[e] * H + h], gt * f[i]);
    });
  }
  float* g_nepre = c.A("g_nepre", (size_t)N * H);
  ln_bwd(c, N, H, act("ne_pre"), nullptr, nullptr, g_ne, g_nepre, false, nullptr, nullptr);
  lin_bwd(c, N, H, C, h_in, C, c.w("neighbor_emb.embedding.weight"), C, g_nepre, H, g_hin, C, 0.f,
          c.g("neighbor_emb.embedding.weight"), c.g("neighbor_emb.embedding.bias"));
  lin_bwd(c, N, H, C, h_in, C, c.w("embedding.weight"), C, gs, H, g_hin, C, 1.f, c.g("embedding.weight"), c.g("embedding.bias"));
  // ---- radial_lin: f = rb * (W2 silu(W1 rbf + b1) + b2)
  par_for(c.stream, (size_t)E * H, OARD_LAMBDA(size_t i) { g_f[i] *= rb[i / H]; });
  float* g_rlh = c.A("g_rlh", (size_t)E * H);
  lin_bwd(c, E, H, H, act("rl_h"), H, c.w("radial_lin.2.weight"), H, g_f, H, g_rlh, H, 0.f, c.g("radial_lin.2.weight"), c.g("radial_lin.2.bias"));
  silu_bwd(c, (size_t)E * H, act("rl_pre"), g_rlh, g_rlh);
  lin_bwd(c, E, H, R, rbf, R, c.w("radial_lin.0.weight"), R, g_rlh, H, nullptr, 0, 0.f, c.g("radial_lin.0.weight"), c.g("radial_lin.0.bias"));
}

}  // namespace oard_train
