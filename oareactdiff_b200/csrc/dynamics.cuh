// Device-resident EGNNDynamics wrapper and reverse-diffusion step around the LEFTNet forward (sm_100a).
// Citations: reference oa_reactdiff/dynamics/egnn_dynamics.py (prologue :91-119, epilogue :137-168),
// oa_reactdiff/diffusion/en_diffusion.py (sample_p_zs_given_zt :562-632, noise :281-304), diffusion/_utils.py:9-12.
//
// The reference runs ~95 tiny ATen launches and several host syncs around the denoiser at every reverse step; here the
// whole step is three kernels around the forward, so that one step is one CUDA-graph launch.  A "segment" is the set of
// nodes of one (fragment, sample) pair — contiguous in the reference's node order — and every centre-of-mass removal
// of the path is a mean over a segment: one warp per segment, fixed summation order, no atomics.
#pragma once
#include "common.cuh"

namespace oard {

constexpr int DYN_MAX_FRAG = 8;
constexpr int DYN_MAX_D = 16;  // feature width (node_nf - pos_dim), hidden width 2 d <= 32, embed width <= 16

struct DynCodec {  // encoders[f] / decoders[f]: MLP(d -> 2d -> out), SiLU between (dynamics/_base.py:91-109)
  const float *ew0[DYN_MAX_FRAG], *eb0[DYN_MAX_FRAG], *ew1[DYN_MAX_FRAG], *eb1[DYN_MAX_FRAG];
  const float *dw0[DYN_MAX_FRAG], *db0[DYN_MAX_FRAG], *dw1[DYN_MAX_FRAG], *db1[DYN_MAX_FRAG];
};

// step_params[0..3] = t, alpha_ts, coef, sigma ; [4] = nan-guard draw counter (as int bits) ; [5..6] = alpha_s, sigma_s of the
// RePaint blend (q(z_s | x) of the clamped fragments)
__global__ void k_set_params(float* __restrict__ prm, float t, float alpha_ts, float coef, float sigma, int counter,
                             float alpha_s, float sigma_s) {
  prm[0] = t; prm[1] = alpha_ts; prm[2] = coef; prm[3] = sigma;
  reinterpret_cast<int*>(prm)[4] = counter;
  prm[5] = alpha_s; prm[6] = sigma_s;
}

// Prologue (egnn_dynamics.py:91-119): pos = xh[:, :3]; h = [encoder_f(xh[:, 3:]) | t | condition[sample]].
// t: per-sample device array t_dev[B] if given, else the scalar prm[0].  One thread per node.  Also clears the NaN flag.
__global__ void k_dyn_pre(int N, int nf, int d, int emb, int C, int cnd, int cond_time, const float* __restrict__ xh,
                          const int* __restrict__ node_frag, const int* __restrict__ node_sample, DynCodec cd,
                          const float* __restrict__ t_dev, const float* __restrict__ prm, const float* __restrict__ cond,
                          float* __restrict__ pos, float* __restrict__ h_in, int* __restrict__ nan_flag) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) *nan_flag = 0;
  if (n >= N) return;
  const float* x = xh + (size_t)n * nf;
  pos[n * 3 + 0] = x[0]; pos[n * 3 + 1] = x[1]; pos[n * 3 + 2] = x[2];
  const int f = node_frag[n], b = node_sample[n];
  float in[DYN_MAX_D], hid[2 * DYN_MAX_D];
  for (int k = 0; k < d; k++) in[k] = x[3 + k];
  const float *w0 = cd.ew0[f], *b0 = cd.eb0[f], *w1 = cd.ew1[f], *b1 = cd.eb1[f];
  for (int j = 0; j < 2 * d; j++) {
    float a = b0[j];
    for (int k = 0; k < d; k++) a = fmaf(w0[j * d + k], in[k], a);
    hid[j] = silu(a);
  }
  float* o = h_in + (size_t)n * C;
  for (int j = 0; j < emb; j++) {
    float a = b1[j];
    for (int k = 0; k < 2 * d; k++) a = fmaf(w1[j * 2 * d + k], hid[k], a);
    o[j] = a;
  }
  int c = emb;
  if (cond_time) o[c++] = t_dev ? t_dev[b] : prm[0];
  for (int k = 0; k < cnd; k++) o[c++] = cond[(size_t)b * cnd + k];
}

// vel = pos_final - pos with pos_final = pos + dpos rounded to fp32 first, as the reference computes it
// (leftnet.py:878, egnn_dynamics.py:137); flag = any NaN (egnn_dynamics.py:138-143).
__global__ void k_dyn_vel(int N3, const float* __restrict__ pos, const float* __restrict__ dpos, float* __restrict__ vel,
                          int* __restrict__ nan_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N3) return;
  const float v = __fsub_rn(__fadd_rn(pos[i], dpos[i]), pos[i]);
  vel[i] = v;
  if (v != v) atomicOr(nan_flag, 1);
}

// Counter-based standard normal for the NaN guard's replacement noise (private stream: the caller's generator is not
// touched; the reference draws torch.randn_like there, egnn_dynamics.py:143).
__device__ __forceinline__ uint32_t dyn_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float dyn_gauss(uint32_t counter, uint32_t idx) {
  const uint32_t a = dyn_hash(idx * 2u + 0x9e3779b9u * (counter + 1u)), b = dyn_hash(a ^ (idx * 2u + 1u));
  const float u1 = ((a >> 8) + 1u) * (1.0f / 16777216.0f), u2 = (b >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// Epilogue of the dynamics (egnn_dynamics.py:137-168): NaN guard, per-(fragment, sample) centre-of-mass removal of the
// velocity, decoder_f on the first `emb` output channels -> eps = [vel | decoder(h)].
// MODE 0: write eps[N, nf].
// MODE 1: the reverse step z_t -> z_s in place (en_diffusion.py:562-632 with one (s, t) pair for the whole batch):
//   mu = z / alpha_ts - eps * coef ; z_s = mu + sigma * noise ; noise positions and z_s positions are projected on the
//   zero-CoM subspace per segment (_utils.py:9-12, en_diffusion.py:281-304, 627-631); noise_h == NULL: feature noise is
//   zero (pos_only); h0 != NULL: the features are overwritten by h0 (en_diffusion.py:524-527).  Same fp32 operation order as the host formulas
//   (division kept, no contraction).  One warp per segment.
// MODE 2: MODE 1 + the RePaint blend (en_diffusion.py:803-851): segments of the fragments in `known_bits` are not stepped but
//   re-drawn from q(z_s | x_fixed) = alpha_s x_fixed + sigma_s eps' (noised_representation, :260-279) with the second set of
//   raw draws noise2_x / noise2_h (position noise CoM-free per segment; noise2_h == NULL: zero), then the h0 overwrite.
template <int MODE>
__global__ void k_dyn_post(int S, int nf, int d, int emb, int C, const int* __restrict__ seg_ptr,
                           const int* __restrict__ seg_frag, DynCodec cd, const float* __restrict__ vel,
                           const float* __restrict__ h_out, const int* __restrict__ nan_flag,
                           const float* __restrict__ prm, float* __restrict__ eps_out, float* __restrict__ z,
                           const float* __restrict__ noise_x, const float* __restrict__ noise_h,
                           const float* __restrict__ h0, const float* __restrict__ x_fixed = nullptr, int known_bits = 0,
                           const float* __restrict__ noise2_x = nullptr, const float* __restrict__ noise2_h = nullptr) {
  const int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (sgm >= S) return;
  const int n0 = seg_ptr[sgm], n1 = seg_ptr[sgm + 1], f = seg_frag[sgm];
  if (MODE == 2 && ((known_bits >> f) & 1)) {  // clamped fragment: z_s ~ q(z_s | x_fixed)
    const float alpha_s = prm[5], sigma_s = prm[6];
    const float cntk = (float)max(n1 - n0, 1);
    float sk[3] = {0.f, 0.f, 0.f};
    for (int n = n0 + lane; n < n1; n += 32)
      for (int c = 0; c < 3; c++) sk[c] += noise2_x[n * 3 + c];
    for (int c = 0; c < 3; c++) sk[c] = warp_sum(sk[c]) / cntk;
    for (int n = n0 + lane; n < n1; n += 32)
      for (int k = 0; k < nf; k++) {
        const float nz = k < 3 ? noise2_x[n * 3 + k] - sk[k] : (noise2_h ? noise2_h[(size_t)n * d + (k - 3)] : 0.f);
        const float v = __fadd_rn(__fmul_rn(alpha_s, x_fixed[(size_t)n * nf + k]), __fmul_rn(sigma_s, nz));
        z[(size_t)n * nf + k] = (k >= 3 && h0) ? h0[(size_t)n * (nf - 3) + (k - 3)] : v;
      }
    return;
  }
  const bool bad = *nan_flag != 0;
  const uint32_t ctr = (uint32_t)reinterpret_cast<const int*>(prm)[4];
  const float cnt = (float)max(n1 - n0, 1);
  auto vload = [&](int n, int c) { return bad ? dyn_gauss(ctr, (uint32_t)(n * 3 + c)) : vel[n * 3 + c]; };
  // pass 1: segment sums of the velocity (and of the raw position noise)
  float sv[3] = {0.f, 0.f, 0.f}, sn[3] = {0.f, 0.f, 0.f};
  for (int n = n0 + lane; n < n1; n += 32)
    for (int c = 0; c < 3; c++) {
      sv[c] += vload(n, c);
      if (MODE >= 1) sn[c] += noise_x[n * 3 + c];
    }
  for (int c = 0; c < 3; c++) { sv[c] = warp_sum(sv[c]) / cnt; if (MODE >= 1) sn[c] = warp_sum(sn[c]) / cnt; }
  const float *w0 = cd.dw0[f], *b0 = cd.db0[f], *w1 = cd.dw1[f], *b1 = cd.db1[f];
  auto eps_of = [&](int n, float* e) {  // e[0..nf)
    for (int c = 0; c < 3; c++) e[c] = vload(n, c) - sv[c];
    const float* hh = h_out + (size_t)n * C;
    float hid[2 * DYN_MAX_D];
    for (int j = 0; j < 2 * d; j++) {
      float a = b0[j];
      for (int k = 0; k < emb; k++) a = fmaf(w0[j * emb + k], hh[k], a);
      hid[j] = silu(a);
    }
    for (int j = 0; j < d; j++) {
      float a = b1[j];
      for (int k = 0; k < 2 * d; k++) a = fmaf(w1[j * 2 * d + k], hid[k], a);
      e[3 + j] = a;
    }
  };
  if (MODE == 0) {
    for (int n = n0 + lane; n < n1; n += 32) {
      float e[3 + DYN_MAX_D];
      eps_of(n, e);
      for (int k = 0; k < nf; k++) eps_out[(size_t)n * nf + k] = e[k];
    }
    return;
  }
  const float alpha_ts = prm[1], coef = prm[2], sigma = prm[3];
  auto zs_of = [&](int n, float* o) {
    float e[3 + DYN_MAX_D];
    eps_of(n, e);
    for (int k = 0; k < nf; k++) {
      const float nz = k < 3 ? noise_x[n * 3 + k] - sn[k] : (noise_h ? noise_h[(size_t)n * d + (k - 3)] : 0.f);
      const float mu = __fsub_rn(__fdiv_rn(z[(size_t)n * nf + k], alpha_ts), __fmul_rn(e[k], coef));
      o[k] = __fadd_rn(mu, __fmul_rn(sigma, nz));
    }
  };
  // pass 2: segment mean of the new positions
  float sz[3] = {0.f, 0.f, 0.f};
  for (int n = n0 + lane; n < n1; n += 32) {
    float o[3 + DYN_MAX_D];
    zs_of(n, o);
    for (int c = 0; c < 3; c++) sz[c] += o[c];
  }
  for (int c = 0; c < 3; c++) sz[c] = warp_sum(sz[c]) / cnt;
  __syncwarp();
  // pass 3: write (each lane rewrites only the nodes it read)
  for (int n = n0 + lane; n < n1; n += 32) {
    float o[3 + DYN_MAX_D];
    zs_of(n, o);
    for (int c = 0; c < 3; c++) z[(size_t)n * nf + c] = o[c] - sz[c];
    for (int k = 3; k < nf; k++) z[(size_t)n * nf + k] = h0 ? h0[(size_t)n * (nf - 3) + (k - 3)] : o[k];
  }
}

// The RePaint jump-back z_s -> z_t (en_diffusion.py:1050-1074, sample_p_zt_given_zs) IN PLACE: z = alpha_ts z + sigma_ts eps with
// eps = the caller's raw draws (position noise CoM-free per segment, feature noise NULL = zero), then the centre of mass of
// the new positions is removed per segment.  One warp per segment, fixed summation order.
__global__ void k_dyn_jump(int S, int nf, int d, const int* __restrict__ seg_ptr, float alpha_ts, float sigma_ts,
                           float* __restrict__ z, const float* __restrict__ noise_x, const float* __restrict__ noise_h) {
  const int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (sgm >= S) return;
  const int n0 = seg_ptr[sgm], n1 = seg_ptr[sgm + 1];
  const float cnt = (float)max(n1 - n0, 1);
  float sn[3] = {0.f, 0.f, 0.f};
  for (int n = n0 + lane; n < n1; n += 32)
    for (int c = 0; c < 3; c++) sn[c] += noise_x[n * 3 + c];
  for (int c = 0; c < 3; c++) sn[c] = warp_sum(sn[c]) / cnt;
  auto zt_of = [&](int n, int k) {
    const float nz = k < 3 ? noise_x[n * 3 + k] - sn[k] : (noise_h ? noise_h[(size_t)n * d + (k - 3)] : 0.f);
    return __fadd_rn(__fmul_rn(alpha_ts, z[(size_t)n * nf + k]), __fmul_rn(sigma_ts, nz));
  };
  float sz[3] = {0.f, 0.f, 0.f};
  for (int n = n0 + lane; n < n1; n += 32)
    for (int c = 0; c < 3; c++) sz[c] += zt_of(n, c);
  for (int c = 0; c < 3; c++) sz[c] = warp_sum(sz[c]) / cnt;
  __syncwarp();
  for (int n = n0 + lane; n < n1; n += 32) {
    float o[3 + DYN_MAX_D];
    for (int k = 0; k < nf; k++) o[k] = zt_of(n, k);
    for (int k = 0; k < nf; k++) z[(size_t)n * nf + k] = k < 3 ? o[k] - sz[k] : o[k];
  }
}

}  // namespace oard
