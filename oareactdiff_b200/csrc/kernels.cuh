// Message-passing / geometry / node kernels of the object-aware LEFTNet forward (sm_100a).
// Citations: reference oa_reactdiff/model/leftnet.py (file:line in each kernel's comment).
//
// Conventions: edges are CSR-ordered by source (row_ptr/esrc/ecol), rev[e] is the position of the transposed edge,
// "channel kernels" run one block per node (or edge) with one thread per hidden channel h < H (blockDim = H rounded
// up to a warp multiple).  Aggregations at edge_index[1] (PyG 'target') are done as row sums over transposed edges,
// so every reduction is a fixed-order segmented sum: no atomics, bitwise reproducible.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace oard {

// ---------------------------------------------------------------------------------------------------------------
// (B) raw distance + cutoff/same-fragment edge mask.  leftnet.py:747-753
__global__ void k_edge_mask(int E, const int* __restrict__ esrc, const int* __restrict__ ecol,
                            const float* __restrict__ pos, const int64_t* __restrict__ sub, float cutoff,
                            uint8_t* __restrict__ mask, uint8_t* __restrict__ sub8) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  sub8[e] = (sub == nullptr || sub[e] > 0);  // same-fragment relation (an equivalence): groups of k_equi_tgt
  const int i = esrc[e], j = ecol[e];
  const float dx = pos[i * 3 + 0] - pos[j * 3 + 0], dy = pos[i * 3 + 1] - pos[j * 3 + 1],
              dz = pos[i * 3 + 2] - pos[j * 3 + 2];
  // same rounding sequence as torch: pow(2) (separately rounded), sequential sum, sqrt
  const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  mask[e] = (d < cutoff) && (sub == nullptr || sub[e] > 0);
}

// ---------------------------------------------------------------------------------------------------------------
// (C)+(D)+(J) per connected component of the unmasked graph (= one reaction): greedy group labelling with the
// reference's overwrite semantics (leftnet.py:707-722), per-group centre-of-mass removal (:760-761, :26-29) and the
// legacy node frame (:812-834).  CoM and node frame are evaluated in fp64: in exact arithmetic b = -a/(n-1) and
// y1 = a x b vanishes, so the reference's fp32 y1 is pure rounding noise; fp64 here tracks the fp64
// oracle instead of inventing different noise.
template <int MAXC>
__global__ void __launch_bounds__(128) k_group_frame(
    const int* __restrict__ comp_ptr, const int* __restrict__ comp_nodes, const int* __restrict__ node_local,
    const int* __restrict__ row_ptr, const int* __restrict__ ecol, const uint8_t* __restrict__ mask,
    const float* __restrict__ pos, float* __restrict__ pf, double* __restrict__ pf64, float* __restrict__ nodeframe,
    float* __restrict__ pos_prjt, int* __restrict__ owner, uint8_t* __restrict__ opener) {
  __shared__ int lab[MAXC];
  __shared__ double p[MAXC][3];
  __shared__ double q[MAXC][3];
  const int c = blockIdx.x, base = comp_ptr[c], nc = comp_ptr[c + 1] - base;
  const int* nodes = comp_nodes + base;
  for (int n = threadIdx.x; n < nc; n += blockDim.x) {
    lab[n] = -1;
    const int t = nodes[n];
    p[n][0] = pos[t * 3 + 0]; p[n][1] = pos[t * 3 + 1]; p[n][2] = pos[t * 3 + 2];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int a = 0; a < nc; a++) {  // centres in ascending global node id
      if (lab[a] < 0) {
        const int t = nodes[a];
        for (int e = row_ptr[t] + lane; e < row_ptr[t + 1]; e += 32)
          if (mask[e]) lab[node_local[ecol[e]]] = a;
        if (lane == 0) lab[a] = a;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < nc; n += blockDim.x) {
    const int own = lab[n];
    double sx = 0, sy = 0, sz = 0;
    int cnt = 0;
    for (int m = 0; m < nc; m++)
      if (lab[m] == own) { sx += p[m][0]; sy += p[m][1]; sz += p[m][2]; cnt++; }
    q[n][0] = p[n][0] - sx / cnt; q[n][1] = p[n][1] - sy / cnt; q[n][2] = p[n][2] - sz / cnt;
    const int t = nodes[n];
    pf[t * 3 + 0] = (float)q[n][0]; pf[t * 3 + 1] = (float)q[n][1]; pf[t * 3 + 2] = (float)q[n][2];
    pf64[t * 3 + 0] = q[n][0]; pf64[t * 3 + 1] = q[n][1]; pf64[t * 3 + 2] = q[n][2];
    owner[t] = nodes[own];
    opener[t] = (own == n);
  }
  __syncthreads();
  for (int n = threadIdx.x; n < nc; n += blockDim.x) {
    const int t = nodes[n];
    double b0 = 0, b1 = 0, b2 = 0;
    const int r0 = row_ptr[t], r1 = row_ptr[t + 1];
    for (int e = r0; e < r1; e++) {
      const int m = node_local[ecol[e]];
      b0 += q[m][0]; b1 += q[m][1]; b2 += q[m][2];
    }
    const double cnt = (r1 - r0) > 0 ? (double)(r1 - r0) : 1.0;
    b0 /= cnt; b1 /= cnt; b2 /= cnt;
    const double a0 = q[n][0], a1 = q[n][1], a2 = q[n][2];
    double x0 = a0 - b0, x1 = a1 - b1, x2 = a2 - b2;
    const double nx = sqrt(x0 * x0 + x1 * x1 + x2 * x2) + (double)OARD_EPS;
    x0 /= nx; x1 /= nx; x2 /= nx;
    double y0 = a1 * b2 - a2 * b1, y1 = a2 * b0 - a0 * b2, y2 = a0 * b1 - a1 * b0;
    const double ny = sqrt(y0 * y0 + y1 * y1 + y2 * y2) + (double)OARD_EPS;
    y0 /= ny; y1 /= ny; y2 /= ny;
    const double z0 = x1 * y2 - x2 * y1, z1 = x2 * y0 - x0 * y2, z2 = x0 * y1 - x1 * y0;
    float* nf = nodeframe + (size_t)t * 9;  // [xyz c][k]: k = 0:x1, 1:y1, 2:z1
    nf[0] = (float)x0; nf[1] = (float)y0; nf[2] = (float)z0;
    nf[3] = (float)x1; nf[4] = (float)y1; nf[5] = (float)z1;
    nf[6] = (float)x2; nf[7] = (float)y2; nf[8] = (float)z2;
    pos_prjt[t * 3 + 0] = (float)(a0 * x0 + a1 * x1 + a2 * x2);
    pos_prjt[t * 3 + 1] = (float)(a0 * y0 + a1 * y1 + a2 * y2);
    pos_prjt[t * 3 + 2] = (float)(a0 * z0 + a1 * z1 + a2 * z2);
  }
}

// group ids in the reference's numbering: rank of the opening centre among all opening centres (debug/parity only)
__global__ void k_group_ids(int N, const int* __restrict__ owner, const uint8_t* __restrict__ opener,
                            int* __restrict__ rank_tmp, int* __restrict__ group) {
  // single block
  __shared__ int sm[1025];
  const int T = blockDim.x, tid = threadIdx.x, per = (N + T - 1) / T;
  const int b = tid * per, e = min(N, b + per);
  int s = 0;
  for (int i = b; i < e; i++) s += opener[i];
  sm[tid + 1] = s;
  if (tid == 0) sm[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int i = 1; i <= T; i++) sm[i] += sm[i - 1];
  __syncthreads();
  int run = sm[tid];
  for (int i = b; i < e; i++) { rank_tmp[i] = run; run += opener[i]; }
  __syncthreads();
  __threadfence_block();
  for (int i = tid; i < N; i += T) group[i] = rank_tmp[owner[i]];
}

// ---------------------------------------------------------------------------------------------------------------
// (E) edge geometry on pos_frame, masked (leftnet.py:693-705,764-771,785): geo[e] = (u_x,u_y,u_z,dist), rb[e],
// ecross[e] = unit cross of pos_frame_i x pos_frame_j; also counts the row's active edges.  One warp per CSR row.
// Evaluated in fp64 on the fp64 frame positions: inside a group of two (or a collinear group) the centred positions are
// exactly (anti)parallel, so the cross product is pure rounding noise in fp32 — against the 1e-6 of its normalisation that
// noise is O(1) in the reference's fp32, while fp64 gives the clean ~0 of the fp64 oracle (same argument as the node frame).
__global__ void k_edge_geom(int N, const int* __restrict__ row_ptr, const int* __restrict__ ecol,
                            const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sub8,
                            const double* __restrict__ pf64, float cutoff, float4* __restrict__ geo,
                            float4* __restrict__ ecross, float* __restrict__ rb,
                            int* __restrict__ row_cnt, int* __restrict__ leader, int* __restrict__ glocal,
                            int* __restrict__ gsize) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= N) return;
  const double ax = pf64[t * 3 + 0], ay = pf64[t * 3 + 1], az = pf64[t * 3 + 2];
  int cnt = 0, lead = t, below = 0, gsz = 0;
  for (int e = row_ptr[t] + lane; e < row_ptr[t + 1]; e += 32) {
    if (sub8[e]) {  // same-group neighbour: the group's leader is its smallest node id, glocal = rank inside the group
      const int j = ecol[e];
      lead = min(lead, j);
      below += j < t;
      gsz++;
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), cr = g;
    float r = 1.0f;  // 0.5*(cos(0)+1)
    if (mask[e]) {
      const int j = ecol[e];
      const double bx = pf64[j * 3 + 0], by = pf64[j * 3 + 1], bz = pf64[j * 3 + 2];
      const double dx = ax - bx, dy = ay - by, dz = az - bz;
      const double d = sqrt(dx * dx + dy * dy + dz * dz);
      const double inv = 1.0 / (d + (double)OARD_EPS);
      g = make_float4((float)(dx * inv), (float)(dy * inv), (float)(dz * inv), (float)d);
      const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      const double cinv = 1.0 / (sqrt(cx * cx + cy * cy + cz * cz) + (double)OARD_EPS);
      cr = make_float4((float)(cx * cinv), (float)(cy * cinv), (float)(cz * cinv), 0.f);
      r = 0.5f * (cosf(g.w * (float)OARD_PI / cutoff) + 1.0f);
      cnt++;
    }
    geo[e] = g;
    ecross[e] = cr;
    rb[e] = r;
  }
  cnt = (int)warp_sum((float)cnt);
  below = (int)warp_sum((float)below);
  gsz = (int)warp_sum((float)gsz);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lead = min(lead, __shfl_xor_sync(0xffffffffu, lead, o));
  if (lane == 0) { row_cnt[t] = cnt; leader[t] = lead; glocal[t] = below; gsize[t] = gsz + 1; }
}

// exclusive scan of row_cnt -> row_act_ptr[N+1], n_act = total.  Single block.
__global__ void k_scan_rows(int N, const int* __restrict__ row_cnt, int* __restrict__ row_act_ptr,
                            int* __restrict__ n_act, const int* __restrict__ leader, const int* __restrict__ gsize,
                            int* __restrict__ lead_list, int2* __restrict__ lead_info, int* __restrict__ n_lead,
                            int* __restrict__ work_ctr, int n_ctr) {
  __shared__ int sm[1025];
  const int T = blockDim.x, tid = threadIdx.x, per = (N + T - 1) / T;
  const int b = min(N, tid * per), e = min(N, b + per);
  int s = 0;
  for (int i = b; i < e; i++) s += row_cnt[i];
  sm[tid + 1] = s;
  if (tid == 0) sm[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int i = 1; i <= T; i++) sm[i] += sm[i - 1];
  __syncthreads();
  int run = sm[tid];
  for (int i = b; i < e; i++) { row_act_ptr[i] = run; run += row_cnt[i]; }
  if (tid == T - 1) { row_act_ptr[N] = sm[T]; *n_act = sm[T]; }
  // ordered compaction of the group leaders (leader[t] == t) and reset of the per-layer work counters of k_equi_tgt
  __syncthreads();
  s = 0;
  for (int i = b; i < e; i++) s += leader[i] == i;
  sm[tid + 1] = s;
  __syncthreads();
  if (tid == 0)
    for (int i = 1; i <= T; i++) sm[i] += sm[i - 1];
  __syncthreads();
  run = sm[tid];
  for (int i = b; i < e; i++)
    if (leader[i] == i) lead_list[run++] = i;
  if (tid == T - 1) *n_lead = sm[T];
  if (tid < n_ctr) work_ctr[tid] = 0;
  // offsets of the groups' member lists (exclusive scan of the group sizes in leader order)
  __syncthreads();
  s = 0;
  for (int i = b; i < e; i++) s += leader[i] == i ? gsize[i] : 0;
  const int first = sm[tid];  // index of this thread's first leader in lead_list
  __syncthreads();
  sm[tid + 1] = s;
  if (tid == 0) sm[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int i = 1; i <= T; i++) sm[i] += sm[i - 1];
  __syncthreads();
  int off = sm[tid], li = first;
  for (int i = b; i < e; i++)
    if (leader[i] == i) { lead_info[li++] = make_int2(off, gsize[i]); off += gsize[i]; }  // (member-list offset, size)
}

// ordered compaction of active edges: act_idx[p] = e, act_pos[e] = p or -1.  One warp per row.
__global__ void k_compact(int N, const int* __restrict__ row_ptr, const uint8_t* __restrict__ mask,
                          const int* __restrict__ row_act_ptr, int* __restrict__ act_idx, int* __restrict__ act_pos,
                          int* __restrict__ act_pos_t) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= N) return;
  int base = row_act_ptr[t];
  const int r0 = row_ptr[t], r1 = row_ptr[t + 1];
  for (int e0 = r0; e0 < r1; e0 += 32) {
    const int e = e0 + lane;
    const bool a = e < r1 && mask[e];
    const unsigned bal = __ballot_sync(0xffffffffu, a);
    const int off = __popc(bal & ((1u << lane) - 1u));
    if (e < r1) {
      if (a) { act_idx[base + off] = e; act_pos[e] = base + off; }
      else { act_pos[e] = -1; act_pos_t[e] = -1; }
    }
    base += __popc(bal);
  }
}

// (F) radial basis on active edges (leftnet.py:63-69; mask == 1 there): rbf_act[p, r]
__global__ void k_rbf(const int* __restrict__ n_act, int cap, int R, const int* __restrict__ act_idx,
                      const float4* __restrict__ geo, const float* __restrict__ means, const float* __restrict__ betas,
                      float cutoff, float* __restrict__ rbf_act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = (int)(idx / R), r = (int)(idx % R);
  if (p >= min(*n_act, cap)) return;
  const float d = geo[act_idx[p]].w;
  float rbc = 0.5f * (cosf(d * (float)OARD_PI / cutoff) + 1.0f);
  rbc = d < cutoff ? rbc : 0.f;
  const float x = expf(-d) - means[r];
  rbf_act[idx] = rbc * expf(-betas[r] * x * x);
}

// constants of masked edges: f0 = radial_lin(0) (leftnet.py:784-786 with rbf = 0, rbounds = 1),
// c3 = lin3(0) (leftnet.py:798-805 with zero frame).  Single block, one thread per channel.
__global__ void k_masked_consts(int H, int Hq, const float* __restrict__ rl0_b, const float* __restrict__ rl2_w,
                                const float* __restrict__ rl2_b, const float* __restrict__ l3_b0,
                                const float* __restrict__ l3_w2, const float* __restrict__ l3_b2,
                                float* __restrict__ f0, float* __restrict__ c3) {
  extern __shared__ float t1[];
  const int h = threadIdx.x;
  if (h < H) t1[h] = silu(rl0_b[h]);
  __syncthreads();
  if (h < H) {
    float a = rl2_b[h];
    for (int k = 0; k < H; k++) a = fmaf(rl2_w[h * H + k], t1[k], a);
    f0[h] = a;
  }
  if (h == 0) {
    float a = l3_b2[0];
    for (int q = 0; q < Hq; q++) a = fmaf(l3_w2[q], silu(l3_b0[q]), a);
    *c3 = a;
  }
}

// (A)+(G prologue) z_emb = embedding(h); ne = LN0(neighbor_emb.embedding(h)).  leftnet.py:744, 82
__global__ void k_node_init(int H, int C, const float* __restrict__ hin, const float* __restrict__ w_emb,
                            const float* __restrict__ b_emb, const float* __restrict__ w_ne,
                            const float* __restrict__ b_ne, float* __restrict__ z_emb, float* __restrict__ ne) {
  __shared__ float sm[40];
  __shared__ float x[32];
  const int t = blockIdx.x, h = threadIdx.x;
  if (h < C) x[h] = hin[(size_t)t * C + h];
  __syncthreads();
  const bool ok = h < H;
  float z = 0.f, n = 0.f;
  if (ok) {
    z = b_emb[h]; n = b_ne[h];
    for (int c = 0; c < C; c++) { z = fmaf(w_emb[h * C + c], x[c], z); n = fmaf(w_ne[h * C + c], x[c], n); }
    z_emb[(size_t)t * H + h] = z;
  }
  float mean, rstd;
  block_ln_stats(n, ok, H, sm, mean, rstd);
  if (ok) ne[(size_t)t * H + h] = (n - mean) * rstd;
}

// (G) NeighborEmb: s_t = z_emb_t + sum_{e: tgt = t} f_e * ne[src]   (leftnet.py:81-89), all edges incl. masked (f = f0)
__global__ void k_neighbor(int H, const int* __restrict__ row_ptr, const int* __restrict__ ecol,
                           const int* __restrict__ rev, const int* __restrict__ act_pos,
                           const float* __restrict__ f_act, const float* __restrict__ f0,
                           const float* __restrict__ z_emb, const float* __restrict__ ne, float* __restrict__ s) {
  const int t = blockIdx.x, h = threadIdx.x;
  if (h >= H) return;
  float acc = z_emb[(size_t)t * H + h];
  const float f0h = f0[h];
  for (int e = row_ptr[t]; e < row_ptr[t + 1]; e++) {
    const int p = act_pos[rev[e]];
    const float fv = p >= 0 ? f_act[(size_t)p * H + h] : f0h;
    acc = fmaf(fv, ne[(size_t)ecol[e] * H + h], acc);
  }
  s[(size_t)t * H + h] = acc;
}

// row-wise LayerNorm variants.  x: [rows, H] (ldx).  y = act(LN(x (+ add)) * gamma + beta) -> out (ldo)
__global__ void k_layernorm(int H, const float* __restrict__ x, int ldx, const float* __restrict__ add,
                            const float* __restrict__ gamma, const float* __restrict__ beta, int act_silu,
                            float* __restrict__ out, int ldo) {
  __shared__ float sm[40];
  const int t = blockIdx.x, h = threadIdx.x;
  const bool ok = h < H;
  float v = 0.f;
  if (ok) {
    v = x[(size_t)t * ldx + h];
    if (add) v += add[(size_t)t * H + h];
  }
  float mean, rstd;
  block_ln_stats(v, ok, H, sm, mean, rstd);
  if (ok) {
    float y = (v - mean) * rstd;
    if (gamma) y = y * gamma[h] + beta[h];
    if (act_silu) y = silu(y);
    out[(size_t)t * ldo + h] = y;
  }
}

// Warp-per-row LayerNorm (H <= 256): no block barriers, 8 rows per 256-thread block.  Same arithmetic as k_layernorm.
__global__ void __launch_bounds__(256) k_layernorm_w(int rows, int H, const float* __restrict__ x, int ldx,
                                                      const float* __restrict__ add, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int act_silu,
                                                      float* __restrict__ out, int ldo) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= rows) return;
  float v[8];
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int h = lane + 32 * q;
    v[q] = 0.f;
    if (h < H) {
      v[q] = x[(size_t)t * ldx + h];
      if (add) v[q] += add[(size_t)t * H + h];
    }
    sum += v[q];
  }
  const float mean = warp_sum(sum) / (float)H;
  float var = 0.f;
#pragma unroll
  for (int q = 0; q < 8; q++) { const float d = (lane + 32 * q < H) ? v[q] - mean : 0.f; var += d * d; }
  const float rstd = rsqrtf(warp_sum(var) / (float)H + 1e-5f);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int h = lane + 32 * q;
    if (h < H) {
      float y = (v[q] - mean) * rstd;
      if (gamma) y = y * gamma[h] + beta[h];
      if (act_silu) y = silu(y);
      out[(size_t)t * ldo + h] = y;
    }
  }
}

// (H) CFConvS2V aggregation: NE1_t[c,:] = sum_{e: tgt = t, active} f_e * u_e[c] * q[src]   (leftnet.py:116-125)
__global__ void k_s2v(int H, const int* __restrict__ row_ptr, const int* __restrict__ ecol,
                      const int* __restrict__ rev, const int* __restrict__ act_pos, const float* __restrict__ f_act,
                      const float4* __restrict__ geo, const float* __restrict__ q, float* __restrict__ NE1) {
  const int t = blockIdx.x, h = threadIdx.x;
  if (h >= H) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int e = row_ptr[t]; e < row_ptr[t + 1]; e++) {
    const int r = rev[e], p = act_pos[r];
    if (p < 0) continue;
    const float fq = f_act[(size_t)p * H + h] * q[(size_t)ecol[e] * H + h];
    const float4 g = geo[r];
    a0 = fmaf(fq, g.x, a0); a1 = fmaf(fq, g.y, a1); a2 = fmaf(fq, g.z, a2);
  }
  float* o = NE1 + (size_t)t * 3 * H;
  o[h] = a0; o[H + h] = a1; o[2 * H + h] = a2;
}

// GCL attention gate + mean aggregation at the edge source (leftnet.py:169-183, util_funcs.py:27-45).
// One block per node; warps take the row's edges round-robin.  m2 is NOT rewritten: the gate is a per-edge scalar, so
// edge_out_trans applies it to its accumulator rows (W (att m) = att (W m): GemmArgs::prescale) and this kernel only
// stores att[e] — half the HBM traffic of the in-place version.
__global__ void k_att_agg(int H, const int* __restrict__ row_ptr, const float* __restrict__ m2,
                          const float* __restrict__ w_att, const float* __restrict__ b_att, float* __restrict__ att_out,
                          float* __restrict__ xa, int ldxa) {
  extern __shared__ float part[];  // [nw][H]
  const int t = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int r0 = row_ptr[t], r1 = row_ptr[t + 1];
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = 0.f;
  const float ba = b_att[0];
  for (int e = r0 + w; e < r1; e += nw) {
    const float* row = m2 + (size_t)e * H;
    float v[8], d = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int h = lane + 32 * k;
      v[k] = h < H ? row[h] : 0.f;
      d = fmaf(v[k], h < H ? w_att[h] : 0.f, d);
    }
    d = warp_sum(d);
    const float att = silu(d + ba);
    if (lane == 0) att_out[e] = att;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int h = lane + 32 * k;
      if (h < H) acc[k] += v[k] * att;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int h = lane + 32 * k;
    if (h < H) part[w * H + h] = acc[k];
  }
  __syncthreads();
  const int h = threadIdx.x;
  if (h < H) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ww++) s += part[ww * H + h];
    const int cnt = r1 - r0;
    xa[(size_t)t * ldxa + H + h] = s / (float)(cnt > 0 ? cnt : 1);
  }
}

// Per-step lists over the compact active edges: for p = (t -> a), the neighbour a and the geometry of p (the transposed
// edge's unit vector is its exact negation), and per EDGE e the compact position of its transposed edge, act_pos_t[e]: the
// row where edge_out puts e's compact copy, so that the compact rows are in TARGET order (rows [row_act_ptr[t],
// row_act_ptr[t+1]) = the edges arriving at t) and the target-side aggregation walks one contiguous block.  The distance
// and the same-fragment relation are symmetric, so e is active iff its transposed edge is.
__global__ void k_act_lists(const int* __restrict__ n_act, int cap, const int* __restrict__ act_idx,
                            const int* __restrict__ act_pos, const int* __restrict__ rev, const int* __restrict__ ecol,
                            const float4* __restrict__ geo, const float4* __restrict__ ecross,
                            const int* __restrict__ glocal, int* __restrict__ act_pos_t,
                            int* __restrict__ act_col, float4* __restrict__ act_geo, float4* __restrict__ act_cross,
                            int2* __restrict__ act_rec) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= min(*n_act, cap)) return;
  const int e = act_idx[p];
  const int tr = act_pos[rev[e]], a = ecol[e];
  act_pos_t[e] = tr;
  act_col[p] = a;
  act_geo[p] = geo[e];
  act_cross[p] = ecross[e];
  act_rec[p] = make_int2(tr, glocal[a]);  // k_equi_tgt: (source-ordered position of the message edge, source's rank inside its group)
}

// Member lists of the groups, once per forward: gm_node[off + i] = i-th member (ascending), gm_rap[off + i] = its range of
// compact active edges.  One warp per group.
__global__ void k_group_members(const int* __restrict__ n_lead, const int* __restrict__ lead_list,
                                const int2* __restrict__ lead_info, const int* __restrict__ row_ptr,
                                const int* __restrict__ ecol, const uint8_t* __restrict__ sub8,
                                const int* __restrict__ row_act_ptr, int* __restrict__ gm_node, int2* __restrict__ gm_rap) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gi >= *n_lead) return;
  const int b = lead_list[gi], off = lead_info[gi].x;
  if (lane == 0) { gm_node[off] = b; gm_rap[off] = make_int2(row_act_ptr[b], row_act_ptr[b + 1]); }
  int cnt = 1;
  const int r0 = row_ptr[b], r1 = row_ptr[b + 1];
  for (int e0 = r0; e0 < r1; e0 += 32) {
    const int e = e0 + lane;
    const bool a = e < r1 && sub8[e];
    const unsigned bal = __ballot_sync(0xffffffffu, a);
    if (a) {
      const int j = ecol[e], o = off + cnt + __popc(bal & ((1u << lane) - 1u));
      gm_node[o] = j;
      gm_rap[o] = make_int2(row_act_ptr[j], row_act_ptr[j + 1]);
    }
    cnt += __popc(bal);
  }
}

// EquiMessage message + aggregation at the edge target + residual (leftnet.py:263-284, 857-859): the message-passing
// kernel proper (HBM-bound: streams G[E_act, 3H] once).  One block per target node t; NG groups of 64 threads take the
// node's active incoming edges round-robin (one thread = 4 channels, float4 loads, two edges in flight per group), the
// NG partial sums are combined through shared memory in a fixed order (bitwise reproducible), then
// s = (s + dx)/sqrt2, vec_out = vec_in + dvec.
template <int NG>
__global__ void __launch_bounds__(NG * 64, 1024 / (NG * 64)) k_equi_reduce(
    int H, int reflect, const int* __restrict__ row_act_ptr, const int* __restrict__ act_col, const float4* __restrict__ act_geo, const float4* __restrict__ act_cross,
    const float* __restrict__ G, const float* __restrict__ X, const float* __restrict__ vec_in,
    float* __restrict__ vec_out, float* __restrict__ s) {
  extern __shared__ __align__(16) float4 eq_part[];  // [NG][4][H/4]
  const int t = blockIdx.x, grp = threadIdx.x >> 6, h = (threadIdx.x & 63) * 4, H4 = H / 4;
  const bool on = h < H;
  const float inv_sqrt_3 = 0.57735026918962576f, inv_sqrt_h = rsqrtf((float)H), inv_sqrt_2 = 0.70710678118654752f;
  auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
  float4 dx = make_float4(0.f, 0.f, 0.f, 0.f), d0 = dx, d1 = dx, d2 = dx;
  const int p0 = row_act_ptr[t], p1 = row_act_ptr[t + 1];
  if (on) {
    const float* Xt = X + (size_t)t * 3 * H;
    const float4 x0 = ld4(Xt + h), x1 = ld4(Xt + H + h), x2 = ld4(Xt + 2 * H + h);
    auto edge = [&](int p) {
      const int a = act_col[p];
      const float4 gm = act_geo[p];  // geometry of (t -> a); the message edge (a -> t) has the negated unit vector
      const float ux = -gm.x, uy = -gm.y, uz = -gm.z;
      const float* g = G + (size_t)p * 3 * H;  // compact rows are in target order: row p = message (a -> t) of edge p = (t -> a)
      const float* Xa = X + (size_t)a * 3 * H;
      const float* va = vec_in + (size_t)a * 3 * H;
      const float4 g0 = ld4(g + h), g1 = ld4(g + H + h), g2 = ld4(g + 2 * H + h);
      const float4 a0 = ld4(Xa + h), a1 = ld4(Xa + H + h), a2 = ld4(Xa + 2 * H + h);
      const float4 v0 = ld4(va + h), v1 = ld4(va + H + h), v2 = ld4(va + 2 * H + h);
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (!reflect) {  // + x * edge_cross (leftnet.py:268-269): unit cross of pos_frame_a x pos_frame_t = -(t x a) of edge p
        const float4 cr = act_cross[p];
        cx = -cr.x; cy = -cr.y; cz = -cr.z;
      }
#define OARD_EQ(c)                                                                         \
      {                                                                                    \
        const float al = (a0.c + x0.c) * g0.c;                                             \
        const float be = (a1.c + x1.c) * g1.c * inv_sqrt_3;                                \
        const float ga = (a2.c + x2.c) * g2.c;                                             \
        float m0 = fmaf(v0.c, be, ga * ux), m1 = fmaf(v1.c, be, ga * uy), m2 = fmaf(v2.c, be, ga * uz); \
        if (!reflect) { m0 = fmaf(al, cx, m0); m1 = fmaf(al, cy, m1); m2 = fmaf(al, cz, m2); }          \
        dx.c += al;                                                                        \
        d0.c = fmaf(m0, inv_sqrt_h, d0.c); d1.c = fmaf(m1, inv_sqrt_h, d1.c); d2.c = fmaf(m2, inv_sqrt_h, d2.c); \
      }
      OARD_EQ(x) OARD_EQ(y) OARD_EQ(z) OARD_EQ(w)
#undef OARD_EQ
    };
    for (int p = p0 + grp; p < p1; p += NG) edge(p);
    float4* mine = eq_part + (size_t)grp * 4 * H4 + (h >> 2);
    mine[0] = dx; mine[H4] = d0; mine[2 * H4] = d1; mine[3 * H4] = d2;
  }
  __syncthreads();
  if (grp == 0 && on) {
    auto add4 = [](float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; };
    for (int gq = 1; gq < NG; gq++) {
      const float4* o = eq_part + (size_t)gq * 4 * H4 + (h >> 2);
      add4(dx, o[0]); add4(d0, o[H4]); add4(d1, o[2 * H4]); add4(d2, o[3 * H4]);
    }
    const size_t o = (size_t)t * 3 * H + h;
    float4 sv = ld4(s + (size_t)t * H + h);
    sv.x = (sv.x + dx.x) * inv_sqrt_2; sv.y = (sv.y + dx.y) * inv_sqrt_2; sv.z = (sv.z + dx.z) * inv_sqrt_2; sv.w = (sv.w + dx.w) * inv_sqrt_2;
    *reinterpret_cast<float4*>(s + (size_t)t * H + h) = sv;
    const float4 w0 = ld4(vec_in + o), w1 = ld4(vec_in + o + H), w2 = ld4(vec_in + o + 2 * H);
    *reinterpret_cast<float4*>(vec_out + o) = make_float4(w0.x + d0.x, w0.y + d0.y, w0.z + d0.z, w0.w + d0.w);
    *reinterpret_cast<float4*>(vec_out + o + H) = make_float4(w1.x + d1.x, w1.y + d1.y, w1.z + d1.z, w1.w + d1.w);
    *reinterpret_cast<float4*>(vec_out + o + 2 * H) = make_float4(w2.x + d2.x, w2.y + d2.y, w2.z + d2.z, w2.w + d2.w);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// EquiMessage message + aggregation at the target, round-2 form (reflect_equiv only; same arithmetic as k_equi_reduce).
// The compact rows of the active edges are kept in TARGET order (edge_out scatters the row of edge e to the compact
// position of its transposed edge), so the G rows of all messages arriving at target t are the contiguous block
// [row_act_ptr[t], row_act_ptr[t+1]) and G row p belongs to the message (a -> t) of the compact edge p = (t -> a).
// Work item = (group, CH-channel slice), one item per CTA turn (atomic counter); the group's X / vec rows of the slice are
// staged once in shared memory.  Thread = (target slot, float4 column): it walks the target's block of G rows with the next
// row's three 16-byte loads in flight and accumulates in registers — no cross-thread reduction, no shuffles, fixed edge
// order (bitwise reproducible).  Latency is hidden by the co-resident CTAs (4 per SM), not by an in-CTA pipeline.
// (Edge slots, deeper register staging, a loader warp and L2 prefetches were all measured slower: profiles/r2_experiments.md.)
// 16-byte read-only load with an L2 eviction-priority policy (createpolicy): the G rows are written by dir_proj2 with evict_last
// so that they survive in L2 until this kernel, and are read here with evict_first so that they leave it afterwards.
__device__ __forceinline__ float4 ldg4_policy(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int CH>
inline size_t et_smem_bytes(int gmax) {
  return (size_t)gmax * 2 * 3 * CH * 4 + (size_t)((gmax + 3) & ~3) * (4 + 8);
}
constexpr int ET_THREADS = 192;
template <int CH>
__global__ void __launch_bounds__(ET_THREADS, 4) k_equi_tgt(
    int H, int NS, int gmax, const int* __restrict__ n_lead, const int2* __restrict__ lead_info,
    int* __restrict__ work_ctr, const int* __restrict__ gm_node, const int2* __restrict__ gm_rap,
    const int2* __restrict__ act_rec, const float4* __restrict__ act_geo, const float* __restrict__ G,
    const float* __restrict__ X, const float* __restrict__ vec_in, float* __restrict__ vec_out, float* __restrict__ s) {
  constexpr int Q = CH / 4;              // float4 columns per slice
  constexpr int NSLOT = ET_THREADS / Q;  // targets in flight per CTA
  extern __shared__ __align__(16) float4 et_sm[];
  float4* Xs = et_sm;                         // [gmax][3][Q]
  float4* Vs = et_sm + (size_t)gmax * 3 * Q;  // [gmax][3][Q]
  const int gpad = (gmax + 3) & ~3;
  int2* mem_rap = reinterpret_cast<int2*>(Vs + (size_t)gmax * 3 * Q);  // [gpad]
  int* mem_node = reinterpret_cast<int*>(mem_rap + gpad);               // [gpad]
  __shared__ int w_sm;
  const int tid = threadIdx.x;
  const int slot = tid / Q, q = tid - slot * Q;
  const float inv_sqrt_3 = 0.57735026918962576f, inv_sqrt_h = rsqrtf((float)H), inv_sqrt_2 = 0.70710678118654752f;
  auto ld4 = [](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
  const uint64_t polG = l2_evict_first_policy();
  const int total = *n_lead * NS;
  const size_t H3 = (size_t)3 * H;
  for (;;) {
    if (tid == 0) w_sm = atomicAdd(work_ctr, 1);
    __syncthreads();
    const int w = w_sm;
    if (w >= total) break;
    const int gi = w / NS, h0 = (w - gi * NS) * CH;
    const int2 inf = lead_info[gi];
    const int gs = inf.y;
    for (int i = tid; i < gs; i += ET_THREADS) { mem_node[i] = gm_node[inf.x + i]; mem_rap[i] = gm_rap[inf.x + i]; }
    for (int i = tid; i < gs * 3 * Q; i += ET_THREADS) {
      const int m = i / (3 * Q), r = i - m * 3 * Q, c = r / Q, qq = r - c * Q;
      const size_t o = (size_t)gm_node[inf.x + m] * H3 + (size_t)c * H + h0 + 4 * qq;
      Xs[i] = ld4(X + o);
      Vs[i] = ld4(vec_in + o);
    }
    __syncthreads();
    if (slot < NSLOT) {
      for (int tl = slot; tl < gs; tl += NSLOT) {
        const int2 rap = mem_rap[tl];
        const int t = mem_node[tl];
        const float4 x0 = Xs[(tl * 3 + 0) * Q + q], x1 = Xs[(tl * 3 + 1) * Q + q], x2 = Xs[(tl * 3 + 2) * Q + q];
        float4 dx = make_float4(0.f, 0.f, 0.f, 0.f), d0 = dx, d1 = dx, d2 = dx;
        const float* gp = G + (size_t)rap.x * H3 + h0 + 4 * q;
        float4 g0, g1, g2, gm;
        int al = 0;
        if (rap.x < rap.y) {
          g0 = ldg4_policy(gp, polG); g1 = ldg4_policy(gp + H, polG); g2 = ldg4_policy(gp + 2 * H, polG);
          al = act_rec[rap.x].y; gm = act_geo[rap.x];
        }
        for (int p = rap.x; p < rap.y; p++) {
          const float4 c0 = g0, c1 = g1, c2 = g2, cm = gm;
          const int ca = al;
          if (p + 1 < rap.y) {  // next row in flight while this one is evaluated
            gp += H3;
            g0 = ldg4_policy(gp, polG); g1 = ldg4_policy(gp + H, polG); g2 = ldg4_policy(gp + 2 * H, polG);
            al = act_rec[p + 1].y; gm = act_geo[p + 1];
          }
          const float ux = -cm.x, uy = -cm.y, uz = -cm.z;  // the message edge (a -> t) has the negated unit vector of (t -> a)
          const float4 a0 = Xs[(ca * 3 + 0) * Q + q], a1 = Xs[(ca * 3 + 1) * Q + q], a2 = Xs[(ca * 3 + 2) * Q + q];
          const float4 v0 = Vs[(ca * 3 + 0) * Q + q], v1 = Vs[(ca * 3 + 1) * Q + q], v2 = Vs[(ca * 3 + 2) * Q + q];
#define OARD_EQT(c)                                                                          \
          {                                                                                  \
            const float al_ = (a0.c + x0.c) * c0.c;                                          \
            const float be = (a1.c + x1.c) * c1.c * inv_sqrt_3;                              \
            const float ga = (a2.c + x2.c) * c2.c;                                           \
            const float m0 = fmaf(v0.c, be, ga * ux), m1 = fmaf(v1.c, be, ga * uy), m2 = fmaf(v2.c, be, ga * uz); \
            dx.c += al_;                                                                     \
            d0.c = fmaf(m0, inv_sqrt_h, d0.c); d1.c = fmaf(m1, inv_sqrt_h, d1.c); d2.c = fmaf(m2, inv_sqrt_h, d2.c); \
          }
          OARD_EQT(x) OARD_EQT(y) OARD_EQT(z) OARD_EQT(w)
#undef OARD_EQT
        }
        float* sp = s + (size_t)t * H + h0 + 4 * q;
        float4 sv = *reinterpret_cast<const float4*>(sp);
        sv.x = (sv.x + dx.x) * inv_sqrt_2; sv.y = (sv.y + dx.y) * inv_sqrt_2; sv.z = (sv.z + dx.z) * inv_sqrt_2; sv.w = (sv.w + dx.w) * inv_sqrt_2;
        *reinterpret_cast<float4*>(sp) = sv;
        const size_t o = (size_t)t * H3 + h0 + 4 * q;
        const float4 w0 = Vs[(tl * 3 + 0) * Q + q], w1 = Vs[(tl * 3 + 1) * Q + q], w2 = Vs[(tl * 3 + 2) * Q + q];
        *reinterpret_cast<float4*>(vec_out + o) = make_float4(w0.x + d0.x, w0.y + d0.y, w0.z + d0.z, w0.w + d0.w);
        *reinterpret_cast<float4*>(vec_out + o + H) = make_float4(w1.x + d1.x, w1.y + d1.y, w1.z + d1.z, w1.w + d1.w);
        *reinterpret_cast<float4*>(vec_out + o + 2 * H) = make_float4(w2.x + d2.x, w2.y + d2.y, w2.z + d2.z, w2.w + d2.w);
      }
    }
    __syncthreads();  // tiles dead before the next item stages
  }
}

// lin3 weights of the EquiUpdate (3 -> 48 -> 8 -> 1, leftnet.py:305-312) / of the edge scalarisation (3 -> H/4 -> 1, :626-630)
// passed BY VALUE: kernel parameters live in the constant bank, so with fully unrolled loops every weight is an immediate
// constant operand of its FFMA — no load instruction and no shared-memory traffic.  (The shared-memory float4-broadcast
// versions were bound by the LDS return bandwidth: 3 LDS.128 = 1.5 KB per warp per hidden unit, 0.37-0.45 IPC.)
// Every weight is stored as the pair (w, w): one 64-bit constant operand of an FFMA2 that evaluates two items at once.
struct Lin3U { float2 w0[48 * 3], b0[48], w2[8 * 48], b2[8]; float w4[8], b4; };
struct Lin3E { float2 w0[64 * 3], b0[64], w2[64]; float b2; int hq; };

constexpr int US_NT = 2;  // nodes per thread = the two halves of every packed operation
__global__ void __launch_bounds__(256) k_upd_scalar_c(int N, int H, int reflect, const float* __restrict__ VP,
                                                       const float* __restrict__ nodeframe, const float* __restrict__ s,
                                                       const __grid_constant__ Lin3U W, float* __restrict__ sx,
                                                       float* __restrict__ vd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int tg = idx / H, h = idx - tg * H;
  if (tg * US_NT >= N) return;
  const float inv_sqrt_h = rsqrtf((float)H);
  float s0[2], s1[2], s2[2], vdv[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int t = min(tg * 2 + j, N - 1);
    const float* nf = nodeframe + (size_t)t * 9;
    const float* vp = VP + (size_t)t * 3 * 2 * H;
    const float v10 = vp[h], v11 = vp[2 * H + h], v12 = vp[4 * H + h];
    const float v20 = vp[H + h], v21 = vp[3 * H + h], v22 = vp[5 * H + h];
    s0[j] = v10 * nf[0] + v11 * nf[3] + v12 * nf[6];
    s1[j] = v10 * nf[1] + v11 * nf[4] + v12 * nf[7];
    s2[j] = v10 * nf[2] + v11 * nf[5] + v12 * nf[8];
    if (reflect) s1[j] = fabsf(s1[j]);
    vdv[j] = (v10 * v20 + v11 * v21 + v12 * v22) * inv_sqrt_h;
  }
  const f32x2 p0 = pk2(s0[0], s0[1]), p1 = pk2(s1[0], s1[1]), p2 = pk2(s2[0], s2[1]);
  f32x2 a[8];
#pragma unroll
  for (int q = 0; q < 8; q++) a[q] = ld2(W.b2[q]);
#pragma unroll
  for (int k = 0; k < 48; k += 4) {
    f32x2 u[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      u[i] = fma2(ld2(W.w0[(k + i) * 3]), p0, fma2(ld2(W.w0[(k + i) * 3 + 1]), p1, fma2(ld2(W.w0[(k + i) * 3 + 2]), p2, ld2(W.b0[k + i]))));
    silu4_shared_rcp2(u[0], u[1], u[2], u[3]);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int q = 0; q < 8; q++) a[q] = fma2(ld2(W.w2[q * 48 + k + i]), u[i], a[q]);
  }
  silu4_shared_rcp2(a[0], a[1], a[2], a[3]);
  silu4_shared_rcp2(a[4], a[5], a[6], a[7]);
  float o0 = W.b4, o1 = W.b4;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    float x, y;
    upk2(a[q], x, y);
    o0 = fmaf(W.w4[q], x, o0); o1 = fmaf(W.w4[q], y, o1);
  }
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int t = tg * 2 + j;
    if (t < N) {
      sx[(size_t)t * 2 * H + h] = s[(size_t)t * H + h];
      sx[(size_t)t * 2 * H + H + h] = j ? o1 : o0;
      vd[(size_t)t * H + h] = vdv[j];
    }
  }
}

// EquiUpdate scalarisation on the node frame + lin3 (3->48->8->1) + vec_dot  (leftnet.py:326-336).
// One thread per (node, channel).  The lin3 weights are broadcast from shared memory as float4 ((w0,w1,w2,b0) per hidden
// unit, the 8 second-layer weights of a hidden unit as two float4): 3 LDS.128 per hidden unit instead of 12 LDS.32 (the
// first version was bound by the shared-memory pipe), and four SiLUs share one reciprocal.
__global__ void k_upd_scalar(int N, int H, int reflect, const float* __restrict__ VP, const float* __restrict__ nodeframe,
                             const float* __restrict__ s, const float* __restrict__ w0, const float* __restrict__ b0,
                             const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w4,
                             const float* __restrict__ b4, float* __restrict__ sx, float* __restrict__ vd) {
  __shared__ __align__(16) float4 Wq[48];
  __shared__ __align__(16) float4 W2t[48][2];
  __shared__ float B2[8], W4[8];
  for (int k = threadIdx.x; k < 48; k += blockDim.x) {
    Wq[k] = make_float4(w0[k * 3], w0[k * 3 + 1], w0[k * 3 + 2], b0[k]);
    W2t[k][0] = make_float4(w2[k], w2[48 + k], w2[96 + k], w2[144 + k]);
    W2t[k][1] = make_float4(w2[192 + k], w2[240 + k], w2[288 + k], w2[336 + k]);
  }
  for (int k = threadIdx.x; k < 8; k += blockDim.x) { B2[k] = b2[k]; W4[k] = w4[k]; }
  __syncthreads();
  const int h = threadIdx.x;
  if (h >= H) return;
  const float bias4 = b4[0], inv_sqrt_h = rsqrtf((float)H);
  for (int t = blockIdx.x; t < N; t += gridDim.x) {  // persistent: the weights are staged once per block
  const float* nf = nodeframe + (size_t)t * 9;
  const float* vp = VP + (size_t)t * 3 * 2 * H;
  const float v10 = vp[h], v11 = vp[2 * H + h], v12 = vp[4 * H + h];
  const float v20 = vp[H + h], v21 = vp[3 * H + h], v22 = vp[5 * H + h];
  const float s0 = v10 * nf[0] + v11 * nf[3] + v12 * nf[6];
  float s1 = v10 * nf[1] + v11 * nf[4] + v12 * nf[7];
  const float s2 = v10 * nf[2] + v11 * nf[5] + v12 * nf[8];
  if (reflect) s1 = fabsf(s1);
  float a[8];
#pragma unroll
  for (int q = 0; q < 8; q++) a[q] = B2[q];
#pragma unroll 2
  for (int k = 0; k < 48; k += 4) {
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 q = Wq[k + i];
      u[i] = fmaf(q.x, s0, fmaf(q.y, s1, fmaf(q.z, s2, q.w)));
    }
    silu4_shared_rcp(u[0], u[1], u[2], u[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 wa = W2t[k + i][0], wb = W2t[k + i][1];
      a[0] = fmaf(wa.x, u[i], a[0]); a[1] = fmaf(wa.y, u[i], a[1]); a[2] = fmaf(wa.z, u[i], a[2]); a[3] = fmaf(wa.w, u[i], a[3]);
      a[4] = fmaf(wb.x, u[i], a[4]); a[5] = fmaf(wb.y, u[i], a[5]); a[6] = fmaf(wb.z, u[i], a[6]); a[7] = fmaf(wb.w, u[i], a[7]);
    }
  }
  silu4_shared_rcp(a[0], a[1], a[2], a[3]);
  silu4_shared_rcp(a[4], a[5], a[6], a[7]);
  float out = bias4;
#pragma unroll
  for (int q = 0; q < 8; q++) out = fmaf(W4[q], a[q], out);
  sx[(size_t)t * 2 * H + h] = s[(size_t)t * H + h];
  sx[(size_t)t * 2 * H + H + h] = out;
  vd[(size_t)t * H + h] = (v10 * v20 + v11 * v21 + v12 * v22) * inv_sqrt_h;
  }
}

// EquiUpdate apply: s += (xv1 + xv2 + vec_dot)/sqrt2 ; vec += xv3 * vec2   (leftnet.py:338-346, 863-864)
__global__ void k_upd_apply(int H, const float* __restrict__ XV, const float* __restrict__ VP,
                            const float* __restrict__ vd, float* __restrict__ s, float* __restrict__ vec) {
  const int t = blockIdx.x, h = threadIdx.x;
  if (h >= H) return;
  const float* xv = XV + (size_t)t * 3 * H;
  const float* vp = VP + (size_t)t * 3 * 2 * H;
  s[(size_t)t * H + h] += (xv[h] + xv[H + h] + vd[(size_t)t * H + h]) * 0.70710678118654752f;
  const float x3 = xv[2 * H + h];
  float* v = vec + (size_t)t * 3 * H;
  v[h] = fmaf(x3, vp[H + h], v[h]);
  v[H + h] = fmaf(x3, vp[3 * H + h], v[H + h]);
  v[2 * H + h] = fmaf(x3, vp[5 * H + h], v[2 * H + h]);
}

// GatedEquivariantBlock prologue: sn = [ s | ||vec1_proj(vec)||_xyz ]   (leftnet.py:567-570)
__global__ void k_out_norm(int H, const float* __restrict__ O1, const float* __restrict__ s, float* __restrict__ sn) {
  const int t = blockIdx.x, h = threadIdx.x;
  if (h >= H) return;
  const float* o = O1 + (size_t)t * 3 * H;
  const float a = o[h], b = o[H + h], c = o[2 * H + h];
  sn[(size_t)t * 2 * H + h] = s[(size_t)t * H + h];
  sn[(size_t)t * 2 * H + H + h] = sqrtf(a * a + b * b + c * c);
}

// Output head: gate = update_net.2(t)[1]; dpos = gate * vec2_proj(vec); h_out = embedding_out(s)  (:571-572, 878-887)
__global__ void k_final(int H, int C, const float* __restrict__ tu, const float* __restrict__ w_u2,
                        const float* __restrict__ b_u2, const float* __restrict__ vec, const float* __restrict__ w_o2,
                        const float* __restrict__ s, const float* __restrict__ w_eout, const float* __restrict__ b_eout,
                        float* __restrict__ dpos, float* __restrict__ h_out) {
  __shared__ float sm[40];
  const int t = blockIdx.x, h = threadIdx.x;
  const bool ok = h < H;
  const float tv = ok ? tu[(size_t)t * H + h] : 0.f;
  const float gate = block_sum(ok ? tv * w_u2[H + h] : 0.f, sm) + b_u2[1];
  const float* v = vec + (size_t)t * 3 * H;
  const float wo = ok ? w_o2[h] : 0.f;
  for (int c = 0; c < 3; c++) {
    const float d = block_sum(ok ? v[c * H + h] * wo : 0.f, sm);
    if (h == 0) dpos[t * 3 + c] = gate * d;
  }
  const float sv = ok ? s[(size_t)t * H + h] : 0.f;
  for (int c = 0; c < C; c++) {
    const float d = block_sum(ok ? sv * w_eout[c * H + h] : 0.f, sm);
    if (h == 0) h_out[(size_t)t * C + c] = d + b_eout[c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pair16 storage (gemm_p16.cuh): every 16 consecutive values of a row are stored as 64 bytes [16 bf16 hi | 16 bf16 lo].
__device__ __forceinline__ void pair16_store8(float* row_base, int c8, const float* x) {
  uint32_t hh[4], ll[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 f = __bfloat1622float2(a);
    const __nv_bfloat162 b = __floats2bfloat162_rn(x[2 * i] - f.x, x[2 * i + 1] - f.y);
    hh[i] = *reinterpret_cast<const uint32_t*>(&a);
    ll[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  uint8_t* p = reinterpret_cast<uint8_t*>(row_base) + (size_t)(c8 >> 4) * 64 + (size_t)(c8 & 15) * 2;
  *reinterpret_cast<uint4*>(p) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  *reinterpret_cast<uint4*>(p + 32) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}
__device__ __forceinline__ void pair16_load8(const float* row_base, int c8, float* x) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(row_base) + (size_t)(c8 >> 4) * 64 + (size_t)(c8 & 15) * 2;
  const uint4 hi = *reinterpret_cast<const uint4*>(p), lo = *reinterpret_cast<const uint4*>(p + 32);
  const uint32_t hh[4] = {hi.x, hi.y, hi.z, hi.w}, ll[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    x[2 * i] = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
    x[2 * i + 1] = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
  }
}

// Constant row of masked edges in the storage format of the edge state: [c3 x H | c3 x H | f0 | 0 x R | pad].  One block.
template <bool PAIR>
__global__ void k_const_row(int H, int R, int ld, const float* __restrict__ f0, const float* __restrict__ c3,
                            float* __restrict__ crow) {
  const float c = *c3;
  for (int c8 = threadIdx.x * 8; c8 < ld; c8 += blockDim.x * 8) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int col = c8 + k;
      x[k] = col < 2 * H ? c : (col < 3 * H ? f0[col - 2 * H] : 0.f);
    }
    if (PAIR) pair16_store8(crow, c8, x);
    else {
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (c8 + k < ld) crow[c8 + k] = x[k];
    }
  }
}

// (I) initial edge state, masked edges: copy of the constant row.  One warp per edge, 16-byte copies (ld % 4 == 0).
__global__ void k_edge_init_masked(int E, int ld, const int* __restrict__ act_pos, const float* __restrict__ crow,
                                   float* __restrict__ ew) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= E || act_pos[e] >= 0) return;
  const uint4* src = reinterpret_cast<const uint4*>(crow);
  uint4* dst = reinterpret_cast<uint4*>(ew + (size_t)e * ld);
  for (int k = lane; k < ld / 4; k += 32) dst[k] = src[k];
}

// (I) initial edge state of ACTIVE edges e0 = [ (sc3 | sc4) * rb | f | rbf ]  (leftnet.py:792-809).  Persistent blocks
// stride over the compact active list; thread t < 2H evaluates lin3 (3 -> Hq -> 1, SiLU) for side t / H, channel t % H,
// with the weights read as float4 (w0, w1, w2, b0) shared-memory broadcasts and four SiLUs per reciprocal.  The inputs
// of the next edge are fetched before the current one is evaluated (the index -> geometry -> NE1 chain is three
// dependent loads deep); the row is staged in shared memory and written once, in pair16 (PAIR) or fp32.
struct EdgeInitIn { int e; float a0, a1, a2, b0, b1, b2, gx, gy, gz, cx, cy, cz, rbe, fv, rv; };
// Thread h < H evaluates lin3 for BOTH sides (NE1 of the source and of the target) of channel h: the two evaluations share
// every weight operand, which halves the constant-bank traffic that bounds this kernel.
// HQ4 > 0: the hidden width rounded up to a multiple of 4 as a compile-time loop bound (the padding units have zero weights and
// contribute SiLU(0) * 0): without the per-group branch the unrolled groups interleave, which hides the MUFU latency of one
// group's exp2 / rcp chain behind the FMAs of the next (the branchy form ran at IPC 0.75 per SM, 2.4x its MUFU floor).
template <bool PAIR, int HQ4>
__global__ void __launch_bounds__(256, 4) k_edge_init_act(
    int H, int R, int reflect, int ld, int cap, const int* __restrict__ n_act, const int* __restrict__ act_idx,
    const int* __restrict__ esrc, const int* __restrict__ ecol, const float4* __restrict__ ecross,
    const float4* __restrict__ geo, const float* __restrict__ rb, const float* __restrict__ NE1,
    const float* __restrict__ f_act, const float* __restrict__ rbf_act, const __grid_constant__ Lin3E W,
    float* __restrict__ ew) {
  extern __shared__ __align__(16) float sm_ei[];  // row[2][ld]: the row of edge i leaves while edge i + 1 is evaluated (one barrier per edge)
  const int na = min(*n_act, cap);
  const int t = threadIdx.x;
  for (int k = 3 * H + R + t; k < ld; k += blockDim.x) { sm_ei[k] = 0.f; sm_ei[ld + k] = 0.f; }
  auto load = [&](int p) {
    EdgeInitIn in;
    in.e = act_idx[p];
    in.fv = t < H ? f_act[(size_t)p * H + t] : 0.f;
    in.rv = t < R ? rbf_act[(size_t)p * R + t] : 0.f;
    in.a0 = in.a1 = in.a2 = in.b0 = in.b1 = in.b2 = in.gx = in.gy = in.gz = in.cx = in.cy = in.cz = in.rbe = 0.f;
    if (t < H) {
      const int i = esrc[in.e], j = ecol[in.e];
      const float4 g = geo[in.e];
      in.gx = g.x; in.gy = g.y; in.gz = g.z;
      in.rbe = rb[in.e];
      // edge frame columns: u (unit diff), c (unit cross of pos_frame_i x pos_frame_j, k_edge_geom), v = u x c   (:693-705)
      const float4 cr = ecross[in.e];
      in.cx = cr.x; in.cy = cr.y; in.cz = cr.z;
      const float* ni = NE1 + (size_t)i * 3 * H;
      const float* nj = NE1 + (size_t)j * 3 * H;
      in.a0 = ni[t]; in.a1 = ni[H + t]; in.a2 = ni[2 * H + t];
      in.b0 = nj[t]; in.b1 = nj[H + t]; in.b2 = nj[2 * H + t];
    }
    return in;
  };
  int p = blockIdx.x;
  EdgeInitIn cur{};
  if (p < na) cur = load(p);
  __syncthreads();
  for (int par = 0; p < na; p += gridDim.x, par ^= 1) {
    float* row = sm_ei + par * ld;
    EdgeInitIn nxt{};
    if (p + (int)gridDim.x < na) nxt = load(p + gridDim.x);
    if (t < H) row[2 * H + t] = cur.fv;
    if (t < R) row[3 * H + t] = cur.rv;
    if (t < H) {
      const float cx = cur.cx, cy = cur.cy, cz = cur.cz;
      const float vx = cur.gy * cz - cur.gz * cy, vy = cur.gz * cx - cur.gx * cz, vz = cur.gx * cy - cur.gy * cx;
      const float s0 = cur.a0 * cur.gx + cur.a1 * cur.gy + cur.a2 * cur.gz;
      float s1 = cur.a0 * cx + cur.a1 * cy + cur.a2 * cz;
      const float s2 = cur.a0 * vx + cur.a1 * vy + cur.a2 * vz;
      const float r0 = cur.b0 * cur.gx + cur.b1 * cur.gy + cur.b2 * cur.gz;
      float r1 = cur.b0 * cx + cur.b1 * cy + cur.b2 * cz;
      const float r2 = cur.b0 * vx + cur.b1 * vy + cur.b2 * vz;
      if (reflect) { s1 = fabsf(s1); r1 = fabsf(r1); }
      // the two sides are the two halves of every packed operation
      const f32x2 p0 = pk2(s0, r0), p1 = pk2(s1, r1), p2 = pk2(s2, r2);
      f32x2 acc2 = pk2(W.b2, W.b2);
#pragma unroll
      for (int k = 0; k < (HQ4 > 0 ? HQ4 : 64); k += 4) {
        if (HQ4 > 0 || k < W.hq) {  // uniform; entries hq .. hq4 are zero-filled on the host
          f32x2 u[4];
#pragma unroll
          for (int i = 0; i < 4; i++)
            u[i] = fma2(ld2(W.w0[(k + i) * 3]), p0, fma2(ld2(W.w0[(k + i) * 3 + 1]), p1, fma2(ld2(W.w0[(k + i) * 3 + 2]), p2, ld2(W.b0[k + i]))));
          silu4_shared_rcp2(u[0], u[1], u[2], u[3]);
          acc2 = fma2(ld2(W.w2[k]), u[0], fma2(ld2(W.w2[k + 1]), u[1], fma2(ld2(W.w2[k + 2]), u[2], fma2(ld2(W.w2[k + 3]), u[3], acc2))));
        }
      }
      float acc, bcc;
      upk2(acc2, acc, bcc);
      row[t] = (acc + s0) * cur.rbe;
      row[H + t] = (bcc + r0) * cur.rbe;
    }
    __syncthreads();
    float* out = ew + (size_t)cur.e * ld;
    if (PAIR) {
      for (int c8 = t * 8; c8 < ld; c8 += blockDim.x * 8) pair16_store8(out, c8, row + c8);
    } else {
      for (int k = t; k < ld; k += blockDim.x) out[k] = row[k];
    }
    // (no second barrier: the next edge fills the other buffer, and this one is refilled only after the next barrier)
    cur = nxt;
  }
}

// Mean aggregation at the edge source from the per-run partial sums of the fused GCL tail (gcl_tail.cuh): node t's edges
// are the rows [row_ptr[t], row_ptr[t+1]) and span 1-3 groups of 32 rows; in every group they form exactly one run, found by
// its source id.  Summed in group order (bitwise reproducible).  One block per node, thread = channel.
__global__ void k_agg_runs(int H, const int* __restrict__ row_ptr, const float* __restrict__ P, int ldp,
                           const int* __restrict__ Psrc, float* __restrict__ xa, int ldxa) {
  const int t = blockIdx.x, h = threadIdx.x, lane = threadIdx.x & 31;
  const int r0 = row_ptr[t], r1 = row_ptr[t + 1];
  float sum = 0.f;
  if (r1 > r0) {
    for (int gq = r0 >> 5; gq <= (r1 - 1) >> 5; gq++) {
      const unsigned hit = __ballot_sync(0xffffffffu, Psrc[(size_t)gq * 32 + lane] == t);
      const int slot = __ffs(hit) - 1;  // (every warp of the block finds the same slot)
      if (slot >= 0 && h < H) sum += P[((size_t)gq * 32 + slot) * ldp + h];
    }
  }
  if (h < H) xa[(size_t)t * ldxa + H + h] = sum / (float)(r1 > r0 ? r1 - r0 : 1);
}

// GCL attention gate + mean aggregation at the edge source on a pair16 m2 (row pitch ld floats, H <= ld).
// One block per node; warps take the row's edges round-robin; lane l < ld/8 owns 8 consecutive columns.
__global__ void k_att_agg_p16(int H, int ld, const int* __restrict__ row_ptr, const float* __restrict__ m2,
                              const float* __restrict__ w_att, const float* __restrict__ b_att,
                              float* __restrict__ att_out, float* __restrict__ xa, int ldxa) {
  extern __shared__ float part[];  // [nw][ld]
  const int t = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int r0 = row_ptr[t], r1 = row_ptr[t + 1];
  const int c8 = lane * 8;
  const bool own = c8 < ld;
  float wa[8], acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { acc[k] = 0.f; wa[k] = (own && c8 + k < H) ? w_att[c8 + k] : 0.f; }
  const float ba = b_att[0];
  constexpr int NE = 2;  // edges in flight per warp (NE = 4 measured slower: 47 vs 40 us, fewer busy warps per row), same summation order
  for (int e0 = r0 + w; e0 < r1; e0 += NE * nw) {
    float v[NE][8], d[NE];
#pragma unroll
    for (int j = 0; j < NE; j++) {
      const int e = e0 + j * nw;
#pragma unroll
      for (int k = 0; k < 8; k++) v[j][k] = 0.f;
      if (own && e < r1) pair16_load8(m2 + (size_t)e * ld, c8, v[j]);
    }
#pragma unroll
    for (int j = 0; j < NE; j++) {
      d[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) d[j] = fmaf(v[j][k], wa[k], d[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < NE; j++) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
#pragma unroll
    for (int j = 0; j < NE; j++) {
      const int e = e0 + j * nw;
      if (e < r1) {  // warp-uniform
        const float att = silu(d[j] + ba);
        if (lane == 0) att_out[e] = att;
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] += v[j][k] * att;
      }
    }
  }
  if (own) {
#pragma unroll
    for (int k = 0; k < 8; k++) part[w * ld + c8 + k] = acc[k];
  }
  __syncthreads();
  const int h = threadIdx.x;
  if (h < H) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ww++) s += part[ww * ld + h];
    const int cnt = r1 - r0;
    xa[(size_t)t * ldxa + H + h] = s / (float)(cnt > 0 ? cnt : 1);
  }
}

}  // namespace oard
