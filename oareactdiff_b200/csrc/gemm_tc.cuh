// tcgen05 / TMEM GEMM for the per-edge contractions of LEFTNet (sm_100a only).
//
//   C[m, n] = epi( sum_k A[arow(m), k] * W[n, k] )      A: fp32 in HBM (edge state / activations), W: nn.Linear weight
//
// fp32-grade accuracy on the bf16 tensor pipe by error-compensated splitting (bf16x3):
//   a = a_hi + a_lo, w = w_hi + w_lo (bf16 each);  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi   (dropped term ~2^-16)
// accumulated in fp32 in TMEM by three tcgen05.mma.kind::f16 per K step.
//
// Persistent, warp-specialised CTA (320 threads, 1 CTA/SM):
//   warps 0-3  epilogue : tcgen05.ld accumulator -> bias / gathered row adds / SiLU / scale / mul / residual -> HBM
//   warp  4    MMA      : one elected lane issues tcgen05.mma; owns TMEM alloc/dealloc (2 accumulators, double buffered)
//   warp  5    W loader : cp.async.bulk (TMA bulk copy, UBLKCP) of pre-split, pre-tiled weight slabs, mbarrier complete_tx
//   warps 6-9  A producer: fp32 rows (optionally gathered) from HBM/L2 -> bf16 hi/lo -> shared memory in the UMMA
//                          K-major core-matrix layout (no swizzle): 8 rows x 16 B per core matrix, K-adjacent contiguous
// Pipelines: S-stage smem ring (full_a / full_w / empty mbarriers) and a 2-deep TMEM ring (acc_full / acc_empty).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_simt.cuh"  // GemmArgs

namespace oard {

constexpr int TC_BM = 128;      // rows per tile (UMMA M)
constexpr int TC_KC = 32;       // K elements per pipeline stage (2 x UMMA_K)
constexpr int TC_THREADS = 320;
constexpr int TC_CORE_BYTES = 128;                   // one 8x8 bf16 core matrix
constexpr int TC_SBO = (TC_KC / 8) * TC_CORE_BYTES;  // byte stride between 8-row groups (K-adjacent cores contiguous)

// Pre-split, pre-tiled weight: [n_tiles][k_chunks][hi, lo][BN x KC in core-matrix layout], zero padded.
struct TcWeight {
  const __nv_bfloat16* data;
  int N, K, BN, n_tiles, k_chunks;
};

__host__ __device__ inline size_t tc_weight_elems(int N, int K, int BN) {
  const int n_tiles = (N + BN - 1) / BN, k_chunks = (K + TC_KC - 1) / TC_KC;
  return (size_t)n_tiles * k_chunks * 2 * BN * TC_KC;
}

// byte offset of element (r, k) inside one [rows x KC] operand block
__host__ __device__ inline int tc_core_off(int r, int k) {
  return (r >> 3) * TC_SBO + (k >> 3) * TC_CORE_BYTES + (r & 7) * 16 + (k & 7) * 2;
}

__global__ void k_tc_pack_weight(const float* __restrict__ W, int ldw, int N, int K, int BN,
                                 __nv_bfloat16* __restrict__ out) {
  const int n_tiles = (N + BN - 1) / BN, k_chunks = (K + TC_KC - 1) / TC_KC;
  const size_t total = (size_t)n_tiles * k_chunks * BN * TC_KC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % TC_KC);
    const int nn = (int)((i / TC_KC) % BN);
    const int kc = (int)((i / ((size_t)TC_KC * BN)) % k_chunks);
    const int nt = (int)(i / ((size_t)TC_KC * BN * k_chunks));
    const int n = nt * BN + nn, k = kc * TC_KC + kk;
    const float w = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const size_t blk = ((size_t)nt * k_chunks + kc) * 2 * BN * TC_KC;  // elements
    const int off = tc_core_off(nn, kk) / 2;
    out[blk + off] = hi;
    out[blk + (size_t)BN * TC_KC + off] = lo;
  }
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
}  // namespace ptx

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version(=1) <<46 | layout_type(=0) <<61.   LBO = byte distance between K-adjacent core matrices, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}
// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major, N>>3 at 17, M>>4 at 24
__host__ __device__ inline uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


struct TcDebugOpts {  // layout experiments for bring-up (normally all zero)
  int swap_lbo_sbo;
};

// Epilogue modes (compile-time, keeps the register ring of each variant small):
//   0 plain: bias / SiLU / row scale      1: + two gathered row adds (GCL: P[src] + Q[dst])
//   2: * mul[m, n] (EquiMessage: rbf_proj gate)      3: + resid[m, n] (edge-state residual, may alias C)
template <int STAGES, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const GemmArgs g, const TcWeight w, const TcDebugOpts dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int BN = w.BN;
  const int A_PART = TC_BM * TC_KC * 2;  // bytes of one A part (hi or lo)
  const int W_PART = BN * TC_KC * 2;
  const int STAGE_BYTES = 2 * A_PART + 2 * W_PART;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full_a = bars;
  uint64_t* full_w = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* acc_full = bars + 3 * STAGES;
  uint64_t* acc_empty = bars + 3 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  const int m_tiles = (M + TC_BM - 1) / TC_BM;
  const int total_tiles = m_tiles * w.n_tiles;
  const int k_chunks = w.k_chunks;
  const int k16_total = (g.K + 15) / 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      ptx::mbar_init(&full_a[s], 4);  // one elected lane per producer warp
      ptx::mbar_init(&full_w[s], 1);  // arrive.expect_tx by the loader lane
      ptx::mbar_init(&empty[s], 1);   // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      ptx::mbar_init(&acc_full[b], 1);   // tcgen05.commit
      ptx::mbar_init(&acc_empty[b], 4);  // one elected lane per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 4) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 6) {
    // ===================== A producer: one row per thread, loads run 2 chunks ahead of the conversion =====================
    const int p = threadIdx.x - 192;  // 0..127
    const int row_off = (p >> 3) * TC_SBO + (p & 7) * 16;
    struct It { int tile, kc; const float* arow; bool ok; };
    auto init_row = [&](It& it) {
      const int m = (it.tile / w.n_tiles) * TC_BM + p;
      it.ok = it.tile < total_tiles && m < M;
      it.arow = g.A + (size_t)(it.ok ? (g.aidx ? g.aidx[m] : m) : 0) * g.lda;
    };
    auto advance = [&](It& it) {
      if (++it.kc == k_chunks) { it.kc = 0; it.tile += gridDim.x; init_row(it); }
    };
    auto issue = [&](const It& it, float4* v) {
#pragma unroll
      for (int j = 0; j < TC_KC / 4; j++) {
        const int k = it.kc * TC_KC + j * 4;
        v[j] = (it.ok && k < g.K) ? __ldg(reinterpret_cast<const float4*>(it.arow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto consume = [&](uint32_t gchunk, const float4* v) {
      const int s = gchunk % STAGES;
      const uint32_t ph = (gchunk / STAGES) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_hi = smem + (size_t)s * STAGE_BYTES + row_off;
      uint8_t* a_lo = a_hi + A_PART;
#pragma unroll
      for (int j = 0; j < TC_KC / 8; j++) {  // one 16-byte unit (8 bf16) per core-matrix column
        const float x[8] = {v[2 * j].x, v[2 * j].y, v[2 * j].z, v[2 * j].w,
                            v[2 * j + 1].x, v[2 * j + 1].y, v[2 * j + 1].z, v[2 * j + 1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
          const float2 hf = __bfloat1622float2(h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
          hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        *reinterpret_cast<uint4*>(a_hi + j * TC_CORE_BYTES) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + j * TC_CORE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_a[s]);
    };
    const int my_tiles = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t nchunks = (uint32_t)my_tiles * k_chunks;
    It ld{(int)blockIdx.x, 0, nullptr, false};
    init_row(ld);
    float4 b0[TC_KC / 4], b1[TC_KC / 4], b2[TC_KC / 4];
    issue(ld, b0); advance(ld);
    issue(ld, b1); advance(ld);
    for (uint32_t c = 0; c < nchunks; c += 3) {
      issue(ld, b2); advance(ld);
      consume(c, b0);
      if (c + 1 < nchunks) { issue(ld, b0); advance(ld); consume(c + 1, b1); }
      if (c + 2 < nchunks) { issue(ld, b1); advance(ld); consume(c + 2, b2); }
    }
  } else if (warp == 5) {
    // ===================== W loader: TMA bulk copies of pre-tiled slabs =====================
    if (lane == 0) {
      uint32_t gchunk = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % w.n_tiles;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(w.data) + (size_t)nt * k_chunks * 2 * W_PART;
        for (int kc = 0; kc < k_chunks; kc++, gchunk++) {
          const int s = gchunk % STAGES;
          const uint32_t ph = (gchunk / STAGES) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full_w[s], 2 * W_PART);
          ptx::bulk_g2s(smem + (size_t)s * STAGE_BYTES + 2 * A_PART, src + (size_t)kc * 2 * W_PART, 2 * W_PART,
                        &full_w[s]);
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc_idesc(TC_BM, BN);
      const uint32_t lbo = dbg.swap_lbo_sbo ? TC_SBO : TC_CORE_BYTES, sbo = dbg.swap_lbo_sbo ? TC_CORE_BYTES : TC_SBO;
      uint32_t gchunk = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int kc = 0; kc < k_chunks; kc++, gchunk++) {
          const int s = gchunk % STAGES;
          const uint32_t ph = (gchunk / STAGES) & 1;
          ptx::mbar_wait(&full_a[s], ph);
          ptx::mbar_wait(&full_w[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_hi = ptx::smem_u32(smem + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_PART;
          const uint32_t w_hi = a_hi + 2 * A_PART, w_lo = w_hi + W_PART;
          const int steps = min(TC_KC / 16, k16_total - kc * (TC_KC / 16));
          for (int j = 0; j < steps; j++) {
            const uint32_t ko = j * 2 * TC_CORE_BYTES;  // two K-adjacent core matrices per UMMA_K = 16
            const uint64_t dah = tc_smem_desc(a_hi + ko, lbo, sbo), dal = tc_smem_desc(a_lo + ko, lbo, sbo);
            const uint64_t dwh = tc_smem_desc(w_hi + ko, lbo, sbo), dwl = tc_smem_desc(w_lo + ko, lbo, sbo);
            ptx::umma_bf16(d_tmem, dah, dwh, idesc, (kc | j) != 0);
            ptx::umma_bf16(d_tmem, dah, dwl, idesc, 1);
            ptx::umma_bf16(d_tmem, dal, dwh, idesc, 1);
          }
          ptx::umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
        }
        ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue warps 0-3: TMEM lane quarter = warp; one output row per thread =====================
    // 16-column groups; the row-wise operands of the epilogue (gathered adds / mul / resid) are fetched RING-1 groups
    // ahead into a register ring so their HBM/L2 latency overlaps the math of earlier groups.
    constexpr int NA = MODE == 1 ? 2 : (MODE == 0 ? 0 : 1);  // row-operand arrays per group
    constexpr int RING = 4;
    const int ngroups = BN / 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const int m = (tile / w.n_tiles) * TC_BM + warp * 32 + lane;
      const int n0 = (tile % w.n_tiles) * BN;
      const bool ok = m < M;
      const float* ra = nullptr;
      const float* rb2 = nullptr;
      if (ok) {
        if (MODE == 1) { ra = g.radd1 + (size_t)(g.ridx1 ? g.ridx1[m] : m) * g.ld1; rb2 = g.radd2 + (size_t)(g.ridx2 ? g.ridx2[m] : m) * g.ld2; }
        if (MODE == 2) ra = g.mul + (size_t)m * g.ldmul;
        if (MODE == 3) ra = g.resid + (size_t)m * g.ldres;
      }
      const float rs = (ok && g.rowscale) ? g.rowscale[g.rsidx ? g.rsidx[m] : m] : 1.f;
      float* crow = g.C + (size_t)(ok ? m : 0) * g.ldc;
      float4 aux[RING][NA > 0 ? NA * 4 : 1];
      auto issue = [&](int grp, float4* a) {
        if (NA == 0) return;
        const int n = n0 + grp * 16;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const bool in = ok && grp < ngroups && (n + q * 4) < g.N;
          a[q] = in ? *reinterpret_cast<const float4*>(ra + n + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (NA == 2) a[4 + q] = in ? *reinterpret_cast<const float4*>(rb2 + n + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto process = [&](int grp, const float4* a) {
        float v[16];
        ptx::tmem_ld16(tmem_base + buf * 256 + ((uint32_t)(warp * 32) << 16) + grp * 16, v);  // warp-collective
        const int n = n0 + grp * 16;
        if (ok && n < g.N) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int nq = n + q * 4;
            if (nq < g.N) {  // N % 4 == 0 (checked on the host)
              float4 x = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
              if (g.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + nq)); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
              if (MODE == 1) {
                x.x += a[q].x + a[4 + q].x; x.y += a[q].y + a[4 + q].y; x.z += a[q].z + a[4 + q].z; x.w += a[q].w + a[4 + q].w;
              }
              if (g.act == 1) { x.x = silu_fast(x.x); x.y = silu_fast(x.y); x.z = silu_fast(x.z); x.w = silu_fast(x.w); }
              x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs;
              if (MODE == 2) { x.x *= a[q].x; x.y *= a[q].y; x.z *= a[q].z; x.w *= a[q].w; }
              if (MODE == 3) { x.x += a[q].x; x.y += a[q].y; x.z += a[q].z; x.w += a[q].w; }
              *reinterpret_cast<float4*>(crow + nq) = x;
            }
          }
        }
        __syncwarp();
      };
      // the row operands do not depend on the accumulator: start fetching before waiting for the MMAs
#pragma unroll
      for (int u = 0; u < RING - 1; u++) issue(u, aux[u]);
      ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      for (int g0 = 0; g0 < ngroups; g0 += RING) {
#pragma unroll
        for (int u = 0; u < RING; u++) {
          if (g0 + u < ngroups) {
            issue(g0 + u + RING - 1, aux[(u + RING - 1) % RING]);
            process(g0 + u, aux[u]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

inline size_t tc_smem_bytes(int BN, int stages) {
  return (size_t)stages * (2 * TC_BM * TC_KC * 2 + 2 * BN * TC_KC * 2) + (3 * stages + 4) * 8 + 16;
}

// Requirements (checked): K % 4 == 0, lda % 4 == 0, N % 4 == 0, 16-byte aligned operands, BN % 16 == 0, BN <= 256,
// at most one of {radd1+radd2, mul, resid}.
template <int STAGES, int MODE>
inline cudaError_t launch_gemm_tc_inst(const GemmArgs& g, const TcWeight& w, int grid, size_t smem, TcDebugOpts dbg,
                                       cudaStream_t st) {
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  gemm_tc_kernel<STAGES, MODE><<<grid, TC_THREADS, smem, st>>>(g, w, dbg);
  return cudaGetLastError();
}

inline cudaError_t launch_gemm_tc(const GemmArgs& g, const TcWeight& w, int num_sms, cudaStream_t st,
                                  int swap_lbo_sbo = 0) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.K % 4 || g.lda % 4 || g.N % 4 || g.ldc % 4 || w.BN % 16 || w.BN > 256 || w.BN < 16 || g.K != w.K || g.N != w.N)
    return cudaErrorInvalidValue;
  const int nmode = (g.radd1 ? 1 : 0) + (g.mul ? 1 : 0) + (g.resid ? 1 : 0);
  if (nmode > 1 || (g.radd1 && !g.radd2) || (!g.radd1 && g.radd2)) return cudaErrorInvalidValue;
  const int mode = g.radd1 ? 1 : (g.mul ? 2 : (g.resid ? 3 : 0));
  const int m_tiles = (g.M + TC_BM - 1) / TC_BM;
  const int total = m_tiles * w.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  TcDebugOpts dbg{swap_lbo_sbo};
  const int stages = tc_smem_bytes(w.BN, 4) <= 227 * 1024 ? 4 : 3;
  const size_t smem = tc_smem_bytes(w.BN, stages);
#define OARD_TC_CASE(S, MD) \
  if (stages == S && mode == MD) return launch_gemm_tc_inst<S, MD>(g, w, grid, smem, dbg, st);
  OARD_TC_CASE(4, 0) OARD_TC_CASE(4, 1) OARD_TC_CASE(4, 2) OARD_TC_CASE(4, 3)
  OARD_TC_CASE(3, 0) OARD_TC_CASE(3, 1) OARD_TC_CASE(3, 2) OARD_TC_CASE(3, 3)
#undef OARD_TC_CASE
  return cudaErrorInvalidValue;
}

}  // namespace oard
