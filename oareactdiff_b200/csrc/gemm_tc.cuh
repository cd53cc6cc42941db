// tcgen05 / TMEM GEMM for the dense contractions of LEFTNet (sm_100a only).
//
//   C[m, n] = epi( sum_k A[arow(m), k] * W[n, k] )      A: fp32 in HBM (edge state / activations), W: nn.Linear weight
//
// fp32-grade accuracy on the bf16 tensor pipe by error-compensated splitting (bf16x3):
//   a = a_hi + a_lo, w = w_hi + w_lo (bf16 each);  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi   (dropped term ~2^-16)
// accumulated in fp32 in TMEM by three tcgen05.mma.kind::f16 per K step.
//
// Persistent, warp-specialised CTA (576 threads, 1 CTA/SM):
//   warps 0-7   epilogue : tcgen05.ld accumulator (thread = row) -> smem transpose staging -> coalesced
//                          bias / gathered row adds / SiLU / scale / mul / residual -> coalesced float4 stores
//   warp  8     MMA      : one elected lane issues tcgen05.mma; owns TMEM alloc/dealloc (2 accumulators, double buffered)
//   warp  9     W loader : cp.async.bulk (TMA bulk copy, UBLKCP) of pre-split, pre-tiled weight slabs, mbarrier complete_tx
//   warps 10-17 A producer: coalesced float4 loads of fp32 rows (optionally gathered) -> bf16 hi/lo -> shared memory in the
//                          UMMA K-major core-matrix layout (SWIZZLE_NONE), K-stride padded to 144 B so the 8-byte
//                          stores of a warp are bank-conflict free; loads run two K-chunks ahead of the conversion
// Pipelines: S-stage smem ring (full_a / full_w / empty mbarriers) and a 2-deep TMEM ring (acc_full / acc_empty).
// Every global access of the kernel is a full 128-byte line per 8 threads: the first version (row per thread) was bound
// by L1 tag throughput (32 lines per warp instruction), see profiles/.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm_simt.cuh"  // GemmArgs

namespace oard {

constexpr int TC_BM = 128;      // rows per tile (UMMA M)
constexpr int TC_KC = 32;       // K elements per pipeline stage (2 x UMMA_K)
constexpr int TC_EPI_WARPS = 8, TC_PROD_WARPS = 8;
constexpr int TC_CTRL_WARPS = 3;  // MMA issuer, W loader, A loader (TMA)
constexpr int TC_THREADS = (TC_EPI_WARPS + TC_CTRL_WARPS + TC_PROD_WARPS) * 32;  // 608
constexpr int TC_RAW_BYTES = TC_BM * TC_KC * 4;  // one raw fp32 A box (128 rows x 32 floats) as landed by TMA
constexpr int TC_CORE_BYTES = 128;                   // one 8x8 bf16 core matrix
constexpr int TC_SBO = (TC_KC / 8) * TC_CORE_BYTES;  // W operand: byte stride between 8-row groups (K-adjacent cores contiguous)
constexpr int TC_A_LBO = 144;                        // A operand: K-adjacent core matrices 144 B apart (bank-conflict-free stores)
constexpr int TC_A_SBO = (TC_KC / 8) * TC_A_LBO;     // 576
constexpr int TC_A_PART = (TC_BM / 8) * TC_A_SBO;    // 9216 bytes per A part (hi or lo)
constexpr int TC_IO_BYTES = 32 * 32 * 4;             // one 32x32 fp32 epilogue block (TMA box, 128B-swizzled)
constexpr int TC_STG_LD = 36;                        // epilogue staging row pitch in floats (32 + 4 pad)
constexpr int TC_STG_BYTES = TC_EPI_WARPS * 32 * TC_STG_LD * 4;

// Pre-split, pre-tiled weight: [n_tiles][k_chunks][hi, lo][BN x KC in core-matrix layout], zero padded.
struct TcWeight {
  const __nv_bfloat16* data;
  int N, K, BN, n_tiles, k_chunks;
};

// Output-column tile width: one tile (<= 256 columns, multiple of 16) when N fits, else equal tiles whose width is a multiple
// of 32 so that the epilogue's 32-column TMA boxes of neighbouring tiles never overlap.
inline int tc_choose_bn(int N) {
  const int n_tiles = (N + 255) / 256;
  const int per = (N + n_tiles - 1) / n_tiles;
  const int q = n_tiles > 1 ? 32 : 16;
  return (per + q - 1) / q * q;
}

__host__ __device__ inline size_t tc_weight_elems(int N, int K, int BN) {
  const int n_tiles = (N + BN - 1) / BN, k_chunks = (K + TC_KC - 1) / TC_KC;
  return (size_t)n_tiles * k_chunks * 2 * BN * TC_KC;
}

// byte offset of element (r, k) inside one [rows x KC] W operand block
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
// Same, for a waiter that usually finds the phase complete (the MMA issuer, whose every cycle is on the critical path):
// a plain test first, the potentially suspending try_wait only if the phase is still open.
__device__ __forceinline__ void mbar_wait_hot(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  if (!done) mbar_wait(bar, parity);
}
// One lane of the (converged) warp, the same one every time.  Code that issues tcgen05 / TMA instructions runs warp-uniform
// and predicates only those instructions with this: inside an `if (lane == 0)` region ptxas cannot prove a single active
// thread and wraps every UTCHMMA / UTCBAR / UTMALDG in an ELECT ... BRA.U.ANY loop (~40 cycles per instruction, which made
// the MMA issuer, not the tensor pipe, the bound of the edge GEMMs: profiles/r2_issue_loop_notes.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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

// 2-D TMA tile load (SASS: UTMALDG): box {32 floats, 128 rows} at element coordinates (k, row); out-of-bounds -> zeros
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// 2-D TMA tile store (SASS: UTMASTG): shared (dense box, optionally swizzled) -> global; out-of-bounds elements are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(smem_u32(src))
               : "memory");
}
// L2 eviction-priority hints for the TMA streams (0 none, 1 evict_first: data read / written once and too large to stay
// resident — it should leave L2 before the activations the NEXT kernel reads; 2 evict_last).  createpolicy + .L2::cache_hint.
__device__ __forceinline__ uint64_t l2_policy(int kind) {
  uint64_t p = 0;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_h(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, int kind, uint64_t pol) {
  if (kind == 0) { tma_load_2d(dst, tm, c0, c1, bar); return; }
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_h(const CUtensorMap* tm, int c0, int c1, const void* src, int kind, uint64_t pol) {
  if (kind == 0) { tma_store_2d(tm, c0, c1, src); return; }
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tm),
               "r"(c0), "r"(c1), "r"(smem_u32(src)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
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
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
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

struct TcDebugOpts {  // bring-up / ablation knobs (normally zero)
  int swap_lbo_sbo;
  int ablate;  // bit 0: no A global loads, bit 1: no epilogue global traffic, bit 2: no W bulk copies, bit 3: no MMAs
  long long* ts;  // timeline probe of CTA 0 (tools/tc_timeline.py): [0] / [15] globaltimer at entry / exit, [1..] clock64 marks
};
__device__ __forceinline__ void tc_mark(const TcDebugOpts& dbg, int i) {
  if (dbg.ts && blockIdx.x == 0) dbg.ts[i] = clock64();
}
__device__ __forceinline__ long long tc_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Epilogue modes (compile-time):
//   0 plain: bias / SiLU / row scale      1: + two gathered row adds (GCL: P[src] + Q[dst])
//   2: * mul[m, n] (EquiMessage: rbf_proj gate)      3: + resid[m, n] (edge-state residual, may alias C)
// RAW > 0: the A operand is contiguous (no row gather) and arrives by 2-D TMA into a RAW-deep ring of raw fp32 boxes;
// RAW == 0: gathered / unaligned A through the LSU (coalesced float4 loads, register ring).
// NIO > 0: the epilogue moves its blocks with TMA (residual / multiplier boxes in, output boxes out; NIO 4 KB buffers per
// epilogue warp); NIO == 0: LSU epilogue with an smem transpose (fallback for unaligned outputs).
template <int STAGES, int RAW, int MODE, int NIO>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const GemmArgs g, const TcWeight w, const TcDebugOpts dbg, const __grid_constant__ CUtensorMap tmA,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int BN = w.BN;
  const int W_PART = BN * TC_KC * 2;
  const int STAGE_BYTES = 2 * TC_A_PART + 2 * W_PART;
  uint8_t* raw_base = smem + (size_t)STAGES * STAGE_BYTES;
  float* stg_all = reinterpret_cast<float*>(raw_base + (size_t)RAW * TC_RAW_BYTES);  // NIO == 0: transpose staging
  uint8_t* io_all = raw_base + (size_t)RAW * TC_RAW_BYTES;                             // NIO > 0: TMA io blocks (1024-aligned)
  constexpr int EPI_SMEM = NIO > 0 ? TC_EPI_WARPS * NIO * TC_IO_BYTES : TC_STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)RAW * TC_RAW_BYTES + EPI_SMEM);
  uint64_t* full_a = bars;
  uint64_t* full_w = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* acc_full = bars + 3 * STAGES;
  uint64_t* acc_empty = bars + 3 * STAGES + 2;
  uint64_t* raw_full = bars + 3 * STAGES + 4;
  uint64_t* raw_empty = raw_full + (RAW > 0 ? RAW : 1);
  uint64_t* aux_full = raw_empty + (RAW > 0 ? RAW : 1);  // [TC_EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + TC_EPI_WARPS * 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_chunks = w.k_chunks;
  const int k16_total = (g.K + 15) / 16;
  if (dbg.ts && blockIdx.x == 0 && threadIdx.x == 0) { dbg.ts[0] = tc_globaltimer(); dbg.ts[1] = clock64(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      ptx::mbar_init(&full_a[s], TC_PROD_WARPS);  // one elected lane per producer warp
      ptx::mbar_init(&full_w[s], 1);              // arrive.expect_tx by the loader lane
      ptx::mbar_init(&empty[s], 1);               // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      ptx::mbar_init(&acc_full[b], 1);              // tcgen05.commit
      ptx::mbar_init(&acc_empty[b], TC_EPI_WARPS);  // one elected lane per epilogue warp
    }
    for (int i = 0; i < TC_EPI_WARPS * 2; i++) ptx::mbar_init(&aux_full[i], 1);  // arrive.expect_tx by the warp's lane 0
    for (int r = 0; r < RAW; r++) {
      ptx::mbar_init(&raw_full[r], 1);               // arrive.expect_tx by the A loader lane
      ptx::mbar_init(&raw_empty[r], TC_PROD_WARPS);  // one elected lane per producer warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == TC_EPI_WARPS) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  const int m_tiles = (M + TC_BM - 1) / TC_BM;
  const int total_tiles = m_tiles * w.n_tiles;
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tc_mark(dbg, 2);  // set-up done

  if (warp >= TC_EPI_WARPS + TC_CTRL_WARPS) {
    // ===================== A producer =====================
    // warp pw owns rows 16*pw .. 16*pw+15 of the tile.  One warp-wide float4 load covers 4 rows x 128 contiguous bytes
    // (rows r0, r0+4, r0+2, r0+6 so that, with the 144-byte K pitch, the 8-byte smem stores are bank-conflict free).
    const int pw = warp - (TC_EPI_WARPS + TC_CTRL_WARPS);
    const int kq = lane & 7;
    struct It { int tile, kc; const float* rp[4]; };
    // row offsets {0,4,2,6} per 8-lane group: the two groups of each half-warp (the unit a 64-bit shared store is
    // processed in) then land in disjoint bank groups with the 144-byte K pitch
    auto row_of = [&](int i) { return 16 * pw + ((i >> 1) << 3) + (i & 1) + (((lane >> 3) & 1) << 2) + ((lane >> 4) << 1); };
    auto init_rows = [&](It& it) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int m = (it.tile / w.n_tiles) * TC_BM + row_of(i);
        const bool ok = it.tile < total_tiles && m < M;
        it.rp[i] = ok ? g.A + (size_t)(g.aidx ? g.aidx[m] : m) * g.lda + kq * 4 : nullptr;
      }
    };
    auto advance = [&](It& it) {
      if (++it.kc == k_chunks) { it.kc = 0; it.tile += gridDim.x; init_rows(it); }
    };
    auto issue = [&](const It& it, float4* v) {
      const int k = it.kc * TC_KC + kq * 4;
#pragma unroll
      for (int i = 0; i < 4; i++)
        v[i] = (it.rp[i] && k < g.K && !(dbg.ablate & 1)) ? __ldg(reinterpret_cast<const float4*>(it.rp[i] + it.kc * TC_KC))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int soff[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = row_of(i);
      soff[i] = (r >> 3) * TC_A_SBO + (kq >> 1) * TC_A_LBO + (r & 7) * 16 + (kq & 1) * 8;
    }
    auto consume = [&](uint32_t gchunk, const float4* v) {
      const int s = gchunk % STAGES;
      const uint32_t ph = (gchunk / STAGES) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_hi = smem + (size_t)s * STAGE_BYTES;
      uint8_t* a_lo = a_hi + TC_A_PART;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x, v[i].y), h1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v[i].x - f0.x, v[i].y - f0.y);
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(v[i].z - f1.x, v[i].w - f1.y);
        *reinterpret_cast<uint2*>(a_hi + soff[i]) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(a_lo + soff[i]) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_a[s]);
    };
    if constexpr (RAW > 0) {
      // TMA path: the raw fp32 box is already in shared memory; read it (one 128-byte row segment per 8 lanes), split, store
      const int my_tiles_r = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      const uint32_t nchunks_r = (uint32_t)my_tiles_r * k_chunks;
      int roff[4];
#pragma unroll
      for (int i = 0; i < 4; i++) roff[i] = row_of(i) * (TC_KC * 4) + kq * 16;
      for (uint32_t c = 0; c < nchunks_r; c++) {
        const int r = c % RAW;
        ptx::mbar_wait(&raw_full[r], (c / RAW) & 1);
        if (c == 0 && pw == 0 && lane == 0) tc_mark(dbg, 4);  // first A box landed
        const uint8_t* raw = raw_base + (size_t)r * TC_RAW_BYTES;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = *reinterpret_cast<const float4*>(raw + roff[i]);
        // cross-proxy WAR: these generic-proxy reads must be ordered before the TMA engine (async proxy) refills the box.
        // Without a fence the refill can overtake the reads (seen on hardware: stale/corrupt rows in later tiles).  The fence
        // consume() issues before it publishes the converted stage orders them too, so the box is handed back after it:
        // one proxy fence per chunk instead of two (the chunk pipeline of small-M GEMMs runs at the producers' pace).
        consume(c, v);
        if (lane == 0) ptx::mbar_arrive(&raw_empty[r]);  // box consumed
      }
    } else {
    const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t nchunks = (uint32_t)my_tiles * k_chunks;
    It ld;
    ld.tile = blockIdx.x; ld.kc = 0;
    init_rows(ld);
    // (an explicit prefetch.global.L2 of the next tile was tried and REMOVED: DRAM reads rose 65 % — lines were
    //  evicted before use — and the kernel slowed down; see profiles/r1_tc_notes.md)
    float4 b0[4], b1[4], b2[4];
    issue(ld, b0); advance(ld);
    issue(ld, b1); advance(ld);
    for (uint32_t c = 0; c < nchunks; c += 3) {
      issue(ld, b2); advance(ld);
      consume(c, b0);
      if (c + 1 < nchunks) { issue(ld, b0); advance(ld); consume(c + 1, b1); }
      if (c + 2 < nchunks) { issue(ld, b1); advance(ld); consume(c + 2, b2); }
    }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
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
          if (dbg.ablate & 4) { ptx::mbar_arrive(&full_w[s]); continue; }
          ptx::mbar_arrive_expect_tx(&full_w[s], 2 * W_PART);
          ptx::bulk_g2s(smem + (size_t)s * STAGE_BYTES + 2 * TC_A_PART, src + (size_t)kc * 2 * W_PART, 2 * W_PART,
                        &full_w[s]);
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 2) {
    // ===================== A loader: 2-D TMA boxes {32 floats, 128 rows} into the raw ring =====================
    if (RAW > 0 && lane == 0) {
      const uint64_t polA = ptx::l2_policy(g.hintA);
      ptx::tma_prefetch_desc(&tmA);
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / w.n_tiles) * TC_BM;
        for (int kc = 0; kc < k_chunks; kc++, c++) {
          const int r = c % (RAW > 0 ? RAW : 1);
          ptx::mbar_wait(&raw_empty[r], ((c / (RAW > 0 ? RAW : 1)) & 1) ^ 1);
          if (dbg.ablate & 1) { ptx::mbar_arrive(&raw_full[r]); continue; }
          ptx::mbar_arrive_expect_tx(&raw_full[r], TC_RAW_BYTES);
          ptx::tma_load_2d_h(raw_base + (size_t)r * TC_RAW_BYTES, &tmA, kc * TC_KC, m0, &raw_full[r], g.hintA, polA);
          if (c == 0) tc_mark(dbg, 3);  // first A box requested
        }
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // ===================== MMA issuer (converged warp; one elected lane issues the tcgen05 instructions) =====================
    {
      const uint32_t idesc = tc_idesc(TC_BM, BN);
      const uint32_t wl = dbg.swap_lbo_sbo ? TC_SBO : TC_CORE_BYTES, ws = dbg.swap_lbo_sbo ? TC_CORE_BYTES : TC_SBO;
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
          if (gchunk == 0 && lane == 0) tc_mark(dbg, 5);  // first converted A stage
          ptx::mbar_wait(&full_w[s], ph);
          if (gchunk == 0 && lane == 0) tc_mark(dbg, 6);  // first weight slab
          ptx::tc_fence_after();
          const uint32_t a_hi = ptx::smem_u32(smem + (size_t)s * STAGE_BYTES), a_lo = a_hi + TC_A_PART;
          const uint32_t w_hi = a_hi + 2 * TC_A_PART, w_lo = w_hi + W_PART;
          const int steps = (dbg.ablate & 8) ? 0 : min(TC_KC / 16, k16_total - kc * (TC_KC / 16));
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < TC_KC / 16; j++) {  // two K-adjacent core matrices per UMMA_K = 16
              if (j >= steps) break;
              const uint32_t ka = j * 2 * TC_A_LBO, kw = j * 2 * TC_CORE_BYTES;
              const uint64_t dah = tc_smem_desc(a_hi + ka, TC_A_LBO, TC_A_SBO), dal = tc_smem_desc(a_lo + ka, TC_A_LBO, TC_A_SBO);
              const uint64_t dwh = tc_smem_desc(w_hi + kw, wl, ws), dwl = tc_smem_desc(w_lo + kw, wl, ws);
              ptx::umma_bf16(d_tmem, dah, dwh, idesc, (kc | j) != 0);
              ptx::umma_bf16(d_tmem, dah, dwl, idesc, 1);
              ptx::umma_bf16(d_tmem, dal, dwh, idesc, 1);
            }
            ptx::umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(&acc_full[buf]);
        __syncwarp();
        if (lane == 0) tc_mark(dbg, 7);  // last MMA of the (latest) tile issued
      }
    }
  } else {
    // ===================== epilogue warps 0-7 =====================
    // warp e reads TMEM lanes 32*(e%4).. (its row quarter) and takes the 32-column blocks with parity e/4.
    if constexpr (NIO > 0) {
      const uint64_t polX = ptx::l2_policy(g.hintX), polC = ptx::l2_policy(g.hintC);
      // ---- TMA epilogue: thread = output row.  A block is 32 rows x 32 columns (4 KB, 128-byte rows, 128B-swizzled so
      // that both the row-per-thread accesses here and the TMA engine are bank-conflict free).  Residual / multiplier blocks
      // are fetched by TMA one block ahead; results leave by TMA store.  No LSU traffic except bias and gathered rows.
      const int rq = warp & 3, half = warp >> 2;
      uint8_t* io = io_all + (size_t)warp * NIO * TC_IO_BYTES;
      uint64_t* xbar = aux_full + warp * 2;
      const int nblocks = (BN + 31) / 32;
      const int sw = lane & 7;
      constexpr bool HAS_AUX = (MODE == 2 || MODE == 3);
      // block iterator over (tile, blk) in this warp's processing order (used to prefetch the next aux block)
      int pf_tile = blockIdx.x, pf_blk = half;
      auto pf_valid = [&]() { return pf_tile < total_tiles && pf_blk < nblocks; };
      auto pf_next = [&]() {
        pf_blk += 2;
        if (pf_blk >= nblocks) { pf_blk = half; pf_tile += gridDim.x; }
      };
      uint32_t nb = 0;   // blocks processed by this warp
      uint32_t npf = 0;  // aux blocks requested by this warp
      auto issue_aux = [&]() {  // lane 0 only
        if (!HAS_AUX) return;
        while (pf_tile < total_tiles && pf_blk >= nblocks) pf_next();  // half == 1 with a single block
        if (!pf_valid()) return;
        const int b = npf % NIO;
        ptx::mbar_arrive_expect_tx(&xbar[b], TC_IO_BYTES);
        ptx::tma_load_2d_h(io + (size_t)b * TC_IO_BYTES, &tmX, (pf_tile % w.n_tiles) * BN + pf_blk * 32,
                           (pf_tile / w.n_tiles) * TC_BM + rq * 32, &xbar[b], g.hintX, polX);
        npf++;
        pf_next();
      };
      if (lane == 0) {
        ptx::tma_prefetch_desc(&tmC);
        if (HAS_AUX) { ptx::tma_prefetch_desc(&tmX); issue_aux(); }
      }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        const int m0r = (tile / w.n_tiles) * TC_BM + rq * 32;
        const int m = m0r + lane;
        const int n0 = (tile % w.n_tiles) * BN;
        const bool ok = m < M;
        const float* pr = nullptr;
        const float* qr = nullptr;
        if (MODE == 1 && ok) {
          pr = g.radd1 + (size_t)(g.ridx1 ? g.ridx1[m] : m) * g.ld1;
          qr = g.radd2 + (size_t)(g.ridx2 ? g.ridx2[m] : m) * g.ld2;
        }
        const float rs = (ok && g.rowscale) ? g.rowscale[g.rsidx ? g.rsidx[m] : m] : 1.f;
        const float ps = (ok && g.prescale) ? g.prescale[m] : 1.f;
        const int c2 = (ok && g.C2) ? g.c2idx[m] : -1;
        // the bias segment of this warp's first block into L1 while the MMAs run (a cold read costs ~700 cycles of the
        // epilogue's latency chain; node-level GEMMs have one block per warp and tile)
        if (g.bias && lane == 0 && n0 + half * 32 < g.N)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(g.bias + n0 + half * 32));
        ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
        ptx::tc_fence_after();
        if (warp == 0 && lane == 0) tc_mark(dbg, 8);  // accumulator of the (latest) tile complete
        for (int blk = half; blk < nblocks; blk += 2, nb++) {
          const int b = nb % NIO;
          uint8_t* iob = io + (size_t)b * TC_IO_BYTES + lane * 128;
          float v[32];
          ptx::tmem_ld32(tmem_base + buf * 256 + ((uint32_t)(rq * 32) << 16) + blk * 32, v);  // warp-collective
          if (warp == 0 && lane == 0) tc_mark(dbg, 13);  // first block in registers
          if (HAS_AUX) {
            // request the NEXT block's aux box (its buffer was last read by the store issued two blocks ago)
            if (lane == 0 && NIO > 1) { ptx::bulk_wait_read0(); issue_aux(); }
            ptx::mbar_wait(&xbar[b], (nb / NIO) & 1);
          } else {
            if (lane == 0) ptx::bulk_wait_read0();  // the previous store has finished reading this buffer
            __syncwarp();
          }
          const int nblk = n0 + blk * 32;
          // three branch-free passes over the block's 16 packed pairs (operands in, activation, out): inside one loop with the
          // column guards the 32 SiLU chains (mul -> ex2 -> add -> rcp -> mul) ran one after the other, 2 850 cycles per block
          f32x2 x[16];
          {
            const f32x2 ps2 = pk2(ps, ps);  // (ps = 1 without a prescale: fma(v, 1, b) = v + b exactly)
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int n = nblk + q * 4;
              const bool nin = n < g.N && !(dbg.ablate & 2);
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (g.bias && nin) t = __ldg(reinterpret_cast<const float4*>(g.bias + n));
              f32x2 t0 = pk2(t.x, t.y), t1 = pk2(t.z, t.w);
              if (MODE == 1 && nin && ok) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(pr + n));
                const float4 u = __ldg(reinterpret_cast<const float4*>(qr + n));
                t0 = add2(t0, add2(pk2(a.x, a.y), pk2(u.x, u.y)));
                t1 = add2(t1, add2(pk2(a.z, a.w), pk2(u.z, u.w)));
              }
              x[2 * q] = fma2(pk2(v[4 * q], v[4 * q + 1]), ps2, t0);
              x[2 * q + 1] = fma2(pk2(v[4 * q + 2], v[4 * q + 3]), ps2, t1);
            }
          }
          if (g.act == 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = silu2(x[k]);
          }
          if (g.rowscale) {
            const f32x2 rs2 = pk2(rs, rs);
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = mul2(x[k], rs2);
          }
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int n = nblk + q * 4;
            const bool nin = n < g.N && !(dbg.ablate & 2);
            float4* cell = reinterpret_cast<float4*>(iob + ((q ^ sw) << 4));
            if (HAS_AUX) {
              const float4 t = *cell;
              if (MODE == 2) { x[2 * q] = mul2(x[2 * q], pk2(t.x, t.y)); x[2 * q + 1] = mul2(x[2 * q + 1], pk2(t.z, t.w)); }
              else { x[2 * q] = add2(x[2 * q], pk2(t.x, t.y)); x[2 * q + 1] = add2(x[2 * q + 1], pk2(t.z, t.w)); }
            }
            float4 o;
            upk2(x[2 * q], o.x, o.y);
            upk2(x[2 * q + 1], o.z, o.w);
            *cell = o;
            if (c2 >= 0 && nin) *reinterpret_cast<float4*>(g.C2 + (size_t)c2 * g.ldc2 + n) = o;
          }
          if (warp == 0 && lane == 0) tc_mark(dbg, 14);  // block computed and staged
          ptx::fence_proxy_async();  // generic smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            if (!(dbg.ablate & 2)) ptx::tma_store_2d_h(&tmC, nblk, m0r, io + (size_t)b * TC_IO_BYTES, g.hintC, polC);
            ptx::bulk_commit();
            if (HAS_AUX && NIO == 1) { ptx::bulk_wait_read0(); issue_aux(); }
          }
          __syncwarp();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
      }
      if (warp == 0 && lane == 0) tc_mark(dbg, 9);  // last store issued
      if (lane == 0) ptx::bulk_wait_all();
      if (warp == 0 && lane == 0) tc_mark(dbg, 10);  // stores complete
    } else {
    // thread = row after tcgen05.ld; a padded smem transpose turns that into 4 rows x 128 contiguous bytes per warp access.
    const int rq = warp & 3, half = warp >> 2;
    float* stg = stg_all + warp * 32 * TC_STG_LD;
    const int nblocks = (BN + 31) / 32;
    const int cq = lane & 7, rsub = lane >> 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int buf = it & 1;
      const int m_base = (tile / w.n_tiles) * TC_BM + rq * 32;
      const int n0 = (tile % w.n_tiles) * BN;
      // row-wise operands of this thread's 8 rows (rows rsub + 4 j)
      const float* ra[8];
      const float* rb[MODE == 1 ? 8 : 1];
      float rs[8];
      int c2row[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int m = m_base + rsub + 4 * j;
        const bool ok = m < M;
        c2row[j] = (ok && g.C2) ? g.c2idx[m] : -1;
        ra[j] = nullptr;
        if (MODE == 1) rb[j] = nullptr;
        rs[j] = 1.f;
        if (ok) {
          if (MODE == 1) { ra[j] = g.radd1 + (size_t)(g.ridx1 ? g.ridx1[m] : m) * g.ld1; rb[j] = g.radd2 + (size_t)(g.ridx2 ? g.ridx2[m] : m) * g.ld2; }
          if (MODE == 2) ra[j] = g.mul + (size_t)m * g.ldmul;
          if (MODE == 3) ra[j] = g.resid + (size_t)m * g.ldres;
          if (g.rowscale) rs[j] = g.rowscale[g.rsidx ? g.rsidx[m] : m];
        }
      }
      ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      for (int blk = half; blk < nblocks; blk += 2) {
        const int n = n0 + blk * 32 + cq * 4;
        const bool ncol = n < g.N && (blk * 32 + cq * 4) < BN && !(dbg.ablate & 2);  // N % 4 == 0 (checked on the host)
        // single-operand modes: all 8 row-operand loads (HBM for the residual) fly while TMEM is read and transposed
        float4 a1[8];
        if (MODE == 2 || MODE == 3) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            a1[j] = (ncol && ra[j]) ? *reinterpret_cast<const float4*>(ra[j] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ncol && g.bias) bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
        {
          float v[32];
          ptx::tmem_ld32(tmem_base + buf * 256 + ((uint32_t)(rq * 32) << 16) + blk * 32, v);  // warp-collective
          if (g.prescale) {  // here thread = accumulator row (TMEM lane)
            const int mrow = m_base + lane;
            const float ps = mrow < M ? g.prescale[mrow] : 1.f;
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] *= ps;
          }
#pragma unroll
          for (int q = 0; q < 8; q++)
            *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int jh = 0; jh < 2; jh++) {
          float4 p1[MODE == 1 ? 4 : 1], p2[MODE == 1 ? 4 : 1];
          if (MODE == 1) {
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {  // gathered rows are L1/L2 resident: four rows in flight suffice
              const int j = jh * 4 + jj;
              const bool in = ncol && ra[j];
              p1[jj] = in ? *reinterpret_cast<const float4*>(ra[j] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
              p2[jj] = in ? *reinterpret_cast<const float4*>(rb[j] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 4; jj++) {
            const int j = jh * 4 + jj;
            const int m = m_base + rsub + 4 * j;
            if (ncol && m < M) {
              float4 x = *reinterpret_cast<const float4*>(stg + (rsub + 4 * j) * TC_STG_LD + cq * 4);
              x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
              if (MODE == 1) {
                x.x += p1[jj].x + p2[jj].x; x.y += p1[jj].y + p2[jj].y; x.z += p1[jj].z + p2[jj].z; x.w += p1[jj].w + p2[jj].w;
              }
              if (g.act == 1) { x.x = silu_fast(x.x); x.y = silu_fast(x.y); x.z = silu_fast(x.z); x.w = silu_fast(x.w); }
              const float r = rs[j];
              x.x *= r; x.y *= r; x.z *= r; x.w *= r;
              if (MODE == 2) { x.x *= a1[j].x; x.y *= a1[j].y; x.z *= a1[j].z; x.w *= a1[j].w; }
              if (MODE == 3) { x.x += a1[j].x; x.y += a1[j].y; x.z += a1[j].z; x.w += a1[j].w; }
              *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = x;
              if (c2row[j] >= 0) *reinterpret_cast<float4*>(g.C2 + (size_t)c2row[j] * g.ldc2 + n) = x;
            }
          }
        }
        __syncwarp();  // staging is reused by the next block
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
      }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) tc_mark(dbg, 11);
  if (warp == TC_EPI_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
  if (dbg.ts && blockIdx.x == 0 && threadIdx.x == TC_EPI_WARPS * 32) { dbg.ts[12] = clock64(); dbg.ts[15] = tc_globaltimer(); }
}

inline size_t tc_smem_bytes(int BN, int stages, int raw, int nio) {
  const size_t epi = nio > 0 ? (size_t)TC_EPI_WARPS * nio * TC_IO_BYTES : (size_t)TC_STG_BYTES;
  return (size_t)stages * (2 * TC_A_PART + 2 * BN * TC_KC * 2) + (size_t)raw * TC_RAW_BYTES + epi +
         (3 * stages + 4 + 2 * (raw > 0 ? raw : 1) + 2 * TC_EPI_WARPS) * 8 + 16;
}

// cuTensorMapEncodeTiled through the runtime (no libcuda link)
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline tc_encode_fn tc_get_encoder() {
  static tc_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tc_encode_fn>(p);
  }
  return fn;
}
// fp32 [rows, cols] row-major with leading dimension ld (floats); box {box_cols, box_rows}
inline bool tc_make_map(CUtensorMap* tm, const float* base, int rows, int cols, int ld, int box_cols, int box_rows,
                        bool swizzle128) {
  tc_encode_fn enc = tc_get_encoder();
  if (!enc || !base || (reinterpret_cast<uintptr_t>(base) & 15) || ((size_t)ld * 4) % 16 || rows <= 0 || cols <= 0)
    return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Requirements (checked): K % 4 == 0, lda % 4 == 0, N % 4 == 0, 16-byte aligned operands, BN % 16 == 0, BN <= 256,
// at most one of {radd1+radd2, mul, resid}.
template <int STAGES, int RAW, int MODE, int NIO>
inline cudaError_t launch_gemm_tc_inst(const GemmArgs& g, const TcWeight& w, int grid, size_t smem, TcDebugOpts dbg,
                                       const CUtensorMap& tmA, const CUtensorMap& tmC, const CUtensorMap& tmX,
                                       cudaStream_t st) {
  static PerDeviceOnce attr;  // per instantiation
  if (attr.first_time()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<STAGES, RAW, MODE, NIO>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  gemm_tc_kernel<STAGES, RAW, MODE, NIO><<<grid, TC_THREADS, smem, st>>>(g, w, dbg, tmA, tmC, tmX);
  return cudaGetLastError();
}

// ablate bits 4/5 (16, 32) force the LSU paths for the A operand / the epilogue (for A/B comparisons)
inline cudaError_t launch_gemm_tc(const GemmArgs& g, const TcWeight& w, int num_sms, cudaStream_t st,
                                  int swap_lbo_sbo = 0, int ablate = 0, long long* ts = nullptr) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.K % 4 || g.lda % 4 || g.N % 4 || g.ldc % 4 || w.BN % 16 || w.BN > 256 || w.BN < 16 || g.K != w.K || g.N != w.N)
    return cudaErrorInvalidValue;
  const int nmode = (g.radd1 ? 1 : 0) + (g.mul ? 1 : 0) + (g.resid ? 1 : 0);
  if (nmode > 1 || (g.radd1 && !g.radd2) || (!g.radd1 && g.radd2)) return cudaErrorInvalidValue;
  const int mode = g.radd1 ? 1 : (g.mul ? 2 : (g.resid ? 3 : 0));
  const int m_tiles = (g.M + TC_BM - 1) / TC_BM;
  const int total = m_tiles * w.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  {  // OARD_TC_ABLATE=<bits> ORs ablation bits into every launch (16: LSU A operand, 32: LSU epilogue) for A/B debugging
    static int env_bits = -1;
    if (env_bits < 0) { const char* e = getenv("OARD_TC_ABLATE"); env_bits = e ? atoi(e) : 0; }
    ablate |= env_bits;
  }
  TcDebugOpts dbg{swap_lbo_sbo, ablate, ts};
  CUtensorMap tmA, tmC, tmX;
  memset(&tmA, 0, sizeof tmA); memset(&tmC, 0, sizeof tmC); memset(&tmX, 0, sizeof tmX);
  const bool atma = !g.aidx && !(ablate & 16) && tc_make_map(&tmA, g.A, g.M, g.K, g.lda, TC_KC, TC_BM, false);
  bool etma = !(ablate & 32) && tc_make_map(&tmC, g.C, g.M, g.N, g.ldc, 32, 32, true);
  if (etma && mode == 2) etma = tc_make_map(&tmX, g.mul, g.M, g.N, g.ldmul, 32, 32, true);
  if (etma && mode == 3) etma = tc_make_map(&tmX, g.resid, g.M, g.N, g.ldres, 32, 32, true);
  const size_t lim = 227 * 1024;
  // (stages, raw, nio) candidates in order of preference
  int cand[4][3];
  int nc = 0;
  auto add = [&](int s_, int r_, int n_) { cand[nc][0] = s_; cand[nc][1] = r_; cand[nc][2] = n_; nc++; };
  if (etma && mode >= 2) { if (atma) { add(3, 2, 2); add(2, 2, 2); } else { add(3, 0, 2); add(2, 0, 2); } }
  else if (etma)         { if (atma) { add(3, 3, 1); add(3, 2, 1); add(2, 2, 1); } else { add(4, 0, 1); add(3, 0, 1); } }
  else                   { if (atma) { add(3, 3, 0); add(3, 2, 0); } else { add(4, 0, 0); add(3, 0, 0); } }
  int stages = 0, raw = 0, nio = 0;
  for (int i = 0; i < nc; i++)
    if (tc_smem_bytes(w.BN, cand[i][0], cand[i][1], cand[i][2]) <= lim) { stages = cand[i][0]; raw = cand[i][1]; nio = cand[i][2]; break; }
  if (!stages) return cudaErrorInvalidValue;
  const size_t smem = tc_smem_bytes(w.BN, stages, raw, nio);
#define OARD_TC_CASE(S, R, MD, NI) \
  if (stages == S && raw == R && mode == MD && nio == NI) \
    return launch_gemm_tc_inst<S, R, MD, NI>(g, w, grid, smem, dbg, tmA, tmC, tmX, st);
#define OARD_TC_AUX(S, R, NI) OARD_TC_CASE(S, R, 2, NI) OARD_TC_CASE(S, R, 3, NI)
#define OARD_TC_PLAIN(S, R, NI) OARD_TC_CASE(S, R, 0, NI) OARD_TC_CASE(S, R, 1, NI)
  OARD_TC_AUX(3, 2, 2) OARD_TC_AUX(2, 2, 2) OARD_TC_AUX(3, 0, 2) OARD_TC_AUX(2, 0, 2)
  OARD_TC_PLAIN(3, 3, 1) OARD_TC_PLAIN(3, 2, 1) OARD_TC_PLAIN(2, 2, 1) OARD_TC_PLAIN(4, 0, 1) OARD_TC_PLAIN(3, 0, 1)
  OARD_TC_AUX(3, 3, 0) OARD_TC_AUX(3, 2, 0) OARD_TC_AUX(4, 0, 0) OARD_TC_AUX(3, 0, 0)
  OARD_TC_PLAIN(3, 3, 0) OARD_TC_PLAIN(3, 2, 0) OARD_TC_PLAIN(4, 0, 0) OARD_TC_PLAIN(3, 0, 0)
#undef OARD_TC_PLAIN
#undef OARD_TC_AUX
#undef OARD_TC_CASE
  return cudaErrorInvalidValue;
}

}  // namespace oard
