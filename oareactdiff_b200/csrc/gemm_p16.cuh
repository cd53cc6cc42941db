// tcgen05 GEMM whose A operand is already stored in HBM in the tensor core's own split-bf16 form ("pair16").
//
// pair16 layout of a row-major activation X[M, K] (row pitch ld floats, ld % 16 == 0, same bytes as fp32):
//   every 16 consecutive values x[16g .. 16g+15] occupy 64 bytes:  [16 x bf16 hi | 16 x bf16 lo],  x ~= hi + lo
//   (hi = bf16_rn(x), lo = bf16_rn(x - hi): 16 mantissa bits, exactly what the bf16x3 product consumes anyway).
// A 2-D TMA box {32 floats = 128 B, 128 rows} with SWIZZLE_128B therefore lands a ready K-major UMMA operand for a
// K chunk of 32: k-step j uses the sub-atom offsets  hi = 64 j,  lo = 64 j + 32  of every 128-byte row.  No producer
// warps, no in-kernel fp32 -> bf16 conversion, no extra shared-memory copy (compare gemm_tc.cuh, whose per-chunk
// convert/fence skeleton bounded it at 30-45 % tensor-pipe activity: profiles/r1_tc_notes.md).
// The epilogue (thread = output row, 32x32 blocks, TMA in/out) can emit pair16 as well, so chains of edge GEMMs
// (edge state -> GCL edge MLP -> edge state -> dir_proj) never touch fp32 in HBM.
//
//   C[m, n] = epi( sum_k A[m, k] * W[n, k] ),  three tcgen05.mma.kind::f16 per K step (a_hi w_hi + a_hi w_lo + a_lo w_hi)
//
// Warp roles (EW + 3 warps, 1 CTA/SM, persistent over tiles):
//   warps 0..EW-1 epilogue (EW = 8 or 16; warp e owns TMEM lanes 32 (e % 4).. and column blocks e/4, e/4 + EW/4, ..)
//   warp  EW      MMA issuer (one lane), owns TMEM alloc/dealloc, 2 accumulators
//   warp  EW+1    A loader (one lane): 2-D TMA boxes into an SA-deep ring (the HBM stream: deep, 16 KB per slot)
//   warp  EW+2    W loader (one lane): TMA bulk copies of the pre-tiled weight slabs into an SW-deep ring (L2-resident)
// The two rings are separate because the A stream comes from HBM (needs ~64 KB in flight per SM to cover the latency)
// while W comes from L2: with one shared ring the 26-32 KB W slab of every stage capped the A bytes in flight at 32 KB.
#pragma once
#include "gemm_tc.cuh"

namespace oard {

constexpr int P16_A_BYTES = TC_BM * 128;  // one A stage: 128 rows x 128 B

// ---- pair16 encode / decode of 8 consecutive values (one 16-byte cell of hi and one of lo)
__device__ __forceinline__ void p16_split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 f = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - f.x, x[2 * i + 1] - f.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void p16_join8(const uint4& hi, const uint4& lo, float* x) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    x[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
    x[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
  }
}
// byte offsets of value column c inside a pair16 row
__host__ __device__ inline size_t p16_off_hi(int c) { return (size_t)(c >> 4) * 64 + (size_t)(c & 15) * 2; }
__host__ __device__ inline int p16_ld(int K) { return (K + 15) / 16 * 16; }

// fp32 [M, K] (ld_src) -> pair16 [M, p16_ld(K)] (ld_dst floats); pad columns are written as zeros.  One thread per 8 values.
// m_dev (optional): the row count lives in device memory (M is then the cap).
__global__ void k_p16_pack(const float* __restrict__ src, int ld_src, int M, int K, float* __restrict__ dst, int ld_dst,
                           const int* __restrict__ m_dev = nullptr) {
  const int cells = ld_dst / 8;
  if (m_dev) M = min(M, *m_dev);
  const size_t total = (size_t)M * cells;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / cells), c8 = (int)(i % cells) * 8;
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = (c8 + k) < K ? src[(size_t)m * ld_src + c8 + k] : 0.f;
    uint4 hi, lo;
    p16_split8(x, hi, lo);
    uint8_t* row = reinterpret_cast<uint8_t*>(dst + (size_t)m * ld_dst);
    *reinterpret_cast<uint4*>(row + p16_off_hi(c8)) = hi;
    *reinterpret_cast<uint4*>(row + p16_off_hi(c8) + 32) = lo;
  }
}
// pair16 [M, ld_src] -> fp32 [M, K] (ld_dst)
__global__ void k_p16_unpack(const float* __restrict__ src, int ld_src, int M, int K, float* __restrict__ dst, int ld_dst) {
  const int cells = (K + 7) / 8;
  const size_t total = (size_t)M * cells;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / cells), c8 = (int)(i % cells) * 8;
    const uint8_t* row = reinterpret_cast<const uint8_t*>(src + (size_t)m * ld_src);
    const uint4 hi = *reinterpret_cast<const uint4*>(row + p16_off_hi(c8));
    const uint4 lo = *reinterpret_cast<const uint4*>(row + p16_off_hi(c8) + 32);
    float x[8];
    p16_join8(hi, lo, x);
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c8 + k < K) dst[(size_t)m * ld_dst + c8 + k] = x[k];
  }
}


// ---- CTA pairs (tcgen05 cta_group::2): one MMA spans two SMs of a TPC.  Each CTA holds its own 128 rows of A and HALF of
// the weight slab (N/2 rows); the leader (cluster rank 0) issues the MMAs for both, D rows 0-127 land in the leader's tensor
// memory and rows 128-255 in the peer's.  The weight bytes every SM pulls from L2 halve (the slab stream, not the tensor
// pipe, bounds these kernels: profiles/r2_experiments.md).
namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// 2-D tensor load whose completion bytes are counted on an mbarrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster,
                                                 int kind, uint64_t pol) {
  if (kind == 0) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar_cluster)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar_cluster), "l"(pol)
        : "memory");
  }
}
}  // namespace ptx

// K-major SWIZZLE_128B shared-memory descriptor: LBO field 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B),
// version 1 (bit 46), layout type 2 = SWIZZLE_128B (bits 61-63).  The start address may carry a 32/64/96-byte sub-atom
// offset (K advance inside the swizzle atom); the stage base is 1024-byte aligned.
__device__ __forceinline__ uint64_t p16_a_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

__device__ __forceinline__ void p16_mark(const GemmArgs& g, bool on, int i) {
  if (g.ts && on && blockIdx.x == 0) g.ts[i] = clock64();
}

// MODE: 0 plain (bias / SiLU / row scale), 1 + gathered row adds P[src] + Q[dst], 2 * mul[m, n] (fp32 aux),
//       3 + resid[m, n] (aux in pair16 when OUT_PAIR, else fp32; may alias C).
// OUT_PAIR: C (and the optional row-scattered copy C2) are written in pair16, else fp32.
// CTAS: 1, or 2 = CTA pairs (256-row tiles, launched as clusters of two; the weight slabs then arrive as 2-D tensor loads of
// 256-byte rows through tmW, each CTA taking its half of the hi and of the lo part).
template <int SA, int SW, int MODE, bool OUT_PAIR, int EW, int NIO, int CTAS>
__global__ void __launch_bounds__((EW + 3) * 32, 1)
gemm_p16_kernel(const GemmArgs g, const TcWeight w, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NCG = EW / 4;  // column groups of epilogue warps
  const int BN = w.BN;
  const int W_PART = BN * TC_KC * 2;                           // hi (or lo) part of a whole slab
  const int W_MINE = W_PART / CTAS;                            // the rows of it this CTA holds
  uint8_t* a_ring = smem;                                      // [SA][16 KB], 1024-aligned slots
  uint8_t* w_ring = smem + (size_t)SA * P16_A_BYTES;           // [SW][hi (BN / CTAS) x 64 B | lo (BN / CTAS) x 64 B]
  uint8_t* io_all = w_ring + (size_t)SW * 2 * W_MINE;          // 1024-aligned (BN % 16 == 0, BN % 32 == 0 for pairs)
  uint64_t* bars = reinterpret_cast<uint64_t*>(io_all + (size_t)EW * NIO * TC_IO_BYTES);
  uint64_t* full_a = bars;                 // [SA] A box landed (complete_tx)
  uint64_t* empty_a = full_a + SA;         // [SA] MMAs that read the slot retired (tcgen05.commit)
  uint64_t* full_w = empty_a + SA;         // [SW]
  uint64_t* empty_w = full_w + SW;         // [SW]
  uint64_t* acc_full = empty_w + SW;       // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* aux_full = acc_empty + 2;      // [EW][NIO]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + EW * NIO);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_chunks = w.k_chunks;
  const int k16_total = (g.K + 15) / 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; s++) { ptx::mbar_init(&full_a[s], 1); ptx::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < SW; s++) { ptx::mbar_init(&full_w[s], 1); ptx::mbar_init(&empty_w[s], 1); }
    for (int b = 0; b < 2; b++) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], EW * CTAS); }
    for (int i = 0; i < EW * NIO; i++) ptx::mbar_init(&aux_full[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == EW) { if (CTAS == 2) ptx::tmem_alloc2(tmem_slot, 512); else ptx::tmem_alloc(tmem_slot, 512); }
  ptx::tc_fence_before();
  if (CTAS == 2) ptx::cluster_sync_all(); else __syncthreads();  // (pairs: the peer's barriers exist before anything signals them)
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  constexpr int BMT = TC_BM * CTAS;  // rows per tile of the pair
  const int rank = CTAS == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit = blockIdx.x / CTAS, nunits = gridDim.x / CTAS;  // persistent workers: CTAs or CTA pairs
  const int m_tiles = (M + BMT - 1) / BMT;
  const int total_tiles = m_tiles * w.n_tiles;

  if (warp == EW + 1) {
    // ===================== A loader: 2-D TMA boxes {32 floats, 128 rows}, SWIZZLE_128B =====================
    if (lane == 0) {
      const uint64_t polA = ptx::l2_policy(g.hintA);
      ptx::tma_prefetch_desc(&tmA);
      uint32_t gchunk = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits) {
        const int m0 = (tile / w.n_tiles) * BMT + rank * TC_BM;
        for (int kc = 0; kc < k_chunks; kc++, gchunk++) {
          const int s = gchunk % SA;
          ptx::mbar_wait(&empty_a[s], ((gchunk / SA) & 1) ^ 1);
          if (g.ablate & 1) { if (rank == 0) ptx::mbar_arrive(&full_a[s]); continue; }
          if (CTAS == 2) {  // both boxes of the pair are counted on the leader's barrier
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_a[s], 2 * P16_A_BYTES);
            ptx::tma_load_2d_pair(a_ring + (size_t)s * P16_A_BYTES, &tmA, kc * TC_KC, m0,
                                  ptx::mapa_u32(ptx::smem_u32(&full_a[s]), 0), g.hintA, polA);
          } else {
            ptx::mbar_arrive_expect_tx(&full_a[s], P16_A_BYTES);
            ptx::tma_load_2d_h(a_ring + (size_t)s * P16_A_BYTES, &tmA, kc * TC_KC, m0, &full_a[s], g.hintA, polA);
          }
        }
      }
    }
  } else if (warp == EW + 2) {
    // ===================== W loader: TMA bulk copies of pre-tiled slabs =====================
    if (lane == 0) {
      uint32_t gchunk = 0;
      if (CTAS == 2) ptx::tma_prefetch_desc(&tmW);
      for (int tile = unit; tile < total_tiles; tile += nunits) {
        const int nt = tile % w.n_tiles;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w.data) + (size_t)nt * k_chunks * 2 * W_PART;
        for (int kc = 0; kc < k_chunks; kc++, gchunk++) {
          const int s = gchunk % SW;
          ptx::mbar_wait(&empty_w[s], ((gchunk / SW) & 1) ^ 1);
          if (g.ablate & 4) { if (rank == 0) ptx::mbar_arrive(&full_w[s]); continue; }
          if (CTAS == 2) {
            // the slab as rows of 256 bytes: this CTA's half of the hi part and of the lo part, counted on the leader's barrier
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_w[s], 2 * W_PART);
            const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&full_w[s]), 0);
            const int row_hi = (int)((((size_t)nt * k_chunks + kc) * 2 * W_PART + (size_t)rank * W_MINE) >> 8);
            uint8_t* dst = w_ring + (size_t)s * 2 * W_MINE;
            ptx::tma_load_2d_pair(dst, &tmW, 0, row_hi, bar, 0, 0);
            ptx::tma_load_2d_pair(dst + W_MINE, &tmW, 0, row_hi + (W_PART >> 8), bar, 0, 0);
          } else {
            ptx::mbar_arrive_expect_tx(&full_w[s], 2 * W_PART);
            ptx::bulk_g2s(w_ring + (size_t)s * 2 * W_PART, wsrc + (size_t)kc * 2 * W_PART, 2 * W_PART, &full_w[s]);
          }
        }
      }
    }
  } else if (warp == EW) {
    // ===================== MMA issuer (warp-uniform; the tcgen05 instructions are issued by one elected lane) ==========
    if (rank == 0) {
      const uint32_t idesc = tc_idesc(BMT, BN);
      uint32_t gchunk = 0, it = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, it++) {
        const int buf = it & 1;
        ptx::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int kc = 0; kc < k_chunks; kc++, gchunk++) {
          const int sa = gchunk % SA, sw_ = gchunk % SW;
          ptx::mbar_wait(&full_w[sw_], (gchunk / SW) & 1);
          ptx::mbar_wait(&full_a[sa], (gchunk / SA) & 1);
          ptx::tc_fence_after();
          const uint32_t a0 = ptx::smem_u32(a_ring + (size_t)sa * P16_A_BYTES);
          const uint32_t w_hi = ptx::smem_u32(w_ring + (size_t)sw_ * 2 * W_MINE), w_lo = w_hi + W_MINE;
          const int steps = (g.ablate & 8) ? 0 : min(TC_KC / 16, k16_total - kc * (TC_KC / 16));
          // descriptors of K step 0; step 1 is a constant further (the address field counts 16-byte units)
          const uint64_t dah0 = p16_a_desc(a0), dal0 = p16_a_desc(a0 + 32);
          const uint64_t dwh0 = tc_smem_desc(w_hi, TC_CORE_BYTES, TC_SBO), dwl0 = tc_smem_desc(w_lo, TC_CORE_BYTES, TC_SBO);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < TC_KC / 16; j++) {
              if (j >= steps) break;
              const uint64_t dah = dah0 + (uint64_t)(j * (64 >> 4)), dal = dal0 + (uint64_t)(j * (64 >> 4));
              const uint64_t dwh = dwh0 + (uint64_t)(j * ((2 * TC_CORE_BYTES) >> 4)), dwl = dwl0 + (uint64_t)(j * ((2 * TC_CORE_BYTES) >> 4));
              if (CTAS == 2) {
                ptx::umma_bf16_2cta(d_tmem, dah, dwh, idesc, (kc | j) != 0);
                ptx::umma_bf16_2cta(d_tmem, dah, dwl, idesc, 1);
                ptx::umma_bf16_2cta(d_tmem, dal, dwh, idesc, 1);
              } else {
                ptx::umma_bf16(d_tmem, dah, dwh, idesc, (kc | j) != 0);
                ptx::umma_bf16(d_tmem, dah, dwl, idesc, 1);
                ptx::umma_bf16(d_tmem, dal, dwh, idesc, 1);
              }
            }
            // both rings are released (in both CTAs of a pair) when these MMAs retire
            if (CTAS == 2) { ptx::umma_commit_pair(&empty_a[sa]); ptx::umma_commit_pair(&empty_w[sw_]); }
            else { ptx::umma_commit(&empty_a[sa]); ptx::umma_commit(&empty_w[sw_]); }
          }
          __syncwarp();
        }
        if (ptx::elect_one()) {
          if (CTAS == 2) ptx::umma_commit_pair(&acc_full[buf]); else ptx::umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps: thread = output row, 32x32 blocks, TMA in / out =====================
    const int rq = warp & 3, cg = warp >> 2;
    uint8_t* io = io_all + (size_t)warp * NIO * TC_IO_BYTES;
    uint64_t* xbar = aux_full + warp * NIO;
    const int nblocks = (BN + 31) / 32;
    const int sw = lane & 7;
    constexpr bool HAS_AUX = (MODE == 2 || MODE == 3);
    const uint64_t polX = ptx::l2_policy(g.hintX), polC = ptx::l2_policy(g.hintC);
    // Buffer ring per warp: block i uses buffer i % NIO from its aux load until its store has read it.  Aux loads run PF
    // blocks ahead; the buffer of block nb + PF was last read by the store of block nb + PF - NIO, so at most
    // PEND = NIO - 1 - PF newer stores may still be reading when it is refilled (NIO 2: PF 1, PEND 0; NIO 3: PF 1, PEND 1).
    static_assert(NIO >= 1 && NIO <= 3 && (NIO >= 2 || !HAS_AUX), "NIO");
    constexpr int PF = 1, PEND = HAS_AUX ? NIO - 1 - PF : 0;
    int pf_tile = unit, pf_blk = cg;
    auto pf_next = [&]() {
      pf_blk += NCG;
      if (pf_blk >= nblocks) { pf_blk = cg; pf_tile += nunits; }
    };
    uint32_t nb = 0, npf = 0;
    auto issue_aux = [&]() {  // lane 0 only
      if (!HAS_AUX) return;
      while (pf_tile < total_tiles && pf_blk >= nblocks) pf_next();
      if (pf_tile >= total_tiles) return;
      const int b = npf % NIO;
      if (g.ablate & 2) { ptx::mbar_arrive(&xbar[b]); npf++; pf_next(); return; }
      ptx::mbar_arrive_expect_tx(&xbar[b], TC_IO_BYTES);
      ptx::tma_load_2d_h(io + (size_t)b * TC_IO_BYTES, &tmX, (pf_tile % w.n_tiles) * BN + pf_blk * 32,
                         (pf_tile / w.n_tiles) * BMT + rank * TC_BM + rq * 32, &xbar[b], g.hintX, polX);
      npf++;
      pf_next();
    };
    if (lane == 0) {
      ptx::tma_prefetch_desc(&tmC);
      if (HAS_AUX) {
        ptx::tma_prefetch_desc(&tmX);
        for (int i = 0; i < PF; i++) issue_aux();
      }
    }
    uint32_t it = 0;
    const uint32_t acc_empty_leader = CTAS == 2 ? ptx::mapa_u32(ptx::smem_u32(&acc_empty[0]), 0) : 0;
    for (int tile = unit; tile < total_tiles; tile += nunits, it++) {
      const int buf = it & 1;
      const int m0r = (tile / w.n_tiles) * BMT + rank * TC_BM + rq * 32;
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
      const int c2 = (ok && g.C2 && !(g.ablate & 2)) ? g.c2idx[m] : -1;
      ptx::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      for (int blk = cg; blk < nblocks; blk += NCG, nb++) {
        const int b = nb % NIO;
        uint8_t* iob = io + (size_t)b * TC_IO_BYTES + lane * 128;
        float v[32];
        ptx::tmem_ld32(tmem_base + buf * 256 + ((uint32_t)(rq * 32) << 16) + blk * 32, v);  // warp-collective
        if (HAS_AUX) {
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PEND) : "memory");
            issue_aux();
          }
          ptx::mbar_wait(&xbar[b], (nb / NIO) & 1);
        } else {
          // buffer b was last read by the store of block nb - NIO: NIO - 1 newer stores may still be in flight
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NIO - 1) : "memory");
          __syncwarp();
        }
        const int nblk = n0 + blk * 32;
        // The block's 32 values of this row as 16 packed pairs (FFMA2 / FMUL2 / FADD2: the epilogue is an instruction-issue
        // chain per warp — with every load and MMA switched off edge_out still took 91 of its 155 us, 16 warps x ~770
        // instructions per block saturate the four schedulers — so its arithmetic runs two values per instruction).
        f32x2 x[16];
        {
          const f32x2 ps2 = pk2(ps, ps);  // (ps = 1 without a prescale: fma(v, 1, b) = v + b exactly)
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int n = nblk + q * 4;
            const bool nin = n < g.N;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias && nin) t = __ldg(reinterpret_cast<const float4*>(g.bias + n));
            f32x2 t0 = pk2(t.x, t.y), t1 = pk2(t.z, t.w);
            if (MODE == 1 && nin && ok && !(g.ablate & 2)) {
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
        if (MODE == 2 || (MODE == 3 && !OUT_PAIR)) {  // fp32 aux block: cell q = columns 4q .. 4q+3
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 t = *reinterpret_cast<const float4*>(iob + ((q ^ sw) << 4));
            if (MODE == 2) { x[2 * q] = mul2(x[2 * q], pk2(t.x, t.y)); x[2 * q + 1] = mul2(x[2 * q + 1], pk2(t.z, t.w)); }
            else { x[2 * q] = add2(x[2 * q], pk2(t.x, t.y)); x[2 * q + 1] = add2(x[2 * q + 1], pk2(t.z, t.w)); }
          }
        }
        if (OUT_PAIR) {
          // row bytes: cells 0,1 = hi(v0..15), 2,3 = lo(v0..15), 4,5 = hi(v16..31), 6,7 = lo(v16..31)
          uint4 cell[8];
          // hi cell of values 8 g .. 8 g + 7: (g & 1) + 4 (g >> 1); its lo cell is + 2
          if (MODE == 3) {
#pragma unroll
            for (int q = 0; q < 8; q++) cell[q] = *reinterpret_cast<const uint4*>(iob + ((q ^ sw) << 4));
#pragma unroll
            for (int gq = 0; gq < 4; gq++) {
              const uint4 hi = cell[((gq & 1) + 4 * (gq >> 1))], lo = cell[((gq & 1) + 4 * (gq >> 1)) + 2];
              const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const f32x2 eh = pk2(__uint_as_float(hw[i] << 16), __uint_as_float(hw[i] & 0xffff0000u));
                const f32x2 el = pk2(__uint_as_float(lw[i] << 16), __uint_as_float(lw[i] & 0xffff0000u));
                x[gq * 4 + i] = add2(x[gq * 4 + i], add2(eh, el));
              }
            }
          }
          const f32x2 neg1 = pk2(-1.0f, -1.0f);
#pragma unroll
          for (int gq = 0; gq < 4; gq++) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              float a, bb;
              upk2(x[gq * 4 + i], a, bb);
              const __nv_bfloat162 hh = __floats2bfloat162_rn(a, bb);
              const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh);
              const f32x2 back = pk2(__uint_as_float(hu << 16), __uint_as_float(hu & 0xffff0000u));
              float ra, rb2;
              upk2(fma2(back, neg1, x[gq * 4 + i]), ra, rb2);  // x - hi, exact
              const __nv_bfloat162 ll = __floats2bfloat162_rn(ra, rb2);
              hw[i] = hu;
              lw[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            cell[((gq & 1) + 4 * (gq >> 1))] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            cell[((gq & 1) + 4 * (gq >> 1)) + 2] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
#pragma unroll
          for (int q = 0; q < 8; q++) {
            *reinterpret_cast<uint4*>(iob + ((q ^ sw) << 4)) = cell[q];
            if (c2 >= 0 && (nblk + (q >> 2) * 16) < g.ldc2)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(g.C2 + (size_t)c2 * g.ldc2 + nblk) + q * 16) = cell[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; q++) {
            float4 o;
            upk2(x[2 * q], o.x, o.y);
            upk2(x[2 * q + 1], o.z, o.w);
            *reinterpret_cast<float4*>(iob + ((q ^ sw) << 4)) = o;
            if (c2 >= 0 && (nblk + q * 4) < g.N) *reinterpret_cast<float4*>(g.C2 + (size_t)c2 * g.ldc2 + nblk + q * 4) = o;
          }
        }
        ptx::fence_proxy_async();  // generic smem writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (!(g.ablate & 2)) ptx::tma_store_2d_h(&tmC, nblk, m0r, io + (size_t)b * TC_IO_BYTES, g.hintC, polC);
          ptx::bulk_commit();
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) ptx::mbar_arrive_cluster(acc_empty_leader + buf * 8); else ptx::mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  if (CTAS == 2) ptx::cluster_sync_all(); else __syncthreads();  // (pairs: no CTA leaves while its peer may still signal it)
  if (warp == EW) {
    ptx::tc_fence_after();
    if (CTAS == 2) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

inline size_t p16_smem_bytes(int BN, int sa, int sw, int ew, int nio, int ctas = 1) {
  return (size_t)sa * P16_A_BYTES + (size_t)sw * 2 * BN * TC_KC * 2 / ctas + (size_t)ew * nio * TC_IO_BYTES +
         (size_t)(2 * sa + 2 * sw + 4 + ew * nio) * 8 + 16;
}

template <int SA, int SW, int MODE, bool OUT_PAIR, int EW, int NIO, int CTAS>
inline cudaError_t launch_gemm_p16_inst(const GemmArgs& g, const TcWeight& w, int grid, size_t smem, const CUtensorMap& tmA,
                                        const CUtensorMap& tmC, const CUtensorMap& tmX, const CUtensorMap& tmW,
                                        cudaStream_t st) {
  static PerDeviceOnce attr;  // per instantiation
  auto kern = gemm_p16_kernel<SA, SW, MODE, OUT_PAIR, EW, NIO, CTAS>;
  if (attr.first_time()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (CTAS == 1) {
    kern<<<grid, (EW + 3) * 32, smem, st>>>(g, w, tmA, tmC, tmX, tmW);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((EW + 3) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, g, w, tmA, tmC, tmX, tmW);
}

// A: pair16 [M, lda] with lda % 16 == 0 and lda >= K.  out_pair: C (ldc % 16 == 0, ldc >= N) and C2 in pair16, and the
// residual (mode 3) is read as pair16; otherwise C / resid are fp32.  mul (mode 2) is always fp32.
// ew_pref: 8 or 16 epilogue warps (16 only for aux modes); returns cudaErrorInvalidValue for unsupported shapes.
// CTA pairs (cta_group::2, 256-row tiles) are used when every SM gets at least one 128-row tile and K >= 512;
// OARD_P16_CTAS=1|2 overrides.
inline cudaError_t launch_gemm_p16(const GemmArgs& g, const TcWeight& w, int num_sms, cudaStream_t st, bool out_pair,
                                   int ew_pref = 0, int ctas_pref = 0) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.K % 4 || g.lda % 16 || g.lda < g.K || g.N % 4 || g.ldc % 4 || w.BN % 16 || w.BN > 256 || w.BN < 16 ||
      g.K != w.K || g.N != w.N || g.aidx)
    return cudaErrorInvalidValue;
  if (out_pair && (g.ldc % 16 || g.ldc < g.N || (g.C2 && g.ldc2 != g.ldc))) return cudaErrorInvalidValue;
  const int nmode = (g.radd1 ? 1 : 0) + (g.mul ? 1 : 0) + (g.resid ? 1 : 0);
  if (nmode > 1 || (g.radd1 && !g.radd2) || (!g.radd1 && g.radd2)) return cudaErrorInvalidValue;
  const int mode = g.radd1 ? 1 : (g.mul ? 2 : (g.resid ? 3 : 0));
  if (mode == 2 && out_pair) return cudaErrorInvalidValue;
  if (mode == 3 && out_pair && g.ldres != g.ldc) return cudaErrorInvalidValue;
  static int env_ew = -1, env_ctas = -1;
  if (env_ew < 0) { const char* e = getenv("OARD_P16_EW"); env_ew = e ? atoi(e) : 0; }
  if (env_ctas < 0) { const char* e = getenv("OARD_P16_CTAS"); env_ctas = e ? atoi(e) : 0; }
  // Epilogue shape.  The epilogue is a per-warp latency chain (TMEM load -> activation -> split -> staging -> TMA store,
  // ~700 instructions per 32x32 block at low ILP; ncu r1o), so the cure is more warps: 16 epilogue warps.  Plain modes
  // (no streamed aux block) need one staging buffer per warp (NIO 1); the in-place residual update keeps two (the aux
  // block of the next output block is in flight while the current one is processed).  The multiplier mode measured
  // better with 8 warps and deeper operand rings.  OARD_P16_EW=8|16 overrides (A/B runs).
  int ew = ew_pref ? ew_pref : (mode == 2 ? 8 : 16);
  if (!ew_pref && (env_ew == 8 || env_ew == 16)) ew = env_ew;
  int nio = (mode < 2 && ew == 16) ? 1 : 2;
  // CTA pairs: only the epilogue shapes the edge-level GEMMs of the path use are instantiated for them
  const bool pair_shape = (mode == 0 && ew == 16) || (mode == 1 && out_pair && ew == 16) || (mode == 2 && ew == 8) ||
                          (mode == 3 && out_pair && ew == 16);
  // Pairs pay where the weight slab is the larger operand stream (K >= 512: edge_mlp layer 1, dir_proj: -9 ... -10 % per launch);
  // with K = 196 (edge_mlp layer 2, edge_out_trans) the halved slab does not make up for the pair's coupling (+5 ... +9 %).
  int ctas = (((g.M + TC_BM - 1) / TC_BM) * w.n_tiles >= num_sms && g.K >= 512) ? 2 : 1;
  if (env_ctas == 1 || env_ctas == 2) ctas = env_ctas;
  if (ctas_pref == 1 || ctas_pref == 2) ctas = ctas_pref;
  if (!pair_shape || (num_sms & 1) || num_sms < 2) ctas = 1;
  const size_t lim = 227 * 1024;
  // ring depths (A, W) in order of preference
  static const int rings1[3][2] = {{5, 3}, {4, 3}, {2, 2}};
  static const int rings2[3][2] = {{5, 4}, {4, 3}, {3, 3}};
  int sa = 0, sw = 0;
  auto pick = [&](int e_, int ct) {
    const int (*rings)[2] = ct == 2 ? rings2 : rings1;
    for (int i = 0; i < 3; i++)
      if (p16_smem_bytes(w.BN, rings[i][0], rings[i][1], e_, nio, ct) <= lim) { sa = rings[i][0]; sw = rings[i][1]; return true; }
    return false;
  };
  if (ctas == 2 && !pick(ew, 2)) ctas = 1;
  if (ctas == 1 && !pick(ew, 1)) {
    ew = 8; nio = 2;
    if (!pick(ew, 1)) return cudaErrorInvalidValue;
  }
  CUtensorMap tmA, tmC, tmX, tmW;
  memset(&tmA, 0, sizeof tmA); memset(&tmC, 0, sizeof tmC); memset(&tmX, 0, sizeof tmX); memset(&tmW, 0, sizeof tmW);
  if (!tc_make_map(&tmA, g.A, g.M, p16_ld(g.K), g.lda, TC_KC, TC_BM, true)) return cudaErrorInvalidValue;
  if (!tc_make_map(&tmC, g.C, g.M, out_pair ? p16_ld(g.N) : g.N, g.ldc, 32, 32, true)) return cudaErrorInvalidValue;
  if (mode == 2 && !tc_make_map(&tmX, g.mul, g.M, g.N, g.ldmul, 32, 32, true)) return cudaErrorInvalidValue;
  if (mode == 3 && !tc_make_map(&tmX, g.resid, g.M, out_pair ? p16_ld(g.N) : g.N, g.ldres, 32, 32, true))
    return cudaErrorInvalidValue;
  if (ctas == 2) {  // the packed weight as rows of 256 bytes; one box = this CTA's half of a hi or lo part (BN x 32 bytes)
    const size_t wbytes = tc_weight_elems(w.N, w.K, w.BN) * 2;
    if (wbytes % 256 || !tc_make_map(&tmW, reinterpret_cast<const float*>(w.data), (int)(wbytes / 256), 64, 64, 64, w.BN / 8, false))
      return cudaErrorInvalidValue;
  }
  const size_t smem = p16_smem_bytes(w.BN, sa, sw, ew, nio, ctas);
  const int m_tiles = (g.M + TC_BM * ctas - 1) / (TC_BM * ctas);
  const int total = m_tiles * w.n_tiles;
  const int units = total < num_sms / ctas ? total : num_sms / ctas;
  const int grid = units * ctas;
#define OARD_P16_CASE(A_, W_, MD, OP, E, NI, CT)                                                      \
  if (sa == A_ && sw == W_ && mode == MD && out_pair == OP && ew == E && nio == NI && ctas == CT)     \
    return launch_gemm_p16_inst<A_, W_, MD, OP, E, NI, CT>(g, w, grid, smem, tmA, tmC, tmX, tmW, st);
#define OARD_P16_RINGS(MD, OP, E, NI) \
  OARD_P16_CASE(5, 3, MD, OP, E, NI, 1) OARD_P16_CASE(4, 3, MD, OP, E, NI, 1) OARD_P16_CASE(2, 2, MD, OP, E, NI, 1)
#define OARD_P16_PAIRS(MD, OP, E, NI) \
  OARD_P16_CASE(5, 4, MD, OP, E, NI, 2) OARD_P16_CASE(4, 3, MD, OP, E, NI, 2) OARD_P16_CASE(3, 3, MD, OP, E, NI, 2)
  OARD_P16_RINGS(0, true, 8, 2) OARD_P16_RINGS(0, false, 8, 2) OARD_P16_RINGS(1, true, 8, 2)
  OARD_P16_RINGS(0, true, 16, 1) OARD_P16_RINGS(0, false, 16, 1) OARD_P16_RINGS(1, true, 16, 1)
  OARD_P16_RINGS(2, false, 8, 2) OARD_P16_RINGS(2, false, 16, 2)
  OARD_P16_RINGS(3, true, 8, 2) OARD_P16_RINGS(3, true, 16, 2) OARD_P16_RINGS(3, false, 8, 2)
  OARD_P16_PAIRS(0, true, 16, 1) OARD_P16_PAIRS(0, false, 16, 1) OARD_P16_PAIRS(1, true, 16, 1)
  OARD_P16_PAIRS(2, false, 8, 2) OARD_P16_PAIRS(3, true, 16, 2)
#undef OARD_P16_PAIRS
#undef OARD_P16_RINGS
#undef OARD_P16_CASE
  return cudaErrorInvalidValue;
}

}  // namespace oard
