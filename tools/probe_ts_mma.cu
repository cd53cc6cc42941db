// Probe for DESIGN §11.1, open point (1): tcgen05.mma with the A operand in TENSOR MEMORY.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_ts tools/probe_ts_mma.cu && timeout 60 /tmp/probe_ts
// One CTA, 128 threads.  A [128 x 16] (bf16) is written into TMEM by the four warps with tcgen05.st (thread = row / TMEM
// lane, two bf16 per 32-bit column: hypothesis (pack = 0) k even in the LOW half; pack = 1 tries the other order), B [32 x 16]
// sits in shared memory in the K-major SWIZZLE_NONE core-matrix layout the library's W operand uses, D [128 x 32] fp32 comes
// back with tcgen05.ld.  Prints the error against the host reference for both packings; "OK" on the one the hardware uses.
// Written without access to a GPU (it assembles for sm_100a); the layout hypotheses are what the probe is for.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 32, K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ inline uint32_t idesc(int m, int n) {  // D = f32, A = B = bf16, K-major, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                 int pack) {
  __shared__ __align__(128) __nv_bfloat16 Bs[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B -> smem, core matrices of 8 rows x 16 bytes: element (n, k) at ((n/8) * (K/8) + k/8) * 128 + (n%8) * 16 + (k%8) * 2 bytes
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    Bs[((n / 8) * (K / 8) + k / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2bfloat16(B[n * K + k]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of Bs -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t_d = tmem_base, t_a = tmem_base + 32;  // D: columns 0..31, A: columns 32..39
  // A row `tid` -> TMEM lane `tid`, 8 columns of packed bf16 pairs
  uint32_t r[8];
  for (int j = 0; j < 8; j++) {
    const __nv_bfloat16 lo = __float2bfloat16(A[tid * K + 2 * j + (pack ? 1 : 0)]);
    const __nv_bfloat16 hi = __float2bfloat16(A[tid * K + 2 * j + (pack ? 0 : 1)]);
    r[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
  }
  const uint32_t lane_addr = t_a + ((uint32_t)(warp * 32) << 16);  // warp w owns lanes 32 w .. 32 w + 31
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t bdesc = smem_desc(smem_u32(Bs), 128, (K / 8) * 128);  // LBO: K-adjacent cores, SBO: 8-row groups
    const uint32_t id = idesc(M, N);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(t_d),
        "r"(t_a), "l"(bdesc), "r"(id), "r"(0)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {  // wait for the MMA (bounded: a probe must not hang the box)
    uint32_t done = 0;
    for (int spin = 0; !done && spin < (1 << 22); spin++)
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0)
          : "memory");
    if (!done && tid == 0) printf("MMA did not complete within the spin budget\n");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t o[32];
  const uint32_t d_addr = t_d + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
        "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]), "=r"(o[17]), "=r"(o[18]),
        "=r"(o[19]), "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]), "=r"(o[25]), "=r"(o[26]), "=r"(o[27]),
        "=r"(o[28]), "=r"(o[29]), "=r"(o[30]), "=r"(o[31])
      : "r"(d_addr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; n++) D[tid * N + n] = __uint_as_float(o[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 500.0f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 500.0f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int pack = 0; pack < 2; pack++) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128>>>(dA, dB, dD, pack);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pack %d: CUDA error %s\n", pack, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        double ref = 0;
        for (int k = 0; k < K; k++) ref += (double)bf16r(A[m * K + k]) * (double)bf16r(B[n * K + k]);
        worst = fmax(worst, fabs(ref - D[m * N + n]));
        scale = fmax(scale, fabs(ref));
      }
    printf("pack %d (k even in the %s half): max |D - ref| = %.3e of %.3e  %s\n", pack, pack ? "HIGH" : "LOW", worst, scale,
           worst < 1e-3 * scale ? "OK" : "mismatch");
  }
  return 0;
}
