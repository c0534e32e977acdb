// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair, K = 16) as a function of N and
// of where A comes from, back-to-back from one thread of the leader CTA, operands resident (no TMA).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N, bool A_TMEM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate2_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  unsigned char* sb = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sb)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t adesc = make_desc(smem_u32(sb));
    const uint64_t bdesc = make_desc(smem_u32(sb + 16384));
    const uint32_t d_tmem = tmem;
    const uint32_t a_tmem = tmem + 256;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (A_TMEM)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
                       "r"(a_tmem + k * 8), "l"(bdesc + (uint64_t)(k * 2)), "r"(idesc), "r"(1u)
                       : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                       "l"(adesc + (uint64_t)(k * 2)), "l"(bdesc + (uint64_t)(k * 2)), "r"(idesc), "r"(1u)
                       : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                 "h"((uint16_t)1)
                 : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool A_TMEM>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 4096;
  cudaFuncSetAttribute(mma_rate2_kernel<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  mma_rate2_kernel<N, A_TMEM><<<grid, 128, 80 * 1024>>>(64, d);
  cudaDeviceSynchronize();
  mma_rate2_kernel<N, A_TMEM><<<grid, 128, 80 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 4.0);
  printf("pair %-18s grid %3d: %7.1f cycles per MMA (ideal at 4096 MAC/clk/SM: %d)  -> %.0f MAC/clk/SM  %s\n", name, grid, per, N / 2,
         128.0 * N * 16 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {2, 148}) {
    run<64, true>("A=TMEM N=64", grid);
    run<128, true>("A=TMEM N=128", grid);
    run<256, true>("A=TMEM N=256", grid);
    run<128, false>("A=smem N=128", grid);
    run<256, false>("A=smem N=256", grid);
  }
  return 0;
}
