// Microbenchmark: how fast can the accumulators be read out of TMEM? Cycles per tcgen05.ld.32x32b.x32 (one warp reads 32 lanes x
// 32 columns x 4 B = 4 KB per instruction) with 1, 2, 4 or 8 warps of one CTA reading at the same time (warp w reads lane quarter
// w % 4, as the hardware prescribes), each instruction followed by its own tcgen05.wait::ld or with four loads in flight per wait.
// This is the roofline of the filter on SHORT rows: a 128 x 128 fp32 tile is 64 KB to read and 512 cycles of MMA at K = 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read_rate tmem_read_rate.cu && ./tmem_read_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}

template <int BATCH>   // loads in flight per wait
__global__ void __launch_bounds__(256, 1) tmem_read_kernel(int iters, int active_warps, long long* out, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    const uint32_t lane_addr = ((uint32_t)(warp & 3) * 32u) << 16;
    const uint32_t col0 = (uint32_t)(warp >> 2) * 256u;   // the two warps of a quarter read different columns
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float v[BATCH][32];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) ld32(tmem + lane_addr + col0 + (uint32_t)((i * BATCH + b) & 7) * 32u, v[b]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int b = 0; b < BATCH; ++b) acc += v[b][0] + v[b][31];
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int BATCH>
static void run(int warps) {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 8 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 4096 / BATCH;
  tmem_read_kernel<BATCH><<<1, 256>>>(16, warps, d_out, d_sink);
  tmem_read_kernel<BATCH><<<1, 256>>>(iters, warps, d_out, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  long long worst = 0;
  for (int w = 0; w < warps; ++w) worst = h[w] > worst ? h[w] : worst;
  const double loads = (double)iters * BATCH;
  printf("{\"loads_in_flight\": %d, \"warps\": %d, \"cycles_per_ld_x32_per_warp\": %.1f, \"bytes_per_clk_per_SM\": %.1f, \"cuda\": \"%s\"}\n", BATCH, warps,
         worst / loads, warps * loads * 4096.0 / worst, cudaGetErrorString(e));
  cudaFree(d_out);
  cudaFree(d_sink);
}

int main() {
  for (int w : {1, 2, 4, 8}) run<1>(w);
  for (int w : {1, 2, 4, 8}) run<4>(w);
  return 0;
}
