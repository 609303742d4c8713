// mma_rate2.cu — microbenchmark: tcgen05.mma.cta_group::2 issue/execute rate on a CTA pair
// (M = 256) for the shapes the CTA-pair MIPS kernel can use: A from TMEM vs shared memory,
// N = 64/128/256, one accumulator chain vs two alternating. No TMA; operands are zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I retrieval_augmented_mds_b200/csrc -o scripts/mma_rate2.bin scripts/mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

struct Cfg {
  int n;        // MMA N (both CTAs together)
  int a_tmem;   // 1: A from TMEM, 0: A from smem, 2: 2/3 TMEM + 1/3 smem (the kernel's mix for d=768)
  int chains;
  int rounds;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate2_kernel(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc_pair(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish_pair();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(256, c.n);
    const uint64_t bdesc = ptx::smem_desc_sw128(base);              // B half: up to 128 rows x 64 k
    const uint64_t adesc = ptx::smem_desc_sw128(base + 64 * 1024);  // A (SS mode): 128 rows x 64 k
    const uint32_t a_t = tmem + 256;
    long long t0 = 0, t1 = 0;
    __syncwarp();
    if (ptx::elect_one()) t0 = clock64();
    for (int r = 0; r < c.rounds; ++r) {
      if (ptx::elect_one()) {
#pragma unroll 1
        for (int kk = 0; kk < 12; ++kk) {
          const bool ts = c.a_tmem == 1 || (c.a_tmem == 2 && kk < 8);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int step = kk * 4 + j;
            const uint32_t d = tmem + ((c.chains == 2 && (step & 1)) ? c.n : 0);
            if (ts)
              ptx::mma_bf16_ts_pair(d, a_t + j * 8, bdesc + 2u * j, idesc, step > 1 ? 1u : 0u);
            else
              ptx::mma_bf16_ss_pair(d, adesc + 2u * j, bdesc + 2u * j, idesc, step > 1 ? 1u : 0u);
          }
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit_pair(ptx::smem_u32(&bar));
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    if (ptx::elect_one()) {
      t1 = clock64();
      out_cycles[blockIdx.x >> 1] = t1 - t0;
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem, 512);
  }
}

int main() {
  cudaSetDevice(0);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * sms);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Cfg cfgs[] = {{64, 1, 1, 2000},  {64, 1, 2, 2000}, {128, 1, 1, 1000}, {128, 1, 2, 1000}, {256, 1, 1, 500},
                      {64, 0, 1, 2000},  {128, 0, 1, 1000}, {128, 0, 2, 1000}, {256, 0, 1, 500},  {256, 0, 2, 500},
                      {128, 2, 1, 1000}, {256, 2, 1, 500}};
  for (int grid : {2, sms}) {
    for (const Cfg& c : cfgs) {
      mma_rate2_kernel<<<grid, 128, smem>>>(c, d_out);  // warm-up
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      mma_rate2_kernel<<<grid, 128, smem>>>(c, d_out);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(err));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long h[256];
      cudaMemcpy(h, d_out, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      const double n_mma = 48.0 * c.rounds;
      const double cyc = mx / n_mma;
      const double ideal = 128.0 * c.n / 256.0;   // per SM: 128 x N x 16 MACs at 4096 MAC/clk
      const double tflops = 2.0 * 256 * c.n * 16 * n_mma * (grid / 2) / (ms * 1e-3) / 1e12;
      printf("pairs %3d  N=%3d  A=%s  chains=%d : %7.1f cycles/MMA (nominal %5.1f, %5.1f%%)  %8.1f TFLOP/s  %.3f ms\n",
             grid / 2, c.n, c.a_tmem == 1 ? "tmem" : c.a_tmem == 0 ? "smem" : "mix ", c.chains, cyc, ideal,
             100.0 * ideal / cyc, tflops, ms);
    }
  }
  return 0;
}
