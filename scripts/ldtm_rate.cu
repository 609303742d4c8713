// ldtm_rate.cu — microbenchmark: tcgen05.ld 32x32b throughput (TMEM -> registers) with 4 warps,
// alone and while the tensor pipe is busy. Decides whether a single 128-column accumulator can be
// drained fast enough.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

__global__ void __launch_bounds__(192, 1) ldtm_kernel(int rounds, int ncols, int with_mma, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); stop = 0; }
  if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (warp < 4) {
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(lane_addr + c, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(v[j]));
      }
    }
    long long t1 = clock64();
    if (lane == 0 && warp == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
    __syncwarp();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&stop) = 1;
  } else if (warp == 5 && with_mma) {
    const uint32_t idesc = ptx::idesc_bf16_f32(128, 64);
    const uint64_t bdesc = ptx::smem_desc_sw128(base);
    while (*reinterpret_cast<volatile int*>(&stop) == 0) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) ptx::mma_bf16_ts(tmem + 256, tmem + 384 + (j & 3) * 8, bdesc + 2u * (j & 3), idesc, 1u);
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bar));
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, sizeof(long long) * 256); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(ldtm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int with_mma : {0, 1}) for (int ncols : {64, 128}) {
    const int rounds = 2000;
    ldtm_kernel<<<148, 192, 80 * 1024>>>(rounds, ncols, with_mma, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per_round = (double)mx / rounds;
    printf("mma=%d  drain %3d cols x 128 lanes (%3d KiB): %7.1f cycles  -> %6.1f B/clk/SM\n", with_mma, ncols,
           ncols * 128 * 4 / 1024, per_round, ncols * 128 * 4 / per_round);
  }
  return 0;
}
