// umma_rate.cu -- tcgen05.mma issue-rate probe: dependent vs independent accumulators, SS vs TS.
// The issuing warp keeps every operand warp-uniform (uniform registers) and elects one lane.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../soc_matching_b200/csrc/umma.cuh"
using namespace socm::umma;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int M, int N, int A_TMEM, int N_ACC>
__global__ void __launch_bounds__(128) rate_kernel(int n_outer, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  for (int i = tid; i < 50 * 1024; i += 128) ((float*)smem)[i] = 0.001f * (i % 97);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = idesc_tf32(M, N, 0, 0);
    const uint32_t a_base = smem_addr(smem), b_base = smem_addr(smem + 64 * 1024);
    const long long t0 = clock64();
    for (int o = 0; o < n_outer; ++o) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t bd = smem_desc(b_base + (i & 7) * 8192, 128, 256);
          const uint32_t d = tb + (i % N_ACC) * N;
          if (A_TMEM) mma_ts(d, tb + 256 + i * 8, bd, idesc, 1);
          else mma_ss(d, smem_desc(a_base + (i & 7) * 4096, 128, 256), bd, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int M, int N, int A_TMEM, int N_ACC>
void run(long long* dout) {
  const int n_outer = 256;
  auto k = rate_kernel<M, N, A_TMEM, N_ACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 128, 200 * 1024>>>(n_outer, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long cyc; cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
  const double n_mma = 16.0 * n_outer;
  printf("M=%3d A=%s N=%3d acc=%d : %.1f cyc/MMA  %.0f MAC/clk\n", M, A_TMEM ? "TMEM" : "SMEM", N, N_ACC,
         cyc / n_mma, M * N * 8.0 * n_mma / cyc);
}

int main() {
  long long* dout; cudaMalloc(&dout, 8 * 256);
  run<128, 16, 0, 1>(dout);  run<128, 16, 1, 1>(dout);
  run<128, 32, 0, 1>(dout);  run<128, 32, 1, 1>(dout);
  run<128, 64, 0, 1>(dout);  run<128, 64, 1, 1>(dout);  run<128, 64, 0, 2>(dout);  run<128, 64, 1, 2>(dout); run<128, 64, 1, 4>(dout);
  run<128, 128, 0, 1>(dout); run<128, 128, 1, 1>(dout); run<128, 128, 0, 2>(dout); run<128, 128, 1, 2>(dout);
  run<128, 256, 0, 1>(dout); run<128, 256, 1, 1>(dout); run<128, 256, 0, 2>(dout);
  run<64, 64, 0, 1>(dout);   run<64, 128, 0, 1>(dout);  run<64, 256, 0, 1>(dout);  run<64, 256, 1, 1>(dout);
  return 0;
}
