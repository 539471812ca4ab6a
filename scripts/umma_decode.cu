// umma_decode.cu -- decodes which shared-memory address tcgen05.mma reads for element (mn, k) of an
// MN-major operand: the operand region is filled with float(word index) and multiplied by an identity.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../soc_matching_b200/csrc/umma.cuh"
using namespace socm::umma;

// mode 0: A MN-major (decode A), B K-major identity (N=16): D[m][n] = A(m, k=n) n<8
// mode 1: B MN-major (decode B), A K-major identity: D[m][n] = B(n, k=m) for m<8
__global__ void __launch_bounds__(128) decode_kernel(int mode, int lbo, int sbo, int N, int ltype, float* D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* region = (float*)smem;                 // decoded operand: 64 KB of word indices
  float* ident = (float*)(smem + 64 * 1024);    // K-major identity operand, rows x 8 k: RG=256 (2 core matrices), CG=128
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  for (int i = tid; i < 16 * 1024; i += 128) region[i] = (float)i;
  for (int i = tid; i < 16 * 1024; i += 128) ident[i] = 0.f;
  __syncthreads();
  const int rows = mode == 0 ? N : 128;
  for (int i = tid; i < rows * 8; i += 128) {
    const int r = i / 8, k = i % 8;
    ident[((r % 8) * 16 + (k % 4) * 4 + (r / 8) * 256 + (k / 4) * 128) / 4] = (r == k) ? 1.f : 0.f;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot;
  if (tid == 0) {
    const uint64_t idd = smem_desc(smem_addr(ident), 128, 256);
    const uint64_t rd = smem_desc(smem_addr(region), lbo, sbo) | ((uint64_t)ltype << 61);
    if (mode == 0) mma_ss(tb, rd, idd, idesc_tf32(128, N, 1, 0), 0);
    else mma_ss(tb, idd, rd, idesc_tf32(128, N, 0, 1), 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  float* dD; cudaMalloc(&dD, 128 * 256 * 4);
  float* h = (float*)malloc(128 * 256 * 4);
  cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfgs[][3] = {{4096, 1024, 1}, {1024, 4096, 1}, {8192, 512, 1}, {4096, 1024, 2}, {4096, 1024, 6}};
  for (int mode = 0; mode < 2; ++mode)
    for (auto& c : cfgs) {
      const int N = mode == 0 ? 16 : 64;
      cudaMemset(dD, 0, 128 * 256 * 4);
      decode_kernel<<<1, 128, 200 * 1024>>>(mode, c[0], c[1], N, c[2], dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
      printf("mode %d (decode %s MN-major) LBO=%d SBO=%d layout_type=%d : byte offset read for (mn, k)\n", mode, mode == 0 ? "A" : "B", c[0], c[1], c[2]);
      const int mns[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 16, 17, 20, 24, 28, 31, 32, 33, 36, 63, 64, 65, 96, 127};
      for (int mn : mns) {
        if (mode == 0 ? mn >= 128 : mn >= N) continue;
        printf("  mn %3d:", mn);
        for (int k = 0; k < 8; ++k) {
          const float v = mode == 0 ? h[mn * N + k] : h[k * N + mn];
          printf(" %6d", (int)v * 4);
        }
        printf("\n");
      }
    }
  return 0;
}
