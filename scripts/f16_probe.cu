// f16_probe.cu -- hardware probe for the kind::f16 (fp16 inputs, fp32 accumulate) form of tcgen05.mma:
//   * operand conventions: SS K-major (no swizzle, 8 x 16-byte core matrices = 8 rows x 8 halfs) and TS with the
//     A operand in tensor memory (two halfs per 32-bit column: which half is the even k?),
//   * accumulation behaviour over a long K (truncation bias) next to kind::tf32,
//   * issue rate for N = 64 / 128 / 144 / 256, A in TMEM or SMEM,
//   * two co-resident CTAs per SM with 256 TMEM columns each: do their MMA streams interleave?
//   * tcgen05.ld throughput with 8 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16_probe scripts/f16_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../soc_matching_b200/csrc/umma.cuh"

using namespace socm::umma;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}

// ---------------------------------------------------------------- convention check
// mode 0: SS (A smem K-major).  mode 1: TS, low half = even k.  mode 2: TS, high half = even k.
// A[128][K], B[N][K] as floats that are exactly representable in fp16.
__global__ void __launch_bounds__(128) conv_kernel(int mode, int N, int K, const float* __restrict__ A,
                                                   const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sa = smem;
  unsigned char* sb = smem + 64 * 1024;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot, d_t = tb, a_t = tb + 256;
  // K-major no-swizzle, halfs: elem(r, k) -> (r%8)*16 + (k%8)*2 + (k/8)*128 + (r/8)*(K/8)*128   (LBO = 128, SBO = K*16)
  const int sbo = K * 16;
  if (mode == 0) {
    for (int i = tid; i < 128 * K; i += 128) {
      const int r = i / K, k = i % K;
      *(__half*)(sa + (r % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (r / 8) * sbo) = __float2half_rn(A[i]);
    }
  } else {
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) {
        const unsigned short e = __half_as_ushort(__float2half_rn(A[tid * K + k0 + 2 * j]));
        const unsigned short o = __half_as_ushort(__float2half_rn(A[tid * K + k0 + 2 * j + 1]));
        r[j] = mode == 1 ? ((uint32_t)o << 16 | e) : ((uint32_t)e << 16 | o);
      }
      tmem_st16(a_t + ((uint32_t)(warp * 32) << 16) + k0 / 2, r);
    }
    tmem_wait_st();
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    *(__half*)(sb + (r % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (r / 8) * sbo) = __float2half_rn(B[i]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    const uint32_t id = idesc_f16(128, N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bd = smem_desc(smem_addr(sb) + ks * 256, 128, sbo);
      if (mode == 0) mma_ss_f16(d_t, smem_desc(smem_addr(sa) + ks * 256, 128, sbo), bd, id, ks > 0);
      else mma_ts_f16(d_t, a_t + ks * 8, bd, id, ks > 0);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld16(d_t + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static float to_h(float x) { return __half2float(__float2half_rn(x)); }

static void run_conv(const char* name, int mode, int N, int K, int positive) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(99);
  for (auto& v : A) v = to_h(positive ? (float)rand() / RAND_MAX + 0.5f : (float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = to_h(positive ? (float)rand() / RAND_MAX + 0.5f : (float)rand() / RAND_MAX * 2.f - 1.f);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  conv_kernel<<<1, 128, 200 * 1024>>>(mode, N, K, dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-34s : CUDA ERROR %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double emax = 0, rmax = 0, bias = 0, relsum = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      emax = fmax(emax, fabs(D[m * N + n] - s));
      rmax = fmax(rmax, fabs(s));
      bias += (D[m * N + n] - s) / s;
      relsum += fabs((D[m * N + n] - s) / s);
    }
  printf("%-34s : max|D-ref| %.3e  (max|ref| %.2f)  mean rel err %+.3e  mean |rel err| %.3e\n", name, emax, rmax,
         bias / (128 * N), relsum / (128 * N));
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

// tf32 long-K accumulation for comparison (A in TMEM, same data rounded to tf32-exact values)
__global__ void __launch_bounds__(128) acc_tf32_kernel(int N, int K, const float* __restrict__ A,
                                                       const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sb = smem;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot, d_t = tb, a_t = tb + 256;
  // process K in pieces of 128 (A piece in TMEM columns [256, 384), B piece in smem)
  uint32_t phase = 0;
  for (int kp = 0; kp < K; kp += 128) {
    for (int k0 = 0; k0 < 128; k0 += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A[tid * K + kp + k0 + j]);
      tmem_st16(a_t + ((uint32_t)(warp * 32) << 16) + k0, r);
    }
    tmem_wait_st();
    for (int i = tid; i < N * 128; i += 128) {
      const int r = i / 128, k = i % 128;
      *(float*)(sb + (r % 8) * 16 + (k % 4) * 4 + (k / 4) * 128 + (r / 8) * (128 * 32)) = B[r * K + kp + k];
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
      const uint32_t id = idesc_tf32(128, N, 0, 0);
      for (int ks = 0; ks < 16; ++ks)
        mma_ts(d_t, a_t + ks * 8, smem_desc(smem_addr(sb) + ks * 256, 128, 128 * 32), id, (kp > 0 || ks > 0));
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    __syncthreads();
  }
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld16(d_t + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}
__global__ void __launch_bounds__(128) acc_f16_kernel(int N, int K, const float* __restrict__ A,
                                                      const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sb = smem;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot, d_t = tb, a_t = tb + 256;
  uint32_t phase = 0;
  for (int kp = 0; kp < K; kp += 128) {
    for (int k0 = 0; k0 < 128; k0 += 32) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) {
        const unsigned short e = __half_as_ushort(__float2half_rn(A[tid * K + kp + k0 + 2 * j]));
        const unsigned short o = __half_as_ushort(__float2half_rn(A[tid * K + kp + k0 + 2 * j + 1]));
        r[j] = (uint32_t)o << 16 | e;
      }
      tmem_st16(a_t + ((uint32_t)(warp * 32) << 16) + k0 / 2, r);
    }
    tmem_wait_st();
    for (int i = tid; i < N * 128; i += 128) {
      const int r = i / 128, k = i % 128;
      *(__half*)(sb + (r % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (r / 8) * (128 * 16)) = __float2half_rn(B[r * K + kp + k]);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
      const uint32_t id = idesc_f16(128, N);
      for (int ks = 0; ks < 8; ++ks)
        mma_ts_f16(d_t, a_t + ks * 8, smem_desc(smem_addr(sb) + ks * 256, 128, 128 * 16), id, (kp > 0 || ks > 0));
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    __syncthreads();
  }
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld16(d_t + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static void run_acc(int K, int positive) {
  const int N = 64;
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(7);
  // values with <= 10 mantissa bits, exact in fp16 and in tf32
  for (auto& v : A) v = to_h(positive ? (float)rand() / RAND_MAX + 0.5f : (float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = to_h(positive ? (float)rand() / RAND_MAX + 0.5f : (float)rand() / RAND_MAX * 2.f - 1.f);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  for (int which = 0; which < 2; ++which) {
    if (which == 0) {
      CK(cudaFuncSetAttribute(acc_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      acc_tf32_kernel<<<1, 128, 100 * 1024>>>(N, K, dA, dB, dD);
    } else {
      CK(cudaFuncSetAttribute(acc_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      acc_f16_kernel<<<1, 128, 100 * 1024>>>(N, K, dA, dB, dD);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double bias = 0, rms = 0, nref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
        bias += (D[m * N + n] - s);
        rms += (D[m * N + n] - s) * (D[m * N + n] - s);
        nref += s * s;
      }
    printf("accumulate K=%5d %s %s : rel-l2 err %.3e   mean err / rms ref %+.3e\n", K, positive ? "positive" : "signed  ",
           which ? "f16 (K=16/MMA)" : "tf32 (K=8/MMA)", sqrt(rms / nref), bias / (128 * N) / sqrt(nref / (128 * N)));
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

// ---------------------------------------------------------------- rate, optionally with idle gaps; NCOLS TMEM columns
template <int N, int A_TMEM>
__global__ void __launch_bounds__(128) rate_kernel(int n_outer, int gap, int ncols, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, ncols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  for (int i = tid; i < 20 * 1024; i += 128) ((float*)smem)[i] = 0.f;
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot;
  if (warp == 0) {
    constexpr uint32_t id = idesc_f16(128, N);
    const uint32_t a_base = smem_addr(smem), b_base = smem_addr(smem + 32 * 1024);
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int o = 0; o < n_outer; ++o) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t bd = smem_desc(b_base + (i & 3) * 8192, 128, 256);
          if (A_TMEM) mma_ts_f16(tb, tb + ncols / 2 + (i & 7) * 8, bd, id, 1);
          else mma_ss_f16(tb, smem_desc(a_base + (i & 3) * 4096, 128, 256), bd, id, 1);
        }
        if (gap) commit(&bar);
      }
      __syncwarp();
      if (gap) {
        mbar_wait(&bar, ph);
        ph ^= 1;
        const long long t = clock64();
        while (clock64() - t < gap) {
        }
      }
    }
    if (!gap) {
      if (elect_one()) commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
    }
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, ncols);
}

template <int N, int A_TMEM>
static void run_rate(long long* dout) {
  const int n_outer = 256;
  auto k = rate_kernel<N, A_TMEM>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  k<<<1, 128, 100 * 1024>>>(n_outer, 0, 512, dout);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
  const double n_mma = 16.0 * n_outer;
  printf("rate f16 A=%s N=%3d : %.1f cyc/MMA (128xNx16)  %.0f MAC/clk\n", A_TMEM ? "TMEM" : "SMEM", N, cyc / n_mma,
         128.0 * N * 16 * n_mma / cyc);
}

// ---------------------------------------------------------------- tcgen05.ld throughput: 8 warps read 256 columns `rounds` times
__global__ void __launch_bounds__(256) ldtm_kernel(int rounds, int width, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (width == 32) {
      for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(tb + c, v);
        tmem_wait_ld();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
      }
    } else {
      for (int c = 0; c < 256; c += 16) {
        uint32_t v[16];
        tmem_ld16(tb + c, v);
        tmem_wait_ld();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  sink[tid] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  run_conv("SS f16 K-major N=64 K=64", 0, 64, 64, 0);
  run_conv("SS f16 K-major N=256 K=128", 0, 256, 128, 0);
  run_conv("TS f16 low half = even k N=64 K=64", 1, 64, 64, 0);
  run_conv("TS f16 high half = even k N=64 K=64", 2, 64, 64, 0);
  run_conv("TS f16 low half = even k N=256 K=128", 1, 256, 128, 0);
  for (int K : {256, 1024, 4096}) {
    run_acc(K, 1);
    run_acc(K, 0);
  }
  long long* dout;
  CK(cudaMalloc(&dout, 8 * 1024));
  run_rate<64, 1>(dout);
  run_rate<128, 1>(dout);
  run_rate<256, 1>(dout);
  run_rate<64, 0>(dout);
  run_rate<96, 0>(dout);
  run_rate<128, 0>(dout);
  run_rate<144, 0>(dout);
  run_rate<256, 0>(dout);
  // co-residency: bursts of 16 MMAs (N = 128: 16 x 64 = 1024 cycles) followed by a 1024-cycle gap
  {
    auto k = rate_kernel<128, 1>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int ctas : {prop.multiProcessorCount, 2 * prop.multiProcessorCount}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      k<<<ctas, 128, 100 * 1024>>>(64, 1024, 256, dout);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      k<<<ctas, 128, 100 * 1024>>>(4096, 1024, 256, dout);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc;
      CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
      printf("co-residency: %3d CTAs x 256 TMEM columns, 100 KB smem, burst+gap: %.3f ms, CTA 0: %.0f cycles per burst+gap\n",
             ctas, ms, (double)cyc / 4096);
    }
  }
  {
    float* sink;
    CK(cudaMalloc(&sink, 4096));
    for (int width : {16, 32}) {
      ldtm_kernel<<<1, 256>>>(200, width, dout, sink);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
      printf("tcgen05.ld x%d, 8 warps, 512 columns x 128 lanes per round: %.0f cycles per round -> %.1f B/clk\n", width,
             (double)cyc / 200, 512.0 * 128 * 4 * 200 / cyc);
    }
  }
  printf("f16 probe done\n");
  return 0;
}
