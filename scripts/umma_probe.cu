// umma_probe.cu -- stand-alone hardware probe for the tcgen05 layer (csrc/umma.cuh):
//   * checks the shared-memory / tensor-memory operand conventions against a CPU GEMM
//     (SS K-major, TS with A in TMEM, SS MN-major both operands = the weight-gradient form),
//   * tells whether kind::tf32 truncates or rounds its fp32 inputs,
//   * measures the MMA issue rate, the L2 -> SMEM bulk-copy rate per SM and the
//     red.global.add.v4.f32 rate, the three numbers the kernel design in DESIGN.md rests on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe scripts/umma_probe.cu
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

enum { A_SMEM_K = 0, A_TMEM = 1, A_SMEM_MN = 2 };
enum { B_SMEM_K = 0, B_SMEM_MN = 1 };

struct Cfg {
  int N, K;          // M = 128
  int a_mode, b_mode;
  int a_rg, a_cg;    // byte strides of the A core-matrix grid (row group, column group) as stored
  int b_rg, b_cg;
  int swap_a, swap_b;  // swap LBO/SBO in the descriptor (to find the right convention)
};

// logical A[m][k] (m < 128), B[n][k]; D[m][n] = sum_k A*B
__global__ void __launch_bounds__(128) probe_kernel(Cfg c, const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128, N = c.N, K = c.K;
  unsigned char* sa = smem;
  unsigned char* sb = smem + 64 * 1024;
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_slot;
  const uint32_t d_t = tb;           // columns [0, 256)
  const uint32_t a_t = tb + 256;     // columns [256, 512)

  // ---- stage A
  if (c.a_mode == A_SMEM_K) {
    for (int i = tid; i < M * K; i += 128) {
      const int r = i / K, k = i % K;
      *(float*)(sa + (r % 8) * 16 + (k % 4) * 4 + (r / 8) * c.a_rg + (k / 4) * c.a_cg) = A[i];
    }
  } else if (c.a_mode == A_SMEM_MN) {  // storage rows = k, cols = m
    for (int i = tid; i < M * K; i += 128) {
      const int m = i / K, k = i % K;
      *(float*)(sa + (k % 8) * 16 + (m % 4) * 4 + (k / 8) * c.a_rg + (m / 4) * c.a_cg) = A[i];
    }
  } else {  // TMEM: lane = m, column = k; thread tid owns lane tid
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A[tid * K + k0 + j]);
      tmem_st16(a_t + ((uint32_t)(warp * 32) << 16) + k0, r);
    }
    tmem_wait_st();
  }
  // ---- stage B
  if (c.b_mode == B_SMEM_K) {
    for (int i = tid; i < N * K; i += 128) {
      const int r = i / K, k = i % K;
      *(float*)(sb + (r % 8) * 16 + (k % 4) * 4 + (r / 8) * c.b_rg + (k / 4) * c.b_cg) = B[i];
    }
  } else {
    for (int i = tid; i < N * K; i += 128) {
      const int n = i / K, k = i % K;
      *(float*)(sb + (k % 8) * 16 + (n % 4) * 4 + (k / 8) * c.b_rg + (n / 4) * c.b_cg) = B[i];
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(M, N, c.a_mode == A_SMEM_MN, c.b_mode == B_SMEM_MN);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t bd;
      if (c.b_mode == B_SMEM_K) {
        const uint32_t lbo = c.swap_b ? c.b_rg : c.b_cg, sbo = c.swap_b ? c.b_cg : c.b_rg;
        bd = smem_desc(smem_addr(sb) + ks * 2 * c.b_cg, lbo, sbo);
      } else {
        const uint32_t lbo = c.swap_b ? c.b_cg : c.b_rg, sbo = c.swap_b ? c.b_rg : c.b_cg;
        bd = smem_desc(smem_addr(sb) + ks * c.b_rg, lbo, sbo);
      }
      if (c.a_mode == A_TMEM) {
        mma_ts(d_t, a_t + ks * 8, bd, idesc, ks > 0);
      } else {
        uint64_t ad;
        if (c.a_mode == A_SMEM_K) {
          const uint32_t lbo = c.swap_a ? c.a_rg : c.a_cg, sbo = c.swap_a ? c.a_cg : c.a_rg;
          ad = smem_desc(smem_addr(sa) + ks * 2 * c.a_cg, lbo, sbo);
        } else {
          const uint32_t lbo = c.swap_a ? c.a_cg : c.a_rg, sbo = c.swap_a ? c.a_rg : c.a_cg;
          ad = smem_desc(smem_addr(sa) + ks * c.a_rg, lbo, sbo);
        }
        mma_ss(d_t, ad, bd, idesc, ks > 0);
      }
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

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rn_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x00000FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

static void run_cfg(const char* name, Cfg c) {
  const int M = 128, N = c.N, K = c.K;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1234);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 128, 200 * 1024>>>(c, dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s swapA=%d swapB=%d : CUDA ERROR %s\n", name, c.swap_a, c.swap_b, cudaGetErrorString(e));
    exit(2);
  }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double err_t = 0, err_r = 0, err_x = 0, ref_n = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sx = 0;
      for (int k = 0; k < K; ++k) {
        st += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
        sr += (double)rn_tf32(A[m * K + k]) * rn_tf32(B[n * K + k]);
        sx += (double)A[m * K + k] * B[n * K + k];
      }
      const double g = D[m * N + n];
      err_t = fmax(err_t, fabs(g - st));
      err_r = fmax(err_r, fabs(g - sr));
      err_x = fmax(err_x, fabs(g - sx));
      ref_n = fmax(ref_n, fabs(sx));
    }
  printf("%-28s swapA=%d swapB=%d : max|D-ref| trunc %.3e  rn %.3e  exact %.3e  (max|ref| %.2f)\n", name, c.swap_a,
         c.swap_b, err_t, err_r, err_x, ref_n);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

// ---------------------------------------------------------------- timing probes
// T1: back-to-back tcgen05.mma (A in TMEM or SMEM), N columns, n_mma instructions, cycles per MMA
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int n_mma, int a_tmem, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  for (int i = tid; i < 48 * 1024; i += 128) ((float*)smem)[i] = 0.001f * (i % 97);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 0, 0);
    const uint64_t bd = smem_desc(smem_addr(smem + 96 * 1024), 128, 256);
    const uint64_t ad = smem_desc(smem_addr(smem), 128, 256);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (a_tmem)
        mma_ts(tb, tb + 256 + (i & 15) * 8, bd, idesc, 1);
      else
        mma_ss(tb, ad, bd, idesc, 1);
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// T4: every CTA streams the same `tape_bytes` (L2 resident) `rounds` times through a ring of
// `stages` x `chunk` bytes with cp.async.bulk; nothing consumes the data.
__global__ void __launch_bounds__(128) stream_kernel(const unsigned char* tape, int tape_bytes, int chunk, int stages,
                                                     int rounds) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int per_round = tape_bytes / chunk;
    const long long total = (long long)per_round * rounds;
    for (long long i = 0; i < stages && i < total; ++i) {
      mbar_expect_tx(&full[i % stages], chunk);
      bulk_g2s(smem + (i % stages) * chunk, tape + (size_t)(i % per_round) * chunk, chunk, &full[i % stages]);
    }
    for (long long i = 0; i < total; ++i) {
      const int s = (int)(i % stages);
      mbar_wait(&full[s], (uint32_t)((i / stages) & 1));
      if (i + stages < total) {
        mbar_expect_tx(&full[s], chunk);
        bulk_g2s(smem + s * chunk, tape + (size_t)((i + stages) % per_round) * chunk, chunk, &full[s]);
      }
    }
  }
}

// T3: red.global.add.v4.f32: each CTA adds `floats` floats to its buffer (buffer index = cta % n_buf)
__global__ void __launch_bounds__(256) red_kernel(float* buf, int floats, int n_buf, int rounds) {
  float* b = buf + (size_t)(blockIdx.x % n_buf) * floats;
  for (int r = 0; r < rounds; ++r)
    for (int i = threadIdx.x * 4; i < floats; i += 256 * 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(b + i), "f"(1.0f) : "memory");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d smem_optin %zu clock %d kHz\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin, prop.clockRate);

  // --- convention checks
  for (int sw = 0; sw < 2; ++sw) {
    // SS, both K-major, A grid: row groups contiguous (RG=128) then k groups (CG=2048); B: CG=128, RG=K*32
    run_cfg("SS K-major N=64 K=32", Cfg{64, 32, A_SMEM_K, B_SMEM_K, 128, 2048, 32 * 32, 128, sw, sw});
  }
  run_cfg("SS K-major N=256 K=64", Cfg{256, 64, A_SMEM_K, B_SMEM_K, 128, 2048, 64 * 32, 128, 0, 0});
  run_cfg("TS A=TMEM N=256 K=128", Cfg{256, 128, A_TMEM, B_SMEM_K, 0, 0, 128 * 32, 128, 0, 0});
  run_cfg("TS A=TMEM N=64 K=16", Cfg{64, 16, A_TMEM, B_SMEM_K, 0, 0, 128, 1024, 0, 0});
  for (int sw = 0; sw < 2; ++sw) {
    // MN-major both: storage rows = k (points), cols = features; k groups contiguous (RG=128), col groups CG = (K/8)*128
    run_cfg("SS MN-major N=128 K=64", Cfg{128, 64, A_SMEM_MN, B_SMEM_MN, 128, 8 * 128, 128, 8 * 128, sw, sw});
  }
  run_cfg("SS MN-major N=256 K=128", Cfg{256, 128, A_SMEM_MN, B_SMEM_MN, 128, 16 * 128, 128, 16 * 128, 0, 0});
  run_cfg("SS A=K-major B=MN N=64 K=32", Cfg{64, 32, A_SMEM_K, B_SMEM_MN, 128, 2048, 128, 4 * 128, 0, 0});

  // --- T1: MMA rate
  long long* dout;
  CK(cudaMalloc(&dout, 8 * 256));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {64, 128, 256}) {
      const int n_mma = 4096;
      mma_rate_kernel<<<1, 128, 200 * 1024>>>(N, n_mma, a_tmem, dout);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost));
      printf("T1 mma rate: A=%s N=%3d : %.1f cycles / MMA (128xNx8 tf32) -> %.0f MAC/clk/SM\n",
             a_tmem ? "TMEM" : "SMEM", N, (double)cyc / n_mma, 128.0 * N * 8 * n_mma / cyc);
    }
  // all SMs at once (power / clock effects)
  {
    const int n_mma = 1 << 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mma_rate_kernel<<<prop.multiProcessorCount, 128, 200 * 1024>>>(256, n_mma, 1, dout);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    mma_rate_kernel<<<prop.multiProcessorCount, 128, 200 * 1024>>>(256, n_mma, 1, dout);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("T1 full chip TS N=256: %.3f ms -> %.1f TFLOP/s tf32 dense\n", ms,
           2.0 * 128 * 256 * 8 * n_mma * prop.multiProcessorCount / (ms * 1e-3) / 1e12);
  }

  // --- T4: L2 -> SMEM streaming
  {
    const int tape = 1310720;  // 163840 weights x 8 bytes
    unsigned char* dt;
    CK(cudaMalloc(&dt, tape));
    CK(cudaMemset(dt, 1, tape));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int chunk : {16384, 32768})
      for (int stages : {2, 4}) {
        const int rounds = 200;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        stream_kernel<<<prop.multiProcessorCount, 128, 200 * 1024>>>(dt, tape, chunk, stages, 10);
        cudaEventRecord(e0);
        stream_kernel<<<prop.multiProcessorCount, 128, 200 * 1024>>>(dt, tape, chunk, stages, rounds);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)tape * rounds * prop.multiProcessorCount;
        printf("T4 stream chunk %5d stages %d: %.3f ms  %.1f GB/s chip  %.1f GB/s per SM\n", chunk, stages, ms,
               bytes / ms / 1e6, bytes / ms / 1e6 / prop.multiProcessorCount);
      }
    // one SM alone
    {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      stream_kernel<<<1, 128, 200 * 1024>>>(dt, tape, 32768, 4, 200);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("T4 stream single SM: %.1f GB/s\n", (double)tape * 200 / ms / 1e6);
    }
  }

  // --- T3: red.global.add.v4.f32
  {
    const int floats = 170624;
    float* db;
    CK(cudaMalloc(&db, (size_t)floats * 4 * prop.multiProcessorCount));
    CK(cudaMemset(db, 0, (size_t)floats * 4 * prop.multiProcessorCount));
    for (int n_buf : {prop.multiProcessorCount, 37, 8, 1}) {
      const int rounds = 50;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      red_kernel<<<prop.multiProcessorCount, 256>>>(db, floats, n_buf, 5);
      cudaEventRecord(e0);
      red_kernel<<<prop.multiProcessorCount, 256>>>(db, floats, n_buf, rounds);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fl = (double)floats * rounds * prop.multiProcessorCount;
      printf("T3 red.v4 n_buf %3d: %.3f ms  %.1f Gfloat/s  (%.1f us per 170k-float flush per SM)\n", n_buf, ms,
             fl / ms / 1e6, ms * 1e3 / rounds);
    }
  }
  printf("probe done\n");
  return 0;
}
