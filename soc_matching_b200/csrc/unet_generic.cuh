// unet_generic.cuh -- shape-generic FullyConnectedUNet forward / backward, one WARP per point.
// Serves any hdims / d <= SOCM_MAX_DIM (the tiled kernels in unet_tile.cuh are specialised
// for the default hdims [256,128,64]).  Semantics: models.py:202-242.
#pragma once
#include "common.cuh"

namespace socm {
namespace generic {

// Per-warp activation workspace (floats), carved from dynamic shared memory.
struct FwdBuf {
  float *xin, *r1, *r2, *r3, *o2, *o1, *o0;  // activations
  float *y2, *y1, *y0;                        // pre-ReLU values of the up layers (masks for backward)
};

__host__ __device__ inline int fwd_floats(int d, int h0, int h1, int h2) {
  return (d + 1) + h0 + h1 + h2 + h1 + h0 + d + h1 + h0 + d;
}
__host__ __device__ inline int bwd_floats(int d, int h0, int h1, int h2) {
  // d_o1, d_r1 (h0 each), d_o2, d_r2 (h1 each), d_r3 (h2), d_y0/d_o0 (d each)
  return 2 * h0 + 2 * h1 + h2 + 2 * d;
}

__device__ __forceinline__ FwdBuf carve_fwd(float* base, int d, int h0, int h1, int h2) {
  FwdBuf b;
  float* p = base;
  b.xin = p; p += d + 1;
  b.r1 = p; p += h0;
  b.r2 = p; p += h1;
  b.r3 = p; p += h2;
  b.o2 = p; p += h1;
  b.o1 = p; p += h0;
  b.o0 = p; p += d;
  b.y2 = p; p += h1;
  b.y1 = p; p += h0;
  b.y0 = p; p += d;
  return b;
}

// out[n] = bias[n] + sum_k W[n][k] in[k]; each lane owns rows n = lane, lane+32, ...
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ bias, int nout,
                                      int nin, const float* in, float* out, int lane) {
  for (int n = lane; n < nout; n += 32) {
    float acc = __ldg(bias + n);
    const float* row = W + (size_t)n * nin;
    for (int k = 0; k < nin; ++k) acc = fmaf(__ldg(row + k), in[k], acc);
    out[n] = acc;
  }
}

__device__ __forceinline__ void relu_inplace(float* v, int n, int lane) {
  for (int i = lane; i < n; i += 32) v[i] = fmaxf(v[i], 0.f);
}

// Forward for one point; b.xin must be filled ([t, x]).  Result in b.o0.
__device__ __forceinline__ void forward(const socm_unet& net, const FwdBuf& b, int lane) {
  const int d = net.d, h0 = net.h0, h1 = net.h1, h2 = net.h2;
  dense(net.w[0], net.b[0], h0, d + 1, b.xin, b.r1, lane);
  __syncwarp();
  relu_inplace(b.r1, h0, lane);
  __syncwarp();
  dense(net.w[1], net.b[1], h1, h0, b.r1, b.r2, lane);
  __syncwarp();
  relu_inplace(b.r2, h1, lane);
  __syncwarp();
  dense(net.w[2], net.b[2], h2, h1, b.r2, b.r3, lane);
  __syncwarp();
  relu_inplace(b.r3, h2, lane);
  __syncwarp();
  // o2 = relu(up_2 r3) + res_2 r2
  dense(net.w[6], net.b[6], h1, h2, b.r3, b.y2, lane);
  dense(net.w[5], net.b[5], h1, h1, b.r2, b.o2, lane);
  __syncwarp();
  for (int i = lane; i < h1; i += 32) b.o2[i] += fmaxf(b.y2[i], 0.f);
  __syncwarp();
  // o1 = relu(up_1 o2) + res_1 r1
  dense(net.w[7], net.b[7], h0, h1, b.o2, b.y1, lane);
  dense(net.w[4], net.b[4], h0, h0, b.r1, b.o1, lane);
  __syncwarp();
  for (int i = lane; i < h0; i += 32) b.o1[i] += fmaxf(b.y1[i], 0.f);
  __syncwarp();
  // o0 = relu(up_0 o1) + res_0 xin
  dense(net.w[8], net.b[8], d, h0, b.o1, b.y0, lane);
  dense(net.w[3], net.b[3], d, d + 1, b.xin, b.o0, lane);
  __syncwarp();
  for (int i = lane; i < d; i += 32) b.o0[i] += fmaxf(b.y0[i], 0.f);
  __syncwarp();
}

// din[k] (+)= sum_n W[n][k] dout[n]; lanes own k (coalesced rows of W)
__device__ __forceinline__ void dense_t(const float* __restrict__ W, int nout, int nin, const float* dout,
                                        float* din, bool accumulate, int lane) {
  for (int k = lane; k < nin; k += 32) {
    float acc = accumulate ? din[k] : 0.f;
    for (int n = 0; n < nout; ++n) acc = fmaf(__ldg(W + (size_t)n * nin + k), dout[n], acc);
    din[k] = acc;
  }
}

// gW[n][k] += dout[n] in[k], gb[n] += dout[n]   (global atomics; generic path only)
__device__ __forceinline__ void wgrad(float* gW, float* gb, int nout, int nin, const float* dout,
                                      const float* in, int lane) {
  for (int n = 0; n < nout; ++n) {
    const float dn = dout[n];
    if (dn == 0.f) continue;
    for (int k = lane; k < nin; k += 32) atomicAdd(gW + (size_t)n * nin + k, dn * in[k]);
    if (lane == 0) atomicAdd(gb + n, dn);
  }
}

struct GradPtrs {
  float* w[9];
  float* b[9];
};

__host__ __device__ inline void layer_dims(int d, int h0, int h1, int h2, int nout[9], int nin[9]) {
  const int o[9] = {h0, h1, h2, d, h0, h1, h1, h0, d};
  const int i[9] = {d + 1, h0, h1, d + 1, h0, h1, h2, h1, h0};
  for (int l = 0; l < 9; ++l) {
    nout[l] = o[l];
    nin[l] = i[l];
  }
}

__device__ __forceinline__ GradPtrs grad_ptrs(float* flat, int d, int h0, int h1, int h2) {
  int nout[9], nin[9];
  layer_dims(d, h0, h1, h2, nout, nin);
  GradPtrs g;
  float* p = flat;
  for (int l = 0; l < 9; ++l) {
    g.w[l] = p;
    p += (size_t)nout[l] * nin[l];
    g.b[l] = p;
    p += nout[l];
  }
  return g;
}

// Backward for one point given d_o0 (in s[0..d) of `bw`): accumulates parameter gradients.
// bw layout: d_o0[d] d_y0[d] d_o1[h0] d_r1[h0] d_o2[h1] d_r2[h1] d_r3[h2]
__device__ __forceinline__ void backward(const socm_unet& net, const FwdBuf& b, float* bw, const GradPtrs& g,
                                         int lane) {
  const int d = net.d, h0 = net.h0, h1 = net.h1, h2 = net.h2;
  float* d_o0 = bw;
  float* d_y0 = d_o0 + d;
  float* d_o1 = d_y0 + d;
  float* d_r1 = d_o1 + h0;
  float* d_o2 = d_r1 + h0;
  float* d_r2 = d_o2 + h1;
  float* d_r3 = d_r2 + h1;
  for (int i = lane; i < d; i += 32) d_y0[i] = b.y0[i] > 0.f ? d_o0[i] : 0.f;
  __syncwarp();
  wgrad(g.w[8], g.b[8], d, h0, d_y0, b.o1, lane);        // up_0
  wgrad(g.w[3], g.b[3], d, d + 1, d_o0, b.xin, lane);    // res_0
  dense_t(net.w[8], d, h0, d_y0, d_o1, false, lane);
  __syncwarp();
  // o1 = relu(y1) + res_1 r1
  wgrad(g.w[4], g.b[4], h0, h0, d_o1, b.r1, lane);       // res_1
  dense_t(net.w[4], h0, h0, d_o1, d_r1, false, lane);
  __syncwarp();
  for (int i = lane; i < h0; i += 32) d_o1[i] = b.y1[i] > 0.f ? d_o1[i] : 0.f;  // now d_y1
  __syncwarp();
  wgrad(g.w[7], g.b[7], h0, h1, d_o1, b.o2, lane);       // up_1
  dense_t(net.w[7], h0, h1, d_o1, d_o2, false, lane);
  __syncwarp();
  // o2 = relu(y2) + res_2 r2
  wgrad(g.w[5], g.b[5], h1, h1, d_o2, b.r2, lane);       // res_2
  dense_t(net.w[5], h1, h1, d_o2, d_r2, false, lane);
  __syncwarp();
  for (int i = lane; i < h1; i += 32) d_o2[i] = b.y2[i] > 0.f ? d_o2[i] : 0.f;  // now d_y2
  __syncwarp();
  wgrad(g.w[6], g.b[6], h1, h2, d_o2, b.r3, lane);       // up_2
  dense_t(net.w[6], h1, h2, d_o2, d_r3, false, lane);
  __syncwarp();
  for (int i = lane; i < h2; i += 32) d_r3[i] = b.r3[i] > 0.f ? d_r3[i] : 0.f;  // d_z3
  __syncwarp();
  wgrad(g.w[2], g.b[2], h2, h1, d_r3, b.r2, lane);       // down_2
  dense_t(net.w[2], h2, h1, d_r3, d_r2, true, lane);
  __syncwarp();
  for (int i = lane; i < h1; i += 32) d_r2[i] = b.r2[i] > 0.f ? d_r2[i] : 0.f;  // d_z2
  __syncwarp();
  wgrad(g.w[1], g.b[1], h1, h0, d_r2, b.r1, lane);       // down_1
  dense_t(net.w[1], h1, h0, d_r2, d_r1, true, lane);
  __syncwarp();
  for (int i = lane; i < h0; i += 32) d_r1[i] = b.r1[i] > 0.f ? d_r1[i] : 0.f;  // d_z1
  __syncwarp();
  wgrad(g.w[0], g.b[0], h0, d + 1, d_r1, b.xin, lane);   // down_0
  __syncwarp();
}

}  // namespace generic
}  // namespace socm
