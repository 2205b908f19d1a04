// Training-step kernels (tf_raft/model.py:126-144): backward of the pyramid lookup, global gradient norm, AdamW.
#pragma once
#include "kernels.cuh"

namespace raft {

// ------------------------------------------------------------------------------------------------
// Backward of CorrBlock.retrieve (corr.py:116-152 with bilinear_sampler corr.py:28-69).
//   out[q, l, a, b] = sum over the 4 floor/ceil corners of  wy * wx * P_l[q, iy, ix],
//   gx = clip(cx / 2^l + (a - r), 0, W-1),  wx1 = ceil(gx) - gx (corner x0), wx0 = gx - floor(gx) (corner x1), same in y.
// TensorFlow's gradients: floor / ceil / gather indices contribute nothing; clip_by_value passes the gradient where the
// value lies inside [lo, hi]; so with go = d loss / d out[q, l, a, b]
//   d P_l[q, iy, ix] += wy * wx * go                                  (4 scatter-adds)
//   d gx = go * ( wy1 * (P01 - P00) + wy0 * (P11 - P10) ),  d cx += d gx / 2^l  when 0 <= cx/2^l + (a-r) <= W-1
//   d gy = go * ( wx1 * (P10 - P00) + wx0 * (P11 - P01) ),  d cy likewise.
// One warp per (query, level); lanes stride over the taps; coordinate gradients are reduced in the warp and added with
// one atomic per (query, level); pyramid gradients are scatter-added with atomics (floating-point order not fixed).
// ------------------------------------------------------------------------------------------------
struct LookupBwdParams {
  const float* pyr[RAFT_MAX_LEVELS];
  float* gpyr[RAFT_MAX_LEVELS];
  int lh[RAFT_MAX_LEVELS], lw[RAFT_MAX_LEVELS];
  const float* coords;
  const float* gout; int gout_stride;
  float* gcoords;
  int nq, levels, radius;
};

__global__ void __launch_bounds__(256) corr_lookup_bwd_kernel(const LookupBwdParams p) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int side = 2 * p.radius + 1, ntap = side * side;
  const size_t nwork = (size_t)p.nq * p.levels;
  for (size_t wi = (size_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wi < nwork;
       wi += (size_t)gridDim.x * warps_per_block) {
    const int q = (int)(wi / p.levels), l = (int)(wi % p.levels);
    const int H = p.lh[l], W = p.lw[l];
    const float* img = p.pyr[l] + (size_t)q * H * W;
    float* gimg = p.gpyr[l] + (size_t)q * H * W;
    const float inv = 1.0f / (float)(1 << l);
    const float cx = __ldg(p.coords + 2 * (size_t)q) * inv, cy = __ldg(p.coords + 2 * (size_t)q + 1) * inv;
    float gcx = 0.0f, gcy = 0.0f;
    for (int t = lane; t < ntap; t += 32) {
      const int a = t / side, b = t - a * side;
      const float go = __ldg(p.gout + (size_t)q * p.gout_stride + l * ntap + t);
      const float ux = cx + (float)(a - p.radius), uy = cy + (float)(b - p.radius);      // before the clip
      const float gx = fminf(fmaxf(ux, 0.0f), (float)(W - 1)), gy = fminf(fmaxf(uy, 0.0f), (float)(H - 1));
      const float x0 = floorf(gx), x1 = ceilf(gx), y0 = floorf(gy), y1 = ceilf(gy);
      const float wx1 = x1 - gx, wx0 = gx - x0, wy1 = y1 - gy, wy0 = gy - y0;
      const int o00 = (int)y0 * W + (int)x0, o01 = (int)y0 * W + (int)x1, o10 = (int)y1 * W + (int)x0, o11 = (int)y1 * W + (int)x1;
      const float p00 = __ldg(img + o00), p01 = __ldg(img + o01), p10 = __ldg(img + o10), p11 = __ldg(img + o11);
      if (go != 0.0f) {
        const float c00 = wy1 * wx1, c01 = wy1 * wx0, c10 = wy0 * wx1, c11 = wy0 * wx0;
        if (c00 != 0.0f) atomicAdd(gimg + o00, c00 * go);
        if (c01 != 0.0f) atomicAdd(gimg + o01, c01 * go);
        if (c10 != 0.0f) atomicAdd(gimg + o10, c10 * go);
        if (c11 != 0.0f) atomicAdd(gimg + o11, c11 * go);
        const float dgx = go * (wy1 * (p01 - p00) + wy0 * (p11 - p10));
        const float dgy = go * (wx1 * (p10 - p00) + wx0 * (p11 - p01));
        if (ux >= 0.0f && ux <= (float)(W - 1)) gcx += dgx;
        if (uy >= 0.0f && uy <= (float)(H - 1)) gcy += dgy;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      gcx += __shfl_xor_sync(0xffffffffu, gcx, off);
      gcy += __shfl_xor_sync(0xffffffffu, gcy, off);
    }
    if (lane == 0) {
      atomicAdd(p.gcoords + 2 * (size_t)q, gcx * inv);
      atomicAdd(p.gcoords + 2 * (size_t)q + 1, gcy * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tf.linalg.global_norm / tf.clip_by_global_norm (model.py:135) and tfa AdamW (train_chairs.py:87-90) on a flat buffer.
// sumsq: deterministic two-stage reduction (fixed grid, fixed tree) -> out[0] = sum g^2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += g[i] * g[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.0f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[0] = t;
  }
}

// g <- g * clip / max(||g||, clip)   (clip_norm > 0; tf.clip_by_global_norm), then
// var -= wd * var;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  var -= lr_t * m / (sqrt(v) + eps)
// with lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) computed by the caller (Keras Adam, non-amsgrad).
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             size_t n, const float* __restrict__ sumsq, float clip_norm, float lr_t, float b1, float b2,
                             float eps, float wd) {
  float scale = 1.0f;
  if (clip_norm > 0.0f) {
    const float norm = sqrtf(*sumsq);
    scale = clip_norm / fmaxf(norm, clip_norm);
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    float w = p[i];
    w -= wd * w;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = w - lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace raft
