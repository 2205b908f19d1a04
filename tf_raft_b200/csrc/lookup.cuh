// CorrBlock.retrieve (tf_raft/layers/corr.py:116-152) for the model's configurations (radius 4 / 3, 4 levels): the
// HBM gather of the recurrent loop.
//
// One warp per (query pixel, pyramid level).  The (2r+1)^2 floor/ceil-bilinear taps of one such item read a footprint
// of at most (2r+2) x (2r+2) texels of the query's H x W correlation plane, so the warp
//   1. computes the 2*(2r+1) per-axis tap set-ups (clamp, floor, ceil, weights -- corr.py:40-60) once, lanes 0..2S-1;
//   2. copies the footprint (<= 11 rows x 16 columns, 16-byte aligned row segments) from global to shared memory with
//      two 128-bit loads per lane -- every texel is fetched once, coalesced, instead of four scalar gathers per tap;
//   3. evaluates the taps from shared memory: lane = (x offset a, b mod 3), three taps per lane, the x set-up of the
//      lane is read once.  Each tap is the reference expression op by op with individually rounded fp32 operations
//      (__fmul_rn / __fadd_rn), so the output is bit-identical to the oracle, quirks included (integer or clamped
//      coordinate => 0, x-major tap order, fl(fl(c / 2^l) + d)).
// About 140 warp instructions per item instead of the 469 of the gather-per-tap kernel (kernels.cuh, kept for other
// radii / level counts and as the per-item escape when a footprint does not fit the window).
#pragma once
#include "kernels.cuh"

namespace raft {

constexpr int kWinPitch = 20;     // floats per window row: 16 columns + 4 (banks of the three row groups stay apart)
constexpr int kWinRows = 11;

template <int R, int L, bool kVec>
__global__ void __launch_bounds__(256, 5) corr_lookup_win_kernel(const LookupParams p) {
  constexpr int S = 2 * R + 1, NT = S * S;
  constexpr int NJ = (S + 2) / 3;                                   // taps per lane
  static_assert(S <= 10 && 3 * S <= 32, "lane = (a, b mod 3) needs 3 * (2r+1) lanes");
  __shared__ float4 ax_s[8][32];                                    // per warp: [0,16) x set-ups, [16,32) y set-ups
  __shared__ __align__(16) float win_s[8][kWinRows * kWinPitch];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float4* ax = ax_s[wib];
  float* win = win_s[wib];
  const int a = lane / 3, bq = lane - 3 * a;                        // this lane's taps: (a, bq + 3j)
  const unsigned nwork = (unsigned)p.nq * L;                        // host guarantees nq * levels < 2^31
  const unsigned stride = gridDim.x * 8u;
  unsigned wi = blockIdx.x * 8u + wib;
  float2 cnext = make_float2(0.f, 0.f);
  if (wi < nwork) cnext = __ldg(reinterpret_cast<const float2*>(p.coords) + wi / L);
  for (; wi < nwork; wi += stride) {
    const int q = (int)(wi / L), l = (int)(wi % L);
    const float2 c = cnext;
    if (wi + stride < nwork) cnext = __ldg(reinterpret_cast<const float2*>(p.coords) + (wi + stride) / L);   // next item
    const int H = p.lh[l], W = p.lw[l];
    const float* img = p.pyr[l] + (size_t)q * H * W;
    const float inv = 1.0f / (float)(1 << l);                       // exact power of two
    const float cx = __fmul_rn(c.x, inv), cy = __fmul_rn(c.y, inv); // coords / 2**i  (corr.py:141)

    // ---- 1. axis set-ups: lanes [0,S) x, [S,2S) y (the other lanes compute a harmless duplicate) ----
    const bool isy = lane >= S;
    const int i = min(isy ? lane - S : lane, S - 1);
    const int dim = isy ? H : W;
    const float g = fminf(fmaxf(__fadd_rn(isy ? cy : cx, (float)(i - R)), 0.0f), (float)(dim - 1));   // centroid + delta, clamp
    const float g0 = floorf(g), g1 = ceilf(g);
    const int i0 = (int)g0, i1 = (int)g1;
    const int bx = __shfl_sync(0xffffffffu, i0, 0), by = __shfl_sync(0xffffffffu, i0, S);
    const int ex = __shfl_sync(0xffffffffu, i1, S - 1), ey = __shfl_sync(0xffffffffu, i1, 2 * S - 1);
    const int bxa = kVec ? (bx & ~3) : bx;                          // first window column (16-byte aligned when vectorised)
    float* o = p.out ? p.out + (size_t)q * p.out_stride + l * NT : nullptr;
    __half* oh = p.out_hi ? p.out_hi + (size_t)q * p.h_stride + l * NT : nullptr;
    __half* ol = p.out_hi ? p.out_lo + (size_t)q * p.h_stride + l * NT : nullptr;

    if (ex - bxa < 16 && ey - by < kWinRows) {                      // always, by construction of the set-ups (<= 10 apart)
      __syncwarp();                                                 // previous item's readers are done with ax / win
      if (lane < 2 * S)
        ax[(isy ? 16 : 0) + i] = make_float4(__fsub_rn(g1, g), __fsub_rn(g, g0),
                                             __int_as_float(isy ? (i0 - by) * kWinPitch : i0 - bxa),
                                             __int_as_float(isy ? (i1 - by) * kWinPitch : i1 - bxa));
      // ---- 2. footprint -> shared memory ----
      if constexpr (kVec) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int e = lane + 32 * k, r = e >> 2, c4 = (e & 3) << 2;
          const int y = by + r, x = bxa + c4;
          if (r < kWinRows && y <= ey && x <= ex)
            *reinterpret_cast<float4*>(win + r * kWinPitch + c4) = __ldg(reinterpret_cast<const float4*>(img + (size_t)y * W + x));
        }
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {                               // 11 rows x 16 columns, one texel per lane and step
          const int e = lane + 32 * k, r = e >> 4, cc = e & 15;
          const int y = by + r, x = bxa + cc;
          if (r < kWinRows && y <= ey && x <= ex) win[r * kWinPitch + cc] = __ldg(img + (size_t)y * W + x);
        }
      }
      __syncwarp();
      // ---- 3. taps ----
      if (a < S) {
        const float4 sx = ax[a];                                    // (w1, w0, off0, off1)
        const int ox0 = __float_as_int(sx.z), ox1 = __float_as_int(sx.w);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int b = bq + 3 * j;
          if (b < S) {
            const float4 sy = ax[16 + b];
            const int oy0 = __float_as_int(sy.z), oy1 = __float_as_int(sy.w);
            const float x00 = win[oy0 + ox0], x01 = win[oy0 + ox1], x10 = win[oy1 + ox0], x11 = win[oy1 + ox1];
            float v = __fmul_rn(__fmul_rn(sy.x, sx.x), x00);        // corr.py:68, left to right
            v = __fadd_rn(v, __fmul_rn(__fmul_rn(sy.x, sx.y), x01));
            v = __fadd_rn(v, __fmul_rn(__fmul_rn(sy.y, sx.x), x10));
            v = __fadd_rn(v, __fmul_rn(__fmul_rn(sy.y, sx.y), x11));
            const int t = a * S + b;                                // x-major tap order (corr.py:133-143)
            if (o) o[t] = v;
            if (oh) {
              __half hh, ll;
              split_f16(v, hh, ll);
              oh[t] = hh;
              ol[t] = ll;
            }
          }
        }
      }
    } else {                                                        // escape: gather per tap (never taken in practice)
      for (int t = lane; t < NT; t += 32) {
        const int ta = t / S, tb = t - ta * S;
        const float v = sample_floor_ceil(img, H, W, __fadd_rn(cx, (float)(ta - R)), __fadd_rn(cy, (float)(tb - R)));
        if (o) o[t] = v;
        if (oh) {
          __half hh, ll;
          split_f16(v, hh, ll);
          oh[t] = hh;
          ol[t] = ll;
        }
      }
    }
    if (oh && l == L - 1) {
      const __half zero = __float2half_rn(0.f);
      for (int cpad = NT + lane; cpad < p.h_pad - l * NT; cpad += 32) {   // channels [levels * ntap, h_pad) of the operand planes
        oh[cpad] = zero;
        ol[cpad] = zero;
      }
    }
  }
}

// Launch for (radius, levels) in {(4,4), (3,4)}; returns false when the configuration has no window instantiation.
inline bool lookup_win_launch(const LookupParams& p, int levels, int radius, cudaStream_t st) {
  const size_t nwork = (size_t)p.nq * levels;
  if (levels != 4 || (radius != 4 && radius != 3) || nwork >= (1u << 31)) return false;
  bool vec = true;                                                  // 128-bit loads need 16-byte aligned rows on every level
  for (int l = 0; l < levels; ++l)
    vec = vec && (p.lw[l] % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pyr[l]) & 15) == 0);
  const int grid = grid_for(nwork * 32, 256, kNumSMs * 8);
  if (radius == 4) {
    if (vec) corr_lookup_win_kernel<4, 4, true><<<grid, 256, 0, st>>>(p);
    else corr_lookup_win_kernel<4, 4, false><<<grid, 256, 0, st>>>(p);
  } else {
    if (vec) corr_lookup_win_kernel<3, 4, true><<<grid, 256, 0, st>>>(p);
    else corr_lookup_win_kernel<3, 4, false><<<grid, 256, 0, st>>>(p);
  }
  return true;
}

}  // namespace raft
