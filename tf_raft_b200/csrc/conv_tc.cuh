// tcgen05 implicit-GEMM kernel: one kernel serves the all-pairs correlation (a 1x1 "conv" whose
// weights are the other image's features) and every stride-1 convolution of the update blocks.
//
//   D[128 px, bn cout] = sum over (tap, 64-channel chunk) of  A_tap[128 px, 64] * W_tap[bn, 64]^T
//
// * A operand: a TH x TW pixel patch of an NHWC fp16 activation plane, fetched by ONE 4-D TMA box per
//   (tap, chunk) at coordinates shifted by the tap offset; out-of-image rows/cols are zero-filled by
//   the TMA unit, which *is* Keras 'same' padding -- no im2col buffer, no halo logic.
// * B operand: packed weights [tap][cout][cin] fp16, one 3-D TMA box per (tap, chunk).
// * Both operands land in shared memory K-major with the 128-byte swizzle and are consumed in place
//   by tcgen05.mma (kind::f16, M=128, N=bn, K=16); the fp32 accumulator lives in TMEM.
// * fp32-grade arithmetic from fp16 tensor cores: each operand is a (hi, lo) fp16 pair and every
//   K step issues hi*hi, lo*hi, hi*lo into the same accumulator (DESIGN.md "Precision").
// * Tensor-core fp32 accumulation truncates (measured: ~0.6 ulp of bias per K=16 step, all towards
//   zero), so a 1920-deep GRU contraction issued as one 360-step chain is ~60x less accurate than an
//   FFMA chain.  The K loop is therefore cut into groups of `group_chunks` 64-channel chunks; each
//   group accumulates into one of two TMEM buffers (ping-pong) and is then promoted -- added in IEEE
//   fp32 -- into per-thread register accumulators while the next group runs on the tensor core.
// * Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one
//   elected lane), warps 2..9 = promotion + epilogue (TMEM -> registers, then fused bias / activation /
//   GRU gating / hi-lo re-split -> global).  mbarrier rings: smem full/empty (producer <-> issuer),
//   TMEM full/empty (issuer <-> promotion warps).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "tmap.cuh"

namespace raft {

enum TcEpilogue : int {
  EPI_LINEAR = 0,  // v = act(acc*inv_scale + bias) * out_scale  -> optional fp32 and/or fp16 hi/lo planes
  EPI_GRU_ZR = 1,  // cols [0,hid): z = sigmoid(v) -> z plane; cols [hid,2hid): r = sigmoid(v), r*h -> hi/lo planes
  EPI_GRU_Q = 2,   // q = tanh(v); h = (1-z)*h + z*q -> h fp32 (in place) + hi/lo planes
};
enum TcAct : int { ACT_NONE = 0, ACT_RELU = 1 };

constexpr int kEpiWarpsConv = 16;         // promotion/epilogue warps of the convolution instantiation (4 per TMEM lane quarter)
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // fp16 elements per 128-byte swizzled row
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KiB per A plane per stage
constexpr int kEpiPatchBytes = 16 * 2048;         // EPI_GRU_Q: transposition patches of the 16 epilogue warps
constexpr int kSmemBudget = 227 * 1024 - 2048;
constexpr int kARow3Pixels = 136;                 // kRow3: activation box width (1 + 128 + 1 pixels, padded to a multiple of 8)
constexpr int kARow3Bytes = 2 * kARow3Pixels * kChunkK * 2;   // hi + lo planes of one 136-pixel row chunk: 34 KB

struct alignas(64) TcConvParams {
  CUtensorMap a_map[2];           // (hi, lo) plane pair of up to two channel-concatenated sources (K segments)
  CUtensorMap b_map;              // (hi, lo) plane pair of the packed weights
  int nseg, seg_chunks[2], seg_c0[2];
  int kh, kw, ph, pw;             // taps and 'same' padding (pad before)
  int stride;                     // 1 or 2: input pixel = output pixel * stride + tap - pad (TMA elementStrides)
  int B, H, W, TH, TW, tiles_x, tiles_y;
  int bn, n_total;                // N per CTA (multiple of 16, <= 256); total valid output columns
  int n_tiles_n;                  // column tiles (tile id = n_tile * pixel_tiles + pixel_tile)
  int nstages, stage_bytes, tmem_cols;
  int group_chunks;               // K chunks per promotion group (accumulation chain = 12 * group_chunks MMAs)
  int mode, act;
  const float* bias;              // [n_total padded to bn multiple]; may be null
  const float* inv_scale;         // device scalar: 1 / (2^k weight scale); may be null (=1)
  float out_scale;                // applied after the activation (0.25 for the mask head)
  float* out_f32; int f32_stride, f32_c0;
  __half* out_hi; __half* out_lo; int h_stride, h_c0;
  const float* post_scale;        // EPI_LINEAR: optional per-column affine after the bias (folded BatchNorm):
  const float* post_shift;        //   v = v * post_scale[col] + post_shift[col]
  const float* residual; int res_stride, res_c0;   // EPI_LINEAR: optional skip input: v = relu(act(v) + residual)
  const float* concat_src; int concat_n;   // EPI_LINEAR: fp32 (px, concat_n) appended at columns [n_total, n_total+concat_n)
  float* z; int hid;                       // GRU: z plane (px, hid) fp32
  float* h;                                // GRU: hidden state (px, hid) fp32, updated in place by EPI_GRU_Q
  long long* dbg;                          // optional timeline of CTA 0 (tools/timeline.py): [4][512] clock64 stamps
  // Programmatic dependent launch (default; RAFT_B200_PDL=0 disables): the launch carries the programmatic-serialization
  // attribute, so this grid's CTAs may be scheduled -- and run their prologue -- while the previous kernel in the stream
  // drains; every thread then executes griddepcontrol.wait before touching global memory.
  int pdl;
  // kRow3 instantiation (3x3, stride 1, 1 x 128 pixel tiles, cout <= 96: the wide encoder layers).  The three taps of a
  // kernel row read the SAME image row shifted by one pixel, so a stage holds ONE activation box of 136 pixels (x0-1 ..
  // x0+134; 136 * 128 B = 17 KB per plane keeps the lo plane on a 1024-byte boundary) and ONE weight box with the row's
  // three taps; tap dx issues its MMAs on the activation rows [dx, dx + 128) through a descriptor whose start is shifted by
  // dx * 128 bytes.  3 + 3 boxes per 64-channel chunk instead of 9 + 9: these layers are bound by TMA box delivery.
  int row3;
  // EPI_LINEAR with n_total == 2 (flow_head.conv2) inside the iteration loop: coords1 += delta_flow and
  // flow = coords1 - coords0 (model.py:102, :97) are applied by the thread that holds the pixel's two output columns.
  float* adv_coords;                       // (px, 2) coords1, updated in place; null = no fused advance
  float* adv_flow;                         // (px, 2) coords1 - pixel grid
  // update_mega_kernel<true> only: the tile is one half of a CTA pair's M = 256 MMA; a stage holds this CTA's 128 activation
  // rows and HALF of the bn weight rows (b_map's box is bn / 2 rows).
  int pair;
};

#if defined(__CUDA_ARCH__)
// 16-byte global accesses and the fast activation forms used by every epilogue.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 32-byte global store (STG.256, sm_100): the thread-per-row epilogues write a full 32-byte sector per lane and instruction
// instead of half of one -- half as many store instructions and LSU wavefronts for the same bytes.  `p` must be 32-byte aligned.
__device__ __forceinline__ void st8u(void* p, const uint32_t (&r)[8]) {
#if defined(RAFT_EPI_EXP) && (RAFT_EPI_EXP & 1)
  if (r[0] != 0x7fc12345u) return;          // experiment: no epilogue stores (tools/epi_exp.sh)
#endif
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 32-byte loads: L2-coherent (tensors another CTA of the same grid may have written) and read-only forms.
__device__ __forceinline__ void ldcg8(const float* p, float (&r)[8]) {
#if defined(RAFT_EPI_EXP) && (RAFT_EPI_EXP & 2)
  if (p != nullptr) {                        // experiment: no epilogue operand loads
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = 0.5f;
    return;
  }
#endif
  asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void ldnc8(const float* p, float (&r)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st8f(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  const uint32_t r[8] = {__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d),
                         __float_as_uint(e), __float_as_uint(f), __float_as_uint(g), __float_as_uint(h)};
  st8u(p, r);
}
// L2-coherent 16-byte load (ld.global.cg): z and h may have been written by another CTA of the SAME grid (update_mega_kernel),
// so they must not be served from this SM's L1 nor through the non-coherent path.
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

#if defined(RAFT_EPI_EXP) && (RAFT_EPI_EXP & 4)
__device__ __forceinline__ float fast_sigmoid(float x) { return x; }      // experiment: no MUFU
#else
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
#endif
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// ------------------------------------------------------------------------------------------------
// Register-resident epilogue of one 32-column chunk of one pixel row (thread == row).
//
// Measured: with 227 KB of the unified L1/shared array configured as shared memory, per-thread LOCAL memory does not
// stay in L1, so the earlier out-of-line routine (accumulators staged through a local buffer, rolled loops) paid an
// L2 round trip per access: ~32-37 k cycles per tile regardless of the tile width.  Here nothing leaves registers:
// the routine is inlined and specialised on the epilogue mode at compile time, loops are fully unrolled so the
// independent 16-byte global loads (bias, h, z, residual) are all in flight together, and the activation math uses
// the fast exp / reciprocal units (|error| ~1e-7, far inside the parity budget).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void tc_epilogue_regs(const TcConvParams& p, float (&v)[32], size_t pix, int col, int ncol,
                                                 float inv_scale) {
  // bias + folded BatchNorm affine (arrays are zero-padded past the last column)
  if (p.bias) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 bq = ldg4(p.bias + col + 4 * q);
      v[4 * q] = v[4 * q] * inv_scale + bq.x; v[4 * q + 1] = v[4 * q + 1] * inv_scale + bq.y;
      v[4 * q + 2] = v[4 * q + 2] * inv_scale + bq.z; v[4 * q + 3] = v[4 * q + 3] * inv_scale + bq.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= inv_scale;
  }
  if (MODE == EPI_LINEAR && p.post_scale) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 sc = ldg4(p.post_scale + col + 4 * q), sh = ldg4(p.post_shift + col + 4 * q);
      v[4 * q] = v[4 * q] * sc.x + sh.x; v[4 * q + 1] = v[4 * q + 1] * sc.y + sh.y;
      v[4 * q + 2] = v[4 * q + 2] * sc.z + sh.z; v[4 * q + 3] = v[4 * q + 3] * sc.w + sh.w;
    }
  }

  __half* dhi = nullptr;
  __half* dlo = nullptr;
  if (MODE == EPI_LINEAR) {
    const bool full = col + 32 <= p.n_total;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (p.act == ACT_RELU) v[j] = fmaxf(v[j], 0.0f);
      v[j] *= p.out_scale;
    }
    if (full) {
      if (p.residual) {
        const float* rp = p.residual + pix * (size_t)p.res_stride + p.res_c0 + col;
        if (((p.res_stride | p.res_c0) & 7) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float r8[8];
            ldnc8(rp + 8 * q, r8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 * q + e] = fmaxf(v[8 * q + e] + r8[e], 0.f);
          }
        } else if (((p.res_stride | p.res_c0) & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 r4 = ldg4(rp + 4 * q);
            v[4 * q] = fmaxf(v[4 * q] + r4.x, 0.f); v[4 * q + 1] = fmaxf(v[4 * q + 1] + r4.y, 0.f);
            v[4 * q + 2] = fmaxf(v[4 * q + 2] + r4.z, 0.f); v[4 * q + 3] = fmaxf(v[4 * q + 3] + r4.w, 0.f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + __ldg(rp + j), 0.f);
        }
      }
    } else {                                           // ragged tail: concat columns / zeros / residual per column
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cj = col + j - p.n_total;
        if (cj >= 0) v[j] = (p.concat_src && cj < p.concat_n) ? __ldg(p.concat_src + pix * p.concat_n + cj) : 0.0f;
        else if (p.residual) v[j] = fmaxf(v[j] + __ldg(p.residual + pix * (size_t)p.res_stride + p.res_c0 + col + j), 0.0f);
      }
    }
    if (p.out_f32) {
      const int nvalid = min(ncol, p.n_total - col);
      float* dst = p.out_f32 + pix * (size_t)p.f32_stride + p.f32_c0 + col;
      if (nvalid == 32 && ((p.f32_stride | p.f32_c0) & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st8f(dst + 8 * q, v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
      } else if (nvalid == 32 && ((p.f32_stride | p.f32_c0) & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) st4(dst + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = v[j];
      }
    }
    if (p.out_hi && ncol == 32) {
      const size_t o = pix * (size_t)p.h_stride + p.h_c0 + col;
      dhi = p.out_hi + o;
      dlo = p.out_lo + o;
    }
    if (p.adv_coords && col == 0) {                    // model.py:102: coords1 += delta; flow = coords1 - coords0 (pixel grid)
      float2 c1 = __ldcg(reinterpret_cast<const float2*>(p.adv_coords) + pix);
      c1.x = __fadd_rn(c1.x, v[0]);
      c1.y = __fadd_rn(c1.y, v[1]);
      reinterpret_cast<float2*>(p.adv_coords)[pix] = c1;
      const float gx = (float)(pix % p.W), gy = (float)((pix / p.W) % p.H);
      reinterpret_cast<float2*>(p.adv_flow)[pix] = make_float2(__fsub_rn(c1.x, gx), __fsub_rn(c1.y, gy));
    }
  } else if (MODE == EPI_GRU_ZR) {
    if (col < p.hid) {                                 // z gate -> fp32 plane
      float* dst = p.z + pix * (size_t)p.hid + col;
      if ((p.hid & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st8f(dst + 8 * q, fast_sigmoid(v[8 * q]), fast_sigmoid(v[8 * q + 1]), fast_sigmoid(v[8 * q + 2]), fast_sigmoid(v[8 * q + 3]),
               fast_sigmoid(v[8 * q + 4]), fast_sigmoid(v[8 * q + 5]), fast_sigmoid(v[8 * q + 6]), fast_sigmoid(v[8 * q + 7]));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st4(dst + 4 * q, make_float4(fast_sigmoid(v[4 * q]), fast_sigmoid(v[4 * q + 1]), fast_sigmoid(v[4 * q + 2]),
                                       fast_sigmoid(v[4 * q + 3])));
      }
    } else {                                           // r gate -> r*h, re-split for the q convolution
      const int hc = col - p.hid;
      const float* hp = p.h + pix * (size_t)p.hid + hc;
      if ((p.hid & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float h8[8];
          ldcg8(hp + 8 * q, h8);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 * q + e] = fast_sigmoid(v[8 * q + e]) * h8[e];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 hv = ldcg4(hp + 4 * q);
          v[4 * q] = fast_sigmoid(v[4 * q]) * hv.x; v[4 * q + 1] = fast_sigmoid(v[4 * q + 1]) * hv.y;
          v[4 * q + 2] = fast_sigmoid(v[4 * q + 2]) * hv.z; v[4 * q + 3] = fast_sigmoid(v[4 * q + 3]) * hv.w;
        }
      }
      const size_t o = pix * (size_t)p.h_stride + p.h_c0 + hc;
      dhi = p.out_hi + o;
      dlo = p.out_lo + o;
    }
  } else if (MODE == EPI_GRU_Q) {                      // h = (1-z)*h + z*tanh(v), in place
    float* hrow = p.h + pix * (size_t)p.hid + col;
    const float* zp = p.z + pix * (size_t)p.hid + col;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 zv = ldcg4(zp + 4 * q), hv = ldcg4(hrow + 4 * q);
      v[4 * q] = (1.0f - zv.x) * hv.x + zv.x * fast_tanh(v[4 * q]);
      v[4 * q + 1] = (1.0f - zv.y) * hv.y + zv.y * fast_tanh(v[4 * q + 1]);
      v[4 * q + 2] = (1.0f - zv.z) * hv.z + zv.z * fast_tanh(v[4 * q + 2]);
      v[4 * q + 3] = (1.0f - zv.w) * hv.w + zv.w * fast_tanh(v[4 * q + 3]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) st4(hrow + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    const size_t o = pix * (size_t)p.h_stride + p.h_c0 + col;
    dhi = p.out_hi + o;
    dlo = p.out_lo + o;
  }

  if (dhi) {                                           // fp16 hi/lo re-split, 16 channels (32 bytes) per store
    // (operand planes: channel strides are multiples of 64 and slices start at multiples of 32 channels -> 32-byte aligned;
    //  a slice that starts elsewhere takes the 16-byte form)
    const bool wide = ((reinterpret_cast<uintptr_t>(dhi) | reinterpret_cast<uintptr_t>(dlo)) & 31) == 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_f16x2(v[16 * q + 2 * e], v[16 * q + 2 * e + 1], ph[e], pl[e]);
      if (wide) {
        st8u(dhi + 16 * q, ph);
        st8u(dlo + 16 * q, pl);
      } else {
        reinterpret_cast<uint4*>(dhi)[2 * q] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        reinterpret_cast<uint4*>(dhi)[2 * q + 1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        reinterpret_cast<uint4*>(dlo)[2 * q] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        reinterpret_cast<uint4*>(dlo)[2 * q + 1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Coalesced GRU-q epilogue of a 32-row x 16-column accumulator block, AFTER transposition through shared memory:
// lane l holds rows (l>>2) + 8k (k = 0..3) and columns col .. col+3 of the block, so every global access of a warp
// touches 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes.  Measured on the thread-per-row form: the LSU
// retires about one distinct 128-byte line per cycle, which made the epilogue of a 256-column tile cost 14-22 k cycles.
// pixr[k] is the flat pixel index of row k (-1 = outside the image).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split4(__half* hi, __half* lo, size_t o, bool aligned, const float (&v)[4]) {
  uint32_t h01, l01, h23, l23;
  split_f16x2(v[0], v[1], h01, l01);
  split_f16x2(v[2], v[3], h23, l23);
  if (aligned) {
    *reinterpret_cast<uint2*>(hi + o) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(lo + o) = make_uint2(l01, l23);
  } else {
    const uint32_t hh[2] = {h01, h23}, ll[2] = {l01, l23};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[o + e] = __ushort_as_half((unsigned short)(hh[e >> 1] >> (16 * (e & 1))));
      lo[o + e] = __ushort_as_half((unsigned short)(ll[e >> 1] >> (16 * (e & 1))));
    }
  }
}

__device__ __forceinline__ void tc_epilogue_q_t(const TcConvParams& p, float (&v)[4][4], const int (&pixr)[4], int col,
                                              float inv_scale) {
  if (p.bias) {
    const float4 bq = ldg4(p.bias + col);            // bias / affine arrays are zero-padded past the last column
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k][0] = v[k][0] * inv_scale + bq.x; v[k][1] = v[k][1] * inv_scale + bq.y;
      v[k][2] = v[k][2] * inv_scale + bq.z; v[k][3] = v[k][3] * inv_scale + bq.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[k][e] *= inv_scale;
  }

  // h = (1-z)*h + z*tanh(v), in place
  {
    const bool h_vec = ((p.h_stride | p.h_c0) & 3) == 0;
#pragma unroll
    for (int k0 = 0; k0 < 4; k0 += 2) {                // two rows at a time: loads in flight together, registers bounded
      float4 zv[2], hv[2];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const bool ok = pixr[k0 + kk] >= 0;
        zv[kk] = ok ? ldcg4(p.z + (size_t)pixr[k0 + kk] * p.hid + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        hv[kk] = ok ? ldcg4(p.h + (size_t)pixr[k0 + kk] * p.hid + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = k0 + kk;
        if (pixr[k] < 0) continue;
        v[k][0] = (1.0f - zv[kk].x) * hv[kk].x + zv[kk].x * fast_tanh(v[k][0]);
        v[k][1] = (1.0f - zv[kk].y) * hv[kk].y + zv[kk].y * fast_tanh(v[k][1]);
        v[k][2] = (1.0f - zv[kk].z) * hv[kk].z + zv[kk].z * fast_tanh(v[k][2]);
        v[k][3] = (1.0f - zv[kk].w) * hv[kk].w + zv[kk].w * fast_tanh(v[k][3]);
        st4(p.h + (size_t)pixr[k] * p.hid + col, make_float4(v[k][0], v[k][1], v[k][2], v[k][3]));
        store_split4(p.out_hi, p.out_lo, (size_t)pixr[k] * p.h_stride + p.h_c0 + col, h_vec, v[k]);
      }
    }
  }
}

#endif

// kRowEpi selects the epilogue form at compile time: thread-per-row registers (EPI_LINEAR, EPI_GRU_ZR) or transposed through
// shared-memory patches (EPI_GRU_Q), so that neither costs the other registers or code.
template <bool kRowEpi, bool kRow3 = false>
__global__ void __launch_bounds__(64 + 32 * kEpiWarpsConv, 1) conv_tc_kernel(const __grid_constant__ TcConvParams p) {
#if defined(__CUDA_ARCH__)
  // Persistent: CTA c processes output tiles c, c + gridDim.x, ...  A tile is (pixel tile, column tile).  All
  // three roles walk the same tile sequence; the smem ring and the two TMEM buffers carry straight across
  // tile boundaries, so the loads and MMAs of tile i+1 overlap the epilogue of tile i.
  constexpr int kEpiWarps = kEpiWarpsConv;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nst = p.nstages;
  const int b_bytes = p.bn * kChunkK * 2;
  uint8_t* stages = smem;                                     // ring of nst stages
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stages + (size_t)nst * p.stage_bytes);
  uint64_t* empty_bar = full_bar + nst;
  uint64_t* acc_full = empty_bar + nst;      // [2] issuer -> promotion warps
  uint64_t* acc_empty = acc_full + 2;        // [2] promotion warps -> issuer
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* patches = reinterpret_cast<float*>(stages + (size_t)nst * p.stage_bytes + 256);   // transposition patches (GRU q)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int mtiles = p.B * p.tiles_y * p.tiles_x;
  const int ntiles = mtiles * p.n_tiles_n;
  const int ntaps = p.kh * p.kw;
  const int chunks_per_tap = p.seg_chunks[0] + (p.nseg > 1 ? p.seg_chunks[1] : 0);
  const int total = (kRow3 ? p.kh : ntaps) * chunks_per_tap;   // stages per tile (kRow3: one per kernel row and chunk)
  const int gsz = p.group_chunks;
  const int ngroups = (total + gsz - 1) / gsz;            // promotion groups per tile
  const int nchunks32 = (p.bn + 31) >> 5;                 // 32-column accumulator chunks
  constexpr int kParts = kEpiWarps / 4;                   // warps per TMEM lane quarter: each takes a slice of the columns
  constexpr int kMaxCh = 8 / kParts;                      // accumulator chunks per thread (2)
  const int chunks_per_part = (nchunks32 + kParts - 1) / kParts;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4 * ((nchunks32 + chunks_per_part - 1) / chunks_per_part));   // one arrival per participating warp
    }
    fence_mbar_init();
    prefetch_tmap(&p.a_map[0]);
    prefetch_tmap(&p.b_map);
    if (p.nseg > 1) prefetch_tmap(&p.a_map[1]);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (p.pdl && warp != 0) {
    // barriers, TMEM and tensor-map prefetch above touch no global data; from here on the previous grid's results are
    // read (and its inputs overwritten), so wait for it, then let the next grid in the stream start its own prologue.
    // (Warp 0 = the producer thread waits further down, after it has started fetching the weights of the first stages.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      int npre = 0;                // stages whose weight box was issued before the dependency wait (PDL only)
      if (p.pdl) {
        // Convolution weights do not depend on the previous grid: fetch them for the first stages of this CTA's first
        // tile while that grid is still draining, then wait, then fetch the activations.
        if (!kRow3 && (int)blockIdx.x < ntiles) {
          const int n0 = ((int)blockIdx.x / mtiles) * p.bn;
          npre = min(nst, total);
#pragma unroll 1
          for (int i = 0; i < npre; ++i) {
            mbar_arrive_expect_tx(&full_bar[i], (uint32_t)p.stage_bytes);
            tma_load_4d(stages + (size_t)i * p.stage_bytes + 2 * kABytes, &p.b_map, &full_bar[i], (i % chunks_per_tap) * kChunkK,
                        n0, i / chunks_per_tap, 0);
          }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      }
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int nt = t / mtiles;
        int mt = t - nt * mtiles;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int b = mt / p.tiles_y;
        const int x0 = tx * p.TW * p.stride, y0 = ty * p.TH * p.stride, n0 = nt * p.bn;
        if constexpr (kRow3) {
          for (int ky = 0; ky < p.kh; ++ky) {
            for (int ch = 0; ch < p.seg_chunks[0]; ++ch, ++it) {
              const int s = it % nst;
              mbar_wait(&empty_bar[s], ((uint32_t)(it / nst) & 1u) ^ 1u);
              if (p.dbg && blockIdx.x == 0 && it < 512) p.dbg[it] = clock64();              // slot free
              uint8_t* st = stages + (size_t)s * p.stage_bytes;
              mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
              tma_load_5d(st, &p.a_map[0], &full_bar[s], p.seg_c0[0] + ch * kChunkK, x0 - p.pw, y0 + ky - p.ph, b, 0);
              tma_load_4d(st + kARow3Bytes, &p.b_map, &full_bar[s], ch * kChunkK, n0, ky * p.kw, 0);
            }
          }
        } else {
          for (int tap = 0; tap < ntaps; ++tap) {
            const int dy = tap / p.kw - p.ph, dx = tap % p.kw - p.pw;
            int kc = 0;
            for (int seg = 0; seg < p.nseg; ++seg) {
              for (int ch = 0; ch < p.seg_chunks[seg]; ++ch, ++kc, ++it) {
                const int s = it % nst;
                const uint32_t phase = (uint32_t)(it / nst) & 1u;
                mbar_wait(&empty_bar[s], phase ^ 1u);
                if (p.dbg && blockIdx.x == 0 && it < 512) p.dbg[it] = clock64();              // slot free
                uint8_t* st = stages + (size_t)s * p.stage_bytes;
                const int c = p.seg_c0[seg] + ch * kChunkK;
#if defined(RAFT_TC_EXP) && (RAFT_TC_EXP & 6)        // mainloop experiments (tools/tc_exp.sh; run with RAFT_B200_PDL=0)
                if (RAFT_TC_EXP & 2) {               // activations only
                  mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * kABytes));
                  tma_load_5d(st, &p.a_map[seg], &full_bar[s], c, x0 + dx, y0 + dy, b, 0);
                } else {                             // weights only
                  mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(p.stage_bytes - 2 * kABytes));
                  tma_load_4d(st + 2 * kABytes, &p.b_map, &full_bar[s], kc * kChunkK, n0, tap, 0);
                }
#else
                if (it >= npre) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
                // two boxes per stage: [A_hi | A_lo] and [B_hi | B_lo] (TMA cost is per box, not per byte)
                tma_load_5d(st, &p.a_map[seg], &full_bar[s], c, x0 + dx, y0 + dy, b, 0);
                if (it >= npre) tma_load_4d(st + 2 * kABytes, &p.b_map, &full_bar[s], kc * kChunkK, n0, tap, 0);
#endif
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_f16(kTileM, p.bn);
    int it = 0, gg = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int done = 0;
      for (int g = 0; g < ngroups; ++g, ++gg) {
        const int buf = gg & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(gg >> 1) & 1u) ^ 1u);   // promotion warps drained this buffer
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && gg < 256 && lane == 0) p.dbg[1024 + 256 + gg] = clock64();   // issuer owns the buffer
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.bn);
        const int gend = min(total, done + gsz);
        for (int first = 1; done < gend; ++done, ++it, first = 0) {
          const int s = it % nst;
          const uint32_t phase = (uint32_t)(it / nst) & 1u;
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && it < 512 && lane == 0) p.dbg[512 + it] = clock64();   // data landed
          if (elect_one()) {
            const uint32_t sa = smem_u32(stages + (size_t)s * p.stage_bytes);
            if constexpr (kRow3) {
              // stage = [A hi: 136 rows | A lo: 136 rows | W hi: tap 0, 1, 2 | W lo: tap 0, 1, 2]
#pragma unroll 1
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t a_hi = make_desc_sw128(sa + dx * 128);     // shifted start, base-offset field 0 (common.cuh)
                const uint64_t a_lo = make_desc_sw128(sa + kARow3Bytes / 2 + dx * 128);
                const uint64_t b_hi = make_desc_sw128(sa + kARow3Bytes + dx * b_bytes);
                const uint64_t b_lo = make_desc_sw128(sa + kARow3Bytes + (3 + dx) * b_bytes);
                umma_chunk3<false>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, first != 0 && dx == 0);
              }
            } else {
              const uint64_t a_hi = make_desc_sw128(sa);
              const uint64_t a_lo = make_desc_sw128(sa + kABytes);
              const uint64_t b_hi = make_desc_sw128(sa + 2 * kABytes);
              const uint64_t b_lo = make_desc_sw128(sa + 2 * kABytes + b_bytes);
#if defined(RAFT_TC_EXP) && (RAFT_TC_EXP & 1)        // experiment: no MMAs, only the commits
              if (sa != 0xffffffffu) goto tc_exp_skip_mma;
#endif
              umma_chunk3<false>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, first != 0);   // (+32 bytes per K=16 slice == +2 in 16-byte units)
#if defined(RAFT_TC_EXP) && (RAFT_TC_EXP & 1)
            tc_exp_skip_mma:;
#endif
            }
            umma_commit(&empty_bar[s]);                      // frees the smem slot once these MMAs retire
            if (done == gend - 1) umma_commit(&acc_full[buf]);   // group complete -> promotion warps
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== promotion + epilogue (warps 2..17) =====================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int part_id = (warp - 2) >> 2;             // column slice of this warp within its lane quarter
    const int chunk0 = part_id * chunks_per_part;
    const int my_chunks = max(0, min(chunks_per_part, nchunks32 - chunk0));
    const int m = quarter * 32 + lane;               // tile row == TMEM lane
    const int xl = m % p.TW, yl = m / p.TW;
    const float inv_scale = p.inv_scale ? __ldg(p.inv_scale) : 1.0f;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);

    if (my_chunks > 0) {
      int gg = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        float racc[kMaxCh][32];
#pragma unroll
        for (int ci = 0; ci < kMaxCh; ++ci)
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[ci][j] = 0.0f;

        for (int g = 0; g < ngroups; ++g, ++gg) {
          const int buf = gg & 1;
          mbar_wait(&acc_full[buf], (uint32_t)(gg >> 1) & 1u);
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && gg < 256 && warp == 2 && lane == 0) p.dbg[1024 + gg] = clock64();   // group retired
#pragma unroll
          for (int ci = 0; ci < kMaxCh; ++ci) {
            if (ci < my_chunks) {
              const int c0 = (chunk0 + ci) * 32;
              uint32_t r[32];
              if (p.bn - c0 >= 32) {
                tmem_ld_32x32(trow + (uint32_t)(buf * p.bn + c0), r);
              } else {
                uint32_t r16[16];
                tmem_ld_32x16(trow + (uint32_t)(buf * p.bn + c0), r16);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = r16[j];
#pragma unroll
                for (int j = 16; j < 32; ++j) r[j] = 0u;
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) racc[ci][j] += __uint_as_float(r[j]);   // IEEE fp32 promotion
            }
          }
          tc_fence_before();
          __syncwarp();
          if (p.dbg && blockIdx.x == 0 && gg < 256 && warp == 2 && lane == 0) p.dbg[1536 + gg] = clock64();   // group drained
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }

        // ---- epilogue of this tile (the issuer is already accumulating the next one) ----
        const int nt = t / mtiles;
        int mt = t - nt * mtiles;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int b = mt / p.tiles_y;
        const int x = tx * p.TW + xl, y = ty * p.TH + yl;
        if constexpr (kRowEpi) {    // thread-per-row register epilogue (EPI_LINEAR, EPI_GRU_ZR)
          if (x < p.W && y < p.H) {
            const size_t pix = ((size_t)b * p.H + y) * p.W + x;
#pragma unroll
            for (int ci = 0; ci < kMaxCh; ++ci) {
              if (ci < my_chunks) {
                const int c0 = (chunk0 + ci) * 32;
                const int ncol = min(32, p.bn - c0);
                if (p.mode == EPI_LINEAR) tc_epilogue_regs<EPI_LINEAR>(p, racc[ci], pix, nt * p.bn + c0, ncol, inv_scale);
                else tc_epilogue_regs<EPI_GRU_ZR>(p, racc[ci], pix, nt * p.bn + c0, ncol, inv_scale);
              }
            }
          }
        } else {                           // EPI_GRU_Q: coalesced epilogue through a 32 x 16 transposition patch (measured:
                                           // 13 k cycles vs 23 k thread-per-row; the other modes measured slower transposed)
          float4* patch4 = reinterpret_cast<float4*>(patches + (warp - 2) * 512);
          const int pix_own = (x < p.W && y < p.H) ? (int)(((size_t)b * p.H + y) * p.W + x) : -1;
          int pixr[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) pixr[k] = __shfl_sync(0xffffffffu, pix_own, (lane >> 2) + 8 * k);
          const int c4 = lane & 3;
          const int wsw = (lane >> 1) & 3, rsw = (lane >> 3) & 3;   // XOR swizzles: conflict-free 16-byte writes and reads
#pragma unroll
          for (int ci = 0; ci < kMaxCh; ++ci) {
            if (ci < my_chunks) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int c0 = (chunk0 + ci) * 32 + hh * 16;
                if (c0 < p.bn) {
                  __syncwarp();
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    patch4[lane * 4 + (q ^ wsw)] = make_float4(racc[ci][hh * 16 + 4 * q], racc[ci][hh * 16 + 4 * q + 1],
                                                               racc[ci][hh * 16 + 4 * q + 2], racc[ci][hh * 16 + 4 * q + 3]);
                  __syncwarp();
                  float v[4][4];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float4 t4 = patch4[((lane >> 2) + 8 * k) * 4 + (c4 ^ rsw)];
                    v[k][0] = t4.x; v[k][1] = t4.y; v[k][2] = t4.z; v[k][3] = t4.w;
                  }
                  const int col = nt * p.bn + c0 + 4 * c4;
                  tc_epilogue_q_t(p, v, pixr, col, inv_scale);
                }
              }
            }
          }
        }
        if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) {
          p.dbg[2047] = clock64();                                           // epilogue of the (last) tile done
          const int tl = (t - (int)blockIdx.x) / (int)gridDim.x;
          if (tl < 255) p.dbg[1536 + 256 + tl] = clock64();                   // ... of every tile
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
#endif
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
inline bool tc_uses_patch(int mode) { return mode == EPI_GRU_Q; }

inline void tc_pick_tile(int W, int H, int* tw, int* th) {
  // TW*TH = 128 with TW a power of two; minimise padded area, prefer wide tiles on ties.
  long best = -1;
  for (int t = 128; t >= 8; t >>= 1) {
    const int hh = 128 / t;
    const long area = (long)round_up(W, t) * round_up(H, hh);
    if (best < 0 || area < best) {
      best = area;
      *tw = t;
      *th = hh;
    }
  }
}

// Fills the derived launch fields (tile grid, stages, TMEM columns) of `p`; returns bytes of
// dynamic shared memory.  Caller has set bn, B, H, W, TH, TW.
inline int tc_finalize(TcConvParams& p) {
  p.tiles_x = ceil_div(p.W, p.TW);
  p.tiles_y = ceil_div(p.H, p.TH);
  p.stage_bytes = 2 * kABytes + 2 * (p.pair ? p.bn / 2 : p.bn) * kChunkK * 2;
  if (p.row3) p.stage_bytes = kARow3Bytes + 3 * 2 * p.bn * kChunkK * 2;
  const int patch = tc_uses_patch(p.mode) ? kEpiPatchBytes : 0;   // 16 x 2 KB patches (GRU q)
  int nst = (kSmemBudget - patch) / p.stage_bytes;
  if (nst > 8) nst = 8;
  p.nstages = nst;
  int cols = 32;
  while (cols < 2 * p.bn) cols <<= 1;           // two accumulator buffers (ping-pong promotion)
  p.tmem_cols = cols;
  if (p.group_chunks <= 0) p.group_chunks = 2;
  return nst * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + patch;
}

// RAFT_B200_PDL=0 disables programmatic dependent launch of the per-layer kernel (A/B timing).
inline bool tc_pdl_enabled() {
  static const int pdl = [] { const char* e = getenv("RAFT_B200_PDL"); return e ? atoi(e) : 1; }();
  return pdl != 0;
}

inline int tc_launch(TcConvParams& p, int n_tiles_n, cudaStream_t stream) {
  if (p.bn % 16 != 0 || p.bn < 16 || p.bn > 256 || p.TW * p.TH != kTileM) return RAFT_ERR_BAD_SHAPE;
  if (p.mode != EPI_LINEAR && p.mode != EPI_GRU_ZR && p.mode != EPI_GRU_Q) return RAFT_ERR_UNSUPPORTED;
  if (p.stride < 1) p.stride = 1;
  const int smem = tc_finalize(p);
  if (p.nstages < 2) return RAFT_ERR_UNSUPPORTED;
  // the attribute is per device: one bit per device ordinal (benign race: the calls are idempotent)
  int dev = 0;
  RAFT_CUDA_TRY(cudaGetDevice(&dev));
  const unsigned long long dev_bit = 1ull << (dev & 63);
  static unsigned long long attr_set_mask = 0;
  if (!(attr_set_mask & dev_bit)) {
    RAFT_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RAFT_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RAFT_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set_mask |= dev_bit;
  }
  const int mtiles = p.B * p.tiles_y * p.tiles_x;
  p.n_tiles_n = n_tiles_n;
  const long ntiles = (long)mtiles * n_tiles_n;
  const unsigned grid = (unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs);   // one persistent CTA per SM
  p.pdl = tc_pdl_enabled() ? 1 : 0;
  const int threads = 64 + 32 * kEpiWarpsConv;
  if (p.row3 && (p.mode != EPI_LINEAR || p.kh != 3 || p.kw != 3 || p.stride != 1 || p.TH != 1 || p.TW != kTileM || p.nseg != 1 ||
                 p.bn % 8 != 0 || n_tiles_n != 1))
    return RAFT_ERR_UNSUPPORTED;
  void (*kern)(TcConvParams) = p.mode == EPI_GRU_Q ? conv_tc_kernel<false> : (p.row3 ? conv_tc_kernel<true, true> : conv_tc_kernel<true>);
  if (!p.pdl) {
    kern<<<grid, threads, smem, stream>>>(p);
  } else {                                             // programmatic-serialization attribute: see the kernel prologue
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RAFT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  }
  return raft_launch_status();
}

}  // namespace raft
