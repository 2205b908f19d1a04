// All-pairs correlation pyramid (tf_raft/layers/corr.py:100-114, 154-162) on the tensor cores: ONE persistent
// warp-specialised tcgen05 kernel writes every level of the pyramid,
//
//   pyr[l][b, q, n] = < fmap1[b, q, :], avgpool^l(fmap2)[b, n, :] > / sqrt(C)          (pooling is linear: the 256-channel
//                                                                                       features are pooled, not the volume)
// plus one preparation kernel that pools fmap2 and splits both feature maps into the fp16 (hi, lo) operand planes.
//
// Tile = 128 queries x bn <= 128 targets, K = C in 64-channel chunks of a 3-stage ring (64 KB stages), three fp16 passes
// per chunk (hi*hi, lo*hi, hi*lo; fp32-grade).  Accumulation chains are 24 MMAs long, exactly the
// arithmetic of every other tensor-core layer: the K chunks are issued in groups of two, each group into its OWN 128-column
// TMEM buffer, and the store warps add the group results in IEEE fp32 when they read them out ((0 + g0) + g1 + ...).  Four
// TMEM buffers = two tiles in flight, so the MMAs of tile i+1 run while tile i is read out and stored.  (A single 48-MMA
// chain per tile is inside the pyramid tolerance, but its ~1e-6 relative difference moved one query of the benchmark pair
// across a discontinuity of the reference sampler at iteration 4 -- DESIGN.md section 4 -- so the chains stay at 24.)
// Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = read-out + store: each store warp owns 32 queries (its TMEM lane
// quarter) x 64 columns, transposes 32 x 32 blocks through a swizzled 4 KB shared-memory patch and writes full 128-byte
// lines of the pyramid rows.  The tile list runs over all levels (level 0 first), consecutive CTAs take consecutive query
// tiles of the same target tile, so the target features are shared through L2.  Everything in the store path is inlined
// and register-resident: the round-1 form of this epilogue lived in a non-inlined routine whose call made ptxas spill
// accumulators to local memory (which, with 227 KB of the L1/shared array configured as shared memory, is an L2 round trip
// per access).
#pragma once
#include "conv_tc.cuh"
#include "kernels.cuh"

namespace raft {

constexpr int kCorrStoreWarps = 8;
constexpr int kCorrThreads = 64 + 32 * kCorrStoreWarps;
constexpr int kCorrBn = 128;                                              // target columns per tile (one TMEM buffer)
constexpr int kCorrStageBytes = 2 * kABytes + 2 * kCorrBn * kChunkK * 2;  // 64 KB: A (hi|lo) 32 KB + B (hi|lo) up to 32 KB
constexpr int kCorrStages = 3;
constexpr int kCorrPatchBytes = kCorrStoreWarps * 4096;
constexpr int kCorrSmemBytes = 1024 /*align slack*/ + kCorrStages * kCorrStageBytes + kCorrPatchBytes + 256 /*barriers*/;
static_assert(kCorrSmemBytes <= 227 * 1024, "correlation kernel shared memory");

struct alignas(64) CorrTcParams {
  CUtensorMap a_map;                        // fmap1 (hi, lo): (C, N, 1, B, plane), box {64, 128, 1, 1, 2}
  CUtensorMap b_map[RAFT_MAX_LEVELS];       // level-l target features (hi, lo): (C, N2_l, B, plane), box {64, bn_l, 1, 2}
  float* out[RAFT_MAX_LEVELS];              // pyr[l]: (B * N, N2_l) fp32
  int n2[RAFT_MAX_LEVELS], bn[RAFT_MAX_LEVELS];
  int tile0[RAFT_MAX_LEVELS + 1];           // first tile index of each level; tile0[levels] = total
  int levels, B, N, chunks;                 // chunks = C / 64
  int mtiles_img;                           // query tiles per image = ceil(N / 128)
  float corr_mul, corr_div;                 // 1/sqrt(C) when that is an exact power of two, else 0 and the divisor is used
  long long* dbg;                           // optional clock64 timeline of CTA 0 (tools/timeline_corr.py): [4][512]
};

__global__ void __launch_bounds__(kCorrThreads, 1) corr_tc_kernel(const __grid_constant__ CorrTcParams p) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  float* patches = reinterpret_cast<float*>(smem + kCorrStages * kCorrStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kCorrStages * kCorrStageBytes + kCorrPatchBytes);
  uint64_t* empty_bar = full_bar + kCorrStages;
  uint64_t* acc_full = empty_bar + kCorrStages;      // [2] issuer -> store warps: the buffer PAIR of a tile holds two groups
  uint64_t* acc_empty = acc_full + 2;                // [2] store warps -> issuer: the pair has been read out
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = p.B * p.mtiles_img;
  const int ntiles = p.tile0[p.levels];
  const int chunks = p.chunks;
  const int npairs = (chunks + 3) >> 2;              // group pairs (4 K chunks) per tile: 1 for C <= 256

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kCorrStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kCorrStoreWarps);
    }
    fence_mbar_init();
    prefetch_tmap(&p.a_map);
    for (int l = 0; l < p.levels; ++l) prefetch_tmap(&p.b_map[l]);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // tile t -> (level, target tile, query tile): levels in order, query tile fastest
  auto decode = [&](int t, int& l, int& nt, int& mt) {
    l = 0;
    while (l + 1 < p.levels && t >= p.tile0[l + 1]) ++l;
    const int r = t - p.tile0[l];
    nt = r / mtiles;
    mt = r - nt * mtiles;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int l, nt, mt;
        decode(t, l, nt, mt);
        const int b = mt / p.mtiles_img, m0 = (mt - b * p.mtiles_img) * kTileM, n0 = nt * p.bn[l];
        const uint32_t bytes = (uint32_t)(2 * kABytes + p.bn[l] * kChunkK * 4);
        for (int kc = 0; kc < chunks; ++kc, ++it) {
          const int s = it % kCorrStages;
          mbar_wait(&empty_bar[s], ((uint32_t)(it / kCorrStages) & 1u) ^ 1u);
          if (p.dbg && blockIdx.x == 0 && it < 512) p.dbg[it] = clock64();
          uint8_t* st = stages + (size_t)s * kCorrStageBytes;
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          tma_load_5d(st, &p.a_map, &full_bar[s], kc * kChunkK, m0, 0, b, 0);
          tma_load_4d(st + 2 * kABytes, &p.b_map[l], &full_bar[s], kc * kChunkK, n0, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int it = 0, uu = 0;                                // uu: uses of the buffer pairs (tile * npairs + pair)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int l, nt, mt;
      decode(t, l, nt, mt);
      const int bn = p.bn[l];
      const uint32_t idesc = make_idesc_f16(kTileM, bn);
      const uint32_t b_lo_off = (uint32_t)(bn * kChunkK * 2);
      int done = 0;
      for (int gp = 0; gp < npairs; ++gp, ++uu) {
        const int pr = uu & 1;                         // buffer pair: TMEM columns [256 pr, 256 pr + 256)
        mbar_wait(&acc_empty[pr], ((uint32_t)(uu >> 1) & 1u) ^ 1u);          // store warps have read this pair out
        tc_fence_after();
        const int pend = min(chunks, done + 4);
        for (; done < pend; ++done, ++it) {
          const int s = it % kCorrStages;
          const int g = (done >> 1) & 1;               // group inside the pair: chunks {0,1} -> buffer 0, {2,3} -> buffer 1
          const uint32_t d_tmem = tmem_base + (uint32_t)(pr * 256 + g * 128);
          const bool first = (done & 1) == 0;
          mbar_wait(&full_bar[s], (uint32_t)(it / kCorrStages) & 1u);
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && it < 512 && lane == 0) p.dbg[512 + it] = clock64();
          if (elect_one()) {
            const uint32_t sa = smem_u32(stages + (size_t)s * kCorrStageBytes);
            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + kABytes);
            const uint64_t b_hi = make_desc_sw128(sa + 2 * kABytes), b_lo = make_desc_sw128(sa + 2 * kABytes + b_lo_off);
            umma_chunk3<false>(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, first != 0);
            umma_commit(&empty_bar[s]);
            if (done == pend - 1) umma_commit(&acc_full[pr]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== read-out + store (warps 2..9) =====================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;                // columns [part * 64, part * 64 + 64) of the tile
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(part * 64);
    float* patch = patches + (warp - 2) * 1024;
    const int rsub = lane >> 3, q4 = lane & 7;
    int uu = 0, tt = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tt) {
      int l, nt, mt;
      decode(t, l, nt, mt);
      const int bn = p.bn[l], n2 = p.n2[l];
      const int b = mt / p.mtiles_img, m0 = (mt - b * p.mtiles_img) * kTileM, n0 = nt * bn;
      // 32-column blocks of this warp in this tile that hold columns of the level (warp-uniform)
      const int nblk = max(0, min(2, (min(bn, n2 - n0) - part * 64 + 31) >> 5));

      float racc[2][32];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int j = 0; j < 32; ++j) racc[ci][j] = 0.0f;

#pragma unroll 1
      for (int gp = 0; gp < npairs; ++gp, ++uu) {
        const int pr = uu & 1;
        const int ng = min(2, ((chunks - 4 * gp) + 1) >> 1);               // groups in this pair (1 or 2)
        mbar_wait(&acc_full[pr], (uint32_t)(uu >> 1) & 1u);
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && gp == npairs - 1 && tt < 256 && warp == 2 && lane == 0) p.dbg[1024 + tt] = clock64();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < ng) {
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
              if (ci < nblk) {
                uint32_t r[32];
                tmem_ld_32x32(trow + (uint32_t)(pr * 256 + g * 128 + ci * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) racc[ci][j] += __uint_as_float(r[j]);   // IEEE fp32 sum of the 24-MMA groups
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[pr]);                        // the issuer may reuse the pair
        if (p.dbg && blockIdx.x == 0 && gp == npairs - 1 && tt < 256 && warp == 2 && lane == 0) p.dbg[1536 + tt] = clock64();
      }

      // ---- store (the issuer is already two tiles ahead at most) ----
      const int row_base = m0 + quarter * 32;                              // first query of this warp's 32 rows
      float* dst0 = p.out[l] + ((size_t)b * p.N + row_base + rsub) * n2 + n0 + part * 64 + 4 * q4;
      const int col_end = min(n2, n0 + bn);                                // bn need not be a multiple of 32: the last block
                                                                           // of a tile stops at the tile's own last column
      const bool vec = (n2 & 3) == 0 && (bn & 3) == 0;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        if (ci < nblk) {
          __syncwarp();                                                    // previous block's readers are done with the patch
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            float4 a4 = make_float4(racc[ci][4 * qq], racc[ci][4 * qq + 1], racc[ci][4 * qq + 2], racc[ci][4 * qq + 3]);
            if (p.corr_mul != 0.0f) {
              a4.x *= p.corr_mul; a4.y *= p.corr_mul; a4.z *= p.corr_mul; a4.w *= p.corr_mul;
            } else {
              a4.x = __fdiv_rn(a4.x, p.corr_div); a4.y = __fdiv_rn(a4.y, p.corr_div);
              a4.z = __fdiv_rn(a4.z, p.corr_div); a4.w = __fdiv_rn(a4.w, p.corr_div);
            }
            *reinterpret_cast<float4*>(patch + lane * 32 + ((qq ^ (lane & 7)) << 2)) = a4;
          }
          __syncwarp();
          const int col = n0 + part * 64 + ci * 32 + 4 * q4;
          float* dst = dst0 + ci * 32;
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 4) {                              // four 16-byte loads in flight, then four stores
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 4 * (i0 + i) + rsub;
              v[i] = *reinterpret_cast<const float4*>(patch + rr * 32 + ((q4 ^ (rr & 7)) << 2));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = row_base + 4 * (i0 + i) + rsub;
              if (row < p.N && col < col_end) {
                float* d = dst + (size_t)(4 * (i0 + i)) * n2;
                if (vec) {
                  *reinterpret_cast<float4*>(d) = v[i];
                } else {
                  d[0] = v[i].x;
                  if (col + 1 < col_end) d[1] = v[i].y;
                  if (col + 2 < col_end) d[2] = v[i].z;
                  if (col + 3 < col_end) d[3] = v[i].w;
                }
              }
            }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && tt < 255) p.dbg[1536 + 256 + tt] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
#endif
}

// ------------------------------------------------------------------------------------------------
// Preparation: fmap1 -> (hi, lo) planes; fmap2 -> pooled levels 1..levels-1 (2x2 mean, VALID, successively, the same
// fp32 arithmetic as avgpool2x2_kernel) and the (hi, lo) planes of every level.  One launch:
//   blocks [0, npatch): one 8 x 8 patch of fmap2 pixels (a complete pooling tree down to level 3), thread = channel pair;
//   blocks [npatch, ...): elementwise split of fmap1, 4 channels per thread.
// levels <= 4 here (deeper pyramids take the generic kernels).
// ------------------------------------------------------------------------------------------------
struct CorrPrepParams {
  const float* f1; const float* f2;
  __half *f1_hi, *f1_lo;
  __half *f2_hi[4], *f2_lo[4];
  int B, h, w, C, levels;
  int patches_x, patches_y, npatch;
};

__device__ __forceinline__ void store_split2(__half* hi, __half* lo, size_t o, float x, float y) {
  uint32_t h2, l2;
  split_f16x2(x, y, h2, l2);
  *reinterpret_cast<uint32_t*>(hi + o) = h2;
  *reinterpret_cast<uint32_t*>(lo + o) = l2;
}
__device__ __forceinline__ float pool4(float a, float b, float c, float d) {
  return __fmul_rn(__fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d)), 0.25f);
}

__global__ void __launch_bounds__(128) corr_prep_kernel(const CorrPrepParams p) {
  if ((int)blockIdx.x >= p.npatch) {                 // ---- fmap1: elementwise split ----
    const size_t n4 = (size_t)p.B * p.h * p.w * p.C / 4;
    const size_t nb = gridDim.x - p.npatch;
    for (size_t i = (size_t)(blockIdx.x - p.npatch) * blockDim.x + threadIdx.x; i < n4; i += nb * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.f1) + i);
      uint32_t h01, l01, h23, l23;
      split_f16x2(v.x, v.y, h01, l01);
      split_f16x2(v.z, v.w, h23, l23);
      reinterpret_cast<uint2*>(p.f1_hi)[i] = make_uint2(h01, h23);
      reinterpret_cast<uint2*>(p.f1_lo)[i] = make_uint2(l01, l23);
    }
    return;
  }
  int t = blockIdx.x;
  const int px = t % p.patches_x; t /= p.patches_x;
  const int py = t % p.patches_y;
  const int b = t / p.patches_y;
  const int C = p.C;
  int hl[4], wl[4];
  hl[0] = p.h; wl[0] = p.w;
#pragma unroll
  for (int l = 1; l < 4; ++l) { hl[l] = hl[l - 1] / 2; wl[l] = wl[l - 1] / 2; }
  for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
    float2 l1p[4], l2p[2];                            // level-1 row / level-2 row waiting for their partner
#pragma unroll 1
    for (int rp = 0; rp < 4; ++rp) {                  // row pairs of the patch (rolled: 16 loads in flight at a time)
      const int y0 = py * 8 + rp * 2;
      float2 v[2][8];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const int yy = y0 + dy, xx = px * 8 + x;
          v[dy][x] = (yy < hl[0] && xx < wl[0])
                         ? __ldg(reinterpret_cast<const float2*>(p.f2 + (((size_t)b * hl[0] + yy) * wl[0] + xx) * C + c))
                         : make_float2(0.f, 0.f);
        }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const int yy = y0 + dy, xx = px * 8 + x;
          if (yy < hl[0] && xx < wl[0])
            store_split2(p.f2_hi[0], p.f2_lo[0], (((size_t)b * hl[0] + yy) * wl[0] + xx) * C + c, v[dy][x].x, v[dy][x].y);
        }
      float2 l1[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        l1[x].x = pool4(v[0][2 * x].x, v[0][2 * x + 1].x, v[1][2 * x].x, v[1][2 * x + 1].x);
        l1[x].y = pool4(v[0][2 * x].y, v[0][2 * x + 1].y, v[1][2 * x].y, v[1][2 * x + 1].y);
        const int y1 = py * 4 + rp, x1 = px * 4 + x;
        if (p.levels > 1 && y1 < hl[1] && x1 < wl[1])
          store_split2(p.f2_hi[1], p.f2_lo[1], (((size_t)b * hl[1] + y1) * wl[1] + x1) * C + c, l1[x].x, l1[x].y);
      }
      if (rp & 1) {
        float2 l2[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          l2[x].x = pool4(l1p[2 * x].x, l1p[2 * x + 1].x, l1[2 * x].x, l1[2 * x + 1].x);
          l2[x].y = pool4(l1p[2 * x].y, l1p[2 * x + 1].y, l1[2 * x].y, l1[2 * x + 1].y);
          const int y2 = py * 2 + (rp >> 1), x2 = px * 2 + x;
          if (p.levels > 2 && y2 < hl[2] && x2 < wl[2])
            store_split2(p.f2_hi[2], p.f2_lo[2], (((size_t)b * hl[2] + y2) * wl[2] + x2) * C + c, l2[x].x, l2[x].y);
        }
        if (rp == 3) {
          if (p.levels > 3 && py < hl[3] && px < wl[3])
            store_split2(p.f2_hi[3], p.f2_lo[3], (((size_t)b * hl[3] + py) * wl[3] + px) * C + c,
                         pool4(l2p[0].x, l2p[1].x, l2[0].x, l2[1].x), pool4(l2p[0].y, l2p[1].y, l2[0].y, l2[1].y));
        } else {
          l2p[0] = l2[0]; l2p[1] = l2[1];
        }
      } else {
#pragma unroll
        for (int x = 0; x < 4; ++x) l1p[x] = l1[x];
      }
    }
  }
}

}  // namespace raft
