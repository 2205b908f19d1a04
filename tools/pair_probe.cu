// CTA-pair (cta_group::2) tcgen05 probe for sm_100a: D[256 x N] = A[256 x K] * B[N x K]^T, fp16 operands, fp32 accumulate,
// one cluster of two CTAs.  Each CTA stages its own 128 rows of A and its own HALF of the N rows of B; the leader CTA issues
// the M=256 MMAs for both; completion is committed to mbarriers of both CTAs.  Exercises exactly the mechanics a paired
// version of the convolution kernels needs (2-CTA TMEM allocation, TMA loads that signal the leader's barrier, multicast
// commits, remote arrives) and checks the result against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pair_probe.cu -o tools/pair_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../tf_raft_b200/csrc/tmap.cuh"
using namespace raft;

constexpr int kM = 256, kK = 256, kStages = 2;

struct alignas(64) PairParams {
  CUtensorMap a_map, b_map;   // A: (K, 256) box {64, 128};  B: (K, N) box {64, N/2}
  float* d;
  int n;
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) pair_kernel(const __grid_constant__ PairParams p, long long* cyc) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nh = p.n / 2;                                  // rows of B staged by each CTA
  const int a_bytes = 128 * 128, b_bytes = nh * 128, stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint32_t* holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem2_alloc(holder, 256u);
    tmem2_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const int chunks = kK / 64;
  const long long t0 = clock64();
  if (warp == 0) {
    if (elect_one()) {
      for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc % kStages;
        mbar_wait(&empty_bar[s], ((uint32_t)(kc / kStages) & 1u) ^ 1u);
        const uint32_t lead = mapa_u32(smem_u32(&full_bar[s]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * stage_bytes));   // bytes of BOTH CTAs
        uint8_t* st = smem + (size_t)s * stage_bytes;
        tma2_load_2d(st, &p.a_map, lead, kc * 64, (int)rank * 128);
        tma2_load_2d(st + a_bytes, &p.b_map, lead, kc * 64, (int)rank * nh);
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_f16(256, p.n);
      for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc % kStages;
        mbar_wait(&full_bar[s], (uint32_t)(kc / kStages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma2_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
          umma2_commit(&empty_bar[s]);
          if (kc == chunks - 1) umma2_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp & 3;                        // warps 2..5 -> lane quarters 2, 3, 0, 1
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int row = (int)rank * 128 + quarter * 32 + lane;
    for (int c0 = 0; c0 < p.n; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.d[(size_t)row * p.n + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) cyc[rank] = clock64() - t0;
  if (warp == 1) tmem2_dealloc(tmem_base, 256u);
#endif
}

int main() {
  int bad_total = 0;
  for (int n : {256, 128, 192, 96, 64, 32, 16}) {
    std::vector<__half> ha((size_t)kM * kK), hb((size_t)n * kK);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(1234 + n);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 17 - 8) / 8.f; ha[i] = __float2half(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 13 - 6) / 4.f; hb[i] = __float2half(fb[i]); }
    __half *da, *db;
    float* dd;
    long long* cyc;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, (size_t)kM * n * 4); cudaMalloc(&cyc, 16);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, (size_t)kM * n * 4);
    PairParams p;
    memset(&p, 0, sizeof(p));
    uint64_t adims[2] = {(uint64_t)kK, (uint64_t)kM}, astr[1] = {(uint64_t)kK * 2};
    uint32_t abox[2] = {64, 128};
    uint64_t bdims[2] = {(uint64_t)kK, (uint64_t)n}, bstr[1] = {(uint64_t)kK * 2};
    uint32_t bbox[2] = {64, (uint32_t)(n / 2)};
    if (make_tmap_f16(&p.a_map, da, 2, adims, astr, abox) || make_tmap_f16(&p.b_map, db, 2, bdims, bstr, bbox)) { printf("tensor map failed\n"); return 1; }
    p.d = dd; p.n = n;
    const int smem = 1024 + kStages * (128 * 128 + (n / 2) * 128) + 256;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_kernel<<<2, 192, smem>>>(p, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: kernel failed: %s\n", n, cudaGetErrorString(e)); return 1; }
    std::vector<float> hd((size_t)kM * n);
    long long hc[2];
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < kM; ++m)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < kK; ++k) s += (double)fa[(size_t)m * kK + k] * fb[(size_t)j * kK + k];
        const double err = fabs(s - hd[(size_t)m * n + j]);
        if (err > maxerr) maxerr = err;
        if (!(err <= 1e-3)) { if (bad < 4) printf("  mismatch m=%d n=%d want %.4f got %.4f\n", m, j, s, hd[(size_t)m * n + j]); ++bad; }
      }
    printf("N=%3d: %s  mismatches %d / %d  max |err| %.3g  cycles (rank0, rank1) %lld %lld\n", n, bad ? "FAIL" : "ok", bad, kM * n, maxerr, hc[0], hc[1]);
    bad_total += bad;
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(cyc);
  }
  return bad_total ? 2 : 0;
}
