// What does one TMA box cost, and what does the cost depend on?  (sm_100a; follow-up of tma_probe2: every box of the
// convolution kernels costs 566-840 cycles whatever its size.)  An L2-resident activation tensor (4 x 56 x 64 pixels x 384
// channels, fp16, two planes) is read through tensor maps of rank 2..5 with boxes of 8-64 KB, by one or by TWO issuing
// threads (different warps, own barriers), with 2-8 boxes in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_probe3.cu -o tools/tma_probe3
#include <cstdio>
#include <cstring>
#include <vector>
#include "../tf_raft_b200/csrc/tmap.cuh"
using namespace raft;

struct alignas(64) P3 {
  CUtensorMap map;
  int rank, box_bytes, nslots, loads, issuers, W, H, B, C, bw, bh;
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
#endif

__global__ void __launch_bounds__(128, 1) probe3(const __grid_constant__ P3 p, unsigned long long* cyc) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int total_slots = p.nslots * p.issuers;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)total_slots * p.box_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < total_slots; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int who = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && who < p.issuers) {
    uint8_t* base = smem + (size_t)who * p.nslots * p.box_bytes;
    uint64_t* mybar = bar + who * p.nslots;
    const int loads = p.loads / p.issuers;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    const int tiles_y = p.H / p.bh, tiles = p.B * tiles_y;
    while (done < loads) {
      while (issued < loads && issued - done < p.nslots) {
        const int s = issued % p.nslots;
        const int i = issued * p.issuers + who;
        const int t = ((int)blockIdx.x + (i / 6) * (int)gridDim.x) % tiles, ch = (i % 6) * 64;   // six 64-channel chunks per tile
        const int b = t / tiles_y, y = (t % tiles_y) * p.bh;
        mbar_arrive_expect_tx(&mybar[s], (uint32_t)p.box_bytes);
        void* dst = base + (size_t)s * p.box_bytes;
        if (p.rank == 2) tma_load_2d(dst, &p.map, &mybar[s], ch, (b * p.H + y) * p.W);
        else if (p.rank == 3) tma_load_3d(dst, &p.map, &mybar[s], ch, 0, b * p.H + y);
        else if (p.rank == 4) tma_load_4d(dst, &p.map, &mybar[s], ch, 0, y, b);
        else tma_load_5d(dst, &p.map, &mybar[s], ch, 0, y, b, 0);
        ++issued;
      }
      const int s = done % p.nslots;
      mbar_wait(&mybar[s], (uint32_t)(done / p.nslots) & 1u);
      ++done;
    }
    cyc[blockIdx.x * 4 + who] = (unsigned long long)(clock64() - t0);
  }
#endif
}

int main() {
  const int B = 4, H = 56, W = 64, C = 384;
  const size_t plane = (size_t)B * H * W * C;
  __half* act;
  cudaMalloc(&act, plane * 2 * 2);
  cudaMemset(act, 0, plane * 2 * 2);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 32);
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { const char* name; int rank, rows_h, planes; };   // box = 64 channels x 64 px x rows_h rows (x planes)
  const Cfg cfgs[] = {
      {"rank 2 (C, B*H*W)        box {64,128}      16 KB", 2, 2, 1},
      {"rank 2 (C, B*H*W)        box {64,256}      32 KB", 2, 4, 1},
      {"rank 3 (C, W, B*H)       box {64,64,2}     16 KB", 3, 2, 1},
      {"rank 3 (C, W, B*H)       box {64,64,4}     32 KB", 3, 4, 1},
      {"rank 4 (C, W, H, B)      box {64,64,2,1}   16 KB", 4, 2, 1},
      {"rank 5 (C, W, H, B, pl)  box {64,64,2,1,1} 16 KB", 5, 2, 1},
      {"rank 5 (C, W, H, B, pl)  box {64,64,2,1,2} 32 KB", 5, 2, 2},
      {"rank 5 (C, W, H, B, pl)  box {64,64,4,1,2} 64 KB", 5, 4, 2},
  };
  printf("%-52s %7s %7s %6s %12s %10s\n", "map / box", "issuers", "slots", "CTAs", "cycles/box", "B/clk/SM");
  for (const Cfg& c : cfgs) {
    P3 p;
    memset(&p, 0, sizeof(p));
    int rc = 0;
    if (c.rank == 2) {
      uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * H * W}, str[1] = {(uint64_t)C * 2};
      uint32_t box[2] = {64, (uint32_t)(64 * c.rows_h)};
      rc = make_tmap_f16(&p.map, act, 2, dims, str, box);
    } else if (c.rank == 3) {
      uint64_t dims[3] = {(uint64_t)C, (uint64_t)W, (uint64_t)B * H}, str[2] = {(uint64_t)C * 2, (uint64_t)W * C * 2};
      uint32_t box[3] = {64, 64, (uint32_t)c.rows_h};
      rc = make_tmap_f16(&p.map, act, 3, dims, str, box);
    } else if (c.rank == 4) {
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B}, str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
      uint32_t box[4] = {64, 64, (uint32_t)c.rows_h, 1};
      rc = make_tmap_f16(&p.map, act, 4, dims, str, box);
    } else {
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, 2};
      uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)plane * 2};
      uint32_t box[5] = {64, 64, (uint32_t)c.rows_h, 1, (uint32_t)c.planes};
      rc = make_tmap_f16(&p.map, act, 5, dims, str, box);
    }
    if (rc) { printf("map failed for %s\n", c.name); continue; }
    p.rank = c.rank; p.W = W; p.H = H; p.B = B; p.C = C; p.bh = c.rows_h; p.box_bytes = 64 * c.rows_h * 128 * c.planes;
    for (int issuers : {1, 2, 3, 4}) {
      for (int slots : {2, 4}) {
        for (int ctas : {1, 148}) {
          p.nslots = slots; p.issuers = issuers; p.loads = 480;   // (divisible by 1..4)
          const int smem = slots * issuers * p.box_bytes + 1024 + 256;
          if (smem > 227 * 1024) continue;
          probe3<<<ctas, 128, smem>>>(p, cyc);
          probe3<<<ctas, 128, smem>>>(p, cyc);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          std::vector<unsigned long long> h(ctas * 4);
          cudaMemcpy(h.data(), cyc, ctas * 32, cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < ctas; ++i) {
            unsigned long long m = 0;
            for (int w = 0; w < issuers; ++w) m = h[4 * i + w] > m ? h[4 * i + w] : m;
            avg += (double)m;
          }
          avg /= ctas;
          printf("%-52s %7d %7d %6d %12.0f %10.1f\n", c.name, issuers, slots, ctas, avg / p.loads, (double)p.loads * p.box_bytes / avg);
        }
      }
    }
  }
  return 0;
}
