// TMA delivery probe for the encoder-layer box shapes (sm_100a): persistent CTAs walk the tiles of a (8, 224, 256, 64)
// fp16 hi/lo activation (2 x 58.7 MB: streams from HBM like the real layer) or a small weight tensor, load one box per
// step into a shared-memory ring and do nothing else.  Reports cycles per box and B/clk per SM for each shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_probe2.cu -o tools/tma_probe2
#include <cstdio>
#include <cstring>
#include <vector>
#include "../tf_raft_b200/csrc/tmap.cuh"
using namespace raft;

struct alignas(64) ProbeParams {
  CUtensorMap map;
  int box_bytes, nslots, loads_per_cta, mode, xoff, W, H, B, tile_w;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p, unsigned long long* cyc) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.nslots * p.box_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslots; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int issued = 0, done = 0;
    const int tiles_x = p.W / p.tile_w;
    while (done < p.loads_per_cta) {
      while (issued < p.loads_per_cta && issued - done < p.nslots) {
        const int s = issued % p.nslots;
        mbar_arrive_expect_tx(&bar[s], (uint32_t)p.box_bytes);
        if (p.mode == 0) {                       // activation: tile t = blockIdx + k * grid, three rows per tile
          const int t = (int)blockIdx.x + (issued / 3) * (int)gridDim.x, ky = issued % 3;
          const int tx = t % tiles_x, ty = (t / tiles_x) % p.H, b = (t / tiles_x / p.H) % p.B;
          tma_load_5d(smem + (size_t)s * p.box_bytes, &p.map, &bar[s], 0, tx * p.tile_w + p.xoff, ty + ky - 1, b, 0);
        } else {                                 // weights: (cin, cout, tap, plane)
          tma_load_4d(smem + (size_t)s * p.box_bytes, &p.map, &bar[s], 0, 0, (issued % 3) * p.xoff, 0);
        }
        ++issued;
      }
      const int s = done % p.nslots;
      mbar_wait(&bar[s], (uint32_t)(done / p.nslots) & 1u);
      ++done;
    }
    cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
#endif
}

int main() {
  const int B = 8, H = 224, W = 256, C = 64;
  const size_t plane = (size_t)B * H * W * C;
  __half* act;
  cudaMalloc(&act, plane * 2 * 2);
  cudaMemset(act, 0, plane * 2 * 2);
  __half* wgt;
  const size_t wplane = 9 * 64 * 64;
  cudaMalloc(&wgt, wplane * 2 * 2);
  cudaMemset(wgt, 0, wplane * 2 * 2);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { const char* name; int mode, bw, bh, planes, xoff, tile_w; };
  const Cfg cfgs[] = {
      {"act {64,128,1,1,2} x0      (32 KB)", 0, 128, 1, 2, 0, 128},
      {"act {64,128,1,1,2} x0-1    (32 KB)", 0, 128, 1, 2, -1, 128},
      {"act {64,136,1,1,2} x0-1    (34 KB)", 0, 136, 1, 2, -1, 128},
      {"act {64,130,1,1,2} x0-1  (32.5 KB)", 0, 130, 1, 2, -1, 128},
      {"act {64,136,1,1,1} x0-1    (17 KB)", 0, 136, 1, 1, -1, 128},
      {"act {64,256,1,1,2} x0      (64 KB)", 0, 256, 1, 2, 0, 256},
      {"act {64,64,2,1,2}  x0      (32 KB)", 0, 64, 2, 2, 0, 64},
      {"wgt {64,64,1,2} one tap    (16 KB)", 1, 64, 1, 2, 1, 0},
      {"wgt {64,64,3,2} three taps (48 KB)", 1, 64, 3, 2, 3, 0},
  };
  printf("%-40s %6s %6s %12s %10s\n", "box", "slots", "CTAs", "cycles/box", "B/clk/SM");
  for (const Cfg& c : cfgs) {
    ProbeParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if (c.mode == 0) {
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, 2};
      uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)plane * 2};
      uint32_t box[5] = {64, (uint32_t)c.bw, (uint32_t)c.bh, 1, (uint32_t)c.planes};
      rc = make_tmap_f16(&p.map, act, 5, dims, str, box);
      p.box_bytes = c.bw * c.bh * 128 * c.planes;
    } else {
      uint64_t dims[4] = {64, 64, 9, 2};
      uint64_t str[3] = {128, 64 * 128, (uint64_t)wplane * 2};
      uint32_t box[4] = {64, 64, (uint32_t)c.bh, 2};
      rc = make_tmap_f16(&p.map, wgt, 4, dims, str, box);
      p.box_bytes = 64 * 128 * c.bh * 2;
    }
    if (rc) { printf("map failed for %s\n", c.name); continue; }
    p.mode = c.mode; p.xoff = c.xoff; p.W = W; p.H = H; p.B = B; p.tile_w = c.tile_w ? c.tile_w : 128;
    for (int slots : {2, 4}) {
      for (int ctas : {1, 148}) {
        p.nslots = slots; p.loads_per_cta = 600;
        const int smem = slots * p.box_bytes + 1024 + slots * 8 + 64;
        if (smem > 227 * 1024) continue;
        probe_kernel<<<ctas, 128, smem>>>(p, cyc);
        probe_kernel<<<ctas, 128, smem>>>(p, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<unsigned long long> h(ctas);
        cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= ctas;
        printf("%-40s %6d %6d %12.0f %10.1f\n", c.name, slots, ctas, avg / p.loads_per_cta, (double)p.loads_per_cta * p.box_bytes / avg);
      }
    }
  }
  return 0;
}
