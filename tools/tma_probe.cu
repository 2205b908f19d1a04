// TMA delivery-rate probe (sm_100a): persistent CTAs stream boxes of an L2-resident fp16 tensor into a shared
// memory ring and do nothing else.  Reports GB/s per SM for several box shapes / promotion modes / depths.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_probe.cu -o tools/tma_probe
#include <cstdio>
#include <cstring>
#include <vector>
#include "../tf_raft_b200/csrc/tmap.cuh"
using namespace raft;

struct alignas(64) ProbeParams {
  CUtensorMap map;
  int box_bytes, nslots, loads_per_cta, rows_total, box_rows, ncols_chunks, rank;
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
    const int nboxes_r = p.rows_total / p.box_rows;
    unsigned seed = blockIdx.x * 7919u + 13u;
    while (done < p.loads_per_cta) {
      while (issued < p.loads_per_cta && issued - done < p.nslots) {
        const int s = issued % p.nslots;
        seed = seed * 1664525u + 1013904223u;
        const int r = (int)((seed >> 8) % (unsigned)nboxes_r) * p.box_rows;
        const int c = (int)((seed >> 4) % (unsigned)p.ncols_chunks) * 64;
        mbar_arrive_expect_tx(&bar[s], (uint32_t)p.box_bytes);
        if (p.rank == 3) tma_load_3d(smem + (size_t)s * p.box_bytes, &p.map, &bar[s], c, r, 0);
        else tma_load_4d(smem + (size_t)s * p.box_bytes, &p.map, &bar[s], c, r % 64, (r / 64) * 2 % 56, 0);
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

static int make_map(CUtensorMap* m, const void* base, int rank, int cols, int rows, int box_rows, CUtensorMapL2promotion promo,
                    int th = 1) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return -1;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], es[4] = {1, 1, 1, 1};
  if (rank == 3) {
    gdim[0] = cols; gdim[1] = rows; gdim[2] = 1;
    gstr[0] = (cuuint64_t)cols * 2; gstr[1] = (cuuint64_t)cols * 2 * rows;
    box[0] = 64; box[1] = box_rows; box[2] = 1;
  } else {           // NHWC plane (C=cols, W=64, H=56, B=rows/3584)
    gdim[0] = cols; gdim[1] = 64; gdim[2] = 56; gdim[3] = rows / 3584;
    gstr[0] = (cuuint64_t)cols * 2; gstr[1] = gstr[0] * 64; gstr[2] = gstr[1] * 56;
    box[0] = 64; box[1] = box_rows / th; box[2] = th; box[3] = 1;
  }
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

int main() {
  const int rows = 14336;
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%-44s %8s %8s %10s\n", "config", "slots", "CTAs", "GB/s/SM");
  struct Cfg { const char* name; int rank, cols, box_rows, th; CUtensorMapL2promotion promo; };
  const Cfg cfgs[] = {
      {"3D rows contiguous (C=64)  box 128 rows", 3, 64, 128, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3D rows strided   (C=256) box 128 rows", 3, 256, 128, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3D rows strided   (C=256) box 256 rows", 3, 256, 256, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"3D strided C=256 box128, promo NONE", 3, 256, 128, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"3D strided C=256 box128, promo 128B", 3, 256, 128, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"4D NHWC C=256 box 64x2 (conv A tile)", 4, 256, 128, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"4D NHWC C=384 box 64x2", 4, 384, 128, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
  };
  for (const Cfg& c : cfgs) {
    __half* buf;
    const size_t bytes = (size_t)rows * c.cols * 2;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    for (int slots : {1, 2, 4, 8, 12}) {
      for (int ctas : {1, 148}) {
        ProbeParams p;
        memset(&p, 0, sizeof(p));
        if (make_map(&p.map, buf, c.rank, c.cols, rows, c.box_rows, c.promo, c.th)) { printf("map failed\n"); return 1; }
        p.box_bytes = c.box_rows * 128; p.nslots = slots; p.loads_per_cta = 2000; p.rows_total = rows; p.box_rows = c.box_rows;
        p.ncols_chunks = c.cols / 64; p.rank = c.rank;
        const int smem = slots * p.box_bytes + 1024 + slots * 8 + 64;
        if (smem > 227 * 1024) continue;
        probe_kernel<<<ctas, 128, smem>>>(p, cyc);      // warm (L2 fill)
        probe_kernel<<<ctas, 128, smem>>>(p, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<unsigned long long> h(ctas);
        cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= ctas;
        const double sec = avg / (clk_khz * 1e3);
        printf("%-44s %8d %8d %10.1f   (%.1f B/clk)\n", c.name, slots, ctas, (double)p.loads_per_cta * p.box_bytes / sec / 1e9,
               (double)p.loads_per_cta * p.box_bytes / avg);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
