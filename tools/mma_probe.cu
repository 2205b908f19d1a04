// tcgen05.mma issue-rate probe (sm_100a): one CTA, operands resident in shared memory (zeros), R back-to-back
// M=128 x N x K=16 fp16 MMAs from ONE thread, or split between TWO threads of different warps accumulating into different
// TMEM columns.  Reports cycles per MMA against the tensor-pipe time N/2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mma_probe.cu -o tools/mma_probe
#include <cstdio>
#include <cstring>
#include "../tf_raft_b200/csrc/tmap.cuh"
using namespace raft;

#if defined(__CUDA_ARCH__)
// A-operand collector: ::fill keeps the A tile in the tensor core's collector buffer, ::lastuse reuses it (no shared-memory read)
__device__ __forceinline__ void umma_f16_afill(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_alast(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#endif

__global__ void __launch_bounds__(128, 1) mma_probe(int n, int reps, int issuers, int distinct, int commit_every, long long* out) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1 << 20);   // sink of the in-loop commits (never completes a phase)
    mbar_init(&bar[3], 1 << 20);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(holder, 512u);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder;
  const uint32_t idesc = make_idesc_f16(128, n);
  if (warp < issuers) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem);
      const uint64_t a0 = make_desc_sw128(sa), b0 = make_desc_sw128(sa + 32 * 1024);
      const uint32_t d = tmem + (uint32_t)(warp * 256);
      const int mine = reps / issuers;
      const long long t0 = clock64();
      for (int i = 0; i < mine; ++i) {
        // distinct: walk the four K=16 slices of a 64-channel chunk like the real mainloop; else the same operands every time
        const uint64_t off = distinct ? (uint64_t)(2 * (i & 3)) : 0;
        if (commit_every == -1) {                        // pairs of MMAs share A: fill, then lastuse with another B slice
          const uint64_t offa = (uint64_t)(2 * ((i >> 1) & 3));
          if (i & 1) umma_f16_alast(d, a0 + offa, b0 + off, idesc, 1u);
          else umma_f16_afill(d, a0 + offa, b0 + off, idesc, i > 0 ? 1u : 0u);
        } else
        umma_f16(d, a0 + off, b0 + off, idesc, i > 0 ? 1u : 0u);
        if (commit_every > 0 && (i + 1) % commit_every == 0) umma_commit(&bar[2 + warp]);   // like the mainloop: frees a stage
      }
      const long long t1 = clock64();
      umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], 0);
      const long long t2 = clock64();
      out[warp * 2] = t1 - t0;
      out[warp * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512u);
#endif
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("%5s %8s %9s %8s %14s %16s %12s\n", "N", "issuers", "operands", "commit/", "issue cyc/MMA", "complete cyc/MMA", "pipe N/2");
  for (int n : {16, 32, 64, 128, 192, 256})
    for (int issuers : {1, 2})
      for (int commit_every : {0, -1})
      for (int distinct : {1}) {
        const int reps = 4080;
        mma_probe<<<1, 128, 100 * 1024>>>(n, reps, issuers, distinct, commit_every, out);
        mma_probe<<<1, 128, 100 * 1024>>>(n, reps, issuers, distinct, commit_every, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[4];
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        const long long issue = issuers == 2 ? (h[0] > h[2] ? h[0] : h[2]) : h[0], done = issuers == 2 ? (h[1] > h[3] ? h[1] : h[3]) : h[1];
        printf("%5d %8d %9s %8d %14.1f %16.1f %12d\n", n, issuers, distinct ? "4 slices" : "same", commit_every, (double)issue / reps, (double)done / reps, n / 2);
      }
  return 0;
}
