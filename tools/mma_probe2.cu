// Micro-probe #2 of tcgen05.mma issue behaviour (bring-up tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../audiolcm_b200/csrc mma_probe2.cu -o mma_probe2
// Question: where does the fixed ~300-cycle cost per "issue group" come from?  Variants (template V):
//   0  group loop over all lanes, `if (leader) { g MMAs }` + __syncwarp per group (current kernel structure)
//   1  whole group loop inside ONE leader region, nothing between groups
//   2  all lanes run the loop; the MMA itself is predicated by elect.sync inside the asm (no branch)
//   3  like 1, accumulators alternate between two TMEM regions per group
//   4  like 1, + tcgen05.commit per group
//   5  like 0, + mbarrier try_wait (already complete) + fence + commit per group
//   6  like 1, + commit + try_wait per group (inside the leader region)
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

using namespace alcm;

struct ProbeArgs {
  int N, G;
  long long* cycles;
};

__device__ __forceinline__ void mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b32 rx;\n\t"
      "elect.sync rx|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int V, int g>
__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t sA = smem_u32(smem), sB = sA + 48 * 1024;
    const uint32_t done_bar = smem_u32(&bars[0]), scratch_bar = smem_u32(&bars[1]), ready_bar = smem_u32(&bars[2]);
    if (leader) mbar_arrive(ready_bar);
    __syncwarp();
    // V=7: no-swizzle A start shifted by one 16-byte row (a conv tap);  V=8: SWIZZLE_128B A and B, aligned;
    // V=9: SWIZZLE_128B, A start shifted by one 128-byte row; V=10: like 9 with base_offset = 1
    uint64_t ad = umma_desc_kmajor(sA, 160 * 16, 128);
    uint64_t bd = umma_desc_kmajor(sB, p.N * 16, 128);
    if (V == 7) ad = umma_desc_kmajor(sA + 16, 178 * 16, 128);
    if (V >= 8) {
      const uint32_t sh = (V >= 9) ? 128u : 0u;
      ad = umma_desc_kmajor(sA + sh, 16, 1024) | ((uint64_t)2 << 61);
      bd = umma_desc_kmajor(sB, 16, 1024) | ((uint64_t)2 << 61);
      if (V == 10) ad |= (uint64_t)1 << 49;
    }
    const uint32_t idesc = umma_idesc(1, p.N);
    const int nG = p.G;
    const long long t0 = clock64();
    uint32_t acc = 0;
    if (V == 0 || V == 5) {
      for (int G = 0; G < nG; ++G) {
        if (V == 5) { mbar_wait(ready_bar, 0); tc_fence_after(); }
        if (leader) {
#pragma unroll
          for (int i = 0; i < g; ++i) { umma_ss<0>(tmem, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, acc); acc = 1; }
          if (V == 5) tc_commit(scratch_bar);
        }
        __syncwarp();
      }
    } else if (V == 2) {
      for (int G = 0; G < nG; ++G) {
#pragma unroll
        for (int i = 0; i < g; ++i) { mma_elect(tmem, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, acc); acc = 1; }
      }
    } else {
      if (leader) {
        for (int G = 0; G < nG; ++G) {
          const uint32_t d = (V == 3) ? tmem + (uint32_t)((G & 1) * 256) : tmem;
          if (V == 6) mbar_wait(ready_bar, 0);
#pragma unroll
          for (int i = 0; i < g; ++i) {
            umma_ss<0>(d, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, (V == 3) ? (uint32_t)(G >= 2 || i > 0) : acc);
            acc = 1;
          }
          if (V == 4 || V == 6) tc_commit(scratch_bar);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) tc_commit(done_bar);
    __syncwarp();
    mbar_wait(done_bar, 0);
    const long long t2 = clock64();
    if (leader && blockIdx.x == 0) {
      p.cycles[0] = t1 - t0;
      p.cycles[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int V, int g>
static void run(int grid, int N, long long* d) {
  ProbeArgs p{N, 512 / g, d};
  cudaFuncSetAttribute(probe_kernel<V, g>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  probe_kernel<V, g><<<grid, 128, 96 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf(" ERR %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("  g=%2d %6.1f/%6.1f", g, h[0] / 512.0, h[1] / 512.0);
}

template <int V>
static void run_v(int grid, long long* d) {
  const int Ns[] = {64, 128, 256};
  for (int N : Ns) {
    printf("V=%d N=%3d:", V, N);
    run<V, 1>(grid, N, d); run<V, 2>(grid, N, d); run<V, 4>(grid, N, d); run<V, 8>(grid, N, d); run<V, 16>(grid, N, d);
    printf("\n");
  }
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  long long* d;
  cudaMalloc(&d, 16);
  printf("grid=%d  (cycles per MMA: issue-loop / until-complete)\n", grid);
  run_v<1>(grid, d); run_v<5>(grid, d); run_v<7>(grid, d); run_v<8>(grid, d); run_v<9>(grid, d); run_v<10>(grid, d);
  return 0;
}
