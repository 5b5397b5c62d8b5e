// Micro-probe of tcgen05.mma issue behaviour (bring-up tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../audiolcm_b200/csrc mma_probe.cu -o mma_probe
// One CTA per SM (grid given), warp 1 issues G groups of g MMAs (M=128, N, K=16, bf16, no-swizzle K-major
// operands on zeroed smem).  Options per group: commit to an mbarrier, try_wait on an already
// completed mbarrier, tcgen05 fence.  Reports cycles per MMA measured with clock64 in the issuing warp
// (from first issue to the completion of the final commit).
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

using namespace alcm;

struct ProbeArgs {
  int N, g, G, commit, wait, fence, swz, single;
  long long* cycles;
};

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
    tmem_alloc(smem_u32(&tmem_slot), 256);
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
    if (leader) mbar_arrive(ready_bar);  // phase 0 of ready_bar is complete -> try_wait(parity 0) succeeds at once
    __syncwarp();
    uint64_t ad, bd;
    if (p.swz) {  // SWIZZLE_128B K-major: SBO = 1024, layout_type 2
      ad = umma_desc_kmajor(sA, 16, 1024) | ((uint64_t)2 << 61);
      bd = umma_desc_kmajor(sB, 16, 1024) | ((uint64_t)2 << 61);
    } else {
      ad = umma_desc_kmajor(sA, 160 * 16, 128);
      bd = umma_desc_kmajor(sB, p.N * 16, 128);
    }
    const uint32_t idesc = umma_idesc(1, p.N);
    const long long t0 = clock64();
    uint32_t acc = 0;
    if (p.single) {  // variant B: the whole loop inside ONE elected-thread region, parameters in locals
      if (leader) {
        const int nG = p.G, ng = p.g, do_commit = p.commit, do_wait = p.wait;
        for (int G = 0; G < nG; ++G) {
          if (do_wait) mbar_wait(ready_bar, 0);
          tc_fence_after();
          for (int i = 0; i < ng; ++i) {
            umma_ss<0>(tmem, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, acc);
            acc = 1;
          }
          if (do_commit) tc_commit(scratch_bar);
        }
      }
      __syncwarp();
    } else
    for (int G = 0; G < p.G; ++G) {
      if (p.wait) mbar_wait(ready_bar, 0);
      if (p.fence) tc_fence_after();
      if (leader) {
        for (int i = 0; i < p.g; ++i) {
          umma_ss<0>(tmem, ad + (uint64_t)(2 * i), bd + (uint64_t)(2 * i), idesc, acc);
          acc = 1;
        }
        if (p.commit) tc_commit(scratch_bar);
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
    tmem_dealloc(tmem, 256);
  }
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int Ns[] = {64, 128, 256};
  const int gs[] = {1, 2, 4, 8, 16};
  printf("grid=%d  (cycles per MMA: issue-loop / until-complete)\n", grid);
  for (int single = 0; single < 2; ++single)
    for (int N : Ns)
      for (int mode = 0; mode < 4; mode += 2) {  // 0: bare, 2: +commit+wait
        printf("single=%d N=%3d mode=%d:", single, N, mode);
        const int swz = 0;
        for (int g : gs) {
          ProbeArgs p{N, g, 512 / g, mode >= 1, mode >= 2, mode >= 3, swz, single, d};
          probe_kernel<<<grid, 128, 96 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" ERR %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("  g=%2d %6.1f/%6.1f", g, h[0] / 512.0, h[1] / 512.0);
        }
        printf("\n");
      }
  return 0;
}
