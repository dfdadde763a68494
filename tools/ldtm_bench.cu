// tools/ldtm_bench.cu -- how fast can the epilogue warps of ONE SM read tensor memory?
// W warps (4, 8, 16; warp w reads the lane quarter w & 3) each issue `iters` tcgen05.ld.32x32b.x32 (128 B per lane,
// 4 KB per warp instruction) over the 512 columns, waiting for each load (or for every second one) like the conv
// epilogue does.  Prints bytes / clk / SM: the ceiling of any epilogue that reads s32 accumulators, 4 B per output
// (8 with two limbs, 12 in the fused block tail).  Diagnostic only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I <pkg>/csrc -o /tmp/ldtm_bench tools/ldtm_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "umma_ptx.cuh"

using namespace slq;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// mode 0: ld x32, wait, (consume)      mode 1: two ld x32 in flight, wait       mode 2: ld x16, wait
// (alu > 0 adds dependent FFMAs per loaded value; the run-time `alu` loop makes mode 0's consume code branchy --
// its numbers measure that code, not the load: read modes 1 and 2)
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int warps, int iters, int mode, int alu, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t tmem_slot;
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)((w & 3) * 32) << 16);
  uint32_t acc = 0;
  float facc = 1.0f;
  __syncthreads();
  const long long t0 = clock64();
  if (w < warps) {
    int col = (w >> 2) * 64;
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        uint32_t r[32];
        tmem_ld32(trow + (col & 511), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (alu == 0) acc ^= r[j];
          else {
            float v = (float)(int)r[j];
            for (int k = 0; k < alu; ++k) v = __fmaf_rn(v, facc, 1.0f);
            acc ^= __float_as_uint(v);
          }
        }
        col += 32;
      } else if (mode == 1) {
        uint32_t r[32], q[32];
        tmem_ld32(trow + (col & 511), r);
        tmem_ld32(trow + ((col + 32) & 511), q);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j];
        col += 64;
        ++it;
      } else {
        uint32_t r[16];
        tmem_ld16(trow + (col & 511), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[j];
        col += 16;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long *cyc;
  uint32_t *sink;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaMalloc(&sink, 4));
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16})
      for (int alu : {0, 3}) {
        if (alu && mode) continue;
        for (int rep = 0; rep < 2; ++rep) {
          ldtm_kernel<<<148, 512>>>(warps, iters, mode, alu, cyc, sink);
          CK(cudaDeviceSynchronize());
        }
        long long h[148];
        CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        const double bytes = (double)warps * iters * 32.0 * (mode == 2 ? 64.0 : 128.0);
        printf("ldtm mode %d (%s) warps %2d alu %d : %8lld clk  %7.1f B/clk/SM  = %5.1f s32 outputs/clk/SM\n", mode,
               mode == 0 ? "x32,wait" : (mode == 1 ? "2 x32,wait" : "x16,wait"), warps, alu, h[0], bytes / (double)h[0],
               bytes / 4.0 / (double)h[0]);
      }
  return 0;
}
