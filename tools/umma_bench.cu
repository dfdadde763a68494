// tools/umma_bench.cu -- tcgen05.mma kind::i8 issue/execute rate from ONE thread on resident smem
// operands (no TMA), to separate the tensor-pipe floor from the feed pipeline of conv_umma.cu.
// Diagnostic only.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I <pkg>/csrc -I include
//                   -o /tmp/umma_bench tools/umma_bench.cu -lcuda && /tmp/umma_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "umma_ptx.cuh"

using namespace slq;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// kb_per_commit MMAs groups of (SWZ/32) instructions, then one commit; `stages` distinct smem tiles
__global__ void __launch_bounds__(128, 1) umma_kernel(int n, int iters, int stages, int commit_every, int issuers, int nomask, long long *cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[2];
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int who = threadIdx.x >> 5;  // issuer = lane 0 of warp `who`
  if ((threadIdx.x & 31) == 0 && who < issuers) {
    const uint32_t bar = smem_u32(&bars[who]);
    const uint32_t idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int stage_bytes = 16384 + 36864;  // A 128x128 + B up to 288x128
    const long long t0 = clock64();
    uint32_t commits = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t sa = base + ((i + who) % stages) * stage_bytes;
      const uint64_t da = make_smem_desc<128>(sa);
      const uint64_t db = make_smem_desc<128>(sa + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (nomask) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + who * 256), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(1) : "memory");
        } else {
          umma_i8(tmem + who * 256, da + 2 * k, db + 2 * k, idesc, 1);
        }
      }
      if ((i + 1) % commit_every == 0) {
        umma_commit(bar);
        ++commits;
        if (commits % 16 == 0) mbar_wait(bar, (commits - 1) & 1);  // drain now and then
      }
    }
    if (commits % 16 != 0) mbar_wait(bar, (commits - 1) & 1);
    umma_commit(bar);
    ++commits;
    mbar_wait(bar, (commits - 1) & 1);
    if (who == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  long long *cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int iters = 4000;
  for (int n : {80, 144, 256}) {
    for (int issuers : {1, 2}) {
      for (int nomask : {0, 1}) {
        umma_kernel<<<148, 128, 220 * 1024>>>(n, iters, 4, 4, issuers, nomask, cyc);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148;
        const double per = avg / iters / issuers;  // cycles per K block of the whole SM
        printf("N=%3d issuers=%d nomask=%d : %7.1f clk per K block (4 MMAs of K=32B) = %6.1f clk/MMA, %5.0f MAC/clk\n", n,
               issuers, nomask, per, per / 4, 128.0 * n * 128 / per);
      }
    }
  }
  return 0;
}
