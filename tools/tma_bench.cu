// tools/tma_bench.cu -- measures what one SM's TMA engine sustains on B200 for the box shapes the
// conv kernel uses (tiled / im2col, 64- and 128-byte rows, L2-resident and DRAM-sized tensors).
// Diagnostic only (not part of the library).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_bench tools/tma_bench.cu -lcuda
//   /tmp/tma_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

struct Args {
  int mode;       // 0 tiled load, 1 im2col load, 2 tiled store
  int row_bytes;  // 64 / 128
  int rows;       // box rows (pixels)
  int stages;
  int iters;      // loads per CTA
  long long tiles;  // distinct tile positions
  int W, H;       // im2col geometry (Wo = W, Ho = H, 3x3 pad 1)
  int warps;      // issuing warps per CTA (each with its own ring)
  int lanes;      // issuing lanes per warp (each with its own ring)
  int share;      // >1: groups of `share` consecutive CTAs load the SAME tiles (weights-like reuse)
};

__global__ void __launch_bounds__(256, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, Args a, long long *cycles) {
  extern __shared__ uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = (threadIdx.x >> 5) * a.lanes + lane;  // issuer id
  const int box_bytes = a.row_bytes * a.rows;
  const uint32_t base = ((smem_u32(smem_raw) + 1023u) & ~1023u) + warp * a.stages * box_bytes;
  const uint32_t bars = ((smem_u32(smem_raw) + 1023u) & ~1023u) + 200 * 1024 + warp * 128;
  if (lane < a.lanes && (threadIdx.x >> 5) < a.warps) {
    for (int s = 0; s < a.stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x >> 5) >= a.warps) return;
  if (a.lanes == 0) {  // convergent issue: the whole warp walks the loop, one elected lane issues
    const int w = threadIdx.x >> 5;
    const uint32_t cbase = ((smem_u32(smem_raw) + 1023u) & ~1023u) + w * a.stages * box_bytes;
    const uint32_t cbars = ((smem_u32(smem_raw) + 1023u) & ~1023u) + 200 * 1024 + w * 128;
    if (lane == 0) {
      for (int s = 0; s < a.stages; ++s) mbar_init(cbars + 8 * s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < a.iters; ++i) {
      const int s = i % a.stages;
      const uint32_t dst = cbase + s * box_bytes;
      const long long tile = ((long long)(blockIdx.x * a.warps + w) + (long long)i * gridDim.x * a.warps) % a.tiles;
      if (i >= a.stages) mbar_wait(cbars + 8 * s, ((i / a.stages) - 1) & 1);
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
        mbar_expect_tx(cbars + 8 * s, box_bytes);
        if (a.mode == 0) {
          tma_load_2d(dst, &tm, cbars + 8 * s, 0, (int)(tile * a.rows));
        } else {
          const long long m0 = tile * a.rows;
          const int q = (int)(m0 % a.W), p = (int)((m0 / a.W) % a.H), n = (int)(m0 / ((long long)a.W * a.H));
          const int tap = i % 9;
          tma_load_im2col_4d(dst, &tm, cbars + 8 * s, 0, q - 1, p - 1, n, (uint16_t)(tap % 3), (uint16_t)(tap / 3));
        }
      }
      __syncwarp();
    }
    for (int i = a.iters - a.stages < 0 ? 0 : a.iters - a.stages; i < a.iters; ++i)
      mbar_wait(cbars + 8 * (i % a.stages), (i / a.stages) & 1);
    if (w == 0 && lane == 0) cycles[blockIdx.x] = clock64() - t0;
    return;
  }
  if (lane >= a.lanes) return;
  const int issuers = a.warps * a.lanes;
  const long long t0 = clock64();
  for (int i = 0; i < a.iters; ++i) {
    const int s = i % a.stages;
    const uint32_t dst = base + s * box_bytes;
    const int bid = a.share > 1 ? (int)(blockIdx.x % a.share == 0 ? blockIdx.x / a.share : blockIdx.x / a.share) : (int)blockIdx.x;
    const long long tile = ((long long)(bid * issuers + warp) + (long long)i * gridDim.x * issuers) % a.tiles;
    if (a.mode == 2) {
      if (i >= a.stages) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(7) : "memory");
      tma_store_2d(&tm, dst, 0, (int)(tile * a.rows));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      continue;
    }
    if (i >= a.stages) mbar_wait(bars + 8 * s, ((i / a.stages) - 1) & 1);
    mbar_expect_tx(bars + 8 * s, box_bytes);
    if (a.mode == 0) {
      tma_load_2d(dst, &tm, bars + 8 * s, 0, (int)(tile * a.rows));
    } else {
      const long long m0 = tile * a.rows;
      const int q = (int)(m0 % a.W), p = (int)((m0 / a.W) % a.H), n = (int)(m0 / ((long long)a.W * a.H));
      const int tap = i % 9;
      tma_load_im2col_4d(dst, &tm, bars + 8 * s, 0, q - 1, p - 1, n, (uint16_t)(tap % 3), (uint16_t)(tap / 3));
    }
  }
  if (a.mode == 2) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    for (int i = a.iters - a.stages < 0 ? 0 : a.iters - a.stages; i < a.iters; ++i)
      mbar_wait(bars + 8 * (i % a.stages), (i / a.stages) & 1);
  }
  if (warp == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *p1 = nullptr, *p2 = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p1, cudaEnableDefault, &qr));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p2, cudaEnableDefault, &qr));
  EncodeTiledFn enc_tiled = (EncodeTiledFn)p1;
  EncodeIm2colFn enc_im2col = (EncodeIm2colFn)p2;
  const size_t big = 1ull << 30;
  uint8_t *buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  long long *cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));

  struct Case { const char *name; int mode, row_bytes, rows, stages; size_t bytes; int W; int warps = 1; int lanes = 1; int share = 1; };
  const size_t l2 = 24u << 20, a28 = 256ull * 28 * 28 * 128, a56 = 256ull * 56 * 56 * 64;
  std::vector<Case> cases = {
      {"tiled 128B x128 L2 S=2 l4 distinct", 0, 128, 128, 2, l2, 0, 1, 4, 1},
      {"tiled 128B x128 L2 S=2 l4 share 2", 0, 128, 128, 2, l2, 0, 1, 4, 2},
      {"tiled 128B x128 L2 S=2 l4 share 8", 0, 128, 128, 2, l2, 0, 1, 4, 8},
      {"tiled 128B x128 L2 S=2 l4 share 74", 0, 128, 128, 2, l2, 0, 1, 4, 74},
      {"tiled 128B x128 L2 S=2 l4 share 148", 0, 128, 128, 2, l2, 0, 1, 4, 148},
      {"tiled 128B x128 1MB S=2 l4 share 74", 0, 128, 128, 2, 1u << 20, 0, 1, 4, 74},
      {"tiled 128B x128 1MB S=2 l4 distinct", 0, 128, 128, 2, 1u << 20, 0, 1, 4, 1},
  };
  for (auto &c : cases) {
    CUtensorMap tm;
    Args a{};
    a.mode = c.mode; a.row_bytes = c.row_bytes; a.rows = c.rows; a.stages = c.stages; a.warps = c.warps; a.lanes = c.lanes; a.share = c.share;
    const CUtensorMapSwizzle sw = c.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r;
    if (c.mode != 1) {
      const cuuint64_t M = c.bytes / c.row_bytes;
      cuuint64_t dims[2] = {(cuuint64_t)c.row_bytes, M};
      cuuint64_t strides[1] = {(cuuint64_t)c.row_bytes};
      cuuint32_t box[2] = {(cuuint32_t)c.row_bytes, (cuuint32_t)c.rows};
      cuuint32_t es[2] = {1, 1};
      r = enc_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      a.tiles = (long long)(M / c.rows);
    } else {
      const int W = c.W, H = c.W, C = c.row_bytes, N = 256;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)H * W * C};
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc_im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, buf, dims, strides, lower, upper, (cuuint32_t)C,
                     (cuuint32_t)c.rows, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      a.W = W; a.H = H;
      a.tiles = (long long)N * H * W / c.rows;
    }
    if (r != CUDA_SUCCESS) { printf("%-44s encode failed %d\n", c.name, (int)r); continue; }
    const long long box_bytes = (long long)c.row_bytes * c.rows;
    a.iters = (int)((768ll << 20) / 148 / box_bytes / (c.warps * (c.lanes ? c.lanes : 1)));  // ~768 MB moved per run
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      tma_kernel<<<148, 256, 227 * 1024>>>(tm, a, cyc);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const int nl = c.lanes ? c.lanes : 1;
    const double bytes = 148.0 * a.iters * box_bytes * c.warps * nl;
    printf("%-40s %7.2f TB/s  %6.1f B/clk/SM  %7.1f clk/box/issuer  (%.3f ms)\n", c.name, bytes / best / 1e9,
           (double)a.iters * box_bytes * c.warps * nl / avg, avg / a.iters, best);
  }
  return 0;
}
