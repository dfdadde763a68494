// tools/issue_bench.cu -- what ONE warp can issue per unit time on B200 when the issue loop carries no
// integer divisions and the issuing instruction sits in convergent code (elect.sync), for
//   (1) cp.async.bulk.tensor (tiled and im2col boxes)   (2) tcgen05.mma kind::i8
// Diagnostic only.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I <pkg>/csrc
//                   -o /tmp/issue_bench tools/issue_bench.cu -lcuda && /tmp/issue_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "umma_ptx.cuh"

using namespace slq;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

struct TArgs {
  int mode;        // 0 tiled, 1 im2col
  int rows;        // box rows
  int row_bytes;   // 64 / 128
  int stages;      // boxes in flight per warp
  int iters;       // boxes per warp
  int warps;       // issuing warps
  int n_tiles;     // tiled: distinct tile rows / rows ; im2col: images
  int W;           // im2col: image width == height
  int ctot;        // im2col: channels of the tensor (row_bytes of them per box); 0 = row_bytes
  int walk;        // im2col: 1 = walk real 128-pixel tiles of the images (base pixel from the tile index)
  int noise;       // other warps of the CTA spin on an mbarrier meanwhile: 0 none, 1 try_wait loop with clock64
                   // checks (the library's mbar_wait), 2 try_wait loop with __nanosleep back-off
};

// every issuing warp owns `stages` slots; the whole warp walks the loop, one elected lane issues
__global__ void __launch_bounds__(640, 1) tma_issue_kernel(const __grid_constant__ CUtensorMap tm, TArgs a,
                                                           long long *cycles, long long *issue_clk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = a.rows * a.row_bytes;
  const uint32_t base = base0 + w * a.stages * box_bytes;
  const uint32_t bars = base0 + 210 * 1024 + w * 128;
  if (w < a.warps && lane == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  __shared__ __align__(8) uint64_t never_bar;
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&never_bar), 1); done_flag = 0; }
  __syncthreads();
  if (w >= a.warps) {
    if (a.noise == 0) return;
    const uint32_t nb = smem_u32(&never_bar);
    while (!done_flag) {
      if (a.noise == 1) {
        const long long t0 = clock64();
        for (int k = 0; k < 64 && !mbar_try_wait(nb, 0); ++k) {
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      } else {
        if (!mbar_try_wait(nb, 0)) __nanosleep(a.noise == 2 ? 100 : 500);
      }
    }
    return;
  }
  int tile = blockIdx.x * a.warps + w;  // tiled: row block ; im2col: image index
  const int tile_step = gridDim.x * a.warps;
  while (tile >= a.n_tiles) tile -= a.n_tiles;
  int stage = 0;
  uint32_t phase = 0;
  int tap_r = 0, tap_s = 0, row0 = 0;
  int cch = 0, bw = -1, bh = -1, bn = 0, wt = blockIdx.x * a.warps + w;  // walk state
  (void)row0;
  long long issue_sum = 0;
  const long long t0 = clock64();
  for (int i = 0; i < a.iters; ++i) {
    if (i >= a.stages) spin_wait(bars + 8 * stage, phase ^ 1);
    const uint32_t dst = base + stage * box_bytes;
    const uint32_t bar = bars + 8 * stage;
    if (elect_one()) {
      mbar_expect_tx(bar, box_bytes);
      const long long c0 = clock64();
      if (a.mode == 0) tma_load_2d(dst, &tm, bar, 0, tile * a.rows);
      else tma_load_im2col_4d(dst, &tm, bar, cch, bw, bh, a.walk ? bn : tile, (uint16_t)tap_s, (uint16_t)tap_r);
      issue_sum += clock64() - c0;
    }
    __syncwarp();
    if (a.mode == 1) {  // the 9 taps of one tile, then another image
      if (a.ctot > a.row_bytes && cch + a.row_bytes < a.ctot) { cch += a.row_bytes; }
      else {
        cch = 0;
        if (++tap_s == 3) { tap_s = 0; if (++tap_r == 3) {
          tap_r = 0; tile += tile_step;
          if (a.walk) {  // next 128-pixel tile: base pixel of output pixel m0 = wt * 128
            wt += tile_step;
            const int hw = a.W * a.W, tiles = 256 * hw / 128;
            if (wt >= tiles) wt -= tiles;
            const int m0 = wt * 128;
            bn = m0 / hw; const int rem = m0 - bn * hw; const int p = rem / a.W;
            bw = rem - p * a.W - 1; bh = p - 1;
          }
        } }
      }
    } else {
      tile += tile_step;
    }
    if (tile >= a.n_tiles) tile -= a.n_tiles;
    if (++stage == a.stages) { stage = 0; phase ^= 1; }
  }
  // drain
  for (int s = 0; s < a.stages && s < a.iters; ++s) {
    const int idx = a.iters - 1 - s;  // the last `stages` boxes
    spin_wait(bars + 8 * (idx % a.stages), (uint32_t)((idx / a.stages) & 1));
  }
  const long long t1 = clock64();
  if (w == 0) {
    const long long v = __shfl_sync(0xffffffffu, issue_sum, 0);
    long long m = issue_sum;
    for (int o = 16; o; o >>= 1) { long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
    (void)v;
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; issue_clk[blockIdx.x] = m; }
  }
  __syncwarp();
  if (w == 0 && lane == 0) done_flag = 1;
}

// ------------------------------------------------------------------------------------------------
struct MArgs { int n; int iters; int accs; int per_commit; int swz; int mode; };  // mode: 0 SS i8, 1 SS f16, 2 TS i8 (A in TMEM), 3 tcgen05.cp only, 4 cp + TS mma
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}


__global__ void __launch_bounds__(128, 1) umma_issue_kernel(MArgs a, long long *cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[1];
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t bar = smem_u32(&bars[0]);
    const uint32_t idesc = (2u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t d128 = make_smem_desc<128>(0), d64 = make_smem_desc<64>(0);
    const uint64_t dflags = a.swz == 128 ? d128 : d64;
    const int kper = a.swz / 32;  // MMAs per K block
    const uint32_t stage16 = (16384 + 36864) >> 4;
    uint32_t commits = 0;
    uint32_t st = 0;
    const long long t0 = clock64();
    for (int i = 0; i < a.iters; ++i) {
      const uint32_t lo = ((base & 0x3FFFFu) >> 4) + st * stage16;
      const uint64_t da = dflags | lo, db = dflags | (lo + (16384 >> 4));
      const uint32_t acc = tmem + (a.accs == 2 ? (uint32_t)(i & 1) * 256u : 0u);
      if (elect_one()) {
        const uint32_t idesc_f16 = (1u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // D=f32, A=B=f16
        const uint32_t ta = tmem + 448 + (uint32_t)(i & 1) * 32;  // A slices in TMEM (8 columns each)
        for (int k = 0; k < kper; ++k) {
          if (a.mode == 0) umma_i8(acc, da + 2 * k, db + 2 * k, idesc, 1);
          else if (a.mode == 1) umma_f16(acc, da + 2 * k, db + 2 * k, idesc_f16, 1);
          else if (a.mode == 2) umma_i8_ts(acc, ta + 8 * k, db + 2 * k, idesc);
          else if (a.mode == 3) tmem_cp_128x256b(ta + 8 * k, da + 2 * k);
          else { tmem_cp_128x256b(ta + 8 * k, da + 2 * k); umma_i8_ts(acc, ta + 8 * k, db + 2 * k, idesc); }
        }
        if ((i + 1) % a.per_commit == 0) umma_commit(bar);
      }
      __syncwarp();
      if ((i + 1) % a.per_commit == 0) {
        ++commits;
        if ((commits & 15) == 0) spin_wait(bar, (commits - 1) & 1);
      }
      if (++st == 3) st = 0;
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    ++commits;
    spin_wait(bar, (commits - 1) & 1);
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Interference: warps 0/1 issue tcgen05.mma (same or different accumulators) while warp 2 issues TMA boxes.
struct CArgs { int n; int mma_warps; int same_acc; int tma_on; int iters; };

__global__ void __launch_bounds__(128, 1) combo_kernel(const __grid_constant__ CUtensorMap tm, CArgs a, long long *cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[16];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
  if (w == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64(), t1 = t0;
  if (w < a.mma_warps) {
    const uint32_t bar = smem_u32(&bars[w]);
    const uint32_t idesc = (2u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dflags = make_smem_desc<128>(0);
    const uint32_t acc = tmem + ((a.same_acc || w == 0) ? 0u : 256u);
    uint32_t commits = 0;
    for (int i = 0; i < a.iters; ++i) {
      const uint32_t lo = ((base & 0x3FFFFu) >> 4) + (uint32_t)((i + w) % 2) * ((16384 + 36864) >> 4);
      const uint64_t da = dflags | lo, db = dflags | (lo + (16384 >> 4));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_i8(acc, da + 2 * k, db + 2 * k, idesc, 1);
        if ((i & 3) == 3) umma_commit(bar);
      }
      __syncwarp();
      if ((i & 3) == 3) { ++commits; if ((commits & 15) == 0) spin_wait(bar, (commits - 1) & 1); }
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    ++commits;
    spin_wait(bar, (commits - 1) & 1);
    t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 4 + w] = t1 - t0;
  } else if (w == 2 && a.tma_on) {
    const uint32_t tb = base + 110 * 1024;
    int stage = 0; uint32_t phase = 0; int tile = blockIdx.x;
    for (int i = 0; i < a.iters; ++i) {
      const uint32_t bar = smem_u32(&bars[4 + stage]);
      if (i >= 4) spin_wait(bar, phase ^ 1);
      if (elect_one()) { mbar_expect_tx(bar, 16384); tma_load_2d(tb + stage * 16384, &tm, bar, 0, tile * 128); }
      __syncwarp();
      tile += gridDim.x; if (tile >= 1536) tile -= 1536;
      if (++stage == 4) { stage = 0; phase ^= 1; }
    }
    for (int s = 0; s < 4; ++s) { const int idx = a.iters - 1 - s; spin_wait(smem_u32(&bars[4 + (idx & 3)]), (uint32_t)((idx >> 2) & 1)); }
    t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 4 + 2] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (w == 3) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory"); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static double avg148(long long *d) {
  long long h[148];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  return s / 148;
}

int main() {
  void *p1 = nullptr, *p2 = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p1, cudaEnableDefault, &qr));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p2, cudaEnableDefault, &qr));
  EncodeTiledFn enc_tiled = (EncodeTiledFn)p1;
  EncodeIm2colFn enc_im2col = (EncodeIm2colFn)p2;
  uint8_t *buf;
  const size_t big = 256ull << 20;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  long long *cyc, *iss;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaMalloc(&iss, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(tma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
  CK(cudaFuncSetAttribute(umma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));

  struct TC { int mode, rows, row_bytes, stages, warps; size_t bytes; int W; int noise = 0; int ctot = 0; int walk = 0; };
  const TC tcs[] = {
      {0, 128, 128, 8, 1, 24u << 20, 0},  {0, 128, 128, 4, 1, 24u << 20, 0}, {0, 32, 128, 8, 1, 24u << 20, 0},
      {0, 256, 128, 4, 1, 24u << 20, 0},  {0, 128, 64, 8, 1, 24u << 20, 0},  {0, 128, 128, 4, 2, 24u << 20, 0},
      {0, 128, 128, 3, 4, 24u << 20, 0},  {0, 128, 128, 8, 1, 256u << 20, 0}, {0, 128, 128, 3, 4, 256u << 20, 0},
      {1, 128, 128, 8, 1, 0, 28},         {1, 128, 128, 4, 2, 0, 28},        {1, 128, 128, 3, 4, 0, 28},
      {1, 128, 64, 8, 1, 0, 56},          {1, 128, 64, 4, 2, 0, 56},         {1, 128, 64, 3, 4, 0, 56},
      {1, 256, 64, 4, 1, 0, 56},          {1, 256, 64, 3, 2, 0, 56},
      {1, 128, 128, 4, 1, 0, 28, 1},      {1, 128, 128, 4, 1, 0, 28, 2},     {1, 128, 128, 4, 1, 0, 28, 3},
      {1, 128, 128, 4, 2, 0, 28, 1},      {1, 128, 128, 4, 2, 0, 28, 2},     {0, 128, 128, 4, 1, 24u << 20, 0, 1},
      {0, 128, 128, 4, 1, 24u << 20, 0, 2},
      {1, 128, 128, 4, 1, 0, 14, 0, 256, 1}, {1, 128, 128, 4, 2, 0, 14, 0, 256, 1}, {1, 128, 128, 3, 4, 0, 14, 0, 256, 1},
      {1, 128, 128, 4, 1, 0, 28, 0, 128, 1}, {1, 128, 128, 4, 2, 0, 28, 0, 128, 1},
      {1, 128, 128, 4, 1, 0, 7, 0, 512, 1},  {1, 128, 128, 4, 2, 0, 7, 0, 512, 1},
      {1, 128, 128, 4, 1, 0, 14, 0, 256, 0}, {1, 128, 128, 4, 2, 0, 14, 0, 256, 0},
  };
  for (const TC &c : tcs) {
    CUtensorMap tm;
    TArgs a{};
    a.mode = c.mode; a.rows = c.rows; a.row_bytes = c.row_bytes; a.stages = c.stages; a.warps = c.warps; a.noise = c.noise; a.ctot = c.ctot ? c.ctot : c.row_bytes; a.walk = c.walk;
    a.iters = 1800;
    const CUtensorMapSwizzle sw = c.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r;
    if (c.mode == 0) {
      const cuuint64_t M = c.bytes / c.row_bytes;
      cuuint64_t dims[2] = {(cuuint64_t)c.row_bytes, M};
      cuuint64_t strides[1] = {(cuuint64_t)c.row_bytes};
      cuuint32_t box[2] = {(cuuint32_t)c.row_bytes, (cuuint32_t)c.rows};
      cuuint32_t es[2] = {1, 1};
      r = enc_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      a.n_tiles = (int)(M / c.rows);
    } else {
      const int W = c.W, C = c.ctot ? c.ctot : c.row_bytes, N = 256;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)W, (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)W * W * C};
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc_im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, buf, dims, strides, lower, upper, (cuuint32_t)c.row_bytes,
                     (cuuint32_t)c.rows, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      a.n_tiles = N; a.W = W;
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    tma_issue_kernel<<<148, 640, 216 * 1024>>>(tm, a, cyc, iss);
    CK(cudaDeviceSynchronize());
    const double clk = avg148(cyc), icl = avg148(iss);
    printf("%s %3dB x%3d rows %s S=%d w%d noise=%d : %7.1f clk/box/warp  %6.1f B/clk/SM   issue instr %6.1f clk\n",
           c.mode ? "im2col" : "tiled ", c.row_bytes, c.rows, c.mode ? (c.W == 28 ? "28x28" : c.W == 56 ? "56x56" : c.W == 14 ? (c.walk ? "14x14 walk" : "14x14 fix ") : " 7x7  walk") : (c.bytes > (64u << 20) ? "DRAM " : "L2   "),
           c.stages, c.warps, c.noise, clk / a.iters, (double)a.iters * c.warps * c.rows * c.row_bytes / clk, icl / a.iters);
  }

  if (getenv("COMBO")) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {128, (cuuint64_t)(24u << 20) / 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    CK(cudaFuncSetAttribute(combo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long *cyc4;
    CK(cudaMalloc(&cyc4, 148 * 4 * sizeof(long long)));
    for (int n : {144, 256}) for (int mw : {1, 2}) for (int same : {0, 1}) for (int tma : {0, 1}) {
      if (mw == 1 && same == 1) continue;
      CArgs a{n, mw, same, tma, 2000};
      CK(cudaMemset(cyc4, 0, 148 * 4 * sizeof(long long)));
      combo_kernel<<<148, 128, 200 * 1024>>>(tm, a, cyc4);
      CK(cudaDeviceSynchronize());
      long long h[148 * 4];
      CK(cudaMemcpy(h, cyc4, sizeof(h), cudaMemcpyDeviceToHost));
      double m0 = 0, m1 = 0, t2 = 0;
      for (int i = 0; i < 148; ++i) { m0 += h[4 * i]; m1 += h[4 * i + 1]; t2 += h[4 * i + 2]; }
      printf("combo N=%3d mma_warps=%d same_acc=%d tma=%d : warp0 %6.1f clk/MMA  warp1 %6.1f clk/MMA  tma %6.1f clk/box\n", n, mw, same, tma,
             m0 / 148 / a.iters / 4, m1 / 148 / a.iters / 4, t2 / 148 / a.iters);
    }
    return 0;
  }
  if (getenv("SKIP_UMMA")) return 0;
  const char *mname[] = {"SS i8 ", "SS f16", "TS i8 ", "cp    ", "cp+TS "};
  for (int mode = 0; mode < 5; ++mode) {
    for (int swz : {128, 64}) {
      for (int n : {64, 80, 128, 144, 256}) {
        if (mode == 3 && n != 64) continue;
        for (int accs : {1, 2}) {
          if (accs == 2 && (mode == 3 || n > 128)) continue;
          MArgs m{n, 4000, accs, 4, swz, mode};
          umma_issue_kernel<<<148, 128, 220 * 1024>>>(m, cyc);
          CK(cudaDeviceSynchronize());
          const double clk = avg148(cyc) / m.iters;
          const int kper = swz / 32;
          printf("umma %s SWZ=%3d N=%3d accs=%d : %7.1f clk per K block = %6.1f clk/MMA  %5.0f MAC/clk (N/2 = %5.1f)\n",
                 mname[mode], swz, n, accs, clk, clk / kper, 128.0 * n * 32 * kper / clk, n / 2.0);
        }
      }
    }
  }
  return 0;
}
