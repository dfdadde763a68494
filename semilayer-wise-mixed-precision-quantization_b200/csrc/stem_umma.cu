// csrc/stem_umma.cu -- the un-quantised stem on Blackwell tensor cores, ONE fused kernel.
//
// Replaces resnet.py:206-209 (conv1 7x7 s2 p3, 3->64, fp32 weights; bn1; relu; maxpool 3x3 s2 p1).
// The reference keeps these weights in fp32, so this layer is NOT quantised to 8 bit:
// operands go through the tensor core as fp16 (11-bit significand, fp32 accumulation in TMEM),
// a relative error ~3e-4, an order of magnitude below the u8 activation quantisation that follows.
//
// Data path: the fp32 NCHW image is read from HBM exactly once and the pooled u8 NHWC tensor is
// written exactly once; im2col, the conv output and the pre-pool rows never leave the SM.
//   work unit  = (image, 14 pooled rows) -> 29 conv rows (one recomputed at the seam), units are
//                dealt round-robin to one persistent CTA per SM
//   builders   (warps 0-15) fp32 input rows -> fp16 row ring in smem (next row's loads are in flight
//                while this row is built) -> implicit-GEMM A tile of one conv row in the UMMA
//                K-major SWIZZLE_128B layout: row = output column q, K = (c, r, s padded to 8):
//                every 16-byte chunk is the 8 consecutive input pixels x[c][2p-3+r][2q-3 .. 2q+4]
//   MMA        (warps 24,25) D[128 x 64] = A[128 x 192] * W[64 x 192]^T, tcgen05.mma kind::f16, weights
//                resident in smem, A and the TMEM accumulator double-buffered.  One thread gets a
//                tcgen05.mma out only every ~100 cycles whatever its size (32 tensor cycles here), so
//                even and odd conv rows are issued by two different warps
//   epilogue   (warps 16-23) tcgen05.ld -> folded BN + ReLU -> u8 (scale of the POOLED tensor: max
//                commutes with a monotone map) -> 4-row ring in smem -> 3x3/s2 max-pool -> coalesced
//                stores of the pooled row
// Calibration (SLQ_OUT_F32) writes the fp32 conv rows to a scratch tensor and pools them with the
// fp32 pool kernel of layers.cu.
#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>
#include <new>

#include "conv_common.cuh"
#include "umma_ptx.cuh"

struct slq_stem {
  int N, H, W, Hc, Wc, Hp, Wp;
  __half *wh;  // [64, 192] fp16, K order (c*7 + r)*8 + s, zero padded
  int num_ctas;
};

namespace slq {

constexpr int kSfK = 192;                       // padded K: 24 chunks of 8 halfs (21 real)
constexpr int kSfChunks = 24;
constexpr int kSfABytes = 3 * 16384;            // one A tile: 3 K blocks x (128 rows x 128 B)
constexpr int kSfBBytes = 3 * 8192;             // weights: 3 K blocks x (64 rows x 128 B)
constexpr int kSfRowP = 264;                    // halfs per ring row: column cc = w + 3, w in [-3, 2*127+4]
constexpr int kSfRing = 16;                     // input rows kept (7 live + 2 arriving fit twice)
constexpr int kSfRingBytes = kSfRing * 3 * kSfRowP * 2;
constexpr int kSfConvRing = 4;                  // u8 conv rows kept for pooling
constexpr int kSfConvRowBytes = 128 * 64;
constexpr int kSfAOff = 0;
constexpr int kSfBOff = kSfAOff + 2 * kSfABytes;
constexpr int kSfRingOff = kSfBOff + kSfBBytes;
constexpr int kSfConvOff = kSfRingOff + ((kSfRingBytes + 1023) / 1024) * 1024;
constexpr int kSfPrmOff = kSfConvOff + kSfConvRing * kSfConvRowBytes;
constexpr int kSfBarOff = kSfPrmOff + 64 * 8;
constexpr int kSfSmemBytes = 1024 + kSfBarOff + 128;
constexpr int kSfThreads = 26 * 32;  // 16 builder + 8 epilogue + 2 MMA warps
constexpr int kSfBuilders = 512, kSfEpi = 256;
constexpr int kSfBW = kSfBuilders / 32;          // builder warps; epilogue = warps kSfBW..kSfBW+7, MMA = the last two
constexpr int kSfUnitRows = 14;                 // pooled rows per work unit
constexpr int kSfTmemCols = 128;                // 2 accumulators x 64 columns
static_assert(kSfSmemBytes <= 232448, "stem kernel exceeds 227 KB of shared memory");

// w fp32 [64, 3, 7, 7] -> wh fp16 [64, 192]: wh[oc][(c*7 + r)*8 + s], zero for s == 7 and the pad chunks
__global__ void stem_weights_kernel(const float *__restrict__ w, __half *__restrict__ wh) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * kSfK) return;
  const int k = idx % kSfK, oc = idx / kSfK;
  const int chunk = k >> 3, s = k & 7;
  float v = 0.f;
  if (chunk < 21 && s < 7) {
    const int c = chunk / 7, r = chunk % 7;
    v = w[((oc * 3 + c) * 7 + r) * 7 + s];
  }
  wh[idx] = __float2half_rn(v);
}

struct StemArgs {
  const void *x;       // [N,3,H,W]: fp32 | fp16 | u8 (slq_stem_launch_in)
  float nmean[3], nstd[3];  // u8 input: value = (u/255 - mean[c]) / std[c]  (imagenet.py:14-15 ToTensor + Normalize)
  int N, H, W, Hc, Wc, Hp, Wp;
  const __half *wh;
  const float *bn_a, *bn_b, *act_scales;
  int out_id, out_mode;
  void *out;  // u8 pooled [N,Hp,Wp,64]  |  fp32 conv rows [N,Hc,Wc,64] (SLQ_OUT_F32)
  uint32_t *out_rowsum;  // [N*Hp*Wp] channel sum of every pooled u8 pixel (or NULL): slq_epilogue.in_rowsum of layer 1
  int units_per_img, total_units;
  long long *trace;  // debug timeline of CTA 0 (slq_debug_set_trace): 24 issuers x cap/24 events
  int trace_cap;
  int dbg;  // $SLQ_STEM_DBG bit mask (timing experiments only): 1 skip A build, 2 skip pooling, 4 skip epilogue math
};

#define SF_DBG(a) (kDebugTrace ? (a).dbg : 0)
__device__ __forceinline__ void stem_trace(const StemArgs &a, int issuer, int &n, int ev, int idx) {
  if (!kDebugTrace || a.trace == nullptr || blockIdx.x != 0) return;
  const int per = a.trace_cap / 24;
  if (n < per) {
    long long *p = a.trace + 3LL * (issuer * per + n);
    p[0] = ev + 1;
    p[1] = idx;
    p[2] = clock64();
    ++n;
  }
}

// A load whose ISSUE POINT the compiler must keep: __ldg() is an invariant load that gets sunk to its
// first use (after the A-tile build), which exposes the whole DRAM latency once per conv row.
// IN selects the element type of the image in HBM (SLQ_IN_*); the value is returned RAW (fp32 bits, fp16
// bits or the byte) and turned into the fp32 pixel by stem_pixel() after the build, so that only the
// load itself is pinned.
template <int IN>
__device__ __forceinline__ uint32_t ldg_pinned(const void *base, long long idx) {
  uint32_t v;
  if (IN == SLQ_IN_F32) asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<const float *>(base) + idx) : "memory");
  else if (IN == SLQ_IN_F16) {
    uint16_t h;
    asm volatile("ld.global.nc.b16 %0, [%1];" : "=h"(h) : "l"(reinterpret_cast<const uint16_t *>(base) + idx) : "memory");
    v = h;
  } else asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<const uint8_t *>(base) + idx) : "memory");
  return v;
}
// raw element -> the fp32 pixel value the reference's model would see (then rounded to fp16 by the caller)
template <int IN>
__device__ __forceinline__ float stem_pixel(uint32_t raw, float mean, float stdv) {
  if (IN == SLQ_IN_F32) return __uint_as_float(raw);
  if (IN == SLQ_IN_F16) return __half2float(__ushort_as_half((unsigned short)raw));
  // torchvision ToTensor: u8 -> fp32 / 255 ; Normalize: (t - mean) / std -- three separately rounded fp32 ops
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)raw, 255.0f), mean), stdv);
}

// conv rows [p0, p1) of unit u and the pooled rows [j0, j1) it owns
__device__ __forceinline__ void unit_rows(const StemArgs &a, int u, int &n, int &j0, int &j1, int &p0, int &p1) {
  n = u / a.units_per_img;
  const int q = u - n * a.units_per_img;
  j0 = q * kSfUnitRows;
  j1 = min(j0 + kSfUnitRows, a.Hp);
  p0 = max(2 * j0 - 1, 0);
  p1 = min(2 * j1, a.Hc);  // last conv row needed is 2*(j1-1)+1
}

// Chunks [J0, J1) of output column q of one conv row: every index that depends on the chunk is a
// compile-time constant after unrolling (c, r, K block, chunk-in-block), so a chunk costs 4 LDS.32 +
// 1 STS.128 + a handful of integer ops.  ring_q = ring + 2q halfs; atile_q = tile + q*128; qx = q & 7.
template <int J0, int J1>
__device__ __forceinline__ void stem_build_chunks(const __half *ring_q, uint8_t *atile_q, int qx, int p) {
#pragma unroll
  for (int j = J0; j < J1; ++j) {
    constexpr int dummy = 0;
    (void)dummy;
    const int c = j / 7, r = j % 7;
    const int slot = (2 * p + 13 + r) & (kSfRing - 1);  // (2p - 3 + r + 16) mod 16
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ring_q + (slot * 3 + c) * kSfRowP);
    const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
    const int kb = j >> 3, cj = j & 7;
    *reinterpret_cast<uint4 *>(atile_q + kb * 16384 + ((cj ^ qx) << 4)) = v;
  }
}

template <int IN>
__global__ void __launch_bounds__(kSfThreads, 1) stem_fused_kernel(const StemArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kSfBarOff;
  auto afull_bar = [&](int b) { return bar_base + 8u * b; };         // A tile built      (256 arrivals)
  auto aempty_bar = [&](int b) { return bar_base + 8u * (2 + b); };  // A tile consumed   (commit)
  auto tfull_bar = [&](int b) { return bar_base + 8u * (4 + b); };   // accumulator ready (commit)
  auto tempty_bar = [&](int b) { return bar_base + 8u * (6 + b); };  // accumulator drained (256)
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSfBarOff + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(afull_bar(b), kSfBuilders);
      mbar_init(aempty_bar(b), 1);
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kSfEpi);
    }
    fence_barrier_init();
  }
  if (warp == kSfBW + 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void *)tmem_slot)),
                 "r"(kSfTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the input ring (its pad columns are never written again) and stage weights + BN constants
  for (int i = threadIdx.x; i < kSfRingBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(smem + kSfRingOff)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 64 * kSfChunks; i += blockDim.x) {
    const int oc = i / kSfChunks, j = i % kSfChunks;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.wh + oc * kSfK + j * 8));
    const int kb = j >> 3, cj = j & 7;
    *reinterpret_cast<uint4 *>(smem + kSfBOff + kb * 8192 + oc * 128 + ((cj ^ (oc & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < 2 * 128 * 3; i += blockDim.x) {  // K-pad chunks 21..23 of both A tiles stay zero
    const int ab = i / 384, rem = i % 384, q = rem / 3, j = 21 + rem % 3;
    *reinterpret_cast<uint4 *>(smem + kSfAOff + ab * kSfABytes + (j >> 3) * 16384 + q * 128 + (((j & 7) ^ (q & 7)) << 4)) =
        make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x < 64) {  // u8 output: the re-quantisation multiply is folded into the BN constants
    const float inv = a.out_mode == SLQ_OUT_F32 ? 1.f : __fdiv_rn(1.0f, a.act_scales[a.out_id]);
    reinterpret_cast<float2 *>(smem + kSfPrmOff)[threadIdx.x] =
        make_float2(__fmul_rn(a.bn_a[threadIdx.x], inv), __fmul_rn(a.bn_b[threadIdx.x], inv));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kSfBW) {
    // ================================ builders ================================================
    const int t = threadIdx.x;                       // 0..511
    const int q = t & 127, part = t >> 7;            // output column, which quarter of the 21 chunks
    const int lt = t & 255, lrow = t >> 8;           // loader role: input column, which of the two new rows
    __half *ring = reinterpret_cast<__half *>(smem + kSfRingOff);
    constexpr int kMaxPer = 3;                       // ring elements per thread per step: 3 channels of one row
    int tn = 0;
    int rc = 0;                                      // conv rows built so far (A buffer parity / phase)
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
      int n, j0, j1, p0, p1;
      unit_rows(a, u, n, j0, j1, p0, p1);
      const long long xn = (long long)n * 3 * a.H * a.W;  // element index of image n
      // rows [lo, hi) of the image (may be outside: zeros) -> ring, through registers
      auto load_rows = [&](int lo, int hi, uint32_t (&v)[kMaxPer], bool first) {
        const int per_row = 3 * a.W;
        const int total = (hi - lo) * per_row;
        if (first) {  // unit start: 7 rows at once, written straight away
          for (int e = t; e < total; e += kSfBuilders) {
            const int ri = e / per_row, rem = e - ri * per_row;
            const int c = rem / a.W, w = rem - c * a.W;
            const int h = lo + ri;
            const float f = (h >= 0 && h < a.H)
                                ? stem_pixel<IN>(ldg_pinned<IN>(a.x, xn + ((long long)c * a.H + h) * a.W + w), a.nmean[c], a.nstd[c])
                                : 0.f;
            ring[(((h + 16) & (kSfRing - 1)) * 3 + c) * kSfRowP + w + 3] = __float2half_rn(f);
          }
          return;
        }
        // steady state: thread = (one of the two new rows, one column), three channels: no divisions
        const int h = lo + lrow;
        const bool ok = lt < a.W && lrow < hi - lo && h >= 0 && h < a.H;
        const long long src = xn + (long long)h * a.W + lt;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = ok ? ldg_pinned<IN>(a.x, src + (long long)c * a.H * a.W) : 0u;
      };
      auto store_rows = [&](int lo, int hi, const uint32_t (&v)[kMaxPer]) {
        const int h = lo + lrow;
        const bool inside = h >= 0 && h < a.H;  // rows outside the image are the conv's zero padding
        if (lt < a.W && lrow < hi - lo) {
          __half *dst = ring + ((h + 16) & (kSfRing - 1)) * 3 * kSfRowP + lt + 3;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            dst[c * kSfRowP] = __float2half_rn(inside ? stem_pixel<IN>(v[c], a.nmean[c], a.nstd[c]) : 0.f);
        }
      };
      uint32_t nxt[kMaxPer];
      named_bar_sync(1, kSfBuilders);  // nobody still builds from the previous unit's rows
      load_rows(2 * p0 - 3, 2 * p0 + 4, nxt, true);
      named_bar_sync(1, kSfBuilders);
      for (int p = p0; p < p1; ++p, ++rc) {
        const int ab = rc & 1;
        if (t == 0) stem_trace(a, 0, tn, 0, rc);
        if (p + 1 < p1 && !(SF_DBG(a) & 16)) load_rows(2 * p + 4, 2 * p + 6, nxt, false);  // in flight during the build
        mbar_wait(aempty_bar(ab), (uint32_t)(((rc >> 1) & 1) ^ 1));
        if (t == 0) stem_trace(a, 0, tn, 1, rc);
        uint8_t *atile = smem + kSfAOff + ab * kSfABytes;
        if (!(SF_DBG(a) & 1)) {
          if (part == 0) stem_build_chunks<0, 6>(ring + 2 * q, atile + q * 128, q & 7, p);
          else if (part == 1) stem_build_chunks<6, 11>(ring + 2 * q, atile + q * 128, q & 7, p);
          else if (part == 2) stem_build_chunks<11, 16>(ring + 2 * q, atile + q * 128, q & 7, p);
          else stem_build_chunks<16, 21>(ring + 2 * q, atile + q * 128, q & 7, p);
        }
        if (t == 0) stem_trace(a, 0, tn, 2, rc);
        if (!(SF_DBG(a) & 8)) fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(afull_bar(ab));
        if (t == 0) stem_trace(a, 0, tn, 3, rc);
        if (p + 1 < p1) {
          if (!(SF_DBG(a) & 16)) store_rows(2 * p + 4, 2 * p + 6, nxt);
          named_bar_sync(1, kSfBuilders);  // the next row's inputs are complete
        }
        if (t == 0) stem_trace(a, 0, tn, 8, rc);
      }
    }
  } else if (warp >= kSfBW + 8) {
    // ================================ MMA issuers (convergent warps, elected lane) =============
    const int my_par = warp - (kSfBW + 8);  // this warp issues the rows with rc % 2 == my_par
    int tn = 0;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int rc = 0;
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
      int n, j0, j1, p0, p1;
      unit_rows(a, u, n, j0, j1, p0, p1);
      for (int p = p0; p < p1; ++p, ++rc) {
        const int ab = rc & 1;
        if (ab != my_par) continue;
        const uint32_t ph = (uint32_t)((rc >> 1) & 1);
        mbar_wait(tempty_bar(ab), ph ^ 1);
        mbar_wait(afull_bar(ab), ph);
        tc_fence_after();
        if (lane == 0) stem_trace(a, 16 + my_par, tn, 4, rc);
        const uint32_t tmem_d = tmem_u + ab * 64;
        const uint64_t da = make_smem_desc<128>(smem_base + kSfAOff + ab * kSfABytes);
        const uint64_t db = make_smem_desc<128>(smem_base + kSfBOff);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 3; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)  // UMMA_K = 16 halfs = 32 bytes
              umma_f16(tmem_d, da + (uint64_t)(kb * (16384 >> 4) + 2 * k), db + (uint64_t)(kb * (8192 >> 4) + 2 * k),
                       idesc, (kb | k) != 0);
          umma_commit(aempty_bar(ab));
          umma_commit(tfull_bar(ab));
        }
        __syncwarp();
        if (lane == 0) stem_trace(a, 16 + my_par, tn, 9, rc);
      }
    }
  } else {
    // ================================ epilogue + pooling (warps 8-15) ==========================
    const int et = threadIdx.x - kSfBuilders;        // 0..255
    const int wq = warp & 3, half = (warp - kSfBW) >> 2; // TMEM lane quarter, channel half
    const int q = wq * 32 + lane;
    const float2 *prm = reinterpret_cast<const float2 *>(smem + kSfPrmOff);
    uint8_t *cring = smem + kSfConvOff;
    const bool f32_out = a.out_mode == SLQ_OUT_F32;
    int tn = 0;
    int rc = 0;
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x) {
      int n, j0, j1, p0, p1;
      unit_rows(a, u, n, j0, j1, p0, p1);
      for (int p = p0; p < p1; ++p, ++rc) {
        const int ab = rc & 1;
        mbar_wait(tfull_bar(ab), (uint32_t)((rc >> 1) & 1));
        tc_fence_after();
        if (et == 0) stem_trace(a, 18, tn, 5, rc);
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + ab * 64 + half * 32, acc);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tempty_bar(ab));  // accumulator is in registers: the next row may start
        if (SF_DBG(a) & 4) continue;
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 p2 = prm[half * 32 + j];
          y[j] = __fmaf_rn(__uint_as_float(acc[j]), p2.x, p2.y);  // u8: in units of the output scale
        }
        if (f32_out) {
          if (q < a.Wc) {
            float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(a.out) +
                                                   (((long long)n * a.Hc + p) * a.Wc + q) * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(fmaxf(y[4 * j], 0.f), fmaxf(y[4 * j + 1], 0.f), fmaxf(y[4 * j + 2], 0.f),
                                 fmaxf(y[4 * j + 3], 0.f));
          }
          continue;
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)  // saturation at 0 is the ReLU
          pk[j] = epi_pack4<false>(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        uint4 *dst = reinterpret_cast<uint4 *>(cring + (p & (kSfConvRing - 1)) * kSfConvRowBytes + q * 64 + half * 32);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (et == 0) stem_trace(a, 18, tn, 10, rc);
        named_bar_sync(2, kSfEpi);  // conv row p is complete in the ring
        if (et == 0) stem_trace(a, 18, tn, 11, rc);
        // pooled row j = max over conv rows 2j-1..2j+1: complete after an odd row or the last row
        if (!((p & 1) || p == a.Hc - 1) || (SF_DBG(a) & 2)) continue;
        const int j = p >> 1;
        if (j < j0) continue;  // the seam row only feeds this unit's first pooled row
        // one thread per (pooled column, 16 channels): 9 clamped 16-byte loads (a duplicated row or
        // column does not change a max), byte-wise max, one 16-byte coalesced store
        const int r0 = max(2 * j - 1, 0), r1 = 2 * j, r2 = min(2 * j + 1, a.Hc - 1);
        const uint8_t *row0 = cring + (r0 & (kSfConvRing - 1)) * kSfConvRowBytes;
        const uint8_t *row1 = cring + (r1 & (kSfConvRing - 1)) * kSfConvRowBytes;
        const uint8_t *row2 = cring + (r2 & (kSfConvRing - 1)) * kSfConvRowBytes;
        uint4 *orow = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(a.out) +
                                                (((long long)n * a.Hp + j) * a.Wp) * 64);
        uint32_t *rsrow = a.out_rowsum ? a.out_rowsum + ((long long)n * a.Hp + j) * a.Wp : nullptr;
        for (int base = et & ~31; base < a.Wp * 4; base += kSfEpi) {  // whole warps walk the loop (shuffles below)
          const int idx = base + lane;
          const bool live = idx < a.Wp * 4;
          const int i = idx >> 2, g = (idx & 3) * 16;
          uint4 m = make_uint4(0, 0, 0, 0);
          if (live) {
            const int c0 = max(2 * i - 1, 0) * 64 + g, c1 = 2 * i * 64 + g, c2 = min(2 * i + 1, a.Wc - 1) * 64 + g;
            m = *reinterpret_cast<const uint4 *>(row0 + c0);
            auto mx = [&](const uint8_t *ptr) {
              const uint4 v = *reinterpret_cast<const uint4 *>(ptr);
              m.x = __vmaxu4(m.x, v.x); m.y = __vmaxu4(m.y, v.y); m.z = __vmaxu4(m.z, v.z); m.w = __vmaxu4(m.w, v.w);
            };
            mx(row0 + c1); mx(row0 + c2);
            mx(row1 + c0); mx(row1 + c1); mx(row1 + c2);
            mx(row2 + c0); mx(row2 + c1); mx(row2 + c2);
            orow[idx] = m;
          }
          // channel sum of the pooled pixel: 16 channels per thread, four neighbouring lanes per pixel
          uint32_t ps = __dp4a(m.x, 0x01010101u, __dp4a(m.y, 0x01010101u, __dp4a(m.z, 0x01010101u, __dp4a(m.w, 0x01010101u, 0u))));
          ps += __shfl_xor_sync(0xffffffffu, ps, 1);
          ps += __shfl_xor_sync(0xffffffffu, ps, 2);
          if (live && rsrow && (idx & 3) == 0) rsrow[i] = ps;
        }
        if (et == 0) stem_trace(a, 18, tn, 6, rc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSfBW + 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kSfTmemCols) : "memory");
  }
}

}  // namespace slq

using namespace slq;

static void stem_dims(int H, int W, int *Hc, int *Wc, int *Hp, int *Wp) {
  *Hc = (H + 6 - 7) / 2 + 1;
  *Wc = (W + 6 - 7) / 2 + 1;
  *Hp = (*Hc + 2 - 3) / 2 + 1;
  *Wp = (*Wc + 2 - 3) / 2 + 1;
}

extern "C" int64_t slq_stem_workspace_bytes(int32_t N, int32_t H, int32_t W) {
  if (N <= 0 || H < 7 || W < 7) return -1;
  return 64 * kSfK * 2;  // the fp16 weight matrix; activations never leave the SM
}

extern "C" int slq_stem_create(int32_t N, int32_t H, int32_t W, void *workspace, slq_stem **out) {
  SLQ_CHECK_ARG(workspace && out, "slq_stem_create: null pointer argument");
  SLQ_CHECK_ARG(N > 0 && H >= 7 && W >= 7, "slq_stem_create: bad shape");
  SLQ_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "slq_stem_create: workspace must be 256-byte aligned");
  slq_stem *s = new (std::nothrow) slq_stem();
  SLQ_CHECK_ARG(s != nullptr, "slq_stem_create: out of host memory");
  s->N = N; s->H = H; s->W = W;
  stem_dims(H, W, &s->Hc, &s->Wc, &s->Hp, &s->Wp);
  if (s->Wc > 128 || 3 * W > 3 * 256) {
    delete s;
    set_error("slq_stem_create: the tcgen05 stem maps one output row to one 128-pixel tile (W <= 256); use slq_stem_forward");
    return SLQ_ERR_UNSUPPORTED;
  }
  s->wh = reinterpret_cast<__half *>(workspace);
  const int units_per_img = (s->Hp + kSfUnitRows - 1) / kSfUnitRows;
  s->num_ctas = (int)std::min<long long>((long long)N * units_per_img, sm_count());
  static bool attr_done[kMaxDevices] = {false};
  const int cur_dev = current_device();
  if (cur_dev >= kMaxDevices || !attr_done[cur_dev]) {
    cudaError_t e = cudaFuncSetAttribute(stem_fused_kernel<SLQ_IN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSfSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem_fused_kernel<SLQ_IN_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSfSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem_fused_kernel<SLQ_IN_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSfSmemBytes);
    if (e != cudaSuccess) {
      delete s;
      set_error("slq_stem_create: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return SLQ_ERR_CUDA;
    }
    if (cur_dev < kMaxDevices) attr_done[cur_dev] = true;
  }
  *out = s;
  return SLQ_OK;
}

extern "C" void slq_stem_destroy(slq_stem *s) { delete s; }

extern "C" int slq_stem_set_weights(slq_stem *s, const float *w, void *stream) {
  SLQ_CHECK_ARG(s && w, "slq_stem_set_weights: null pointer argument");
  stem_weights_kernel<<<(64 * kSfK + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, s->wh);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_stem_launch(slq_stem *s, const float *x, const float *bn_a, const float *bn_b,
                               const float *act_scales, int32_t out_id, void *out, int32_t out_mode,
                               float *f32_scratch, uint32_t *out_rowsum, void *stream) {
  return slq_stem_launch_in(s, x, SLQ_IN_F32, nullptr, bn_a, bn_b, act_scales, out_id, out, out_mode, f32_scratch,
                            out_rowsum, stream);
}

extern "C" int slq_stem_launch_in(slq_stem *s, const void *x, int32_t in_kind, const float *norm, const float *bn_a,
                                  const float *bn_b, const float *act_scales, int32_t out_id, void *out,
                                  int32_t out_mode, float *f32_scratch, uint32_t *out_rowsum, void *stream) {
  SLQ_CHECK_ARG(s && x && bn_a && bn_b && out, "slq_stem_launch: null pointer argument");
  SLQ_CHECK_ARG(in_kind == SLQ_IN_F32 || in_kind == SLQ_IN_F16 || in_kind == SLQ_IN_U8, "slq_stem_launch: in_kind %d", in_kind);
  SLQ_CHECK_ARG(in_kind != SLQ_IN_U8 || norm != nullptr, "slq_stem_launch: u8 input needs norm = {mean[3], std[3]}");
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_U8 || out_mode == SLQ_OUT_F32, "slq_stem_launch: out_mode %d", out_mode);
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_U8 ? act_scales != nullptr : f32_scratch != nullptr,
                "slq_stem_launch: act_scales (u8) / f32_scratch (fp32) required");
  cudaStream_t st = (cudaStream_t)stream;
  StemArgs a;
  a.x = x;
  for (int c = 0; c < 3; ++c) {
    a.nmean[c] = in_kind == SLQ_IN_U8 ? norm[c] : 0.f;
    a.nstd[c] = in_kind == SLQ_IN_U8 ? norm[3 + c] : 1.f;
  }
  a.N = s->N; a.H = s->H; a.W = s->W; a.Hc = s->Hc; a.Wc = s->Wc; a.Hp = s->Hp; a.Wp = s->Wp;
  a.wh = s->wh;
  a.bn_a = bn_a; a.bn_b = bn_b; a.act_scales = act_scales; a.out_id = out_id; a.out_mode = out_mode;
  a.out = out_mode == SLQ_OUT_U8 ? out : (void *)f32_scratch;
  a.out_rowsum = out_mode == SLQ_OUT_U8 ? out_rowsum : nullptr;
  a.units_per_img = (s->Hp + kSfUnitRows - 1) / kSfUnitRows;
  a.total_units = s->N * a.units_per_img;
  {
    int cap = 0;
    debug_trace_buffer(&a.trace, &cap);
    a.trace_cap = cap;
#if SLQ_DEBUG_TRACE
    const char *d = getenv("SLQ_STEM_DBG");
    a.dbg = d ? atoi(d) : 0;
#else
    a.dbg = 0;
#endif
  }
  if (in_kind == SLQ_IN_F32) stem_fused_kernel<SLQ_IN_F32><<<s->num_ctas, kSfThreads, kSfSmemBytes, st>>>(a);
  else if (in_kind == SLQ_IN_F16) stem_fused_kernel<SLQ_IN_F16><<<s->num_ctas, kSfThreads, kSfSmemBytes, st>>>(a);
  else stem_fused_kernel<SLQ_IN_U8><<<s->num_ctas, kSfThreads, kSfSmemBytes, st>>>(a);
  SLQ_LAUNCH_CHECK();
  if (out_mode == SLQ_OUT_F32)
    return launch_stem_pool(f32_scratch, s->N, s->Hc, s->Wc, s->Hp, s->Wp, act_scales, out_id, out, SLQ_OUT_F32, nullptr, st);
  return SLQ_OK;
}
