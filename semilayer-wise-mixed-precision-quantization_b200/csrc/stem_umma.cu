// csrc/stem_umma.cu -- the un-quantised stem on Blackwell tensor cores.
//
// Replaces resnet.py:206-209 (conv1 7x7 s2 p3, 3->64, fp32 weights; bn1; relu; maxpool 3x3 s2 p1).
// The reference keeps these weights in fp32, so this layer is NOT quantised to 8 bit:
// operands go through the tensor core as fp16 (11-bit significand, fp32 accumulation in TMEM),
// a relative error ~3e-4, an order of magnitude below the u8 activation quantisation that follows.
//
// Data path (3 launches):
//   1. stem_prep_kernel   x fp32 NCHW -> xr fp16 [N, Hr, Wc, 32]: for every input row hh and output
//      column q the 7 taps x 4 channels (3 + zero pad) window = 32 contiguous halfs.  An implicit
//      GEMM over a 3-channel NHWC image cannot be fed by TMA directly (6-byte pixels), this
//      row-wise expansion (2.7x of the fp32 image) is what makes every A tile one contiguous
//      7 KB TMA box.
//   2. stem_umma_kernel   one tile = one output row (n, p): D[q, oc] = sum_r xr[n, 2p+r, q, :] . W[oc, r, :]
//      7 K-blocks of 32 halfs, tcgen05.mma kind::f16 M=128 N=64 K=16, epilogue = folded BN + ReLU
//      (+ u8 quantisation with the POOLED tensor's scale: max-pooling commutes with a monotone map).
//   3. stem_pool_u8_kernel  3x3 s2 max-pool on the u8 NHWC tensor.
#include <algorithm>
#include <cuda_fp16.h>
#include <new>

#include "conv_common.cuh"
#include "umma_ptx.cuh"

struct slq_stem {
  int N, H, W, Hc, Wc, Hp, Wp, Hr;
  __half *xr;        // [N, Hr, Wc, 32]
  uint8_t *conv_u8;  // [N, Hc, Wc, 64]
  __half *wh;        // [64, 7*32]
  CUtensorMap tmA, tmB;
  int num_ctas;
};

namespace slq {

constexpr int kStemStages = 8;
constexpr int kStemABytes = 128 * 64;  // 128 pixel rows x 32 halfs
constexpr int kStemBBytes = 64 * 64;   // 64 output channels x 32 halfs
constexpr int kStemStageBytes = kStemABytes + kStemBBytes;
constexpr int kStemSmemBytes = 1024 + kStemStages * kStemStageBytes + 64 * 8 + 256;
constexpr int kStemTmemCols = 128;  // 2 accumulator buffers x 64 columns

__global__ void __launch_bounds__(256) stem_prep_kernel(const float *__restrict__ x, int N, int H, int W,
                                                        int Hr, int Wc, __half *__restrict__ xr) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (n, hh, q)
  const long long total = (long long)N * Hr * Wc;
  if (idx >= total) return;
  const int q = (int)(idx % Wc);
  const int hh = (int)((idx / Wc) % Hr);
  const int n = (int)(idx / ((long long)Wc * Hr));
  const int h = hh - 3;
  __align__(16) __half v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float2half_rn(0.f);
  if (h >= 0 && h < H) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int w = 2 * q + s - 3;
      if (w >= 0 && w < W) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[s * 4 + c] = __float2half_rn(__ldg(x + (((long long)n * 3 + c) * H + h) * W + w));
      }
    }
  }
  uint4 *dst = reinterpret_cast<uint4 *>(xr + idx * 32);
  const uint4 *src = reinterpret_cast<const uint4 *>(v);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = src[i];
}

// w fp32 [64, 3, 7, 7] -> wh fp16 [64, 7, 32]: wh[oc][r][s*4 + c]
__global__ void stem_weights_kernel(const float *__restrict__ w, __half *__restrict__ wh) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 7 * 32) return;
  const int j = idx & 31, r = (idx >> 5) % 7, oc = idx / (7 * 32);
  const int s = j >> 2, c = j & 3;
  float v = 0.f;
  if (s < 7 && c < 3) v = w[((oc * 3 + c) * 7 + r) * 7 + s];
  wh[idx] = __float2half_rn(v);
}

struct StemArgs {
  int N, Hc, Wc, Hr;
  const float *bn_a, *bn_b, *act_scales;
  int out_id, out_mode;
  void *out;  // u8 [N,Hc,Wc,64] or fp32 [N,Hc,Wc,64]
};

__global__ void __launch_bounds__(256, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const StemArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  float2 *prm = reinterpret_cast<float2 *>(smem + kStemStages * kStemStageBytes);  // {bn_a, bn_b} x 64
  const uint32_t bar_base = smem_base + kStemStages * kStemStageBytes + 64 * 8;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStemStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kStemStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kStemStages + 2 + b); };
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(
      smem + kStemStages * kStemStageBytes + 64 * 8 + 8 * (2 * kStemStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_tiles = (long long)a.N * a.Hc;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStemStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void *)tmem_slot)),
                 "r"(kStemTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) prm[threadIdx.x] = make_float2(a.bn_a[threadIdx.x], a.bn_b[threadIdx.x]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = (int)(tile / a.Hc), p = (int)(tile % a.Hc);
        for (int r = 0; r < 7; ++r) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * kStemStageBytes;
          mbar_expect_tx(full_bar(stage), kStemStageBytes);
          tma_load_3d(sa, &tmA, full_bar(stage), 0, 0, n * a.Hr + 2 * p + r);
          tma_load_2d(sa + kStemABytes, &tmB, full_bar(stage), r * 32, 0);
          if (++stage == kStemStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer: D=f32, A=B=f16, K-major, M=128, N=64
      const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        mbar_wait(tempty_bar(buf), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 64;
        for (int r = 0; r < 7; ++r) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStemStageBytes;
          const uint64_t da = make_smem_desc<64>(sa);
          const uint64_t db = make_smem_desc<64>(sa + kStemABytes);
#pragma unroll
          for (int k = 0; k < 2; ++k)  // UMMA_K = 16 halfs = 32 bytes
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (r | k) != 0);
          umma_commit(empty_bar(stage));
          if (++stage == kStemStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else if (warp >= 4) {  // ---- epilogue: folded BN + ReLU (+ u8 quantisation)
    const int wq = warp & 3;
    const float inv_out = a.out_mode == SLQ_OUT_F32 ? 1.f : __fdiv_rn(1.0f, a.act_scales[a.out_id]);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      mbar_wait(tfull_bar(buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * 64;
      const int q = wq * 32 + lane;
      const bool valid = q < a.Wc;
      const long long pix = tile * a.Wc + q;  // tile == n*Hc + p
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t acc[32];
        tmem_ld32(trow + ch * 32, acc);
        tmem_ld_wait();
        if (!valid) continue;
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 p2 = prm[ch * 32 + j];
          y[j] = fmaxf(__fadd_rn(__fmul_rn(__uint_as_float(acc[j]), p2.x), p2.y), 0.f);
        }
        if (a.out_mode == SLQ_OUT_F32) {
          float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(a.out) + pix * 64 + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        } else {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = epi_quant_u8(__fmul_rn(y[4 * j], inv_out)) | (epi_quant_u8(__fmul_rn(y[4 * j + 1], inv_out)) << 8) |
                    (epi_quant_u8(__fmul_rn(y[4 * j + 2], inv_out)) << 16) | (epi_quant_u8(__fmul_rn(y[4 * j + 3], inv_out)) << 24);
          uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(a.out) + pix * 64 + ch * 32);
          o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kStemTmemCols) : "memory");
  }
}

// 3x3 stride-2 pad-1 max-pool on u8 NHWC (C = 64): one thread per (pixel, 16 channels)
__global__ void __launch_bounds__(256) stem_pool_u8_kernel(const uint8_t *__restrict__ y, int N, int Hc, int Wc,
                                                           int Hp, int Wp, uint8_t *__restrict__ out) {
  const long long total = (long long)N * Hp * Wp * 4;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx & 3);
  const long long pix = idx >> 2;
  const int wp = (int)(pix % Wp), hp = (int)((pix / Wp) % Hp), n = (int)(pix / ((long long)Wp * Hp));
  uint4 m = make_uint4(0, 0, 0, 0);
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * hp - 1 + r;
    if (h < 0 || h >= Hc) continue;
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wp - 1 + s;
      if (w < 0 || w >= Wc) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(y + (((long long)n * Hc + h) * Wc + w) * 64) + g);
      m.x = __vmaxu4(m.x, v.x); m.y = __vmaxu4(m.y, v.y); m.z = __vmaxu4(m.z, v.z); m.w = __vmaxu4(m.w, v.w);
    }
  }
  reinterpret_cast<uint4 *>(out + pix * 64)[g] = m;
}

}  // namespace slq

using namespace slq;

static void stem_dims(int H, int W, int *Hc, int *Wc, int *Hp, int *Wp, int *Hr) {
  *Hc = (H + 6 - 7) / 2 + 1;
  *Wc = (W + 6 - 7) / 2 + 1;
  *Hp = (*Hc + 2 - 3) / 2 + 1;
  *Wp = (*Wc + 2 - 3) / 2 + 1;
  *Hr = 2 * (*Hc - 1) + 7;
}

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" int64_t slq_stem_workspace_bytes(int32_t N, int32_t H, int32_t W) {
  if (N <= 0 || H < 7 || W < 7) return -1;
  int Hc, Wc, Hp, Wp, Hr;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp, &Hr);
  return align256((int64_t)N * Hr * Wc * 32 * 2) + align256((int64_t)N * Hc * Wc * 64) + align256(64 * 7 * 32 * 2);
}

extern "C" int slq_stem_create(int32_t N, int32_t H, int32_t W, void *workspace, slq_stem **out) {
  SLQ_CHECK_ARG(workspace && out, "slq_stem_create: null pointer argument");
  SLQ_CHECK_ARG(N > 0 && H >= 7 && W >= 7, "slq_stem_create: bad shape");
  SLQ_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "slq_stem_create: workspace must be 256-byte aligned");
  slq_stem *s = new (std::nothrow) slq_stem();
  SLQ_CHECK_ARG(s != nullptr, "slq_stem_create: out of host memory");
  s->N = N; s->H = H; s->W = W;
  stem_dims(H, W, &s->Hc, &s->Wc, &s->Hp, &s->Wp, &s->Hr);
  if (s->Wc > 128) {
    delete s;
    set_error("slq_stem_create: the tcgen05 stem maps one output row to one 128-pixel tile (W <= 256); use slq_stem_forward");
    return SLQ_ERR_UNSUPPORTED;
  }
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  s->xr = reinterpret_cast<__half *>(ws);
  ws += align256((int64_t)N * s->Hr * s->Wc * 32 * 2);
  s->conv_u8 = ws;
  ws += align256((int64_t)N * s->Hc * s->Wc * 64);
  s->wh = reinterpret_cast<__half *>(ws);
  // tensor maps
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
  if (ce != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) {
    delete s;
    set_error("slq_stem_create: cuTensorMapEncodeTiled unavailable");
    return SLQ_ERR_CUDA;
  }
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  {
    cuuint64_t dims[3] = {32, (cuuint64_t)s->Wc, (cuuint64_t)N * s->Hr};
    cuuint64_t strides[2] = {64, (cuuint64_t)s->Wc * 64};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&s->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, s->xr, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete s;
      set_error("slq_stem_create: cuTensorMapEncodeTiled(A) failed: CUresult %d", (int)r);
      return SLQ_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {7 * 32, 64};
    cuuint64_t strides[1] = {7 * 32 * 2};
    cuuint32_t box[2] = {32, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&s->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, s->wh, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete s;
      set_error("slq_stem_create: cuTensorMapEncodeTiled(B) failed: CUresult %d", (int)r);
      return SLQ_ERR_CUDA;
    }
  }
  s->num_ctas = (int)std::min<long long>((long long)N * s->Hc, sm_count());
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmemBytes);
    if (e != cudaSuccess) {
      delete s;
      set_error("slq_stem_create: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return SLQ_ERR_CUDA;
    }
    attr_done = true;
  }
  *out = s;
  return SLQ_OK;
}

extern "C" void slq_stem_destroy(slq_stem *s) { delete s; }

extern "C" int slq_stem_set_weights(slq_stem *s, const float *w, void *stream) {
  SLQ_CHECK_ARG(s && w, "slq_stem_set_weights: null pointer argument");
  stem_weights_kernel<<<(64 * 7 * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, s->wh);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_stem_launch(slq_stem *s, const float *x, const float *bn_a, const float *bn_b,
                               const float *act_scales, int32_t out_id, void *out, int32_t out_mode,
                               float *f32_scratch, void *stream) {
  SLQ_CHECK_ARG(s && x && bn_a && bn_b && out, "slq_stem_launch: null pointer argument");
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_U8 || out_mode == SLQ_OUT_F32, "slq_stem_launch: out_mode %d", out_mode);
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_U8 ? act_scales != nullptr : f32_scratch != nullptr,
                "slq_stem_launch: act_scales (u8) / f32_scratch (fp32) required");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nprep = (long long)s->N * s->Hr * s->Wc;
  stem_prep_kernel<<<(unsigned)ceil_div(nprep, 256), 256, 0, st>>>(x, s->N, s->H, s->W, s->Hr, s->Wc, s->xr);
  SLQ_LAUNCH_CHECK();
  StemArgs a;
  a.N = s->N; a.Hc = s->Hc; a.Wc = s->Wc; a.Hr = s->Hr;
  a.bn_a = bn_a; a.bn_b = bn_b; a.act_scales = act_scales; a.out_id = out_id; a.out_mode = out_mode;
  a.out = out_mode == SLQ_OUT_U8 ? (void *)s->conv_u8 : (void *)f32_scratch;
  stem_umma_kernel<<<s->num_ctas, 256, kStemSmemBytes, st>>>(s->tmA, s->tmB, a);
  SLQ_LAUNCH_CHECK();
  if (out_mode == SLQ_OUT_U8) {
    const long long total = (long long)s->N * s->Hp * s->Wp * 4;
    stem_pool_u8_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(s->conv_u8, s->N, s->Hc, s->Wc, s->Hp, s->Wp,
                                                                       reinterpret_cast<uint8_t *>(out));
  } else {
    return launch_stem_pool(f32_scratch, s->N, s->Hc, s->Wc, s->Hp, s->Wp, act_scales, out_id, out, SLQ_OUT_F32, st);
  }
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}
