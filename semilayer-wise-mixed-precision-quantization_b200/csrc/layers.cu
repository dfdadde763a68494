// csrc/layers.cu -- everything around the tcgen05 conv kernel:
//   * GEMM-ready weight builder (packed 2/4/8/16-bit rows -> u8 limbs, K reordered to (r,s,c))
//   * SIMT (dp4a) convolution with the same fused epilogue: the on-device cross-check of the
//     tcgen05 kernel (tests) -- NOT the product path
//   * stem   (resnet.py:206-209: conv 7x7 s2 + bn + relu + maxpool 3x3 s2), fp32 weights
//   * activation-scale calibration helpers (abs-max -> scale, fp32 -> u8/s8)
#include "conv_common.cuh"

namespace slq {

int validate_desc(const slq_conv_desc *d) {
  SLQ_CHECK_ARG(d != nullptr, "conv desc is NULL");
  SLQ_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "conv desc: N/H/W must be positive");
  SLQ_CHECK_ARG(d->Cin > 0 && d->Cin % 64 == 0, "conv desc: Cin=%d must be a multiple of 64", d->Cin);
  SLQ_CHECK_ARG(d->Cout > 0 && d->Cout % 16 == 0, "conv desc: Cout=%d must be a multiple of 16", d->Cout);
  SLQ_CHECK_ARG((d->kh == 1 && d->kw == 1) || (d->kh == 3 && d->kw == 3), "conv desc: only 1x1 and 3x3 kernels");
  SLQ_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv desc: stride %d", d->stride);
  SLQ_CHECK_ARG(d->pad == d->kh / 2, "conv desc: pad must be kh/2");
  SLQ_CHECK_ARG(d->H + 2 * d->pad >= d->kh && d->W + 2 * d->pad >= d->kw, "conv desc: input smaller than kernel");
  return SLQ_OK;
}

// ------------------------------------------------------------------------------------------
// GEMM-ready weights
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int unpack_code(const uint8_t *row, int64_t e, int64_t K, int bit) {
  if (bit == 4) return (row[e >> 1] >> ((e & 1) * 4)) & 15;
  if (bit == 2) return (row[e >> 2] >> ((e & 3) * 2)) & 3;
  if (bit == 16) return (int)row[e] | ((int)row[K + e] << 8);
  return row[e];
}

__global__ void build_gemm_weights_kernel(ConvGeom g, const uint8_t *__restrict__ codes,
                                          const int64_t *__restrict__ code_offsets,
                                          const int32_t *__restrict__ bit,
                                          uint8_t *__restrict__ wg) {
  // one thread per 4 consecutive k (same tap, 4 channels): one u32 store
  const int64_t kq = g.Ktot / 4;
  const int64_t total = (int64_t)g.gemm_rows * kq;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / kq);
    const int k0 = (int)(idx % kq) * 4;
    int oc, limb;
    if (g.w16) {
      oc = (row >> 7) * 64 + (row & 63);
      limb = (row >> 6) & 1;
    } else {
      oc = row;
      limb = 0;
    }
    uint32_t word = 0;
    if (oc < g.Cout) {
      const int b = bit[oc];
      const uint8_t *src = codes + code_offsets[oc];
      const int tap = k0 / g.Cin, c0 = k0 % g.Cin;  // tap = r*kw + s
      const int taps = g.kh * g.kw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t e = (int64_t)(c0 + j) * taps + tap;  // OIHW element (c, r, s)
        const int code = unpack_code(src, e, g.Ktot, b);
        word |= (uint32_t)((limb ? (code >> 8) : (code & 255)) & 255) << (8 * j);
      }
    }
    *reinterpret_cast<uint32_t *>(wg + (int64_t)row * g.Ktot + k0) = word;
  }
}

// GEMM-ready weights in PACKED form (slq_build_packed_gemm_weights): per n-tile and K block (k_block codes of
// K in (r, s, c) order) the rows' codes back to back -- rows of <= 4 bits two codes per byte (low nibble first),
// wider rows one per byte.  One thread per byte of the result.
__global__ void build_packed_gemm_weights_kernel(ConvGeom g, int swz, const uint8_t *__restrict__ codes,
                                                 const int64_t *__restrict__ code_offsets,
                                                 const int32_t *__restrict__ bit, const int64_t *__restrict__ tile_base,
                                                 const int32_t *__restrict__ seg_bytes,
                                                 const uint16_t *__restrict__ row_offsets, uint8_t *__restrict__ wgp) {
  const int num_kb = g.Ktot / swz;
  const int row = blockIdx.x;  // GEMM row == output channel (one-limb layers only)
  const int t = row / g.bn_cols, rl = row - t * g.bn_cols;
  const uint16_t *ro = row_offsets + (long long)t * (g.bn_cols + 1);
  const int o0 = ro[rl], len = ro[rl + 1] - o0;  // bytes of this row per K block: swz/2 or swz
  const bool half = len == swz / 2;
  const int b = row < g.Cout ? bit[row] : 4;
  const uint8_t *src = row < g.Cout ? codes + code_offsets[row] : nullptr;
  const int taps = g.kh * g.kw;
  for (int i = threadIdx.x; i < num_kb * len; i += blockDim.x) {
    const int kb = i / len, j = i - kb * len;
    int v = 0;
    if (src) {
      const int per = half ? 2 : 1;
      for (int e = 0; e < per; ++e) {
        const int k = kb * swz + j * per + e;             // GEMM K index, (r, s, c) order
        const int tap = k / g.Cin, c = k - tap * g.Cin;
        const int code = unpack_code(src, (int64_t)c * taps + tap, g.Ktot, b);  // OIHW element (c, r, s)
        v |= (code & (half ? 15 : 255)) << (4 * e);
      }
    }
    wgp[tile_base[t] + (long long)kb * seg_bytes[t] + o0 + j] = (uint8_t)v;
  }
}

// ------------------------------------------------------------------------------------------
// SIMT convolution (cross-check): thread = (output pixel, output channel)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_simt_kernel(ConvGeom g, const uint8_t *__restrict__ in,
                                                        const uint8_t *__restrict__ wg, EpiDev e) {
  const int oc = blockIdx.y * 32 + (threadIdx.x & 31);
  const long long m = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= g.M || oc >= g.Cout) return;
  const int wo = (int)(m % g.Wo);
  const int ho = (int)((m / g.Wo) % g.Ho);
  const int n = (int)(m / ((long long)g.Wo * g.Ho));
  const uint8_t *wlo = wg + (long long)gemm_row_of(oc, 0, g.w16) * g.Ktot;
  const uint8_t *whi = wg + (long long)gemm_row_of(oc, 1, g.w16) * g.Ktot;
  unsigned acc_lo = 0, acc_hi = 0, S = 0;
  for (int r = 0; r < g.kh; ++r) {
    const int hi = ho * g.stride - g.pad + r;
    if (hi < 0 || hi >= g.H) continue;
    for (int s = 0; s < g.kw; ++s) {
      const int wi = wo * g.stride - g.pad + s;
      if (wi < 0 || wi >= g.W) continue;
      const uint32_t *xp = reinterpret_cast<const uint32_t *>(in + (((long long)n * g.H + hi) * g.W + wi) * g.Cin);
      const int kbase = (r * g.kw + s) * g.Cin;
      const uint32_t *wl = reinterpret_cast<const uint32_t *>(wlo + kbase);
      const uint32_t *wh = reinterpret_cast<const uint32_t *>(whi + kbase);
      for (int c4 = 0; c4 < g.Cin / 4; ++c4) {
        const uint32_t xv = __ldg(xp + c4);
        acc_lo = __dp4a(xv, __ldg(wl + c4), acc_lo);
        if (g.w16) acc_hi = __dp4a(xv, __ldg(wh + c4), acc_hi);
        S = __dp4a(xv, 0x01010101u, S);
      }
    }
  }
  if (e.out_mode == SLQ_OUT_ACC) {
    int32_t *o = reinterpret_cast<int32_t *>(e.out);
    const long long ld = (long long)(g.w16 ? 2 : 1) * g.Cout;
    o[m * ld + oc] = (int)acc_lo;
    if (g.w16) o[m * ld + g.Cout + oc] = (int)acc_hi;
    if (e.out_S && oc == 0) e.out_S[m] = (int)S;
    return;
  }
  const float s_in = e.act_scales[e.in_id];
  const bool quantised = e.out_mode != SLQ_OUT_F32;
  const float inv = quantised ? __fdiv_rn(1.0f, e.act_scales[e.out_id]) : 1.0f;
  const ChanParam cp = make_chan_param(e.wscale[oc], e.zf[oc], e.bias[oc], s_in, inv, quantised);
  const float Sf = (float)(int)S;
  float y = g.w16 ? epi_value<true>((int)acc_lo, (int)acc_hi, Sf, cp) : epi_value<false>((int)acc_lo, 0, Sf, cp);
  if (e.res != nullptr) {
    float sr = e.act_scales[e.res_id];
    if (quantised) sr = __fmul_rn(sr, inv);
    y = epi_add_res(y, e.res[m * g.Cout + oc], e.res_signed != 0, sr);
  }
  if (e.out_mode == SLQ_OUT_F32) {
    reinterpret_cast<float *>(e.out)[m * g.Cout + oc] = e.relu ? fmaxf(y, 0.f) : y;
  } else {
    const uint32_t q = e.out_mode == SLQ_OUT_S8 ? epi_quant_s8(y) : epi_quant_u8(y);
    reinterpret_cast<uint8_t *>(e.out)[m * g.Cout + oc] = (uint8_t)q;
  }
}

int launch_conv_simt(const ConvGeom &g, const uint8_t *in, const uint8_t *wg, const EpiDev &e,
                     cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(g.M, 4), (unsigned)ceil_div(g.Cout, 32));
  conv_simt_kernel<<<grid, 128, 0, st>>>(g, in, wg, e);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

// ------------------------------------------------------------------------------------------
// Stem: conv 7x7 s2 p3 (3 -> 64) + folded BN + ReLU  (fp32, CUDA cores), then maxpool 3x3 s2 p1
// ------------------------------------------------------------------------------------------
constexpr int kStemTile = 8;                      // conv-output tile edge per CTA
constexpr int kStemPatch = 2 * kStemTile + 5;     // input rows/cols needed: 21

__global__ void __launch_bounds__(256) stem_conv_kernel(const float *__restrict__ x, int N, int H, int W,
                                                        int Hc, int Wc, const float *__restrict__ w,
                                                        const float *__restrict__ bn_a,
                                                        const float *__restrict__ bn_b,
                                                        float *__restrict__ y) {
  __shared__ __align__(16) float wS[147][64];                 // [c*49 + r*7 + s][oc]
  __shared__ float patch[3][kStemPatch][kStemPatch + 1];
  const int n = blockIdx.z;
  const int ho0 = blockIdx.y * kStemTile, wo0 = blockIdx.x * kStemTile;
  for (int i = threadIdx.x; i < 147 * 64; i += 256) {
    const int oc = i / 147, k = i % 147;
    wS[k][oc] = w[i];
  }
  const int hi0 = 2 * ho0 - 3, wi0 = 2 * wo0 - 3;
  for (int i = threadIdx.x; i < 3 * kStemPatch * kStemPatch; i += 256) {
    const int c = i / (kStemPatch * kStemPatch);
    const int rr = (i / kStemPatch) % kStemPatch, cc = i % kStemPatch;
    const int hi = hi0 + rr, wi = wi0 + cc;
    float v = 0.f;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = x[(((long long)n * 3 + c) * H + hi) * W + wi];
    patch[c][rr][cc] = v;
  }
  __syncthreads();
  const int og = threadIdx.x & 15;  // 4 output channels og*4..
  const int pg = threadIdx.x >> 4;  // 16 groups of 4 pixels
  const int prow = pg >> 1, pcol0 = (pg & 1) * 4;
  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 7; ++r)
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float4 w4 = *reinterpret_cast<const float4 *>(&wS[c * 49 + r * 7 + s][og * 4]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = patch[c][2 * prow + r][2 * (pcol0 + p) + s];
          acc[p][0] = fmaf(xv, w4.x, acc[p][0]);
          acc[p][1] = fmaf(xv, w4.y, acc[p][1]);
          acc[p][2] = fmaf(xv, w4.z, acc[p][2]);
          acc[p][3] = fmaf(xv, w4.w, acc[p][3]);
        }
      }
  const float4 a4 = *reinterpret_cast<const float4 *>(bn_a + og * 4);
  const float4 b4 = *reinterpret_cast<const float4 *>(bn_b + og * 4);
  const int ho = ho0 + prow;
  if (ho >= Hc) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int wo = wo0 + pcol0 + p;
    if (wo >= Wc) continue;
    float4 o;
    o.x = fmaxf(__fadd_rn(__fmul_rn(acc[p][0], a4.x), b4.x), 0.f);
    o.y = fmaxf(__fadd_rn(__fmul_rn(acc[p][1], a4.y), b4.y), 0.f);
    o.z = fmaxf(__fadd_rn(__fmul_rn(acc[p][2], a4.z), b4.z), 0.f);
    o.w = fmaxf(__fadd_rn(__fmul_rn(acc[p][3], a4.w), b4.w), 0.f);
    *reinterpret_cast<float4 *>(y + ((((long long)n * Hc + ho) * Wc + wo) * 64 + og * 4)) = o;
  }
}

__global__ void __launch_bounds__(256) stem_pool_kernel(const float *__restrict__ y, int N, int Hc, int Wc,
                                                        int Hp, int Wp, const float *__restrict__ act_scales,
                                                        int out_id, void *__restrict__ out, int out_mode,
                                                        uint32_t *__restrict__ out_rowsum) {
  const long long total = (long long)N * Hp * Wp * 16;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int og = (int)(idx & 15);
  const long long pix = idx >> 4;
  const int wp = (int)(pix % Wp), hp = (int)((pix / Wp) % Hp), n = (int)(pix / ((long long)Wp * Hp));
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);  // inputs are post-ReLU (>= 0) and padding never wins
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * hp - 1 + r;
    if (h < 0 || h >= Hc) continue;
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wp - 1 + s;
      if (w < 0 || w >= Wc) continue;
      const float4 v = *reinterpret_cast<const float4 *>(y + ((((long long)n * Hc + h) * Wc + w) * 64 + og * 4));
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  if (out_mode == SLQ_OUT_F32) {
    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + pix * 64 + og * 4) = m;
  } else {
    const float inv = __fdiv_rn(1.0f, act_scales[out_id]);
    const uint32_t q = epi_quant_u8(__fmul_rn(m.x, inv)) | (epi_quant_u8(__fmul_rn(m.y, inv)) << 8) |
                       (epi_quant_u8(__fmul_rn(m.z, inv)) << 16) | (epi_quant_u8(__fmul_rn(m.w, inv)) << 24);
    *reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(out) + pix * 64 + og * 4) = q;
    if (out_rowsum) {  // the 16 threads of a pixel are 16 neighbouring lanes (total is a multiple of 16: no lane is missing)
      uint32_t ps = __dp4a(q, 0x01010101u, 0u);
      const unsigned live = __activemask();  // threads past the end left in whole groups of 16
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(live, ps, o, 16);
      if (og == 0) out_rowsum[pix] = ps;
    }
  }
}

int launch_stem_pool(const float *y, int N, int Hc, int Wc, int Hp, int Wp, const float *act_scales,
                     int out_id, void *out, int out_mode, uint32_t *out_rowsum, cudaStream_t st) {
  const long long total = (long long)N * Hp * Wp * 16;
  stem_pool_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(y, N, Hc, Wc, Hp, Wp, act_scales, out_id, out, out_mode,
                                                                  out_mode == SLQ_OUT_U8 ? out_rowsum : nullptr);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

// ------------------------------------------------------------------------------------------
// Calibration helpers
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ y, long long n,
                                                     uint32_t *__restrict__ tmp) {
  float m = 0.f;
  const long long n4 = n >> 2;
  const float4 *y4 = reinterpret_cast<const float4 *>(y);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(y4 + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(y[i]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(tmp, __float_as_uint(m));
}

__global__ void absmax_finalize_kernel(const uint32_t *tmp, float *act_scales, int id, float qmax) {
  const float m = __uint_as_float(*tmp);
  act_scales[id] = (m == 0.f) ? 1.0f : __fdiv_rn(m, qmax);
}

__global__ void __launch_bounds__(256) quantize_act_kernel(const float *__restrict__ y, long long n,
                                                           const float *__restrict__ act_scales, int id,
                                                           int is_signed, uint8_t *__restrict__ out) {
  const float inv = __fdiv_rn(1.0f, act_scales[id]);
  const long long n4 = n >> 2;
  const float4 *y4 = reinterpret_cast<const float4 *>(y);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(y4 + i);
    uint32_t q;
    if (is_signed)
      q = epi_quant_s8(__fmul_rn(v.x, inv)) | (epi_quant_s8(__fmul_rn(v.y, inv)) << 8) | (epi_quant_s8(__fmul_rn(v.z, inv)) << 16) |
          (epi_quant_s8(__fmul_rn(v.w, inv)) << 24);
    else
      q = epi_quant_u8(__fmul_rn(v.x, inv)) | (epi_quant_u8(__fmul_rn(v.y, inv)) << 8) | (epi_quant_u8(__fmul_rn(v.z, inv)) << 16) |
          (epi_quant_u8(__fmul_rn(v.w, inv)) << 24);
    reinterpret_cast<uint32_t *>(out)[i] = q;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
      out[i] = (uint8_t)(is_signed ? epi_quant_s8(__fmul_rn(y[i], inv)) : epi_quant_u8(__fmul_rn(y[i], inv)));
}

}  // namespace slq

using namespace slq;

extern "C" int64_t slq_gemm_weight_rows(const slq_conv_desc *d) {
  if (!d) return -1;
  return make_geom(*d).gemm_rows;
}

extern "C" int slq_build_gemm_weights(const slq_conv_desc *d, const uint8_t *codes,
                                      const int64_t *code_offsets, const int32_t *bit, uint8_t *wg,
                                      void *stream) {
  int rc = validate_desc(d);
  if (rc != SLQ_OK) return rc;
  SLQ_CHECK_ARG(codes && code_offsets && bit && wg, "slq_build_gemm_weights: null pointer argument");
  const ConvGeom g = make_geom(*d);
  const int64_t total = (int64_t)g.gemm_rows * (g.Ktot / 4);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  build_gemm_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, codes, code_offsets, bit, wg);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_stem_forward(const float *x, int32_t N, int32_t H, int32_t W, const float *w,
                                const float *bn_a, const float *bn_b, const float *act_scales,
                                int32_t out_id, float *scratch, void *out, int32_t out_mode,
                                uint32_t *out_rowsum, void *stream) {
  SLQ_CHECK_ARG(x && w && bn_a && bn_b && scratch && out, "slq_stem_forward: null pointer argument");
  SLQ_CHECK_ARG(N > 0 && H >= 7 && W >= 7, "slq_stem_forward: bad shape");
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_U8 || out_mode == SLQ_OUT_F32, "slq_stem_forward: out_mode %d", out_mode);
  SLQ_CHECK_ARG(out_mode == SLQ_OUT_F32 || act_scales, "slq_stem_forward: act_scales required");
  const int Hc = (H + 6 - 7) / 2 + 1, Wc = (W + 6 - 7) / 2 + 1;
  const int Hp = (Hc + 2 - 3) / 2 + 1, Wp = (Wc + 2 - 3) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(Wc, kStemTile), (unsigned)ceil_div(Hc, kStemTile), (unsigned)N);
  stem_conv_kernel<<<grid, 256, 0, st>>>(x, N, H, W, Hc, Wc, w, bn_a, bn_b, scratch);
  SLQ_LAUNCH_CHECK();
  return launch_stem_pool(scratch, N, Hc, Wc, Hp, Wp, act_scales, out_id, out, out_mode, out_rowsum, st);
}

extern "C" int slq_absmax_scale(const float *y, int64_t n, float *act_scales, int32_t id,
                                int32_t qmax, uint32_t *tmp, void *stream) {
  SLQ_CHECK_ARG(y && act_scales && tmp && n > 0 && qmax > 0, "slq_absmax_scale: bad argument");
  SLQ_CHECK_ARG(reinterpret_cast<uintptr_t>(y) % 16 == 0, "slq_absmax_scale: y must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  SLQ_CUDA(cudaMemsetAsync(tmp, 0, sizeof(uint32_t), st));
  const int blocks = (int)std::min<int64_t>(ceil_div(n / 4 + 1, 256), (int64_t)sm_count() * 8);
  absmax_kernel<<<blocks, 256, 0, st>>>(y, n, tmp);
  SLQ_LAUNCH_CHECK();
  absmax_finalize_kernel<<<1, 1, 0, st>>>(tmp, act_scales, id, (float)qmax);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_quantize_act(const float *y, int64_t n, const float *act_scales, int32_t id,
                                int32_t is_signed, uint8_t *out, void *stream) {
  SLQ_CHECK_ARG(y && act_scales && out && n > 0, "slq_quantize_act: bad argument");
  SLQ_CHECK_ARG(reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0,
                "slq_quantize_act: misaligned buffers");
  const int blocks = (int)std::min<int64_t>(ceil_div(n / 4 + 1, 256), (int64_t)sm_count() * 8);
  quantize_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, n, act_scales, id, is_signed, out);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_build_packed_gemm_weights(const slq_conv_desc *d, const uint8_t *codes, const int64_t *code_offsets,
                                             const int32_t *bit, const int64_t *tile_base, const int32_t *seg_bytes,
                                             const uint16_t *row_offsets, uint8_t *wgp, void *stream) {
  int rc = validate_desc(d);
  if (rc != SLQ_OK) return rc;
  SLQ_CHECK_ARG(codes && code_offsets && bit && tile_base && seg_bytes && row_offsets && wgp,
                "slq_build_packed_gemm_weights: null pointer argument");
  SLQ_CHECK_ARG(!d->w16, "slq_build_packed_gemm_weights: two-limb (never-quantised) layers are not packed");
  const ConvGeom g = make_geom(*d);
  const int swz = (d->Cin % 128 == 0) ? 128 : 64;
  build_packed_gemm_weights_kernel<<<g.gemm_rows, 128, 0, (cudaStream_t)stream>>>(g, swz, codes, code_offsets, bit, tile_base,
                                                                                 seg_bytes, row_offsets, wgp);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}
