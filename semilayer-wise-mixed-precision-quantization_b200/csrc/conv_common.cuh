// csrc/conv_common.cuh -- geometry of one quantised conv layer viewed as an implicit GEMM.
//   M = N*Ho*Wo output pixels, N_gemm = Cout (x2 limbs in w16 mode), K = kh*kw*Cin ordered (r,s,c).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace slq {

constexpr int kTileM = 128;  // output pixels per tile == TMEM lanes == UMMA_M

struct ConvGeom {
  int N, H, W, Cin, Cout, kh, kw, stride, pad;
  int Ho, Wo, w16;
  int bn_ch;      // output channels per N tile: 64, 128 or (wide tiles of K-heavy layers) 256; always 64 in w16 mode
  int bn_cols;    // GEMM columns per N tile == UMMA N (bn_ch, or 2*bn_ch in w16 mode)
  int n_tiles;    // tiles along N
  int gemm_rows;  // rows of the GEMM-ready weight matrix = n_tiles * bn_cols
  long long M;    // output pixels
  int Ktot;       // kh*kw*Cin
};

// wide: 256-channel tiles (UMMA N = 256, the shape at which ONE tcgen05.mma keeps the tensor pipe busy for as
// long as it takes to issue the next one).  Only for one-limb layers whose Cout is a multiple of 256; the
// GEMM-ready weight matrix is the same for both tilings (row oc = channel oc).
inline ConvGeom make_geom(const slq_conv_desc &d, int wide = 0) {
  ConvGeom g{};
  g.N = d.N; g.H = d.H; g.W = d.W; g.Cin = d.Cin; g.Cout = d.Cout;
  g.kh = d.kh; g.kw = d.kw; g.stride = d.stride; g.pad = d.pad; g.w16 = d.w16 ? 1 : 0;
  g.Ho = (d.H + 2 * d.pad - d.kh) / d.stride + 1;
  g.Wo = (d.W + 2 * d.pad - d.kw) / d.stride + 1;
  g.bn_ch = g.w16 ? 64 : (d.Cout > 64 ? 128 : 64);
  if (wide && !g.w16 && d.Cout % 256 == 0) g.bn_ch = 256;
  g.bn_cols = g.w16 ? 128 : g.bn_ch;
  g.n_tiles = (d.Cout + g.bn_ch - 1) / g.bn_ch;
  g.gemm_rows = g.n_tiles * g.bn_cols;
  g.M = (long long)d.N * g.Ho * g.Wo;
  g.Ktot = d.kh * d.kw * d.Cin;
  return g;
}

// row of the GEMM-ready weight matrix that holds limb `limb` (0 = low, 1 = high) of channel oc
__host__ __device__ inline int gemm_row_of(int oc, int limb, int w16) {
  return w16 ? ((oc >> 6) * 128 + limb * 64 + (oc & 63)) : oc;
}

int validate_desc(const slq_conv_desc *d);  // SLQ_OK or error (message set)

// debug timeline buffer installed by slq_debug_set_trace (conv_umma.cu); NULL when tracing is off
void debug_trace_buffer(long long **buf, int *cap);

// SIMT (dp4a) launcher, layers.cu
int launch_conv_simt(const ConvGeom &g, const uint8_t *in, const uint8_t *wg, const EpiDev &e,
                     cudaStream_t st);

// fp32 3x3 s2 max-pool of the stem (+ optional u8 quantisation), layers.cu
int launch_stem_pool(const float *y, int N, int Hc, int Wc, int Hp, int Wp, const float *act_scales,
                     int out_id, void *out, int out_mode, uint32_t *out_rowsum, cudaStream_t st);

}  // namespace slq

struct slq_conv {
  slq_conv_desc desc;
  slq::ConvGeom g;
  slq::ConvGeom g_wide;  // 256-channel tiling (valid when wide_ok)
  int wide_ok;           // the layer streams its weights and Cout % 256 == 0: launches without residual use g_wide
  CUtensorMap tmB_wide;
  // packed weights (slq_conv_set_packed_weights): device pointers, used when the layer keeps its weights resident
  const uint8_t *wgp;
  const long long *wgp_tile;
  const int *wgp_seg;
  const uint16_t *wgp_rowoff;
  const uint8_t *in;
  const uint8_t *wg;
  int a_im2col;     // 1: A operand through im2col-mode TMA, 0: tiled TMA over [M, Cin]
  int swizzle;      // 64 or 128: bytes of K per pipeline stage == TMA/UMMA swizzle span
  CUtensorMap tmA;  // activations
  CUtensorMap tmB;  // GEMM-ready weights
  CUtensorMap tmO;  // [M, Cout] output (TMA store), encoded for out_ptr at launch
  CUtensorMap tmR;  // [M, Cout] residual (TMA load), encoded for res_ptr at launch
  const void *out_ptr;
  const void *res_ptr;
  int num_ctas;
  int smem_bytes;
};
