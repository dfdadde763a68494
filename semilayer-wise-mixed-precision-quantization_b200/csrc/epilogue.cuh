// csrc/epilogue.cuh -- the fused conv epilogue shared by the tcgen05 kernel and the SIMT checker.
//
// Replaces, per output element, the reference's separate passes
//   native_batch_norm (resnet.py:58,61,100,104,108)  -> folded scale/bias
//   add_ residual     (resnet.py:65,113)
//   relu_             (resnet.py:59,66,101,105,114)
// and adds what an integer conv needs: zero-point correction z[oc]*S[m], de-quantisation by
// s32[oc]*s_in, and re-quantisation of the activation to u8/s8 with the static scale of the
// output tensor.
//
// Arithmetic contract (oracle/slq_oracle.py `epilogue` / `epilogue_q` restate it exactly; fma = one
// rounding, every other op a separately rounded fp32 op):
//   per channel : wsc = wscale*s_in ; zw = zf*wsc
//                 fp32 out  : A = wsc        Z = zw        B = bias        sr = s_res
//                 u8/s8 out : A = wsc*inv    Z = zw*inv    B = bias*inv    sr = s_res*inv
//                             with inv = 1/act_scales[out_id]: the re-quantisation multiply is folded
//                             into the per-channel constants, so an element costs 3 FMAs + converts
//   per element : accf = f32(acc)            [two limbs: fma(f32(hi), 256, f32(lo))]
//                 y    = fma(accf, A, fma(f32(S), Z, B))
//                 y    = fma(f32(res), sr, y)                           (if residual)
//   fp32 out    : relu ? max(y, 0) : y
//   u8 out      : sat_u8(rint(y))      saturation at 0 IS the ReLU
//   s8 out      : sat_s8(rint(y))      (tensors that are not post-ReLU)
#pragma once
#include <cstdint>

#include "slq.h"

namespace slq {

struct EpiDev {  // device-side view of slq_epilogue (+ layer constants)
  const float *wscale, *zf, *bias, *act_scales;
  const uint8_t *res;
  void *out;
  int32_t *out_S;
  const uint32_t *in_rowsum;  // [in_planes][in_plane_stride] per-pixel channel sums of the INPUT activation: the sum of the planes
  uint32_t *out_rowsum;       // [n_tiles][M] per-pixel sums over each n-tile's channels of the u8 OUTPUT (or NULL)
  int in_planes;
  long long in_plane_stride;
  int in_id, out_id, res_id;
  int out_mode, relu, res_signed;
  int Cout, w16;
  long long M;
};

struct ChanParam {  // per output channel, staged in shared memory: (A, Z, B) of the contract above
  float wsc, zw, bias, pad;
};

// inv_out = 1/act_scales[out_id] for quantised outputs; pass quantised = false for fp32 / raw outputs
__device__ __forceinline__ ChanParam make_chan_param(float wscale, float zf, float bias, float s_in,
                                                     float inv_out, bool quantised) {
  ChanParam p;
  p.wsc = __fmul_rn(wscale, s_in);
  p.zw = __fmul_rn(zf, p.wsc);
  p.bias = bias;
  p.pad = 0.f;
  if (quantised) {
    p.wsc = __fmul_rn(p.wsc, inv_out);
    p.zw = __fmul_rn(p.zw, inv_out);
    p.bias = __fmul_rn(p.bias, inv_out);
  }
  return p;
}

template <bool W16>
__device__ __forceinline__ float epi_value(int acc_lo, int acc_hi, float Sf, const ChanParam &p) {
  float accf = (float)acc_lo;
  if (W16) accf = __fmaf_rn((float)acc_hi, 256.0f, accf);
  const float c2 = __fmaf_rn(Sf, p.zw, p.bias);
  return __fmaf_rn(accf, p.wsc, c2);
}

__device__ __forceinline__ float epi_add_res(float y, uint32_t res_byte, bool res_signed, float s_res) {
  const float r = res_signed ? (float)(int)(int8_t)res_byte : (float)res_byte;
  return __fmaf_rn(r, s_res, y);
}

// two independent fused multiply-adds in ONE instruction (sm_100 FFMA2): d = a * b + c per component, each
// rounded once, i.e. bit-identical to two __fmaf_rn calls
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#ifdef SLQ_NO_FFMA2  // A/B timing only (tools/gpu_*.sh build a variant library with it)
  return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#endif
  uint64_t ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// y is already in units of the output scale (see the contract)
__device__ __forceinline__ uint32_t epi_quant_u8(float y) {
  uint32_t q;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q) : "f"(y));
  return q;
}
__device__ __forceinline__ uint32_t epi_quant_s8(float y) {
  int q;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(y));
  return (uint32_t)q & 0xffu;
}
// four values -> four packed bytes: rint to s32, then two saturating pack instructions
template <bool SIGNED>
__device__ __forceinline__ uint32_t epi_pack4(float y0, float y1, float y2, float y3) {
  int i0, i1, i2, i3;
  asm("cvt.rni.s32.f32 %0, %1;" : "=r"(i0) : "f"(y0));
  asm("cvt.rni.s32.f32 %0, %1;" : "=r"(i1) : "f"(y1));
  asm("cvt.rni.s32.f32 %0, %1;" : "=r"(i2) : "f"(y2));
  asm("cvt.rni.s32.f32 %0, %1;" : "=r"(i3) : "f"(y3));
  uint32_t hi2, w;
  if (SIGNED) {
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi2) : "r"(i3), "r"(i2));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(i1), "r"(i0), "r"(hi2));
  } else {
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi2) : "r"(i3), "r"(i2));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(i1), "r"(i0), "r"(hi2));
  }
  return w;
}

}  // namespace slq
