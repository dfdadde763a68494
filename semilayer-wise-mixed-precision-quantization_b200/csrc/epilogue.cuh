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
// Arithmetic contract (oracle/slq_oracle.py `epilogue_v2` restates it exactly; fma = one rounding):
//   per channel : wsc = wscale*s_in ; zw = zf*wsc                       (fp32 mul each)
//   per element : accf = f32(acc)            [two limbs: fma(f32(hi), 256, f32(lo))]
//                 c2   = fma(f32(S), zw, bias)
//                 y    = fma(accf, wsc, c2)
//                 y    = fma(f32(res), s_res, y)                        (if residual)
//   fp32 out    : relu ? max(y, 0) : y
//   u8 out      : cvt.rni.sat.u8(y * inv_out)      saturation at 0 IS the ReLU
//   s8 out      : cvt.rni.sat.s8(y * inv_out)      (tensors that are not post-ReLU)
// ~8 issue slots per element instead of ~17 for the separately-rounded form.
#pragma once
#include <cstdint>

#include "slq.h"

namespace slq {

struct EpiDev {  // device-side view of slq_epilogue (+ layer constants)
  const float *wscale, *zf, *bias, *act_scales;
  const uint8_t *res;
  void *out;
  int32_t *out_S;
  int in_id, out_id, res_id;
  int out_mode, relu, res_signed;
  int Cout, w16;
  long long M;
};

struct ChanParam {  // per output channel, staged in shared memory
  float wsc, zw, bias, pad;
};

__device__ __forceinline__ ChanParam make_chan_param(float wscale, float zf, float bias, float s_in) {
  ChanParam p;
  p.wsc = __fmul_rn(wscale, s_in);
  p.zw = __fmul_rn(zf, p.wsc);
  p.bias = bias;
  p.pad = 0.f;
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

__device__ __forceinline__ uint32_t epi_quant_u8(float y, float inv_s_out) {
  uint32_t q;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q) : "f"(__fmul_rn(y, inv_s_out)));
  return q;
}
__device__ __forceinline__ uint32_t epi_quant_s8(float y, float inv_s_out) {
  int q;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(__fmul_rn(y, inv_s_out)));
  return (uint32_t)q & 0xffu;
}

}  // namespace slq
