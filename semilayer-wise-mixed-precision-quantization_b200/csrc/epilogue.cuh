// csrc/epilogue.cuh -- the fused conv epilogue shared by the tcgen05 kernel and the SIMT checker.
//
// Replaces, per output element, the reference's separate passes
//   native_batch_norm (resnet.py:58,61,100,104,108)  -> folded scale/bias
//   add_ residual     (resnet.py:65,113)
//   relu_             (resnet.py:59,66,101,105,114)
// and adds what an integer conv needs: zero-point correction z[oc]*S[m], de-quantisation by
// s32[oc]*s_in, and re-quantisation of the activation to u8/s8 with the static scale of the
// output tensor.  Every fp32 op is separately rounded (no FMA) so that oracle/slq_oracle.py
// reproduces the bytes exactly with numpy float32.
#pragma once
#include <cstdint>

#include "slq.h"

namespace slq {

constexpr int SLQ_OUT_S8_INTERNAL = 3;  // == SLQ_OUT_S8 in slq.h

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

// de-quantised, BN-folded pre-activation of one output element
__device__ __forceinline__ float epi_value(int acc_lo, int acc_hi, bool w16, float Sf, float zf,
                                           float wsc /* wscale[oc] * s_in */, float bias) {
  float accf = (float)acc_lo;
  if (w16) accf = __fadd_rn(__fmul_rn((float)acc_hi, 256.0f), accf);
  const float v = __fadd_rn(accf, __fmul_rn(zf, Sf));
  return __fadd_rn(__fmul_rn(v, wsc), bias);
}

__device__ __forceinline__ float epi_residual_relu(float y, bool has_res, int res_raw,
                                                   bool res_signed, float s_res, bool relu) {
  if (has_res) {
    const float r = res_signed ? (float)(int8_t)res_raw : (float)res_raw;
    y = __fadd_rn(y, __fmul_rn(r, s_res));
  }
  return relu ? fmaxf(y, 0.0f) : y;
}

__device__ __forceinline__ uint32_t epi_quant_u8(float y, float inv_s_out) {
  const float q = rintf(__fmul_rn(y, inv_s_out));
  return (uint32_t)fminf(fmaxf(q, 0.0f), 255.0f);
}
__device__ __forceinline__ uint32_t epi_quant_s8(float y, float inv_s_out) {
  const float q = rintf(__fmul_rn(y, inv_s_out));
  return (uint32_t)(int)fminf(fmaxf(q, -127.0f), 127.0f) & 0xffu;
}

}  // namespace slq
