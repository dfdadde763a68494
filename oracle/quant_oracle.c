/*
 * oracle/quant_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement, in plain C, of the reference's per-output-channel affine fake-quantizer:
 *     /root/reference/functions.py:25-43   quantize_wgt(tensor, bit)
 *     /root/reference/functions.py:9-23    channel_wise_quantizationperchan(tensor, bit, i)
 * following the step list A.1-A.9 of SURVEY.md Appendix A.
 *
 * Parity pin: tests/test_oracle_pins.py checks this file against tests/golden/quant_rows_*.npz,
 * which oracle/gen_golden.py produced by importing the unmodified reference functions.py in the
 * build container (torch 2.11.0 CPU).  The reference ships no tests / golden vectors of its own.
 *
 * Compile (see oracle/Makefile):  gcc -O2 -ffp-contract=off -fPIC -shared
 * -ffp-contract=off matters: every fp32 intermediate of functions.py:41 is a separate ATen
 * kernel, i.e. a separate IEEE rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define SLQ_ORACLE_OK 0
#define SLQ_ORACLE_ZERO_RANGE 1 /* reference raises ZeroDivisionError (functions.py:40) */
#define SLQ_ORACLE_CODE_RANGE 2 /* a code fell outside [0, 2^bit-1]; value clamped in codes_out */

/* div_mode 0: true IEEE fp32 divide      (ATen CPU path of `tensor/scale`, functions.py:41)
 * div_mode 1: multiply by float32(1.0/scale64) (ATen CUDA div_true kernel with a CPU-scalar divisor) */
int slq_oracle_quantize_row(const float *w, int64_t K, int bit, int div_mode,
                            float *q_out,      /* K  fake-quantised fp32 values (may alias w)   */
                            int32_t *codes_out, /* K  stored codes u = k - z (may be NULL)        */
                            int32_t *z_out, float *s32_out, double *scale64_out)
{
    /* A.1  functions.py:35-36  min/max, .item() widens exactly to double */
    float mnf = w[0], mxf = w[0];
    for (int64_t i = 1; i < K; ++i) {
        if (w[i] < mnf) mnf = w[i];
        if (w[i] > mxf) mxf = w[i];
    }
    double mn = (double)mnf, mx = (double)mxf;
    /* A.2  functions.py:39 */
    double levels = (double)((1LL << bit) - 1);
    double scale = (mx - mn) / levels;
    if (scale64_out) *scale64_out = scale;
    if (scale == 0.0) return SLQ_ORACLE_ZERO_RANGE;
    /* A.3  functions.py:40  Python round() == round-half-to-even == rint() in the default mode */
    double zd = rint(mn / scale);
    /* A.4  the python float meets an fp32 tensor: demoted with round-to-nearest-even */
    float s32 = (float)scale;
    float zf = (float)zd;
    /* ATen CUDA div by a CPU scalar: inv_b = float(1.0 / double(scalar)), computed on the host in
     * double from the python float (measured on B200, torch 2.11: tools/diag_div.py, 0 of 21M off) */
    volatile float inv = (float)(1.0 / scale);
    int status = SLQ_ORACLE_OK;
    int32_t maxcode = (int32_t)((1LL << bit) - 1);
    for (int64_t i = 0; i < K; ++i) {
        volatile float t1 = div_mode ? (w[i] * inv) : (w[i] / s32); /* A.5 */
        volatile float t2 = t1 + zf;                                  /* A.6 */
        volatile float t3 = rintf(t2);                                /* A.7 torch.round = half-even */
        volatile float k = t3 - zf;                                   /* A.8 */
        volatile float q = k * s32;                                   /* A.9 */
        q_out[i] = q;
        if (codes_out) {
            /* w/s lies in [z, z+2^bit-1], so t3 = round(w/s + z) lies in [2z, 2z+2^bit-1] and the
             * level the reference keeps is k = t3 - z (functions.py:41 "(...).round() - z").
             * min k == z and max k == z + 2^bit - 1, so the storage code is
             * u = k - z = t3 - 2z in [0, 2^bit-1] and the real weight is (u + z) * s32.       */
            double u = (double)t3 - 2.0 * zd;
            int32_t ui = (int32_t)u;
            if (u < 0.0) { ui = 0; status = SLQ_ORACLE_CODE_RANGE; }
            else if (u > (double)maxcode) { ui = maxcode; status = SLQ_ORACLE_CODE_RANGE; }
            codes_out[i] = ui;
        }
    }
    if (z_out) *z_out = (int32_t)zd;
    if (s32_out) *s32_out = s32;
    return status;
}

/* functions.py:9-23: in-place overwrite of row i of a contiguous [rows, K] weight tensor. */
int slq_oracle_channel_wise(float *tensor, int64_t K, int64_t i, int bit, int div_mode)
{
    float *row = tensor + i * K;
    return slq_oracle_quantize_row(row, K, bit, div_mode, row, NULL, NULL, NULL, NULL);
}
