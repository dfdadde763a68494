// csrc/eval_tail.cu -- the evaluation tail of the reference as ONE pass over the logits.
//
// Replaces, per batch, the ATen chain of functions.py:109-122 (evaluate_acc_loss_softmax)
//     output.max(1)                      -> argmax / correct count
//     CrossEntropyLoss()(output, y)      -> batch-mean negative log-likelihood
//     Softmax(dim=1)(output)             -> probabilities (kept: the mains hand them to KLdiv)
// and the Python double loop of functions.py:142-146 (KLdiv: one .sum() and one list append per SAMPLE)
//     kl_m = sum_c p[m,c] * log(p[m,c] / q[m,c])
// with two launches: a warp per sample (three passes over its row, which stays in L1), then one CTA that
// folds the per-sample results in a FIXED order into four double accumulators on the device.  Nothing
// comes back to the host until the caller reads the accumulators once per evaluation
// (the reference synchronises three times per evaluation and once per sample inside KLdiv).
//
// Arithmetic is fp32 per element like ATen's (exp(x - max) / sum ; x - max - log(sum)); the per-sample
// and per-batch sums are carried in fp32 / double in a fixed order, so results are deterministic and
// agree with ATen's to a few ulp (tests state the tolerance).  0 * log(0 / 0) stays NaN and p * log(p / 0)
// stays +inf exactly as in the reference.
#include "common.cuh"

namespace slq {

constexpr int kEvalWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rows[3*B]: [0,B) cross-entropy per sample, [B,2B) KL per sample, [2B,3B) 1.0 if argmax == label
__global__ void __launch_bounds__(kEvalWarps * 32) eval_rows_kernel(
    const float *__restrict__ logits, const int64_t *__restrict__ labels, int B, int C,
    float *__restrict__ probs_out, const float *__restrict__ ref_probs, float *__restrict__ rows) {
  const int b = blockIdx.x * kEvalWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float *x = logits + (long long)b * C;
  // pass 1: max and first index of the max (torch.max returns the first maximal index on ties)
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = x[c];
    if (v > m || (v == m && c < am)) { m = v; am = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  // pass 2: sum of exp(x - max)
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(__fsub_rn(x[c], m));
  s = warp_sum(s);
  // pass 3: probabilities, KL against the reference probabilities
  float kl = 0.f;
  const float *p = ref_probs ? ref_probs + (long long)b * C : nullptr;
  float *q_out = probs_out ? probs_out + (long long)b * C : nullptr;
  if (p || q_out) {
    for (int c = lane; c < C; c += 32) {
      const float q = __fdiv_rn(expf(__fsub_rn(x[c], m)), s);
      if (q_out) q_out[c] = q;
      if (p) {
        const float pc = p[c];
        kl += __fmul_rn(pc, logf(__fdiv_rn(pc, q)));  // functions.py:145, element for element
      }
    }
    kl = warp_sum(kl);
  }
  if (lane == 0) {
    float ce = 0.f, ok = 0.f;
    if (labels) {
      const long long y = labels[b];
      if (y >= 0 && y < C) ce = -(__fsub_rn(__fsub_rn(x[y], m), logf(s)));  // -log_softmax(x)[y]
      ok = (y == (long long)am) ? 1.f : 0.f;
    }
    rows[b] = ce;
    rows[B + b] = kl;
    rows[2 * B + b] = ok;
  }
}

// same per-sample KL for two PROBABILITY tensors (functions.KLdiv called on stored softmax outputs)
__global__ void __launch_bounds__(kEvalWarps * 32) kl_rows_kernel(const float *__restrict__ p,
                                                                  const float *__restrict__ q, int B, int C,
                                                                  float *__restrict__ rows) {
  const int b = blockIdx.x * kEvalWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float kl = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float pc = p[(long long)b * C + c];
    kl += __fmul_rn(pc, logf(__fdiv_rn(pc, q[(long long)b * C + c])));
  }
  kl = warp_sum(kl);
  if (lane == 0) { rows[b] = 0.f; rows[B + b] = kl; rows[2 * B + b] = 0.f; }
}

// accum[0] += #correct ; accum[1] += mean_b CE (functions.py:115: one batch-mean loss per batch) ;
// accum[2] += sum_b KL_b ; accum[3] += B.  One CTA, fixed summation order.
__global__ void __launch_bounds__(256) eval_fold_kernel(const float *__restrict__ rows, int B, int has_labels,
                                                        double *__restrict__ accum) {
  __shared__ double sh[3][256];
  double ce = 0.0, kl = 0.0, ok = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) {
    ce += (double)rows[i];
    kl += (double)rows[B + i];
    ok += (double)rows[2 * B + i];
  }
  sh[0][threadIdx.x] = ce; sh[1][threadIdx.x] = kl; sh[2][threadIdx.x] = ok;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (has_labels) {
      accum[0] += sh[2][0];
      accum[1] += (double)(float)(sh[0][0] / (double)B);  // the batch mean is an fp32 value in the reference
    }
    accum[2] += sh[1][0];
    accum[3] += (double)B;
  }
}

}  // namespace slq

using namespace slq;

extern "C" int slq_eval_tail(const float *logits, const int64_t *labels, int32_t B, int32_t C, float *probs_out,
                             const float *ref_probs, float *rows, double *accum, void *stream) {
  SLQ_CHECK_ARG(logits && rows && accum, "slq_eval_tail: null pointer argument");
  SLQ_CHECK_ARG(B > 0 && C > 0, "slq_eval_tail: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  eval_rows_kernel<<<(unsigned)ceil_div(B, kEvalWarps), kEvalWarps * 32, 0, st>>>(logits, labels, B, C, probs_out,
                                                                                 ref_probs, rows);
  SLQ_LAUNCH_CHECK();
  eval_fold_kernel<<<1, 256, 0, st>>>(rows, B, labels != nullptr, accum);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_kl_rows(const float *p, const float *q, int32_t B, int32_t C, float *rows, double *accum,
                           void *stream) {
  SLQ_CHECK_ARG(p && q && rows && accum, "slq_kl_rows: null pointer argument");
  SLQ_CHECK_ARG(B > 0 && C > 0, "slq_kl_rows: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  kl_rows_kernel<<<(unsigned)ceil_div(B, kEvalWarps), kEvalWarps * 32, 0, st>>>(p, q, B, C, rows);
  SLQ_LAUNCH_CHECK();
  eval_fold_kernel<<<1, 256, 0, st>>>(rows, B, 0, accum);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}
