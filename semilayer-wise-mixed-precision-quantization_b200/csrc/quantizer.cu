// csrc/quantizer.cu -- per-output-channel affine weight quantizer for sm_100a (HBM-bound).
//
// Replaces the ATen call chain of the reference's
//   functions.py:25-43  quantize_wgt            (7 launches + 2 .item() syncs per row)
//   functions.py:9-23   channel_wise_quantizationperchan (row write-back)
// with ONE launch per work-list of (row, bit) jobs.  Arithmetic follows SURVEY.md Appendix A
// step by step: fp64 scale / zero-point, then five separately rounded fp32 ops (no FMA
// contraction: every op below is an explicit __f*_rn intrinsic).
//
// Data movement: a row (K <= 4608 fp32) is read from HBM exactly once into registers
// (float4 loads, 32 threads per row for K <= 1152, 128 threads per row above), reduced for
// min/max, transformed, and written once (fake-quant fp32 write-back + packed codes).
#include <algorithm>

#include "common.cuh"

namespace slq {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 ? dev : 0;
}
int sm_count() {  // cached per device: a process may drive several GPUs
  static int cached[kMaxDevices] = {0};
  const int dev = current_device();
  if (dev >= kMaxDevices) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      return 148;
  }
  return cached[dev];
}

constexpr int kBlock = 128;  // threads per CTA for every quantizer kernel
constexpr int kVpt = 9;      // float4 per thread held in registers: 36 floats

struct Affine {  // result of Appendix A.1-A.4 for one row
  float s32;
  float inv;  // SLQ_DIV_RECIP multiplier: float32(1.0 / scale64), what ATen computes on the host
  float zf;
  long long z;
  int ok;  // 0: zero range
};

__device__ __forceinline__ Affine derive_affine(float mnf, float mxf, int bit) {
  Affine a;
  const double mn = (double)mnf, mx = (double)mxf;
  const double levels = (double)((1LL << bit) - 1);
  const double scale = __ddiv_rn(__dsub_rn(mx, mn), levels);  // functions.py:39
  a.ok = (scale != 0.0);
  const double zd = a.ok ? rint(__ddiv_rn(mn, scale)) : 0.0;   // functions.py:40 (half-even)
  a.s32 = __double2float_rn(scale);
  a.inv = __double2float_rn(__ddiv_rn(1.0, scale));
  a.zf = __double2float_rn(zd);
  a.z = (long long)zd;
  return a;
}

// functions.py:41, one element.  Returns the fake-quantised value; *t3 is round(w/s + z).
template <int DIV>
__device__ __forceinline__ float fake_quant(float w, float s32, float inv, float zf, float *t3) {
  const float t1 = DIV == SLQ_DIV_RECIP ? __fmul_rn(w, inv) : __fdiv_rn(w, s32);  // A.5
  const float t2 = __fadd_rn(t1, zf);                                             // A.6
  *t3 = rintf(t2);                                                                // A.7
  const float k = __fsub_rn(*t3, zf);                                             // A.8
  return __fmul_rn(k, s32);                                                       // A.9
}

// Packs the 4 codes of one float4 group and stores them: 8/6 bit -> 4 bytes, 4 bit -> 2, 2 bit -> 1,
// 16 bit -> low limbs at p[0..4), high limbs at p[K..K+4).
__device__ __forceinline__ void store_codes4(uint8_t *row_codes, int64_t elem, int64_t K, int bit,
                                             const int u[4]) {
  if (bit == 4) {
    uchar2 v = make_uchar2((uint8_t)(u[0] | (u[1] << 4)), (uint8_t)(u[2] | (u[3] << 4)));
    *reinterpret_cast<uchar2 *>(row_codes + (elem >> 1)) = v;
  } else if (bit == 2) {
    row_codes[elem >> 2] = (uint8_t)(u[0] | (u[1] << 2) | (u[2] << 4) | (u[3] << 6));
  } else if (bit == 16) {
    *reinterpret_cast<uchar4 *>(row_codes + elem) =
        make_uchar4(u[0] & 255, u[1] & 255, u[2] & 255, u[3] & 255);
    *reinterpret_cast<uchar4 *>(row_codes + K + elem) =
        make_uchar4(u[0] >> 8, u[1] >> 8, u[2] >> 8, u[3] >> 8);
  } else {
    *reinterpret_cast<uchar4 *>(row_codes + elem) = make_uchar4(u[0], u[1], u[2], u[3]);
  }
}

// min/max across the GROUP threads that share a row.
template <int GROUP>
__device__ __forceinline__ void group_minmax(float &mn, float &mx, float *smem) {
#pragma unroll
  for (int o = (GROUP < 32 ? GROUP : 32) / 2; o > 0; o >>= 1) {  // xor offsets below GROUP stay inside the group
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (GROUP > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      smem[warp] = mn;
      smem[8 + warp] = mx;
    }
    __syncthreads();
    mn = smem[0];
    mx = smem[8];
#pragma unroll
    for (int i = 1; i < GROUP / 32; ++i) {
      mn = fminf(mn, smem[i]);
      mx = fmaxf(mx, smem[8 + i]);
    }
    __syncthreads();
  }
}

// A row held in registers by GROUP threads: float4 #i of thread t covers elements 4*(t + i*GROUP).
template <int GROUP>
struct RowRegs {
  float4 v[kVpt];
  int nvec, t;
  __device__ __forceinline__ void load(const float *row, int64_t K) {
    nvec = (int)(K >> 2);
    t = threadIdx.x % GROUP;
    const float4 *p = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int i = 0; i < kVpt; ++i) {
      const int idx = t + i * GROUP;
      if (idx < nvec) v[i] = __ldg(p + idx);
    }
  }
  __device__ __forceinline__ void minmax(float &mn, float &mx) const {
    mn = INFINITY;
    mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kVpt; ++i) {
      if (t + i * GROUP < nvec) {
        mn = fminf(mn, fminf(fminf(v[i].x, v[i].y), fminf(v[i].z, v[i].w)));
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// K1: quantize a work-list of (row, bit) jobs (reference path, bit-exact)
// ------------------------------------------------------------------------------------------
template <int GROUP, int DIV>
__global__ void __launch_bounds__(kBlock) quantize_rows_kernel(
    float *__restrict__ w, int64_t K, const int32_t *__restrict__ rows,
    const int32_t *__restrict__ bits, int n_jobs, int write_back, uint8_t *__restrict__ codes,
    const int64_t *__restrict__ code_offsets, int32_t *__restrict__ z_out,
    float *__restrict__ s_out, int32_t *__restrict__ status) {
  __shared__ float red[16];
  const int job = blockIdx.x * (kBlock / GROUP) + threadIdx.x / GROUP;
  if (GROUP == 32 && job >= n_jobs) return;  // whole warp exits together
  const int jobc = job < n_jobs ? job : n_jobs - 1;
  float *row = w + (int64_t)rows[jobc] * K;
  const int bit = bits[jobc];
  RowRegs<GROUP> r;
  r.load(row, K);
  float mn, mx;
  r.minmax(mn, mx);
  group_minmax<GROUP>(mn, mx, red);
  const Affine a = derive_affine(mn, mx, bit);
  const bool leader = (threadIdx.x % GROUP) == 0;
  if (!a.ok) {  // reference: ZeroDivisionError; leave the row untouched
    if (leader) {
      status[job] = SLQ_ROW_ZERO_RANGE;
      if (z_out) z_out[job] = 0;
      if (s_out) s_out[job] = 0.f;
    }
    return;
  }
  const float inv = a.inv;
  const int maxcode = (1 << bit) - 1;
  uint8_t *row_codes = codes ? codes + code_offsets[job] : nullptr;
  float4 *out4 = reinterpret_cast<float4 *>(row);
  int bad = 0;
#pragma unroll
  for (int i = 0; i < kVpt; ++i) {
    const int idx = r.t + i * GROUP;
    if (idx < r.nvec) {
      float t3[4];
      float4 q;
      q.x = fake_quant<DIV>(r.v[i].x, a.s32, inv, a.zf, &t3[0]);
      q.y = fake_quant<DIV>(r.v[i].y, a.s32, inv, a.zf, &t3[1]);
      q.z = fake_quant<DIV>(r.v[i].z, a.s32, inv, a.zf, &t3[2]);
      q.w = fake_quant<DIV>(r.v[i].w, a.s32, inv, a.zf, &t3[3]);
      if (write_back) out4[idx] = q;
      if (row_codes) {
        int u[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // stored code u = k - z = t3 - 2z  (level k = t3 - z lies in [z, z + 2^bit - 1])
          long long c = (long long)t3[e] - 2 * a.z;
          if (c < 0) { c = 0; bad = 1; }
          if (c > maxcode) { c = maxcode; bad = 1; }
          u[e] = (int)c;
        }
        store_codes4(row_codes, (int64_t)idx * 4, K, bit, u);
      }
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  if (GROUP > 32) bad = __syncthreads_or(bad);
  if (leader) {
    status[job] = bad ? SLQ_ROW_CODE_RANGE : SLQ_ROW_OK;
    if (z_out) z_out[job] = (int32_t)a.z;
    if (s_out) s_out[job] = a.s32;
  }
}

// Generic fallback: any K (no alignment assumption), one CTA per job, row re-read from L1/L2.
template <int DIV>
__global__ void __launch_bounds__(kBlock) quantize_rows_generic_kernel(
    float *__restrict__ w, int64_t K, const int32_t *__restrict__ rows,
    const int32_t *__restrict__ bits, int n_jobs, int write_back, uint8_t *__restrict__ codes,
    const int64_t *__restrict__ code_offsets, int32_t *__restrict__ z_out,
    float *__restrict__ s_out, int32_t *__restrict__ status) {
  __shared__ float red[16];
  const int job = blockIdx.x;
  float *row = w + (int64_t)rows[job] * K;
  const int bit = bits[job];
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < K; i += kBlock) {
    const float v = row[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  group_minmax<kBlock>(mn, mx, red);
  const Affine a = derive_affine(mn, mx, bit);
  if (!a.ok) {
    if (threadIdx.x == 0) {
      status[job] = SLQ_ROW_ZERO_RANGE;
      if (z_out) z_out[job] = 0;
      if (s_out) s_out[job] = 0.f;
    }
    return;
  }
  const float inv = a.inv;
  const int maxcode = (1 << bit) - 1;
  uint8_t *row_codes = codes ? codes + code_offsets[job] : nullptr;
  int bad = 0;
  // groups of 4 consecutive elements so that sub-byte codes of one byte belong to one thread
  for (int64_t g = threadIdx.x; g * 4 < K; g += kBlock) {
    int u[4] = {0, 0, 0, 0};
    for (int e = 0; e < 4; ++e) {
      const int64_t i = g * 4 + e;
      if (i >= K) break;
      float t3;
      const float q = fake_quant<DIV>(row[i], a.s32, inv, a.zf, &t3);
      if (write_back) row[i] = q;
      long long c = (long long)t3 - 2 * a.z;
      if (c < 0) { c = 0; bad = 1; }
      if (c > maxcode) { c = maxcode; bad = 1; }
      u[e] = (int)c;
    }
    if (row_codes) {
      if (bit == 4) {
        row_codes[g * 2] = (uint8_t)(u[0] | (u[1] << 4));
        if (g * 4 + 2 < K) row_codes[g * 2 + 1] = (uint8_t)(u[2] | (u[3] << 4));
      } else if (bit == 2) {
        row_codes[g] = (uint8_t)(u[0] | (u[1] << 2) | (u[2] << 4) | (u[3] << 6));
      } else {
        for (int e = 0; e < 4 && g * 4 + e < K; ++e) row_codes[g * 4 + e] = (uint8_t)u[e];
      }
    }
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    status[job] = bad ? SLQ_ROW_CODE_RANGE : SLQ_ROW_OK;
    if (z_out) z_out[job] = (int32_t)a.z;
    if (s_out) s_out[job] = a.s32;
  }
}

// ------------------------------------------------------------------------------------------
// K2: content-derived classification (smallest exact grid in {2,4,6,8}, else 16 bit)
// ------------------------------------------------------------------------------------------
constexpr float kEncodeTol = 0.02f;

template <int GROUP>
__global__ void __launch_bounds__(kBlock) classify_rows_kernel(const float *__restrict__ w,
                                                               int64_t K, int n_rows,
                                                               int32_t *__restrict__ bit_out,
                                                               int32_t *__restrict__ z_out,
                                                               float *__restrict__ s_out,
                                                               int32_t *__restrict__ exact_out) {
  __shared__ float red[16];
  const int job = blockIdx.x * (kBlock / GROUP) + threadIdx.x / GROUP;
  if (GROUP == 32 && job >= n_rows) return;
  const int jobc = job < n_rows ? job : n_rows - 1;
  RowRegs<GROUP> r;
  r.load(w + (int64_t)jobc * K, K);
  int res_exact = 0;
  float mn, mx;
  r.minmax(mn, mx);
  group_minmax<GROUP>(mn, mx, red);
  const bool leader = (threadIdx.x % GROUP) == 0;
  int res_bit = 16;
  long long res_z = 1;
  float res_s = mn;
  bool found = false;
  if (mx == mn) {  // constant row: (0 + z) * s == mn
    if (mn == 0.f) { res_z = 0; res_s = 1.f; }
    found = true;
  }
  const int cand[5] = {2, 4, 6, 8, 16};
#pragma unroll 1
  for (int ci = 0; ci < 5 && !found; ++ci) {
    const int bit = cand[ci];
    const double levels = (double)((1LL << bit) - 1);
    const double scale = __ddiv_rn(__dsub_rn((double)mx, (double)mn), levels);
    const float s32 = __double2float_rn(scale);
    const float inv = __fdiv_rn(1.0f, s32);
    if (s32 == 0.f || isinf(inv)) continue;
    const double zd = rint(__ddiv_rn((double)mn, scale));
    // |z| must stay exactly representable in the epilogue's fp32 arithmetic; the 16-bit grid of a
    // never-quantised row may sit far from zero (narrow range around an offset), so it gets the full 2^23
    if (fabs(zd) > (bit == 16 ? 8.0e6 : 1.0e6)) continue;
    int fail = 0;
    if (bit != 16) {
#pragma unroll
      for (int i = 0; i < kVpt; ++i) {
        if (r.t + i * GROUP < r.nvec) {
          const float e[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float t = __fdiv_rn(e[j], s32);
            fail |= (fabsf(__fsub_rn(t, rintf(t))) > kEncodeTol);
          }
        }
      }
      fail = __any_sync(0xffffffffu, fail);
    }
    if (GROUP > 32) fail = __syncthreads_or(fail);  // uniform: every thread of the CTA gets here
    if (!fail) {
      res_bit = bit;
      res_z = (long long)zd;
      res_s = s32;
      found = true;
      if (bit != 16) {
        // Refinement: the scale re-derived from the row's own min/max can be an ulp or two off the scale
        // the row was quantised with.  Look, among its nearest neighbours, for the scale that REPRODUCES
        // every element exactly -- fp32(rint(w / s) * s) == w, functions.py:41's last two operations -- so
        // that the packed row decodes bit for bit (slq_decode_rows: snapshots, packed files).
        const int maxcode = (1 << bit) - 1;
#pragma unroll 1
        for (int d = 0; d < 5 && !res_exact; ++d) {
          const int delta = d == 0 ? 0 : (d == 1 ? -1 : (d == 2 ? 1 : (d == 3 ? -2 : 2)));
          const float sc = __int_as_float(__float_as_int(s32) + delta);
          const float kmin = rintf(__fdiv_rn(mn, sc));
          int bad = 0;
#pragma unroll
          for (int i = 0; i < kVpt; ++i) {
            if (r.t + i * GROUP < r.nvec) {
              const float e[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float k = rintf(__fdiv_rn(e[j], sc));
                const float c = __fsub_rn(k, kmin);
                bad |= !(__fmul_rn(k, sc) == e[j]) || c < 0.f || c > (float)maxcode;
              }
            }
          }
          bad = __any_sync(0xffffffffu, bad);
          if (GROUP > 32) bad = __syncthreads_or(bad);
          if (!bad && fabsf(kmin) <= 1.0e6f) {
            res_s = sc;
            res_z = (long long)kmin;
            res_exact = 1;
          }
        }
      }
    }
  }
  if (!found) {  // degenerate (denormal) range: treat as the constant mn
    res_bit = 16;
    res_z = 1;
    res_s = (mn != 0.f) ? mn : 1.f;
    if (mn == 0.f) res_z = 0;
  }
  if (leader && job < n_rows) {
    bit_out[job] = res_bit;
    z_out[job] = (int32_t)res_z;
    s_out[job] = res_s;
    if (exact_out) exact_out[job] = res_exact;
  }
}

// ------------------------------------------------------------------------------------------
// K3: pack rows whose (bit, z, s) are known: code = clamp(rint(w / s) - z, 0, 2^bit - 1)
// ------------------------------------------------------------------------------------------
template <int GROUP>
__global__ void __launch_bounds__(kBlock) encode_rows_kernel(
    const float *__restrict__ w, int64_t K, int n_rows, const int32_t *__restrict__ bit_in,
    const int32_t *__restrict__ z_in, const float *__restrict__ s_in, uint8_t *__restrict__ codes,
    const int64_t *__restrict__ code_offsets) {
  const int job = blockIdx.x * (kBlock / GROUP) + threadIdx.x / GROUP;
  if (job >= n_rows) return;
  RowRegs<GROUP> r;
  r.load(w + (int64_t)job * K, K);
  const int bit = bit_in[job];
  const float s = s_in[job];
  const float zf = (float)z_in[job];
  const float maxc = (float)((1 << bit) - 1);
  uint8_t *row_codes = codes + code_offsets[job];
#pragma unroll
  for (int i = 0; i < kVpt; ++i) {
    const int idx = r.t + i * GROUP;
    if (idx < r.nvec) {
      const float e[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
      int u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float k = rintf(__fdiv_rn(e[j], s));
        u[j] = (int)fminf(fmaxf(__fsub_rn(k, zf), 0.f), maxc);
      }
      store_codes4(row_codes, (int64_t)idx * 4, K, bit, u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1m: the same quantizer over a MULTI-TENSOR job table (slq_quantize_jobs): every job names its own row
// pointer, length and code destination, so ALL layers of a model are one launch (the reference issues
// 7 launches + 2 syncs per row, functions.py:35-41; resnet50_main.py:189-197 walks 22,656 rows per phase).
// Jobs are sorted by the caller into three classes, which are three block ranges of one grid:
//   K <= 288  : 8 lanes per row, 16 rows per CTA   (K = 64 / 128 / 256: a whole warp per row left 50-75 %
//               of its lanes without a float4 to load)
//   K <= 1152 : one warp per row, 4 rows per CTA
//   K <= 4608 : one CTA per row
// No early exits below warp granularity: sub-warp groups share shuffles and ballots.
// ------------------------------------------------------------------------------------------
template <int GROUP, int DIV>
__device__ __forceinline__ void quantize_job(const slq_qjob &jb, bool active, int job, int write_back,
                                             int32_t *__restrict__ z_out, float *__restrict__ s_out,
                                             int32_t *__restrict__ status, float *red) {
  const int64_t K = jb.K;
  float *row = jb.row;
  const int bit = jb.bit;
  RowRegs<GROUP> r;
  r.load(row, K);
  float mn, mx;
  r.minmax(mn, mx);
  group_minmax<GROUP>(mn, mx, red);
  const Affine a = derive_affine(mn, mx, bit);
  const bool leader = (threadIdx.x % GROUP) == 0;
  const bool go = active && a.ok;
  const int maxcode = (1 << bit) - 1;
  uint8_t *row_codes = jb.codes;
  float4 *out4 = reinterpret_cast<float4 *>(row);
  int bad = 0;
  if (go) {
#pragma unroll
    for (int i = 0; i < kVpt; ++i) {
      const int idx = r.t + i * GROUP;
      if (idx < r.nvec) {
        float t3[4];
        float4 q;
        q.x = fake_quant<DIV>(r.v[i].x, a.s32, a.inv, a.zf, &t3[0]);
        q.y = fake_quant<DIV>(r.v[i].y, a.s32, a.inv, a.zf, &t3[1]);
        q.z = fake_quant<DIV>(r.v[i].z, a.s32, a.inv, a.zf, &t3[2]);
        q.w = fake_quant<DIV>(r.v[i].w, a.s32, a.inv, a.zf, &t3[3]);
        if (write_back) out4[idx] = q;
        if (row_codes) {
          int u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            long long c = (long long)t3[e] - 2 * a.z;
            if (c < 0) { c = 0; bad = 1; }
            if (c > maxcode) { c = maxcode; bad = 1; }
            u[e] = (int)c;
          }
          store_codes4(row_codes, (int64_t)idx * 4, K, bit, u);
        }
      }
    }
  }
  if (GROUP <= 32) {
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned kOnes = GROUP < 32 ? ((1u << (GROUP & 31)) - 1u) : 0xffffffffu;  // GROUP lanes
    const unsigned gmask = kOnes << (lane & ~(unsigned)((GROUP - 1) & 31));
    bad = (__ballot_sync(0xffffffffu, bad) & gmask) != 0;
  } else {
    bad = __syncthreads_or(bad);
  }
  if (leader && active) {
    status[job] = !a.ok ? SLQ_ROW_ZERO_RANGE : (bad ? SLQ_ROW_CODE_RANGE : SLQ_ROW_OK);
    if (z_out) z_out[job] = a.ok ? (int32_t)a.z : 0;
    if (s_out) s_out[job] = a.ok ? a.s32 : 0.f;
  }
}

template <int DIV>
__global__ void __launch_bounds__(kBlock) quantize_jobs_kernel(const slq_qjob *__restrict__ jobs, int n_small,
                                                               int n_mid, int n_large, int blocks_small,
                                                               int blocks_mid, int write_back,
                                                               int32_t *__restrict__ z_out, float *__restrict__ s_out,
                                                               int32_t *__restrict__ status) {
  __shared__ float red[16];
  const int b = blockIdx.x;
  if (b < blocks_small) {
    const int job = b * (kBlock / 8) + threadIdx.x / 8;
    const bool active = job < n_small;
    quantize_job<8, DIV>(jobs[active ? job : n_small - 1], active, job, write_back, z_out, s_out, status, red);
  } else if (b < blocks_small + blocks_mid) {
    const int j = (b - blocks_small) * (kBlock / 32) + threadIdx.x / 32;
    const bool active = j < n_mid;
    const int job = n_small + j;
    quantize_job<32, DIV>(jobs[active ? job : n_small + n_mid - 1], active, job, write_back, z_out, s_out, status, red);
  } else {
    const int job = n_small + n_mid + (b - blocks_small - blocks_mid);
    quantize_job<128, DIV>(jobs[job], true, job, write_back, z_out, s_out, status, red);
  }
}

// K4: packed codes -> fp32 rows, w = fp32((code + z) * s): the exact inverse of the quantizer's write-back
// for rows of <= 8 bits ((code + z) is the integer level k of functions.py:41, k * scale is its last op).
__global__ void __launch_bounds__(256) decode_rows_kernel(const uint8_t *__restrict__ codes,
                                                          const int64_t *__restrict__ code_offsets,
                                                          const int32_t *__restrict__ bit_in,
                                                          const int32_t *__restrict__ z_in,
                                                          const float *__restrict__ s_in, int n_rows, int64_t K,
                                                          float *__restrict__ w) {
  const int64_t kq = (K + 3) >> 2;
  const int64_t total = (int64_t)n_rows * kq;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / kq);
    const int64_t e0 = (idx % kq) * 4;
    const int bit = bit_in[row];
    const float zf = (float)z_in[row], s = s_in[row];
    const uint8_t *src = codes + code_offsets[row];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = e0 + j;
      if (e >= K) break;
      int c;
      if (bit == 4) c = (src[e >> 1] >> ((e & 1) * 4)) & 15;
      else if (bit == 2) c = (src[e >> 2] >> ((e & 3) * 2)) & 3;
      else if (bit == 16) c = (int)src[e] | ((int)src[K + e] << 8);
      else c = src[e];
      w[(int64_t)row * K + e] = __fmul_rn(__fadd_rn((float)c, zf), s);
    }
  }
}

static bool fast_path_ok(const void *w, int64_t K, int group) {
  return (K % 4 == 0) && (reinterpret_cast<uintptr_t>(w) % 16 == 0) && (K <= (int64_t)group * kVpt * 4);
}

}  // namespace slq

using namespace slq;

extern "C" const char *slq_last_error(void) { return g_err; }
extern "C" int slq_abi_version(void) { return SLQ_ABI_VERSION; }

extern "C" int slq_device_info(int32_t *sms, int32_t *major, int32_t *minor) {
  int dev = 0;
  SLQ_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sms) {
    SLQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sms = v;
  }
  if (major) {
    SLQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *major = v;
  }
  if (minor) {
    SLQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *minor = v;
  }
  return SLQ_OK;
}

extern "C" int64_t slq_packed_row_bytes(int64_t K, int32_t bit) {
  if (bit == 16) return 2 * K;
  if (bit == 4) return (K + 1) / 2;
  if (bit == 2) return (K + 3) / 4;
  return K;
}

extern "C" int slq_quantize_rows(float *w, int64_t n_rows, int64_t K, const int32_t *rows,
                                 const int32_t *bits, int32_t n_jobs, int32_t div_mode,
                                 int32_t write_back, uint8_t *codes, const int64_t *code_offsets,
                                 int32_t *z, float *s32, int32_t *status, void *stream) {
  SLQ_CHECK_ARG(w && rows && bits && status, "slq_quantize_rows: null pointer argument");
  SLQ_CHECK_ARG(K > 0 && n_rows > 0, "slq_quantize_rows: K=%lld n_rows=%lld", (long long)K,
                (long long)n_rows);
  SLQ_CHECK_ARG(div_mode == SLQ_DIV_TRUE || div_mode == SLQ_DIV_RECIP,
                "slq_quantize_rows: div_mode %d", div_mode);
  SLQ_CHECK_ARG(!codes || code_offsets, "slq_quantize_rows: codes given without code_offsets");
  if (n_jobs <= 0) return SLQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define SLQ_QLAUNCH(KERNEL, GRID)                                                              \
  do {                                                                                         \
    if (div_mode == SLQ_DIV_TRUE)                                                              \
      KERNEL<SLQ_DIV_TRUE><<<(GRID), kBlock, 0, st>>>(w, K, rows, bits, n_jobs, write_back,     \
                                                     codes, code_offsets, z, s32, status);     \
    else                                                                                       \
      KERNEL<SLQ_DIV_RECIP><<<(GRID), kBlock, 0, st>>>(w, K, rows, bits, n_jobs, write_back,    \
                                                      codes, code_offsets, z, s32, status);    \
  } while (0)
  if (fast_path_ok(w, K, 32)) {
    if (div_mode == SLQ_DIV_TRUE)
      quantize_rows_kernel<32, SLQ_DIV_TRUE><<<(unsigned)ceil_div(n_jobs, 4), kBlock, 0, st>>>(
          w, K, rows, bits, n_jobs, write_back, codes, code_offsets, z, s32, status);
    else
      quantize_rows_kernel<32, SLQ_DIV_RECIP><<<(unsigned)ceil_div(n_jobs, 4), kBlock, 0, st>>>(
          w, K, rows, bits, n_jobs, write_back, codes, code_offsets, z, s32, status);
  } else if (fast_path_ok(w, K, 128)) {
    if (div_mode == SLQ_DIV_TRUE)
      quantize_rows_kernel<128, SLQ_DIV_TRUE><<<n_jobs, kBlock, 0, st>>>(
          w, K, rows, bits, n_jobs, write_back, codes, code_offsets, z, s32, status);
    else
      quantize_rows_kernel<128, SLQ_DIV_RECIP><<<n_jobs, kBlock, 0, st>>>(
          w, K, rows, bits, n_jobs, write_back, codes, code_offsets, z, s32, status);
  } else {
    SLQ_QLAUNCH(quantize_rows_generic_kernel, n_jobs);
  }
#undef SLQ_QLAUNCH
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_quantize_rows_host(float *w, int64_t n_rows, int64_t K, const int32_t *rows,
                                      const int32_t *bits, int32_t n_jobs, int32_t div_mode,
                                      int32_t write_back, uint8_t *codes,
                                      const int64_t *code_offsets, int32_t *z, float *s32,
                                      int32_t *status) {
  SLQ_CHECK_ARG(w && rows && bits && status, "slq_quantize_rows_host: null pointer argument");
  SLQ_CHECK_ARG(K > 0 && n_rows > 0 && n_jobs > 0, "slq_quantize_rows_host: empty problem");
  SLQ_CHECK_ARG(!codes || code_offsets, "slq_quantize_rows_host: codes without code_offsets");
  // Only the rows that are named by the job list travel: they are gathered into a compact
  // device tensor [n_jobs, K] (job j -> device row j).
  int64_t code_bytes = 0;
  if (codes)
    for (int j = 0; j < n_jobs; ++j) {
      SLQ_CHECK_ARG(bits[j] >= 1 && bits[j] <= 8, "slq_quantize_rows_host: bit %d", bits[j]);
      const int64_t end = code_offsets[j] + slq_packed_row_bytes(K, bits[j]);
      if (end > code_bytes) code_bytes = end;
    }
  cudaStream_t st = nullptr;
  SLQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  float *d_w = nullptr;
  int32_t *d_i32 = nullptr;  // rows | bits | z | status
  float *d_s = nullptr;
  int64_t *d_off = nullptr;
  uint8_t *d_codes = nullptr;
  int rc = SLQ_OK;
  auto fail = [&](cudaError_t e, const char *what) {
    set_error("slq_quantize_rows_host: %s: %s", what, cudaGetErrorString(e));
    rc = SLQ_ERR_CUDA;
  };
  cudaError_t e;
  int32_t *h_rows_dev = (int32_t *)malloc(sizeof(int32_t) * n_jobs);
  for (int j = 0; j < n_jobs; ++j) h_rows_dev[j] = j;
  do {
    if ((e = cudaMalloc(&d_w, sizeof(float) * n_jobs * K)) != cudaSuccess) { fail(e, "cudaMalloc w"); break; }
    if ((e = cudaMalloc(&d_i32, sizeof(int32_t) * 4 * n_jobs)) != cudaSuccess) { fail(e, "cudaMalloc meta"); break; }
    if ((e = cudaMalloc(&d_s, sizeof(float) * n_jobs)) != cudaSuccess) { fail(e, "cudaMalloc s"); break; }
    if (codes) {
      if ((e = cudaMalloc(&d_off, sizeof(int64_t) * n_jobs)) != cudaSuccess) { fail(e, "cudaMalloc off"); break; }
      if ((e = cudaMalloc(&d_codes, code_bytes)) != cudaSuccess) { fail(e, "cudaMalloc codes"); break; }
      cudaMemcpyAsync(d_off, code_offsets, sizeof(int64_t) * n_jobs, cudaMemcpyHostToDevice, st);
    }
    for (int j = 0; j < n_jobs; ++j)
      if (rows[j] < 0 || rows[j] >= n_rows) { set_error("slq_quantize_rows_host: row %d out of range", rows[j]); rc = SLQ_ERR_INVALID; break; }
    if (rc != SLQ_OK) break;
    // one copy per run of consecutive rows (a whole layer is a single copy)
    for (int j = 0; j < n_jobs;) {
      int e2 = j + 1;
      while (e2 < n_jobs && rows[e2] == rows[e2 - 1] + 1) ++e2;
      cudaMemcpyAsync(d_w + (int64_t)j * K, w + (int64_t)rows[j] * K, sizeof(float) * K * (e2 - j), cudaMemcpyHostToDevice, st);
      j = e2;
    }
    cudaMemcpyAsync(d_i32, h_rows_dev, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_i32 + n_jobs, bits, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, st);
    rc = slq_quantize_rows(d_w, n_jobs, K, d_i32, d_i32 + n_jobs, n_jobs, div_mode, write_back,
                           d_codes, d_off, d_i32 + 2 * n_jobs, d_s, d_i32 + 3 * n_jobs, st);
    if (rc != SLQ_OK) break;
    if (write_back)
      for (int j = 0; j < n_jobs;) {
        int e2 = j + 1;
        while (e2 < n_jobs && rows[e2] == rows[e2 - 1] + 1) ++e2;
        cudaMemcpyAsync(w + (int64_t)rows[j] * K, d_w + (int64_t)j * K, sizeof(float) * K * (e2 - j), cudaMemcpyDeviceToHost, st);
        j = e2;
      }
    if (codes) cudaMemcpyAsync(codes, d_codes, code_bytes, cudaMemcpyDeviceToHost, st);
    if (z) cudaMemcpyAsync(z, d_i32 + 2 * n_jobs, sizeof(int32_t) * n_jobs, cudaMemcpyDeviceToHost, st);
    if (s32) cudaMemcpyAsync(s32, d_s, sizeof(float) * n_jobs, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(status, d_i32 + 3 * n_jobs, sizeof(int32_t) * n_jobs, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { fail(e, "synchronize"); break; }
    for (int j = 0; j < n_jobs; ++j)
      if (status[j] & SLQ_ROW_ZERO_RANGE) {
        set_error("slq_quantize_rows_host: job %d (row %d) is constant: float division by zero", j, rows[j]);
        rc = SLQ_ERR_ZERO_RANGE;
        break;
      }
  } while (0);
  free(h_rows_dev);
  cudaFree(d_w);
  cudaFree(d_i32);
  cudaFree(d_s);
  cudaFree(d_off);
  cudaFree(d_codes);
  cudaStreamDestroy(st);
  return rc;
}

extern "C" int slq_classify_rows(const float *w, int64_t n_rows, int64_t K, int32_t *bit,
                                 int32_t *z, float *s, int32_t *exact, void *stream) {
  SLQ_CHECK_ARG(w && bit && z && s, "slq_classify_rows: null pointer argument");
  SLQ_CHECK_ARG(K > 0 && n_rows > 0, "slq_classify_rows: empty problem");
  cudaStream_t st = (cudaStream_t)stream;
  if (fast_path_ok(w, K, 32))
    classify_rows_kernel<32><<<(unsigned)ceil_div(n_rows, 4), kBlock, 0, st>>>(w, K, (int)n_rows, bit, z, s, exact);
  else if (fast_path_ok(w, K, 128))
    classify_rows_kernel<128><<<(unsigned)n_rows, kBlock, 0, st>>>(w, K, (int)n_rows, bit, z, s, exact);
  else {
    set_error("slq_classify_rows: K=%lld unsupported (need K %% 4 == 0, K <= 4608, 16B-aligned)", (long long)K);
    return SLQ_ERR_UNSUPPORTED;
  }
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_encode_rows(const float *w, int64_t n_rows, int64_t K, const int32_t *bit,
                               const int32_t *z, const float *s, uint8_t *codes,
                               const int64_t *code_offsets, void *stream) {
  SLQ_CHECK_ARG(w && bit && z && s && codes && code_offsets, "slq_encode_rows: null pointer argument");
  SLQ_CHECK_ARG(K > 0 && n_rows > 0, "slq_encode_rows: empty problem");
  cudaStream_t st = (cudaStream_t)stream;
  if (fast_path_ok(w, K, 32))
    encode_rows_kernel<32><<<(unsigned)ceil_div(n_rows, 4), kBlock, 0, st>>>(w, K, (int)n_rows, bit, z, s, codes, code_offsets);
  else if (fast_path_ok(w, K, 128))
    encode_rows_kernel<128><<<(unsigned)n_rows, kBlock, 0, st>>>(w, K, (int)n_rows, bit, z, s, codes, code_offsets);
  else {
    set_error("slq_encode_rows: K=%lld unsupported (need K %% 4 == 0, K <= 4608, 16B-aligned)", (long long)K);
    return SLQ_ERR_UNSUPPORTED;
  }
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}


extern "C" int slq_quantize_jobs(const slq_qjob *jobs, int32_t n_small, int32_t n_mid, int32_t n_large,
                                 int32_t div_mode, int32_t write_back, int32_t *z, float *s32,
                                 int32_t *status, void *stream) {
  SLQ_CHECK_ARG(jobs && status, "slq_quantize_jobs: null pointer argument");
  SLQ_CHECK_ARG(n_small >= 0 && n_mid >= 0 && n_large >= 0, "slq_quantize_jobs: negative job count");
  SLQ_CHECK_ARG(div_mode == SLQ_DIV_TRUE || div_mode == SLQ_DIV_RECIP, "slq_quantize_jobs: div_mode %d", div_mode);
  const int blocks_small = (int)ceil_div(n_small, kBlock / 8), blocks_mid = (int)ceil_div(n_mid, kBlock / 32);
  const int64_t grid = (int64_t)blocks_small + blocks_mid + n_large;
  if (grid == 0) return SLQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (div_mode == SLQ_DIV_TRUE)
    quantize_jobs_kernel<SLQ_DIV_TRUE><<<(unsigned)grid, kBlock, 0, st>>>(jobs, n_small, n_mid, n_large, blocks_small,
                                                                        blocks_mid, write_back, z, s32, status);
  else
    quantize_jobs_kernel<SLQ_DIV_RECIP><<<(unsigned)grid, kBlock, 0, st>>>(jobs, n_small, n_mid, n_large, blocks_small,
                                                                         blocks_mid, write_back, z, s32, status);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}

extern "C" int slq_decode_rows(const uint8_t *codes, const int64_t *code_offsets, const int32_t *bit,
                               const int32_t *z, const float *s, int64_t n_rows, int64_t K, float *w,
                               void *stream) {
  SLQ_CHECK_ARG(codes && code_offsets && bit && z && s && w, "slq_decode_rows: null pointer argument");
  SLQ_CHECK_ARG(K > 0 && n_rows > 0, "slq_decode_rows: empty problem");
  const int64_t total = n_rows * ((K + 3) / 4);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  decode_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(codes, code_offsets, bit, z, s, (int)n_rows, K, w);
  SLQ_LAUNCH_CHECK();
  return SLQ_OK;
}
