// csrc/common.cuh -- shared host/device helpers for libslq_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "slq.h"

namespace slq {

// thread-local last-error message behind slq_last_error()
void set_error(const char *fmt, ...);

#define SLQ_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      ::slq::set_error(__VA_ARGS__);  \
      return SLQ_ERR_INVALID;         \
    }                                 \
  } while (0)

#define SLQ_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::slq::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                        \
      return SLQ_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define SLQ_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      ::slq::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return SLQ_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

constexpr int kMaxDevices = 64;
int current_device();  // cudaGetDevice, 0 on error
int sm_count();  // multiprocessor count of the current device (148 on B200), cached per device

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace slq
