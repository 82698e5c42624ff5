// Shared helpers for libscflow_sm100a.so (error plumbing, launch checks). sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/scflow_b200.h"

namespace scf {

void set_error(const char* fmt, ...);
// counts kernel launches issued by the calling thread (bench.py's gpu_launches)
extern thread_local long long g_launches;

inline int check_launch(const char* what) {
  g_launches++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define SCF_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      scf::set_error(__VA_ARGS__);          \
      return (code);                        \
    }                                       \
  } while (0)

#define SCF_TRY(expr)                       \
  do {                                      \
    int _rc = (expr);                       \
    if (_rc != 0) return _rc;               \
  } while (0)

#define SCF_CUDA(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) {                                             \
      scf::set_error("%s: %s", #expr, cudaGetErrorString(_e));           \
      return (int)_e;                                                    \
    }                                                                    \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case SCF_ACT_RELU: return fmaxf(v, 0.f);
    case SCF_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case SCF_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

}  // namespace scf
