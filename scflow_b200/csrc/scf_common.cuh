// Shared helpers for libscflow_sm100a.so (error plumbing, launch checks). sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <utility>
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

// Launch with the programmatic-dependent-launch attribute when SCFLOW_PDL_SMALL=1: the kernel may then be scheduled while its
// predecessor on the stream is still running and MUST execute scf_pdl_enter() (griddepcontrol.launch_dependents + .wait) before
// its first global-memory access.  Tried on the small kernels of the refinement loop (lookup, split copy, predict gather, x-fold,
// GroupNorm, pose projection / update, re-projection) to hide one launch latency each: measured 9.08 - 9.14 ms per step against
// 8.99 ms with plain launches (the early-scheduled blocks only spin in griddepcontrol.wait next to the predecessor's), so the
// default is a plain launch; the tensor-core kernels keep their own PDL launches (SCFLOW_PDL), where the prologue is worth hiding.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL_SMALL"); return e ? atoi(e) != 0 : false; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// first statements of a kernel launched through launch_pdl (no-ops under a plain launch)
__device__ __forceinline__ void scf_pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

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
