// Optimizer step of the training path (BASELINE config 5): global gradient norm -> clip coefficient -> AdamW, over the FLAT
// parameter / gradient / moment buffers the Python Trainer keeps (scflow_b200/training.py).
//   reference: mmcv OptimizerHook(grad_clip=dict(max_norm=10.)) = torch.nn.utils.clip_grad_norm_  followed by
//              torch.optim.AdamW(lr, betas, eps, weight_decay)            (configs/refine_models/scflow.py:117-125)
// Two launches, HBM-bound (20 B read + 12 B written per parameter):
//   1. clip_norm_partial_kernel: per-block sums of g^2 in a fixed order (deterministic), grid = SM multiple;
//   2. clip_adamw_kernel: every block first folds the partial sums in the same fixed order (so all blocks agree bit for bit on
//      the norm - no third launch, no atomics), then updates its slice.  `gscale` (1 / world size) turns the all-reduced SUM into
//      the DDP mean before the norm is taken.
#include "scf_common.cuh"

namespace scf {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) clip_norm_partial_kernel(const float* __restrict__ g, long long n, float gscale,
                                                                        float* __restrict__ partial) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float a = v.x * gscale, b = v.y * gscale, c = v.z * gscale, d = v.w * gscale;
    acc += (a * a + b * b) + (c * c + d * d);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                 float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                                 float wd, float bc1, float bc2_sqrt, float max_norm, float gscale,
                                                                 const float* __restrict__ partial, int nparts, float* __restrict__ stats) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < nparts; ++i) tot += partial[i];                    // same order in every block
    const float norm = sqrtf(tot);
    // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.f) : 1.f;
    s_coef = coef * gscale;
    if (blockIdx.x == 0) { stats[0] = norm; stats[1] = coef; }
  }
  __syncthreads();
  const float gs = s_coef;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * gs;
      // torch.optim.AdamW (single tensor): p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
      //                                   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
      float pv = P[j] * (1.f - lr * wd);
      M[j] = b1 * M[j] + (1.f - b1) * gr;
      V[j] = b2 * V[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[j]) / bc2_sqrt + eps;
      P[j] = pv - (lr / bc1) * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace scf

extern "C" {

int scf_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float max_norm, float grad_scale, float* scratch,
                   int scratch_floats, float* stats2, void* stream) {
  using namespace scf;
  SCF_REQUIRE(params && grads && exp_avg && exp_avg_sq && scratch && stats2, SCF_ERR_ARG, "scf_clip_adamw: null pointer");
  SCF_REQUIRE(n > 0 && n % 4 == 0 && step >= 1, SCF_ERR_ARG, "scf_clip_adamw: n must be a positive multiple of 4, step >= 1");
  for (const void* q : {(const void*)params, (const void*)grads, (const void*)exp_avg, (const void*)exp_avg_sq})
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(q) % 16 == 0, SCF_ERR_ALIGN, "scf_clip_adamw: buffers must be 16B aligned");
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  long long want = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  int blocks = (int)(want < (long long)num_sms * 4 ? want : (long long)num_sms * 4);
  if (blocks > scratch_floats) blocks = scratch_floats;
  SCF_REQUIRE(blocks >= 1, SCF_ERR_ARG, "scf_clip_adamw: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  clip_norm_partial_kernel<<<blocks, OPT_THREADS, 0, st>>>(grads, n, grad_scale, scratch);
  SCF_TRY(check_launch("clip_norm_partial_kernel"));
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  clip_adamw_kernel<<<blocks, OPT_THREADS, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                    sqrtf(bc2), max_norm, grad_scale, scratch, blocks, stats2);
  return check_launch("clip_adamw_kernel");
}

}  // extern "C"
