// Pose-regressor tail (models/head/pose_head.py:201-211): GroupNorm+ReLU on NHWC maps, the FC chain and the
// class-selected rotation/translation projection. All tiny, latency-bound; weights stay L2-resident.
#include "scf_common.cuh"

namespace scf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per (sample, group); NHWC [B,HW,C]; two-pass mean / biased variance like torch.group_norm
__global__ void __launch_bounds__(256) group_norm_relu_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int B, int HW, int C,
                                                              int G, float eps) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= B * G) return;
  const int b = wid / G, g = wid - b * G;
  const int cpg = C / G;
  float* base = x + (long long)b * HW * C + g * cpg;
  const int n = HW * cpg;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += base[(long long)(i / cpg) * C + (i % cpg)];
  const float mean = warp_sum(s) / (float)n;
  float v = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float d = base[(long long)(i / cpg) * C + (i % cpg)] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)n + eps);
  for (int i = lane; i < n; i += 32) {
    const int c = i % cpg;
    float* p = base + (long long)(i / cpg) * C + c;
    const float y = (*p - mean) * rstd * gamma[g * cpg + c] + beta[g * cpg + c];
    *p = fmaxf(y, 0.f);
  }
}

// y[b,o] = act(W[o,:] . x[b,:] + bias[o]); one warp per output row, 8 samples per sweep of the row
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, int B, int I,
                                                     int O, int act) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  const float* wr = w + (long long)o * I;
  const float bo = bias ? bias[o] : 0.f;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int i = lane * 4; i < I; i += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (b0 + j < B) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(b0 + j) * I + i));
          acc[j] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[j]))));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = warp_sum(acc[j]);
      if (lane == 0 && b0 + j < B) y[(long long)(b0 + j) * O + o] = act_apply(s + bo, act);
    }
  }
}

// rotation/translation projection of the class given by label[0] (the reference's index_select(...)[:, 0] quirk,
// pose_head.py:209-210). One warp per (sample, output row).
__global__ void __launch_bounds__(256) pose_project_kernel(const float* __restrict__ x, const float* __restrict__ rot_w,
                                                           const float* __restrict__ rot_b, const float* __restrict__ tr_w,
                                                           const float* __restrict__ tr_b, const int64_t* __restrict__ label,
                                                           float* __restrict__ d_rot, float* __restrict__ d_trs, int B, int I,
                                                           int rot_dim, int num_class) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int rows = rot_dim + 3;
  if (wid >= B * rows) return;
  const int b = wid / rows, r = wid - b * rows;
  long long cls = 0;
  if (num_class > 0) {
    cls = label[0];
    if (cls < 0) cls = 0;
    if (cls >= num_class) cls = num_class - 1;
  }
  const float* wr;
  float bo;
  if (r < rot_dim) { wr = rot_w + (cls * rot_dim + r) * I; bo = rot_b[cls * rot_dim + r]; }
  else { wr = tr_w + (cls * 3 + (r - rot_dim)) * I; bo = tr_b[cls * 3 + (r - rot_dim)]; }
  float acc = 0.f;
  for (int i = lane; i < I; i += 32) acc = fmaf(wr[i], x[(long long)b * I + i], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    if (r < rot_dim) d_rot[b * rot_dim + r] = acc + bo;
    else d_trs[b * 3 + (r - rot_dim)] = acc + bo;
  }
}

}  // namespace scf

extern "C" {

int scf_group_norm_relu(float* x, const float* gamma, const float* beta, int B, int HW, int C, int num_groups, float eps,
                        void* stream) {
  SCF_REQUIRE(x && gamma && beta && B > 0 && HW > 0 && C > 0 && num_groups > 0 && C % num_groups == 0, SCF_ERR_ARG,
              "scf_group_norm_relu: bad args");
  const int warps = B * num_groups;
  scf::group_norm_relu_kernel<<<scf::cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, B, HW, C, num_groups,
                                                                                    eps);
  return scf::check_launch("group_norm_relu_kernel");
}

int scf_linear(const float* x, const float* w, const float* bias, float* y, int B, int I, int O, int act, void* stream) {
  SCF_REQUIRE(x && w && y && B > 0 && I > 0 && O > 0, SCF_ERR_ARG, "scf_linear: bad args");
  SCF_REQUIRE(I % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0,
              SCF_ERR_ALIGN, "scf_linear: I must be a multiple of 4 and x/w 16B aligned");
  scf::linear_kernel<<<scf::cdiv(O, 8), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, I, O, act);
  return scf::check_launch("linear_kernel");
}

int scf_pose_project(const float* x, const float* rot_w, const float* rot_b, const float* tr_w, const float* tr_b,
                     const int64_t* label, float* d_rot, float* d_trs, int B, int I, int rot_dim, int num_class,
                     void* stream) {
  SCF_REQUIRE(x && rot_w && rot_b && tr_w && tr_b && d_rot && d_trs && B > 0 && I > 0 && rot_dim > 0, SCF_ERR_ARG,
              "scf_pose_project: bad args");
  SCF_REQUIRE(num_class <= 0 || label != nullptr, SCF_ERR_ARG, "scf_pose_project: multi-class head needs label");
  const int warps = B * (rot_dim + 3);
  scf::pose_project_kernel<<<scf::cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(x, rot_w, rot_b, tr_w, tr_b, label,
                                                                                 d_rot, d_trs, B, I, rot_dim, num_class);
  return scf::check_launch("pose_project_kernel");
}

}  // extern "C"
