// Training-loss forward of SCFlowRefiner.loss (models/refiner/scflow_refiner.py:204-258) for the shipped configuration
// (configs/refine_models/scflow.py:75-104): ground-truth flow filtering (models/utils/flow.py:6-26), the sequence-weighted
// RAFT flow loss and L1 mask loss (models/loss/sequence_loss.py) and the disentangled point-matching pose loss
// (models/loss/point_matching_loss.py:160-218, loss_type 'l1', disentangle_z) - SURVEY.md §8 row a16 / §8(f) rank 2.
// All reductions are two-stage and deterministic (per-block partials combined in double precision in a fixed order).
// The ground-truth flow itself is scf_unproject (reference pose) + scf_reproject (ground-truth pose, invalid = max_flow).
#include "scf_common.cuh"

namespace scf {

// ---------------------------------------------------------------- filter_flow_by_mask (flow.py:6-26, warp.py:9-28)
// The coordinate arithmetic replays the reference's fp32 operation sequence (no FMA contraction) so that the sampled
// neighbours - and therefore the valid/invalid decision - match ATen's grid_sample(bilinear, zeros, align_corners=False).
__global__ void __launch_bounds__(256) filter_flow_kernel(float* __restrict__ flow, const float* __restrict__ mask, float invalid,
                                                          int B, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  const float dw = (float)(W - 1 > 1 ? W - 1 : 1), dh = (float)(H - 1 > 1 ? H - 1 : 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / HW, r = idx - b * HW;
    const int y = (int)(r / W), x = (int)(r - (long long)y * W);
    float* f = flow + b * 2 * HW + r;
    const float fx = f[0], fy = f[HW];
    bool bad = fx >= invalid && fy >= invalid;
    const float gx = __fadd_rn(__fdiv_rn(__fmul_rn(__fadd_rn((float)x, fx), 2.f), dw), -1.f);
    const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(__fadd_rn((float)y, fy), 2.f), dh), -1.f);
    const float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 2.f);
    const float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 2.f);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float tx = __fadd_rn(ix, -x0f), ty = __fadd_rn(iy, -y0f);     // ix - ix_nw, iy - iy_nw
    const float ux = __fadd_rn(__fadd_rn(x0f, 1.f), -ix), uy = __fadd_rn(__fadd_rn(y0f, 1.f), -iy);   // ix_se - ix, iy_se - iy
    const float* m = mask + b * HW;
    auto tap = [&](float xf, float yf) -> float {
      if (!(xf >= 0.f && xf <= (float)(W - 1) && yf >= 0.f && yf <= (float)(H - 1))) return 0.f;
      return m[(long long)(int)yf * W + (int)xf];
    };
    float v = __fmul_rn(tap(x0f, y0f), __fmul_rn(ux, uy));
    v = __fadd_rn(v, __fmul_rn(tap(x0f + 1.f, y0f), __fmul_rn(tx, uy)));
    v = __fadd_rn(v, __fmul_rn(tap(x0f, y0f + 1.f), __fmul_rn(ux, ty)));
    v = __fadd_rn(v, __fmul_rn(tap(x0f + 1.f, y0f + 1.f), __fmul_rn(tx, ty)));
    bad = bad || v < 0.9f;
    if (bad) { f[0] = invalid; f[HW] = invalid; }
  }
}

// ---------------------------------------------------------------- dense sequence losses
constexpr int LOSS_BLOCKS = 296;   // per iteration: two waves of 148 SMs
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
  __syncthreads();
  return s;   // valid in thread 0
}
// grid (LOSS_BLOCKS, iters). partial[(it*LOSS_BLOCKS + blk)*3 + {0: sum valid*|pred-gt|, 1: sum |mask-occ|, 2: sum valid}]
__global__ void __launch_bounds__(256) seq_loss_partial_kernel(const float* __restrict__ flow_pred, const float* __restrict__ mask_pred,
                                                               const float* __restrict__ gt_flow, const float* __restrict__ valid,
                                                               float max_flow, int B, long long HW, double* __restrict__ partial) {
  __shared__ double sh[8];
  const int it = blockIdx.y;
  const long long total = (long long)B * HW;
  const float* fp = flow_pred + (long long)it * total * 2;
  const float* mp = mask_pred + (long long)it * total;
  float sf = 0.f, sm = 0.f, sv = 0.f;
  double df = 0., dm = 0., dv = 0.;
  int n = 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / HW, r = idx - b * HW;
    const float gx = __ldg(gt_flow + b * 2 * HW + r), gy = __ldg(gt_flow + b * 2 * HW + HW + r);
    const float mag = sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    const float v = (__ldg(valid + idx) >= 0.5f && mag < max_flow) ? 1.f : 0.f;
    const float px = __ldg(fp + b * 2 * HW + r), py = __ldg(fp + b * 2 * HW + HW + r);
    sf += v * (fabsf(px - gx) + fabsf(py - gy));
    const float occ = __fadd_rn(gx, gy) < max_flow ? 1.f : 0.f;
    sm += fabsf(__ldg(mp + idx) - occ);
    sv += v;
    if (++n == 64) { df += sf; dm += sm; dv += sv; sf = sm = sv = 0.f; n = 0; }   // bound the fp32 partial length
  }
  df += sf; dm += sm; dv += sv;
  const double a = block_sum(df, sh), b2 = block_sum(dm, sh), c = block_sum(dv, sh);
  if (threadIdx.x == 0) {
    double* o = partial + ((long long)it * gridDim.x + blockIdx.x) * 3;
    o[0] = a; o[1] = b2; o[2] = c;
  }
}

// ---------------------------------------------------------------- disentangled point-matching loss, one block per (sample, iter)
struct PmParams {
  const float* rot; const float* trs; const float* gt_rot; const float* gt_trs; const long long* label;
  const float* points; const int* num_points; const unsigned char* symmetric; const float* diameter;
  int B, max_points, num_class;
  float* per_sample;   // [iters, B]
};
__global__ void __launch_bounds__(256) point_matching_kernel(const PmParams p) {
  extern __shared__ float pred_pts[];   // symmetric classes: the predicted-rotation points [np][3]
  __shared__ double sh[8];
  const int b = blockIdx.x, it = blockIdx.y;
  long long c = p.label[b];
  if (c < 0) c = 0;
  if (c >= p.num_class) c = p.num_class - 1;
  const int np = p.num_points[c];
  const float* pts = p.points + (long long)c * p.max_points * 3;
  float Rp[9], Rg[9], tg[3], tp[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) { Rp[i] = p.rot[((long long)it * p.B + b) * 9 + i]; Rg[i] = p.gt_rot[(long long)b * 9 + i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { tp[i] = p.trs[((long long)it * p.B + b) * 3 + i]; tg[i] = p.gt_trs[(long long)b * 3 + i]; }
  const bool sym = p.symmetric[c] != 0;
  if (sym) {
    for (int k = threadIdx.x; k < np; k += blockDim.x) {
      const float x = pts[k * 3], y = pts[k * 3 + 1], z = pts[k * 3 + 2];
      pred_pts[k * 3] = Rp[0] * x + Rp[1] * y + Rp[2] * z + tg[0];
      pred_pts[k * 3 + 1] = Rp[3] * x + Rp[4] * y + Rp[5] * z + tg[1];
      pred_pts[k * 3 + 2] = Rp[6] * x + Rp[7] * y + Rp[8] * z + tg[2];
    }
    __syncthreads();
  }
  double acc = 0.;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const float x = pts[j * 3], y = pts[j * 3 + 1], z = pts[j * 3 + 2];
    const float gx = Rg[0] * x + Rg[1] * y + Rg[2] * z + tg[0];
    const float gy = Rg[3] * x + Rg[4] * y + Rg[5] * z + tg[1];
    const float gz = Rg[6] * x + Rg[7] * y + Rg[8] * z + tg[2];
    float qx, qy, qz;
    if (sym) {                       // nearest predicted point (squared L2, first minimum wins)
      float best = 3.4e38f;
      int bi = 0;
      for (int k = 0; k < np; ++k) {
        const float dx = gx - pred_pts[k * 3], dy = gy - pred_pts[k * 3 + 1], dz = gz - pred_pts[k * 3 + 2];
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d < best) { best = d; bi = k; }
      }
      qx = pred_pts[bi * 3]; qy = pred_pts[bi * 3 + 1]; qz = pred_pts[bi * 3 + 2];
    } else {
      qx = Rp[0] * x + Rp[1] * y + Rp[2] * z + tg[0];
      qy = Rp[3] * x + Rp[4] * y + Rp[5] * z + tg[1];
      qz = Rp[6] * x + Rp[7] * y + Rp[8] * z + tg[2];
    }
    acc += (double)(fabsf(qx - gx) + fabsf(qy - gy) + fabsf(qz - gz));
  }
  const double s = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    const float l_rot = (float)(s / (np > 0 ? np : 1));
    // translation part with disentangle_z: every model point moves by the same vector, so the point means reduce to
    // |dz| (depth term) + |dx| + |dy| (xy term)                                   (point_matching_loss.py:196-207)
    const float l_trans = fabsf(tp[2] - tg[2]) + (fabsf(tp[0] - tg[0]) + fabsf(tp[1] - tg[1]));
    p.per_sample[(long long)it * p.B + b] = (l_trans + l_rot) / p.diameter[c];
  }
}

// ---------------------------------------------------------------- finalize: sequence weights, loss weights
// out[0..3] = loss, loss_pose, loss_flow, loss_mask ; out[4 + i], out[4 + iters + i], out[4 + 2*iters + i] = weighted pose /
// flow / mask loss of iteration i (the reference's seq_*_loss_list)
__global__ void loss_finalize_kernel(const float* __restrict__ per_sample, const double* __restrict__ partial, int iters, int B,
                                     long long HW, float gamma, float w_flow, float w_pose, float w_mask, float eps,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double lp = 0., lf = 0., lm = 0.;
  for (int it = 0; it < iters; ++it) {
    double sp = 0.;
    for (int b = 0; b < B; ++b) sp += per_sample[(long long)it * B + b];
    double sf = 0., sm = 0., sv = 0.;
    for (int k = 0; k < LOSS_BLOCKS; ++k) {
      const double* q = partial + ((long long)it * LOSS_BLOCKS + k) * 3;
      sf += q[0]; sm += q[1]; sv += q[2];
    }
    const float pose_i = w_pose * (float)(sp / B);
    const float flow_i = w_flow * (float)sf / ((float)sv + eps);
    const float mask_i = w_mask * (float)(sm / ((double)B * HW));
    out[4 + it] = pose_i; out[4 + iters + it] = flow_i; out[4 + 2 * iters + it] = mask_i;
    const double wgt = pow((double)gamma, (double)(iters - it - 1));
    lp += wgt * pose_i; lf += wgt * flow_i; lm += wgt * mask_i;
  }
  out[1] = (float)lp; out[2] = (float)lf; out[3] = (float)lm;
  out[0] = (float)lp + (float)lf + (float)lm;
}

}  // namespace scf

using namespace scf;

extern "C" {

int scf_filter_flow_by_mask(float* flow, const float* gt_mask, float invalid, int B, int H, int W, void* stream) {
  SCF_REQUIRE(flow && gt_mask && B > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_filter_flow_by_mask: bad args");
  const long long total = (long long)B * H * W;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  filter_flow_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(flow, gt_mask, invalid, B, H, W);
  return check_launch("filter_flow_kernel");
}

size_t scf_refiner_loss_scratch_bytes(int iters, int B) {
  if (iters <= 0 || B <= 0) return 0;
  return align_up((size_t)iters * LOSS_BLOCKS * 3 * sizeof(double), 256) + align_up((size_t)iters * B * sizeof(float), 256);
}

int scf_refiner_loss(const scf_loss_desc* d, void* stream) {
  SCF_REQUIRE(d != nullptr, SCF_ERR_ARG, "scf_refiner_loss: null descriptor");
  SCF_REQUIRE(d->flow_pred && d->mask_pred && d->rotation && d->translation && d->gt_flow && d->valid && d->gt_rotation &&
                  d->gt_translation && d->label && d->points && d->num_points && d->symmetric && d->diameter && d->scratch && d->out,
              SCF_ERR_ARG, "scf_refiner_loss: null pointer");
  SCF_REQUIRE(d->iters > 0 && d->B > 0 && d->H > 0 && d->W > 0 && d->num_class > 0 && d->max_points > 0, SCF_ERR_ARG,
              "scf_refiner_loss: bad sizes");
  SCF_REQUIRE(d->scratch_bytes >= scf_refiner_loss_scratch_bytes(d->iters, d->B) && reinterpret_cast<uintptr_t>(d->scratch) % 256 == 0,
              SCF_ERR_ARG, "scf_refiner_loss: scratch too small or misaligned");
  SCF_REQUIRE((size_t)d->max_points * 12 <= 200 * 1024, SCF_ERR_UNSUPPORTED, "scf_refiner_loss: more than 17066 model points per class");
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = reinterpret_cast<double*>(d->scratch);
  float* per_sample = reinterpret_cast<float*>(reinterpret_cast<char*>(d->scratch) +
                                                align_up((size_t)d->iters * LOSS_BLOCKS * 3 * sizeof(double), 256));
  const long long HW = (long long)d->H * d->W;
  seq_loss_partial_kernel<<<dim3(LOSS_BLOCKS, d->iters), 256, 0, st>>>(d->flow_pred, d->mask_pred, d->gt_flow, d->valid, d->max_flow,
                                                                       d->B, HW, partial);
  SCF_TRY(check_launch("seq_loss_partial_kernel"));
  PmParams p;
  p.rot = d->rotation; p.trs = d->translation; p.gt_rot = d->gt_rotation; p.gt_trs = d->gt_translation;
  p.label = reinterpret_cast<const long long*>(d->label);
  p.points = d->points; p.num_points = d->num_points; p.symmetric = d->symmetric; p.diameter = d->diameter;
  p.B = d->B; p.max_points = d->max_points; p.num_class = d->num_class; p.per_sample = per_sample;
  const size_t smem = (size_t)d->max_points * 12;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    SCF_CUDA(cudaFuncSetAttribute(point_matching_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  point_matching_kernel<<<dim3(d->B, d->iters), 256, smem, st>>>(p);
  SCF_TRY(check_launch("point_matching_kernel"));
  loss_finalize_kernel<<<1, 32, 0, st>>>(per_sample, partial, d->iters, d->B, HW, d->gamma, d->w_flow, d->w_pose, d->w_mask, d->eps, d->out);
  return check_launch("loss_finalize_kernel");
}

}  // extern "C"
