// Pose-regressor tail (models/head/pose_head.py:201-211): GroupNorm+ReLU on NHWC maps, the FC chain and the
// class-selected rotation/translation projection. All tiny, latency-bound; weights stay L2-resident.
#include "scf_common.cuh"
#include "scf_tc.cuh"

namespace scf {

int group_norm_relu_partials(float* x, int nsplit, long long split_stride, const float* gamma, const float* beta, int B, int HW, int C,
                             int num_groups, float eps, void* out_hl, long long plane_stride, cudaStream_t stream);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[j] = this lane's partial sum of sample j (j = 0..31).  Transpose-reduce: after the call lane l holds (return value) the
// sum over all lanes of sample l - 31 shuffles (16 + 8 + 4 + 2 + 1) instead of 32 full butterflies (160).
__device__ __forceinline__ float transpose_reduce32(float (&acc)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = up ? acc[j] : acc[j + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      acc[j] = (up ? acc[j + off] : acc[j]) + recv;
    }
  }
  return acc[0];
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

// Fast path for C == 4*G (one float4 = one group; the pose head: C=128, G=32): grid (B, 4), a 256-thread block owns 8
// groups of one sample: thread t = (group t%8, pixel lane t/8), pixels lane + 32j -> 128 B contiguous per pixel; the
// sample's values stay in registers between the three steps (mean, variance about the mean, normalise), so the map is
// read once.  Optional split-bf16 copy of the result for a following tensor-core convolution.  HW <= 32*GN_MAXP.
constexpr int GN_MAXP = 8;
// ``nsplit`` > 1: x holds nsplit partial maps ``split_stride`` floats apart (split-K convolution); they are added on load
// and the result is written to the first one.
__global__ void __launch_bounds__(256) group_norm_relu_c4_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int HW, float eps,
                                                                 __nv_bfloat16* __restrict__ out_hl, long long plane, int nsplit,
                                                                 long long split_stride) {
  scf_pdl_enter();
  __shared__ float red[32][9];
  __shared__ float stat[8];
  const int b = blockIdx.x, gl = threadIdx.x & 7, r = threadIdx.x >> 3;
  const int g = blockIdx.y * 8 + gl;
  const int C = 128;
  float4* base = reinterpret_cast<float4*>(x + (long long)b * HW * C) + g;
  float4 v[GN_MAXP];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < GN_MAXP; ++j) {
    const int p = r + 32 * j;
    v[j] = p < HW ? base[(long long)p * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < HW)
      for (int k = 1; k < nsplit; ++k) {
        const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base + (long long)p * 32) + k * split_stride);
        v[j].x += u.x; v[j].y += u.y; v[j].z += u.z; v[j].w += u.w;
      }
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  red[r][gl] = s;
  __syncthreads();
  if (r == 0) { float t = 0.f; for (int i = 0; i < 32; ++i) t += red[i][gl]; stat[gl] = t / (float)(HW * 4); }
  __syncthreads();
  const float mean = stat[gl];
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < GN_MAXP; ++j) {
    if (r + 32 * j < HW) {
      const float a = v[j].x - mean, bq = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + bq * bq) + (c * c + d * d);
    }
  }
  __syncthreads();
  red[r][gl] = q;
  __syncthreads();
  if (r == 0) { float t = 0.f; for (int i = 0; i < 32; ++i) t += red[i][gl]; stat[gl] = rsqrtf(t / (float)(HW * 4) + eps); }
  __syncthreads();
  const float rstd = stat[gl];
  const float4 ga = reinterpret_cast<const float4*>(gamma)[g], be = reinterpret_cast<const float4*>(beta)[g];
#pragma unroll
  for (int j = 0; j < GN_MAXP; ++j) {
    const int p = r + 32 * j;
    if (p >= HW) continue;
    float4 o4;
    o4.x = fmaxf((v[j].x - mean) * rstd * ga.x + be.x, 0.f);
    o4.y = fmaxf((v[j].y - mean) * rstd * ga.y + be.y, 0.f);
    o4.z = fmaxf((v[j].z - mean) * rstd * ga.z + be.z, 0.f);
    o4.w = fmaxf((v[j].w - mean) * rstd * ga.w + be.w, 0.f);
    base[(long long)p * 32] = o4;
    if (out_hl) {
      __nv_bfloat16 hi[4], lo[4];
      tc::split_bf16(o4.x, hi[0], lo[0]); tc::split_bf16(o4.y, hi[1], lo[1]);
      tc::split_bf16(o4.z, hi[2], lo[2]); tc::split_bf16(o4.w, hi[3], lo[3]);
      __nv_bfloat16* o = out_hl + ((long long)b * HW + p) * C + g * 4;
      *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(o + plane) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

// y[b,o] = act(W[o,:] . x[b,:] + bias[o]).  Block = 8 warps = 8 output rows; x is staged through shared memory in
// [32 samples x 256] chunks (coalesced, read once per block); each lane keeps 32 per-sample partial sums and the
// weight row is read exactly once per batch chunk. Replaces a first version whose warps each re-read all of x.
constexpr int LIN_KC = 256;
__global__ void __launch_bounds__(256) linear_smem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y, int B, int I,
                                                          int O, int act) {
  __shared__ __align__(16) float xs[32][LIN_KC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int o = blockIdx.x * 8 + wid;
  const float* wr = w + (long long)(o < O ? o : O - 1) * I;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int nb = B - b0 < 32 ? B - b0 : 32;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    // the warp's weight row is fetched up front (I <= 2048: 16 float4 per lane) so its latency overlaps the x staging
    float4 wreg[16];
    const bool wpre = I <= 2048;
    if (wpre) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int k = u * 128 + lane * 4;
        wreg[u] = k < I ? __ldg(reinterpret_cast<const float4*>(wr + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int k0 = 0; k0 < I; k0 += LIN_KC) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < 32 * (LIN_KC / 4); idx += 256) {
        const int bb = idx / (LIN_KC / 4), c4 = idx - bb * (LIN_KC / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bb < nb && k0 + c4 * 4 < I) v = __ldg(reinterpret_cast<const float4*>(x + (long long)(b0 + bb) * I + k0) + c4);
        reinterpret_cast<float4*>(&xs[bb][0])[c4] = v;
      }
      __syncthreads();
#pragma unroll
      for (int h = 0; h < LIN_KC / 128; ++h) {
        const int k = h * 128 + lane * 4;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wpre) {
          const int u = (k0 >> 7) + h;           // k0 / 128 + h  (LIN_KC = 256 -> two register slots per chunk)
#pragma unroll
          for (int uu = 0; uu < 16; ++uu) if (uu == u) wv = wreg[uu];
        } else if (k0 + k < I) {
          wv = __ldg(reinterpret_cast<const float4*>(wr + k0 + k));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[j][k]);
          acc[j] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[j]))));
        }
      }
    }
    // lane j ends up with the full sum of sample j
    const float mine = transpose_reduce32(acc, lane);
    if (o < O && lane < nb) y[(long long)(b0 + lane) * O + o] = act_apply(mine + (bias ? bias[o] : 0.f), act);
  }
}

// Split-K form of the FC layers for small batches: with B <= 32 samples an FC layer is a weight stream (fc0: 8 MB) that a
// handful of blocks cannot pull from L2 fast enough, and one block's serial sweep over I = 2048 is pure latency.  grid =
// (O / 8, KS): block (ob, ks) forms the partial dot products of 8 output rows over input columns [ks * I / KS, ...) and writes
// them raw to part[ks][b][o]; bias and activation of THIS layer are applied by whoever reads the partials next (the next layer's
// x staging below, or pose_project) - `xin` describes that for this layer's own input: x = act(sum_s xin.part[s] + xin.bias).
struct LinIn { const float* p; int nsplit; long long split_stride; const float* bias; int relu; };

__device__ __forceinline__ float4 lin_load4(const LinIn& in, long long off, int col) {
  float4 v = __ldg(reinterpret_cast<const float4*>(in.p + off));
  for (int s = 1; s < in.nsplit; ++s) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(in.p + s * in.split_stride + off));
    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
  }
  if (in.bias) {
    const float4 bq = __ldg(reinterpret_cast<const float4*>(in.bias + col));
    v.x += bq.x; v.y += bq.y; v.z += bq.z; v.w += bq.w;
  }
  if (in.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}

__global__ void __launch_bounds__(256) linear_splitk_kernel(const LinIn xin, const float* __restrict__ w, float* __restrict__ part,
                                                            int B, int I, int O, int krange) {
  __shared__ __align__(16) float xs[32][LIN_KC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int o = blockIdx.x * 8 + wid, ks = blockIdx.y;
  const int kbeg = ks * krange, kend = kbeg + krange < I ? kbeg + krange : I;
  const float* wr = w + (long long)(o < O ? o : O - 1) * I;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += LIN_KC) {
    // this chunk's weights first (their latency overlaps the x staging)
    float4 wv[LIN_KC / 128];
#pragma unroll
    for (int h = 0; h < LIN_KC / 128; ++h) {
      const int k = k0 + h * 128 + lane * 4;
      wv[h] = k < kend ? __ldg(reinterpret_cast<const float4*>(wr + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * (LIN_KC / 4); idx += 256) {
      const int bb = idx / (LIN_KC / 4), c4 = idx - bb * (LIN_KC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bb < B && k0 + c4 * 4 < kend) v = lin_load4(xin, (long long)bb * I + k0 + c4 * 4, k0 + c4 * 4);
      reinterpret_cast<float4*>(&xs[bb][0])[c4] = v;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < LIN_KC / 128; ++h) {
      const int k = h * 128 + lane * 4;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[j][k]);
        acc[j] = fmaf(wv[h].x, xv.x, fmaf(wv[h].y, xv.y, fmaf(wv[h].z, xv.z, fmaf(wv[h].w, xv.w, acc[j]))));
      }
    }
  }
  const float mine = transpose_reduce32(acc, lane);
  if (o < O && lane < B) part[((long long)ks * B + lane) * O + o] = mine;
}

// pose_project over split-K partials of the last FC layer: x = relu(sum_s part[s] + bias)
__global__ void __launch_bounds__(256) pose_project_partials_kernel(const LinIn xin, const float* __restrict__ rot_w,
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
  for (int i = lane * 4; i < I; i += 128) {
    const float4 xv = lin_load4(xin, (long long)b * I + i, i);
    const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
    acc = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc))));
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (r < rot_dim) d_rot[b * rot_dim + r] = acc + bo;
    else d_trs[b * 3 + (r - rot_dim)] = acc + bo;
  }
}

// class-selected projection over the split-K partial maps of the last FC layer (x = relu(sum_s part[s] + bias))
int pose_project_partials(const float* part, int nsplit, long long split_stride, const float* bias, const float* rot_w, const float* rot_b,
                          const float* tr_w, const float* tr_b, const int64_t* label, float* d_rot, float* d_trs, int B, int I,
                          int rot_dim, int num_class, cudaStream_t st) {
  LinIn in = {part, nsplit, split_stride, bias, 1};
  const int warps = B * (rot_dim + 3);
  pose_project_partials_kernel<<<cdiv(warps, 8), 256, 0, st>>>(in, rot_w, rot_b, tr_w, tr_b, label, d_rot, d_trs, B, I, rot_dim, num_class);
  return check_launch("pose_project_partials_kernel");
}

// The pose head's FC tail for B <= 32 (pose_head.py:203-210): fc0 (relu) -> fc1 (relu) -> class-selected projection as three
// launches over split-K partial sums.  part0: [ks0][B][O0], part1: [ks1][B][O1] scratch.
int pose_fc_tail(const float* x, const float* w0, const float* b0, int I0, int O0, const float* w1, const float* b1, int O1,
                 const float* rot_w, const float* rot_b, const float* tr_w, const float* tr_b, const int64_t* label, float* d_rot,
                 float* d_trs, int B, int rot_dim, int num_class, float* part0, float* part1, int ks0, int ks1, cudaStream_t st) {
  SCF_REQUIRE(B >= 1 && B <= 32 && I0 % (ks0 * 4) == 0 && O0 % (ks1 * 4) == 0 && O0 % 4 == 0 && O1 % 4 == 0, SCF_ERR_ARG, "pose_fc_tail: bad shape");
  LinIn in0 = {x, 1, 0, nullptr, 0};
  linear_splitk_kernel<<<dim3(cdiv(O0, 8), ks0), 256, 0, st>>>(in0, w0, part0, B, I0, O0, I0 / ks0);
  SCF_TRY(check_launch("linear_splitk_kernel"));
  LinIn in1 = {part0, ks0, (long long)B * O0, b0, 1};
  linear_splitk_kernel<<<dim3(cdiv(O1, 8), ks1), 256, 0, st>>>(in1, w1, part1, B, O0, O1, O0 / ks1);
  SCF_TRY(check_launch("linear_splitk_kernel"));
  LinIn in2 = {part1, ks1, (long long)B * O1, b1, 1};
  const int warps = B * (rot_dim + 3);
  pose_project_partials_kernel<<<cdiv(warps, 8), 256, 0, st>>>(in2, rot_w, rot_b, tr_w, tr_b, label, d_rot, d_trs, B, O1, rot_dim, num_class);
  return check_launch("pose_project_partials_kernel");
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
  scf_pdl_enter();
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

int scf_group_norm_relu_split(float* x, const float* gamma, const float* beta, int B, int HW, int C, int num_groups, float eps,
                              void* out_hl, long long plane_stride, void* stream) {
  return scf::group_norm_relu_partials(x, 1, 0, gamma, beta, B, HW, C, num_groups, eps, out_hl, plane_stride, (cudaStream_t)stream);
}
}  // extern "C"

namespace scf {
int group_norm_relu_partials(float* x, int nsplit, long long split_stride, const float* gamma, const float* beta, int B, int HW, int C,
                             int num_groups, float eps, void* out_hl, long long plane_stride, cudaStream_t stream) {
  SCF_REQUIRE(x && gamma && beta && B > 0 && HW > 0 && nsplit >= 1 && split_stride % 4 == 0, SCF_ERR_ARG, "scf_group_norm_relu_split: bad args");
  SCF_REQUIRE(C == 128 && num_groups == 32, SCF_ERR_UNSUPPORTED, "scf_group_norm_relu_split: C=128, 32 groups only");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(gamma) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(beta) % 16 == 0, SCF_ERR_ALIGN, "scf_group_norm_relu_split: 16B alignment required");
  if (HW > 32 * scf::GN_MAXP) {     // large maps: the generic one-warp-per-(sample, group) kernel (no split copy available)
    SCF_REQUIRE(out_hl == nullptr && nsplit == 1, SCF_ERR_UNSUPPORTED, "scf_group_norm_relu_split: maps above %d pixels are not supported", 32 * scf::GN_MAXP);
    const int warps = B * num_groups;
    scf::group_norm_relu_kernel<<<scf::cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, B, HW, C, num_groups, eps);
    return scf::check_launch("group_norm_relu_kernel");
  }
  scf::launch_pdl(scf::group_norm_relu_c4_kernel, dim3(B, 4), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, HW, eps,
                                                                     reinterpret_cast<__nv_bfloat16*>(out_hl), plane_stride, nsplit,
                                                                     split_stride);
  return scf::check_launch("group_norm_relu_c4_kernel");
}
}  // namespace scf

extern "C" {

int scf_group_norm_relu(float* x, const float* gamma, const float* beta, int B, int HW, int C, int num_groups, float eps,
                        void* stream) {
  SCF_REQUIRE(x && gamma && beta && B > 0 && HW > 0 && C > 0 && num_groups > 0 && C % num_groups == 0, SCF_ERR_ARG,
              "scf_group_norm_relu: bad args");
  if (C == 128 && num_groups == 32 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(gamma) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(beta) % 16 == 0)
    return scf_group_norm_relu_split(x, gamma, beta, B, HW, C, num_groups, eps, nullptr, 0, stream);
  const int warps = B * num_groups;
  scf::group_norm_relu_kernel<<<scf::cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, B, HW, C, num_groups,
                                                                                    eps);
  return scf::check_launch("group_norm_relu_kernel");
}

int scf_linear(const float* x, const float* w, const float* bias, float* y, int B, int I, int O, int act, void* stream) {
  SCF_REQUIRE(x && w && y && B > 0 && I > 0 && O > 0, SCF_ERR_ARG, "scf_linear: bad args");
  SCF_REQUIRE(I % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0,
              SCF_ERR_ALIGN, "scf_linear: I must be a multiple of 4 and x/w 16B aligned");
  scf::linear_smem_kernel<<<scf::cdiv(O, 8), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, I, O, act);
  return scf::check_launch("linear_smem_kernel");
}

int scf_pose_project(const float* x, const float* rot_w, const float* rot_b, const float* tr_w, const float* tr_b,
                     const int64_t* label, float* d_rot, float* d_trs, int B, int I, int rot_dim, int num_class,
                     void* stream) {
  SCF_REQUIRE(x && rot_w && rot_b && tr_w && tr_b && d_rot && d_trs && B > 0 && I > 0 && rot_dim > 0, SCF_ERR_ARG,
              "scf_pose_project: bad args");
  SCF_REQUIRE(num_class <= 0 || label != nullptr, SCF_ERR_ARG, "scf_pose_project: multi-class head needs label");
  const int warps = B * (rot_dim + 3);
  scf::launch_pdl(scf::pose_project_kernel, dim3(scf::cdiv(warps, 8)), dim3(256), 0, (cudaStream_t)stream, x, rot_w, rot_b, tr_w, tr_b, label,
                                                                                 d_rot, d_trs, B, I, rot_dim, num_class);
  return scf::check_launch("pose_project_kernel");
}

}  // extern "C"
