// Pose / flow geometry: dense un-projection, pose update, pose-induced-flow re-projection, bilinear resize.
// HBM-bound streaming kernels (one pixel per thread, float4 point records, coalesced NCHW stores).
#include "scf_common.cuh"

namespace scf {

// general 3x3 inverse via the adjugate, evaluated in fp64 and rounded once (the reference uses torch.inverse,
// an fp32 LU; both are within a few ulp of the true inverse - SURVEY.md Appendix A.10)
__device__ void inverse3x3(const float* m, float* inv) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * Bc + c * C;
  const double r = 1.0 / det;
  inv[0] = (float)(A * r);
  inv[1] = (float)(-(b * i - c * h) * r);
  inv[2] = (float)((b * f - c * e) * r);
  inv[3] = (float)(Bc * r);
  inv[4] = (float)((a * i - c * g) * r);
  inv[5] = (float)(-(a * f - c * d) * r);
  inv[6] = (float)(C * r);
  inv[7] = (float)(-(a * h - b * g) * r);
  inv[8] = (float)((a * e - b * d) * r);
}

// pose.py:26-41,57-64 for every pixel: X_cam = K^-1 (x d, y d, d)^T ; X_obj = R^-1 (X_cam - t)
__global__ void __launch_bounds__(256) unproject_kernel(const float* __restrict__ depth, const float* __restrict__ K,
                                                        const float* __restrict__ rot, const float* __restrict__ trs,
                                                        float4* __restrict__ pts4, int H, int W) {
  __shared__ float kinv[9], rinv[9], t[3];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) inverse3x3(K + b * 9, kinv);
  if (threadIdx.x == 32) inverse3x3(rot + b * 9, rinv);
  if (threadIdx.x >= 64 && threadIdx.x < 67) t[threadIdx.x - 64] = trs[b * 3 + threadIdx.x - 64];
  __syncthreads();
  const int HW = H * W;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    const int y = pix / W, x = pix - y * W;
    const float d = depth[(long long)b * HW + pix];
    const float hx = (float)x * d, hy = (float)y * d, hz = d;
    const float cx = kinv[0] * hx + kinv[1] * hy + kinv[2] * hz - t[0];
    const float cy = kinv[3] * hx + kinv[4] * hy + kinv[5] * hz - t[1];
    const float cz = kinv[6] * hx + kinv[7] * hy + kinv[8] * hz - t[2];
    float4 o;
    o.x = rinv[0] * cx + rinv[1] * cy + rinv[2] * cz;
    o.y = rinv[3] * cx + rinv[4] * cy + rinv[5] * cz;
    o.z = rinv[6] * cx + rinv[7] * cy + rinv[8] * cz;
    o.w = d > 0.f ? 1.f : 0.f;
    pts4[(long long)b * HW + pix] = o;
  }
}

// pose.py:79-87 for one pixel: u = K (R X + t) ; flow = (u_x/u_z - x, u_y/u_z - y) on depth>0, `invalid` elsewhere
__device__ __forceinline__ float2 pose_flow_at(const float4 p, const float* k, const float* r, const float* t, int x, int y, float invalid) {
  float2 o = make_float2(invalid, invalid);
  if (p.w > 0.f) {
    const float vx = r[0] * p.x + r[1] * p.y + r[2] * p.z + t[0];
    const float vy = r[3] * p.x + r[4] * p.y + r[5] * p.z + t[1];
    const float vz = r[6] * p.x + r[7] * p.y + r[8] * p.z + t[2];
    const float ux = k[0] * vx + k[1] * vy + k[2] * vz;
    const float uy = k[3] * vx + k[4] * vy + k[5] * vz;
    const float uz = k[6] * vx + k[7] * vy + k[8] * vz;
    o.x = ux / uz - (float)x;
    o.y = uy / uz - (float)y;
  }
  return o;
}

// Pose-induced flow at full resolution (NCHW) and, in the trailing blocks of the same launch, the next iteration's 1/8-resolution
// flow  flow8 = 1/s * F.interpolate(flow, 1/s, bilinear, align_corners=True)  (scflow_decoder.py:196-197) as NHWC [B,H8,W8,2]:
// each coarse pixel blends the flows of its four full-resolution taps, evaluated with the same function as the dense map
// (same values as interpolating the stored map, without re-reading it and without a second launch).
__global__ void __launch_bounds__(256) reproject_kernel(const float4* __restrict__ pts4, const float* __restrict__ K,
                                                        const float* __restrict__ rot, const float* __restrict__ trs,
                                                        float invalid, float* __restrict__ flow, int H, int W, int full_blocks,
                                                        float* __restrict__ flow8, int H8, int W8, float ry, float rx, float scale8) {
  scf_pdl_enter();
  __shared__ float k[9], r[9], t[3];
  const int b = blockIdx.y;
  if (threadIdx.x < 9) { k[threadIdx.x] = K[b * 9 + threadIdx.x]; r[threadIdx.x] = rot[b * 9 + threadIdx.x]; }
  if (threadIdx.x < 3) t[threadIdx.x] = trs[b * 3 + threadIdx.x];
  __syncthreads();
  const int HW = H * W;
  const float4* pb = pts4 + (long long)b * HW;
  if ((int)blockIdx.x >= full_blocks) {
    const int P8 = H8 * W8;
    for (int q = ((int)blockIdx.x - full_blocks) * blockDim.x + threadIdx.x; q < P8; q += ((int)gridDim.x - full_blocks) * blockDim.x) {
      const int yo = q / W8, xo = q - yo * W8;
      const float sy = ry * (float)yo, sx = rx * (float)xo;
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
      const float ly1 = sy - (float)y0, ly0 = 1.f - ly1, lx1 = sx - (float)x0, lx0 = 1.f - lx1;
      const float2 v00 = pose_flow_at(__ldg(pb + y0 * W + x0), k, r, t, x0, y0, invalid);
      const float2 v01 = pose_flow_at(__ldg(pb + y0 * W + x1), k, r, t, x1, y0, invalid);
      const float2 v10 = pose_flow_at(__ldg(pb + y1 * W + x0), k, r, t, x0, y1, invalid);
      const float2 v11 = pose_flow_at(__ldg(pb + y1 * W + x1), k, r, t, x1, y1, invalid);
      float2 o;
      o.x = scale8 * (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
      o.y = scale8 * (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
      reinterpret_cast<float2*>(flow8)[(long long)b * P8 + q] = o;
    }
    return;
  }
  float* fx = flow + (long long)b * 2 * HW;
  float* fy = fx + HW;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += full_blocks * blockDim.x) {
    const int y = pix / W, x = pix - y * W;
    const float2 o = pose_flow_at(__ldg(pb + pix), k, r, t, x, y, invalid);
    fx[pix] = o.x;
    fy[pix] = o.y;
  }
}

// pose.py:124-169 (ortho6d, depth_transform='exp'): one thread per sample
__global__ void pose_update_kernel(const float* __restrict__ d_rot, const float* __restrict__ d_trs,
                                   const float* __restrict__ rot_in, const float* __restrict__ trs_in,
                                   float* __restrict__ rot_out, float* __restrict__ trs_out, int B) {
  scf_pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* o = d_rot + b * 6;
  float x[3] = {o[0], o[1], o[2]}, yr[3] = {o[3], o[4], o[5]};
  float n = fmaxf(sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), 1e-12f);       // F.normalize eps
  x[0] /= n; x[1] /= n; x[2] /= n;
  float z[3] = {x[1] * yr[2] - x[2] * yr[1], x[2] * yr[0] - x[0] * yr[2], x[0] * yr[1] - x[1] * yr[0]};
  n = fmaxf(sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]), 1e-12f);
  z[0] /= n; z[1] /= n; z[2] /= n;
  const float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
  // dR columns are (x, y, z):  dR[i][0]=x[i], dR[i][1]=y[i], dR[i][2]=z[i]
  const float* R = rot_in + b * 9;
  float Rn[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = x[i] * R[j] + y[i] * R[3 + j] + z[i] * R[6 + j];
  const float* t = trs_in + b * 3;
  const float* dt = d_trs + b * 3;
  const float tz = t[2] / expf(dt[2]);
  const float tx = tz * (dt[0] / 10.f + t[0] / t[2]);
  const float ty = tz * (dt[1] / 10.f + t[1] / t[2]);
#pragma unroll
  for (int i = 0; i < 9; ++i) rot_out[b * 9 + i] = Rn[i];
  trs_out[b * 3 + 0] = tx;
  trs_out[b * 3 + 1] = ty;
  trs_out[b * 3 + 2] = tz;
}

struct ResizeParams {
  const float* src; const float* add;
  long long s_b, s_c, s_y, s_x;
  int Hi, Wi;
  float* dst;
  long long d_b, d_c, d_y, d_x;
  int Ho, Wo, B, C;
  float scale, ry, rx;
};

// F.interpolate(mode='bilinear', align_corners=True): src = dst*(in-1)/(out-1), ATen's lambda formulation
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const ResizeParams p) {
  const long long total = (long long)p.B * p.C * p.Ho * p.Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % p.Wo);
    long long r = idx / p.Wo;
    const int yo = (int)(r % p.Ho);
    r /= p.Ho;
    const int c = (int)(r % p.C);
    const int b = (int)(r / p.C);
    const float sy = p.ry * (float)yo, sx = p.rx * (float)xo;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
    const float ly1 = sy - (float)y0, ly0 = 1.f - ly1, lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const long long base = b * p.s_b + c * p.s_c;
    const long long i00 = base + y0 * p.s_y + x0 * p.s_x, i01 = base + y0 * p.s_y + x1 * p.s_x;
    const long long i10 = base + y1 * p.s_y + x0 * p.s_x, i11 = base + y1 * p.s_y + x1 * p.s_x;
    float v00 = p.src[i00], v01 = p.src[i01], v10 = p.src[i10], v11 = p.src[i11];
    if (p.add) { v00 += p.add[i00]; v01 += p.add[i01]; v10 += p.add[i10]; v11 += p.add[i11]; }
    const float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    p.dst[b * p.d_b + c * p.d_c + yo * p.d_y + xo * p.d_x] = p.scale * v;
  }
}

// Fast path for contiguous-x destinations (NCHW up-sampling, the two x8 resizes of every iteration): one thread
// produces 4 consecutive output pixels of one row (float4 store), 32-bit index math, row weights computed once.
__global__ void __launch_bounds__(256) resize_bilinear_rows_kernel(const ResizeParams p) {
  const int wo4 = p.Wo >> 2;
  const int total = p.B * p.C * p.Ho * wo4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int xq = idx % wo4;
    int r = idx / wo4;
    const int yo = r % p.Ho;
    r /= p.Ho;
    const int c = r % p.C, b = r / p.C;
    const float sy = p.ry * (float)yo;
    const int y0 = (int)sy, y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0);
    const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
    const long long base = b * p.s_b + c * p.s_c;
    const float* r0 = p.src + base + y0 * p.s_y;
    const float* r1 = p.src + base + y1 * p.s_y;
    const float* a0 = p.add ? p.add + base + y0 * p.s_y : nullptr;
    const float* a1 = p.add ? p.add + base + y1 * p.s_y : nullptr;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sx = p.rx * (float)(xq * 4 + j);
      const int x0 = (int)sx, x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
      const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
      float v00 = __ldg(r0 + x0 * p.s_x), v01 = __ldg(r0 + x1 * p.s_x), v10 = __ldg(r1 + x0 * p.s_x), v11 = __ldg(r1 + x1 * p.s_x);
      if (a0) { v00 += __ldg(a0 + x0 * p.s_x); v01 += __ldg(a0 + x1 * p.s_x); v10 += __ldg(a1 + x0 * p.s_x); v11 += __ldg(a1 + x1 * p.s_x); }
      o[j] = p.scale * (ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11));
    }
    *reinterpret_cast<float4*>(p.dst + b * p.d_b + c * p.d_c + yo * p.d_y + xq * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace scf

extern "C" {

int scf_unproject(const float* depth, const float* K, const float* rot, const float* trs, float* pts4, int B, int H,
                  int W, void* stream) {
  SCF_REQUIRE(depth && K && rot && trs && pts4 && B > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_unproject: bad args");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(pts4) % 16 == 0, SCF_ERR_ALIGN, "scf_unproject: pts4 must be 16B aligned");
  dim3 grid(scf::cdiv(H * W, 256 * 4), B);
  scf::unproject_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, K, rot, trs, reinterpret_cast<float4*>(pts4), H, W);
  return scf::check_launch("unproject_kernel");
}

static int reproject_launch(const float* pts4, const float* K, const float* rot, const float* trs, float invalid, float* flow,
                            int B, int H, int W, float* flow8, int H8, int W8, void* stream) {
  SCF_REQUIRE(pts4 && K && rot && trs && flow && B > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_reproject: bad args");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(pts4) % 16 == 0, SCF_ERR_ALIGN, "scf_reproject: pts4 must be 16B aligned");
  SCF_REQUIRE(!flow8 || (H8 > 0 && W8 > 0 && H8 <= H && W8 <= W && reinterpret_cast<uintptr_t>(flow8) % 8 == 0), SCF_ERR_ARG,
              "scf_reproject_down: bad coarse size / alignment");
  const int full_blocks = scf::cdiv(H * W, 256 * 4), coarse_blocks = flow8 ? scf::cdiv(H8 * W8, 256) : 0;
  dim3 grid(full_blocks + coarse_blocks, B);
  // ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 when out == 1 (as scf_resize_bilinear)
  const float ry = H8 > 1 ? (float)(H - 1) / (float)(H8 - 1) : 0.f, rx = W8 > 1 ? (float)(W - 1) / (float)(W8 - 1) : 0.f;
  const float scale8 = flow8 ? 1.0f / (float)(H / H8) : 1.f;
  scf::launch_pdl(scf::reproject_kernel, grid, dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(pts4), K, rot, trs, invalid, flow, H, W,
                  full_blocks, flow8, H8, W8, ry, rx, scale8);
  return scf::check_launch("reproject_kernel");
}

int scf_reproject(const float* pts4, const float* K, const float* rot, const float* trs, float invalid, float* flow,
                  int B, int H, int W, void* stream) {
  return reproject_launch(pts4, K, rot, trs, invalid, flow, B, H, W, nullptr, 0, 0, stream);
}

int scf_reproject_down(const float* pts4, const float* K, const float* rot, const float* trs, float invalid, float* flow,
                       int B, int H, int W, float* flow8, int H8, int W8, void* stream) {
  SCF_REQUIRE(flow8 != nullptr && H8 > 0 && H % H8 == 0 && W8 > 0 && W / W8 == H / H8, SCF_ERR_ARG,
              "scf_reproject_down: the coarse map must be the full one divided by one integer factor");
  return reproject_launch(pts4, K, rot, trs, invalid, flow, B, H, W, flow8, H8, W8, stream);
}

int scf_pose_update(const float* d_rot, const float* d_trs, const float* rot_in, const float* trs_in, float* rot_out,
                    float* trs_out, int B, void* stream) {
  SCF_REQUIRE(d_rot && d_trs && rot_in && trs_in && rot_out && trs_out && B > 0, SCF_ERR_ARG, "scf_pose_update: bad args");
  scf::launch_pdl(scf::pose_update_kernel, dim3(scf::cdiv(B, 64)), dim3(64), 0, (cudaStream_t)stream, d_rot, d_trs, rot_in, trs_in, rot_out, trs_out, B);
  return scf::check_launch("pose_update_kernel");
}

int scf_resize_bilinear(const float* src, const float* add, long long s_b, long long s_c, long long s_y, long long s_x,
                        int Hi, int Wi, float* dst, long long d_b, long long d_c, long long d_y, long long d_x, int Ho,
                        int Wo, int B, int C, float scale, void* stream) {
  SCF_REQUIRE(src && dst && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, SCF_ERR_ARG,
              "scf_resize_bilinear: bad args");
  scf::ResizeParams p;
  p.src = src; p.add = add; p.s_b = s_b; p.s_c = s_c; p.s_y = s_y; p.s_x = s_x; p.Hi = Hi; p.Wi = Wi;
  p.dst = dst; p.d_b = d_b; p.d_c = d_c; p.d_y = d_y; p.d_x = d_x; p.Ho = Ho; p.Wo = Wo; p.B = B; p.C = C;
  p.scale = scale;
  // ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 when out == 1
  p.ry = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  p.rx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)B * C * Ho * Wo;
  if (d_x == 1 && Wo % 4 == 0 && d_y % 4 == 0 && d_c % 4 == 0 && d_b % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0 &&
      total < (1ll << 31)) {
    const int blocks = (int)((total / 4 + 255) / 256 < 148 * 16 ? (total / 4 + 255) / 256 : 148 * 16);
    scf::resize_bilinear_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return scf::check_launch("resize_bilinear_rows_kernel");
  }
  const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  scf::resize_bilinear_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return scf::check_launch("resize_bilinear_kernel");
}

}  // extern "C"
