// RAFT 'Basic' feature / context encoder (models/encoder/raft_encoder.py:286-314, models/backbone/resnet.py:14-94,
// 678-773) on the library's kernels: the 7x7 stride-2 stem as an x-folded 7x1 tcgen05 convolution, every 3x3 / 1x1
// convolution on the tcgen05 split-bf16 kernel (stride 2 through TMA element strides), InstanceNorm as deterministic
// per-tile statistics emitted by the convolution's epilogue + one fused normalise/ReLU/residual pass, eval-mode BatchNorm
// folded into the packed weights.
// This is SURVEY.md §8(f) rank 1: the component feeding the refinement loop.
#include "scf_common.cuh"
#include <stdlib.h>
#include "scf_tc.cuh"

namespace scf {

int conv2d_f32(const scf_conv_desc& d, cudaStream_t st);
int conv2d_thin(const scf_conv_desc& d, int in_nchw, cudaStream_t st);
int conv2d_tc(const scf_tc_conv_desc& d, cudaStream_t st);
bool conv2d_stem_rows_fold_ok(const scf_tc_conv_desc& d, const float* images);
bool conv2d_rows_eligible(const scf_tc_conv_desc& d);
int conv2d_stem_rows_fold(const float* images, const scf_tc_conv_desc& d, cudaStream_t st);
void conv2d_tc_last_tiles(int* m_tiles, int* per_img, int* stat_rows);
int conv2d_tc_max_tiles(int B, int Hout, int Wout);
int im2col_x_split(const float* in, int nchw, int cin, int kw, void* out_hl, long long plane, int N, int H, int Wi, int sx,
                   cudaStream_t st);
int pack_conv_weight_tc_foldx(const float* w_oihw, float* scratch, void* packed, int O, int C, int KH, int KW, int cin_pad,
                              int cout_pad, int o_off, cudaStream_t st);

// ------------------------------------------------------------------ conv-unit table
struct EncUnit { int cin, cout, k, stride; };
static const EncUnit kUnits[SCF_ENC_UNITS] = {
    {3, 64, 7, 2},                                                            // 0 stem conv1
    {64, 64, 3, 1}, {64, 64, 3, 1}, {64, 64, 3, 1}, {64, 64, 3, 1},          // 1-4  res_layer1.{0,1}.{conv1,conv2}
    {64, 96, 3, 2}, {96, 96, 3, 1}, {64, 96, 1, 2}, {96, 96, 3, 1}, {96, 96, 3, 1},      // 5-9  res_layer2 (7 = downsample)
    {96, 128, 3, 2}, {128, 128, 3, 1}, {96, 128, 1, 2}, {128, 128, 3, 1}, {128, 128, 3, 1},  // 10-14 res_layer3 (12 = downsample)
    {128, 256, 1, 1},                                                         // 15 conv2
};

struct EncArena {
  size_t w_f32[SCF_ENC_UNITS];   // byte offsets: folded fp32 OIHW copy (source of both packings)
  size_t bias[SCF_ENC_UNITS];    // folded bias, padded to a multiple of 16 floats
  size_t packed[SCF_ENC_UNITS];  // bf16 [2][taps][cout_pad][cin_pad]; unit 0 (7x7 stem) is folded to 7x1 over kx*3+c channels
  size_t fold0;                  // scratch for the folded stem weight
  size_t packed15_half[2];       // unit 15 (conv2) once more as two 128-output-channel halves (h / context heads of the context encoder)
  int cin_pad[SCF_ENC_UNITS], cout_pad[SCF_ENC_UNITS];
  size_t total;
};

static void build_enc_arena(EncArena& a) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  for (int u = 0; u < SCF_ENC_UNITS; ++u) {
    const EncUnit& e = kUnits[u];
    a.cin_pad[u] = (e.cin + 7) / 8 * 8;
    a.cout_pad[u] = (e.cout + 15) / 16 * 16;
    a.w_f32[u] = take((size_t)e.cout * e.cin * e.k * e.k * 4);
    a.bias[u] = take((size_t)a.cout_pad[u] * 4);
    if (u == 0) {
      a.cin_pad[u] = 32;       // 7 taps x 3 channels = 21, padded to a power-of-two pixel pitch
      a.packed[u] = take((size_t)2 * e.k * a.cout_pad[u] * a.cin_pad[u] * 2);
      a.fold0 = take((size_t)e.cout * e.cin * e.k * e.k * 4);
    } else a.packed[u] = take((size_t)2 * e.k * e.k * a.cout_pad[u] * a.cin_pad[u] * 2);
  }
  for (int i = 0; i < 2; ++i) a.packed15_half[i] = take((size_t)2 * 128 * 128 * 2);
  a.total = off;
}

// w'[o,...] = w[o,...] * g[o] / sqrt(var[o] + eps) ; b'[o] = (b[o] - mean[o]) * g[o] / sqrt(var[o]+eps) + beta[o]
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ g,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ wo, float* __restrict__ bo, int O, int per_o) {
  const int o = blockIdx.x;
  const float s = g ? g[o] * rsqrtf(var[o] + eps) : 1.f;
  for (int i = threadIdx.x; i < per_o; i += blockDim.x) wo[(long long)o * per_o + i] = w[(long long)o * per_o + i] * s;
  if (threadIdx.x == 0) bo[o] = g ? ((b ? b[o] : 0.f) - mean[o]) * s + beta[o] : (b ? b[o] : 0.f);
}

// ---- InstanceNorm, stage 1: per (image, pixel-slice) partial sums of x and x^2 for every channel (deterministic)
constexpr int IN_SLICES = 32;
__global__ void __launch_bounds__(256) instnorm_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int HW, int C) {
  // grid (IN_SLICES, N); thread t owns channel quad (t % (C/4)) on pixels t/(C/4) + k*(256/(C/4))
  __shared__ float4 rs[256], rq[256];
  const int n = blockIdx.y, sl = blockIdx.x;
  const int c4n = C >> 2, rows = 256 / c4n;
  const int cq = threadIdx.x % c4n, r = threadIdx.x / c4n;
  const int p0 = (int)((long long)HW * sl / IN_SLICES), p1 = (int)((long long)HW * (sl + 1) / IN_SLICES);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (r < rows)
    for (int p = p0 + r; p < p1; p += rows) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((long long)n * HW + p) * C) + cq);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
    }
  rs[threadIdx.x] = s; rq[threadIdx.x] = q;
  __syncthreads();
  if (r == 0) {
    for (int i = 1; i < rows; ++i) {
      const float4 a = rs[i * c4n + cq], b = rq[i * c4n + cq];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* o = part + (((long long)n * IN_SLICES + sl) * 2) * C + cq * 4;
    *reinterpret_cast<float4*>(o) = s;
    *reinterpret_cast<float4*>(o + C) = q;
  }
}
// stage 2: mean / rstd per (image, channel), combined in double
__global__ void instnorm_finalize_kernel(const float* __restrict__ part, float* __restrict__ stat, int HW, int C, float eps, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * C) return;
  const int n = idx / C, c = idx - n * C;
  double s = 0., q = 0.;
  for (int sl = 0; sl < IN_SLICES; ++sl) {
    s += part[(((long long)n * IN_SLICES + sl) * 2) * C + c];
    q += part[(((long long)n * IN_SLICES + sl) * 2 + 1) * C + c];
  }
  const double mean = s / HW;
  double var = q / HW - mean * mean;
  if (var < 0.) var = 0.;
  stat[(long long)idx * 2] = (float)mean;
  stat[(long long)idx * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}
// stage 2 for statistics gathered by the convolution's own epilogue (scf_conv2d_tc `stats`): part is
// [N][rows][2][C] (rows = partial-sum rows per image, see conv2d_tc_last_tiles); one block per image, deterministic double-precision combine
__global__ void __launch_bounds__(256) instnorm_finalize_tiles_kernel(const float* __restrict__ part, float* __restrict__ stat,
                                                                      int HW, int C, float eps, int rows) {
  // grid (N, C/8): thread = (channel c of 8, row group g of 32); 8 independent loads in flight per thread; fixed combine order
  __shared__ double ss[256], sq[256];
  const int n = blockIdx.x;
  const int cl = threadIdx.x & 7, c = blockIdx.y * 8 + cl, g = threadIdx.x >> 3;
  const float* base = part + (long long)n * rows * 2 * C + c;
  double s = 0., q = 0.;
  for (int r0 = g; r0 < rows; r0 += 256) {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + 32 * j;
      a[j] = r < rows ? base[(long long)r * 2 * C] : 0.f;
      b[j] = r < rows ? base[(long long)r * 2 * C + C] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; q += b[j]; }
  }
  ss[threadIdx.x] = s; sq[threadIdx.x] = q;
  __syncthreads();
  if (g == 0) {
    for (int i = 1; i < 32; ++i) { s += ss[i * 8 + cl]; q += sq[i * 8 + cl]; }
    const double mean = s / HW;
    double var = q / HW - mean * mean;
    if (var < 0.) var = 0.;
    stat[((long long)n * C + c) * 2] = (float)mean;
    stat[((long long)n * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}
// stage 3: y = [relu]( (x - mean) * rstd  [+ (r - rmean) * rrstd | + r] ) -> fp32 and/or split-bf16.  HBM-streaming: a thread
// owns 8 consecutive channels (two 16 B loads per input, one 16 B store per bf16 plane), two pixels in flight per iteration.
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ stat,
                                                             const float* __restrict__ res, const float* __restrict__ res_stat,
                                                             const __nv_bfloat16* __restrict__ res_hl,
                                                             int relu, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hl,
                                                             long long plane, int HW, int C, long long total8, int reverse) {
  // reverse: sweep the map back to front - the producing convolution wrote its last images most recently (still in L2), and the
  // consuming convolution starts at image 0, which this sweep then writes last
  const int c8n = C >> 3;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < total8; it += (long long)gridDim.x * blockDim.x) {
    const long long idx = reverse ? total8 - 1 - it : it;
    const int cq = (int)(idx % c8n);
    const long long pix = idx / c8n;
    const long long n = pix / HW;
    const float4 v0 = __ldcs(reinterpret_cast<const float4*>(x) + 2 * idx), v1 = __ldcs(reinterpret_cast<const float4*>(x) + 2 * idx + 1);
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float4* st = reinterpret_cast<const float4*>(stat + (n * C + cq * 8) * 2);     // (mean, rstd) pairs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 s2 = __ldg(st + i);
      vv[2 * i] = (vv[2 * i] - s2.x) * s2.y;
      vv[2 * i + 1] = (vv[2 * i + 1] - s2.z) * s2.w;
    }
    if (res) {
      const float4 r0 = __ldcs(reinterpret_cast<const float4*>(res) + 2 * idx), r1 = __ldcs(reinterpret_cast<const float4*>(res) + 2 * idx + 1);
      float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      if (res_stat) {
        const float4* rt = reinterpret_cast<const float4*>(res_stat + (n * C + cq * 8) * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 s2 = __ldg(rt + i);
          rr[2 * i] = (rr[2 * i] - s2.x) * s2.y;
          rr[2 * i + 1] = (rr[2 * i + 1] - s2.z) * s2.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) vv[i] += rr[i];
    } else if (res_hl) {
      // identity branch of a BasicBlock read from the block input's split-bf16 planes (hi + lo carries 16 mantissa bits), so no
      // separate fp32 copy of every block output has to be written and read back
      const uint4 h4 = __ldcs(reinterpret_cast<const uint4*>(res_hl + idx * 8)), l4 = __ldcs(reinterpret_cast<const uint4*>(res_hl + plane + idx * 8));
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h4);
      const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a = __bfloat1622float2(hp[i]), b2 = __bfloat1622float2(lp[i]);
        vv[2 * i] += a.x + b2.x;
        vv[2 * i + 1] += a.y + b2.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) vv[i] = fmaxf(vv[i], 0.f);
    }
    if (out_f32) {
      reinterpret_cast<float4*>(out_f32)[2 * idx] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      reinterpret_cast<float4*>(out_f32)[2 * idx + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
    }
    if (out_hl) {
      __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tc::split_bf16(vv[i], hi[i], lo[i]);
      *reinterpret_cast<uint4*>(out_hl + idx * 8) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(out_hl + plane + idx * 8) = *reinterpret_cast<const uint4*>(lo);
    }
  }
}

struct EncWs {
  size_t t0, raw, raw2, xf[2], xs[2], ts, part, stat, stat2;
  size_t total;
};
static void build_enc_ws(int N, int H, int W, EncWs& w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t p2 = (size_t)N * (H / 2) * (W / 2);
  w.t0 = take((size_t)N * H * (W / 2) * 32 * 2 * 2);     // x-folded image, split-bf16 [2][N,H,W/2,32]
  w.raw = take(p2 * 64 * 4); w.raw2 = take(p2 / 4 * 96 * 4 + 4096);
  for (int i = 0; i < 2; ++i) { w.xf[i] = take(p2 * 64 * 4); w.xs[i] = take(p2 * 64 * 2 * 2); }
  w.ts = take(p2 * 64 * 2 * 2);
  {
    // partial sums: either IN_SLICES per image (separate pass) or one row per (128-pixel tile, epilogue warp) of a convolution
    size_t part = (size_t)N * IN_SLICES * 2 * 128 * 4;
    int hh = H / 2, ww = W / 2;
    const int chans[3] = {64, 96, 128};
    for (int l = 0; l < 3; ++l) {
      const size_t need = (size_t)conv2d_tc_max_tiles(N, hh, ww) * 4 * 2 * chans[l] * 4;
      if (need > part) part = need;
      hh = (hh - 1) / 2 + 1; ww = (ww - 1) / 2 + 1;
    }
    w.part = take(part);
  }
  w.stat = take((size_t)N * 128 * 2 * 4); w.stat2 = take((size_t)N * 128 * 2 * 4);
  w.total = off;
}

}  // namespace scf

using namespace scf;

extern "C" {

size_t scf_encoder_packed_bytes(void) {
  EncArena a;
  build_enc_arena(a);
  return a.total;
}

size_t scf_encoder_workspace_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  EncWs w;
  build_enc_ws(N, H, W, w);
  return w.total;
}

int scf_encoder_pack(int norm, const float* const* h_weights, void* packed, void* stream) {
  SCF_REQUIRE(h_weights && packed, SCF_ERR_ARG, "scf_encoder_pack: null pointer");
  SCF_REQUIRE(norm == SCF_ENC_NORM_IN || norm == SCF_ENC_NORM_BN, SCF_ERR_ARG, "scf_encoder_pack: norm must be IN (0) or BN (1)");
  cudaStream_t st = (cudaStream_t)stream;
  EncArena a;
  build_enc_arena(a);
  char* base = reinterpret_cast<char*>(packed);
  SCF_CUDA(cudaMemsetAsync(packed, 0, a.total, st));
  for (int u = 0; u < SCF_ENC_UNITS; ++u) {
    const EncUnit& e = kUnits[u];
    const float* const* s = h_weights + u * 6;     // w, b, bn_weight, bn_bias, bn_mean, bn_var
    SCF_REQUIRE(s[0] != nullptr, SCF_ERR_ARG, "scf_encoder_pack: conv weight of unit %d is null", u);
    const bool fold = norm == SCF_ENC_NORM_BN && u != SCF_ENC_UNITS - 1;
    if (fold) SCF_REQUIRE(s[2] && s[3] && s[4] && s[5], SCF_ERR_ARG, "scf_encoder_pack: BatchNorm parameters of unit %d missing", u);
    float* wf = reinterpret_cast<float*>(base + a.w_f32[u]);
    float* bf = reinterpret_cast<float*>(base + a.bias[u]);
    fold_bn_kernel<<<e.cout, 128, 0, st>>>(s[0], s[1], fold ? s[2] : nullptr, s[3], s[4], s[5], 1e-5f, wf, bf, e.cout,
                                           e.cin * e.k * e.k);
    SCF_TRY(check_launch("fold_bn_kernel"));
    if (u == 0) SCF_TRY(pack_conv_weight_tc_foldx(wf, reinterpret_cast<float*>(base + a.fold0), base + a.packed[u], e.cout, e.cin, e.k,
                                                  e.k, a.cin_pad[u], a.cout_pad[u], 0, st));
    else SCF_TRY(scf_pack_conv_weight_tc(wf, base + a.packed[u], e.cout, e.cin, e.k, e.k, a.cin_pad[u], a.cout_pad[u], 0, st));
    if (u == SCF_ENC_UNITS - 1)
      for (int i = 0; i < 2; ++i)
        SCF_TRY(scf_pack_conv_weight_tc(wf + (size_t)i * 128 * e.cin, base + a.packed15_half[i], 128, e.cin, 1, 1, 128, 128, 0, st));
  }
  return 0;
}

static int encoder_forward_impl(int norm, const void* packed, const float* images, int N, int H, int W, float* out_nchw,
                                const scf_encoder_out* ex, void* workspace, size_t workspace_bytes, void* stream) {
  SCF_REQUIRE(packed && images && (out_nchw || ex) && workspace, SCF_ERR_ARG, "scf_encoder_forward: null pointer");
  SCF_REQUIRE(norm == SCF_ENC_NORM_IN || norm == SCF_ENC_NORM_BN, SCF_ERR_ARG, "scf_encoder_forward: norm must be IN (0) or BN (1)");
  SCF_REQUIRE(N > 0 && H >= 16 && W >= 16 && H % 8 == 0 && W % 8 == 0, SCF_ERR_ARG, "scf_encoder_forward: H, W must be multiples of 8");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0 && reinterpret_cast<uintptr_t>(packed) % 256 == 0, SCF_ERR_ALIGN,
              "scf_encoder_forward: workspace / arena must be 256B aligned");
  EncArena a;
  build_enc_arena(a);
  EncWs ws;
  build_enc_ws(N, H, W, ws);
  SCF_REQUIRE(workspace_bytes >= ws.total, SCF_ERR_ARG, "scf_encoder_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* wsb = reinterpret_cast<char*>(workspace);
  const char* pk = reinterpret_cast<const char*>(packed);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(wsb + off); };
  auto S = [&](size_t off) { return reinterpret_cast<void*>(wsb + off); };
  auto BIAS = [&](int u) { return reinterpret_cast<const float*>(pk + a.bias[u]); };
  const bool inorm = norm == SCF_ENC_NORM_IN;

  // ---- stem: 7x7 stride-2 convolution on the tensor cores: kernel rows folded into the channel axis (x-im2col with
  // stride 2 -> 21 channels), then a 7x1 convolution with vertical stride 2 (see scf_conv_tc.cu)
  int h = H / 2, w = W / 2;
  long long npix = (long long)N * h * w;
  // (when the rolling-rows stem kernel applies it x-folds the image itself: no im2col launch, no 4.2 MB-per-image intermediate)
  bool stem_folded = false, stem_probed = false;
  auto stem_conv = [&](int act, float* of32, void* ohl, float* stats) -> int {
    scf_tc_conv_desc d = {};
    d.seg[0].ptr = S(ws.t0); d.seg[0].plane_stride = (long long)N * H * w * a.cin_pad[0]; d.seg[0].stride = a.cin_pad[0];
    d.seg[0].coff = 0; d.seg[0].nch = a.cin_pad[0];
    d.nseg = 1;
    d.B = N; d.H = H; d.W = w; d.kh = 7; d.kw = 1; d.stride_x = 1; d.stride_y = 2;
    d.w = pk + a.packed[0]; d.cin_pad = a.cin_pad[0]; d.cout_pad = a.cout_pad[0]; d.cout = 64;
    d.bias = BIAS(0); d.scale = 1.f; d.epi = SCF_EPI_ACT; d.act = act;
    d.out_f32 = of32; d.out_f32_stride = 64;
    d.out_hl = ohl; d.out_hl_plane = npix * 64; d.out_hl_stride = 64;
    d.stats = stats;
    if (!stem_probed) {
      stem_probed = true;
      stem_folded = conv2d_stem_rows_fold_ok(d, images);
      if (!stem_folded) SCF_TRY(im2col_x_split(images, /*nchw=*/1, 3, 7, S(ws.t0), (long long)N * H * w * a.cin_pad[0], N, H, W, 2, st));
    }
    return stem_folded ? conv2d_stem_rows_fold(images, d, st) : conv2d_tc(d, st);
  };
  // InstanceNorm helpers
  auto in_stats = [&](const float* x, float* stat, int hw, int C) -> int {
    instnorm_partial_kernel<<<dim3(IN_SLICES, N), 256, 0, st>>>(x, F(ws.part), hw, C);
    SCF_TRY(check_launch("instnorm_partial_kernel"));
    instnorm_finalize_kernel<<<cdiv(N * C, 128), 128, 0, st>>>(F(ws.part), stat, hw, C, 1e-5f, N);
    return check_launch("instnorm_finalize_kernel");
  };
  static const bool apply_reverse = [] { const char* e = getenv("SCFLOW_IN_REVERSE"); return e ? atoi(e) != 0 : true; }();
  auto in_apply = [&](const float* x, const float* stat, const float* res, const float* res_stat, int relu, float* of32, void* ohl,
                      long long pixels, int hw, int C, const void* res_hl = nullptr) -> int {
    const long long total8 = pixels * C / 8;
    instnorm_apply_kernel<<<cdiv(total8, 256) < 148 * 16 ? cdiv(total8, 256) : 148 * 16, 256, 0, st>>>(
        x, stat, res, res_stat, reinterpret_cast<const __nv_bfloat16*>(res_hl), relu, of32, reinterpret_cast<__nv_bfloat16*>(ohl),
        pixels * C, hw, C, total8, apply_reverse ? 1 : 0);
    return check_launch("instnorm_apply_kernel");
  };
  // tensor-core conv unit u on split input (hin x win), stride from the table
  auto tcconv = [&](int u, void* in_s, int hin, int win, int act, float* of32, void* ohl, const float* residual,
                    float* stats = nullptr, const void* residual_hl = nullptr, bool probe_rows = false) -> int {
    const EncUnit& e = kUnits[u];
    scf_tc_conv_desc d = {};
    const long long in_pix = (long long)N * hin * win;
    d.seg[0].ptr = in_s; d.seg[0].plane_stride = in_pix * e.cin; d.seg[0].stride = e.cin; d.seg[0].coff = 0; d.seg[0].nch = e.cin;
    d.nseg = 1;
    d.B = N; d.H = hin; d.W = win; d.kh = d.kw = e.k; d.stride = e.stride;
    d.w = pk + a.packed[u]; d.cin_pad = a.cin_pad[u]; d.cout_pad = a.cout_pad[u]; d.cout = e.cout;
    d.bias = BIAS(u); d.scale = 1.f; d.epi = SCF_EPI_ACT; d.act = act;
    const int ho = (hin + 2 * (e.k / 2) - e.k) / e.stride + 1, wo = (win + 2 * (e.k / 2) - e.k) / e.stride + 1;
    d.out_f32 = of32; d.out_f32_stride = e.cout;
    d.out_hl = ohl; d.out_hl_plane = (long long)N * ho * wo * e.cout; d.out_hl_stride = e.cout;
    d.aux0 = residual; d.aux0_stride = e.cout;
    d.aux0_hl = residual_hl; d.aux0_hl_plane = (long long)N * ho * wo * e.cout; d.aux0_hl_stride = e.cout;
    d.stats = stats;
    if (probe_rows) return conv2d_rows_eligible(d) ? 1 : 0;       // would this layer run on the rolling-rows kernel?
    return conv2d_tc(d, st);
  };
  // convolution + InstanceNorm statistics of its output: the conv epilogue emits per-tile partial sums when a tile never
  // spans two images (always, except for tiny maps), otherwise a separate statistics pass reads the map back
  auto tcconv_stats = [&](int u, void* in_s, int hin, int win, float* raw, float* stat) -> int {
    const EncUnit& e = kUnits[u];
    const int ho = (hin + 2 * (e.k / 2) - e.k) / e.stride + 1, wo = (win + 2 * (e.k / 2) - e.k) / e.stride + 1;
    int per_img = 0;
    scf_conv2d_tc_tiles(N, ho, wo, &per_img);
    if (per_img > 0 && e.cout % 32 == 0 && ho * wo >= 128) {
      SCF_TRY(tcconv(u, in_s, hin, win, SCF_ACT_NONE, raw, nullptr, nullptr, F(ws.part)));
      int rows = 0;
      conv2d_tc_last_tiles(nullptr, &per_img, &rows);      // the tiling the launch actually used (halo / transposed tiles differ)
      SCF_REQUIRE(per_img > 0, SCF_ERR_UNSUPPORTED, "scf_encoder_forward: statistics need one sample per tile");
      instnorm_finalize_tiles_kernel<<<dim3(N, e.cout / 8), 256, 0, st>>>(F(ws.part), stat, ho * wo, e.cout, 1e-5f, rows);
      return check_launch("instnorm_finalize_tiles_kernel");
    }
    SCF_TRY(tcconv(u, in_s, hin, win, SCF_ACT_NONE, raw, nullptr, nullptr));
    return in_stats(raw, stat, ho * wo, e.cout);
  };

  int cur = 0;                         // block input lives in xf[cur] (fp32) and xs[cur] (split)
  // folded-BatchNorm encoder, 64-channel stage: when its convolutions run on the rolling-rows kernel the identity branch is read
  // from the block input's split-bf16 planes, so no fp32 copy of the stem / block outputs is written at all
  static const bool split_id_on = [] { const char* e = getenv("SCFLOW_ENC_SPLIT_ID"); return e ? atoi(e) != 0 : true; }();
  const bool split_identity = split_id_on && !inorm && tcconv(2, S(ws.ts), h, w, SCF_ACT_RELU, nullptr, S(ws.xs[1]), nullptr, nullptr, S(ws.xs[0]), true) == 1;
  if (!inorm) {
    SCF_TRY(stem_conv(SCF_ACT_RELU, split_identity ? nullptr : F(ws.xf[0]), S(ws.xs[0]), nullptr));
  } else {
    int per_img = 0;
    scf_conv2d_tc_tiles(N, h, w, &per_img);
    if (per_img > 0) {
      SCF_TRY(stem_conv(SCF_ACT_NONE, F(ws.raw), nullptr, F(ws.part)));
      int rows = 0;
      conv2d_tc_last_tiles(nullptr, &per_img, &rows);
      instnorm_finalize_tiles_kernel<<<dim3(N, 8), 256, 0, st>>>(F(ws.part), F(ws.stat), h * w, 64, 1e-5f, rows);
      SCF_TRY(check_launch("instnorm_finalize_tiles_kernel"));
    } else {
      SCF_TRY(stem_conv(SCF_ACT_NONE, F(ws.raw), nullptr, nullptr));
      SCF_TRY(in_stats(F(ws.raw), F(ws.stat), h * w, 64));
    }
    SCF_TRY(in_apply(F(ws.raw), F(ws.stat), nullptr, nullptr, 1, nullptr, S(ws.xs[0]), npix, h * w, 64));
  }
  // ---- residual stages: units (c1, c2, ds) per block
  const int blocks[6][3] = {{1, 2, -1}, {3, 4, -1}, {5, 6, 7}, {8, 9, -1}, {10, 11, 12}, {13, 14, -1}};
  for (int bi = 0; bi < 6; ++bi) {
    const int u1 = blocks[bi][0], u2 = blocks[bi][1], ud = blocks[bi][2];
    const EncUnit& e1 = kUnits[u1];
    const int ho = (h - 1) / e1.stride + 1, wo = (w - 1) / e1.stride + 1;
    const long long opix = (long long)N * ho * wo;
    const int C = e1.cout, nxt = cur ^ 1;
    if (inorm) {
      SCF_TRY(tcconv_stats(u1, S(ws.xs[cur]), h, w, F(ws.raw), F(ws.stat)));
      SCF_TRY(in_apply(F(ws.raw), F(ws.stat), nullptr, nullptr, 1, nullptr, S(ws.ts), opix, ho * wo, C));
      SCF_TRY(tcconv_stats(u2, S(ws.ts), ho, wo, F(ws.raw), F(ws.stat)));
      if (ud >= 0) {
        SCF_TRY(tcconv_stats(ud, S(ws.xs[cur]), h, w, F(ws.raw2), F(ws.stat2)));
        SCF_TRY(in_apply(F(ws.raw), F(ws.stat), F(ws.raw2), F(ws.stat2), 1, nullptr, S(ws.xs[nxt]), opix, ho * wo, C));
      } else {
        SCF_TRY(in_apply(F(ws.raw), F(ws.stat), nullptr, nullptr, 1, nullptr, S(ws.xs[nxt]), opix, ho * wo, C, S(ws.xs[cur])));
      }
    } else {   // eval-mode BatchNorm folded into the convolutions: everything happens in the conv epilogues
      SCF_TRY(tcconv(u1, S(ws.xs[cur]), h, w, SCF_ACT_RELU, nullptr, S(ws.ts), nullptr));
      if (split_identity && bi < 2) {
        SCF_TRY(tcconv(u2, S(ws.ts), ho, wo, SCF_ACT_RELU, nullptr, S(ws.xs[nxt]), nullptr, nullptr, S(ws.xs[cur])));
      } else {
        const float* identity = F(ws.xf[cur]);
        if (ud >= 0) {
          SCF_TRY(tcconv(ud, S(ws.xs[cur]), h, w, SCF_ACT_NONE, F(ws.raw2), nullptr, nullptr));
          identity = F(ws.raw2);
        }
        SCF_TRY(tcconv(u2, S(ws.ts), ho, wo, SCF_ACT_RELU, F(ws.xf[nxt]), S(ws.xs[nxt]), identity));
      }
    }
    cur = nxt; h = ho; w = wo;
  }
  if (ex) {
    // ---- conv2 writing the consumer's layout directly (pixel-major split-bf16 / fp32, optional activation per channel half):
    // no NCHW round trip between the encoder and the refinement loop
    SCF_REQUIRE(ex->split == 256 || ex->split == 128, SCF_ERR_ARG, "scf_encoder_forward_ex: split must be 128 or 256");
    const long long opix = (long long)N * h * w;
    for (int part = 0; part < (ex->split == 256 ? 1 : 2); ++part) {
      scf_tc_conv_desc d = {};
      d.seg[0].ptr = S(ws.xs[cur]); d.seg[0].plane_stride = opix * 128; d.seg[0].stride = 128; d.seg[0].coff = 0; d.seg[0].nch = 128;
      d.nseg = 1;
      d.B = N; d.H = h; d.W = w; d.kh = d.kw = 1; d.stride = 1;
      const int cout = ex->split == 256 ? 256 : 128;
      d.w = ex->split == 256 ? pk + a.packed[15] : pk + a.packed15_half[part];
      d.cin_pad = 128; d.cout_pad = cout; d.cout = cout;
      d.bias = BIAS(15) + part * 128; d.scale = 1.f; d.epi = SCF_EPI_ACT;
      d.act = part == 0 ? ex->act0 : ex->act1;
      d.out_f32 = part == 0 ? ex->f32_0 : ex->f32_1; d.out_f32_stride = part == 0 ? ex->f32_stride0 : ex->f32_stride1;
      d.out_hl = part == 0 ? ex->hl0 : ex->hl1; d.out_hl_plane = part == 0 ? ex->plane0 : ex->plane1;
      d.out_hl_stride = part == 0 ? ex->stride0 : ex->stride1;
      SCF_REQUIRE(d.out_f32 || d.out_hl, SCF_ERR_ARG, "scf_encoder_forward_ex: output %d has no destination", part);
      SCF_TRY(conv2d_tc(d, st));
    }
    return 0;
  }
  // ---- conv2: 1x1 128 -> out channels, bias only; NHWC -> NCHW for the reference's output layout
  SCF_TRY(tcconv(15, S(ws.xs[cur]), h, w, SCF_ACT_NONE, F(ws.raw), nullptr, nullptr));
  SCF_TRY(scf_nhwc_to_nchw(F(ws.raw), 256, 0, out_nchw, N, 256, h, w, st));
  return 0;
}

int scf_encoder_forward(int norm, const void* packed, const float* images, int N, int H, int W, float* out_nchw,
                        void* workspace, size_t workspace_bytes, void* stream) {
  SCF_REQUIRE(out_nchw != nullptr, SCF_ERR_ARG, "scf_encoder_forward: null output");
  return encoder_forward_impl(norm, packed, images, N, H, W, out_nchw, nullptr, workspace, workspace_bytes, stream);
}
int scf_encoder_forward_ex(int norm, const void* packed, const float* images, int N, int H, int W, const scf_encoder_out* out,
                           void* workspace, size_t workspace_bytes, void* stream) {
  SCF_REQUIRE(out != nullptr, SCF_ERR_ARG, "scf_encoder_forward_ex: null output descriptor");
  return encoder_forward_impl(norm, packed, images, N, H, W, nullptr, out, workspace, workspace_bytes, stream);
}

}  // extern "C"
