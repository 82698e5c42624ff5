// Direct convolution for thin inputs (Cin <= 4): the 7x7 stride-2 RGB stem of the RAFT encoder, the 7x7 flow / delta-flow
// encoders (Cin = 2) and the 3x3 mask encoder (Cin = 1).  K = k*k*Cin is 9..147 - too thin for the tensor cores - so
// this is a register-tiled fp32 FMA kernel: a CTA computes 32x8 output pixels x 64 output channels from an input
// patch and a weight slab staged in shared memory; each thread owns 4 consecutive pixels x 16 channels.
#include "scf_common.cuh"
#include "scf_tc.cuh"

namespace scf {

constexpr int TH_TPX = 32, TH_TPY = 8, TH_CO = 64;

struct ThinParams {
  const float* in; int in_nchw; int in_stride, in_coff;    // NHWC: floats per pixel / first channel; NCHW: planes
  int B, Hi, Wi, Ho, Wo, pad;
  const float* w; int ldw;          // packed [k*k*CIN][ldw]
  const float* bias; int cout, act;
  float* out; int out_stride, out_coff;
  __nv_bfloat16* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff;
};

template <int CIN, int K, int STRIDE>
__global__ void __launch_bounds__(256) conv_thin_kernel(const ThinParams p) {
  constexpr int PW = (TH_TPX - 1) * STRIDE + K, PH = (TH_TPY - 1) * STRIDE + K;   // input patch
  constexpr int KK = K * K * CIN;
  extern __shared__ __align__(16) float smem[];
  float* patch = smem;                        // [PH][PW][CIN]
  float* wsm = smem + ((PH * PW * CIN + 3) & ~3);   // [KK][TH_CO]
  const int tiles_x = (p.Wo + TH_TPX - 1) / TH_TPX, tiles_y = (p.Ho + TH_TPY - 1) / TH_TPY;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int tr = blockIdx.x - b * tiles_x * tiles_y;
  const int ox0 = (tr % tiles_x) * TH_TPX, oy0 = (tr / tiles_x) * TH_TPY;
  const int co0 = blockIdx.y * TH_CO;
  const int ix0 = ox0 * STRIDE - p.pad, iy0 = oy0 * STRIDE - p.pad;
  // ---- stage the patch (zero padded) and the weight slab
  for (int idx = threadIdx.x; idx < PH * PW * CIN; idx += 256) {
    int c, x, y;
    if (p.in_nchw) { x = idx % PW; const int r = idx / PW; y = r % PH; c = r / PH; }
    else { c = idx % CIN; const int r = idx / CIN; x = r % PW; y = r / PW; }
    const int iy = iy0 + y, ix = ix0 + x;
    float v = 0.f;
    if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
      v = p.in_nchw ? __ldg(p.in + (((long long)b * CIN + c) * p.Hi + iy) * p.Wi + ix)
                    : __ldg(p.in + (((long long)b * p.Hi + iy) * p.Wi + ix) * p.in_stride + p.in_coff + c);
    patch[(y * PW + x) * CIN + c] = v;
  }
  for (int idx = threadIdx.x; idx < KK * (TH_CO / 4); idx += 256) {
    const int k = idx / (TH_CO / 4), c4 = idx - k * (TH_CO / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co0 + c4 * 4 < p.ldw) v = __ldg(reinterpret_cast<const float4*>(p.w + (long long)k * p.ldw + co0) + c4);
    reinterpret_cast<float4*>(wsm)[idx] = v;
  }
  __syncthreads();
  // ---- thread tile: 4 consecutive x-pixels x 16 channels; channel group is warp-uniform (weight loads broadcast)
  const int cg = threadIdx.x >> 6, pg = threadIdx.x & 63;
  const int py = pg >> 3, px = (pg & 7) * 4;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  const float* prow = patch + (py * STRIDE * PW + px * STRIDE) * CIN;
  const float* wcol = wsm + cg * 16;
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float* wp = wcol + ((ky * K + kx) * CIN + c) * TH_CO;
        float wv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(wp + 4 * j);
          wv[4 * j] = t.x; wv[4 * j + 1] = t.y; wv[4 * j + 2] = t.z; wv[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xv = prow[(ky * PW + kx + i * STRIDE) * CIN + c];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
        }
      }
    }
  }
  // ---- epilogue: bias + activation, fp32 and/or split-bf16 NHWC
  const int oy = oy0 + py;
  if (oy >= p.Ho) return;
  const int nb = co0 + cg * 16;
  const bool full = nb + 16 <= p.cout;
  float bv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) bv[j] = (p.bias && nb + j < p.cout) ? __ldg(p.bias + nb + j) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + px + i;
    if (ox >= p.Wo) continue;
    const long long pix = ((long long)b * p.Ho + oy) * p.Wo + ox;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_apply(acc[i][j] + bv[j], p.act);
    if (p.out) {
      float* o = p.out + pix * p.out_stride + p.out_coff + nb;
      if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (nb + j < p.cout) o[j] = v[j];
      }
    }
    if (p.out_hl) {
      __nv_bfloat16* o = p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __bfloat1622float2(h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && ((p.out_hl_plane * 2) & 15) == 0) {
        reinterpret_cast<uint4*>(o)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        reinterpret_cast<uint4*>(o + p.out_hl_plane)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        reinterpret_cast<uint4*>(o + p.out_hl_plane)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      } else {
        unsigned short* oh = reinterpret_cast<unsigned short*>(o);
        unsigned short* ol = reinterpret_cast<unsigned short*>(o + p.out_hl_plane);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.cout) {
            oh[j] = (unsigned short)(hi[j >> 1] >> ((j & 1) * 16));
            ol[j] = (unsigned short)(lo[j >> 1] >> ((j & 1) * 16));
          }
      }
    }
  }
}

template <int CIN, int K, int STRIDE>
static int launch_thin(const ThinParams& p, cudaStream_t st) {
  constexpr int PW = (TH_TPX - 1) * STRIDE + K, PH = (TH_TPY - 1) * STRIDE + K;
  const int smem = (((PH * PW * CIN + 3) & ~3) + K * K * CIN * TH_CO) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    SCF_CUDA(cudaFuncSetAttribute(conv_thin_kernel<CIN, K, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int tiles = cdiv(p.Wo, TH_TPX) * cdiv(p.Ho, TH_TPY) * p.B;
  conv_thin_kernel<CIN, K, STRIDE><<<dim3(tiles, cdiv(p.cout, TH_CO)), 256, smem, st>>>(p);
  return check_launch("conv_thin_kernel");
}

// Returns -100 when the shape is not covered (caller falls back to the generic implicit-GEMM kernel).
int conv2d_thin(const scf_conv_desc& d, int in_nchw, cudaStream_t st) {
  if (d.nseg != 1 || d.epi != SCF_EPI_ACT || d.w_batch_stride != 0 || d.scale != 1.f || d.kh != d.kw || d.sh != d.sw ||
      d.ph != d.pw || d.ph != d.kh / 2)
    return -100;
  if (d.ldw % 4 != 0 || reinterpret_cast<uintptr_t>(d.w) % 16 != 0) return -100;
  ThinParams p;
  p.in = d.seg[0].ptr; p.in_nchw = in_nchw; p.in_stride = d.seg[0].stride; p.in_coff = d.seg[0].coff;
  p.B = d.B; p.Hi = d.Hi; p.Wi = d.Wi; p.Ho = d.Ho; p.Wo = d.Wo; p.pad = d.ph;
  p.w = d.w; p.ldw = d.ldw; p.bias = d.bias; p.cout = d.cout; p.act = d.act;
  p.out = d.out; p.out_stride = d.out_stride; p.out_coff = d.out_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.out_hl_plane = d.out_hl_plane; p.out_hl_stride = d.out_hl_stride;
  p.out_hl_coff = d.out_hl_coff;
  const int cin = d.seg[0].nch, k = d.kh, s = d.sh;
  if (cin == 3 && k == 7 && s == 2) return launch_thin<3, 7, 2>(p, st);
  if (cin == 2 && k == 7 && s == 1) return launch_thin<2, 7, 1>(p, st);
  if (cin == 1 && k == 3 && s == 1) return launch_thin<1, 3, 1>(p, st);
  return -100;
}

}  // namespace scf
