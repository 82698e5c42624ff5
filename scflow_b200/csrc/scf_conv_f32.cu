// fp32 implicit-GEMM convolution on CUDA cores (exact fp32 FMA accumulate), NHWC activations.
//
// Role in the design: the general-shape convolution of the library. It carries the layers
// whose contraction is too thin for tensor cores (7x7 2->128: K=98, 3x3 1->64: K=9, 3x3 256->2, 1x1 256->1), the
// stride-2 pose-head convolutions, and - with precision=0 - every convolution of the decoder as the exact-fp32
// comparison path for the tcgen05 split-bf16 kernels (scf_conv_tc.cu).
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin, k = (ky*kw + kx)*Cin + c.  The input may be
// the channel-concatenation of up to three NHWC buffers ("segments"), which removes every torch.cat of the
// reference loop (raft_decoder.py:165,166,248,251; scflow_decoder.py:207,219).
#include "scf_common.cuh"
#include "scf_tc.cuh"

namespace scf {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int APAD = 4;

struct ConvParams {
  scf_conv_desc d;
  int M, K, cin;
  int seg_end0, seg_end1;  // cumulative channel ends of segments 0 and 1
};

__device__ __forceinline__ const float* seg_ptr(const ConvParams& p, int c, long long pix) {
  // channel c of the concatenated input at input pixel `pix`
  if (c < p.seg_end0) return p.d.seg[0].ptr + pix * p.d.seg[0].stride + p.d.seg[0].coff + c;
  if (c < p.seg_end1) return p.d.seg[1].ptr + pix * p.d.seg[1].stride + p.d.seg[1].coff + (c - p.seg_end0);
  return p.d.seg[2].ptr + pix * p.d.seg[2].stride + p.d.seg[2].coff + (c - p.seg_end1);
}

template <int VEC>
__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvParams p) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const int n0 = blockIdx.y * BN;
  const scf_conv_desc& d = p.d;
  const int HoWo = d.Ho * d.Wo;
  // per-sample weights (correlation build): tiles never straddle samples; blockIdx.z = sample
  const bool per_sample = d.w_batch_stride != 0;
  const int m0 = per_sample ? blockIdx.z * HoWo + blockIdx.x * BM : blockIdx.x * BM;
  const int m_end = per_sample ? (blockIdx.z + 1) * HoWo : p.M;

  // ---- A-load bookkeeping: which output pixel / k-lanes this thread fetches
  int a_ml, a_k[4];
  if (VEC == 4) {
    a_ml = t >> 2;
    a_k[0] = (t & 3) * 4;
  } else {
    a_ml = t & 63;
#pragma unroll
    for (int j = 0; j < 4; ++j) a_k[j] = (t >> 6) + 4 * j;
  }
  const int a_gm = m0 + a_ml;
  const bool a_valid = a_gm < m_end;
  int a_b = 0, a_oy = 0, a_ox = 0;
  if (a_valid) {
    a_b = a_gm / HoWo;
    int r = a_gm - a_b * HoWo;
    a_oy = r / d.Wo;
    a_ox = r - a_oy * d.Wo;
  }
  const int iy0 = a_oy * d.sh - d.ph, ix0 = a_ox * d.sw - d.pw;

  // ---- B-load bookkeeping
  const int b_row = t >> 4, b_n4 = (t & 15) * 4;
  const float* wbase = d.w;
  if (per_sample) wbase += (long long)blockIdx.z * d.w_batch_stride;

  float a_reg[4];
  float4 b_reg;

  auto load_tile = [&](int k0) {
    if (VEC == 4) {
      const int k = k0 + a_k[0];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_valid && k < p.K) {
        const int tap = k / p.cin, c = k - tap * p.cin;
        const int ky = tap / d.kw, kx = tap - ky * d.kw;
        const int iy = iy0 + ky, ix = ix0 + kx;
        if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
          v = __ldg(reinterpret_cast<const float4*>(seg_ptr(p, c, ((long long)a_b * d.Hi + iy) * d.Wi + ix)));
      }
      a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + a_k[j];
        float v = 0.f;
        if (a_valid && k < p.K) {
          const int tap = k / p.cin, c = k - tap * p.cin;
          const int ky = tap / d.kw, kx = tap - ky * d.kw;
          const int iy = iy0 + ky, ix = ix0 + kx;
          if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
            v = __ldg(seg_ptr(p, c, ((long long)a_b * d.Hi + iy) * d.Wi + ix));
        }
        a_reg[j] = v;
      }
    }
    const int kb = k0 + b_row, n = n0 + b_n4;
    b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kb < p.K && n < d.ldw) b_reg = __ldg(reinterpret_cast<const float4*>(wbase + (long long)kb * d.ldw + n));
  };
  auto store_tile = [&](int buf) {
    if (VEC == 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][a_k[0] + j][a_ml] = a_reg[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][a_k[j]][a_ml] = a_reg[j];
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_row][b_n4]) = b_reg;
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const int half = d.cout >> 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gm = m0 + ty * 4 + i;
    if (gm >= m_end) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.cout) continue;
      float v = acc[i][j] * d.scale + (d.bias ? __ldg(d.bias + n) : 0.f);
      if (d.epi == SCF_EPI_ACT) {
        v = act_apply(v, d.act);
        if (d.out) d.out[gm * d.out_stride + d.out_coff + n] = v;
        if (d.out_hl) {
          __nv_bfloat16 hi, lo;
          tc::split_bf16(v, hi, lo);
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(d.out_hl) + gm * d.out_hl_stride + d.out_hl_coff + n;
          o[0] = hi;
          o[d.out_hl_plane] = lo;
        }
      } else if (d.epi == SCF_EPI_GRU_ZR) {
        const float s = 1.f / (1.f + expf(-v));
        if (n < half) d.out[gm * d.out_stride + d.out_coff + n] = s;
        else d.out2[gm * d.out2_stride + (n - half)] = s * __ldg(d.aux0 + gm * d.aux0_stride + (n - half));
      } else {  // SCF_EPI_GRU_Q
        const float q = tanhf(v);
        const float h = __ldg(d.aux0 + gm * d.aux0_stride + n);
        const float z = __ldg(d.aux1 + gm * d.aux1_stride + n);
        d.out[gm * d.out_stride + d.out_coff + n] = (1.f - z) * h + z * q;
      }
    }
  }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int O, int I, int taps,
                                        int ldw, int o_off) {
  // packed[(tap*I + i)*ldw + o_off + o] = w[(o*I + i)*taps + tap]
  const long long total = (long long)O * I * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % O);
    const long long r = idx / O;
    const int i = (int)(r % I);
    const int tap = (int)(r / I);
    packed[((long long)tap * I + i) * ldw + o_off + o] = w[((long long)o * I + i) * taps + tap];
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW,
                                    int dst_stride, int dst_coff) {
  // 32x32 smem transpose per (b, c-tile, p-tile)
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = src[((long long)b * C + c) * HW + p];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[((long long)b * HW + p) * dst_stride + dst_coff + c] = tile[threadIdx.x][i];
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int src_stride, int src_coff, float* __restrict__ dst,
                                    int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = src[((long long)b * HW + p) * src_stride + src_coff + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((long long)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

int conv2d_thin(const scf_conv_desc& d, int in_nchw, cudaStream_t st);

int conv2d_f32(const scf_conv_desc& d, cudaStream_t st) {
  SCF_REQUIRE(d.nseg >= 1 && d.nseg <= 3, SCF_ERR_ARG, "scf_conv2d: nseg must be 1..3");
  if (d.nseg == 1 && d.seg[0].nch <= 4 && d.seg[0].ptr && d.w && (d.out || d.out_hl) && d.B > 0 && d.cout > 0) {
    const int rc = conv2d_thin(d, 0, st);       // register-tiled direct kernel for thin inputs
    if (rc != -100) return rc;
  }
  SCF_REQUIRE(d.w && (d.out || (d.out_hl && d.epi == SCF_EPI_ACT)) && d.B > 0 && d.cout > 0, SCF_ERR_ARG,
              "scf_conv2d: null pointer or empty shape");
  SCF_REQUIRE(d.ldw >= d.cout && d.ldw % 4 == 0, SCF_ERR_ARG, "scf_conv2d: ldw must be >= cout and a multiple of 4");
  SCF_REQUIRE(d.epi >= SCF_EPI_ACT && d.epi <= SCF_EPI_GRU_Q, SCF_ERR_ARG, "scf_conv2d: bad epilogue");
  if (d.epi == SCF_EPI_GRU_ZR) SCF_REQUIRE(d.aux0 && d.out2 && d.cout % 2 == 0, SCF_ERR_ARG, "scf_conv2d: GRU_ZR needs aux0/out2");
  if (d.epi == SCF_EPI_GRU_Q) SCF_REQUIRE(d.aux0 && d.aux1, SCF_ERR_ARG, "scf_conv2d: GRU_Q needs aux0 (h) and aux1 (z)");
  ConvParams p;
  p.d = d;
  p.cin = 0;
  bool vec = true;
  for (int s = 0; s < d.nseg; ++s) {
    SCF_REQUIRE(d.seg[s].ptr && d.seg[s].nch > 0, SCF_ERR_ARG, "scf_conv2d: bad segment %d", s);
    p.cin += d.seg[s].nch;
    vec = vec && d.seg[s].nch % 4 == 0 && d.seg[s].coff % 4 == 0 && d.seg[s].stride % 4 == 0 &&
          (reinterpret_cast<uintptr_t>(d.seg[s].ptr) % 16 == 0);
  }
  p.seg_end0 = d.seg[0].nch;
  p.seg_end1 = d.nseg > 1 ? p.seg_end0 + d.seg[1].nch : 0x7fffffff;
  if (d.nseg == 1) p.seg_end0 = 0x7fffffff;
  p.M = d.B * d.Ho * d.Wo;
  p.K = d.kh * d.kw * p.cin;
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d: packed weight must be 16B aligned");
  dim3 grid(cdiv(p.M, BM), cdiv(d.cout, BN));
  if (d.w_batch_stride != 0) {
    SCF_REQUIRE(d.w_batch_stride % 4 == 0, SCF_ERR_ALIGN, "scf_conv2d: w_batch_stride must be a multiple of 4");
    grid = dim3(cdiv(d.Ho * d.Wo, BM), cdiv(d.cout, BN), d.B);
  }
  if (vec) conv_f32_kernel<4><<<grid, 256, 0, st>>>(p);
  else conv_f32_kernel<1><<<grid, 256, 0, st>>>(p);
  return check_launch("conv_f32_kernel");
}

}  // namespace scf

extern "C" {

int scf_conv2d(const scf_conv_desc* d, void* stream) {
  SCF_REQUIRE(d != nullptr, SCF_ERR_ARG, "scf_conv2d: null descriptor");
  return scf::conv2d_f32(*d, (cudaStream_t)stream);
}

int scf_pack_conv_weight(const float* w_oihw, float* packed, int O, int I, int kh, int kw, int ldw, int o_off,
                         void* stream) {
  SCF_REQUIRE(w_oihw && packed && O > 0 && I > 0 && kh > 0 && kw > 0, SCF_ERR_ARG, "scf_pack_conv_weight: bad args");
  SCF_REQUIRE(ldw >= o_off + O, SCF_ERR_ARG, "scf_pack_conv_weight: ldw < o_off + O");
  const long long total = (long long)O * I * kh * kw;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  scf::pack_conv_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, packed, O, I, kh * kw, ldw, o_off);
  return scf::check_launch("pack_conv_weight_kernel");
}

int scf_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, int dst_stride, int dst_coff,
                     void* stream) {
  SCF_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_nchw_to_nhwc: bad args");
  dim3 grid(scf::cdiv(H * W, 32), scf::cdiv(C, 32), B);
  scf::nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, C, H * W, dst_stride, dst_coff);
  return scf::check_launch("nchw_to_nhwc_kernel");
}

int scf_nhwc_to_nchw(const float* src, int src_stride, int src_coff, float* dst, int B, int C, int H, int W,
                     void* stream) {
  SCF_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_nhwc_to_nchw: bad args");
  dim3 grid(scf::cdiv(H * W, 32), scf::cdiv(C, 32), B);
  scf::nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, src_stride, src_coff, dst, C, H * W);
  return scf::check_launch("nhwc_to_nchw_kernel");
}

}  // extern "C"
