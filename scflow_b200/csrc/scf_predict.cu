// Predict layers of the flow and mask heads (models/decoder/raft_decoder.py:292-294, scflow_decoder.py:210-213) in one
// kernel: delta-flow = conv3x3(256 -> 2)(hidden_flow) and mask = sigmoid(conv1x1(256 -> 1)(hidden_mask)), both reading the
// merged split-bf16 hidden map [2][B*P][512] written by the heads convolution.  With 2 + 1 output channels these layers
// are pure activation streaming (a tensor-core tile would re-read the map once per tap for N = 16), so this is an fp32
// CUDA-core kernel: an 8x8 pixel tile + halo is staged per 64-channel chunk, 4 threads share a pixel.
#include "scf_common.cuh"
#include <cuda_bf16.h>

namespace scf {

constexpr int HP_T = 8, HP_HALO = HP_T + 2, HP_CH = 64;
// shared-memory layouts are padded so that the 8 lanes of a quarter-warp (2 pixels x 4 channel quarters) hit 8 distinct
// 16 B bank groups: a pixel row is 4 quarters x (16 + 4) floats, a weight tap row 4 quarters x (32 + 4) floats
constexpr int HP_QS = 20, HP_RS = 4 * HP_QS, HP_WQ = 36, HP_WT = 4 * HP_WQ;

__device__ __forceinline__ void bf16x8_sum(const uint4& h, const uint4& l, float* o) {
  const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
  const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __bfloat1622float2(hp[i]), b = __bfloat1622float2(lp[i]);
    o[2 * i] = a.x + b.x; o[2 * i + 1] = a.y + b.y;
  }
}

__global__ void __launch_bounds__(256) heads_predict_kernel(const __nv_bfloat16* __restrict__ hd, long long plane, int stride,
                                                            const float* __restrict__ wf, int ldwf, const float* __restrict__ bf,
                                                            const float* __restrict__ wm, int ldwm, const float* __restrict__ bm,
                                                            float* __restrict__ dflow, float* __restrict__ mask8, int B, int H, int W,
                                                            int hidden) {
  __shared__ __align__(16) float tile[HP_HALO * HP_HALO * HP_RS];
  __shared__ __align__(16) float wsm[9 * HP_WT];
  const int tiles_x = (W + HP_T - 1) / HP_T, tiles_y = (H + HP_T - 1) / HP_T;
  const int b = blockIdx.x / (tiles_x * tiles_y), tr = blockIdx.x - b * tiles_x * tiles_y;
  const int x0 = (tr % tiles_x) * HP_T, y0 = (tr / tiles_x) * HP_T;
  const int quarter = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const int px = pl & 7, py = pl >> 3;
  const int x = x0 + px, y = y0 + py;
  const bool valid = x < W && y < H;
  float a0 = 0.f, a1 = 0.f, am = 0.f;
  // Per 64-channel chunk every global load (patch, weights, the mask head's own pixel) is issued before the first use, and the
  // next chunk's loads are issued before the current chunk's arithmetic: the kernel is latency-bound (0.3 GFLOP, 67 MB from
  // L2), so what matters is the number of serialised L2 round trips - one per chunk instead of ~7.
  constexpr int NT = (HP_HALO * HP_HALO * (HP_CH / 8) + 255) / 256;      // patch items per thread (4)
  constexpr int NW = (9 * HP_CH + 255) / 256;                            // weight rows per thread (3)
  uint4 th[NT], tl[NT], mh[2], ml[2];
  float w0[NW], w1[NW], mw[16];
  auto issue = [&](int c0) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int idx = threadIdx.x + k * 256;
      const int c8 = idx % (HP_CH / 8), pp = idx / (HP_CH / 8);
      const int ty = pp / HP_HALO, tx = pp - ty * HP_HALO;
      const int iy = y0 + ty - 1, ix = x0 + tx - 1;
      th[k] = tl[k] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < HP_HALO * HP_HALO * (HP_CH / 8) && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const __nv_bfloat16* src = hd + (((long long)b * H + iy) * W + ix) * stride + c0 + c8 * 8;
        th[k] = __ldg(reinterpret_cast<const uint4*>(src));
        tl[k] = __ldg(reinterpret_cast<const uint4*>(src + plane));
      }
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int idx = threadIdx.x + k * 256;
      w0[k] = w1[k] = 0.f;
      if (idx < 9 * HP_CH) {
        const int tap = idx / HP_CH, c = idx - tap * HP_CH;
        const float* src = wf + (long long)(tap * hidden + c0 + c) * ldwf;
        w0[k] = __ldg(src); w1[k] = __ldg(src + 1);
      }
    }
    if (valid) {
      const __nv_bfloat16* src = hd + (((long long)b * H + y) * W + x) * stride + hidden + c0 + quarter * 16;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        mh[j] = __ldg(reinterpret_cast<const uint4*>(src + 8 * j));
        ml[j] = __ldg(reinterpret_cast<const uint4*>(src + plane + 8 * j));
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) mw[i] = __ldg(wm + (long long)(c0 + quarter * 16 + i) * ldwm);
  };
  issue(0);
  for (int c0 = 0; c0 < hidden; c0 += HP_CH) {
    __syncthreads();                       // the previous chunk's arithmetic has finished reading the shared-memory patch
    // stage the (8+2)^2 pixel patch of channels [c0, c0+64) as fp32 (hi + lo), zero outside the image
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int idx = threadIdx.x + k * 256;
      if (idx < HP_HALO * HP_HALO * (HP_CH / 8)) {
        const int c8 = idx % (HP_CH / 8), pp = idx / (HP_CH / 8);
        float v[8];
        bf16x8_sum(th[k], tl[k], v);
        float4* dst = reinterpret_cast<float4*>(tile + pp * HP_RS + (c8 >> 1) * HP_QS + (c8 & 1) * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    // weights of this chunk: wsm[tap][quarter][(c % 16) * 2 + o]
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int idx = threadIdx.x + k * 256;
      if (idx < 9 * HP_CH) {
        const int tap = idx / HP_CH, c = idx - tap * HP_CH;
        float* dst = wsm + tap * HP_WT + (c >> 4) * HP_WQ + (c & 15) * 2;
        dst[0] = w0[k]; dst[1] = w1[k];
      }
    }
    // mask head: 1x1 over channels hidden + [c0, c0+64) of the same pixel (this thread: 16 of them)
    if (valid) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float v[8];
        bf16x8_sum(mh[j], ml[j], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) am = fmaf(v[i], mw[8 * j + i], am);
      }
    }
    if (c0 + HP_CH < hidden) issue(c0 + HP_CH);       // in flight during this chunk's arithmetic
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - ky * 3;
      const float* tp = tile + ((py + ky) * HP_HALO + px + kx) * HP_RS + quarter * HP_QS;
      const float* wp = wsm + tap * HP_WT + quarter * HP_WQ;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(tp + 4 * j);
        const float4 w0v = *reinterpret_cast<const float4*>(wp + 8 * j), w1v = *reinterpret_cast<const float4*>(wp + 8 * j + 4);
        a0 = fmaf(v.x, w0v.x, a0); a1 = fmaf(v.x, w0v.y, a1);
        a0 = fmaf(v.y, w0v.z, a0); a1 = fmaf(v.y, w0v.w, a1);
        a0 = fmaf(v.z, w1v.x, a0); a1 = fmaf(v.z, w1v.y, a1);
        a0 = fmaf(v.w, w1v.z, a0); a1 = fmaf(v.w, w1v.w, a1);
      }
    }
  }
  // combine the 4 channel quarters of a pixel (adjacent lanes)
#pragma unroll
  for (int off = 1; off < 4; off <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, off);
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    am += __shfl_xor_sync(0xffffffffu, am, off);
  }
  if (valid && quarter == 0) {
    const long long pix = ((long long)b * H + y) * W + x;
    dflow[pix * 2] = a0 + __ldg(bf);
    dflow[pix * 2 + 1] = a1 + __ldg(bf + 1);
    mask8[pix] = 1.f / (1.f + expf(-(am + __ldg(bm))));
  }
}

// hd: split-bf16 [2][B*H*W][stride] with the flow-head hidden channels at [0, hidden) and the mask-head hidden channels at
// [hidden, 2*hidden); wf: packed fp32 [9*hidden][ldwf] (scf_pack_conv_weight layout, 2 outputs); wm: [hidden][ldwm] (1 output)
int heads_predict(const void* hd_hl, long long plane, int stride, int hidden, const float* wf, int ldwf, const float* bf,
                  const float* wm, int ldwm, const float* bm, float* dflow, float* mask8, int B, int H, int W, cudaStream_t st) {
  SCF_REQUIRE(hd_hl && wf && bf && wm && bm && dflow && mask8 && B > 0 && H > 0 && W > 0, SCF_ERR_ARG, "heads_predict: bad args");
  SCF_REQUIRE(hidden % HP_CH == 0 && stride % 8 == 0 && plane % 8 == 0 && reinterpret_cast<uintptr_t>(hd_hl) % 16 == 0, SCF_ERR_ALIGN,
              "heads_predict: hidden %% 64, stride %% 8 and 16B alignment required");
  const int blocks = B * ((W + HP_T - 1) / HP_T) * ((H + HP_T - 1) / HP_T);
  heads_predict_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(hd_hl), plane, stride, wf, ldwf, bf, wm, ldwm, bm,
                                               dflow, mask8, B, H, W, hidden);
  return check_launch("heads_predict_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Predict layers, re-associated so that the 512-channel hidden map is read ONCE, without a halo, by the tensor cores:
//   d[p][2 t + o] = <W_flow[o, :, tap t], hidden_flow[p, :]>  (t = 0..8, o = 0..1),   d[p][18] = <w_mask, hidden_mask[p, :]>
// is a 1x1 convolution 512 -> 19 (scf_conv2d_tc, weights packed by pack_predict_tc below); the 3x3 convolution is then the sum
// of its nine taps' partial products at the neighbouring pixels (zero outside the map = the convolution's zero padding):
//   dflow[p][o] = b_o + sum_t d[p + offset_t][2 t + o] ;  mask[p] = sigmoid(d[p][18] + b_m).
__global__ void pack_predict_tc_kernel(const float* __restrict__ wf, const float* __restrict__ wm, __nv_bfloat16* __restrict__ packed,
                                       int hidden, int cout_pad) {
  // packed: [2 planes][cout_pad rows][2 * hidden]
  const int total = cout_pad * 2 * hidden;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % (2 * hidden), o = idx / (2 * hidden);
    float v = 0.f;
    if (o < 18 && c < hidden) v = wf[(((o & 1) * hidden) + c) * 9 + (o >> 1)];       // OIHW [2, hidden, 3, 3], tap = ky*3 + kx
    else if (o == 18 && c >= hidden) v = wm[c - hidden];                              // [1, hidden, 1, 1]
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    packed[idx] = hi;
    packed[(long long)total + idx] = lo;
  }
}

__global__ void __launch_bounds__(256) predict_gather_kernel(const float* __restrict__ d, int ld, const float* __restrict__ bf,
                                                             const float* __restrict__ bm, float* __restrict__ dflow,
                                                             float* __restrict__ mask8, int H, int W, long long npix) {
  scf_pdl_enter();
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    float a0 = __ldg(bf), a1 = __ldg(bf + 1);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(d + (pix + (long long)(t / 3 - 1) * W + (t % 3 - 1)) * ld + 2 * t));
        a0 += v.x; a1 += v.y;
      }
    }
    reinterpret_cast<float2*>(dflow)[pix] = make_float2(a0, a1);
    mask8[pix] = 1.f / (1.f + expf(-(__ldg(d + pix * ld + 18) + __ldg(bm))));
  }
}

int pack_predict_tc(const float* wf_oihw, const float* wm_oihw, void* packed, int hidden, int cout_pad, cudaStream_t st) {
  SCF_REQUIRE(wf_oihw && wm_oihw && packed && hidden > 0 && cout_pad >= 19, SCF_ERR_ARG, "pack_predict_tc: bad args");
  pack_predict_tc_kernel<<<64, 256, 0, st>>>(wf_oihw, wm_oihw, reinterpret_cast<__nv_bfloat16*>(packed), hidden, cout_pad);
  return check_launch("pack_predict_tc_kernel");
}

int predict_gather(const float* d, int ld, const float* bf, const float* bm, float* dflow, float* mask8, int B, int H, int W,
                   cudaStream_t st) {
  SCF_REQUIRE(d && bf && bm && dflow && mask8 && ld >= 19 && ld % 2 == 0, SCF_ERR_ARG, "predict_gather: bad args");
  const long long npix = (long long)B * H * W;
  launch_pdl(predict_gather_kernel, dim3(cdiv(npix, 256)), dim3(256), 0, st, d, ld, bf, bm, dflow, mask8, H, W, npix);
  return check_launch("predict_gather_kernel");
}

}  // namespace scf
