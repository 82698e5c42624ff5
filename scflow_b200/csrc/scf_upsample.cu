// Convex 8x up-sampling of the RAFT baseline decoders (models/decoder/raft_decoder.py:381-416
// RAFTDecoder._upsample with a predicted mask): out[n,c,8h+i,8w+j] = sum_k softmax_k(mask[n,k*64+i*8+j,h,w]) * 8*flow[n,c,h+ky-1,w+kx-1]
// over the 3x3 neighbourhood k = ky*3+kx (zero padding, F.unfold); RAFTDecoderMask.upsample_mask (raft_decoder_mask.py:
// 143-162) is the same combination of a 1-channel occlusion map without the x8 factor.  HBM-bound: 576 mask floats in and 128 floats out per
// coarse pixel.  One block per (sample, coarse row, sub-row i, 32 coarse columns): the 9 x 8 mask rows are read as full
// 128 B lines, the fine output row is written as one contiguous run.
#include "scf_common.cuh"

namespace scf {

__global__ void __launch_bounds__(256) convex_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                                                              float* __restrict__ out, int C, int H, int W, float mul) {
  __shared__ float sm[9][8][33];      // mask logits [k][j][coarse column]
  __shared__ float sf[2][3][34];      // 8 * flow, rows h-1..h+1, columns w0-1..w0+32
  const int w0 = blockIdx.x * 32, h = blockIdx.y >> 3, i = blockIdx.y & 7, n = blockIdx.z;
  const long long HW = (long long)H * W;
  for (int idx = threadIdx.x; idx < 72 * 32; idx += 256) {
    const int r = idx >> 5, wl = idx & 31, k = r >> 3, j = r & 7, w = w0 + wl;
    sm[k][j][wl] = w < W ? __ldg(mask + ((long long)n * 576 + k * 64 + i * 8 + j) * HW + (long long)h * W + w) : 0.f;
  }
  for (int idx = threadIdx.x; idx < C * 3 * 34; idx += 256) {
    const int c = idx / 102, rem = idx - c * 102, ry = rem / 34, cx = rem - ry * 34;
    const int hh = h + ry - 1, ww = w0 + cx - 1;
    const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
    sf[c][ry][cx] = in ? mul * __ldg(flow + ((long long)n * C + c) * HW + (long long)hh * W + ww) : 0.f;
  }
  __syncthreads();
  const int wl = threadIdx.x >> 3, j = threadIdx.x & 7;
  if (w0 + wl >= W) return;
  float e[9], m = sm[0][j][wl];
#pragma unroll
  for (int k = 1; k < 9; ++k) m = fmaxf(m, sm[k][j][wl]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = expf(sm[k][j][wl] - m); s += e[k]; }
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float wgt = e[k] / s;
    o0 = fmaf(wgt, sf[0][k / 3][wl + k % 3], o0);
    if (C > 1) o1 = fmaf(wgt, sf[1][k / 3][wl + k % 3], o1);
  }
  const long long oW = 8LL * W, oHW = 64LL * HW;
  float* op = out + (long long)n * C * oHW + (8LL * h + i) * oW + 8LL * (w0 + wl) + j;
  op[0] = o0;
  if (C > 1) op[oHW] = o1;
}

}  // namespace scf

extern "C" {

int scf_convex_upsample(const float* x, const float* mask, float* out, int B, int C, int H, int W, float mul, void* stream) {
  using namespace scf;
  SCF_REQUIRE(x && mask && out, SCF_ERR_ARG, "scf_convex_upsample: null pointer");
  SCF_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * 8 <= 65535 && (C == 1 || C == 2), SCF_ERR_ARG,
              "scf_convex_upsample: bad shape (C must be 1 or 2)");
  convex_upsample_kernel<<<dim3(cdiv(W, 32), H * 8, B), 256, 0, (cudaStream_t)stream>>>(x, mask, out, C, H, W, mul);
  return check_launch("convex_upsample_kernel");
}

}  // extern "C"
