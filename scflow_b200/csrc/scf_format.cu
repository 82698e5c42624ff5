// Input formatting next to the hot path: what BaseRefiner.format_data_test does with the renderer's
// output before get_pose (models/refiner/base_refiner.py:96-107) - drop the alpha channel, NHWC -> NCHW, normalise with the
// dataset's mean / std, take the nearest face's depth and the silhouette mask - as ONE streaming pass instead of six
// PyTorch kernels.  HBM-bound: per pixel cin + zk floats in, 5 floats out; one pixel per thread, coalesced on both sides.
#include "scf_common.cuh"

namespace scf {

struct FormatNorm { float mean[3], std[3]; };

template <bool VEC4>
__global__ void __launch_bounds__(256) format_rendered_kernel(const float* __restrict__ images, int cin, const float* __restrict__ zbuf,
                                                              int zk, FormatNorm nm, float* __restrict__ out_img,
                                                              float* __restrict__ out_depth, float* __restrict__ out_mask,
                                                              long long HW, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, pix = i - b * HW;
    float r, g, bl;
    if (VEC4) {          // RGBA, 16 B aligned: one 128-bit load per pixel
      const float4 v = __ldg(reinterpret_cast<const float4*>(images) + i);
      r = v.x; g = v.y; bl = v.z;
    } else {
      const float* p = images + i * cin;
      r = __ldg(p); g = __ldg(p + 1); bl = __ldg(p + 2);
    }
    // (x - mean) / std exactly as torch evaluates it: one fp32 subtraction, one fp32 division, no contraction
    float* o = out_img + b * 3 * HW + pix;
    o[0] = __fdiv_rn(__fsub_rn(r, nm.mean[0]), nm.std[0]);
    o[HW] = __fdiv_rn(__fsub_rn(g, nm.mean[1]), nm.std[1]);
    o[2 * HW] = __fdiv_rn(__fsub_rn(bl, nm.mean[2]), nm.std[2]);
    const float d = __ldg(zbuf + i * zk);
    out_depth[i] = d;
    out_mask[i] = d > 0.f ? 1.f : 0.f;
  }
}

}  // namespace scf

extern "C" {

int scf_format_rendered(const float* images, int cin, const float* zbuf, int zk, const float* mean3, const float* std3,
                        float* out_images, float* out_depth, float* out_mask, int B, int H, int W, void* stream) {
  using namespace scf;
  SCF_REQUIRE(images && zbuf && mean3 && std3 && out_images && out_depth && out_mask, SCF_ERR_ARG, "scf_format_rendered: null pointer");
  SCF_REQUIRE(B > 0 && H > 0 && W > 0 && cin >= 3 && zk >= 1, SCF_ERR_ARG, "scf_format_rendered: bad shape (cin >= 3, zk >= 1)");
  FormatNorm nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3[c]; nm.std[c] = std3[c];
    SCF_REQUIRE(std3[c] != 0.f, SCF_ERR_ARG, "scf_format_rendered: std must be non-zero");
  }
  const long long HW = (long long)H * W, total = (long long)B * HW;
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want < 148LL * 16 ? want : 148LL * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (cin == 4 && reinterpret_cast<uintptr_t>(images) % 16 == 0)
    format_rendered_kernel<true><<<blocks, 256, 0, st>>>(images, cin, zbuf, zk, nm, out_images, out_depth, out_mask, HW, total);
  else
    format_rendered_kernel<false><<<blocks, 256, 0, st>>>(images, cin, zbuf, zk, nm, out_images, out_depth, out_mask, HW, total);
  return check_launch("format_rendered_kernel");
}

}  // extern "C"
