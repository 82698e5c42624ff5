// Fully connected layers of the pose regressor (models/head/pose_head.py:203-210: fc0 2048 -> 1024, fc1 1024 -> 256, ReLU after
// each) on tcgen05, split-K.  With at most 32 samples per batch shard an FC layer is a weight stream with almost no arithmetic;
// the CUDA-core form (linear_smem_kernel) is bound by re-staging x in every block and by its shared-memory reads (25 + 14 us per
// refinement iteration).  Here block (mo, ks) contracts the K range [ks * KR, (ks + 1) * KR) for 128 output rows:
//   * A = the weight rows, pre-split bf16 [2][O][I] (K-major), all KR / 64 chunks TMA-loaded up front (<= 128 KB resident);
//   * B = the 32 samples: every thread of the block forms x = act(sum_s partial_s + bias) for 8 consecutive inputs of one sample
//     (the PREVIOUS layer's split-K partial sums, its bias and ReLU are applied here, so no reduction kernel runs in between),
//     splits it to bf16 hi / lo and writes the 16 B unit into the SWIZZLE_128B operand tile;
//   * KR / 16 k-steps x 3 products (hi*hi + hi*lo + lo*hi) of M = 128, N = 32 into one 32-column TMEM accumulator;
//   * epilogue, split-K form: lane = output row, column = sample; the raw partial sums leave as part[ks][b][o] (coalesced along o);
//     bias and activation of THIS layer are applied by the consumer (the next FC layer or the pose projection);
//   * epilogue, reduced form (I / KR = 8): the eight K-range blocks of a row tile are one thread-block cluster; each parks its
//     partial tile in shared memory and block `rank` sums samples [4 rank, 4 rank + 4) over the eight tiles through distributed
//     shared memory in fixed order, adds bias, applies ReLU and writes y[b][o]: one launch = one complete layer, deterministic.
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz);

constexpr int FC_M = 128, FC_N = 32, FC_MAXCH = 4;                 // up to 4 chunks of 64 inputs per block (KR <= 256)
constexpr uint32_t FC_WPLANE = FC_M * 128u, FC_WCHUNK = 2 * FC_WPLANE;      // 32 KB per chunk (hi + lo)
constexpr uint32_t FC_XPLANE = FC_N * 128u, FC_XCHUNK = 2 * FC_XPLANE;      // 8 KB per chunk
constexpr int FC_THREADS = 160;
constexpr int FC_SMEM = 1024 + 1024 + FC_MAXCH * (int)(FC_WCHUNK + FC_XCHUNK);

struct FcIn { const float* p; int nsplit; long long split_stride; const float* bias; int relu; };
struct FcParams {
  FcIn x;
  float* part;           // [ks][B][O] raw partial sums (split-K mode)
  float* y;              // [B][O] = act(sum_ks + bias) (cluster mode: the I / KR blocks of a row tile form one cluster)
  const float* bias; int relu;
  int B, I, O, KR;
};

__device__ __forceinline__ float4 fc_load4(const FcIn& in, long long off, int col) {
  float4 v = __ldg(reinterpret_cast<const float4*>(in.p + off));
  for (int s = 1; s < in.nsplit; ++s) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(in.p + s * in.split_stride + off));
    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
  }
  if (in.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(in.bias + col));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (in.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  return v;
}

__global__ void __launch_bounds__(FC_THREADS, 1)
fc_tc_kernel(const __grid_constant__ CUtensorMap tmW, const FcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_w = sb, bar_d = sb + 8, tmem_slot = sb + 16;
  const uint32_t w0 = sb + 1024, x0 = w0 + FC_MAXCH * FC_WCHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mo = blockIdx.x, ks = blockIdx.y;
  const int nch = p.KR / 64, k0 = ks * p.KR;
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    mbar_init(bar_w, 1);
    mbar_init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 32u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, (uint32_t)nch * FC_WCHUNK);
    for (int c = 0; c < nch; ++c)
      for (int pl = 0; pl < 2; ++pl)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(w0 + c * FC_WCHUNK + pl * FC_WPLANE), "l"(reinterpret_cast<uint64_t>(&tmW)), "r"(bar_w), "r"(k0 + c * 64),
                       "r"(mo * FC_M), "r"(pl) : "memory");
  }
  // ---- B operand: unit = (sample b, 8 consecutive inputs); 16 B per plane into the swizzled [32 rows][128 B] chunk tile
  const int units = FC_N * (p.KR / 8);
  for (int u = threadIdx.x; u < units; u += FC_THREADS) {
    const int b = u / (p.KR / 8), k8 = u - b * (p.KR / 8);
    float v[8];
    if (b < p.B) {
      const int col = k0 + k8 * 8;
      const float4 a = fc_load4(p.x, (long long)b * p.I + col, col), c4 = fc_load4(p.x, (long long)b * p.I + col + 4, col + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * i], h0, l0);
      split_bf16(v[2 * i + 1], h1, l1);
      hw[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lw[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const int c = k8 >> 3, j = k8 & 7;
    const uint32_t addr = x0 + (uint32_t)c * FC_XCHUNK + (uint32_t)b * 128u + (uint32_t)((j ^ (b & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + FC_XPLANE), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(FC_M, FC_N);
    mbar_wait(bar_w, 0);
    tc_fence_after();
    for (int c = 0; c < nch; ++c) {
      const uint32_t wa = w0 + c * FC_WCHUNK, xa = x0 + c * FC_XCHUNK;
      const uint64_t w_hi = make_smem_desc_sw128(wa, 1024), w_lo = make_smem_desc_sw128(wa + FC_WPLANE, 1024);
      const uint64_t x_hi = make_smem_desc_sw128(xa, 1024), x_lo = make_smem_desc_sw128(xa + FC_XPLANE, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ko = (uint64_t)(k * 32 >> 4);
        umma_bf16(tmem_base, w_hi + ko, x_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
        umma_bf16(tmem_base, w_hi + ko, x_lo + ko, idesc, 1u);
        umma_bf16(tmem_base, w_lo + ko, x_hi + ko, idesc, 1u);
      }
    }
    umma_commit(bar_d);
  }
  if (warp >= 1) {
    const int q = warp & 3;
    const int o = mo * FC_M + q * 32 + lane;
    mbar_wait(bar_d, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
    if (!p.y) {
      if (o < p.O) {
        float* dst = p.part + ((long long)ks * p.B) * p.O + o;
#pragma unroll
        for (int b = 0; b < FC_N; ++b)
          if (b < p.B) dst[(long long)b * p.O] = v[b];
      }
    } else {
      // cluster mode: park this block's partial tile [32 samples][128 rows] in shared memory (the operand tiles are dead: every
      // MMA has completed) for the cross-block reduction below
#pragma unroll
      for (int b = 0; b < FC_N; ++b)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(x0 + (uint32_t)((b * FC_M + q * 32 + lane) * 4)), "f"(v[b]) : "memory");
    }
  }
  if (p.y) {
    // split-K reduction through distributed shared memory: block `rank` of the cluster sums samples [4 * rank, 4 * rank + 4) over
    // the eight partial tiles in fixed order (deterministic), adds the bias, applies the activation and writes the layer's output
    __syncwarp();
    cluster_sync_all();
    const uint32_t rank = cluster_ctarank();
    if (warp >= 1) {
      const int ol = (warp & 3) * 32 + lane, o = mo * FC_M + ol;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = (int)rank * 4 + j;
        float acc = 0.f;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
          float t;
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t) : "r"(mapa_shared(x0 + (uint32_t)((b * FC_M + ol) * 4), k)) : "memory");
          acc += t;
        }
        if (b < p.B && o < p.O) {
          acc += p.bias ? __ldg(p.bias + o) : 0.f;
          p.y[(long long)b * p.O + o] = p.relu ? fmaxf(acc, 0.f) : acc;
        }
      }
    }
    __syncwarp();
    cluster_sync_all();          // peers may still be reading this block's tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32u);
}

// part[ks][b][o] = sum over inputs [ks*KR, (ks+1)*KR) of W[o][i] * x[b][i];  x = act(sum_s xin.p[s] + xin.bias) (see FcIn).
// w_packed: split-bf16 [2][O][I] as written by scf_pack_conv_weight_tc(w, ., O, I, 1, 1, I, O, 0).
int fc_tc(const float* x, int x_nsplit, long long x_split_stride, const float* x_bias, int x_relu, const void* w_packed, float* part,
          float* y, const float* bias, int relu, int B, int I, int O, int KR, cudaStream_t st) {
  SCF_REQUIRE(x && w_packed && (part || y), SCF_ERR_ARG, "fc_tc: null pointer");
  SCF_REQUIRE(!y || I / KR == 8, SCF_ERR_ARG, "fc_tc: the reduced form needs exactly 8 K-ranges (one cluster of 8 blocks per row tile): I / KR = %d", I / (KR > 0 ? KR : 1));
  SCF_REQUIRE(B >= 1 && B <= FC_N && KR % 64 == 0 && KR >= 64 && KR <= 64 * FC_MAXCH && I % KR == 0 && O % 4 == 0, SCF_ERR_ARG,
              "fc_tc: needs B <= 32, KR a multiple of 64 up to 256 dividing I (B %d, I %d, O %d, KR %d)", B, I, O, KR);
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w_packed) % 16 == 0 && (x_split_stride % 4) == 0 &&
                  (!x_bias || reinterpret_cast<uintptr_t>(x_bias) % 16 == 0),
              SCF_ERR_ALIGN, "fc_tc: buffers must be 16B aligned");
  FcParams p = {};
  p.x.p = x; p.x.nsplit = x_nsplit < 1 ? 1 : x_nsplit; p.x.split_stride = x_split_stride; p.x.bias = x_bias; p.x.relu = x_relu;
  p.part = part; p.y = y; p.bias = bias; p.relu = relu; p.B = B; p.I = I; p.O = O; p.KR = KR;
  CUtensorMap tmW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)I, (cuuint64_t)O, 2};
    cuuint64_t str[2] = {(cuuint64_t)I * 2, (cuuint64_t)O * I * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)FC_M, 1};
    SCF_TRY(encode_map(&tmW, w_packed, 3, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM); });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(fc_tc_kernel): %s", cudaGetErrorString(attr_err));
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cdiv(O, FC_M), (unsigned)(I / KR)); cfg.blockDim = dim3(FC_THREADS);
  cfg.dynamicSmemBytes = FC_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (y) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 8; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, fc_tc_kernel, tmW, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("fc_tc_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("fc_tc_kernel");
}

}  // namespace scf

extern "C" {

int scf_linear_tc(const float* x, int x_nsplit, long long x_split_stride, const float* x_bias, int x_relu, const void* w_packed,
                  float* part, float* y, const float* bias, int relu, int B, int I, int O, int KR, void* stream) {
  return scf::fc_tc(x, x_nsplit, x_split_stride, x_bias, x_relu, w_packed, part, y, bias, relu, B, I, O, KR, (cudaStream_t)stream);
}

}  // extern "C"
