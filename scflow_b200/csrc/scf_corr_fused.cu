// All-pairs correlation volume AND its pooled pyramid in one kernel (CorrelationPyramid.forward, raft_decoder.py:35-58):
//     level0[b*P + q][k] = <f_render[b, :, q], f_real[b, :, k]> / sqrt(C) ;  level l+1 = AvgPool2d(2, 2)(level l) over the k map
// A dense contraction on tcgen05 (split-bf16: hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) whose output - 5.57 MB of fp32
// per sample against 0.54 GFLOP - makes it HBM-bound: what matters is to write every level exactly once and never read one
// back.  The three pooled levels therefore come out of the SAME accumulator tile in the epilogue instead of three pooling
// kernels that re-read the 134 MB (B = 32) level-0 volume.
//
// Tile = 128 queries (TMEM lanes; one query per epilogue thread) x 256 keys = 8 complete rows of the 32-wide key map, so every
// 2x2 / 4x4 / 8x8 pooling window of the tile is inside ONE thread's accumulator row:
//   * operands: both feature maps are pixel-major split-bf16 [2][B*P][C]; per 32-channel chunk one TMA box of the query rows
//     and one of the key rows (SWIZZLE_64B, four 48 KB stages); 8 chunks x 2 k-steps x 3 products = 48 MMAs (M 128, N 256);
//   * two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i + 1 (the kernel is epilogue / store bound);
//   * epilogue: the two warps of a lane quarter take key rows 0-3 and 4-7.  Per key row: 32 columns -> registers, x 1/sqrt(C),
//     level 0 leaves as a [32 queries][32 keys] block through shared memory and one TMA store; pairs of key
//     rows give a level-1 row (16 floats, TMA store as well), pairs of those a level-2 row (8 floats, one sector store) - same summation order as the reference's
//     successive pools: ((a + b) + c) + d, then x 0.25.  Level 3 needs both warps' level-2 rows: 8 floats per thread cross
//     through shared memory.
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz);

constexpr int CB_M = 128, CB_N = 256, CB_BK = 32, CB_STAGES = 3, CB_EW = 8, CB_W8 = 32;
constexpr uint32_t CB_ROWB = CB_BK * 2;                       // 64 B rows (SWIZZLE_64B)
constexpr uint32_t CB_A_PLANE = CB_M * CB_ROWB, CB_B_PLANE = CB_N * CB_ROWB;
constexpr uint32_t CB_STAGE = 2 * (CB_A_PLANE + CB_B_PLANE);  // 48 KB
constexpr uint32_t CB_XCHG = 2 * 4 * 32 * 8 * 4;              // level-2 rows crossing between the two warps of a quarter, two tile parities
constexpr uint32_t CB_STG = 4096 + 2048;                  // per epilogue warp: a [32 queries][32 keys] fp32 level-0 block + a [32][16] level-1 block
constexpr int CB_SMEM = 1024 + 1024 + (int)CB_XCHG + CB_EW * (int)CB_STG + CB_STAGES * (int)CB_STAGE;
static_assert(CB_SMEM <= 232448, "corr_pyramid_kernel does not fit in shared memory");

struct CorrFusedParams {
  int B, P, C, H8;                 // P = H8 * 32 keys = queries per sample
  int q_tiles, k_tiles, num_tiles; // per sample: P/128 query tiles x P/256 key tiles
  float scale;
  float* lvl0; float* lvl1; float* lvl2; float* lvl3;
  int dbg;                         // timing experiments (SCFLOW_CORR_DBG): 1 no MMAs, 2 no global stores, 4 no operand loads
};

__device__ __forceinline__ void st_global_v8(float* ptr, const float* v, bool on = true) {
  if (!on && v[0] != 123.456f) return;
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

__global__ void __launch_bounds__(64 + 32 * CB_EW, 1)
corr_pyramid_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmL0, const __grid_constant__ CUtensorMap tmL1, const CorrFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 64, bar_tfull = smem_base + 128, bar_tempty = smem_base + 144,
                 tmem_slot = smem_base + 192;
  const uint32_t xchg0 = smem_base + 1024;
  const uint32_t stg0 = xchg0 + CB_XCHG;
  const uint32_t ring0 = stg0 + CB_EW * CB_STG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmL0); prefetch_tmap(&tmL1);
    for (int s = 0; s < CB_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, CB_EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  const int chunks = p.C / CB_BK;
  const int per_sample = p.q_tiles * p.k_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int b = t / per_sample, r = t - b * per_sample;
        const int qt = r / p.k_tiles, kt = r - qt * p.k_tiles;            // key tile fastest: the query rows stay hot in L2
        const int qrow = b * p.P + qt * CB_M, krow = b * p.P + kt * CB_N;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage, dst = ring0 + stage * CB_STAGE;
          if (p.dbg & 4) { mbar_arrive(full); if (++stage == CB_STAGES) { stage = 0; phase ^= 1u; } continue; }
          mbar_arrive_expect_tx(full, CB_STAGE);
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmQ)), "r"(full), "r"(c * CB_BK), "r"(qrow), "r"(0) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst + 2 * CB_A_PLANE), "l"(reinterpret_cast<uint64_t>(&tmK)), "r"(full), "r"(c * CB_BK), "r"(krow), "r"(0) : "memory");
          if (++stage == CB_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(CB_M, CB_N);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_tempty + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CB_N);
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = ring0 + stage * CB_STAGE, b_addr = a_addr + 2 * CB_A_PLANE;
          const uint64_t a_hi = make_smem_desc_sw64(a_addr, 512), a_lo = make_smem_desc_sw64(a_addr + CB_A_PLANE, 512);
          const uint64_t b_hi = make_smem_desc_sw64(b_addr, 512), b_lo = make_smem_desc_sw64(b_addr + CB_B_PLANE, 512);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (p.dbg & 1) break;
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
            umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == CB_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * acc);
      }
    }
  } else {
    // ================= epilogue: thread = query (TMEM lane); par 0 -> key rows 0..3 of the tile, par 1 -> key rows 4..7
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int P2 = p.P >> 4, P3 = p.P >> 6;          // keys per query at levels 1..3 (W8 = 32, H8 % 8 == 0)
    const uint32_t stg = stg0 + (uint32_t)(warp - 2) * CB_STG;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = t / per_sample, r = t - b * per_sample;
      const int qt = r / p.k_tiles, kt = r - qt * p.k_tiles;
      const long long row = (long long)b * p.P + qt * CB_M + q * 32 + lane;       // this thread's query
      float* o2 = p.lvl2 + row * P2 + kt * 16 + par * 8;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CB_N + par * 128);
      mbar_wait(bar_tfull + 8 * acc, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      float p1[16], p2[8], l2[8];
      const int qrow0 = b * p.P + qt * CB_M + q * 32;                 // first query of this warp
#pragma unroll
      for (int kr = 0; kr < 4; ++kr) {
        float v[32];
        __syncwarp();
        tmem_ld32(t_addr + (uint32_t)(kr * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.scale;
        // level 0: [32 queries][32 keys] block -> shared memory (128 B rows, SWIZZLE_128B pattern: conflict-free 16 B stores) ->
        // one TMA store.  Per-thread global stores of a thread's own row measured 19 GB/s per SM: every warp instruction
        // touches 32 different lines.
        const uint32_t blk = stg;
        if (lane == 0) bulk_wait_group_read0();                          // the previous slab's stores have read the staging blocks
        __syncwarp();
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4))), "f"(v[4 * j]),
                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
        }
        if ((kr & 1) == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) p1[j] = v[2 * j] + v[2 * j + 1];
        } else {
          float l1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) l1[j] = ((p1[j] + v[2 * j]) + v[2 * j + 1]) * 0.25f;
          // level 1: [32 queries][16 keys] block, 64 B rows (SWIZZLE_64B pattern)
          const uint32_t blk1 = stg + 4096u;
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk1 + (uint32_t)(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4))), "f"(l1[4 * j]),
                           "f"(l1[4 * j + 1]), "f"(l1[4 * j + 2]), "f"(l1[4 * j + 3]) : "memory");
          }
          if (kr == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) p2[j] = l1[2 * j] + l1[2 * j + 1];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) l2[j] = ((p2[j] + l1[2 * j]) + l1[2 * j + 1]) * 0.25f;
            st_global_v8(o2, l2, !(p.dbg & 2));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(p.dbg & 2)) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&tmL0), "r"(blk), "r"(kt * CB_N + (par * 4 + kr) * 32), "r"(qrow0) : "memory");
          if (kr & 1)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&tmL1), "r"(stg + 4096u), "r"(kt * 64 + (par * 2 + (kr >> 1)) * 16), "r"(qrow0) : "memory");
          bulk_commit_group();
        }
      }
      // the accumulator has been read: hand it back before the (short) level-3 exchange
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      // level 3: row `kt` of the 4-wide map = 2x2 means over this tile's two level-2 rows (par 0: upper, par 1: lower)
      const uint32_t slot = xchg0 + (uint32_t)(((acc * 4 + q) * 32 + lane) * 32);
      if (par == 1) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "f"(l2[0]), "f"(l2[1]), "f"(l2[2]), "f"(l2[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 16), "f"(l2[4]), "f"(l2[5]), "f"(l2[6]), "f"(l2[7]) : "memory");
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");          // the two warps of this lane quarter
      if (par == 0) {
        float lo[8], l3[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(lo[0]), "=f"(lo[1]), "=f"(lo[2]), "=f"(lo[3]) : "r"(slot));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(lo[4]), "=f"(lo[5]), "=f"(lo[6]), "=f"(lo[7]) : "r"(slot + 16));
#pragma unroll
        for (int j = 0; j < 4; ++j) l3[j] = (((l2[2 * j] + l2[2 * j + 1]) + lo[2 * j]) + lo[2 * j + 1]) * 0.25f;
        *reinterpret_cast<float4*>(p.lvl3 + row * P3 + kt * 4) = make_float4(l3[0], l3[1], l3[2], l3[3]);
      }
      // (the slot of this tile parity is rewritten two tiles later, after another bar.sync of the same pair: no hazard)
    }
    if (lane == 0) bulk_wait_group0();          // shared memory must outlive the last TMA stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// f1s / f2s: pixel-major split-bf16 feature maps (render = queries, real = keys), hi -> lo plane stride `plane` elements
bool corr_pyramid_fused_ok(int C, int H8, int W8, int num_levels) {
  const char* e = getenv("SCFLOW_CORR_FUSED");          // read per call (the parity test compares both forms in one process)
  const bool on = e ? atoi(e) != 0 : true;
  return on && W8 == CB_W8 && H8 % 8 == 0 && C % CB_BK == 0 && num_levels == 4;
}

int corr_pyramid_fused(const void* f1s, const void* f2s, long long plane, int B, int C, int H8, int W8, float* const* levels,
                       cudaStream_t st) {
  SCF_REQUIRE(corr_pyramid_fused_ok(C, H8, W8, 4), SCF_ERR_UNSUPPORTED, "corr_pyramid_fused: needs a 32-wide map, H8 %% 8 == 0, 4 levels");
  CorrFusedParams p = {};
  p.B = B; p.P = H8 * W8; p.C = C; p.H8 = H8;
  p.q_tiles = p.P / CB_M; p.k_tiles = p.P / CB_N; p.num_tiles = B * p.q_tiles * p.k_tiles;
  p.scale = 1.0f / sqrtf((float)C);
  p.lvl0 = levels[0]; p.lvl1 = levels[1]; p.lvl2 = levels[2]; p.lvl3 = levels[3];
  { const char* de = getenv("SCFLOW_CORR_DBG"); p.dbg = de ? atoi(de) : 0; }
  for (int l = 0; l < 4; ++l)
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(levels[l]) % 32 == 0, SCF_ERR_ALIGN, "corr_pyramid_fused: level %d must be 32B aligned", l);
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(f1s) % 16 == 0 && reinterpret_cast<uintptr_t>(f2s) % 16 == 0 && (plane * 2) % 16 == 0, SCF_ERR_ALIGN,
              "corr_pyramid_fused: feature maps must be 16B aligned");
  CUtensorMap tmQ, tmK, tmL0, tmL1;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)B * p.P, 2};
    cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)plane * 2};
    cuuint32_t boxq[3] = {(cuuint32_t)CB_BK, (cuuint32_t)CB_M, 2}, boxk[3] = {(cuuint32_t)CB_BK, (cuuint32_t)CB_N, 2};
    SCF_TRY(encode_map(&tmQ, f1s, 3, dims, str, boxq, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
    SCF_TRY(encode_map(&tmK, f2s, 3, dims, str, boxk, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
    // output maps of the epilogue's TMA stores: level 0 [B*P][P] in [32 queries][32 keys] blocks, level 1 [B*P][P/4] in [32][16]
    cuuint64_t d0[2] = {(cuuint64_t)p.P, (cuuint64_t)B * p.P}, s0[1] = {(cuuint64_t)p.P * 4};
    cuuint32_t b0[2] = {32, 32};
    SCF_TRY(encode_map(&tmL0, levels[0], 2, d0, s0, b0, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B));
    cuuint64_t d1[2] = {(cuuint64_t)p.P / 4, (cuuint64_t)B * p.P}, s1[1] = {(cuuint64_t)p.P};
    cuuint32_t b1[2] = {16, 32};
    SCF_TRY(encode_map(&tmL1, levels[1], 2, d1, s1, b1, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(corr_pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM); });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(corr_pyramid_kernel): %s", cudaGetErrorString(attr_err));
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_tiles < num_sms ? p.num_tiles : num_sms); cfg.blockDim = dim3(64 + 32 * CB_EW);
  cfg.dynamicSmemBytes = CB_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, corr_pyramid_kernel, tmQ, tmK, tmL0, tmL1, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("corr_pyramid_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("corr_pyramid_kernel");
}

}  // namespace scf
