// One kernel per SepConvGRU pass (models/decoder/raft_decoder.py:235-253):
//     z = sigmoid(Wz * [h, x]) ; r = sigmoid(Wr * [h, x]) ; q = tanh(Wq * [r.h, x]) ; h' = (1 - z) h + z q
// with x = [context | motion]; the context columns are loop invariant and arrive as fp32 maps `pre_zr` / `pre_q` (bias folded
// in), so the kernel contracts over [h | motion] (K = 256 per tap, 5 taps).
//
// Why it fuses per tile.  Both passes have 1-D taps (1x5, then 5x1).  A 256-pixel tile made of WHOLE rows (pass 0: 8 rows x 32)
// or WHOLE columns (pass 1: 32 rows x 8 columns) needs r.h of that tile only for its q convolution - no halo from a neighbour
// tile, no grid-wide synchronisation - so z, r, r.h, q and the state update of a tile stay on one SM:
//   * transposed tcgen05 form: accumulator lane = output channel (M = 128), accumulator column = pixel (N = 256);
//     z in TMEM columns [0, 256), r in [256, 512); q re-uses r's columns once the gate epilogue has read r out; z never leaves
//     TMEM until the final epilogue.
//   * pixels are ordered (outer, inner) with inner = the 8 positions ACROSS the taps and outer = the 32 positions ALONG the taps
//     (+ 2 zero positions on either side, written by TMA's out-of-bounds fill), so one 8-row core-matrix group of the UMMA
//     B operand = one outer position and a tap is the same shared-memory tile read 8 rows further on: the [h | motion]
//     activations of a 32-channel chunk are fetched ONCE for all five taps of both gates (36 KB instead of 5 x 32 KB).
//     For the horizontal pass the tensor map simply lists y before x, which makes TMA deliver the tile x-major.
//   * gate epilogue (all eight epilogue warps at once): r -> sigmoid -> r.h -> split-bf16, written into shared memory in exactly
//     that layout as the B operand of the q convolution - channels 0..63 into a dedicated buffer, 64..127 into the activation ring,
//     which is idle between phase 1 and the motion half of q's reduction; final epilogue: z = sigmoid, q = tanh, state update,
//     h' leaves through TMA stores as fp32 + split-bf16 (the next convolution's operand).
// Against the two-kernel form this removes one kernel boundary per pass (prologue, pipeline fill, exposed epilogue), the HBM/L2
// round trip of r.h and four of every five activation fetches.
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz);

constexpr int G_OUTER = 32, G_INNER = 8, G_HALO = 2, G_TAPS = 5, G_DIM = 32;     // the map is G_DIM x G_DIM (256x256 crops at 1/8)
constexpr int G_PIX = G_OUTER * G_INNER;                               // 256 pixels per tile
constexpr uint32_t G_ROWB = 64;                                        // 32 channels x bf16 (SWIZZLE_64B rows)
constexpr uint32_t G_ACT_ROWS = (G_OUTER + 2 * G_HALO) * G_INNER;      // 288 pixel rows incl. the zero halo
constexpr uint32_t G_ACT_PLANE = G_ACT_ROWS * G_ROWB;                  // 18432
constexpr uint32_t G_ACT_STAGE = 2 * G_ACT_PLANE;                      // hi + lo
constexpr uint32_t G_W_PLANE = 128 * G_ROWB;                           // 8192
constexpr uint32_t G_W_STAGE = 2 * G_W_PLANE;
constexpr int G_ACT_STAGES = 2, G_W_STAGES = 5;
constexpr uint32_t G_RH_BYTES = 2 * G_ACT_STAGE;                       // r.h, 64 channels: [sub-chunk 2][plane 2][288 rows][64 B]
static_assert(G_RH_BYTES == G_ACT_STAGES * G_ACT_STAGE, "the activation ring doubles as the second r.h buffer");
constexpr int G_EW = 8;                                                // epilogue warps
constexpr int G_SMEM = 1024 + 1024 + (int)G_RH_BYTES + G_ACT_STAGES * (int)G_ACT_STAGE + G_W_STAGES * (int)G_W_STAGE;
static_assert(G_SMEM <= 232448, "fused GRU pass does not fit in shared memory");

struct GruParams {
  int B, num_tiles;
  const float* h_f32;      // [B*P][128]
  const float* pre_zr;     // [B*P][256]  context term + bias of z | r
  const float* pre_q;      // [B*P][128]
  int dbg;                 // timing experiments (SCFLOW_GRU_DBG): 1 = no MMAs, 2 = no epilogue global loads
  long long* dbg_times;    // optional [grid][16] globaltimer stamps of each CTA's first tile (tools/bench_gru.py, TRACE=1)
};

__device__ __forceinline__ float g_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float g_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// two 16-column TMEM loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float* a, float* b) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(tb)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); }
}

// hi / lo bf16 halves of x as 16-bit stores (lanes = consecutive channels: a warp fills 64 contiguous bytes, no shuffles)
__device__ __forceinline__ void st_split_u16(uint32_t addr_hi, uint32_t addr_lo, float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr_hi), "h"(__bfloat16_as_ushort(h)) : "memory");
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr_lo), "h"(__bfloat16_as_ushort(l)) : "memory");
}

// VERT = false: 1x5 pass (taps along x: outer = x, inner = y); VERT = true: 5x1 pass (outer = y, inner = x).  The map is
// G_DIM x G_DIM, so every pixel offset inside a 16-column chunk is a compile-time constant.
template <bool VERT>
__global__ void __launch_bounds__(64 + 32 * G_EW, 1)
gru_pass_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmM,
                const __grid_constant__ CUtensorMap tmWzr, const __grid_constant__ CUtensorMap tmWq,
                const __grid_constant__ CUtensorMap tmOF, const __grid_constant__ CUtensorMap tmOH, const GruParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // header
  const uint32_t bar_wfull = smem_base, bar_wempty = smem_base + 64, bar_pfull = smem_base + 128, bar_pempty = smem_base + 144,
                 bar_zr_full = smem_base + 160, bar_rh_ready = smem_base + 168, bar_rh_consumed = smem_base + 176,
                 bar_q_full = smem_base + 184, bar_tile_free = smem_base + 192, tmem_slot = smem_base + 224;
  const uint32_t rh0 = smem_base + 1024;                   // r.h channels 0..63   (dedicated buffer)
  const uint32_t act0 = rh0 + G_RH_BYTES;                  // activation ring; holds r.h channels 64..127 between phase 1 and phase 2a
  const uint32_t wring0 = act0 + G_ACT_STAGES * G_ACT_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  auto stamp = [&](int slot) {
    if (p.dbg_times) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg_times[(long long)blockIdx.x * 16 + slot] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmH); prefetch_tmap(&tmM); prefetch_tmap(&tmWzr); prefetch_tmap(&tmWq); prefetch_tmap(&tmOF); prefetch_tmap(&tmOH);
    for (int s = 0; s < G_W_STAGES; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
    for (int s = 0; s < G_ACT_STAGES; ++s) { mbar_init(bar_pfull + 8 * s, 1); mbar_init(bar_pempty + 8 * s, 1); }
    mbar_init(bar_zr_full, 1);
    mbar_init(bar_rh_ready, G_EW);
    mbar_init(bar_rh_consumed, 1);
    mbar_init(bar_q_full, 1);
    mbar_init(bar_tile_free, G_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  // zero the dedicated r.h buffer once: its halo rows stay zero for the kernel's lifetime (the epilogue writes interior rows only)
  for (uint32_t o = threadIdx.x * 16u; o < G_RH_BYTES; o += blockDim.x * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rh0 + o), "r"(0u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  if (threadIdx.x == 0) stamp(1);
  constexpr int TPI = G_DIM / G_INNER;                     // tiles per sample

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int ws = 0, as = 0;
      uint32_t wph = 0, aph = 0;
      auto act_load = [&](const CUtensorMap* tm, int k0, int inner0, int b) {
        mbar_wait(bar_pempty + 8 * as, aph ^ 1u);
        mbar_arrive_expect_tx(bar_pfull + 8 * as, G_ACT_STAGE);
        tma_load_5d(act0 + as * G_ACT_STAGE, tm, bar_pfull + 8 * as, k0, inner0, -G_HALO, b, 0);
        if (++as == G_ACT_STAGES) { as = 0; aph ^= 1u; }
      };
      auto w_load = [&](const CUtensorMap* tm, int k0, int row0, int tap) {
        mbar_wait(bar_wempty + 8 * ws, wph ^ 1u);
        const uint32_t full = bar_wfull + 8 * ws, dst = wring0 + ws * G_W_STAGE;
        mbar_arrive_expect_tx(full, G_W_STAGE);
        tma_load_4d(dst, tm, full, k0, row0, tap, 0);
        tma_load_4d(dst + G_W_PLANE, tm, full, k0, row0, tap, 1);
        if (++ws == G_W_STAGES) { ws = 0; wph ^= 1u; }
      };
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int b = t / TPI, inner0 = (t - b * TPI) * G_INNER;
        // phase 1: z | r over [h | motion]  (the rings are guarded by their own barriers, so the next tile's first chunks are
        // prefetched while the previous tile's final epilogue runs)
        for (int c = 0; c < 8; ++c) {
          act_load(c < 4 ? &tmH : &tmM, (c & 3) * 32, inner0, b);
          for (int tap = 0; tap < G_TAPS; ++tap) {
            w_load(&tmWzr, c * 32, 0, tap);
            w_load(&tmWzr, c * 32, 128, tap);
          }
        }
        // phase 2b: q over r.h (operand written into shared memory by the gate epilogue): weights only
        for (int c = 0; c < 4; ++c)
          for (int tap = 0; tap < G_TAPS; ++tap) w_load(&tmWq, c * 32, 0, tap);
        // phase 2a: q over the motion channels (columns 128.. of Wq); the activation ring is r.h's second buffer until 2b is done
        mbar_wait(bar_rh_consumed, (uint32_t)it & 1u);
        for (int c = 0; c < 4; ++c) {
          act_load(&tmM, c * 32, inner0, b);
          for (int tap = 0; tap < G_TAPS; ++tap) w_load(&tmWq, 128 + c * 32, 0, tap);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer: D[channel][pixel] += W[channel][k] * X[pixel][k]   (hi*hi + hi*lo + lo*hi)
      const uint32_t idesc = make_idesc_bf16(128, G_PIX);
      int ws = 0, as = 0;
      uint32_t wph = 0, aph = 0;
      // one (tap, 32-channel chunk) group: 2 k-steps x 3 products, A = weight stage, B = pixel rows starting `tap` outer positions in
      auto mma_group = [&](uint32_t d_tmem, uint32_t x_hi_base, int tap, bool first) {
        mbar_wait(bar_wfull + 8 * ws, wph);
        tc_fence_after();
        const uint32_t w_addr = wring0 + ws * G_W_STAGE;
        const uint64_t w_hi = make_smem_desc_sw64(w_addr, 512), w_lo = make_smem_desc_sw64(w_addr + G_W_PLANE, 512);
        const uint32_t x_addr = x_hi_base + (uint32_t)(tap * G_INNER) * G_ROWB;
        const uint64_t x_hi = make_smem_desc_sw64(x_addr, 512), x_lo = make_smem_desc_sw64(x_addr + G_ACT_PLANE, 512);
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            umma_bf16(d_tmem, w_hi + ko, x_hi + ko, idesc, (!first || k > 0) ? 1u : 0u);
            umma_bf16(d_tmem, w_hi + ko, x_lo + ko, idesc, 1u);
            umma_bf16(d_tmem, w_lo + ko, x_hi + ko, idesc, 1u);
          }
        }
        umma_commit(bar_wempty + 8 * ws);
        if (++ws == G_W_STAGES) { ws = 0; wph ^= 1u; }
      };
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)it & 1u;
        if (it > 0) { mbar_wait(bar_tile_free, (uint32_t)(it - 1) & 1u); tc_fence_after(); }
        // ---- phase 1: z -> columns [0, 256), r -> [256, 512)
        for (int c = 0; c < 8; ++c) {
          mbar_wait(bar_pfull + 8 * as, aph);
          tc_fence_after();
          const uint32_t x_base = act0 + as * G_ACT_STAGE;
          for (int tap = 0; tap < G_TAPS; ++tap) {
            mma_group(tmem_base, x_base, tap, c == 0 && tap == 0);
            mma_group(tmem_base + G_PIX, x_base, tap, c == 0 && tap == 0);
          }
          umma_commit(bar_pempty + 8 * as);
          if (++as == G_ACT_STAGES) { as = 0; aph ^= 1u; }
        }
        umma_commit(bar_zr_full);
        if (it == 0) stamp(2);
        // ---- phase 2b: q (r.h part) into r's columns, once the gate epilogue has read r out and written r.h
        mbar_wait(bar_rh_ready, par);
        tc_fence_after();
        if (it == 0) stamp(4);
        for (int c = 0; c < 4; ++c)
          for (int tap = 0; tap < G_TAPS; ++tap)
            mma_group(tmem_base + G_PIX, (c < 2 ? rh0 : act0) + (uint32_t)(c & 1) * G_ACT_STAGE, tap, c == 0 && tap == 0);
        umma_commit(bar_rh_consumed);         // the activation ring is free again
        if (it == 0) stamp(5);
        // ---- phase 2a: q (motion part)
        for (int c = 0; c < 4; ++c) {
          mbar_wait(bar_pfull + 8 * as, aph);
          tc_fence_after();
          const uint32_t x_base = act0 + as * G_ACT_STAGE;
          for (int tap = 0; tap < G_TAPS; ++tap) mma_group(tmem_base + G_PIX, x_base, tap, false);
          umma_commit(bar_pempty + 8 * as);
          if (++as == G_ACT_STAGES) { as = 0; aph ^= 1u; }
        }
        umma_commit(bar_q_full);
        if (it == 0) stamp(6);
      }
    }
  } else {
    // ================= epilogue warps: thread = channel (TMEM lane), columns = pixels; the two warps of a lane quarter take
    // alternate 16-column chunks (chunk = 2 outer positions x 8 inner positions)
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int c = q * 32 + lane;
    // element offset of column j of a chunk relative to the chunk's first pixel, in pixels (compile-time per j)
    auto pix_off = [](int j) { return VERT ? (j >> 3) * G_DIM + (j & 7) : (j & 7) * G_DIM + (j >> 3); };
    // r.h operand block of this warp's 32 channels: channels 0..63 -> dedicated buffer, 64..127 -> the (idle) activation ring
    const uint32_t blk = ((q >> 1) ? act0 : rh0) + (uint32_t)(q & 1) * G_ACT_STAGE;
    // SWIZZLE_64B: 16-byte unit (lane >> 3) of the 64 B row is XORed with bits 1..2 of the row index; rows of a chunk start at a
    // multiple of 16, so the XOR term depends on j only: four per-thread constants
    uint32_t swz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) swz[k] = ((((uint32_t)lane >> 3) ^ (uint32_t)k) << 4) + ((uint32_t)lane & 7u) * 2u;
    // staging for the final TMA stores: 8 KB per warp inside the INTERIOR rows of the dedicated r.h buffer (idle by then)
    const uint32_t stg = rh0 + (uint32_t)((warp - 2) >> 1) * G_ACT_PLANE + 1024u + (uint32_t)((warp - 2) & 1) * 8192u;
    uint32_t kc = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const uint32_t tpar = (uint32_t)it & 1u;
      const int b = t / TPI, inner0 = (t - b * TPI) * G_INNER;
      // first pixel of chunk ch: outer = 2*ch, inner = inner0
      const int pix0 = b * (G_DIM * G_DIM) + (VERT ? inner0 : inner0 * G_DIM);
      constexpr int CH_STEP = VERT ? 2 * G_DIM : 2;             // pixels between consecutive chunks
      const float* pre_z = p.pre_zr + (size_t)pix0 * 256 + c;
      const float* pre_r = pre_z + 128;
      const float* pre_q = p.pre_q + (size_t)pix0 * 128 + c;
      const float* hp = p.h_f32 + (size_t)pix0 * 128 + c;
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      // (measured: pulling the tile's epilogue inputs into L2 from these idle warps during phase 1 costs the main loop 4 us of
      // TMA bandwidth and saves the epilogues 2 us - not done)
      // ---------------- gate: r = sigmoid(.), r.h -> split-bf16 B operand of the q convolution, in shared memory
      {
        // inputs are requested TWO chunks ahead (two register sets): one chunk of work (~0.7 us) does not cover an L2 / DRAM round
        // trip while the TMA traffic of the main loop is in flight
        float npre[2][16], nh[2][16];
        auto issue = [&](int ch, float (&pr)[16], float (&hr)[16]) {
          const float* a = pre_r + (size_t)(ch * CH_STEP) * 256;
          const float* hh = hp + (size_t)(ch * CH_STEP) * 128;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pr[j] = (p.dbg & 2) ? 0.f : __ldg(a + pix_off(j) * 256);
            hr[j] = (p.dbg & 2) ? 0.5f : __ldg(hh + pix_off(j) * 128);
          }
        };
        issue(par, npre[0], nh[0]);
        issue(par + 2, npre[1], nh[1]);
        mbar_wait(bar_zr_full, tpar);
        tc_fence_after();
        if (it == 0 && warp == 2 && lane == 0) stamp(3);
        if (q >> 1) {
          // the ring held TMA tiles: re-zero the halo rows (2 outer positions = 1 KB at either end of each of the 4 planes);
          // 4 warps x 32 lanes x 4 stores of 16 B = 8 KB
          const int e = ((warp - 2) >> 2) * 2 + (q & 1);                  // 0..3 among the four warps of this half
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t unit = (uint32_t)(e * 128 + i * 32 + lane);    // 0..511 16-byte units
            const uint32_t plane = unit >> 7, r16 = unit & 127u;           // 128 units per plane: 64 top, 64 bottom
            const uint32_t off = plane * G_ACT_PLANE + (r16 < 64 ? r16 * 16u : G_ACT_PLANE - 1024u + (r16 - 64u) * 16u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(act0 + off), "r"(0u) : "memory");
          }
        }
        auto gate_chunk = [&](int ch, float (&pr)[16], float (&hr)[16]) {
          float v[16], pv[16], hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { pv[j] = pr[j]; hv[j] = hr[j]; }
          if (ch + 4 < G_PIX / 16) issue(ch + 4, pr, hr);
          __syncwarp();
          tmem_ld16(t_lane + (uint32_t)(G_PIX + ch * 16), v);
          const uint32_t rowbase = blk + (uint32_t)(G_HALO * G_INNER + ch * 16) * G_ROWB;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = g_sigmoid(v[j] + pv[j]) * hv[j];
            const uint32_t addr = rowbase + (uint32_t)j * G_ROWB + swz[(j >> 1) & 3];
            st_split_u16(addr, addr + G_ACT_PLANE, a);
          }
        };
#pragma unroll 1
        for (int ch = par; ch < G_PIX / 16; ch += 4) {
          gate_chunk(ch, npre[0], nh[0]);
          gate_chunk(ch + 2, npre[1], nh[1]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rh_ready);
      }
      // ---------------- final: z = sigmoid(.), q = tanh(.), h' = (1 - z) h + z q, out as fp32 + split-bf16 through TMA stores
      {
        float npz[2][16], npq[2][16], nh[2][16];
        auto issue = [&](int ch, float (&zr)[16], float (&qr)[16], float (&hr)[16]) {
          const float* a = pre_z + (size_t)(ch * CH_STEP) * 256;
          const float* bq = pre_q + (size_t)(ch * CH_STEP) * 128;
          const float* hh = hp + (size_t)(ch * CH_STEP) * 128;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            zr[j] = (p.dbg & 2) ? 0.f : __ldg(a + pix_off(j) * 256);
            qr[j] = (p.dbg & 2) ? 0.f : __ldg(bq + pix_off(j) * 128);
            hr[j] = (p.dbg & 2) ? 0.5f : __ldg(hh + pix_off(j) * 128);
          }
        };
        issue(par, npz[0], npq[0], nh[0]);
        issue(par + 2, npz[1], npq[1], nh[1]);
        mbar_wait(bar_q_full, tpar);
        tc_fence_after();
        if (it == 0 && warp == 2 && lane == 0) stamp(10);
        auto final_chunk = [&](int ch, float (&zr)[16], float (&qr)[16], float (&hr)[16]) {
          float vz[16], vq[16], o[16];
          __syncwarp();
          tmem_ld16x2(t_lane + (uint32_t)(ch * 16), t_lane + (uint32_t)(G_PIX + ch * 16), vz, vq);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = g_sigmoid(vz[j] + zr[j]);
            o[j] = fmaf(z, g_tanh(vq[j] + qr[j]) - hr[j], hr[j]);          // (1 - z) h + z q
          }
          if (ch + 4 < G_PIX / 16) issue(ch + 4, zr, qr, hr);
          // staging: two 4 KB sets used alternately, each = [16 px][32 ch] fp32 (2 KB) + [2 planes][16 px][32 ch] bf16 (2 KB);
          // a set is rewritten only after the bulk group that read it two chunks ago has completed
          const uint32_t blk_f = stg + (kc & 1u) * 4096u, blk_h = blk_f + 2048u + (uint32_t)lane * 2u;
          ++kc;
          if (lane == 0) bulk_wait_group_read1();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(blk_f + (uint32_t)(j * 128 + lane * 4)), "f"(o[j]) : "memory");
            st_split_u16(blk_h + (uint32_t)(j * 64), blk_h + 1024u + (uint32_t)(j * 64), o[j]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmOF, blk_f, q * 32, inner0, 2 * ch, b);
            tma_store_5d(&tmOH, blk_f + 2048u, q * 32, inner0, 2 * ch, b, 0);
            bulk_commit_group();
          }
        };
#pragma unroll 1
        for (int ch = par; ch < G_PIX / 16; ch += 4) {
          final_chunk(ch, npz[0], npq[0], nh[0]);
          final_chunk(ch + 2, npz[1], npq[1], nh[1]);
        }
        if (lane == 0) bulk_wait_group_read0();           // the staging area is the next tile's r.h operand
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tile_free);
        if (it == 0 && warp == 2 && lane == 0) stamp(11);
      }
    }
    if (lane == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

int gru_pass_fused(const scf_gru_pass_desc& d, cudaStream_t st) {
  SCF_REQUIRE(d.h_hl && d.h_f32 && d.m_hl && d.w_zr && d.w_q && d.pre_zr && d.pre_q && d.out_f32 && d.out_hl, SCF_ERR_ARG,
              "scf_gru_pass_fused: null pointer");
  SCF_REQUIRE(d.B > 0 && d.H > 0 && d.W > 0, SCF_ERR_ARG, "scf_gru_pass_fused: empty shape");
  SCF_REQUIRE(d.H == G_DIM && d.W == G_DIM, SCF_ERR_UNSUPPORTED,
              "scf_gru_pass_fused: whole-row / whole-column tiles need a %d x %d map (256x256 crops at 1/8); got %d x %d", G_DIM,
              G_DIM, d.H, d.W);
  auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
  SCF_REQUIRE(al16(d.h_hl) && al16(d.m_hl) && al16(d.w_zr) && al16(d.w_q) && al16(d.out_f32) && al16(d.out_hl) && d.h_plane % 8 == 0 &&
                  d.m_plane % 8 == 0 && d.out_plane % 8 == 0,
              SCF_ERR_ALIGN, "scf_gru_pass_fused: buffers and plane strides must be 16B aligned");
  SCF_REQUIRE(d.out_f32 != d.h_f32 && d.out_hl != d.h_hl, SCF_ERR_ARG, "scf_gru_pass_fused: the state is not updated in place");
  GruParams p = {};
  p.B = d.B;
  p.num_tiles = d.B * (G_DIM / G_INNER);
  p.h_f32 = d.h_f32; p.pre_zr = d.pre_zr; p.pre_q = d.pre_q;
  { const char* de = getenv("SCFLOW_GRU_DBG"); p.dbg = de ? atoi(de) : 0; }
  { const char* dt = getenv("SCFLOW_GRU_DBG_TIMES"); p.dbg_times = dt ? reinterpret_cast<long long*>(strtoull(dt, nullptr, 16)) : nullptr; }
  // activation maps, dimension order (channel, inner, outer, sample, plane): the horizontal pass lists y before x
  const long long W = d.W, H = d.H;
  auto act_map = [&](CUtensorMap* m, const void* base, long long plane, int box_c, CUtensorMapDataType dt, int esz, CUtensorMapSwizzle swz,
                     int box_outer, bool planes) -> int {
    const cuuint64_t inner_dim = (cuuint64_t)(d.vertical ? W : H), outer_dim = (cuuint64_t)(d.vertical ? H : W);
    const cuuint64_t inner_str = (cuuint64_t)(d.vertical ? 128 : W * 128) * esz, outer_str = (cuuint64_t)(d.vertical ? W * 128 : 128) * esz;
    cuuint64_t dims[5] = {128, inner_dim, outer_dim, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {inner_str, outer_str, (cuuint64_t)(H * W * 128) * esz, (cuuint64_t)plane * esz};
    cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)G_INNER, (cuuint32_t)box_outer, 1, 2};
    return encode_map(m, base, planes ? 5 : 4, dims, str, box, nullptr, dt, swz);
  };
  CUtensorMap tmH, tmM, tmWzr, tmWq, tmOF, tmOH;
  SCF_TRY(act_map(&tmH, d.h_hl, d.h_plane, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B, G_OUTER + 2 * G_HALO, true));
  SCF_TRY(act_map(&tmM, d.m_hl, d.m_plane, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B, G_OUTER + 2 * G_HALO, true));
  SCF_TRY(act_map(&tmOF, d.out_f32, 0, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_NONE, 2, false));
  SCF_TRY(act_map(&tmOH, d.out_hl, d.out_plane, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_NONE, 2, true));
  auto w_map = [&](CUtensorMap* m, const void* base, int rows) -> int {
    cuuint64_t dims[4] = {256, (cuuint64_t)rows, G_TAPS, 2};
    cuuint64_t str[3] = {256 * 2, (cuuint64_t)rows * 256 * 2, (cuuint64_t)G_TAPS * rows * 256 * 2};
    cuuint32_t box[4] = {32, 128, 1, 1};
    return encode_map(m, base, 4, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  SCF_TRY(w_map(&tmWzr, d.w_zr, 256));
  SCF_TRY(w_map(&tmWq, d.w_q, 128));
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(gru_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(gru_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(gru_pass_kernel): %s", cudaGetErrorString(attr_err));
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_tiles < num_sms ? p.num_tiles : num_sms); cfg.blockDim = dim3(64 + 32 * G_EW);
  cfg.dynamicSmemBytes = G_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = d.vertical ? cudaLaunchKernelEx(&cfg, gru_pass_kernel<true>, tmH, tmM, tmWzr, tmWq, tmOF, tmOH, p)
                              : cudaLaunchKernelEx(&cfg, gru_pass_kernel<false>, tmH, tmM, tmWzr, tmWq, tmOF, tmOH, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("gru_pass_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("gru_pass_kernel");
}

}  // namespace scf

extern "C" {

int scf_gru_pass_fused(const scf_gru_pass_desc* d, void* stream) {
  SCF_REQUIRE(d != nullptr, SCF_ERR_ARG, "scf_gru_pass_fused: null descriptor");
  return scf::gru_pass_fused(*d, (cudaStream_t)stream);
}

}  // extern "C"
