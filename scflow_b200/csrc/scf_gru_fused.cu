// One kernel per SepConvGRU pass (models/decoder/raft_decoder.py:235-253):
//     z = sigmoid(Wz * [h, x]) ; r = sigmoid(Wr * [h, x]) ; q = tanh(Wq * [r.h, x]) ; h' = (1 - z) h + z q
// with x = [context | motion]; the context columns are loop invariant and arrive as fp32 maps `pre_zr` / `pre_q` (bias folded
// in), so the kernel contracts over [h | motion] (K = 256 per tap, 5 taps).
//
// Why it fuses per tile.  Both passes have 1-D taps (1x5, then 5x1).  A 256-pixel tile made of WHOLE rows (pass 0: 8 rows x 32)
// or WHOLE columns (pass 1: 32 rows x 8 columns) needs r.h of that tile only for its q convolution - no halo from a neighbour
// tile, no grid-wide synchronisation - so z, r, r.h, q and the state update of a tile stay on one SM:
//   * transposed tcgen05 form: accumulator lane = output channel (M = 128), accumulator column = pixel (N = 256);
//     z in TMEM columns [0, 256), r in [256, 512); q re-uses z's columns once z has been drained.
//   * pixels are ordered (outer, inner) with inner = the 8 positions ACROSS the taps and outer = the 32 positions ALONG the taps
//     (+ 2 zero positions on either side, written by TMA's out-of-bounds fill), so one 8-row core-matrix group of the UMMA
//     B operand = one outer position and a tap is the same shared-memory tile read 8 rows further on: the [h | motion]
//     activations of a 32-channel chunk are fetched ONCE for all five taps of both gates (36 KB instead of 5 x 32 KB).
//     For the horizontal pass the tensor map simply lists y before x, which makes TMA deliver the tile x-major.
//   * gate epilogue: z -> sigmoid -> fp32 scratch (read back by the same thread at the end); r -> sigmoid -> r.h -> split-bf16,
//     written into shared memory in exactly that layout as the B operand of the q convolution (64 channels at a time, the
//     motion half of q's reduction runs on the tensor core meanwhile); final epilogue: tanh, state update, h' leaves through TMA
//     stores as fp32 + split-bf16 (the next convolution's operand).
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

constexpr int G_OUTER = 32, G_INNER = 8, G_HALO = 2, G_TAPS = 5;
constexpr int G_PIX = G_OUTER * G_INNER;                               // 256 pixels per tile
constexpr uint32_t G_ROWB = 64;                                        // 32 channels x bf16 (SWIZZLE_64B rows)
constexpr uint32_t G_ACT_ROWS = (G_OUTER + 2 * G_HALO) * G_INNER;      // 288 pixel rows incl. the zero halo
constexpr uint32_t G_ACT_PLANE = G_ACT_ROWS * G_ROWB;                  // 18432
constexpr uint32_t G_ACT_STAGE = 2 * G_ACT_PLANE;                      // hi + lo
constexpr uint32_t G_W_PLANE = 128 * G_ROWB;                           // 8192
constexpr uint32_t G_W_STAGE = 2 * G_W_PLANE;
constexpr int G_ACT_STAGES = 2, G_W_STAGES = 5;
constexpr uint32_t G_RH_BYTES = 2 * G_ACT_STAGE;                       // r.h, 64 channels: [sub-chunk 2][plane 2][288 rows][64 B]
constexpr int G_EW = 8;                                                // epilogue warps
constexpr int G_SMEM = 1024 + 1024 + (int)G_RH_BYTES + G_ACT_STAGES * (int)G_ACT_STAGE + G_W_STAGES * (int)G_W_STAGE;
static_assert(G_SMEM <= 232448, "fused GRU pass does not fit in shared memory");

struct GruParams {
  int B, H, W, vertical, num_tiles, tiles_per_img;
  const float* h_f32;      // [B*P][128]
  const float* pre_zr;     // [B*P][256]  context term + bias of z | r
  const float* pre_q;      // [B*P][128]
  float* z;                // [B*P][128]  scratch
  int dbg;                 // timing experiments: 1 = no MMAs
};

__device__ __forceinline__ float g_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float g_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__global__ void __launch_bounds__(64 + 32 * G_EW, 1)
gru_pass_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmM,
                const __grid_constant__ CUtensorMap tmWzr, const __grid_constant__ CUtensorMap tmWq,
                const __grid_constant__ CUtensorMap tmOF, const __grid_constant__ CUtensorMap tmOH, const GruParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // header
  const uint32_t bar_wfull = smem_base, bar_wempty = smem_base + 64, bar_pfull = smem_base + 128, bar_pempty = smem_base + 144,
                 bar_zr_full = smem_base + 160, bar_z_drained = smem_base + 168, bar_rh_ready = smem_base + 176 /* 2 */,
                 bar_rh_free = smem_base + 192, bar_q_full = smem_base + 200, bar_tile_free = smem_base + 208,
                 tmem_slot = smem_base + 224;
  const uint32_t rh0 = smem_base + 1024;
  const uint32_t act0 = rh0 + G_RH_BYTES;
  const uint32_t wring0 = act0 + G_ACT_STAGES * G_ACT_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmH); prefetch_tmap(&tmM); prefetch_tmap(&tmWzr); prefetch_tmap(&tmWq); prefetch_tmap(&tmOF); prefetch_tmap(&tmOH);
    for (int s = 0; s < G_W_STAGES; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
    for (int s = 0; s < G_ACT_STAGES; ++s) { mbar_init(bar_pfull + 8 * s, 1); mbar_init(bar_pempty + 8 * s, 1); }
    mbar_init(bar_zr_full, 1);
    mbar_init(bar_z_drained, G_EW);
    mbar_init(bar_rh_ready, G_EW / 2);
    mbar_init(bar_rh_ready + 8, G_EW / 2);
    mbar_init(bar_rh_free, 1);
    mbar_init(bar_q_full, 1);
    mbar_init(bar_tile_free, G_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  // zero the r.h operand buffer once: its halo rows stay zero for the kernel's lifetime (the epilogue writes interior rows only)
  for (uint32_t o = threadIdx.x * 16u; o < G_RH_BYTES; o += blockDim.x * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rh0 + o), "r"(0u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

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
        const int b = t / p.tiles_per_img, inner0 = (t - b * p.tiles_per_img) * G_INNER;
        // (no wait on the previous tile: the operand rings are guarded by their own barriers, so the next tile's first chunks
        // are prefetched while the previous tile's final epilogue runs)
        // phase 1: z | r over [h | motion]
        for (int c = 0; c < 8; ++c) {
          act_load(c < 4 ? &tmH : &tmM, (c & 3) * 32, inner0, b);
          for (int tap = 0; tap < G_TAPS; ++tap) {
            w_load(&tmWzr, c * 32, 0, tap);
            w_load(&tmWzr, c * 32, 128, tap);
          }
        }
        // phase 2a: q over the motion channels (columns 128.. of Wq)
        for (int c = 0; c < 4; ++c) {
          act_load(&tmM, c * 32, inner0, b);
          for (int tap = 0; tap < G_TAPS; ++tap) w_load(&tmWq, 128 + c * 32, 0, tap);
        }
        // phase 2b: q over r.h (operand produced in shared memory by the gate epilogue)
        for (int c = 0; c < 4; ++c)
          for (int tap = 0; tap < G_TAPS; ++tap) w_load(&tmWq, c * 32, 0, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer: D[channel][pixel] += W[channel][k] * X[pixel][k]   (hi*hi + hi*lo + lo*hi)
      const uint32_t idesc = make_idesc_bf16(128, G_PIX);
      int ws = 0, as = 0;
      uint32_t wph = 0, aph = 0;
      // one (tap, 32-channel chunk) group: 2 k-steps x 3 products, A = weight stage, B = pixel rows starting `tap` outer positions in
      auto mma_group = [&](uint32_t d_tmem, uint32_t x_hi_base, uint32_t x_plane, int tap, bool first) {
        mbar_wait(bar_wfull + 8 * ws, wph);
        tc_fence_after();
        const uint32_t w_addr = wring0 + ws * G_W_STAGE;
        const uint64_t w_hi = make_smem_desc_sw64(w_addr, 512), w_lo = make_smem_desc_sw64(w_addr + G_W_PLANE, 512);
        const uint32_t x_addr = x_hi_base + (uint32_t)(tap * G_INNER) * G_ROWB;
        const uint64_t x_hi = make_smem_desc_sw64(x_addr, 512), x_lo = make_smem_desc_sw64(x_addr + x_plane, 512);
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
        // ---- phase 1
        for (int c = 0; c < 8; ++c) {
          mbar_wait(bar_pfull + 8 * as, aph);
          tc_fence_after();
          const uint32_t x_base = act0 + as * G_ACT_STAGE;
          for (int tap = 0; tap < G_TAPS; ++tap) {
            mma_group(tmem_base, x_base, G_ACT_PLANE, tap, c == 0 && tap == 0);            // z
            mma_group(tmem_base + G_PIX, x_base, G_ACT_PLANE, tap, c == 0 && tap == 0);    // r
          }
          umma_commit(bar_pempty + 8 * as);
          if (++as == G_ACT_STAGES) { as = 0; aph ^= 1u; }
        }
        umma_commit(bar_zr_full);
        // ---- phase 2a: q (motion part) into z's columns, once z has been read out
        mbar_wait(bar_z_drained, par);
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          mbar_wait(bar_pfull + 8 * as, aph);
          tc_fence_after();
          const uint32_t x_base = act0 + as * G_ACT_STAGE;
          for (int tap = 0; tap < G_TAPS; ++tap) mma_group(tmem_base, x_base, G_ACT_PLANE, tap, c == 0 && tap == 0);
          umma_commit(bar_pempty + 8 * as);
          if (++as == G_ACT_STAGES) { as = 0; aph ^= 1u; }
        }
        // ---- phase 2b: q (r.h part), 64 channels at a time
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait(bar_rh_ready + 8 * kh, par);
          tc_fence_after();
          for (int sub = 0; sub < 2; ++sub)
            for (int tap = 0; tap < G_TAPS; ++tap) mma_group(tmem_base, rh0 + sub * G_ACT_STAGE, G_ACT_PLANE, tap, false);
          if (kh == 0) umma_commit(bar_rh_free);        // channels 64..127 of r.h may overwrite the buffer
        }
        umma_commit(bar_q_full);
      }
    }
  } else {
    // ================= epilogue warps: thread = channel (TMEM lane), columns = pixels; the two warps of a lane quarter take
    // alternate 16-column chunks (chunk = 2 outer positions x 8 inner positions)
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int c = q * 32 + lane;
    const int odd = lane & 1;
    // staging for the final TMA stores: 8 KB per warp inside the INTERIOR rows of the r.h buffer (idle by then; halo untouched)
    const uint32_t stg = rh0 + (uint32_t)((warp - 2) >> 1) * G_ACT_PLANE + 1024u + (uint32_t)((warp - 2) & 1) * 8192u;
    uint32_t kc = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const uint32_t tpar = (uint32_t)it & 1u;
      const int b = t / p.tiles_per_img, inner0 = (t - b * p.tiles_per_img) * G_INNER;
      // pixel index of column 16*ch + j : outer = 2*ch + (j >> 3), inner = j & 7
      const long long img = (long long)b * p.H * p.W;
      auto pix_of = [&](int ch, int j) -> long long {
        const int outer = 2 * ch + (j >> 3), inner = inner0 + (j & 7);
        return p.vertical ? img + (long long)outer * p.W + inner : img + (long long)inner * p.W + outer;
      };
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      // ---------------- z: sigmoid -> scratch
      {
        float npre[16];
        auto issue = [&](int ch) {
#pragma unroll
          for (int j = 0; j < 16; ++j) npre[j] = __ldg(p.pre_zr + pix_of(ch, j) * 256 + c);
        };
        issue(par);
        mbar_wait(bar_zr_full, tpar);
        tc_fence_after();
#pragma unroll 1
        for (int ch = par; ch < G_PIX / 16; ch += 2) {
          float v[16], pv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pv[j] = npre[j];
          if (ch + 2 < G_PIX / 16) issue(ch + 2);
          __syncwarp();
          tmem_ld16(t_lane + (uint32_t)(ch * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) p.z[pix_of(ch, j) * 128 + c] = g_sigmoid(v[j] + pv[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_z_drained);
      }
      // ---------------- r: sigmoid, r.h -> split-bf16 B operand in shared memory (this warp's 32 channels = sub-chunk q & 1 of
      // channel half q >> 1)
      {
        const int kh = q >> 1;
        float npre[16], nh[16];
        auto issue = [&](int ch) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const long long px = pix_of(ch, j);
            npre[j] = __ldg(p.pre_zr + px * 256 + 128 + c);
            nh[j] = __ldg(p.h_f32 + px * 128 + c);
          }
        };
        issue(par);
        if (kh == 1) mbar_wait(bar_rh_free, tpar);        // the MMAs over channels 0..63 have finished reading the buffer
        const uint32_t blk = rh0 + (uint32_t)(q & 1) * G_ACT_STAGE;
#pragma unroll 1
        for (int ch = par; ch < G_PIX / 16; ch += 2) {
          float v[16], pv[16], hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { pv[j] = npre[j]; hv[j] = nh[j]; }
          if (ch + 2 < G_PIX / 16) issue(ch + 2);
          __syncwarp();
          tmem_ld16(t_lane + (uint32_t)(G_PIX + ch * 16), v);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float a0 = g_sigmoid(v[j] + pv[j]) * hv[j], a1 = g_sigmoid(v[j + 1] + pv[j + 1]) * hv[j + 1];
            const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
            const uint32_t hl0 = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(l0) << 16);
            const uint32_t hl1 = (uint32_t)__bfloat16_as_ushort(h1) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            // lanes 2i / 2i+1 trade: the even lane stores channels (c, c+1) of pixel j, the odd lane channels (c-1, c) of pixel j+1
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? hl0 : hl1, 1);
            const uint32_t mine = odd ? hl1 : hl0;
            const uint32_t lo_ch = odd ? recv : mine, hi_ch = odd ? mine : recv;
            const int col = ch * 16 + j + odd;                                  // tile pixel (= accumulator column)
            const uint32_t row = (uint32_t)(col + G_HALO * G_INNER);             // row of the halo'd operand tile
            const uint32_t byte = (uint32_t)(lane & ~1) * 2u;                    // byte offset of the channel pair inside the 64 B row
            const uint32_t off = row * G_ROWB + ((((byte >> 4) ^ (row >> 1)) & 3u) << 4) + (byte & 15u);      // SWIZZLE_64B
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(blk + off), "r"((lo_ch & 0xffffu) | (hi_ch << 16)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(blk + G_ACT_PLANE + off), "r"((lo_ch >> 16) | (hi_ch & 0xffff0000u)) : "memory");
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rh_ready + 8 * kh);
      }
      // ---------------- q: tanh, state update, h' out (fp32 + split-bf16) through TMA stores
      {
        float npre[16], nh[16], nz[16];
        auto issue = [&](int ch) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const long long px = pix_of(ch, j);
            npre[j] = __ldg(p.pre_q + px * 128 + c);
            nh[j] = __ldg(p.h_f32 + px * 128 + c);
            nz[j] = p.z[px * 128 + c];                    // written above by this very thread
          }
        };
        issue(par);
        mbar_wait(bar_q_full, tpar);
        tc_fence_after();
#pragma unroll 1
        for (int ch = par; ch < G_PIX / 16; ch += 2) {
          float v[16], pv[16], hv[16], zv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { pv[j] = npre[j]; hv[j] = nh[j]; zv[j] = nz[j]; }
          if (ch + 2 < G_PIX / 16) issue(ch + 2);
          __syncwarp();
          tmem_ld16(t_lane + (uint32_t)(ch * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (1.f - zv[j]) * hv[j] + zv[j] * g_tanh(v[j] + pv[j]);
          // staging: two 4 KB sets used alternately, each = [16 px][32 ch] fp32 (2 KB) + [2 planes][16 px][32 ch] bf16 (2 KB);
          // a set is rewritten only after the bulk group that read it two chunks ago has completed
          const uint32_t blk_f = stg + (kc & 1u) * 4096u, blk_h = blk_f + 2048u;
          ++kc;
          if (lane == 0) bulk_wait_group_read1();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(blk_f + (uint32_t)(j * 128 + lane * 4)), "f"(v[j]) : "memory");
          const uint32_t hbase = blk_h + (uint32_t)(odd * 64 + (lane & ~1) * 2);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[j]), h1 = __float2bfloat16_rn(v[j + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[j + 1] - __bfloat162float(h1));
            const uint32_t hl0 = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(l0) << 16);
            const uint32_t hl1 = (uint32_t)__bfloat16_as_ushort(h1) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? hl0 : hl1, 1);
            const uint32_t mine = odd ? hl1 : hl0;
            const uint32_t lo_ch = odd ? recv : mine, hi_ch = odd ? mine : recv;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(hbase + (uint32_t)(j * 64)), "r"((lo_ch & 0xffffu) | (hi_ch << 16)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(hbase + 1024u + (uint32_t)(j * 64)), "r"((lo_ch >> 16) | (hi_ch & 0xffff0000u)) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmOF, blk_f, q * 32, inner0, 2 * ch, b);
            tma_store_5d(&tmOH, blk_h, q * 32, inner0, 2 * ch, b, 0);
            bulk_commit_group();
          }
        }
        if (lane == 0) bulk_wait_group_read0();           // the staging area is the next tile's r.h operand
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tile_free);
      }
    }
    if (lane == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

int gru_pass_fused(const scf_gru_pass_desc& d, cudaStream_t st) {
  SCF_REQUIRE(d.h_hl && d.h_f32 && d.m_hl && d.w_zr && d.w_q && d.pre_zr && d.pre_q && d.z_scratch && d.out_f32 && d.out_hl,
              SCF_ERR_ARG, "scf_gru_pass_fused: null pointer");
  SCF_REQUIRE(d.B > 0 && d.H > 0 && d.W > 0, SCF_ERR_ARG, "scf_gru_pass_fused: empty shape");
  const int outer = d.vertical ? d.H : d.W, inner = d.vertical ? d.W : d.H;
  SCF_REQUIRE(outer == G_OUTER && inner % G_INNER == 0, SCF_ERR_UNSUPPORTED,
              "scf_gru_pass_fused: the pass needs %d positions along the taps and a multiple of %d across (got %d x %d)", G_OUTER,
              G_INNER, outer, inner);
  auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
  SCF_REQUIRE(al16(d.h_hl) && al16(d.m_hl) && al16(d.w_zr) && al16(d.w_q) && al16(d.out_f32) && al16(d.out_hl) && d.h_plane % 8 == 0 &&
                  d.m_plane % 8 == 0 && d.out_plane % 8 == 0,
              SCF_ERR_ALIGN, "scf_gru_pass_fused: buffers and plane strides must be 16B aligned");
  SCF_REQUIRE(d.out_f32 != d.h_f32 && d.out_hl != d.h_hl, SCF_ERR_ARG, "scf_gru_pass_fused: the state is not updated in place");
  GruParams p = {};
  p.B = d.B; p.H = d.H; p.W = d.W; p.vertical = d.vertical ? 1 : 0;
  p.tiles_per_img = inner / G_INNER;
  p.num_tiles = d.B * p.tiles_per_img;
  p.h_f32 = d.h_f32; p.pre_zr = d.pre_zr; p.pre_q = d.pre_q; p.z = d.z_scratch;
  { const char* de = getenv("SCFLOW_GRU_DBG"); p.dbg = de ? atoi(de) : 0; }
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
  std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(gru_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM); });
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
  cudaError_t le = cudaLaunchKernelEx(&cfg, gru_pass_kernel, tmH, tmM, tmWzr, tmWq, tmOF, tmOH, p);
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
