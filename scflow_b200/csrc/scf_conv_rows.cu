// 3x3 stride-1 convolutions of 64 -> 64 channels on large maps: the four residual-block convolutions of the RAFT encoders'
// first stage (models/backbone/resnet.py:14-94 BasicBlock.conv1 / conv2 at 1/2 resolution, 64 channels), the biggest bucket
// of the encoder time.  On the generic pixels-as-rows tile such a layer issues 9 taps x 4 k-steps x 3 split-bf16 products
// = 108 MMAs of N = 64 per 128-pixel tile, and an M = 128 tcgen05.mma costs the same ~61 ns for every N <= 128: half of the
// tensor core's width is idle.  This kernel re-associates the sum instead ("rolling rows"):
//     out[y][x] = sum_ky  Q_ky[y + ky - 1][x],     Q_ky[r][x] = sum_kx W(ky, kx) . in[r][x + kx - 1]
//   * a tile is ONE input row r of 128 pixels (+1 halo pixel each side; TMA zero fill = the convolution's padding), fetched
//     once as two 130 x 128 B SWIZZLE_128B planes (hi, lo); the three x-shifts are row-shifted A descriptors on that tile;
//   * the B operand stacks the three ky taps of one kx along N: [W(2,kx) | W(1,kx) | W(0,kx)] = 192 rows, so ONE MMA of
//     N = 192 yields Q_2[r], Q_1[r], Q_0[r] = the contributions of input row r to out[r-1], out[r], out[r+1]:
//     3 kx x 4 k-steps x 3 products = 36 MMAs per row instead of 108;
//   * TMEM is a ring of eight 64-column slots, one per OUTPUT row in flight; consecutive output rows sit in consecutive
//     slots, so the three row sums accumulate in place in the tensor core: input row r adds into slots of out[r-1], out[r]
//     (accumulate) and opens out[r+1] (first MMA of the step issued separately with accumulate = 0 for that slot).  A step
//     whose three slots wrap around the ring issues two MMAs (N = 128 + 64) per product;
//   * all 9 taps of the weights (2 planes x 72 KB) stay resident in shared memory for the CTA's lifetime; a CTA owns a
//     contiguous range of (image, column strip, row) units, re-reading one halo row at each end of a range;
//   * epilogue: eight warps, one thread per (pixel, 32-channel half) reads its channels from the finished slot, adds bias (+ residual), optional ReLU,
//     writes fp32 and / or split-bf16 NHWC with 32 B vector stores; (InstanceNorm) the sums of the output and of its square stay in
//     registers per pixel column and leave once per CTA segment through a 31-shuffle transpose-reduce - a few rows of partial
//     sums per image, deterministic, no atomics.
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz);
extern thread_local int g_last_m_tiles, g_last_tiles_per_img, g_last_stat_rows_per_img;

constexpr int CR_M = 128, CR_C = 64, CR_EW = 8, CR_ASTAGES = 2, CR_SLOTS = 8;
constexpr uint32_t CR_AROWS = CR_M + 2;                                   // pixels per tile row incl. the x halo
constexpr uint32_t CR_APLANE = (CR_AROWS * 128u + 1023u) & ~1023u;        // 17 KB
constexpr uint32_t CR_ASTAGE = 2 * CR_APLANE;
constexpr uint32_t CR_WBLK = CR_C * 128u;                                 // one tap, one plane: 64 output rows x 128 B
constexpr uint32_t CR_WKX = 3 * CR_WBLK;                                  // the three ky taps of one kx, stacked along N
constexpr uint32_t CR_WPLANE = 3 * CR_WKX;
constexpr uint32_t CR_WBYTES = 2 * CR_WPLANE;                             // 144 KB
constexpr int CR_SMEM = 1024 + 1024 + (int)CR_WBYTES + CR_ASTAGES * (int)CR_ASTAGE;
static_assert(CR_SMEM <= 232448, "conv_rows_kernel does not fit in shared memory");

struct RowsParams {
  int N, H, W, TX;                 // images, OUTPUT map size, 128-pixel column strips per row
  int Hin;                         // stem kernel only: input rows (H = Hin / 2 output rows)
  long long U;                     // work units = N * TX * H output rows of one strip
  const float* bias; int relu;
  float* out_f32; int of_stride, of_coff;
  __nv_bfloat16* out_hl; long long oh_plane; int oh_stride, oh_coff;
  const float* res; int res_stride;
  const __nv_bfloat16* res_hl; long long res_hl_plane; int res_hl_stride;    // residual as split-bf16 planes (instead of res)
  float* stats; int stat_rows;     // rows of [2][64] partial sums per image: (xt * stat_k + segment ordinal) * 4 + warp; zeroed by the launcher
  int stat_k;                      // upper bound of the CTA segments that can touch one (image, strip)
  int al32;                        // 32 B vector accesses allowed: bit 0 fp32 output, bit 1 split output, bit 2 fp32 residual, bit 3 split residual
  int dbg;                         // timing experiments (SCFLOW_ROWS_DBG): 1 no MMAs, 2 no global stores, 4 no activation loads, 8 no epilogue work
  long long* dbg_times;            // optional [grid][8] clock64 totals of the MMA thread's phases (SCFLOW_ROWS_DBG_TIMES = hex pointer)
};

// The three roles (TMA producer, MMA issuer, epilogue) walk the same schedule: the CTA's unit range, cut into segments at
// strip boundaries; a segment with output rows [y0, y1) consumes input rows max(y0-1, 0) .. min(y1, H-1).  `seq` numbers the
// CTA's output rows consecutively (TMEM slot = seq & 7).
template <class F>
__device__ __forceinline__ void rows_for_each_step(const RowsParams& p, F&& f) {
  long long u = p.U * blockIdx.x / gridDim.x;
  const long long u1 = p.U * (blockIdx.x + 1) / gridDim.x;
  uint32_t seq = 0;
  while (u < u1) {
    const int strip = (int)(u / p.H), y0 = (int)(u - (long long)strip * p.H);
    const int y1 = (u1 - u) < (long long)(p.H - y0) ? y0 + (int)(u1 - u) : p.H;
    const int img = strip / p.TX, xt = strip - img * p.TX;
    const int r0 = y0 > 0 ? y0 - 1 : 0, r1 = y1 < p.H ? y1 : p.H - 1;
    for (int r = r0; r <= r1; ++r) f(img, xt, r, y0, y1, seq);
    seq += (uint32_t)(y1 - y0);
    u += y1 - y0;
  }
}

// The same schedule as rows_for_each_step / stem_for_each_step as a resumable iterator (STEM selects the stem's input-row range).  The
// MMA-issuing thread uses it to prepare step i + 1 (barrier waits, descriptors) BETWEEN the MMAs and the commits of step i: the
// tensor pipe's queue is shallow, so scalar work done after a commit would run against an idle pipe.
template <bool STEM>
struct RowsStepIter {
  long long u, u1;
  uint32_t seq;
  int H, Hin, TX, img, xt, y0, y1, r, r1;
  __device__ __forceinline__ RowsStepIter(const RowsParams& p)
      : u(p.U * blockIdx.x / gridDim.x), u1(p.U * (blockIdx.x + 1) / gridDim.x), seq(0), H(p.H), Hin(p.Hin), TX(p.TX), img(0), xt(0), y0(0),
        y1(0), r(1), r1(0) {}
  // advances to the next step; false when the CTA's range is exhausted.  Afterwards img, xt, r, y0, y1, seq describe the step
  __device__ __forceinline__ bool next() {
    if (r < r1) { ++r; return true; }
    if (r1 >= r && y1 > y0) { seq += (uint32_t)(y1 - y0); u += y1 - y0; }       // leaving a segment
    if (u >= u1) return false;
    const int strip = (int)(u / H);
    y0 = (int)(u - (long long)strip * H);
    y1 = (u1 - u) < (long long)(H - y0) ? y0 + (int)(u1 - u) : H;
    img = strip / TX; xt = strip - img * TX;
    if (STEM) { r = 2 * y0 - 3 > 0 ? 2 * y0 - 3 : 0; r1 = 2 * (y1 - 1) + 3 < Hin - 1 ? 2 * (y1 - 1) + 3 : Hin - 1; }
    else { r = y0 > 0 ? y0 - 1 : 0; r1 = y1 < H ? y1 : H - 1; }
    return true;
  }
};

__device__ __forceinline__ float transpose_reduce32_rows(float (&acc)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = up ? acc[j] : acc[j + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      acc[j] = (up ? acc[j + off] : acc[j]) + recv;
    }
  }
  return acc[0];
}

// zero 32 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + 16u), "r"(z) : "memory");
}

// 32 channels (128 B) of the fp32 residual map for one pixel
__device__ __forceinline__ void rows_load_residual(const RowsParams& p, long long pix, bool valid, int h, float4 (&rs)[8]) {
  if (p.res_hl) {
    // hi + lo bf16 planes: 64 B each for this thread's 32 channels (two 32 B loads per plane when the planes are 32 B aligned)
    uint4 hq[4], lq[4];
    const __nv_bfloat16* hp = p.res_hl + pix * p.res_hl_stride + h * 32;
    const __nv_bfloat16* lp = hp + p.res_hl_plane;
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { hq[j] = make_uint4(0u, 0u, 0u, 0u); lq[j] = hq[j]; }
    } else if (p.al32 & 8) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        asm volatile("ld.global.cs.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(hq[2 * j].x), "=r"(hq[2 * j].y), "=r"(hq[2 * j].z), "=r"(hq[2 * j].w), "=r"(hq[2 * j + 1].x), "=r"(hq[2 * j + 1].y),
                       "=r"(hq[2 * j + 1].z), "=r"(hq[2 * j + 1].w)
                     : "l"(hp + 16 * j));
        asm volatile("ld.global.cs.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(lq[2 * j].x), "=r"(lq[2 * j].y), "=r"(lq[2 * j].z), "=r"(lq[2 * j].w), "=r"(lq[2 * j + 1].x), "=r"(lq[2 * j + 1].y),
                       "=r"(lq[2 * j + 1].z), "=r"(lq[2 * j + 1].w)
                     : "l"(lp + 16 * j));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { hq[j] = __ldcs(reinterpret_cast<const uint4*>(hp) + j); lq[j] = __ldcs(reinterpret_cast<const uint4*>(lp) + j); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t hw[4] = {hq[j].x, hq[j].y, hq[j].z, hq[j].w}, lw[4] = {lq[j].x, lq[j].y, lq[j].z, lq[j].w};
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
        f[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
      }
      rs[2 * j] = make_float4(f[0], f[1], f[2], f[3]);
      rs[2 * j + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    return;
  }
  const float* rp = p.res + pix * p.res_stride + h * 32;
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) rs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (p.al32 & 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("ld.global.cs.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(rs[2 * j].x), "=f"(rs[2 * j].y), "=f"(rs[2 * j].z), "=f"(rs[2 * j].w), "=f"(rs[2 * j + 1].x), "=f"(rs[2 * j + 1].y),
                     "=f"(rs[2 * j + 1].z), "=f"(rs[2 * j + 1].w)
                   : "l"(rp + 8 * j));
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) rs[j] = __ldcs(reinterpret_cast<const float4*>(rp) + j);
  }
}

// Epilogue of one finished output row: thread = (pixel of the 128-pixel strip, 32-channel half h); `sq` = the row's sequence number
// in this CTA (TMEM slot sq & 7).  bias (+ residual) (+ ReLU) -> fp32 and / or split-bf16 NHWC, optional InstanceNorm partial sums.
template <bool RES, bool STATS, bool ALIAS = false>
__device__ __forceinline__ void rows_epilogue_row(const RowsParams& p, uint32_t tmem_base, uint32_t bar_tfull, uint32_t bar_tempty,
                                            uint32_t bias_s, uint32_t sq, int img, int xt, int y, int q, int h, int lane,
                                            float (&acc1)[STATS ? 32 : 1], float (&acc2)[STATS ? 32 : 1]) {
  // ALIAS (3x3 kernel): ring of six logical slots in eight physical ones - rows whose window position ran past slot 5 keep part of
  // their sum in the alias slots 6 / 7 (see conv_rows_kernel); every slot is handed back ZEROED, so all MMAs accumulate
  const uint32_t gen = ALIAS ? sq / 6u : sq >> 3, slot = ALIAS ? sq - gen * 6u : sq & 7u;
  const int x = xt * CR_M + q * 32 + lane;
  const bool valid = x < p.W;
  const long long pix = ((long long)img * p.H + y) * p.W + x;
  // residual values are requested before the accumulator is awaited (a one-row look-ahead was measured: no gain - the residual
  // variant is bound by its 2x memory traffic, not by load latency)
  float4 rs[RES ? 8 : 1];
  if constexpr (RES) rows_load_residual(p, pix, valid, h, rs);
  mbar_wait(bar_tfull + 8 * slot, gen & 1u);
  tc_fence_after();
  const bool aliased = ALIAS && slot < 2u;
  if (p.dbg & 8) {                 // timing experiment: hand the slot back without reading it
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar_tempty + 8 * slot); if (aliased) mbar_arrive(bar_tempty + 8 * (slot + 6u)); }
    return;
  }
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 64u;
  {
    float v[32];
    tmem_ld32(taddr + (uint32_t)(h * 32), v);
    if constexpr (ALIAS) {
      if (aliased) {
        float v2[32];
        tmem_ld32(taddr + (uint32_t)(6 * 64 + h * 32), v2);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
        tmem_zero32(taddr + (uint32_t)(6 * 64 + h * 32));
      }
      tmem_zero32(taddr + (uint32_t)(h * 32));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar_tempty + 8 * slot); if (aliased) mbar_arrive(bar_tempty + 8 * (slot + 6u)); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(bias_s + (uint32_t)(h * 128 + j * 16)));
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      if (RES) { v[4 * j] += rs[j].x; v[4 * j + 1] += rs[j].y; v[4 * j + 2] += rs[j].z; v[4 * j + 3] += rs[j].w; }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (valid && !(p.dbg & 2)) {
      if (p.out_f32) {
        float* o = p.out_f32 + pix * p.of_stride + p.of_coff + h * 32;
        if (p.al32 & 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 8 * j), "f"(v[8 * j]), "f"(v[8 * j + 1]),
                         "f"(v[8 * j + 2]), "f"(v[8 * j + 3]), "f"(v[8 * j + 4]), "f"(v[8 * j + 5]), "f"(v[8 * j + 6]), "f"(v[8 * j + 7]) : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (p.out_hl) {
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[2 * j], h0, l0);
          split_bf16(v[2 * j + 1], h1, l1);
          hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        __nv_bfloat16* oh = p.out_hl + pix * p.oh_stride + p.oh_coff + h * 32;
        __nv_bfloat16* ol = oh + p.oh_plane;
        if (p.al32 & 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(oh + 16 * j), "r"(hw[8 * j]), "r"(hw[8 * j + 1]),
                         "r"(hw[8 * j + 2]), "r"(hw[8 * j + 3]), "r"(hw[8 * j + 4]), "r"(hw[8 * j + 5]), "r"(hw[8 * j + 6]), "r"(hw[8 * j + 7]) : "memory");
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ol + 16 * j), "r"(lw[8 * j]), "r"(lw[8 * j + 1]),
                         "r"(lw[8 * j + 2]), "r"(lw[8 * j + 3]), "r"(lw[8 * j + 4]), "r"(lw[8 * j + 5]), "r"(lw[8 * j + 6]), "r"(lw[8 * j + 7]) : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            reinterpret_cast<uint4*>(oh)[j] = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
            reinterpret_cast<uint4*>(ol)[j] = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
          }
        }
      }
    }
    if (STATS) {
      // InstanceNorm partial sums stay in registers (this thread's pixel column, 32 channels) until the CTA's segment ends
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float t = valid ? v[i] : 0.f;
        acc1[i] += t;
        acc2[i] = fmaf(t, t, acc2[i]);
      }
    }
  }
}

// End of a CTA segment (consecutive output rows of one image strip): the warp's per-pixel-column sums are reduced across its 32
// pixels (31-shuffle transpose-reduce) and written as ONE row of partial sums per (segment, warp): a handful of rows per image
// instead of four per output row, deterministic.  ordinal = this CTA's rank among the CTAs that share the strip.
__device__ __forceinline__ void rows_stats_flush(const RowsParams& p, int img, int xt, int strip, int q, int h, int lane, float (&acc1)[32],
                                                 float (&acc2)[32]) {
  const int first = (int)((((long long)strip * p.H + 1) * gridDim.x - 1) / p.U);        // CTA that owns the strip's first unit
  const int ordinal = (int)blockIdx.x - first;
  const float m1 = transpose_reduce32_rows(acc1, lane), m2 = transpose_reduce32_rows(acc2, lane);
  float* o = p.stats + (((long long)img * p.stat_rows + (long long)(xt * p.stat_k + ordinal) * 4 + q) * 2) * CR_C + h * 32 + lane;
  o[0] = m1;
  o[CR_C] = m2;
#pragma unroll
  for (int i = 0; i < 32; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
}

// MODE bit 0: residual input, bit 1: InstanceNorm partial sums, bit 2: ALIAS slot scheme.
// ALIAS: the eight 64-column TMEM slots are a ring of SIX logical slots (output row with sequence number q lives in slot q % 6) plus
// two alias slots.  A step's window of n <= 3 rows starts at logical slot s = q(lo) % 6 and is written to the physical slots
// s .. s + n - 1 WITHOUT wrapping - a position past slot 5 lands in alias slot 6 or 7 - so every product is ONE MMA (the plain
// ring of eight issues N = 128 + 64 on the two of eight steps whose window wraps: +12 % MMA time).  Rows of logical slot 0 / 1
// therefore hold part of their sum in slot 6 / 7; the epilogue adds the two parts.  Slots are handed back zeroed (tcgen05.st), so
// every MMA accumulates and the "first product of a row" special case disappears as well.
template <int MODE>
__global__ void __launch_bounds__(64 + 32 * CR_EW, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const RowsParams p) {
  constexpr bool RES = (MODE & 1) != 0, STATS = (MODE & 2) != 0, ALIAS = (MODE & 4) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = sb, bar_empty = sb + 16, bar_w = sb + 32, bar_tfull = sb + 64, bar_tempty = sb + 128, tmem_slot = sb + 192,
                 bias_s = sb + 256;
  const uint32_t w0 = sb + 1024, a0 = w0 + CR_WBYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
    for (int s = 0; s < CR_ASTAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_w, 1);
    for (int s = 0; s < CR_SLOTS; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, CR_EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // weights, once: tap (ky, kx) of plane pl lands in the kx stack at row block 2 - ky (the slot order out[r-1], out[r], out[r+1])
      mbar_arrive_expect_tx(bar_w, CR_WBYTES);
      for (int pl = 0; pl < 2; ++pl)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            tma_load_4d(w0 + pl * CR_WPLANE + kx * CR_WKX + (2 - ky) * CR_WBLK, &tmW, bar_w, 0, 0, ky * 3 + kx, pl);
      int stage = 0;
      uint32_t phase = 0;
      rows_for_each_step(p, [&](int img, int xt, int r, int, int, uint32_t) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        const uint32_t full = bar_full + 8 * stage, dst = a0 + stage * CR_ASTAGE;
        if (p.dbg & 4) mbar_arrive(full);
        else {
          mbar_arrive_expect_tx(full, 2 * CR_AROWS * 128u);
          tma_load_5d(dst, &tmA, full, 0, xt * CR_M - 1, r, img, 0);
          tma_load_5d(dst + CR_APLANE, &tmA, full, 0, xt * CR_M - 1, r, img, 1);
        }
        if (++stage == CR_ASTAGES) { stage = 0; phase ^= 1u; }
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(CR_M, 64), idesc2 = make_idesc_bf16(CR_M, 128), idesc3 = make_idesc_bf16(CR_M, 192);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(bar_w, 0);
      tc_fence_after();
      if constexpr (ALIAS) {
        // ---- alias-slot scheme, software-pipelined: prep(i + 1) runs between the MMAs and the commits of step i
        struct Prep { uint32_t d1, i1, stage, c0, c1; uint64_t a_hi0, a_lo0, b_hi0, b_lo0; };
        auto prep = [&](const RowsStepIter<false>& s, int stg, uint32_t ph) -> Prep {
          const int r = s.r, y0 = s.y0, y1 = s.y1;
          const int lo = r - 1 > y0 ? r - 1 : y0, hi = r + 1 < y1 - 1 ? r + 1 : y1 - 1;
          // rows that enter the window with this input row (out[r+1]; every row of the window at the segment's first input row)
          // take over their slot - and its alias for logical slots 0 / 1 - once the previous occupant has been read and zeroed
          const bool seg_first = r == (y0 > 0 ? y0 - 1 : 0);
          for (int y = seg_first ? lo : r + 1; y <= hi; ++y) {
            const uint32_t sq = s.seq + (uint32_t)(y - y0), g = sq / 6u, l = sq - g * 6u;
            mbar_wait(bar_tempty + 8 * l, g & 1u);
            if (l < 2u) mbar_wait(bar_tempty + 8 * (l + 6u), g & 1u);
          }
          mbar_wait(bar_full + 8 * stg, ph);
          tc_fence_after();
          Prep P;
          const uint32_t sql = s.seq + (uint32_t)(lo - y0), base = sql - (sql / 6u) * 6u;
          const int n = hi - lo + 1;
          P.d1 = tmem_base + base * 64u; P.i1 = n == 3 ? idesc3 : n == 2 ? idesc2 : idesc1; P.stage = (uint32_t)stg;
          const uint32_t at = a0 + stg * CR_ASTAGE;
          P.a_hi0 = make_smem_desc_sw128(at, 1024); P.a_lo0 = make_smem_desc_sw128(at + CR_APLANE, 1024);
          const uint32_t wrow = w0 + (uint32_t)(lo - r + 1) * CR_WBLK;
          P.b_hi0 = make_smem_desc_sw128(wrow, 1024); P.b_lo0 = make_smem_desc_sw128(wrow + CR_WPLANE, 1024);
          // finished rows -> barrier addresses (0 = none)
          P.c0 = P.c1 = 0u;
          if (lo == r - 1) { const uint32_t sq = s.seq + (uint32_t)(r - 1 - y0); P.c0 = bar_tfull + 8 * (sq - (sq / 6u) * 6u); }
          if (r == p.H - 1 && y1 == p.H) { const uint32_t sq = s.seq + (uint32_t)(r - y0); P.c1 = bar_tfull + 8 * (sq - (sq / 6u) * 6u); }
          return P;
        };
        RowsStepIter<false> itr(p);
        bool have = itr.next();
        Prep P = {};
        if (have) P = prep(itr, stage, phase);
        auto issue_groups = [&](const Prep& P, int kk0, int kk1) {
          if (p.dbg & 1) return;
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            const int kx = kk >> 2, k = kk & 3;
            const uint64_t ao = (uint64_t)((kx * 128 + k * 32) >> 4), wo = (uint64_t)((kx * (int)CR_WKX + k * 32) >> 4);
            umma_bf16(P.d1, P.a_hi0 + ao, P.b_hi0 + wo, P.i1, 1u);
            umma_bf16(P.d1, P.a_hi0 + ao, P.b_lo0 + wo, P.i1, 1u);
            umma_bf16(P.d1, P.a_lo0 + ao, P.b_hi0 + wo, P.i1, 1u);
          }
        };
        while (have) {
          issue_groups(P, 0, 8);
          // the next step is prepared while these MMAs are queued
          if (++stage == CR_ASTAGES) { stage = 0; phase ^= 1u; }
          have = itr.next();
          Prep Pn = {};
          if (have) Pn = prep(itr, stage, phase);
          issue_groups(P, 8, 12);
          umma_commit(bar_empty + 8 * P.stage);
          if (P.c0) umma_commit(P.c0);
          if (P.c1) umma_commit(P.c1);
          P = Pn;
        }
      } else
      rows_for_each_step(p, [&](int, int, int r, int y0, int y1, uint32_t seq) {
        const int lo = r - 1 > y0 ? r - 1 : y0, hi = r + 1 < y1 - 1 ? r + 1 : y1 - 1;
        // output rows touched for the first time by this input row: out[r+1], and out[0] at the top of an image
        const int n_fresh = (hi == r + 1 ? 1 : 0) + ((r == 0 && lo == 0) ? 1 : 0);
        for (int y = hi - n_fresh + 1; y <= hi; ++y) {
          const uint32_t sq = seq + (uint32_t)(y - y0);
          mbar_wait(bar_tempty + 8 * (sq & 7u), ((sq >> 3) & 1u) ^ 1u);
        }
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        // the step's output rows lo..hi sit in consecutive ring slots; a window that wraps is issued as two column ranges.
        // Everything that depends on the step is computed here once: the 36 products below only add immediates.
        const int sa = (int)((seq + (uint32_t)(lo - y0)) & 7u), n = hi - lo + 1;
        const int n1 = n < CR_SLOTS - sa ? n : CR_SLOTS - sa, n2 = n - n1;
        const uint32_t d1 = tmem_base + (uint32_t)(sa * 64), d2 = tmem_base;
        const uint32_t i1 = n1 == 3 ? idesc3 : n1 == 2 ? idesc2 : idesc1, i2 = n2 == 2 ? idesc2 : idesc1;
        const uint32_t at = a0 + stage * CR_ASTAGE;
        const uint64_t a_hi0 = make_smem_desc_sw128(at, 1024), a_lo0 = make_smem_desc_sw128(at + CR_APLANE, 1024);
        const uint32_t wrow = w0 + (uint32_t)(lo - r + 1) * CR_WBLK;                    // ky block of row lo
        const uint64_t b1_hi0 = make_smem_desc_sw128(wrow, 1024), b1_lo0 = make_smem_desc_sw128(wrow + CR_WPLANE, 1024);
        const uint64_t b2_hi0 = b1_hi0 + (uint64_t)(((uint32_t)n1 * CR_WBLK) >> 4), b2_lo0 = b1_lo0 + (uint64_t)(((uint32_t)n1 * CR_WBLK) >> 4);
        if (!(p.dbg & 1)) {
          {
            // first product of the step: rows opened by this input row start from zero, the others accumulate
            const int nf = n_fresh;
            if (nf == 0) {
              umma_bf16(d1, a_hi0, b1_hi0, i1, 1u);
              if (n2) umma_bf16(d2, a_hi0, b2_hi0, i2, 1u);
            } else {
              for (int y = lo; y <= hi; ++y) {      // row by row (N = 64 each): at most three instructions, once per step
                const uint32_t sl = (seq + (uint32_t)(y - y0)) & 7u;
                umma_bf16(tmem_base + sl * 64u, a_hi0, b1_hi0 + (uint64_t)(((uint32_t)(y - lo) * CR_WBLK) >> 4), idesc1, y > hi - nf ? 0u : 1u);
              }
            }
            umma_bf16(d1, a_hi0, b1_lo0, i1, 1u);
            umma_bf16(d1, a_lo0, b1_hi0, i1, 1u);
            if (n2) { umma_bf16(d2, a_hi0, b2_lo0, i2, 1u); umma_bf16(d2, a_lo0, b2_hi0, i2, 1u); }
          }
#pragma unroll
          for (int kk = 1; kk < 12; ++kk) {
            const int kx = kk >> 2, k = kk & 3;
            const uint64_t ao = (uint64_t)((kx * 128 + k * 32) >> 4), wo = (uint64_t)((kx * (int)CR_WKX + k * 32) >> 4);
            umma_bf16(d1, a_hi0 + ao, b1_hi0 + wo, i1, 1u);
            umma_bf16(d1, a_hi0 + ao, b1_lo0 + wo, i1, 1u);
            umma_bf16(d1, a_lo0 + ao, b1_hi0 + wo, i1, 1u);
            if (n2) {
              umma_bf16(d2, a_hi0 + ao, b2_hi0 + wo, i2, 1u);
              umma_bf16(d2, a_hi0 + ao, b2_lo0 + wo, i2, 1u);
              umma_bf16(d2, a_lo0 + ao, b2_hi0 + wo, i2, 1u);
            }
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == CR_ASTAGES) { stage = 0; phase ^= 1u; }
        if (lo == r - 1) umma_commit(bar_tfull + 8 * ((seq + (uint32_t)(r - 1 - y0)) & 7u));
        if (r == p.H - 1 && y1 == p.H) umma_commit(bar_tfull + 8 * ((seq + (uint32_t)(r - y0)) & 7u));
      });
    }
  } else {
    // ================= epilogue: thread = pixel (TMEM lane) of the finished output row
    // the two warps of a lane quarter take one 32-channel half each
    const int q = warp & 3, h = (warp - 2) >> 2, et = (int)threadIdx.x - 64;
    if (et < CR_C) {
      const float b = p.bias ? __ldg(p.bias + et) : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * (uint32_t)et), "f"(b) : "memory");
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * CR_EW) : "memory");
    float acc1[STATS ? 32 : 1], acc2[STATS ? 32 : 1];
#pragma unroll
    for (int i = 0; i < (STATS ? 32 : 1); ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
    if constexpr (ALIAS) {
      // all eight slots start zeroed; this first release completes phase 0 of every slot barrier
      for (uint32_t ps = 0; ps < (uint32_t)CR_SLOTS; ++ps) tmem_zero32(tmem_base + ((uint32_t)(q * 32) << 16) + ps * 64u + (uint32_t)(h * 32));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (uint32_t ps = 0; ps < (uint32_t)CR_SLOTS; ++ps) mbar_arrive(bar_tempty + 8 * ps);
    }
    rows_for_each_step(p, [&](int img, int xt, int r, int y0, int y1, uint32_t seq) {
      auto do_row = [&](int y) {
        rows_epilogue_row<RES, STATS, ALIAS>(p, tmem_base, bar_tfull, bar_tempty, bias_s, seq + (uint32_t)(y - y0), img, xt, y, q, h, lane, acc1, acc2);
        if constexpr (STATS) { if (y == y1 - 1) rows_stats_flush(p, img, xt, img * p.TX + xt, q, h, lane, acc1, acc2); }
      };
      const int lo = r - 1 > y0 ? r - 1 : y0;
      if (lo == r - 1) do_row(r - 1);
      if (r == p.H - 1 && y1 == p.H) do_row(p.H - 1);
    });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------------------------------------------
// The encoders' stem (resnet.py / raft_encoder.py: 7x7 stride-2 convolution 3 -> 64) in the same rolling-rows form.  Its input is
// the x-folded image (im2col along x: 7 taps x 3 channels = 21 -> 32 channels per OUTPUT column), so what remains is a 7x1
// convolution with vertical stride 2:  out[y] = sum_ky W7[ky] . in[2y + ky - 3].  Input row r therefore feeds the output rows
// y = (r + 3 - ky) / 2 for the ky of r's parity class: even rows three of them (ky = 5, 3, 1), odd rows four (ky = 6, 4, 2, 0) - again
// consecutive output rows = consecutive TMEM slots, so ONE MMA per product of N = 192 / 256 over a per-parity weight stack replaces
// 3 / 4 MMAs of N = 64: 12 MMAs per output row instead of 42.  K = 32 per row (two k-steps), tiles are 128 pixels x 64 B
// (SWIZZLE_64B), four 16 KB stages.  The kernel is bound by its epilogue (32 KB of fp32 output per row), which it shares with the
// 3x3 kernel above.
constexpr int CS_ASTAGES = 4;
constexpr uint32_t CS_APLANE = CR_M * 64u, CS_ASTAGE = 2 * CS_APLANE;             // 8 KB per plane
constexpr uint32_t CS_WBLK = CR_C * 64u;                                          // one tap, one plane: 64 rows x 64 B
constexpr uint32_t CS_WODD = 0, CS_WEVEN = 4 * CS_WBLK, CS_WPLANE = 7 * CS_WBLK;   // [ky6|ky4|ky2|ky0] then [ky5|ky3|ky1]
constexpr uint32_t CS_WBYTES = 2 * CS_WPLANE;                                     // 56 KB
constexpr int CS_SMEM = 1024 + 1024 + (int)CS_WBYTES + CS_ASTAGES * (int)CS_ASTAGE;
// FOLD variant: the kernel reads the fp32 NCHW image itself and forms the x-folded split-bf16 tile in shared memory (two converter
// warps), which removes the im2col_x_split launch and its 4.2 MB-per-image intermediate.  Raw stage = two TMA boxes of
// [3 channels][132 columns] fp32 covering input columns 256*xt - 4 .. 256*xt + 259.
constexpr int CS_RAW_COLS = 132, CS_CONV_WARPS = 2;
constexpr uint32_t CS_RAW_BOX = 3 * CS_RAW_COLS * 4, CS_RAW_BOX_PITCH = 1664, CS_RAW_STAGE = 2 * CS_RAW_BOX_PITCH;      // TMA destinations: 128 B aligned
constexpr int CS_SMEM_FOLD = CS_SMEM + CS_ASTAGES * (int)CS_RAW_STAGE;

// schedule of the stem: a segment with output rows [y0, y1) consumes input rows max(2*y0 - 3, 0) .. min(2*(y1-1) + 3, Hin - 1)
template <class F>
__device__ __forceinline__ void stem_for_each_step(const RowsParams& p, F&& f) {
  long long u = p.U * blockIdx.x / gridDim.x;
  const long long u1 = p.U * (blockIdx.x + 1) / gridDim.x;
  uint32_t seq = 0;
  while (u < u1) {
    const int strip = (int)(u / p.H), y0 = (int)(u - (long long)strip * p.H);
    const int y1 = (u1 - u) < (long long)(p.H - y0) ? y0 + (int)(u1 - u) : p.H;
    const int img = strip / p.TX, xt = strip - img * p.TX;
    const int r0 = 2 * y0 - 3 > 0 ? 2 * y0 - 3 : 0, r1 = 2 * (y1 - 1) + 3 < p.Hin - 1 ? 2 * (y1 - 1) + 3 : p.Hin - 1;
    for (int r = r0; r <= r1; ++r) {
      const int m = r >> 1;                                    // window of output rows: m-1 .. m+1 (even r) / m+2 (odd r)
      const int lo = m - 1 > y0 ? m - 1 : y0, top = m + 1 + (r & 1), hi = top < y1 - 1 ? top : y1 - 1;
      if (lo <= hi) f(img, xt, r, y0, y1, lo, hi, seq);
    }
    seq += (uint32_t)(y1 - y0);
    u += y1 - y0;
  }
}

// ALIAS: the slot scheme of conv_rows_kernel (six logical + two alias slots, slots handed back zeroed, every MMA accumulates); a
// four-row window that starts at logical slot 5 puts its last row at that row's home slot 2 with a second MMA (one step in twelve).
template <bool STATS, bool FOLD, bool ALIAS>
__global__ void __launch_bounds__(64 + 32 * CR_EW + (FOLD ? 32 * CS_CONV_WARPS : 0), 1)
conv_stem_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = sb, bar_empty = sb + 32, bar_w = sb + 64, bar_tfull = sb + 128, bar_tempty = sb + 192, tmem_slot = sb + 72,
                 bias_s = sb + 256, bar_rfull = sb + 512, bar_rempty = sb + 544;
  const uint32_t w0 = sb + 1024, a0 = w0 + CS_WBYTES, raw0 = a0 + CS_ASTAGES * CS_ASTAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
    for (int s = 0; s < CS_ASTAGES; ++s) {
      mbar_init(bar_full + 8 * s, FOLD ? CS_CONV_WARPS : 1); mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, CS_CONV_WARPS);
    }
    mbar_init(bar_w, 1);
    for (int s = 0; s < CR_SLOTS; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, CR_EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  if (FOLD && warp >= 2 + CR_EW) {
    // ================= converters: raw fp32 image row -> x-folded split-bf16 tile (element kx * 3 + c = in[c][2x + kx - 3]; 21 of 32)
    const int ct = (int)threadIdx.x - 32 * (2 + CR_EW);          // 0..63: output columns ct and ct + 64 of the strip
    int stage = 0;
    uint32_t phase = 0;
    stem_for_each_step(p, [&](int, int, int, int, int, int, int, uint32_t) {
      mbar_wait(bar_rfull + 8 * stage, phase);
      mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
      const uint32_t rs = raw0 + stage * CS_RAW_STAGE, as = a0 + stage * CS_ASTAGE;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int x = ct + 64 * half;
        float v[24];
#pragma unroll
        for (int e = 0; e < 21; ++e) {
          const int kx = e / 3, c = e - 3 * kx;
          const int idx = 2 * x + kx + 1;                         // column relative to the first box (input x = 256 xt - 4 + idx)
          const uint32_t off = (uint32_t)((c * CS_RAW_COLS + idx) * 4) + (idx >= CS_RAW_COLS ? CS_RAW_BOX_PITCH - CS_RAW_COLS * 4u : 0u);
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[e]) : "r"(rs + off));
        }
        v[21] = v[22] = v[23] = 0.f;
        uint32_t hw[12], lw[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[2 * i], h0, l0);
          split_bf16(v[2 * i + 1], h1, l1);
          hw[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        // SWIZZLE_64B: 16 B unit j of row x sits at unit j ^ ((x >> 1) & 3)
        const uint32_t row = as + (uint32_t)x * 64u, sw = ((uint32_t)x >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = row + ((((uint32_t)j) ^ sw) << 4);
          if (j < 3) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hw[4 * j]), "r"(hw[4 * j + 1]), "r"(hw[4 * j + 2]), "r"(hw[4 * j + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + CS_APLANE), "r"(lw[4 * j]), "r"(lw[4 * j + 1]), "r"(lw[4 * j + 2]), "r"(lw[4 * j + 3]) : "memory");
          } else {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + CS_APLANE), "r"(0u) : "memory");
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_full + 8 * stage); mbar_arrive(bar_rempty + 8 * stage); }
      if (++stage == CS_ASTAGES) { stage = 0; phase ^= 1u; }
    });
  } else if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, CS_WBYTES);
      for (int pl = 0; pl < 2; ++pl)
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t blk = (ky & 1) ? CS_WEVEN + (uint32_t)((5 - ky) >> 1) * CS_WBLK : CS_WODD + (uint32_t)((6 - ky) >> 1) * CS_WBLK;
          tma_load_4d(w0 + pl * CS_WPLANE + blk, &tmW, bar_w, 0, 0, ky, pl);
        }
      int stage = 0;
      uint32_t phase = 0;
      stem_for_each_step(p, [&](int img, int xt, int r, int, int, int, int, uint32_t) {
        if (FOLD) {
          mbar_wait(bar_rempty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_rfull + 8 * stage, dst = raw0 + stage * CS_RAW_STAGE;
          mbar_arrive_expect_tx(full, 2 * CS_RAW_BOX);
          tma_load_4d(dst, &tmA, full, 2 * xt * CR_M - 4, r, 0, img);
          tma_load_4d(dst + CS_RAW_BOX_PITCH, &tmA, full, 2 * xt * CR_M - 4 + CS_RAW_COLS, r, 0, img);
        } else {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8 * stage, dst = a0 + stage * CS_ASTAGE;
          if (p.dbg & 4) mbar_arrive(full);
          else {
            mbar_arrive_expect_tx(full, CS_ASTAGE);
            tma_load_5d(dst, &tmA, full, 0, xt * CR_M, r, img, 0);
            tma_load_5d(dst + CS_APLANE, &tmA, full, 0, xt * CR_M, r, img, 1);
          }
        }
        if (++stage == CS_ASTAGES) { stage = 0; phase ^= 1u; }
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id1 = make_idesc_bf16(CR_M, 64), id2 = make_idesc_bf16(CR_M, 128), id3 = make_idesc_bf16(CR_M, 192), id4 = make_idesc_bf16(CR_M, 256);
      auto idesc_of = [&](int n) { return n == 4 ? id4 : n == 3 ? id3 : n == 2 ? id2 : id1; };
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(bar_w, 0);
      tc_fence_after();
      long long tacc[6] = {0, 0, 0, 0, 0, 0};
      long long tprev = clock64();
      const long long tstart = tprev;
      auto tick = [&](int k) { if (p.dbg_times) { const long long t = clock64(); tacc[k] += t - tprev; tprev = t; } };
      if constexpr (ALIAS) {
        // ---- alias-slot scheme, software-pipelined: prep(i + 1) runs between the MMAs and the commits of step i
        struct Prep { uint32_t d1, d2, in1, n2, stage, c0, c1, c2, c3; uint64_t a_hi0, a_lo0, b_hi0, b_lo0, bo2; };
        auto prep = [&](const RowsStepIter<true>& st, int stg, uint32_t ph) -> Prep {
          const int r = st.r, y0 = st.y0, y1 = st.y1, m = r >> 1;
          const int lo = m - 1 > y0 ? m - 1 : y0, top = m + 1 + (r & 1), hi = top < y1 - 1 ? top : y1 - 1, n = hi - lo + 1;
          const bool seg_first = r == (2 * y0 - 3 > 0 ? 2 * y0 - 3 : 0);
          for (int y = seg_first ? lo : m + 2; y <= hi && (seg_first || (r & 1)); ++y) {
            const uint32_t sq = st.seq + (uint32_t)(y - y0), g = sq / 6u, l = sq - g * 6u;
            mbar_wait(bar_tempty + 8 * l, g & 1u);
            if (l < 2u) mbar_wait(bar_tempty + 8 * (l + 6u), g & 1u);
          }
          tick(1);
          mbar_wait(bar_full + 8 * stg, ph);
          tc_fence_after();
          tick(2);
          Prep P;
          const uint32_t sql = st.seq + (uint32_t)(lo - y0), base = sql - (sql / 6u) * 6u;
          const int n1 = (int)base + n <= CR_SLOTS ? n : CR_SLOTS - (int)base;                   // n2 = 1 only for base 5, n 4
          P.n2 = (uint32_t)(n - n1);
          P.d1 = tmem_base + base * 64u; P.d2 = tmem_base + 2u * 64u; P.in1 = idesc_of(n1); P.stage = (uint32_t)stg;
          const uint32_t at = a0 + stg * CS_ASTAGE;
          P.a_hi0 = make_smem_desc_sw64(at, 512); P.a_lo0 = make_smem_desc_sw64(at + CS_APLANE, 512);
          const uint32_t wrow = w0 + ((r & 1) ? CS_WODD : CS_WEVEN) + (uint32_t)(lo - (m - 1)) * CS_WBLK;
          P.b_hi0 = make_smem_desc_sw64(wrow, 512); P.b_lo0 = make_smem_desc_sw64(wrow + CS_WPLANE, 512);
          P.bo2 = (uint64_t)(((uint32_t)n1 * CS_WBLK) >> 4);
          // finished rows (2y + 3 == r; every open row at the last input row) -> barrier addresses, 0 = none
          auto done_bar = [&](int j) -> uint32_t {
            const int y = lo + j;
            if (y > hi || !(2 * y + 3 == r || r == p.Hin - 1)) return 0u;
            const uint32_t sq = st.seq + (uint32_t)(y - y0);
            return bar_tfull + 8 * (sq - (sq / 6u) * 6u);
          };
          P.c0 = done_bar(0); P.c1 = done_bar(1); P.c2 = done_bar(2); P.c3 = done_bar(3);
          tick(0);
          return P;
        };
        RowsStepIter<true> itr(p);
        bool have = itr.next();
        Prep P = {};
        if (have) P = prep(itr, stage, phase);
        auto issue_k = [&](const Prep& P, int k) {
          if (p.dbg & 1) return;
          const uint64_t ko = (uint64_t)(k * 32 >> 4);
          umma_bf16(P.d1, P.a_hi0 + ko, P.b_hi0 + ko, P.in1, 1u);
          umma_bf16(P.d1, P.a_hi0 + ko, P.b_lo0 + ko, P.in1, 1u);
          umma_bf16(P.d1, P.a_lo0 + ko, P.b_hi0 + ko, P.in1, 1u);
          if (P.n2) {
            umma_bf16(P.d2, P.a_hi0 + ko, P.b_hi0 + P.bo2 + ko, id1, 1u);
            umma_bf16(P.d2, P.a_hi0 + ko, P.b_lo0 + P.bo2 + ko, id1, 1u);
            umma_bf16(P.d2, P.a_lo0 + ko, P.b_hi0 + P.bo2 + ko, id1, 1u);
          }
        };
        while (have) {
          issue_k(P, 0);
          tick(3);
          // the next step is prepared while the first k-step's MMAs are queued
          if (++stage == CS_ASTAGES) { stage = 0; phase ^= 1u; }
          have = itr.next();
          Prep Pn = {};
          if (have) Pn = prep(itr, stage, phase);
          issue_k(P, 1);
          tick(3);
          umma_commit(bar_empty + 8 * P.stage);
          if (P.c0) umma_commit(P.c0);
          if (P.c1) umma_commit(P.c1);
          if (P.c2) umma_commit(P.c2);
          if (P.c3) umma_commit(P.c3);
          tick(4);
          P = Pn;
        }
      } else
      stem_for_each_step(p, [&](int, int, int r, int y0, int, int lo, int hi, uint32_t seq) {
        tick(0);                                  // schedule arithmetic between steps
        const int m = r >> 1, n = hi - lo + 1;
        // rows opened by this input row: y with 2y - 3 == r (odd r: the window's top row), and every row when r == 0
        const int n_fresh = r == 0 ? n : ((r & 1) && hi == m + 2 ? 1 : 0);
        for (int y = hi - n_fresh + 1; y <= hi; ++y) {
          const uint32_t sq = seq + (uint32_t)(y - y0);
          mbar_wait(bar_tempty + 8 * (sq & 7u), ((sq >> 3) & 1u) ^ 1u);
        }
        tick(1);                                  // waiting for TMEM slots
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        tick(2);                                  // waiting for the operand tile
        const int sa = (int)((seq + (uint32_t)(lo - y0)) & 7u);
        const int n1 = n < CR_SLOTS - sa ? n : CR_SLOTS - sa, n2 = n - n1;
        const uint32_t d1 = tmem_base + (uint32_t)(sa * 64), d2 = tmem_base;
        const uint32_t at = a0 + stage * CS_ASTAGE;
        const uint64_t a_hi0 = make_smem_desc_sw64(at, 512), a_lo0 = make_smem_desc_sw64(at + CS_APLANE, 512);
        // weight block of row lo inside this parity's stack: blocks are ordered by ascending output row (descending ky)
        const uint32_t wrow = w0 + ((r & 1) ? CS_WODD : CS_WEVEN) + (uint32_t)(lo - (m - 1)) * CS_WBLK;
        const uint64_t b1_hi0 = make_smem_desc_sw64(wrow, 512), b1_lo0 = make_smem_desc_sw64(wrow + CS_WPLANE, 512);
        const uint64_t bo2 = (uint64_t)(((uint32_t)n1 * CS_WBLK) >> 4);
        const uint32_t in1 = idesc_of(n1), in2 = idesc_of(n2);
        if (!(p.dbg & 1)) {
          if (n_fresh == 0) {
            umma_bf16(d1, a_hi0, b1_hi0, in1, 1u);
            if (n2) umma_bf16(d2, a_hi0, b1_hi0 + bo2, in2, 1u);
          } else {
            for (int y = lo; y <= hi; ++y) {        // first product row by row: fresh rows start from zero
              const uint32_t sl = (seq + (uint32_t)(y - y0)) & 7u;
              umma_bf16(tmem_base + sl * 64u, a_hi0, b1_hi0 + (uint64_t)(((uint32_t)(y - lo) * CS_WBLK) >> 4), id1, y > hi - n_fresh ? 0u : 1u);
            }
          }
          const bool all3 = !(p.dbg & 16);                    // timing experiment: bit 16 = hi*hi products only
          if (all3) {
            umma_bf16(d1, a_hi0, b1_lo0, in1, 1u);
            umma_bf16(d1, a_lo0, b1_hi0, in1, 1u);
            if (n2) { umma_bf16(d2, a_hi0, b1_lo0 + bo2, in2, 1u); umma_bf16(d2, a_lo0, b1_hi0 + bo2, in2, 1u); }
          }
          const uint64_t ko = (uint64_t)(32 >> 4);            // second k-step: channels 16..31
          umma_bf16(d1, a_hi0 + ko, b1_hi0 + ko, in1, 1u);
          if (all3) {
            umma_bf16(d1, a_hi0 + ko, b1_lo0 + ko, in1, 1u);
            umma_bf16(d1, a_lo0 + ko, b1_hi0 + ko, in1, 1u);
          }
          if (n2) {
            umma_bf16(d2, a_hi0 + ko, b1_hi0 + bo2 + ko, in2, 1u);
            if (all3) {
              umma_bf16(d2, a_hi0 + ko, b1_lo0 + bo2 + ko, in2, 1u);
              umma_bf16(d2, a_lo0 + ko, b1_hi0 + bo2 + ko, in2, 1u);
            }
          }
        }
        tick(3);                                  // descriptors + MMA issue
        umma_commit(bar_empty + 8 * stage);
        if (++stage == CS_ASTAGES) { stage = 0; phase ^= 1u; }
        // finished rows: 2y + 3 == r, and at the last input row every row still open
        for (int y = lo; y <= hi; ++y)
          if (2 * y + 3 == r || r == p.Hin - 1) umma_commit(bar_tfull + 8 * ((seq + (uint32_t)(y - y0)) & 7u));
        tick(4);                                  // commits
      });
      if (p.dbg_times) {
        for (int k = 0; k < 5; ++k) p.dbg_times[(long long)blockIdx.x * 8 + k] = tacc[k];
        p.dbg_times[(long long)blockIdx.x * 8 + 5] = clock64() - tstart;
      }
    }
  } else {
    const int q = warp & 3, h = (warp - 2) >> 2, et = (int)threadIdx.x - 64;
    if (et < CR_C) {
      const float b = p.bias ? __ldg(p.bias + et) : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * (uint32_t)et), "f"(b) : "memory");
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * CR_EW) : "memory");
    float acc1[STATS ? 32 : 1], acc2[STATS ? 32 : 1];
#pragma unroll
    for (int i = 0; i < (STATS ? 32 : 1); ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
    if constexpr (ALIAS) {
      for (uint32_t ps = 0; ps < (uint32_t)CR_SLOTS; ++ps) tmem_zero32(tmem_base + ((uint32_t)(q * 32) << 16) + ps * 64u + (uint32_t)(h * 32));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (uint32_t ps = 0; ps < (uint32_t)CR_SLOTS; ++ps) mbar_arrive(bar_tempty + 8 * ps);
    }
    stem_for_each_step(p, [&](int img, int xt, int r, int y0, int y1, int lo, int hi, uint32_t seq) {
      for (int y = lo; y <= hi; ++y)
        if (2 * y + 3 == r || r == p.Hin - 1) {
          rows_epilogue_row<false, STATS, ALIAS>(p, tmem_base, bar_tfull, bar_tempty, bias_s, seq + (uint32_t)(y - y0), img, xt, y, q, h, lane, acc1, acc2);
          if constexpr (STATS) { if (y == y1 - 1) rows_stats_flush(p, img, xt, img * p.TX + xt, q, h, lane, acc1, acc2); }
        }
    });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// 7x1 kernel, vertical stride 2, 32 (x-folded) input channels -> 64: the encoders' stem after im2col_x_split
bool conv2d_stem_rows_eligible(const scf_tc_conv_desc& d) {
  const char* e = getenv("SCFLOW_TC_ROWS");
  if (e && atoi(e) == 0) return false;
  const char* se = getenv("SCFLOW_TC_ROWS_STEM");
  if (se && atoi(se) == 0) return false;
  const int sx = d.stride_x ? d.stride_x : (d.stride == 2 ? 2 : 1), sy = d.stride_y ? d.stride_y : (d.stride == 2 ? 2 : 1);
  if (d.kh != 7 || d.kw != 1 || sx != 1 || sy != 2 || d.nseg != 1 || d.seg[0].nch != 32 || d.cin_pad != 32 || d.cout != CR_C ||
      d.cout_pad != CR_C || d.w_batched || d.ksplit > 1 || d.w_plane_stride != 0)
    return false;
  if (d.epi != SCF_EPI_ACT || (d.act != SCF_ACT_NONE && d.act != SCF_ACT_RELU) || d.pre || d.aux0 || d.aux1 || d.out2_hl || d.scale != 1.f) return false;
  if (d.H % 2 || d.H < 8 || d.W < 96 || (long long)d.B * (d.H / 2) * cdiv(d.W, CR_M) < 148) return false;
  auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
  if (!al16(d.seg[0].ptr) || d.seg[0].stride % 8 || d.seg[0].coff % 8 || (d.seg[0].plane_stride * 2) % 16) return false;
  if (d.out_f32 && (!al16(d.out_f32) || d.out_f32_stride % 4 || d.out_f32_coff % 4)) return false;
  if (d.out_hl && (!al16(d.out_hl) || d.out_hl_stride % 8 || d.out_hl_coff % 8 || (d.out_hl_plane * 2) % 16)) return false;
  return true;
}

// images != nullptr: FOLD variant (the kernel x-folds the fp32 NCHW image itself; d.seg is ignored, d.H x 2*d.W is the image size)
static int conv2d_stem_rows_impl(const scf_tc_conv_desc& d, const float* images, cudaStream_t st) {
  RowsParams p = {};
  p.N = d.B; p.Hin = d.H; p.H = d.H / 2; p.W = d.W; p.TX = cdiv(d.W, CR_M);
  p.U = (long long)d.B * p.TX * p.H;
  p.bias = d.bias; p.relu = d.act == SCF_ACT_RELU ? 1 : 0;
  p.out_f32 = d.out_f32; p.of_stride = d.out_f32_stride; p.of_coff = d.out_f32_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.oh_plane = d.out_hl_plane; p.oh_stride = d.out_hl_stride; p.oh_coff = d.out_hl_coff;
  p.stats = d.stats;
  {
    auto al32 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 32 == 0; };
    if (d.out_f32 && al32(d.out_f32) && d.out_f32_stride % 8 == 0 && d.out_f32_coff % 8 == 0) p.al32 |= 1;
    if (d.out_hl && al32(d.out_hl) && d.out_hl_stride % 16 == 0 && d.out_hl_coff % 16 == 0 && (d.out_hl_plane * 2) % 32 == 0) p.al32 |= 2;
    const char* de = getenv("SCFLOW_ROWS_DBG");
    p.dbg = de ? atoi(de) : 0;
    const char* dt = getenv("SCFLOW_ROWS_DBG_TIMES");
    p.dbg_times = dt ? reinterpret_cast<long long*>(strtoull(dt, nullptr, 16)) : nullptr;
  }
  CUtensorMap tmA, tmW;
  if (images) {
    const cuuint64_t Wi = 2ull * d.W;
    cuuint64_t dims[4] = {Wi, (cuuint64_t)d.H, 3, (cuuint64_t)d.B};
    cuuint64_t str[3] = {Wi * 4, (cuuint64_t)d.H * Wi * 4, 3ull * d.H * Wi * 4};
    cuuint32_t box[4] = {(cuuint32_t)CS_RAW_COLS, 1, 3, 1};
    SCF_TRY(encode_map(&tmA, images, 4, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE));
  } else {
    const scf_tc_seg& sg = d.seg[0];
    const char* base = reinterpret_cast<const char*>(sg.ptr) + (size_t)sg.coff * 2;
    cuuint64_t dims[5] = {32, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {(cuuint64_t)sg.stride * 2, (cuuint64_t)d.W * sg.stride * 2, (cuuint64_t)d.H * d.W * sg.stride * 2,
                         (cuuint64_t)sg.plane_stride * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)CR_M, 1, 1, 1};
    SCF_TRY(encode_map(&tmA, base, 5, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    cuuint64_t wd[4] = {32, (cuuint64_t)CR_C, 7, 2};
    cuuint64_t ws[3] = {32 * 2, (cuuint64_t)CR_C * 32 * 2, (cuuint64_t)7 * CR_C * 32 * 2};
    cuuint32_t wb[4] = {32, (cuuint32_t)CR_C, 1, 1};
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: packed weight must be 16B aligned");
    SCF_TRY(encode_map(&tmW, d.w, 4, wd, ws, wb, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  typedef void (*StemFn)(const CUtensorMap, const CUtensorMap, const RowsParams);
  // index: bit 0 InstanceNorm sums, bit 1 FOLD (x-fold in the kernel), bit 2 alias-slot TMEM scheme
  static const StemFn stem_table[8] = {conv_stem_rows_kernel<false, false, false>, conv_stem_rows_kernel<true, false, false>,
                                       conv_stem_rows_kernel<false, true, false>,  conv_stem_rows_kernel<true, true, false>,
                                       conv_stem_rows_kernel<false, false, true>,  conv_stem_rows_kernel<true, false, true>,
                                       conv_stem_rows_kernel<false, true, true>,   conv_stem_rows_kernel<true, true, true>};
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int i = 0; i < 8 && attr_err == cudaSuccess; ++i)
      attr_err = cudaFuncSetAttribute(stem_table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (i & 2) ? CS_SMEM_FOLD : CS_SMEM);
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(conv_stem_rows_kernel): %s", cudaGetErrorString(attr_err));
  const unsigned grid = (unsigned)(p.U < num_sms ? p.U : num_sms);
  {
    // a strip's H units are shared by at most ceil(H / floor(U / grid)) + 1 consecutive CTAs (never more than H)
    const int upc = (int)(p.U / grid), k = cdiv(p.H, upc) + 1;
    p.stat_k = k < p.H ? k : p.H;
    p.stat_rows = p.TX * p.stat_k * 4;
    if (p.stats) SCF_CUDA(cudaMemsetAsync(p.stats, 0, (size_t)p.N * p.stat_rows * 2 * CR_C * sizeof(float), st));
  }
  g_last_m_tiles = (int)p.U; g_last_tiles_per_img = p.TX * p.H; g_last_stat_rows_per_img = p.stat_rows;
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * CR_EW + (images ? 32 * CS_CONV_WARPS : 0));
  cfg.dynamicSmemBytes = images ? CS_SMEM_FOLD : CS_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  const char* ae = getenv("SCFLOW_ROWS_ALIAS");
  const int which = (d.stats ? 1 : 0) | (images ? 2 : 0) | ((ae ? atoi(ae) != 0 : true) ? 4 : 0);
  cudaError_t le = cudaLaunchKernelEx(&cfg, stem_table[which], tmA, tmW, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("conv_stem_rows_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("conv_stem_rows_kernel");
}

int conv2d_stem_rows(const scf_tc_conv_desc& d, cudaStream_t st) { return conv2d_stem_rows_impl(d, nullptr, st); }

// The stem straight from the fp32 NCHW image [B, 3, d.H, 2 * d.W] (no im2col_x_split launch): SCFLOW_STEM_FOLD (default 1)
bool conv2d_stem_rows_fold_ok(const scf_tc_conv_desc& d, const float* images) {
  const char* fe = getenv("SCFLOW_STEM_FOLD");
  if (fe && atoi(fe) == 0) return false;
  scf_tc_conv_desc t = d;
  static const __nv_bfloat16 dummy[8] = {};
  t.seg[0].ptr = dummy; t.seg[0].stride = 32; t.seg[0].coff = 0; t.seg[0].nch = 32; t.seg[0].plane_stride = 8; t.nseg = 1;
  return images && reinterpret_cast<uintptr_t>(images) % 16 == 0 && (2 * d.W) % 4 == 0 && conv2d_stem_rows_eligible(t);
}
int conv2d_stem_rows_fold(const float* images, const scf_tc_conv_desc& d, cudaStream_t st) { return conv2d_stem_rows_impl(d, images, st); }

// SCFLOW_TC_ROWS (default 1): 3x3 / stride 1 / 64 -> 64 channel layers on maps at least 96 pixels wide take this kernel
bool conv2d_rows_eligible(const scf_tc_conv_desc& d) {
  const char* e = getenv("SCFLOW_TC_ROWS");
  if (e && atoi(e) == 0) return false;
  const int sx = d.stride_x ? d.stride_x : (d.stride == 2 ? 2 : 1), sy = d.stride_y ? d.stride_y : (d.stride == 2 ? 2 : 1);
  if (d.kh != 3 || d.kw != 3 || sx != 1 || sy != 1 || d.nseg != 1 || d.seg[0].nch != CR_C || d.cin_pad != CR_C || d.cout != CR_C ||
      d.cout_pad != CR_C || d.w_batched || d.ksplit > 1 || d.w_plane_stride != 0)
    return false;
  if (d.epi != SCF_EPI_ACT || (d.act != SCF_ACT_NONE && d.act != SCF_ACT_RELU) || d.pre || d.aux1 || d.out2_hl || d.scale != 1.f) return false;
  if (d.W < 96 || (long long)d.B * d.H * cdiv(d.W, CR_M) < 148) return false;
  auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
  if (!al16(d.seg[0].ptr) || d.seg[0].stride % 8 || d.seg[0].coff % 8 || (d.seg[0].plane_stride * 2) % 16) return false;
  if (d.out_f32 && (!al16(d.out_f32) || d.out_f32_stride % 4 || d.out_f32_coff % 4)) return false;
  if (d.out_hl && (!al16(d.out_hl) || d.out_hl_stride % 8 || d.out_hl_coff % 8 || (d.out_hl_plane * 2) % 16)) return false;
  if (d.aux0 && (!al16(d.aux0) || d.aux0_stride % 4)) return false;
  if (d.aux0_hl && (d.aux0 || !al16(d.aux0_hl) || d.aux0_hl_stride % 8 || (d.aux0_hl_plane * 2) % 16)) return false;
  return true;
}

int conv2d_rows(const scf_tc_conv_desc& d, cudaStream_t st) {
  RowsParams p = {};
  p.N = d.B; p.H = d.H; p.W = d.W; p.TX = cdiv(d.W, CR_M);
  p.U = (long long)d.B * p.TX * d.H;
  p.bias = d.bias; p.relu = d.act == SCF_ACT_RELU ? 1 : 0;
  p.out_f32 = d.out_f32; p.of_stride = d.out_f32_stride; p.of_coff = d.out_f32_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.oh_plane = d.out_hl_plane; p.oh_stride = d.out_hl_stride; p.oh_coff = d.out_hl_coff;
  p.res = d.aux0; p.res_stride = d.aux0_stride;
  p.res_hl = reinterpret_cast<const __nv_bfloat16*>(d.aux0_hl); p.res_hl_plane = d.aux0_hl_plane; p.res_hl_stride = d.aux0_hl_stride;
  p.stats = d.stats;
  {
    auto al32 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 32 == 0; };
    if (d.out_f32 && al32(d.out_f32) && d.out_f32_stride % 8 == 0 && d.out_f32_coff % 8 == 0) p.al32 |= 1;
    if (d.out_hl && al32(d.out_hl) && d.out_hl_stride % 16 == 0 && d.out_hl_coff % 16 == 0 && (d.out_hl_plane * 2) % 32 == 0) p.al32 |= 2;
    if (d.aux0 && al32(d.aux0) && d.aux0_stride % 8 == 0) p.al32 |= 4;
    if (d.aux0_hl && al32(d.aux0_hl) && d.aux0_hl_stride % 16 == 0 && (d.aux0_hl_plane * 2) % 32 == 0) p.al32 |= 8;
    const char* de = getenv("SCFLOW_ROWS_DBG");
    p.dbg = de ? atoi(de) : 0;
  }
  CUtensorMap tmA, tmW;
  {
    const scf_tc_seg& sg = d.seg[0];
    const char* base = reinterpret_cast<const char*>(sg.ptr) + (size_t)sg.coff * 2;
    cuuint64_t dims[5] = {(cuuint64_t)CR_C, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {(cuuint64_t)sg.stride * 2, (cuuint64_t)d.W * sg.stride * 2, (cuuint64_t)d.H * d.W * sg.stride * 2,
                         (cuuint64_t)sg.plane_stride * 2};
    cuuint32_t box[5] = {(cuuint32_t)CR_C, CR_AROWS, 1, 1, 1};
    SCF_TRY(encode_map(&tmA, base, 5, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B));
    cuuint64_t wd[4] = {(cuuint64_t)CR_C, (cuuint64_t)CR_C, 9, 2};
    cuuint64_t ws[3] = {(cuuint64_t)CR_C * 2, (cuuint64_t)CR_C * CR_C * 2, (cuuint64_t)9 * CR_C * CR_C * 2};
    cuuint32_t wb[4] = {(cuuint32_t)CR_C, (cuuint32_t)CR_C, 1, 1};
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: packed weight must be 16B aligned");
    SCF_TRY(encode_map(&tmW, d.w, 4, wd, ws, wb, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const RowsParams);
  static const KernelFn table[8] = {conv_rows_kernel<0>, conv_rows_kernel<1>, conv_rows_kernel<2>, conv_rows_kernel<3>,
                                    conv_rows_kernel<4>, conv_rows_kernel<5>, conv_rows_kernel<6>, conv_rows_kernel<7>};
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int i = 0; i < 8 && attr_err == cudaSuccess; ++i)
      attr_err = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM);
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(conv_rows_kernel): %s", cudaGetErrorString(attr_err));
  const unsigned grid = (unsigned)(p.U < num_sms ? p.U : num_sms);
  {
    // a strip's H units are shared by at most ceil(H / floor(U / grid)) + 1 consecutive CTAs (never more than H)
    const int upc = (int)(p.U / grid), k = cdiv(p.H, upc) + 1;
    p.stat_k = k < p.H ? k : p.H;
    p.stat_rows = p.TX * p.stat_k * 4;
    if (p.stats) SCF_CUDA(cudaMemsetAsync(p.stats, 0, (size_t)p.N * p.stat_rows * 2 * CR_C * sizeof(float), st));
  }
  g_last_m_tiles = (int)p.U; g_last_tiles_per_img = p.TX * d.H; g_last_stat_rows_per_img = p.stat_rows;
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * CR_EW);
  cfg.dynamicSmemBytes = CR_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  // SCFLOW_ROWS_ALIAS (default 1): six logical TMEM slots + two alias slots, no wrapped windows
  const char* ae = getenv("SCFLOW_ROWS_ALIAS");
  const int mode = ((d.aux0 || d.aux0_hl) ? 1 : 0) | (d.stats ? 2 : 0) | ((ae ? atoi(ae) != 0 : true) ? 4 : 0);
  cudaError_t le = cudaLaunchKernelEx(&cfg, table[mode], tmA, tmW, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("conv_rows_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("conv_rows_kernel");
}

}  // namespace scf
