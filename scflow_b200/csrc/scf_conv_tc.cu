// tcgen05 implicit-GEMM convolution with split-bf16 operands (bf16x3) and fp32 accumulation in TMEM.
//
// GEMM view per CTA: D[128 pixels x BN couts] += sum over (tap, 64-channel chunk) of A_tap[128 x 64] * W_tap[BN x 64]^T.
//   * A tile  : ONE 5-D TMA box {64 ch, TW, TH, 1 sample, 2 planes} of the NHWC split-bf16 activation, fetched at the
//               tap-shifted pixel origin; out-of-image pixels (the convolution's zero padding) and channels beyond the
//               segment are zero-filled by TMA.  128 B rows + SWIZZLE_128B = the canonical K-major UMMA layout.
//   * W tile  : ONE 4-D TMA box {64 cin, BN cout, 1 tap, 2 planes} of the packed weights.
//   * MMA     : per 16-channel k-step three tcgen05.mma (hi*hi, hi*lo, lo*hi), M=128, N=BN, issued by one thread.
//   * epilogue: 4 warps read the accumulator with tcgen05.ld (one TMEM lane = one pixel per thread), apply
//               bias/activation or the GRU gate math, and write fp32 and/or split-bf16 NHWC outputs.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (coalesced through
// a small shared-memory staging buffer).
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
// threads = 64 + 32*EW: warp 0 TMA, warp 1 MMA + TMEM allocator, EW (4 or 8) epilogue warps
constexpr int TC_STAGE_ROW = 80;  // epilogue staging: 32 rows x 64 B per warp, rows padded to 80 B (conflict-free 16 B accesses)
__host__ __device__ constexpr int tc_header(int ew) { return 1024 + ew * 32 * TC_STAGE_ROW + 2048; }   // barriers + staging + 2 bias buffers (multiple of 1024)
constexpr uint32_t TC_A_PLANE = TC_BM * TC_BK * 2;   // 16 KB
constexpr int TC_MAX_STAGES = 4;      // 64-channel stages; 32-channel stages: 7

struct TcParams {
  int nseg, seg_chunks[3], seg_wcoff[3];
  int seg_last_ks[3];            // 16-channel k-steps of a segment's last (possibly ragged) 64-channel chunk: 1..4
  int B, H, W, kh, kw, ph, pw;   // H, W: OUTPUT spatial size
  int sx, sy;                    // 1 or 2 per axis (input is sampled at stride*out + tap - pad)
  int TW, TH, TB, tiles_x, tiles_y;   // tile = TW x TH pixels x TB samples = 128 rows
  int m_tiles, num_tiles;        // pixel tiles, pixel tiles x N tiles (tile t: n = t / m_tiles, m = t % m_tiles)
  int cluster;                   // CTAs per cluster (1, 2 or 4): consecutive pixel tiles of one N tile share each weight tile,
                                 // every CTA fetching BN/cluster rows of it and multicasting them to its peers
  // halo mode (stride-1 multi-tap convolutions, one sample per tile): the activation tile is fetched ONCE per channel chunk
  // together with its (kw-1) x (kh-1) halo - PW x PH pixel rows of 128 B - and every tap reads it through a row-shifted
  // UMMA descriptor.  The tile is 8 x 16 pixels so that one 8-row core-matrix group = one tile row and the group stride
  // (SBO) is the halo pitch PW*128 B.  Cuts the activation bytes a CTA ingests per chunk from taps*32 KB to ~46 KB; the
  // weight tiles (one per tap) go through their own, deeper ring.
  int halo, PW, PH, a_stages, b_stages;
  int pair;                      // 1 (with cluster == 2): the two CTAs run ONE cta_group::2 MMA of M = 256 - each keeps its
                                 // own 128 A rows and only HALF of the weight rows in shared memory (no multicast), which cuts
                                 // the shared-memory traffic per FLOP (operand reads + TMA fills), the resource a single-CTA
                                 // tile saturates first
  int BN, cout, num_taps, w_batched, stages;
  int bk;                        // channels per stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows; twice the stages in the same
                                 // ring, finer hand-over between TMA and the MMA issuer); halo mode is always 64
  int acc_cols, tmem_cols;       // TMEM columns of one accumulator buffer / allocated (two buffers)
  int dbg_epi;                   // timing experiments only (SCFLOW_TC_DBG_EPI): 1 = skip epilogue loads, 2 = skip stores
  long long* dbg_times;          // optional [grid][8] globaltimer stamps of each CTA's first tile (tools/trace_conv_tc.py)
  // stacked-N mode (BN <= 128): the hi and lo weight planes are adjacent in shared memory, so ONE MMA with N = 2*BN
  // forms A_hi*[W_hi;W_lo] into accumulator columns [0,BN) and [BN,2BN); a second MMA adds A_lo*W_hi into [0,BN).
  // Two instructions per k-step instead of three (less shared-memory operand traffic per FLOP); the epilogue adds
  // the two halves.
  int stackn;
  int dualacc;                   // single-CTA, non-stacked, BN <= 128: the k-steps alternate between two accumulators (columns [0,BN)
                                 // and [BN,2BN), summed by the epilogue) so that consecutive MMAs do not form one dependent chain
  int splitacc;                  // experiment (SCFLOW_TC_SPLITACC, single-CTA stacked-N): the A_lo*W_hi MMA accumulates into its own
                                 // TMEM columns [2BN, 3BN) instead of [0, BN), so consecutive MMAs never depend on each other
  int fast_epi;                  // every global access of the epilogue is 16 B aligned: use the coalesced staged path
  int direct_st;                 // ... and every 32-column output block is 32 B aligned: 256-bit row stores, no staging
  const float* bias; float scale; int epi, act;
  float* out_f32; int out_f32_stride, out_f32_coff;
  __nv_bfloat16* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff;
  const float* aux0; int aux0_stride; const float* aux1; int aux1_stride;
  __nv_bfloat16* out2_hl; long long out2_hl_plane; int out2_hl_stride;
  const float* pre; int pre_stride;   // GRU epilogues: fp32 map added before the gate non-linearity
  float* stats;                       // EPI_ACT: [m_tiles][4 warps][2][cout] partial sums / sums of squares of the output
  // tap split (split-K over the kernel taps, for layers with fewer pixel tiles than SMs): tile t = (split ks, N tile, pixel tile);
  // split ks contracts taps [ks * taps_per_split, ...) only and writes its PARTIAL sums (no bias, no activation) to
  // out_f32 + ks * split_stride; the consumer adds the partial maps (scf_pose_head.cu: group norm)
  int ksplit, taps_per_split; long long split_stride;
};

// ---- compact epilogue helpers (the whole epilogue loop body must stay well inside the instruction cache: a first
// version with runtime-switched activations and local-memory staging arrays was ~50 KB of SASS and ran 10x slower)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if (ACT == SCF_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == SCF_ACT_SIGMOID) return sigmoid_fast(v);
  if (ACT == SCF_ACT_TANH) return tanh_fast(v);
  return v;
}

__device__ __forceinline__ void load16(const float* src, float* d) {     // src 16 B aligned
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
    d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
  }
}

// ---- coalescing through shared memory.  In the accumulator layout a thread owns one pixel row, so direct 16 B accesses
// of a warp touch 32 different cache lines (measured: the epilogue of a 128x256 tile took 5.6 us, LSU-bound).  These
// helpers move a [32 rows x 64 B] block between the warp's registers and global memory with 4 lanes per row, i.e.
// 8 fully used 64 B row segments per instruction.  rp[it] = global pixel index of row it*8 + (lane>>2) of this warp's
// 32 rows, or -1 (row outside the image).
__device__ __forceinline__ void stage_store64(uint32_t sbuf, int lane, char* gbase, long long row_bytes, const int (&rp)[4],
                                              const uint4 (&d)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + lane * TC_STAGE_ROW + i * 16), "r"(d[i].x), "r"(d[i].y),
                 "r"(d[i].z), "r"(d[i].w) : "memory");
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), seg = lane & 3;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sbuf + r * TC_STAGE_ROW + seg * 16) : "memory");
    if (rp[it] >= 0) *reinterpret_cast<uint4*>(gbase + rp[it] * row_bytes + seg * 16) = v;
  }
  __syncwarp();
}
// Same for an fp32 block, additionally accumulating per-column sums and sums of squares of the valid rows:
// after the call lane l holds (in s, q) the partial sums of columns (l&3)*4..+3 over rows (l>>2) + 8*it.
__device__ __forceinline__ void stage_store64_stats(uint32_t sbuf, int lane, char* gbase, long long row_bytes, const int (&rp)[4],
                                                    const uint4 (&d)[4], float (&s)[4], float (&q)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + lane * TC_STAGE_ROW + i * 16), "r"(d[i].x), "r"(d[i].y),
                 "r"(d[i].z), "r"(d[i].w) : "memory");
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), seg = lane & 3;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sbuf + r * TC_STAGE_ROW + seg * 16) : "memory");
    if (rp[it] >= 0) {
      *reinterpret_cast<uint4*>(gbase + rp[it] * row_bytes + seg * 16) = v;
      const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
      s[0] += f0; s[1] += f1; s[2] += f2; s[3] += f3;
      q[0] = fmaf(f0, f0, q[0]); q[1] = fmaf(f1, f1, q[1]); q[2] = fmaf(f2, f2, q[2]); q[3] = fmaf(f3, f3, q[3]);
    }
  }
  __syncwarp();
}
// Epilogue INPUTS (GRU: context term, h, z; ACT: residual) are read straight into the accumulator layout: a thread owns one
// pixel row, so it loads its own 64 B (16 fp32 columns) with four 16 B loads.  A warp-wide load touches 32 rows, but every
// 32 B sector is used completely by two consecutive loads of the same thread (merged in the L1 miss queue), and - unlike the
// staged stores below - no shared-memory bandwidth is taken from the tensor core, which runs the next tile's MMAs out of
// shared memory while this epilogue executes.  row_issue starts the loads (they stay in flight while the previous block is
// processed); the registers are consumed as fp32 later.
__device__ __forceinline__ void row_issue(const float* row_ptr, bool valid, uint4 (&r)[4]) {
  // two 256-bit loads (sm_100): every 32 B sector is requested exactly once.  volatile asm: ptxas otherwise sinks the loads
  // down to their first use (to save registers), which turns the software prefetch into a load-and-wait per block
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    r[2 * j] = make_uint4(0u, 0u, 0u, 0u);
    r[2 * j + 1] = make_uint4(0u, 0u, 0u, 0u);
    if (valid)
      asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[2 * j].x), "=r"(r[2 * j].y), "=r"(r[2 * j].z), "=r"(r[2 * j].w), "=r"(r[2 * j + 1].x), "=r"(r[2 * j + 1].y),
                     "=r"(r[2 * j + 1].z), "=r"(r[2 * j + 1].w)
                   : "l"(reinterpret_cast<const char*>(row_ptr) + 32 * j));
  }
}
__device__ __forceinline__ void row_values(const uint4 (&r)[4], float (&out)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    out[4 * j] = __uint_as_float(r[j].x); out[4 * j + 1] = __uint_as_float(r[j].y);
    out[4 * j + 2] = __uint_as_float(r[j].z); out[4 * j + 3] = __uint_as_float(r[j].w);
  }
}
// Direct row stores with 256-bit instructions (sm_100): the thread writes whole 32 B sectors of its own pixel row, so no
// shared-memory transpose (and none of the tensor core's shared-memory bandwidth) is needed.  Require 32 B alignment.
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
// 32 fp32 values of this thread's row -> 128 B
__device__ __forceinline__ void row_store_f32x32(float* row_ptr, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(v[8 * i + j]);
    st_global_256(row_ptr + 8 * i, w);
  }
}
// 32 fp32 values of this thread's row -> split-bf16: 64 B in the hi plane, 64 B in the lo plane
__device__ __forceinline__ void row_store_split32(__nv_bfloat16* row_ptr, long long plane, const float* v) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t wh[8], wl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { wh[j] = hi[8 * i + j]; wl[j] = lo[8 * i + j]; }
    st_global_256(row_ptr + 16 * i, wh);
    st_global_256(row_ptr + plane + 16 * i, wl);
  }
}
// 16 fp32 values of this thread's row -> one 64 B fp32 block
__device__ __forceinline__ void stage_store_f32(uint32_t sbuf, int lane, float* gbase, int stride, const int (&rp)[4], const float* v) {
  uint4 d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    d[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase), (long long)stride * 4, rp, d);
}
// 32 fp32 values of this thread's row -> split-bf16: one 64 B block in the hi plane, one in the lo plane
__device__ __forceinline__ void stage_store_split32(uint32_t sbuf, int lane, __nv_bfloat16* gbase, long long plane, int stride,
                                                    const int (&rp)[4], const float* v) {
  uint4 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[8 * i + 2 * j], b = v[8 * i + 2 * j + 1];
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
      const float2 hf = __bfloat1622float2(h2);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
      h[j] = *reinterpret_cast<const uint32_t*>(&h2);
      l[j] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi[i] = make_uint4(h[0], h[1], h[2], h[3]);
    lo[i] = make_uint4(l[0], l[1], l[2], l[3]);
  }
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase), (long long)stride * 2, rp, hi);
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase + plane), (long long)stride * 2, rp, lo);
}

__device__ __forceinline__ void store_f32x16(float* dst, const float* v, int nvalid) {
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}

__device__ __forceinline__ void store_split16(__nv_bfloat16* hi_dst, long long plane, const float* v, int nvalid) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(hi_dst) & 15) == 0) {
    uint4* dh = reinterpret_cast<uint4*>(hi_dst);
    uint4* dl = reinterpret_cast<uint4*>(hi_dst + plane);
    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  } else {
    unsigned short* dh = reinterpret_cast<unsigned short*>(hi_dst);
    unsigned short* dl = reinterpret_cast<unsigned short*>(hi_dst + plane);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) {
        dh[i] = (unsigned short)(hi[i >> 1] >> ((i & 1) * 16));
        dl[i] = (unsigned short)(lo[i >> 1] >> ((i & 1) * 16));
      }
  }
}

// PAIR is a template parameter because a kernel that contains cta_group::2 instructions can only be launched as a cluster of
// two (cudaErrorInvalidClusterSize otherwise): the single-CTA variants must not contain them.
template <int EPI, int ACT, int EW, bool PAIR>
__global__ void __launch_bounds__(64 + 32 * EW, EW == 4 ? 2 : 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
  // header: [0,256) barriers + TMEM pointer | [1024, ..) epilogue staging, 2560 B per epilogue warp | two 1 KB bias buffers ;
  // then the operand ring: `stages` x {A hi, A lo, W hi, W lo}
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 64, bar_tfull = smem_base + 128, bar_tempty = smem_base + 144,
                 bar_afull = smem_base + 256, bar_aempty = smem_base + 320, tmem_slot = smem_base + 192;
  const uint32_t stage0 = smem_base + 1024;
  const uint32_t bias0 = smem_base + 1024 + EW * 32 * TC_STAGE_ROW;
  const uint32_t tiles0 = smem_base + tc_header(EW);
  const uint32_t rowb = (uint32_t)p.bk * 2u;                              // bytes of one K-major operand row of a stage (128 or 64)
  const uint32_t ta_plane = 128u * rowb;                                  // one bf16 plane of the activation tile
  const uint32_t b_plane = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * rowb;     // weight rows held by this CTA
  // pair + stacked-N: region X = one whole weight plane (hi in the leader, lo in the peer: together the stacked [W_hi; W_lo]
  // operand of the N = 2*BN MMA), region Y = this CTA's half of W_hi (the B operand of the A_lo * W_hi MMA)
  const bool pstack = PAIR && p.stackn;
  const uint32_t stage_bytes = 2 * ta_plane + (pstack ? 3 * b_plane : 2 * b_plane);
  const uint32_t a_plane = (uint32_t)(p.PW * p.PH) * 128u;                        // halo mode: one bf16 plane of the halo tile
  const uint32_t a_stage = (2u * a_plane + 1023u) & ~1023u;
  const uint32_t bring0 = tiles0 + (uint32_t)p.a_stages * a_stage;                // halo mode: start of the weight ring

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();   // the next kernel's prologue may overlap this grid's tail (it waits before touching memory)
  auto stamp = [&](int slot) {
    if (p.dbg_times) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg_times[(long long)blockIdx.x * 8 + slot] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int tiles_per_split = p.num_tiles / p.ksplit;      // pixel tiles x N tiles

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (p.nseg > 1) prefetch_tmap(&tmA1);
    if (p.nseg > 2) prefetch_tmap(&tmA2);
    prefetch_tmap(&tmW);
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < (p.halo ? p.b_stages : p.stages); ++s) {
      mbar_init(bar_full + 8 * s, 1);
      // multicast mode: every CTA of the cluster releases the stage (peers write into it); pair mode: one multicast commit
      mbar_init(bar_empty + 8 * s, PAIR ? 1u : (uint32_t)p.cluster);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);     // MMA issuer -> epilogue: accumulator a complete
      mbar_init(bar_tempty + 8 * a, PAIR ? 2 * EW : EW);   // epilogue warps (of both CTAs of a pair) -> MMA issuer
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(tmem_slot, (uint32_t)p.tmem_cols); else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  if (p.cluster > 1) cluster_sync_all(); else __syncthreads();   // peers' barriers must exist before any multicast reaches them
  tc_fence_after();
  const int crank = p.cluster > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
  // tile walk: cluster c takes tile groups c, c + #clusters, ...; group g = tiles g*cluster .. +cluster-1 (same N tile)
  const int tile0 = (int)(blockIdx.x / p.cluster) * p.cluster + crank, tile_step = (int)gridDim.x;
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) stamp(1);
  griddep_wait();                // everything above overlapped the previous kernel; its results are visible from here on

  // Persistent CTA: tiles blockIdx.x, +gridDim.x, ...  The three roles walk the same tile sequence; the operand ring and
  // the two TMEM accumulators run on, so tile i+1's loads and MMAs overlap tile i's epilogue.
  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int stage = 0, hsa = 0;
      uint32_t phase = 0, hpa = 0;
      const int w_rows = p.BN / p.cluster;
      for (int t = tile0; t < p.num_tiles; t += tile_step) {
        const int ks = t / tiles_per_split, t2 = t - ks * tiles_per_split;
        const int nt = t2 / p.m_tiles, mt = t2 - nt * p.m_tiles;
        const int b = (mt / tiles_per_img) * p.TB, tr = mt % tiles_per_img;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * p.BN;
        const int tap_lo = ks * p.taps_per_split, tap_hi = tap_lo + p.taps_per_split < p.num_taps ? tap_lo + p.taps_per_split : p.num_taps;
        if (p.halo) {
          // one halo tile per channel chunk, then one weight tile per tap (own ring)
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
            for (int cc = 0; cc < p.seg_chunks[s]; ++cc) {
              mbar_wait(bar_aempty + 8 * hsa, hpa ^ 1u);
              if (p.dbg_epi & 8) mbar_arrive(bar_afull + 8 * hsa);     // timing experiment: no activation loads
              else {
                mbar_arrive_expect_tx(bar_afull + 8 * hsa, 2 * a_plane);
                tma_load_5d(tiles0 + hsa * a_stage, tm, bar_afull + 8 * hsa, cc * TC_BK, x0 - p.pw, y0 - p.ph, b, 0);
              }
              if (++hsa == p.a_stages) { hsa = 0; hpa ^= 1u; }
              const int wk = p.seg_wcoff[s] + cc * TC_BK;
              for (int tap = 0; tap < p.num_taps; ++tap) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t w_dst = bring0 + stage * 2 * b_plane;
                if (p.dbg_epi & 16) mbar_arrive(full);                 // timing experiment: no weight loads
                else {
                  mbar_arrive_expect_tx(full, 2 * b_plane);
                  tma_load_4d(w_dst, &tmW, full, wk, n0, tap, 0);
                  tma_load_4d(w_dst + b_plane, &tmW, full, wk, n0, tap, 1);
                }
                if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
              }
            }
          }
          continue;
        }
        for (int tap = tap_lo; tap < tap_hi; ++tap) {
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          const int cx = x0 * p.sx + kx - p.pw, cy = y0 * p.sy + ky - p.ph;
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
            for (int cc = 0; cc < p.seg_chunks[s]; ++cc) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              const uint32_t full = bar_full + 8 * stage;
              const uint32_t a_dst = tiles0 + stage * stage_bytes;
              const uint32_t w_dst = a_dst + 2 * ta_plane;
              const int wk = p.seg_wcoff[s] + cc * p.bk, wt = p.w_batched ? b : tap;
              if (PAIR) {
                // both CTAs' loads complete on the LEADER's barrier (it issues the pair's MMAs); each brings its A tile and
                // its half of the weight rows
                const uint32_t lfull = mapa_shared(full, 0);
                if (crank == 0) mbar_arrive_expect_tx(full, 2 * stage_bytes);
                tma_load_5d_2cta(a_dst, tm, lfull, cc * p.bk, cx, cy, b, 0);
                if (pstack) {
                  tma_load_4d_2cta(w_dst, &tmW, lfull, wk, n0, wt, crank);                          // X: plane `crank`, rows [0, BN/2)
                  tma_load_4d_2cta(w_dst + b_plane, &tmW, lfull, wk, n0 + w_rows, wt, crank);       //    ... rows [BN/2, BN)
                  tma_load_4d_2cta(w_dst + 2 * b_plane, &tmW, lfull, wk, n0 + crank * w_rows, wt, 0);   // Y: my half of W_hi
                } else {
                  tma_load_4d_2cta(w_dst, &tmW, lfull, wk, n0 + crank * w_rows, wt, 0);
                  tma_load_4d_2cta(w_dst + b_plane, &tmW, lfull, wk, n0 + crank * w_rows, wt, 1);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                continue;
              }
              mbar_arrive_expect_tx(full, stage_bytes);
              tma_load_5d(a_dst, tm, full, cc * p.bk, cx, cy, b, 0);
              if (p.cluster == 1) {
                tma_load_4d(w_dst, &tmW, full, wk, n0, wt, 0);
                tma_load_4d(w_dst + b_plane, &tmW, full, wk, n0, wt, 1);
              } else {       // this CTA's share of both weight planes, delivered to every CTA of the cluster
                const uint32_t off = (uint32_t)(crank * w_rows) * rowb;
                tma_load_4d_mc(w_dst + off, &tmW, full, wk, n0 + crank * w_rows, wt, 0, cmask);
                tma_load_4d_mc(w_dst + b_plane + off, &tmW, full, wk, n0 + crank * w_rows, wt, 1, cmask);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && !(PAIR && crank != 0)) {
      // ================= MMA issuer (pair mode: the leader CTA issues for both)
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, p.BN),
                     idesc2 = make_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, 2 * p.BN);
      int stage = 0, hsa = 0;
      uint32_t phase = 0, hpa = 0;
      int it = 0;
      for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_tempty + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);   // epilogue of tile it-2 has drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
        if (p.halo) {
          // tap (ky, kx) = the halo tile read from row offset ky*PW + kx; SWIZZLE_128B is a function of the absolute shared-
          // memory address for TMA and UMMA alike, so any 128 B row offset is a valid operand start
          const uint32_t a_sbo = (uint32_t)p.PW * 128u;
          bool first = true, init0 = true, init1 = true;
          for (int sg = 0; sg < p.nseg; ++sg) {
            for (int cc = 0; cc < p.seg_chunks[sg]; ++cc) {
              const int ks = (cc == p.seg_chunks[sg] - 1) ? p.seg_last_ks[sg] : TC_BK / 16;
              mbar_wait(bar_afull + 8 * hsa, hpa);
              const uint32_t a_base = tiles0 + hsa * a_stage;
              for (int tap = 0; tap < p.num_taps; ++tap) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (it == 0 && first) stamp(2);
                const int ky = tap / p.kw, kx = tap - ky * p.kw;
                const uint32_t a_addr = a_base + (uint32_t)(ky * p.PW + kx) * 128u;
                const uint64_t a_hi = make_smem_desc_sw128(a_addr, a_sbo), a_lo = make_smem_desc_sw128(a_addr + a_plane, a_sbo);
                const uint32_t w_addr = bring0 + stage * 2 * b_plane;
                const uint64_t b_hi = make_smem_desc_sw128(w_addr, 1024), b_lo = make_smem_desc_sw128(w_addr + b_plane, 1024);
                if (p.dbg_epi & 4) {
                  // timing experiment: no MMAs issued (barrier / load pipeline only)
                } else if (p.stackn) {
                  // MMAs of equal shape are issued back to back: measured (tools/trace_mma_rate.py) 61 ns per MMA when
                  // consecutive instructions share the instruction descriptor, 104 ns when the shape alternates
#pragma unroll
                  for (int k = 0; k < TC_BK / 16; ++k)
                    if (k < ks) umma_bf16(d_tmem, a_hi + (uint64_t)(k * 32 >> 4), b_hi + (uint64_t)(k * 32 >> 4), idesc2, (!first || k > 0) ? 1u : 0u);
#pragma unroll
                  for (int k = 0; k < TC_BK / 16; ++k)
                    if (k < ks) {
                      const uint64_t ko = (uint64_t)(k * 32 >> 4);
                      if (p.splitacc) umma_bf16(d_tmem + 2 * p.BN, a_lo + ko, b_hi + ko, idesc, (!first || k > 0) ? 1u : 0u);
                      else umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                    }
                } else if (p.dualacc) {
                  // two independent accumulation chains, alternating with every instruction
#pragma unroll
                  for (int kp = 0; kp < TC_BK / 16; kp += 2) {
                    const uint64_t k0 = (uint64_t)(kp * 32 >> 4), k1 = (uint64_t)((kp + 1) * 32 >> 4);
                    const bool on0 = kp < ks, on1 = kp + 1 < ks;
                    if (on0) umma_bf16(d_tmem, a_hi + k0, b_hi + k0, idesc, init0 ? 0u : 1u);
                    if (on1) umma_bf16(d_tmem + p.BN, a_hi + k1, b_hi + k1, idesc, init1 ? 0u : 1u);
                    if (on0) { umma_bf16(d_tmem, a_hi + k0, b_lo + k0, idesc, 1u); init0 = false; }
                    if (on1) { umma_bf16(d_tmem + p.BN, a_hi + k1, b_lo + k1, idesc, 1u); init1 = false; }
                    if (on0) umma_bf16(d_tmem, a_lo + k0, b_hi + k0, idesc, 1u);
                    if (on1) umma_bf16(d_tmem + p.BN, a_lo + k1, b_hi + k1, idesc, 1u);
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < TC_BK / 16; ++k) {
                    if (k < ks) {
                      const uint64_t ko = (uint64_t)(k * 32 >> 4);
                      umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, (!first || k > 0) ? 1u : 0u);
                      umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                      umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                    }
                  }
                }
                first = false;
                umma_commit(bar_empty + 8 * stage);
                if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
              }
              umma_commit(bar_aempty + 8 * hsa);      // every tap of this chunk has read the halo tile
              if (++hsa == p.a_stages) { hsa = 0; hpa ^= 1u; }
            }
          }
          umma_commit(bar_tfull + 8 * acc);
          if (it == 0) stamp(3);
          continue;
        }
        int c = 0;
        bool ninit0 = true, ninit1 = true;
        const int mks = t / tiles_per_split;
        const int mtap_lo = mks * p.taps_per_split, mtap_hi = mtap_lo + p.taps_per_split < p.num_taps ? mtap_lo + p.taps_per_split : p.num_taps;
        for (int tap = mtap_lo; tap < mtap_hi; ++tap) {
          for (int sg = 0; sg < p.nseg; ++sg) {
            for (int cc = 0; cc < p.seg_chunks[sg]; ++cc, ++c) {
              const int ks = (cc == p.seg_chunks[sg] - 1) ? p.seg_last_ks[sg] : p.bk / 16;   // skip all-zero k-steps of a ragged chunk
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              if (it == 0 && c == 0) stamp(2);
              const uint32_t a_addr = tiles0 + stage * stage_bytes;
              // 64-channel stages: SWIZZLE_128B rows; 32-channel stages (twice as many, same ring bytes): SWIZZLE_64B rows
              auto mk = [&](uint32_t addr) { return p.bk == 32 ? make_smem_desc_sw64(addr, 512) : make_smem_desc_sw128(addr, 1024); };
              const uint64_t a_hi = mk(a_addr), a_lo = mk(a_addr + ta_plane);
              const uint64_t b_hi = mk(a_addr + 2 * ta_plane);
              const uint64_t b_lo = mk(a_addr + 2 * ta_plane + b_plane);
              if (PAIR && pstack) {
                const uint64_t b_y = mk(a_addr + 2 * ta_plane + 2 * b_plane);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  if (k < ks) umma_bf16_2cta(d_tmem, a_hi + (uint64_t)(k * 32 >> 4), b_hi + (uint64_t)(k * 32 >> 4), idesc2, (c > 0 || k > 0) ? 1u : 0u);   // A_hi * [W_hi; W_lo]
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  if (k < ks) umma_bf16_2cta(d_tmem, a_lo + (uint64_t)(k * 32 >> 4), b_y + (uint64_t)(k * 32 >> 4), idesc, 1u);   // A_lo * W_hi
              } else if (PAIR) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  if (k < ks) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);
                    umma_bf16_2cta(d_tmem, a_hi + ko, b_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    umma_bf16_2cta(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                    umma_bf16_2cta(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                  }
                }
              } else if (p.stackn) {
                // equal-shape MMAs back to back (see the halo path): first all A_hi * [W_hi; W_lo] (N = 2*BN), then all A_lo * W_hi
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)    // 16 bf16 = 32 B along the swizzled 128 B row
                  if (k < ks) umma_bf16(d_tmem, a_hi + (uint64_t)(k * 32 >> 4), b_hi + (uint64_t)(k * 32 >> 4), idesc2, (c > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  if (k < ks) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);
                    if (p.splitacc) umma_bf16(d_tmem + 2 * p.BN, a_lo + ko, b_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    else umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                  }
              } else if (p.dualacc) {
#pragma unroll
                for (int kp = 0; kp < TC_BK / 16; kp += 2) {
                  const uint64_t k0 = (uint64_t)(kp * 32 >> 4), k1 = (uint64_t)((kp + 1) * 32 >> 4);
                  const bool on0 = kp < ks, on1 = kp + 1 < ks;
                  if (on0) umma_bf16(d_tmem, a_hi + k0, b_hi + k0, idesc, ninit0 ? 0u : 1u);
                  if (on1) umma_bf16(d_tmem + p.BN, a_hi + k1, b_hi + k1, idesc, ninit1 ? 0u : 1u);
                  if (on0) { umma_bf16(d_tmem, a_hi + k0, b_lo + k0, idesc, 1u); ninit0 = false; }
                  if (on1) { umma_bf16(d_tmem + p.BN, a_hi + k1, b_lo + k1, idesc, 1u); ninit1 = false; }
                  if (on0) umma_bf16(d_tmem, a_lo + k0, b_hi + k0, idesc, 1u);
                  if (on1) umma_bf16(d_tmem + p.BN, a_lo + k1, b_hi + k1, idesc, 1u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  if (k < ks) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);
                    umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
                    umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
                  }
                }
              }
              if (PAIR) umma_commit_2cta(bar_empty + 8 * stage);        // frees the slot in both CTAs of the pair
              else if (p.cluster == 1) umma_commit(bar_empty + 8 * stage);     // frees the smem slot once these MMAs have read it
              else umma_commit_mc(bar_empty + 8 * stage, cmask);          // ... in every CTA of the cluster
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        if (PAIR) umma_commit_2cta(bar_tfull + 8 * acc); else umma_commit(bar_tfull + 8 * acc);   // accumulator complete
        if (it == 0) stamp(3);
      }
    }
  } else {
    // ================= epilogue: warp w may touch TMEM lanes 32*(w%4)..+31 ; lane = pixel row of the tile.  With EW = 8 two
    // warps share a lane quarter and take alternate 32-column slabs.
    constexpr int NPAR = EW / 4;
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int bb = row / (p.TW * p.TH), rr = row - bb * (p.TW * p.TH);
    const int h = rr / p.TW, w = rr - h * p.TW;
    const int half = p.cout >> 1;
    const int ngroups = p.BN / 16;
    const uint32_t sbuf = stage0 + (uint32_t)(warp - 2) * 32 * TC_STAGE_ROW;
    int it = 0;
    for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
      const int acc = it & 1;
      const int ks = t / tiles_per_split, t2 = t - ks * tiles_per_split;
      const int nt = t2 / p.m_tiles, mt = t2 - nt * p.m_tiles;
      float* const out_f32_t = p.out_f32 ? p.out_f32 + (long long)ks * p.split_stride : nullptr;      // this split's partial map
      const int b = (mt / tiles_per_img) * p.TB, tr = mt % tiles_per_img;
      const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
      const int y = ty * p.TH + h, x = tx * p.TW + w, n0 = nt * p.BN;
      const bool valid = y < p.H && x < p.W && b + bb < p.B;
      const long long pix = ((long long)(b + bb) * p.H + y) * p.W + x;
      int rp[4];                         // pixel index of staging row it*8 + (lane>>2), or -1
#pragma unroll
      for (int i = 0; i < 4; ++i) rp[i] = __shfl_sync(0xffffffffu, valid ? (int)pix : -1, i * 8 + (lane >> 2));
      const uint32_t bias_s = bias0 + (uint32_t)acc * 1024u;
      // stage this tile's bias slice (global-load latency would otherwise sit inside every column slab); the named barrier
      // also orders it against the slowest warp still reading the buffer two tiles ago
      for (int i = threadIdx.x - 64; i < p.BN; i += 32 * EW) {
        const float bvl = (p.bias && n0 + i < p.cout) ? __ldg(p.bias + n0 + i) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4 * i), "f"(bvl) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      int g_begin = 0;
      if (p.fast_epi) {
        // ---- coalesced path: 32-column slabs, every global access staged through shared memory.  The epilogue's inputs
        // (GRU: context term, h, z; ACT: residual) are prefetched one 16-column block ahead into registers (XA..YC).
        const int nslab = (p.cout - n0 < p.BN ? p.cout - n0 : p.BN) / 32;    // full slabs only; the tail uses the plain path
        uint4 XA[4], XB[4], XC[4], YA[4], YB[4], YC[4];
        auto issue = [&](int nb16, uint4 (&A)[4], uint4 (&B)[4], uint4 (&C)[4]) {
          if (p.dbg_epi & 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) A[j] = B[j] = C[j] = make_uint4(0u, 0u, 0u, 0u);
            return;
          }
          if (EPI == SCF_EPI_ACT) {
            if (p.aux0) row_issue(p.aux0 + pix * p.aux0_stride + nb16, valid, B);
          } else {
            if (p.pre) row_issue(p.pre + pix * p.pre_stride + nb16, valid, A);
            if (EPI == SCF_EPI_GRU_ZR) {
              if (nb16 >= half) row_issue(p.aux0 + pix * p.aux0_stride + (nb16 - half), valid, B);
            } else {
              row_issue(p.aux0 + pix * p.aux0_stride + nb16, valid, B);
              row_issue(p.aux1 + pix * p.aux1_stride + nb16, valid, C);
            }
          }
        };
        // applies the epilogue math to 16 accumulator columns (bias already added) using a prefetched block
        auto consume = [&](int nb16, float* v16, const uint4 (&A)[4], const uint4 (&B)[4], const uint4 (&C)[4]) {
          float t0[16];
          if (EPI == SCF_EPI_ACT) {
            if (p.aux0) {                      // residual connection (encoder BasicBlock): added before the activation
              row_values(B, t0);
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] += t0[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v16[i] = act_ct<ACT>(v16[i]);
          } else {
            if (p.pre) {                       // loop-invariant context contribution (bias folded in), computed once per forward
              row_values(A, t0);
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] += t0[i];
            }
            if (EPI == SCF_EPI_GRU_ZR) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] = sigmoid_fast(v16[i]);
              if (nb16 >= half) {              // r gate: r * h feeds the q convolution
                row_values(B, t0);
#pragma unroll
                for (int i = 0; i < 16; ++i) v16[i] *= t0[i];
              }
            } else {                           // h' = (1 - z) h + z tanh(.)
              float t1[16];
              row_values(B, t0);
              row_values(C, t1);
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] = (1.f - t1[i]) * t0[i] + t1[i] * tanh_fast(v16[i]);
            }
          }
        };
        if (par < nslab) issue(n0 + par * 32, XA, XB, XC);       // overlaps the wait for the accumulator
        mbar_wait(bar_tfull + 8 * acc, ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        if (it == 0 && threadIdx.x == 64) stamp(4);
#pragma unroll 1
        for (int sl = par; sl < nslab; sl += NPAR) {
          float v[32];
          __syncwarp();
          tmem_ld32(t_addr + (uint32_t)(sl * 32), v);
          if (p.stackn || p.dualacc) {
            float v2[32];
            tmem_ld32(t_addr + (uint32_t)(p.BN + sl * 32), v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += v2[i];
            if (p.splitacc) {
              tmem_ld32(t_addr + (uint32_t)(2 * p.BN + sl * 32), v2);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += v2[i];
            }
          }
          const int nb = n0 + sl * 32;
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {           // bias from shared memory (same address in every lane: broadcast)
            float4 bq;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bq.x), "=f"(bq.y), "=f"(bq.z), "=f"(bq.w)
                         : "r"(bias_s + (uint32_t)(sl * 32 + i4 * 4) * 4));
            v[4 * i4] = fmaf(v[4 * i4], p.scale, bq.x); v[4 * i4 + 1] = fmaf(v[4 * i4 + 1], p.scale, bq.y);
            v[4 * i4 + 2] = fmaf(v[4 * i4 + 2], p.scale, bq.z); v[4 * i4 + 3] = fmaf(v[4 * i4 + 3], p.scale, bq.w);
          }
          issue(nb + 16, YA, YB, YC);
          consume(nb, v, XA, XB, XC);
          if (sl + NPAR < nslab) issue(nb + NPAR * 32, XA, XB, XC);
          consume(nb + 16, v + 16, YA, YB, YC);
          // ---- stores of the finished 32-column slab
          if (p.dbg_epi & 2) {
            float acc_ = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc_ += v[i];
            if (acc_ == 123.456f) p.out_f32[0] = acc_;
          } else if (EPI == SCF_EPI_GRU_ZR) {
            if (p.direct_st) {
              if (valid) {
                if (nb < half) row_store_f32x32(out_f32_t + pix * p.out_f32_stride + p.out_f32_coff + nb, v);
                else row_store_split32(p.out2_hl + pix * p.out2_hl_stride + (nb - half), p.out2_hl_plane, v);
              }
            } else if (nb < half) {            // z gate -> fp32 (read back by the q convolution's epilogue)
              stage_store_f32(sbuf, lane, out_f32_t + p.out_f32_coff + nb, p.out_f32_stride, rp, v);
              stage_store_f32(sbuf, lane, out_f32_t + p.out_f32_coff + nb + 16, p.out_f32_stride, rp, v + 16);
            } else {
              stage_store_split32(sbuf, lane, p.out2_hl + (nb - half), p.out2_hl_plane, p.out2_hl_stride, rp, v);
            }
          } else {
            if (p.out_f32) {
              if (EPI == SCF_EPI_ACT && p.stats) {
                // InstanceNorm statistics of the map being written: per-column partial sums over this warp's 32 rows
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  uint4 d[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    d[i] = make_uint4(__float_as_uint(v[hh * 16 + 4 * i]), __float_as_uint(v[hh * 16 + 4 * i + 1]),
                                      __float_as_uint(v[hh * 16 + 4 * i + 2]), __float_as_uint(v[hh * 16 + 4 * i + 3]));
                  float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
                  stage_store64_stats(sbuf, lane, reinterpret_cast<char*>(out_f32_t + p.out_f32_coff + nb + hh * 16),
                                      (long long)p.out_f32_stride * 4, rp, d, s4, q4);
#pragma unroll
                  for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      s4[j] += __shfl_xor_sync(0xffffffffu, s4[j], off);
                      q4[j] += __shfl_xor_sync(0xffffffffu, q4[j], off);
                    }
                  }
                  if (lane < 4) {
                    float* o = p.stats + ((long long)(mt * 4 + q) * 2) * p.cout + nb + hh * 16 + lane * 4;
                    *reinterpret_cast<float4*>(o) = make_float4(s4[0], s4[1], s4[2], s4[3]);
                    *reinterpret_cast<float4*>(o + p.cout) = make_float4(q4[0], q4[1], q4[2], q4[3]);
                  }
                }
              } else if (p.direct_st) {
                if (valid) row_store_f32x32(out_f32_t + pix * p.out_f32_stride + p.out_f32_coff + nb, v);
              } else {
                stage_store_f32(sbuf, lane, out_f32_t + p.out_f32_coff + nb, p.out_f32_stride, rp, v);
                stage_store_f32(sbuf, lane, out_f32_t + p.out_f32_coff + nb + 16, p.out_f32_stride, rp, v + 16);
              }
            }
            if (p.out_hl) {
              if (p.direct_st) {
                if (valid) row_store_split32(p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb, p.out_hl_plane, v);
              } else {
                stage_store_split32(sbuf, lane, p.out_hl + p.out_hl_coff + nb, p.out_hl_plane, p.out_hl_stride, rp, v);
              }
            }
          }
        }
        g_begin = nslab * 2;
      } else {
        mbar_wait(bar_tfull + 8 * acc, ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        if (it == 0 && threadIdx.x == 64) stamp(4);
      }
      // ---- plain path (ragged channel tails, unaligned outputs): 16-column groups, direct per-thread stores
#pragma unroll 1
      for (int g = g_begin + par; g < ngroups; g += NPAR) {
        float v[16];
        __syncwarp();
        tmem_ld16(t_addr + (uint32_t)(g * 16), v);
        if (p.stackn || p.dualacc) {           // second half of the stacked accumulator (A_hi * W_lo) / second accumulation chain
          float v2[16];
          tmem_ld16(t_addr + (uint32_t)(p.BN + g * 16), v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += v2[i];
          if (p.splitacc) {
            tmem_ld16(t_addr + (uint32_t)(2 * p.BN + g * 16), v2);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += v2[i];
          }
        }
        const int nb = n0 + g * 16;
        if (!valid || nb >= p.cout) continue;
        const int nvalid = p.cout - nb < 16 ? p.cout - nb : 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float bq;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bq) : "r"(bias_s + (uint32_t)(g * 16 + i) * 4));
          v[i] = fmaf(v[i], p.scale, bq);
        }
        if (EPI == SCF_EPI_ACT) {
          if (p.aux0) {                        // residual connection (encoder BasicBlock): added before the activation
            float rv[16];
            load16(p.aux0 + pix * p.aux0_stride + nb, rv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rv[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = act_ct<ACT>(v[i]);
          if (p.out_f32) store_f32x16(out_f32_t + pix * p.out_f32_stride + p.out_f32_coff + nb, v, nvalid);
          if (p.out_hl) store_split16(p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb, p.out_hl_plane, v, nvalid);
        } else if (EPI == SCF_EPI_GRU_ZR) {    // cout = 2*Ch, Ch % 16 == 0: a group is entirely z or entirely r
          if (p.pre) {
            float pv[16];
            load16(p.pre + pix * p.pre_stride + nb, pv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += pv[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = sigmoid_fast(v[i]);
          if (nb < half) {                     // z gate -> fp32 (read back by the q convolution's epilogue)
            store_f32x16(out_f32_t + pix * p.out_f32_stride + p.out_f32_coff + nb, v, 16);
          } else {                             // r gate -> r*h as split-bf16, the q convolution's first input segment
            float hv[16];
            load16(p.aux0 + pix * p.aux0_stride + (nb - half), hv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= hv[i];
            store_split16(p.out2_hl + pix * p.out2_hl_stride + (nb - half), p.out2_hl_plane, v, 16);
          }
        } else {                               // SCF_EPI_GRU_Q: h' = (1-z) h + z tanh(.)   (cout % 16 == 0)
          float hv[16], zv[16];
          if (p.pre) {
            load16(p.pre + pix * p.pre_stride + nb, hv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += hv[i];
          }
          load16(p.aux0 + pix * p.aux0_stride + nb, hv);
          load16(p.aux1 + pix * p.aux1_stride + nb, zv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (1.f - zv[i]) * hv[i] + zv[i] * tanh_fast(v[i]);
          if (p.out_f32) store_f32x16(out_f32_t + pix * p.out_f32_stride + p.out_f32_coff + nb, v, 16);
          if (p.out_hl) store_split16(p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb, p.out_hl_plane, v, 16);
        }
      }
      // this warp has read everything it needs from accumulator `acc`: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_shared(bar_tempty + 8 * acc, 0));   // the leader's issuer waits for both
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      if (it == 0 && threadIdx.x == 64) stamp(5);
    }
  }
  tc_fence_before();
  if (p.cluster > 1) cluster_sync_all(); else __syncthreads();   // no CTA may exit while peers can still signal or write into it
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (threadIdx.x == 32) stamp(6);
}

// ================================================================== transposed tile (layers of <= 128 output channels)
// Measured (tools/trace_mma_rate.py): one tcgen05.mma of M = 128 costs ~61 ns for every N <= 128 and ~88 ns for N = 256, so a
// pixels-as-rows tile of a 64/96/128-channel layer pays the full instruction time for a fraction of the tensor core's width.
// Here the roles are swapped: the (zero-padded) 128 weight rows are the MMA's M side and 256 pixels its N side - twice the
// pixels per instruction at 1.44x the instruction time.  Accumulator lane = output channel, accumulator column = pixel, so
// in the epilogue a warp's 32 lanes hold 32 consecutive channels of ONE pixel: every NHWC access is naturally coalesced, the
// bias is a per-thread scalar and the InstanceNorm partial sums need no cross-lane reduction.
struct TctParams {
  int nseg, seg_chunks[3], seg_wcoff[3], seg_last_ks[3];
  int B, H, W, kh, kw, ph, pw, sx, sy;      // H, W: output size
  int tw_sh, th_sh;                         // tile = 2^tw_sh x 2^th_sh pixels x (256 >> (tw_sh + th_sh)) samples
  int tiles_x, tiles_y, num_tiles, num_taps, stages, cout;
  const float* bias; float scale;
  float* out_f32; int out_f32_stride, out_f32_coff;
  __nv_bfloat16* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff;
  const float* aux0; int aux0_stride;
  int gru_zr;                               // GRU gates (cout = 256): tile t = (pixel tile t >> 1, gate t & 1); gate 0: z = sigmoid(acc + pre)
                                            // -> out_f32, gate 1: r = sigmoid(acc + pre[128 + c]) -> r * h (aux0) -> out_hl
  int gru_q;                                // GRU state update epilogue: h' = (1 - z) h + z tanh(acc + pre), aux0 = h, aux1 = z
  const float* aux1; int aux1_stride; const float* pre; int pre_stride;
  float* stats;                             // [num_tiles][2 warp parities][2][cout]
  // halo mode (stride-1 multi-tap layers, one sample per tile): the tile is 8 x 32 pixels, its activations are fetched ONCE per
  // channel chunk together with the (kw-1) x (kh-1) halo - PW x PH pixel rows - and every tap reads them through a row-shifted
  // UMMA descriptor (one 8-row core-matrix group = one tile row, group stride = PW rows); weights run through their own ring
  int halo, PW, PH, p_stages, w_stages;
  int cw_sh;                                // log2 of the pixels per row of an epilogue chunk (16 columns = 1 x 16 or 2 x 8 pixels)
  int bk;                                   // channels per stage: 64 (SWIZZLE_128B rows, 2 x 96 KB stages) or 32 (SWIZZLE_64B rows,
                                            // 4 x 48 KB stages: same bytes in the ring, finer hand-over between TMA and MMA)
  int dbg;                                  // timing experiments (SCFLOW_TCT_DBG): 1 no epilogue global accesses, 2 no MMAs, 4 no
                                            // activation loads, 8 no weight loads
};
constexpr int TCT_PIX = 256;
constexpr uint32_t TCT_W_PLANE = 128 * 128, TCT_P_PLANE = TCT_PIX * 128;
constexpr uint32_t TCT_STAGE = 2 * TCT_W_PLANE + 2 * TCT_P_PLANE;      // 96 KB
constexpr int TCT_EW = 8;
constexpr uint32_t TCT_STAGING = 4096;     // per epilogue warp: [16 px][32 ch] fp32 | [2 planes][16 px][32 ch] bf16

template <int ACT>
__global__ void __launch_bounds__(64 + 32 * TCT_EW, 1)
conv_tct_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1,
                const __grid_constant__ CUtensorMap tmP2, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmOF, const __grid_constant__ CUtensorMap tmOH, const TctParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 64, bar_tfull = smem_base + 128, bar_tempty = smem_base + 144,
                 tmem_slot = smem_base + 192, bar_pfull = smem_base + 256, bar_pempty = smem_base + 320;
  const uint32_t staging0 = smem_base + 1024;                         // 4 KB per epilogue warp (TMA store source)
  const uint32_t tiles0 = staging0 + TCT_EW * TCT_STAGING;
  const uint32_t row_bytes = (uint32_t)p.bk * 2u;                     // one K-major operand row of a stage
  const uint32_t w_plane = 128u * row_bytes, p_plane = (uint32_t)TCT_PIX * row_bytes, stage_bytes = 2u * (w_plane + p_plane);
  const uint32_t hp_plane = (uint32_t)(p.PW * p.PH) * row_bytes;      // halo mode: one bf16 plane of the halo tile
  const uint32_t hp_stage = (2u * hp_plane + 1023u) & ~1023u;
  const uint32_t wring0 = tiles0 + (uint32_t)p.p_stages * hp_stage;   // halo mode: weight ring behind the activation ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP0);
    if (p.nseg > 1) prefetch_tmap(&tmP1);
    if (p.nseg > 2) prefetch_tmap(&tmP2);
    prefetch_tmap(&tmW);
    if (p.out_f32) prefetch_tmap(&tmOF);
    if (p.out_hl) prefetch_tmap(&tmOH);
    for (int s = 0; s < (p.halo ? p.w_stages : p.stages); ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < p.p_stages; ++s) {
      mbar_init(bar_pfull + 8 * s, 1);
      mbar_init(bar_pempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, TCT_EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);       // two accumulators of 256 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  const int TW = 1 << p.tw_sh, TH = 1 << p.th_sh, TB = TCT_PIX >> (p.tw_sh + p.th_sh);
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int stage = 0, hs = 0;
      uint32_t phase = 0, hph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int pt = p.gru_zr ? (t >> 1) : t, wrow = p.gru_zr ? (t & 1) * 128 : 0;      // pixel tile, first weight row
        const int b = (pt / tiles_per_img) * TB, tr = pt % tiles_per_img;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int x0 = tx * TW, y0 = ty * TH;
        if (p.halo) {
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmP0 : (s == 1 ? &tmP1 : &tmP2);
            for (int cc = 0; cc < p.seg_chunks[s]; ++cc) {
              mbar_wait(bar_pempty + 8 * hs, hph ^ 1u);
              if (p.dbg & 4) mbar_arrive(bar_pfull + 8 * hs);
              else {
                mbar_arrive_expect_tx(bar_pfull + 8 * hs, 2 * hp_plane);
                tma_load_5d(tiles0 + hs * hp_stage, tm, bar_pfull + 8 * hs, cc * p.bk, x0 - p.pw, y0 - p.ph, b, 0);
              }
              if (++hs == p.p_stages) { hs = 0; hph ^= 1u; }
              const int wk = p.seg_wcoff[s] + cc * p.bk;
              for (int tap = 0; tap < p.num_taps; ++tap) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t w_dst = wring0 + stage * 2 * w_plane;
                if (p.dbg & 8) mbar_arrive(full);
                else {
                  mbar_arrive_expect_tx(full, 2 * w_plane);
                  tma_load_4d(w_dst, &tmW, full, wk, wrow, tap, 0);
                  tma_load_4d(w_dst + w_plane, &tmW, full, wk, wrow, tap, 1);
                }
                if (++stage == p.w_stages) { stage = 0; phase ^= 1u; }
              }
            }
          }
          continue;
        }
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          const int cx = x0 * p.sx + kx - p.pw, cy = y0 * p.sy + ky - p.ph;
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmP0 : (s == 1 ? &tmP1 : &tmP2);
            for (int cc = 0; cc < p.seg_chunks[s]; ++cc) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              const uint32_t full = bar_full + 8 * stage;
              const uint32_t w_dst = tiles0 + stage * stage_bytes, p_dst = w_dst + 2 * w_plane;
              const int wk = p.seg_wcoff[s] + cc * p.bk;
              const uint32_t txb = ((p.dbg & 4) ? 0u : 2 * p_plane) + ((p.dbg & 8) ? 0u : 2 * w_plane);
              if (txb) mbar_arrive_expect_tx(full, txb); else mbar_arrive(full);
              if (!(p.dbg & 4)) tma_load_5d(p_dst, tm, full, cc * p.bk, cx, cy, b, 0);
              if (!(p.dbg & 8)) {
                tma_load_4d(w_dst, &tmW, full, wk, wrow, tap, 0);
                tma_load_4d(w_dst + w_plane, &tmW, full, wk, wrow, tap, 1);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer: D[channel][pixel] += W[channel][k] * P[pixel][k]  (split-bf16: hi*hi + hi*lo + lo*hi)
      const uint32_t idesc = make_idesc_bf16(128, TCT_PIX);
      int stage = 0, it = 0, hs = 0;
      uint32_t phase = 0, hph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_tempty + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TCT_PIX);
        if (p.halo) {
          // tap (ky, kx) = the halo tile read from pixel-row offset ky*PW + kx (the swizzle is a function of the absolute
          // shared-memory address for TMA and UMMA alike, so any row offset is a valid operand start)
          const uint32_t p_sbo = (uint32_t)p.PW * row_bytes;
          bool first = true;
          for (int sg = 0; sg < p.nseg; ++sg) {
            for (int cc = 0; cc < p.seg_chunks[sg]; ++cc) {
              const int ks = (cc == p.seg_chunks[sg] - 1) ? p.seg_last_ks[sg] : p.bk / 16;
              mbar_wait(bar_pfull + 8 * hs, hph);
              const uint32_t p_base = tiles0 + hs * hp_stage;
              for (int tap = 0; tap < p.num_taps; ++tap) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const int ky = tap / p.kw, kx = tap - ky * p.kw;
                const uint32_t p_addr = p_base + (uint32_t)(ky * p.PW + kx) * row_bytes;
                const uint32_t w_addr = wring0 + stage * 2 * w_plane;
                const uint64_t w_hi = make_smem_desc_sw64(w_addr, 512), w_lo = make_smem_desc_sw64(w_addr + w_plane, 512);
                const uint64_t p_hi = make_smem_desc_sw64(p_addr, p_sbo), p_lo = make_smem_desc_sw64(p_addr + hp_plane, p_sbo);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  if (k < ks && !(p.dbg & 2)) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);
                    umma_bf16(d_tmem, w_hi + ko, p_hi + ko, idesc, (!first || k > 0) ? 1u : 0u);
                    umma_bf16(d_tmem, w_hi + ko, p_lo + ko, idesc, 1u);
                    umma_bf16(d_tmem, w_lo + ko, p_hi + ko, idesc, 1u);
                  }
                }
                first = false;
                umma_commit(bar_empty + 8 * stage);
                if (++stage == p.w_stages) { stage = 0; phase ^= 1u; }
              }
              umma_commit(bar_pempty + 8 * hs);        // every tap of this chunk has read the halo tile
              if (++hs == p.p_stages) { hs = 0; hph ^= 1u; }
            }
          }
          umma_commit(bar_tfull + 8 * acc);
          continue;
        }
        int c = 0;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          for (int sg = 0; sg < p.nseg; ++sg) {
            for (int cc = 0; cc < p.seg_chunks[sg]; ++cc, ++c) {
              const int ks = (cc == p.seg_chunks[sg] - 1) ? p.seg_last_ks[sg] : p.bk / 16;
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              const uint32_t w_addr = tiles0 + stage * stage_bytes, p_addr = w_addr + 2 * w_plane;
              const bool sw64 = p.bk == 32;
              const uint64_t w_hi = sw64 ? make_smem_desc_sw64(w_addr, 512) : make_smem_desc_sw128(w_addr, 1024);
              const uint64_t w_lo = sw64 ? make_smem_desc_sw64(w_addr + w_plane, 512) : make_smem_desc_sw128(w_addr + w_plane, 1024);
              const uint64_t p_hi = sw64 ? make_smem_desc_sw64(p_addr, 512) : make_smem_desc_sw128(p_addr, 1024);
              const uint64_t p_lo = sw64 ? make_smem_desc_sw64(p_addr + p_plane, 512) : make_smem_desc_sw128(p_addr + p_plane, 1024);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                if (k < ks && !(p.dbg & 2)) {
                  const uint64_t ko = (uint64_t)(k * 32 >> 4);
                  umma_bf16(d_tmem, w_hi + ko, p_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
                  umma_bf16(d_tmem, w_hi + ko, p_lo + ko, idesc, 1u);
                  umma_bf16(d_tmem, w_lo + ko, p_hi + ko, idesc, 1u);
                }
              }
              umma_commit(bar_empty + 8 * stage);
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        umma_commit(bar_tfull + 8 * acc);
      }
    }
  } else {
    // ================= epilogue: thread = output channel (TMEM lane), accumulator columns = the tile's pixels; the two warps
    // of a lane quarter take alternate 32-pixel slabs
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int c = q * 32 + lane;
    const bool cvalid = p.gru_zr || c < p.cout;
    const bool warp_active = p.gru_zr || q * 32 < p.cout;
    float bias_c = (p.bias && cvalid) ? __ldg(p.bias + c) : 0.f;
    const int tw_mask = TW - 1, th_mask = TH - 1, b_sh = p.tw_sh + p.th_sh, odd = lane & 1, cwm = (1 << p.cw_sh) - 1;
    const uint32_t stg = staging0 + (uint32_t)(warp - 2) * TCT_STAGING;
    const bool both = p.out_f32 && p.out_hl && !p.gru_zr;
    uint32_t kc = 0;                      // chunks staged so far (staging block parity)
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int pt = p.gru_zr ? (t >> 1) : t, gate = p.gru_zr ? (t & 1) : 0;
      const bool do_f32 = p.gru_zr ? gate == 0 : p.out_f32 != nullptr, do_hl = p.gru_zr ? gate == 1 : p.out_hl != nullptr;
      if (p.gru_zr && p.bias) bias_c = __ldg(p.bias + gate * 128 + c);
      const int b = (pt / tiles_per_img) * TB, tr = pt % tiles_per_img;
      const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
      const int x0 = tx * TW, y0 = ty * TH;
      // GRU q: h, z and the context term of a chunk are requested one chunk ahead (the first one before the accumulator is
      // complete), so their latency hides behind the main loop's tail / the previous chunk's stores
      float nh[16], nz[16], npre[16];
      auto gru_issue = [&](int ch) {
        const int col0 = ch * 16;
        const int xb = x0 + (col0 & tw_mask), y = y0 + ((col0 >> p.tw_sh) & th_mask), bi = b + (col0 >> b_sh);
        const int nvx = p.W - xb, nvy = bi < p.B ? p.H - y : 0;
        const long long pix0 = ((long long)bi * p.H + y) * p.W + xb;
        const float* hp = p.aux0 + pix0 * p.aux0_stride + c;
        const float* zp = p.aux1 + pix0 * p.aux1_stride + c;
        const float* pp = p.pre + pix0 * p.pre_stride + gate * 128 + c;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool ok = (j & cwm) < nvx && (j >> p.cw_sh) < nvy && cvalid;
          const long long off = (long long)((j >> p.cw_sh) * p.W + (j & cwm));
          nh[j] = (ok && (p.gru_q || gate == 1)) ? __ldg(hp + off * p.aux0_stride) : 0.f;
          nz[j] = (ok && p.gru_q) ? __ldg(zp + off * p.aux1_stride) : 0.f;
          npre[j] = (ok && p.pre) ? __ldg(pp + off * p.pre_stride) : 0.f;
        }
      };
      if (((ACT == SCF_ACT_TANH && p.gru_q) || (ACT == SCF_ACT_SIGMOID && p.gru_zr)) && warp_active) gru_issue(par);
      mbar_wait(bar_tfull + 8 * acc, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      float ssum = 0.f, qsum = 0.f;
      if (warp_active && !(p.dbg & 1)) {
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TCT_PIX);
#pragma unroll 1
        for (int ch = par; ch < TCT_PIX / 16; ch += TCT_EW / 4) {
          float v[16];
          if (!((ACT == SCF_ACT_TANH && p.gru_q) || (ACT == SCF_ACT_SIGMOID && p.gru_zr))) {
            __syncwarp();
            tmem_ld16(t_addr + (uint32_t)(ch * 16), v);
          }
          // A 16-column chunk is a run of 16 pixels of one tile row (tiles >= 32 wide) or two rows of an 8-wide halo-mode
          // tile.  The warp writes its
          // [16 pixels][32 channels] block to shared memory in NHWC order (conflict-free: lanes = consecutive channels, constant
          // offsets) and one TMA store per output moves it out; the tensor map clips pixels beyond the row / image / batch and
          // channels beyond cout, so the stores need no predicates.  (A first version with per-thread global stores spent
          // ~70 instructions per pixel pair on 64-bit addressing and predicates and was issue-bound at 1.6 TB/s.)
          const int col0 = ch * 16;
          const int xb = x0 + (col0 & tw_mask), y = y0 + ((col0 >> p.tw_sh) & th_mask), bi = b + (col0 >> b_sh);
          const int nvx = p.W - xb, nvy = bi < p.B ? p.H - y : 0;                // valid pixels per chunk row / valid chunk rows
          auto okj = [&](int j) { return (j & cwm) < nvx && (j >> p.cw_sh) < nvy; };
          auto offj = [&](int j) { return (long long)((j >> p.cw_sh) * p.W + (j & cwm)); };   // pixel offset inside the chunk
          if (ACT == SCF_ACT_TANH && p.gru_q) {
            // SepConvGRU state update (raft_decoder.py:235-253)
            float hv[16], zv[16], pv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { hv[j] = nh[j]; zv[j] = nz[j]; pv[j] = npre[j]; }
            if (ch + TCT_EW / 4 < TCT_PIX / 16) gru_issue(ch + TCT_EW / 4);
            __syncwarp();
            tmem_ld16(t_addr + (uint32_t)(ch * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (1.f - zv[j]) * hv[j] + zv[j] * tanh_fast(fmaf(v[j], p.scale, bias_c) + pv[j]);
          } else if (ACT == SCF_ACT_SIGMOID && p.gru_zr) {
            // SepConvGRU gates: z = sigmoid(.) ; r = sigmoid(.), stored as r * h, the q convolution's first input segment
            float hv[16], pv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { hv[j] = nh[j]; pv[j] = npre[j]; }
            if (ch + TCT_EW / 4 < TCT_PIX / 16) gru_issue(ch + TCT_EW / 4);
            __syncwarp();
            tmem_ld16(t_addr + (uint32_t)(ch * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float g = sigmoid_fast(fmaf(v[j], p.scale, bias_c) + pv[j]);
              v[j] = gate ? g * hv[j] : g;
            }
          } else if (p.aux0 && cvalid) {                                           // residual, added before the activation
            const float* ax = p.aux0 + (((long long)bi * p.H + y) * p.W + xb) * p.aux0_stride + c;
            float rv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) rv[j] = okj(j) ? __ldg(ax + offj(j) * p.aux0_stride) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_ct<ACT>(fmaf(v[j], p.scale, bias_c) + rv[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_ct<ACT>(fmaf(v[j], p.scale, bias_c));
          }
          if (p.stats) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (okj(j)) { ssum += v[j]; qsum = fmaf(v[j], v[j], qsum); }
          }
          // Staging: 4 KB per warp.  One output kind: two 2 KB blocks used alternately; both kinds: the fp32 block and the
          // bf16 block each go out as their own bulk group.  Either way a block is rewritten only after the group that read it
          // two groups ago has completed (wait_group.read 1), so one store is always in flight behind the shared-memory writes.
          const uint32_t blk_f = stg + ((!both && (kc & 1)) ? 2048u : 0u);
          const uint32_t blk_h = stg + ((both || (kc & 1)) ? 2048u : 0u);
          ++kc;
          if (lane == 0) bulk_wait_group_read1();
          __syncwarp();
          if (do_f32) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(blk_f + (uint32_t)(j * 128 + lane * 4)), "f"(v[j]) : "memory");
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmOF, blk_f, q * 32, xb, y, bi);
              bulk_commit_group();
              if (both) bulk_wait_group_read1();
            }
            __syncwarp();
          }
          if (do_hl) {
            // lanes 2i, 2i+1 trade values: the even lane stores channels (c, c+1) of pixel j as one 32-bit word per plane, the
            // odd lane channels (c-1, c) of pixel j + 1
            const uint32_t hbase = blk_h + (uint32_t)(odd * 64 + (lane & ~1) * 2);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const __nv_bfloat16 h0 = __float2bfloat16_rn(v[j]), h1 = __float2bfloat16_rn(v[j + 1]);
              const __nv_bfloat16 l0 = __float2bfloat16_rn(v[j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[j + 1] - __bfloat162float(h1));
              const uint32_t hl0 = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(l0) << 16);
              const uint32_t hl1 = (uint32_t)__bfloat16_as_ushort(h1) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? hl0 : hl1, 1);
              const uint32_t mine = odd ? hl1 : hl0;
              const uint32_t lo_ch = odd ? recv : mine, hi_ch = odd ? mine : recv;      // lower / upper channel of the pair
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(hbase + (uint32_t)(j * 64)), "r"((lo_ch & 0xffffu) | (hi_ch << 16)) : "memory");
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(hbase + 1024u + (uint32_t)(j * 64)), "r"((lo_ch >> 16) | (hi_ch & 0xffff0000u)) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_5d(&tmOH, blk_h, q * 32, xb, y, bi, 0);
              bulk_commit_group();
            }
          }
        }
      }
      if (p.stats && cvalid) {
        float* o = p.stats + ((long long)(t * (TCT_EW / 4) + par) * 2) * p.cout + c;
        o[0] = ssum;
        o[p.cout] = qsum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
    if (lane == 0) bulk_wait_group0();       // shared memory must outlive the last TMA stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ------------------------------------------------------------------ prep kernels
// input channels [i_begin, i_begin + I) of a weight with I_total input channels land at packed channels i_dst ..
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, int O, int I, int taps,
                                      int cin_pad, int cout_pad, int o_off, int I_total, int i_begin, int i_dst) {
  const long long total = (long long)O * I * taps;
  const long long plane = (long long)taps * cout_pad * cin_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const long long r = idx / I;
    const int o = (int)(r % O);
    const int tap = (int)(r / O);
    __nv_bfloat16 hi, lo;
    split_bf16(w[((long long)o * I_total + i_begin + i) * taps + tap], hi, lo);
    const long long dst = ((long long)tap * cout_pad + o_off + o) * cin_pad + i_dst + i;
    packed[dst] = hi;
    packed[plane + dst] = lo;
  }
}

__global__ void nchw_to_nhwc_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long plane,
                                          int dst_stride, int dst_coff, float* __restrict__ dst_f32, int f32_stride, int C,
                                          int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    if (c < C && pp < HW) tile[i][threadIdx.x] = src[((long long)b * C + c) * HW + pp];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (c < C && pp < HW) {
      const float v = tile[threadIdx.x][i];
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      const long long o = ((long long)b * HW + pp) * dst_stride + dst_coff + c;
      dst[o] = hi;
      dst[plane + o] = lo;
      if (dst_f32) dst_f32[((long long)b * HW + pp) * f32_stride + c] = v;
    }
  }
}

__global__ void split_copy_kernel(const float* __restrict__ src, int src_stride, int src_coff, __nv_bfloat16* __restrict__ dst,
                                  long long plane, int dst_stride, int dst_coff, long long npix, int nch) {
  scf_pdl_enter();
  const long long total = npix * nch;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long px = idx / nch;
    const int c = (int)(idx - px * nch);
    __nv_bfloat16 hi, lo;
    split_bf16(src[px * src_stride + src_coff + c], hi, lo);
    dst[px * dst_stride + dst_coff + c] = hi;
    dst[plane + px * dst_stride + dst_coff + c] = lo;
  }
}

// ------------------------------------------------------------------ thin-input convolutions on the tensor cores
// A k x k convolution over very few input channels (RGB stem 7x7x3, flow encoders 7x7x2) has K = k*k*Cin <= 147: too
// thin per tap for a 64-channel k-chunk.  The kernel row is therefore folded into the channel axis first:
//   T[n, y, xo, kx*Cin + c] = in[n, c, y, xo*sx + kx - k/2]      (zero outside the image; channels padded to 16 or 32:
//   a power-of-two pixel pitch - measured: the 48 B pitch of 24 channels makes the TMA tile loads 2.3x slower)
// written as split-bf16 NHWC, after which the layer is a k x 1 convolution (stride sy vertically) over KW*Cin channels
// that conv_tc_kernel runs with one ragged k-chunk per tap.
template <int CIN, int KW, bool NCHW>
__global__ void __launch_bounds__(256) im2col_x_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long plane,
                                                             int N, int H, int Wi, int Wo, int sx, long long total) {
  scf_pdl_enter();
  constexpr int KC = KW * CIN <= 16 ? 16 : 32;
  static_assert(KW * CIN <= 32, "folded row must fit 32 channels");
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const long long r = idx / Wo;
    const int y = (int)(r % H);
    const long long n = r / H;
    __nv_bfloat16 hi[KC], lo[KC];
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      const int ix = xo * sx + kx - KW / 2;
      const bool ok = ix >= 0 && ix < Wi;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        float v = 0.f;
        if (ok) v = NCHW ? __ldg(in + ((n * CIN + c) * H + y) * Wi + ix) : __ldg(in + ((n * H + y) * Wi + ix) * CIN + c);
        split_bf16(v, hi[kx * CIN + c], lo[kx * CIN + c]);
      }
    }
#pragma unroll
    for (int i = KW * CIN; i < KC; ++i) { hi[i] = __float2bfloat16_rn(0.f); lo[i] = hi[i]; }
    uint4* oh = reinterpret_cast<uint4*>(out + idx * KC);
    uint4* ol = reinterpret_cast<uint4*>(out + plane + idx * KC);
#pragma unroll
    for (int i = 0; i < KC / 8; ++i) {
      oh[i] = *reinterpret_cast<const uint4*>(hi + 8 * i);
      ol[i] = *reinterpret_cast<const uint4*>(lo + 8 * i);
    }
  }
}

// OIHW [O, C, KH, KW] -> [O, KW*C, KH, 1] with channel kx*C + c (the weight of the folded convolution above)
__global__ void fold_kx_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int C, int KH, int KW) {
  const int total = O * C * KH * KW;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int kx = idx % KW;
    int r = idx / KW;
    const int ky = r % KH; r /= KH;
    const int c = r % C;
    const int o = r / C;
    out[((long long)o * (KW * C) + kx * C + c) * KH + ky] = w[idx];
  }
}

// in: NCHW fp32 [N,cin,H,Wi] (nchw) or NHWC fp32 [N,H,Wi,cin]; out: split-bf16 [2][N,H,Wo,KC], KC = 16 (kw*cin <= 16) or 32,
// Wo = (Wi-1)/sx + 1
int im2col_x_split(const float* in, int nchw, int cin, int kw, void* out_hl, long long plane, int N, int H, int Wi, int sx,
                   cudaStream_t st) {
  SCF_REQUIRE(in && out_hl && N > 0 && H > 0 && Wi > 0 && (sx == 1 || sx == 2), SCF_ERR_ARG, "im2col_x_split: bad args");
  const int Wo = (Wi - 1) / sx + 1;
  const long long total = (long long)N * H * Wo;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_hl);
  if (cin == 3 && kw == 7 && nchw) launch_pdl(im2col_x_split_kernel<3, 7, true>, dim3(blocks), dim3(256), 0, st, in, o, plane, N, H, Wi, Wo, sx, total);
  else if (cin == 2 && kw == 7 && !nchw) launch_pdl(im2col_x_split_kernel<2, 7, false>, dim3(blocks), dim3(256), 0, st, in, o, plane, N, H, Wi, Wo, sx, total);
  else SCF_REQUIRE(false, SCF_ERR_UNSUPPORTED, "im2col_x_split: unsupported (cin %d, kw %d, nchw %d)", cin, kw, nchw);
  return check_launch("im2col_x_split_kernel");
}

// packs an OIHW [O,C,KH,KW] weight for the folded k x 1 convolution: bf16 [2][KH][cout_pad][cin_pad], cin_pad >= pad8(KW*C).
// `scratch` needs O*C*KH*KW floats.
int pack_conv_weight_tc_foldx(const float* w_oihw, float* scratch, void* packed, int O, int C, int KH, int KW, int cin_pad,
                              int cout_pad, int o_off, cudaStream_t st);

// ------------------------------------------------------------------ host side
int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr,
               CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  SCF_REQUIRE(fn != nullptr, SCF_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  elem_strides ? elem_strides : ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCF_REQUIRE(r == CUDA_SUCCESS, SCF_ERR_ARG, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu %llu %llu, box %u %u %u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1],
              box[2]);
  return 0;
}

static void pick_tile(int B, int H, int W, bool one_sample, int& TW, int& TH, int& TB) {
  // TW*TH*TB = 128 rows; minimise the padded volume, prefer wide tiles (longer contiguous runs), then tall ones
  long long best = -1;
  for (int tw = 128; tw >= 2; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      const int tb = 128 / (tw * th);
      if (one_sample && tb != 1) continue;
      const long long vol = (long long)cdiv(W, tw) * tw * cdiv(H, th) * th * cdiv(B, tb) * tb;
      if (best < 0 || vol < best) { best = vol; TW = tw; TH = th; TB = tb; }
    }
  }
}

// tiling of the calling thread's most recent scf_conv2d_tc launch (the `stats` buffer is indexed by its pixel tiles)
thread_local int g_last_m_tiles = 0, g_last_tiles_per_img = 0;
extern thread_local int g_last_stat_rows_per_img;
// per_img: pixel tiles per sample (0 when a tile spans samples); stat_rows: rows of [2][cout] partial sums per sample in `stats`
void conv2d_tc_last_tiles(int* m_tiles, int* per_img, int* stat_rows) {
  if (m_tiles) *m_tiles = g_last_m_tiles;
  if (per_img) *per_img = g_last_tiles_per_img;
  if (stat_rows) *stat_rows = g_last_stat_rows_per_img;
}
// upper bound of the number of pixel tiles any tiling of a [B, Hout, Wout] output can have (sizes the `stats` buffer)
static bool tct_pick_tile(int B, int H, int W, int sx, int sy, bool one_sample, int& tw_sh, int& th_sh);
int conv2d_tc_max_tiles(int B, int Hout, int Wout) {
  int TW = 0, TH = 0, TB = 0;
  pick_tile(B, Hout, Wout, false, TW, TH, TB);
  const int a = cdiv(Wout, TW) * cdiv(Hout, TH) * cdiv(B, TB), h = cdiv(Wout, 8) * cdiv(Hout, 16) * B;
  int best = a > h ? a : h;
  // transposed tiling: 2 rows of partial sums per 256-pixel tile (its 8 x 32 halo form never has more rows than the 8 x 16 one)
  int tw_sh = 0, th_sh = 0;
  if (tct_pick_tile(B, Hout, Wout, 1, 1, true, tw_sh, th_sh)) {
    const int t = cdiv(Wout, 1 << tw_sh) * cdiv(Hout, 1 << th_sh) * B;
    const int eq = (2 * t + 3) / 4;
    if (eq > best) best = eq;
  }
  // rolling-rows kernel (scf_conv_rows.cu): 4 rows per (output row, 128-pixel column strip)
  const int rr = Wout >= 96 ? B * Hout * cdiv(Wout, 128) : 0;      // (the kernel only takes maps at least 96 pixels wide)
  if (rr > best) best = rr;
  return best;
}

// ---- transposed-tile variant (conv_tct_kernel): plain-activation layers of <= 128 output channels
static bool tct_pick_tile(int B, int H, int W, int sx, int sy, bool one_sample, int& tw_sh, int& th_sh) {
  long long best = -1;
  for (int a = 8; a >= 5; --a) {          // >= 32 pixels wide: the epilogue's 32-column slabs must not cross tile rows
    for (int b = 8 - a; b >= 0; --b) {
      const int tw = 1 << a, th = 1 << b, tb = TCT_PIX >> (a + b);
      if (one_sample && tb != 1) continue;
      if ((tw > 1 ? tw * sx : 1) > 256 || (th > 1 ? th * sy : 1) > 256) continue;      // TMA box limit
      const long long vol = (long long)cdiv(W, tw) * tw * cdiv(H, th) * th * cdiv(B, tb) * tb;
      if (best < 0 || vol < best) { best = vol; tw_sh = a; th_sh = b; }
    }
  }
  return best >= 0;
}

thread_local int g_last_stat_rows_per_img = 0;

static bool tct_eligible(const scf_tc_conv_desc& d) {
  // SCFLOW_TC_T: 0 never, 1 (default) where it measured faster than the pixels-as-rows tiling (tools/bench_tct.py: every
  // layer except the 64-channel ones with a short reduction, whose main loop is no longer than the 256-pixel epilogue),
  // 2 every eligible layer
  const char* e = getenv("SCFLOW_TC_T");
  const int mode = e ? atoi(e) : 1;
  if (!mode) return false;
  if (mode == 1) {
    // fewer than ~100 tiles of 256 pixels would leave a third of the SMs idle where the 128-pixel tiling still fills them
    // (measured: config 4's B = 16 shard 9.3 -> 9.7 ms without this rule)
    {
      const int sx = d.stride_x ? d.stride_x : (d.stride == 2 ? 2 : 1), sy = d.stride_y ? d.stride_y : (d.stride == 2 ? 2 : 1);
      const long long wo = (d.W + 2 * (d.kw / 2) - d.kw) / sx + 1, ho = (d.H + 2 * (d.kh / 2) - d.kh) / sy + 1;
      if ((long long)d.B * ho * wo < 100LL * TCT_PIX) return false;
    }
    int cin = 0;
    for (int s = 0; s < d.nseg; ++s) cin += d.seg[s].nch;
    if (d.cout <= 64 && cin * d.kh * d.kw < 1024) {
      // SCFLOW_TCT_SMALL64=1: take them anyway when the halo form applies on a large map (measured alone 289 us against
      // 302 us on the 64-channel 128 x 128 encoder layers, but +0.05 ms on the whole step: off)
      const char* he = getenv("SCFLOW_TCT_HALO");
      const bool halo = (he ? atoi(he) != 0 : true) && d.kh * d.kw > 1 && d.stride != 2 && d.stride_x != 2 && d.stride_y != 2;
      const char* s64 = getenv("SCFLOW_TCT_SMALL64");
      if (!(halo && (s64 ? atoi(s64) != 0 : false) && (long long)d.B * d.H * d.W >= 4LL * 148 * 256)) return false;
    }
  }
  if (d.epi == SCF_EPI_GRU_ZR) {
    // the GRU's gate convolution: two 128-channel weight-row tiles (z, r) per pixel tile
    const char* ze = getenv("SCFLOW_TC_T_GRUZR");
    const char* e = getenv("SCFLOW_TC_T");
    if ((e && atoi(e) == 0) || !(ze ? atoi(ze) != 0 : false)) return false;
    if (d.w_batched || d.cout != 256 || d.cout_pad != 256 || d.act != SCF_ACT_SIGMOID || !d.aux0 || !d.out_f32 || !d.out2_hl || d.stats)
      return false;
    if (d.out2_hl_stride % 8 || d.out2_hl_plane % 8 || reinterpret_cast<uintptr_t>(d.out2_hl) % 16) return false;
    if (d.out_f32_stride % 4 || d.out_f32_coff % 4 || reinterpret_cast<uintptr_t>(d.out_f32) % 16) return false;
    const int sx = d.stride_x ? d.stride_x : (d.stride == 2 ? 2 : 1), sy = d.stride_y ? d.stride_y : (d.stride == 2 ? 2 : 1);
    const long long wo = (d.W + 2 * (d.kw / 2) - d.kw) / sx + 1, ho = (d.H + 2 * (d.kh / 2) - d.kh) / sy + 1;
    return wo >= 24 && (long long)d.B * ho * wo >= 100LL * TCT_PIX;
  }
  if (d.w_batched || d.cout_pad > 128 || d.act < SCF_ACT_NONE || d.act > SCF_ACT_TANH) return false;
  if (d.epi == SCF_EPI_GRU_Q) {
    const char* qe = getenv("SCFLOW_TC_T_GRUQ");
    if (!(qe ? atoi(qe) != 0 : true) || d.act != SCF_ACT_TANH || !d.aux0 || !d.aux1 || d.stats) return false;
  } else if (d.epi == SCF_EPI_GRU_ZR) {
    return false;      // handled by tct_eligible_zr (two 128-channel gates)
  } else if (d.epi != SCF_EPI_ACT || d.pre) {
    return false;
  }
  {
    const int sx = d.stride_x ? d.stride_x : (d.stride == 2 ? 2 : 1);
    if ((d.W + 2 * (d.kw / 2) - d.kw) / sx + 1 < 24) return false;      // tiles are at least 32 pixels wide
  }
  // the epilogue leaves through TMA stores: 16 B aligned bases and pitches
  // ... and they clip the channel axis at 16 B granularity
  if (!d.out_pad_writable && ((d.out_hl && d.cout % 8) || (d.out_f32 && d.cout % 4))) return false;
  if (d.out_pad_writable && ((d.cout + 7) / 8 * 8 > d.cout_pad)) return false;
  if (d.out_hl && (d.cout % 2 || d.out_hl_stride % 8 || d.out_hl_coff % 8 || d.out_hl_plane % 8 || reinterpret_cast<uintptr_t>(d.out_hl) % 16))
    return false;
  if (d.out_f32 && (d.out_f32_stride % 4 || d.out_f32_coff % 4 || reinterpret_cast<uintptr_t>(d.out_f32) % 16)) return false;
  return true;
}

static int conv2d_tct(const scf_tc_conv_desc& d, cudaStream_t st) {
  TctParams p = {};
  p.nseg = d.nseg;
  const int stride = d.stride == 2 ? 2 : 1;
  p.sx = d.stride_x ? d.stride_x : stride;
  p.sy = d.stride_y ? d.stride_y : stride;
  p.kh = d.kh; p.kw = d.kw; p.ph = d.kh / 2; p.pw = d.kw / 2;
  p.B = d.B; p.H = (d.H + 2 * p.ph - d.kh) / p.sy + 1; p.W = (d.W + 2 * p.pw - d.kw) / p.sx + 1;
  SCF_REQUIRE(tct_pick_tile(d.B, p.H, p.W, p.sx, p.sy, d.stats != nullptr, p.tw_sh, p.th_sh), SCF_ERR_UNSUPPORTED,
              "scf_conv2d_tc: no transposed tiling");
  {
    const char* be = getenv("SCFLOW_TCT_BK");
    p.bk = (be && atoi(be) == 64) ? 64 : 32;      // measured: four 48 KB stages beat two 96 KB stages by 3 % of the whole step
  }
  p.stages = p.bk == 32 ? 4 : 2;
  {
    // halo mode: 8 x 32 pixel tiles (one sample each), activations once per channel chunk, weights per tap in their own ring
    const char* he = getenv("SCFLOW_TCT_HALO");
    const bool want = he ? atoi(he) != 0 : true;       // measured: +1 % of the step (encoder layers 2-15 % each)
    // (1-D taps - the GRU's 1x5 / 5x1 - measured slower with the halo form: 58 us against 52 us)
    const char* h1 = getenv("SCFLOW_TCT_HALO_1D");
    const bool two_d = (d.kh > 1 && d.kw > 1) || (h1 && atoi(h1) != 0);
    if (want && p.bk == 32 && p.sx == 1 && p.sy == 1 && d.kh * d.kw > 1 && two_d) {
      const int PW = 8 + d.kw - 1, PH = 32 + d.kh - 1;
      const int hp_stage = (2 * PW * PH * 64 + 1023) / 1024 * 1024;
      int ws = (2 * (int)TCT_STAGE - 2 * hp_stage) / (2 * 128 * 64);
      if (ws > 8) ws = 8;
      if (ws >= 3) {
        p.halo = 1; p.PW = PW; p.PH = PH; p.p_stages = 2; p.w_stages = ws;
        p.tw_sh = 3; p.th_sh = 5;
      }
    }
  }
  p.cw_sh = p.tw_sh < 4 ? p.tw_sh : 4;
  const int TW = 1 << p.tw_sh, TH = 1 << p.th_sh, TB = TCT_PIX >> (p.tw_sh + p.th_sh);
  p.tiles_x = cdiv(p.W, TW); p.tiles_y = cdiv(p.H, TH);
  p.num_tiles = p.tiles_x * p.tiles_y * cdiv(d.B, TB);
  p.num_taps = d.kh * d.kw;
  p.cout = d.cout;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TctParams);
  static const KernelFn table[4] = {conv_tct_kernel<SCF_ACT_NONE>, conv_tct_kernel<SCF_ACT_RELU>, conv_tct_kernel<SCF_ACT_SIGMOID>,
                                    conv_tct_kernel<SCF_ACT_TANH>};
  const int smem = 1024 + 1024 + TCT_EW * (int)TCT_STAGING + 2 * (int)TCT_STAGE;     // ring = 192 KB for both stage sizes
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [smem] {
    for (int i = 0; i < 4; ++i) {
      cudaError_t e = cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) attr_err = e;
    }
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(conv_tct_kernel): %s", cudaGetErrorString(attr_err));
  p.bias = d.bias; p.scale = d.scale;
  p.out_f32 = d.out_f32; p.out_f32_stride = d.out_f32_stride; p.out_f32_coff = d.out_f32_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.out_hl_plane = d.out_hl_plane; p.out_hl_stride = d.out_hl_stride;
  p.out_hl_coff = d.out_hl_coff;
  p.aux0 = d.aux0; p.aux0_stride = d.aux0_stride;
  p.gru_q = d.epi == SCF_EPI_GRU_Q ? 1 : 0;
  p.gru_zr = d.epi == SCF_EPI_GRU_ZR ? 1 : 0;
  if (p.gru_zr) {                      // z -> out_f32 (128 channels), r * h -> out2_hl (128 channels); tile = (pixel tile, gate)
    p.cout = 128;
    p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out2_hl); p.out_hl_plane = d.out2_hl_plane; p.out_hl_stride = d.out2_hl_stride;
    p.out_hl_coff = 0;
    p.num_tiles *= 2;
  }
  p.aux1 = d.aux1; p.aux1_stride = d.aux1_stride; p.pre = d.pre; p.pre_stride = d.pre_stride;
  p.stats = d.stats;
  { const char* de = getenv("SCFLOW_TCT_DBG"); p.dbg = de ? atoi(de) : 0; }
  if (d.stats) SCF_REQUIRE(TB == 1, SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: stats need one sample per tile");
  CUtensorMap tmP[3], tmW;
  int wcoff = 0;
  for (int s = 0; s < 3; ++s) {
    if (s >= d.nseg) { tmP[s] = tmP[0]; continue; }
    const scf_tc_seg& sg = d.seg[s];
    SCF_REQUIRE(sg.ptr && sg.nch > 0 && sg.nch % 8 == 0 && sg.coff % 8 == 0 && sg.stride % 8 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: segment %d channels/offset/stride must be multiples of 8", s);
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(sg.ptr) + sg.coff;
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(base) % 16 == 0 && (sg.plane_stride * 2) % 16 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: segment %d must be 16B aligned", s);
    cuuint64_t dims[5] = {(cuuint64_t)sg.nch, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {(cuuint64_t)sg.stride * 2, (cuuint64_t)d.W * sg.stride * 2, (cuuint64_t)d.H * d.W * sg.stride * 2,
                         (cuuint64_t)sg.plane_stride * 2};
    const int bsx = TW == 1 ? 1 : p.sx, bsy = TH == 1 ? 1 : p.sy;
    cuuint32_t box[5] = {(cuuint32_t)p.bk, (cuuint32_t)(TW * bsx), (cuuint32_t)(TH * bsy), (cuuint32_t)TB, 2};
    if (p.halo) { box[1] = (cuuint32_t)p.PW; box[2] = (cuuint32_t)p.PH; }
    cuuint32_t estr[5] = {1, (cuuint32_t)bsx, (cuuint32_t)bsy, 1, 1};
    SCF_TRY(encode_map(&tmP[s], base, 5, dims, str, box, estr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                       p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
    p.seg_chunks[s] = cdiv(sg.nch, p.bk);
    p.seg_last_ks[s] = cdiv(sg.nch - (p.seg_chunks[s] - 1) * p.bk, 16);
    p.seg_wcoff[s] = wcoff;
    wcoff += sg.nch;
  }
  SCF_REQUIRE(wcoff <= d.cin_pad, SCF_ERR_ARG, "scf_conv2d_tc: segments carry %d channels but the packed weight has cin_pad %d", wcoff, d.cin_pad);
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.cin_pad, (cuuint64_t)d.cout_pad, (cuuint64_t)p.num_taps, 2};
    cuuint64_t str[3] = {(cuuint64_t)d.cin_pad * 2, (cuuint64_t)d.cout_pad * d.cin_pad * 2,
                         d.w_plane_stride > 0 ? (cuuint64_t)d.w_plane_stride * 2 : (cuuint64_t)p.num_taps * d.cout_pad * d.cin_pad * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.bk, 128, 1, 1};       // rows >= cout_pad: zero fill
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: packed weight must be 16B aligned");
    SCF_TRY(encode_map(&tmW, d.w, 4, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                       p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
  }
  // output maps of the epilogue's TMA stores: box = 32 channels x 16 pixels of one row (both bf16 planes in one store)
  CUtensorMap tmOF = tmW, tmOH = tmW;
  if (d.out_f32) {
    cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)d.B};
    cuuint64_t str[3] = {(cuuint64_t)d.out_f32_stride * 4, (cuuint64_t)p.W * d.out_f32_stride * 4, (cuuint64_t)p.H * p.W * d.out_f32_stride * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(1 << p.cw_sh), (cuuint32_t)(16 >> p.cw_sh), 1};
    SCF_TRY(encode_map(&tmOF, d.out_f32 + d.out_f32_coff, 4, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  if (p.out_hl) {
    cuuint64_t dims[5] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {(cuuint64_t)p.out_hl_stride * 2, (cuuint64_t)p.W * p.out_hl_stride * 2, (cuuint64_t)p.H * p.W * p.out_hl_stride * 2,
                         (cuuint64_t)p.out_hl_plane * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)(1 << p.cw_sh), (cuuint32_t)(16 >> p.cw_sh), 1, 2};
    SCF_TRY(encode_map(&tmOH, p.out_hl + p.out_hl_coff, 5, dims, str, box, nullptr,
                       CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  g_last_m_tiles = p.num_tiles;
  g_last_tiles_per_img = TB == 1 ? p.tiles_x * p.tiles_y : 0;
  g_last_stat_rows_per_img = g_last_tiles_per_img * (TCT_EW / 4);
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_tiles < num_sms ? p.num_tiles : num_sms); cfg.blockDim = dim3(64 + 32 * TCT_EW);
  cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, table[d.act], tmP[0], tmP[1], tmP[2], tmW, tmOF, tmOH, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("conv_tct_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("conv_tct_kernel");
}

bool conv2d_rows_eligible(const scf_tc_conv_desc& d);
int conv2d_rows(const scf_tc_conv_desc& d, cudaStream_t st);
bool conv2d_stem_rows_eligible(const scf_tc_conv_desc& d);
int conv2d_stem_rows(const scf_tc_conv_desc& d, cudaStream_t st);

int conv2d_tc(const scf_tc_conv_desc& d, cudaStream_t st) {
  SCF_REQUIRE(d.nseg >= 1 && d.nseg <= 3, SCF_ERR_ARG, "scf_conv2d_tc: nseg must be 1..3");
  SCF_REQUIRE(d.w && d.B > 0 && d.H > 0 && d.W > 0 && d.cout > 0, SCF_ERR_ARG, "scf_conv2d_tc: null pointer or empty shape");
  SCF_REQUIRE(d.kh >= 1 && d.kw >= 1 && d.kh % 2 == 1 && d.kw % 2 == 1, SCF_ERR_ARG, "scf_conv2d_tc: odd kernel sizes only");
  SCF_REQUIRE(d.cin_pad % 8 == 0 && d.cout_pad % 16 == 0 && d.cout <= d.cout_pad, SCF_ERR_ARG,
              "scf_conv2d_tc: cin_pad %% 8, cout_pad %% 16 required");
  SCF_REQUIRE(!d.w_batched || (d.kh == 1 && d.kw == 1), SCF_ERR_ARG, "scf_conv2d_tc: batched weights need a 1x1 kernel");
  SCF_REQUIRE(d.out_f32 || d.out_hl || d.epi == SCF_EPI_GRU_ZR, SCF_ERR_ARG, "scf_conv2d_tc: no output given");
  if (d.epi == SCF_EPI_GRU_ZR)
    SCF_REQUIRE(d.out_f32 && d.aux0 && d.out2_hl && d.cout % 32 == 0, SCF_ERR_ARG, "scf_conv2d_tc: GRU_ZR needs out_f32 (z), aux0 (h), out2_hl (r*h)");
  if (d.epi == SCF_EPI_GRU_Q)
    SCF_REQUIRE(d.aux0 && d.aux1 && d.cout % 16 == 0, SCF_ERR_ARG, "scf_conv2d_tc: GRU_Q needs aux0 (h), aux1 (z), cout %% 16 == 0");
  if (d.epi == SCF_EPI_ACT && d.aux0) SCF_REQUIRE(d.cout % 16 == 0, SCF_ERR_ARG, "scf_conv2d_tc: residual needs cout %% 16 == 0");
  if (d.epi != SCF_EPI_ACT || d.aux0)
    SCF_REQUIRE(d.aux0_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d.aux0) % 16 == 0 &&
                    (d.epi != SCF_EPI_GRU_Q || (d.aux1_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d.aux1) % 16 == 0)),
                SCF_ERR_ALIGN, "scf_conv2d_tc: aux buffers must be 16B aligned with strides %% 4 == 0");
  SCF_REQUIRE(!d.bias || reinterpret_cast<uintptr_t>(d.bias) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: bias must be 16B aligned");
  if (d.out_hl)
    SCF_REQUIRE((d.out_hl_plane * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(d.out_hl) % 16 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: out_hl and its plane stride must be 16B aligned");
  if (d.epi != SCF_EPI_ACT && d.out_f32)
    SCF_REQUIRE(d.out_f32_stride % 4 == 0 && d.out_f32_coff % 4 == 0 && reinterpret_cast<uintptr_t>(d.out_f32) % 16 == 0,
                SCF_ERR_ALIGN, "scf_conv2d_tc: GRU epilogues need 16B-aligned fp32 outputs");

  SCF_REQUIRE(d.stride == 0 || d.stride == 1 || d.stride == 2, SCF_ERR_ARG, "scf_conv2d_tc: stride must be 1 or 2");
  SCF_REQUIRE(d.stride_x >= 0 && d.stride_x <= 2 && d.stride_y >= 0 && d.stride_y <= 2, SCF_ERR_ARG,
              "scf_conv2d_tc: stride_x / stride_y must be 0 (= stride), 1 or 2");
  if (d.stats)
    SCF_REQUIRE(d.cout % 32 == 0 && d.out_f32 && d.epi == SCF_EPI_ACT && reinterpret_cast<uintptr_t>(d.stats) % 16 == 0,
                SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: stats need an fp32 output, cout %% 32 == 0 and the plain epilogue");
  if (conv2d_rows_eligible(d)) return conv2d_rows(d, st);
  if (conv2d_stem_rows_eligible(d)) return conv2d_stem_rows(d, st);
  SCF_REQUIRE(d.aux0_hl == nullptr, SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: a split-bf16 residual (aux0_hl) is served by the rolling-rows kernel only");
  if (d.ksplit <= 1 && tct_eligible(d)) return conv2d_tct(d, st);
  TcParams p = {};
  p.nseg = d.nseg;
  const int stride = d.stride == 2 ? 2 : 1;
  p.sx = d.stride_x ? d.stride_x : stride;
  p.sy = d.stride_y ? d.stride_y : stride;
  p.kh = d.kh; p.kw = d.kw; p.ph = d.kh / 2; p.pw = d.kw / 2;
  p.B = d.B; p.H = (d.H + 2 * p.ph - d.kh) / p.sy + 1; p.W = (d.W + 2 * p.pw - d.kw) / p.sx + 1;
  pick_tile(d.B, p.H, p.W, d.w_batched != 0, p.TW, p.TH, p.TB);
  p.tiles_x = cdiv(p.W, p.TW); p.tiles_y = cdiv(p.H, p.TH);
  p.BN = d.cout_pad <= 256 ? d.cout_pad : 256;
  p.cout = d.cout;
  p.num_taps = d.kh * d.kw;
  p.w_batched = d.w_batched;
  p.m_tiles = p.tiles_x * p.tiles_y * cdiv(d.B, p.TB);
  p.num_tiles = p.m_tiles * cdiv(d.cout_pad, p.BN);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  {
    const char* be = getenv("SCFLOW_TC_BK");
    p.bk = (be && atoi(be) == 32) ? 32 : 64;
  }
  const int rowb = p.bk * 2, max_stages = p.bk == 32 ? 7 : TC_MAX_STAGES;
  int stage_bytes = 2 * 128 * rowb + 2 * p.BN * rowb;
  // epilogue warps: 8 (two per TMEM lane quarter) for the one-CTA-per-SM configuration - the epilogue's global loads
  // (GRU gates, residuals) need the extra memory-level parallelism - and 4 for the small-tile two-CTAs-per-SM mode
  int ew = 8;
  {
    const char* ev = getenv("SCFLOW_TC_EW");
    if (ev && atoi(ev) == 4) ew = 4;
  }
  p.stages = (232448 - 1024 - tc_header(ew)) / stage_bytes;
  if (p.stages > max_stages) p.stages = max_stages;
  SCF_REQUIRE(p.stages >= 2, SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: tile does not fit in shared memory");
  {
    const char* sv = getenv("SCFLOW_TC_STACKN");
    // off by default: measured without any loads (tools/trace_mma_rate.py) an MMA of N <= 128 takes 61 ns when the three
    // products are three equal-shape instructions, but ~100 ns each in the stacked two-instruction form
    p.stackn = (p.BN <= 128 && (sv ? atoi(sv) != 0 : false)) ? 1 : 0;
  }
  {
    const char* sa = getenv("SCFLOW_TC_SPLITACC");
    p.splitacc = (p.stackn && p.BN <= 64 && sa && atoi(sa) != 0) ? 1 : 0;     // 3*BN columns x 2 buffers must fit 512
  }
  {
    const char* da = getenv("SCFLOW_TC_DUALACC");
    p.dualacc = (!p.stackn && p.BN <= 128 && (da ? atoi(da) != 0 : false)) ? 1 : 0;
  }
  p.acc_cols = p.stackn ? (p.splitacc ? 3 : 2) * p.BN : (p.dualacc ? 2 * p.BN : p.BN);
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.acc_cols) p.tmem_cols <<= 1;
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
#define SCF_TC_ROW(EW_, PAIR_)                                                                                        \
  {conv_tc_kernel<SCF_EPI_GRU_ZR, SCF_ACT_SIGMOID, EW_, PAIR_>, conv_tc_kernel<SCF_EPI_GRU_Q, SCF_ACT_TANH, EW_, PAIR_>,  \
   conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_NONE, EW_, PAIR_>, conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_RELU, EW_, PAIR_>,         \
   conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_SIGMOID, EW_, PAIR_>, conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_TANH, EW_, PAIR_>}
  static const KernelFn table[3][6] = {SCF_TC_ROW(4, false), SCF_TC_ROW(8, false), SCF_TC_ROW(8, true)};   // [variant][epilogue]
#undef SCF_TC_ROW
  int ki = -1;
  if (d.epi == SCF_EPI_GRU_ZR) ki = 0;
  else if (d.epi == SCF_EPI_GRU_Q) ki = 1;
  else if (d.epi == SCF_EPI_ACT && d.act >= SCF_ACT_NONE && d.act <= SCF_ACT_TANH) ki = 2 + d.act;
  SCF_REQUIRE(ki >= 0, SCF_ERR_ARG, "scf_conv2d_tc: bad epilogue / activation");
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < 6; ++i) {
        cudaError_t e = cudaFuncSetAttribute(table[a][i], cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) attr_err = e;
      }
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(conv_tc_kernel): %s", cudaGetErrorString(attr_err));
  int ctas_per_sm = 1;
  {
    // Small tiles: keep two CTAs resident per SM (<= 113 KB of shared memory and <= 256 TMEM columns each) so that the two
    // CTAs' main loops interleave; larger tiles run one persistent CTA per SM with double-buffered accumulators.
    const char* ov = getenv("SCFLOW_TC_OCC2");
    const bool occ2 = ov ? atoi(ov) != 0 : true;
    // (the decision is made on the 64-channel stage size so that the stage granularity does not change the mode)
    if (occ2 && 2 * (stage_bytes * 128 / rowb) + 1024 + tc_header(4) <= 115712 && p.tmem_cols <= 256 && p.num_tiles >= 4 * num_sms) {
      ew = 4;
      p.stages = (115712 - 1024 - tc_header(4)) / stage_bytes;
      if (p.stages > max_stages) p.stages = max_stages;
      ctas_per_sm = 2;
    }
  }
  // ---- CTA pair: two CTAs (one TPC) run one cta_group::2 MMA of M = 256; each stages its own A tile and half of the weight
  // rows, so the shared-memory traffic per FLOP (operand reads by the tensor core + TMA fills) drops by a third for
  // BN = 256.  Used for every large-tile layer with an even number of pixel tiles.
  p.pair = 0;
  if (ctas_per_sm == 1) {
    const char* pe = getenv("SCFLOW_TC_PAIR");
    const bool want_pair = pe ? atoi(pe) != 0 : true;
    const char* pm = getenv("SCFLOW_TC_PAIR_MIN_BN");
    const int pair_min_bn = pm ? atoi(pm) : 64;
    if (want_pair && p.m_tiles % 2 == 0 && p.BN % 16 == 0 && p.BN >= pair_min_bn && p.num_tiles >= 4 &&
        (!d.w_batched || (p.tiles_x * p.tiles_y) % 2 == 0)) {
      p.pair = 1;
      ew = 8;
      // BN > 128: three MMAs of N = BN, each CTA holds BN/2 rows of both weight planes.  BN <= 128 (stacked-N): the leader
      // holds W_hi, the peer W_lo (= the two halves of the stacked [W_hi; W_lo] operand) plus half of W_hi each
      stage_bytes = 2 * 128 * rowb + (p.stackn ? 3 : 2) * (p.BN / 2) * rowb;
      p.stages = (232448 - 1024 - tc_header(ew)) / stage_bytes;
      if (p.stages > max_stages) p.stages = max_stages;
      p.splitacc = 0; p.dualacc = 0;
      p.acc_cols = p.stackn ? 2 * p.BN : p.BN;
      p.tmem_cols = 32;
      while (p.tmem_cols < 2 * p.acc_cols) p.tmem_cols <<= 1;
    }
  }
  // ---- halo mode: stride-1 multi-tap convolutions fetch the activation tile once per channel chunk (see TcParams)
  p.halo = 0; p.PW = p.PH = 0; p.a_stages = 0; p.b_stages = 0;
  int smem = 1024 + tc_header(ew) + p.stages * stage_bytes;
  {
    const char* hv = getenv("SCFLOW_TC_HALO");
    const int want = hv ? atoi(hv) : 1;               // 0 off, 1 on for single-CTA one-CTA-per-SM tiles, 2 also instead of two-CTAs-per-SM
    const bool eligible = p.sx == 1 && p.sy == 1 && p.num_taps > 1 && !d.w_batched && !p.pair && p.H * p.W >= 128 &&
                          (ctas_per_sm == 1 || want >= 2);
    if (want && eligible) {
      const int PW = 8 + d.kw - 1, PH = 16 + d.kh - 1;
      const int a_stage = (2 * PW * PH * 128 + 1023) / 1024 * 1024;
      const int b_stage = 2 * p.BN * 128;
      const char* as = getenv("SCFLOW_TC_HALO_ASTAGES");
      int sa = as ? atoi(as) : 2;
      if (sa < 2) sa = 2;
      if (sa > 2 && (232448 - 1024 - tc_header(8) - sa * a_stage) / b_stage < 3) sa = 2;
      int sb = (232448 - 1024 - tc_header(8) - sa * a_stage) / b_stage;
      if (sb > 8) sb = 8;
      if (sb >= 2) {
        p.halo = 1; p.PW = PW; p.PH = PH; p.a_stages = sa; p.b_stages = sb;
        p.bk = 64;
        p.TW = 8; p.TH = 16; p.TB = 1;
        p.tiles_x = cdiv(p.W, 8); p.tiles_y = cdiv(p.H, 16);
        p.m_tiles = p.tiles_x * p.tiles_y * d.B;
        p.num_tiles = p.m_tiles * cdiv(d.cout_pad, p.BN);
        ew = 8; ctas_per_sm = 1;
        smem = 1024 + tc_header(8) + sa * a_stage + sb * b_stage;
      }
    }
  }
  // ---- tap split: requested by the caller (d.ksplit > 1) for plain linear layers whose partial maps it will add itself
  p.ksplit = 1; p.taps_per_split = p.num_taps; p.split_stride = 0;
  if (d.ksplit > 1) {
    SCF_REQUIRE(!p.halo && d.epi == SCF_EPI_ACT && d.act == SCF_ACT_NONE && !d.bias && !d.aux0 && !d.stats && d.out_f32 && !d.out_hl &&
                    !d.w_batched && p.num_taps % d.ksplit == 0 && d.split_stride > 0,
                SCF_ERR_ARG, "scf_conv2d_tc: tap split needs a plain linear fp32-output layer (no bias / activation) and num_taps %% ksplit == 0");
    p.ksplit = d.ksplit; p.taps_per_split = p.num_taps / d.ksplit; p.split_stride = d.split_stride;
    p.num_tiles *= p.ksplit;
  }
  g_last_m_tiles = p.m_tiles;
  g_last_tiles_per_img = p.TB == 1 ? p.tiles_x * p.tiles_y : 0;
  g_last_stat_rows_per_img = g_last_tiles_per_img * 4;
  KernelFn kernel = table[p.pair ? 2 : (ew == 8 ? 1 : 0)][ki];
  // ---- cluster size: CTAs of consecutive pixel tiles (same N tile, same sample when the weights are batched) share each
  // weight tile through TMA multicast, which divides the weight share of the L2 -> shared-memory operand traffic - the
  // resource this kernel saturates first (ncu: 8-9 TB/s) - by the cluster size
  p.cluster = 1;
  int max_clusters = 0;
  if (ctas_per_sm == 1) {
    const char* ce = getenv("SCFLOW_TC_CLUSTER");
    const int want = p.pair ? 2 : (ce ? atoi(ce) : 1);   // measured on B200 (tools/trace_gru.py, bench_conv_tc.py): no gain from 2 or 4 - the main
                                          // loop is MMA-bound at the power-capped clock, not operand-delivery-bound - so off by default
    for (int cs = want >= 4 ? 4 : (want >= 2 ? 2 : 1); cs > 1; cs >>= 1) {
      if (p.m_tiles % cs != 0 || (p.BN / cs) % 8 != 0 || p.num_tiles < 2 * cs) continue;
      if (d.w_batched && (p.tiles_x * p.tiles_y) % cs != 0) continue;
      static std::mutex mu;
      static int cache[3][6][5][8];      // [variant][kernel][cluster][stages] -> max co-resident clusters (+1; 0 = not queried)
      std::lock_guard<std::mutex> lock(mu);
      int& slot = cache[p.pair ? 2 : (ew == 8 ? 1 : 0)][ki][cs][p.stages];
      if (slot == 0) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(num_sms / cs * cs); qc.blockDim = dim3(64 + 32 * ew); qc.dynamicSmemBytes = smem;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &qc) != cudaSuccess) { cudaGetLastError(); n = 0; }
        slot = n + 1;
      }
      if (slot - 1 >= 1) { p.cluster = cs; max_clusters = slot - 1; break; }
    }
    SCF_REQUIRE(!p.pair || p.cluster == 2, SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: CTA pairs cannot be scheduled on this device");
  }
  p.bias = d.bias; p.scale = d.scale; p.epi = d.epi; p.act = d.act;
  p.out_f32 = d.out_f32; p.out_f32_stride = d.out_f32_stride; p.out_f32_coff = d.out_f32_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.out_hl_plane = d.out_hl_plane; p.out_hl_stride = d.out_hl_stride;
  p.out_hl_coff = d.out_hl_coff;
  p.aux0 = d.aux0; p.aux0_stride = d.aux0_stride; p.aux1 = d.aux1; p.aux1_stride = d.aux1_stride;
  p.out2_hl = reinterpret_cast<__nv_bfloat16*>(d.out2_hl); p.out2_hl_plane = d.out2_hl_plane; p.out2_hl_stride = d.out2_hl_stride;
  p.pre = d.pre; p.pre_stride = d.pre_stride; p.stats = d.stats;
  if (d.pre)
    SCF_REQUIRE(d.epi != SCF_EPI_ACT && d.pre_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d.pre) % 16 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: pre needs a GRU epilogue, 16B alignment and stride %% 4 == 0");
  {
    const char* dt = getenv("SCFLOW_TC_DBG_TIMES");       // address of a device buffer, hex (timing experiments only)
    p.dbg_times = dt ? reinterpret_cast<long long*>(strtoull(dt, nullptr, 16)) : nullptr;
    const char* de = getenv("SCFLOW_TC_DBG_EPI");
    p.dbg_epi = de ? atoi(de) : 0;
  }
  CUtensorMap tmA[3], tmW;
  int wcoff = 0;
  for (int s = 0; s < 3; ++s) {
    if (s >= d.nseg) { tmA[s] = tmA[0]; p.seg_chunks[s] = 0; p.seg_wcoff[s] = 0; p.seg_last_ks[s] = 0; continue; }
    const scf_tc_seg& sg = d.seg[s];
    SCF_REQUIRE(sg.ptr && sg.nch > 0 && sg.nch % 8 == 0 && sg.coff % 8 == 0 && sg.stride % 8 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: segment %d channels/offset/stride must be multiples of 8", s);
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(sg.ptr) + sg.coff;
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(base) % 16 == 0 && (sg.plane_stride * 2) % 16 == 0, SCF_ERR_ALIGN,
                "scf_conv2d_tc: segment %d must be 16B aligned", s);
    cuuint64_t dims[5] = {(cuuint64_t)sg.nch, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B, 2};
    cuuint64_t str[4] = {(cuuint64_t)sg.stride * 2, (cuuint64_t)d.W * sg.stride * 2, (cuuint64_t)d.H * d.W * sg.stride * 2,
                         (cuuint64_t)sg.plane_stride * 2};
    // box = elements TRAVERSED per dimension; with element strides (1,s,s,1,1) it deposits TW x TH x TB pixels
    // a one-row (one-column) tile needs no element stride along that axis: the producer's coordinate already carries it
    const int bsx = p.TW == 1 ? 1 : p.sx, bsy = p.TH == 1 ? 1 : p.sy;
    cuuint32_t box[5] = {(cuuint32_t)p.bk, (cuuint32_t)(p.TW * bsx), (cuuint32_t)(p.TH * bsy), (cuuint32_t)p.TB, 2};
    if (p.halo) { box[1] = (cuuint32_t)p.PW; box[2] = (cuuint32_t)p.PH; }
    cuuint32_t estr[5] = {1, (cuuint32_t)bsx, (cuuint32_t)bsy, 1, 1};
    const CUtensorMapSwizzle swz = p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    SCF_TRY(encode_map(&tmA[s], base, 5, dims, str, box, estr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, swz));
    p.seg_chunks[s] = cdiv(sg.nch, p.bk);
    p.seg_last_ks[s] = cdiv(sg.nch - (p.seg_chunks[s] - 1) * p.bk, 16);
    p.seg_wcoff[s] = wcoff;
    wcoff += sg.nch;
  }
  SCF_REQUIRE(wcoff <= d.cin_pad, SCF_ERR_ARG, "scf_conv2d_tc: segments carry %d channels but the packed weight has cin_pad %d", wcoff, d.cin_pad);
  {
    const int third = d.w_batched ? d.B : p.num_taps;
    cuuint64_t dims[4] = {(cuuint64_t)d.cin_pad, (cuuint64_t)d.cout_pad, (cuuint64_t)third, 2};
    cuuint64_t str[3] = {(cuuint64_t)d.cin_pad * 2, (cuuint64_t)d.cout_pad * d.cin_pad * 2,
                         d.w_plane_stride > 0 ? (cuuint64_t)d.w_plane_stride * 2 : (cuuint64_t)third * d.cout_pad * d.cin_pad * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.bk, (cuuint32_t)(p.BN / p.cluster), 1, 1};
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: packed weight must be 16B aligned");
    SCF_TRY(encode_map(&tmW, d.w, 4, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                       p.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    auto al16 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
    bool ok = true;
    if (d.out_f32) ok = ok && al16(d.out_f32) && d.out_f32_stride % 4 == 0 && d.out_f32_coff % 4 == 0;
    if (d.out_hl) ok = ok && al16(d.out_hl) && d.out_hl_stride % 8 == 0 && d.out_hl_coff % 8 == 0 && (d.out_hl_plane * 2) % 16 == 0;
    if (d.aux0) ok = ok && al16(d.aux0) && d.aux0_stride % 4 == 0;
    if (d.aux1) ok = ok && al16(d.aux1) && d.aux1_stride % 4 == 0;
    if (d.out2_hl) ok = ok && al16(d.out2_hl) && d.out2_hl_stride % 8 == 0 && (d.out2_hl_plane * 2) % 16 == 0;
    if (d.epi == SCF_EPI_GRU_ZR) ok = ok && (d.cout / 2) % 32 == 0;
    const char* fe = getenv("SCFLOW_TC_FASTEPI");
    p.fast_epi = (ok && (fe ? atoi(fe) != 0 : true)) ? 1 : 0;
  }
  {
    // 256-bit direct stores need 32 B alignment of every 32-column block of every output
    auto al32 = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 32 == 0; };
    bool ok = p.fast_epi != 0;
    if (d.out_f32) ok = ok && al32(d.out_f32) && d.out_f32_stride % 8 == 0 && d.out_f32_coff % 8 == 0;
    if (d.out_hl) ok = ok && al32(d.out_hl) && d.out_hl_stride % 16 == 0 && d.out_hl_coff % 16 == 0 && (d.out_hl_plane * 2) % 32 == 0;
    if (d.out2_hl) ok = ok && al32(d.out2_hl) && d.out2_hl_stride % 16 == 0 && (d.out2_hl_plane * 2) % 32 == 0;
    const char* ds = getenv("SCFLOW_TC_DIRECT_ST");
    p.direct_st = (ok && (ds ? atoi(ds) != 0 : false)) ? 1 : 0;   // measured: no gain over the staged stores, off by default
  }
  if (d.stats)
    SCF_REQUIRE(p.fast_epi && p.TB == 1 && d.cout % 32 == 0 && d.out_f32 && d.epi == SCF_EPI_ACT &&
                    reinterpret_cast<uintptr_t>(d.stats) % 16 == 0,
                SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: stats need an aligned fp32 output, cout %% 32 == 0 and one sample per tile");
  int grid;
  if (p.cluster > 1) {
    const int groups = p.num_tiles / p.cluster;
    grid = (groups < max_clusters ? groups : max_clusters) * p.cluster;
  } else {
    const int max_ctas = num_sms * ctas_per_sm;
    grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  }
  {
    // programmatic dependent launch: this kernel's prologue (barrier init, TMEM allocation, descriptor prefetch) may overlap
    // the tail of the previous kernel on the stream; the kernel executes griddepcontrol.wait before it touches memory
    static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * ew); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (p.cluster > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = p.cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, tmA[0], tmA[1], tmA[2], tmW, p);
    if (le != cudaSuccess) { cudaGetLastError(); set_error("conv_tc_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  }
  return check_launch("conv_tc_kernel");
}

}  // namespace scf

extern "C" {

int scf_conv2d_tc(const scf_tc_conv_desc* d, void* stream) {
  SCF_REQUIRE(d != nullptr, SCF_ERR_ARG, "scf_conv2d_tc: null descriptor");
  return scf::conv2d_tc(*d, (cudaStream_t)stream);
}

}  // extern "C"

namespace scf {
// packs input channels [i_begin, i_begin + i_count) of an OIHW weight with I_total input channels at packed channels i_dst..
int pack_conv_weight_tc_range(const float* w_oihw, void* packed, int O, int I_total, int i_begin, int i_count, int i_dst, int kh,
                              int kw, int cin_pad, int cout_pad, int o_off, cudaStream_t st) {
  SCF_REQUIRE(w_oihw && packed && O > 0 && i_count > 0 && kh > 0 && kw > 0, SCF_ERR_ARG, "pack_conv_weight_tc_range: bad args");
  SCF_REQUIRE(i_begin >= 0 && i_begin + i_count <= I_total && cin_pad >= i_dst + i_count && cout_pad >= o_off + O, SCF_ERR_ARG,
              "pack_conv_weight_tc_range: range outside the weight / padding");
  const long long total = (long long)O * i_count * kh * kw;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_tc_kernel<<<blocks, 256, 0, st>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(packed), O, i_count, kh * kw, cin_pad,
                                                cout_pad, o_off, I_total, i_begin, i_dst);
  return check_launch("pack_weight_tc_kernel");
}

int pack_conv_weight_tc_foldx(const float* w_oihw, float* scratch, void* packed, int O, int C, int KH, int KW, int cin_pad,
                              int cout_pad, int o_off, cudaStream_t st) {
  SCF_REQUIRE(w_oihw && scratch && packed, SCF_ERR_ARG, "pack_conv_weight_tc_foldx: null pointer");
  fold_kx_weight_kernel<<<cdiv((long long)O * C * KH * KW, 256), 256, 0, st>>>(w_oihw, scratch, O, C, KH, KW);
  SCF_TRY(check_launch("fold_kx_weight_kernel"));
  return pack_conv_weight_tc_range(scratch, packed, O, KW * C, 0, KW * C, 0, KH, 1, cin_pad, cout_pad, o_off, st);
}
}  // namespace scf

extern "C" {

int scf_conv2d_tc_tiles(int B, int Hout, int Wout, int* tiles_per_sample) {
  if (B <= 0 || Hout <= 0 || Wout <= 0) return 0;
  int TW = 0, TH = 0, TB = 0;
  scf::pick_tile(B, Hout, Wout, false, TW, TH, TB);
  const int per = scf::cdiv(Wout, TW) * scf::cdiv(Hout, TH);
  if (tiles_per_sample) *tiles_per_sample = TB == 1 ? per : 0;
  return scf::conv2d_tc_max_tiles(B, Hout, Wout);      // upper bound over every tiling the library may choose
}

int scf_pack_conv_weight_tc(const float* w_oihw, void* packed, int O, int I, int kh, int kw, int cin_pad, int cout_pad,
                            int o_off, void* stream) {
  SCF_REQUIRE(w_oihw && packed && O > 0 && I > 0 && kh > 0 && kw > 0, SCF_ERR_ARG, "scf_pack_conv_weight_tc: bad args");
  SCF_REQUIRE(cin_pad >= I && cout_pad >= o_off + O, SCF_ERR_ARG, "scf_pack_conv_weight_tc: padding smaller than the weight");
  return scf::pack_conv_weight_tc_range(w_oihw, packed, O, I, 0, I, 0, kh, kw, cin_pad, cout_pad, o_off, (cudaStream_t)stream);
}

int scf_nchw_to_nhwc_split(const float* src, void* dst_hl, long long plane_stride, int dst_stride, int dst_coff,
                           float* dst_f32, int dst_f32_stride, int B, int C, int H, int W, void* stream) {
  SCF_REQUIRE(src && dst_hl && B > 0 && C > 0 && H > 0 && W > 0, SCF_ERR_ARG, "scf_nchw_to_nhwc_split: bad args");
  dim3 grid(scf::cdiv(H * W, 32), scf::cdiv(C, 32), B);
  scf::nchw_to_nhwc_split_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst_hl), plane_stride, dst_stride, dst_coff, dst_f32, dst_f32_stride, C, H * W);
  return scf::check_launch("nchw_to_nhwc_split_kernel");
}

int scf_split_copy(const float* src, int src_stride, int src_coff, void* dst_hl, long long plane_stride, int dst_stride,
                   int dst_coff, long long npix, int nch, void* stream) {
  SCF_REQUIRE(src && dst_hl && npix > 0 && nch > 0, SCF_ERR_ARG, "scf_split_copy: bad args");
  const long long total = npix * nch;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  scf::launch_pdl(scf::split_copy_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, src_stride, src_coff, reinterpret_cast<__nv_bfloat16*>(dst_hl),
                                                                  plane_stride, dst_stride, dst_coff, npix, nch);
  return scf::check_launch("split_copy_kernel");
}

}  // extern "C"
