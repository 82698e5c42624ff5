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
constexpr int TC_THREADS = 192;   // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2-5 epilogue
constexpr int TC_STAGE_ROW = 80;  // epilogue staging: 32 rows x 64 B per warp, rows padded to 80 B (conflict-free 16 B accesses)
constexpr int TC_HEADER = 1024 + 4 * 32 * TC_STAGE_ROW + 2048;   // barriers + staging + 2 bias buffers (multiple of 1024)
constexpr uint32_t TC_A_PLANE = TC_BM * TC_BK * 2;   // 16 KB
constexpr int TC_MAX_STAGES = 4;

struct TcParams {
  int nseg, seg_chunks[3], seg_wcoff[3];
  int B, H, W, kh, kw, ph, pw;   // H, W: OUTPUT spatial size
  int stride;                    // 1 or 2 (input is sampled at stride*out + tap - pad)
  int TW, TH, TB, tiles_x, tiles_y;   // tile = TW x TH pixels x TB samples = 128 rows
  int m_tiles, num_tiles;        // pixel tiles, pixel tiles x N tiles (tile t: n = t / m_tiles, m = t % m_tiles)
  int BN, cout, num_taps, w_batched, stages;
  int acc_cols, tmem_cols;       // TMEM columns of one accumulator buffer / allocated (two buffers)
  long long* dbg_times;          // optional [grid][8] globaltimer stamps of each CTA's first tile (tools/trace_conv_tc.py)
  // stacked-N mode (BN <= 128): the hi and lo weight planes are adjacent in shared memory, so ONE MMA with N = 2*BN
  // forms A_hi*[W_hi;W_lo] into accumulator columns [0,BN) and [BN,2BN); a second MMA adds A_lo*W_hi into [0,BN).
  // Two instructions per k-step instead of three (less shared-memory operand traffic per FLOP); the epilogue adds
  // the two halves.
  int stackn;
  int fast_epi;                  // every global access of the epilogue is 16 B aligned: use the coalesced staged path
  const float* bias; float scale; int epi, act;
  float* out_f32; int out_f32_stride, out_f32_coff;
  __nv_bfloat16* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff;
  const float* aux0; int aux0_stride; const float* aux1; int aux1_stride;
  __nv_bfloat16* out2_hl; long long out2_hl_plane; int out2_hl_stride;
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

// ---- coalescing through shared memory.  In the accumulator layout a thread owns one pixel row, so direct 16 B stores
// of a warp touch 32 different cache lines (measured: the epilogue of a 128x256 tile took 5.6 us, LSU-bound).  These
// helpers move a [32 rows x 64 B] block between the warp's registers and global memory with 4 lanes per row, i.e.
// 8 fully used 64 B row segments per instruction.  `mypix` = global pixel index of this thread's row, or -1.
__device__ __forceinline__ void stage_store64(uint32_t sbuf, int lane, char* gbase, long long row_bytes, long long mypix,
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
    const long long pr = __shfl_sync(0xffffffffu, mypix, r);
    if (pr >= 0) *reinterpret_cast<uint4*>(gbase + pr * row_bytes + seg * 16) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void stage_load64(uint32_t sbuf, int lane, const char* gbase, long long row_bytes, long long mypix,
                                             float (&out)[16]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), seg = lane & 3;
    const long long pr = __shfl_sync(0xffffffffu, mypix, r);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pr >= 0) v = __ldg(reinterpret_cast<const uint4*>(gbase + pr * row_bytes + seg * 16));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + r * TC_STAGE_ROW + seg * 16), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w) : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sbuf + lane * TC_STAGE_ROW + i * 16) : "memory");
    out[4 * i] = __uint_as_float(v.x); out[4 * i + 1] = __uint_as_float(v.y);
    out[4 * i + 2] = __uint_as_float(v.z); out[4 * i + 3] = __uint_as_float(v.w);
  }
  __syncwarp();
}
// 16 fp32 values of this thread's row -> one 64 B fp32 block
__device__ __forceinline__ void stage_store_f32(uint32_t sbuf, int lane, float* gbase, int stride, long long mypix, const float* v) {
  uint4 d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    d[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase), (long long)stride * 4, mypix, d);
}
// 32 fp32 values of this thread's row -> split-bf16: one 64 B block in the hi plane, one in the lo plane
__device__ __forceinline__ void stage_store_split32(uint32_t sbuf, int lane, __nv_bfloat16* gbase, long long plane, int stride,
                                                    long long mypix, const float* v) {
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
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase), (long long)stride * 2, mypix, hi);
  stage_store64(sbuf, lane, reinterpret_cast<char*>(gbase + plane), (long long)stride * 2, mypix, lo);
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

template <int EPI, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
  // header: [0,256) barriers + TMEM pointer | [1024, 11264) epilogue staging | [11264, 13312) two bias buffers ; then the
  // operand ring: `stages` x {A hi, A lo, W hi, W lo}
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 64, bar_tfull = smem_base + 128, bar_tempty = smem_base + 144,
                 tmem_slot = smem_base + 192;
  const uint32_t stage0 = smem_base + 1024;
  const uint32_t bias0 = smem_base + 1024 + 4 * 32 * TC_STAGE_ROW;
  const uint32_t tiles0 = smem_base + TC_HEADER;
  const uint32_t b_plane = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = 2 * TC_A_PLANE + 2 * b_plane;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stamp = [&](int slot) {
    if (p.dbg_times) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg_times[(long long)blockIdx.x * 8 + slot] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int chunks_per_tap = p.seg_chunks[0] + p.seg_chunks[1] + p.seg_chunks[2];
  const int num_chunks = p.num_taps * chunks_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (p.nseg > 1) prefetch_tmap(&tmA1);
    if (p.nseg > 2) prefetch_tmap(&tmA2);
    prefetch_tmap(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);     // MMA issuer -> epilogue: accumulator a complete
      mbar_init(bar_tempty + 8 * a, 4);    // 4 epilogue warps -> MMA issuer: accumulator a drained
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) stamp(1);

  // Persistent CTA: tiles blockIdx.x, +gridDim.x, ...  The three roles walk the same tile sequence; the operand ring and
  // the two TMEM accumulators run on, so tile i+1's loads and MMAs overlap tile i's epilogue.
  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
        const int b = (mt / tiles_per_img) * p.TB, tr = mt % tiles_per_img;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * p.BN;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int ky = tap / p.kw, kx = tap - ky * p.kw;
          const int cx = x0 * p.stride + kx - p.pw, cy = y0 * p.stride + ky - p.ph;
          for (int s = 0; s < p.nseg; ++s) {
            const CUtensorMap* tm = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : &tmA2);
            for (int cc = 0; cc < p.seg_chunks[s]; ++cc) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              const uint32_t full = bar_full + 8 * stage;
              mbar_arrive_expect_tx(full, stage_bytes);
              const uint32_t a_dst = tiles0 + stage * stage_bytes;
              tma_load_5d(a_dst, tm, full, cc * TC_BK, cx, cy, b, 0);
              tma_load_4d(a_dst + 2 * TC_A_PLANE, &tmW, full, p.seg_wcoff[s] + cc * TC_BK, n0, p.w_batched ? b : tap, 0);
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer
      const uint32_t idesc = make_idesc_bf16(TC_BM, p.BN), idesc2 = make_idesc_bf16(TC_BM, 2 * p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_tempty + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);   // epilogue of tile it-2 has drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
        for (int c = 0; c < num_chunks; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (it == 0 && c == 0) stamp(2);
          const uint32_t a_addr = tiles0 + stage * stage_bytes;
          const uint64_t a_hi = make_smem_desc_sw128(a_addr, 1024), a_lo = make_smem_desc_sw128(a_addr + TC_A_PLANE, 1024);
          const uint64_t b_hi = make_smem_desc_sw128(a_addr + 2 * TC_A_PLANE, 1024);
          const uint64_t b_lo = make_smem_desc_sw128(a_addr + 2 * TC_A_PLANE + b_plane, 1024);
          if (p.stackn) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 32 >> 4);     // 16 bf16 = 32 B along the swizzled 128 B row
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc2, (c > 0 || k > 0) ? 1u : 0u);   // N = 2*BN: [W_hi;W_lo]
              umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 32 >> 4);
              umma_bf16(d_tmem, a_hi + ko, b_hi + ko, idesc, (c > 0 || k > 0) ? 1u : 0u);
              umma_bf16(d_tmem, a_hi + ko, b_lo + ko, idesc, 1u);
              umma_bf16(d_tmem, a_lo + ko, b_hi + ko, idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * stage);     // frees the smem slot once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * acc);         // accumulator complete
        if (it == 0) stamp(3);
      }
    }
  } else {
    // ================= epilogue: warp w owns TMEM lanes 32*(w%4)..+31 ; lane = pixel row of the tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int bb = row / (p.TW * p.TH), rr = row - bb * (p.TW * p.TH);
    const int h = rr / p.TW, w = rr - h * p.TW;
    const int half = p.cout >> 1;
    const int ngroups = p.BN / 16;
    const uint32_t sbuf = stage0 + (uint32_t)(warp - 2) * 32 * TC_STAGE_ROW;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
      const int b = (mt / tiles_per_img) * p.TB, tr = mt % tiles_per_img;
      const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
      const int y = ty * p.TH + h, x = tx * p.TW + w, n0 = nt * p.BN;
      const bool valid = y < p.H && x < p.W && b + bb < p.B;
      const long long pix = ((long long)(b + bb) * p.H + y) * p.W + x;
      const uint32_t bias_s = bias0 + (uint32_t)acc * 1024u;
      // stage this tile's bias slice (global-load latency would otherwise sit inside every column slab); the named barrier
      // also orders it against the slowest warp still reading the buffer two tiles ago
      for (int i = threadIdx.x - 64; i < p.BN; i += 128) {
        const float bvl = (p.bias && n0 + i < p.cout) ? __ldg(p.bias + n0 + i) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4 * i), "f"(bvl) : "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_tfull + 8 * acc, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      if (it == 0 && threadIdx.x == 64) stamp(4);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      int g_begin = 0;
      if (p.fast_epi) {
        // ---- coalesced path: 32-column slabs, every global access staged through shared memory
        const long long mypix = valid ? pix : -1;
        const int nslab = (p.cout - n0 < p.BN ? p.cout - n0 : p.BN) / 32;    // full slabs only; the tail uses the plain path
#pragma unroll 1
        for (int sl = 0; sl < nslab; ++sl) {
          float v[32];
          __syncwarp();
          tmem_ld32(t_addr + (uint32_t)(sl * 32), v);
          if (p.stackn) {
            float v2[32];
            tmem_ld32(t_addr + (uint32_t)(p.BN + sl * 32), v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += v2[i];
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
          if (EPI == SCF_EPI_ACT) {
            if (p.aux0) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float rv[16];
                stage_load64(sbuf, lane, reinterpret_cast<const char*>(p.aux0 + nb + hh * 16), (long long)p.aux0_stride * 4, mypix, rv);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[hh * 16 + i] += rv[i];
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = act_ct<ACT>(v[i]);
            if (p.out_f32) {
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb, p.out_f32_stride, mypix, v);
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb + 16, p.out_f32_stride, mypix, v + 16);
            }
            if (p.out_hl) stage_store_split32(sbuf, lane, p.out_hl + p.out_hl_coff + nb, p.out_hl_plane, p.out_hl_stride, mypix, v);
          } else if (EPI == SCF_EPI_GRU_ZR) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = sigmoid_fast(v[i]);
            if (nb < half) {
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb, p.out_f32_stride, mypix, v);
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb + 16, p.out_f32_stride, mypix, v + 16);
            } else {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float hv[16];
                stage_load64(sbuf, lane, reinterpret_cast<const char*>(p.aux0 + (nb - half) + hh * 16), (long long)p.aux0_stride * 4, mypix, hv);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[hh * 16 + i] *= hv[i];
              }
              stage_store_split32(sbuf, lane, p.out2_hl + (nb - half), p.out2_hl_plane, p.out2_hl_stride, mypix, v);
            }
          } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float hv[16], zv[16];
              stage_load64(sbuf, lane, reinterpret_cast<const char*>(p.aux0 + nb + hh * 16), (long long)p.aux0_stride * 4, mypix, hv);
              stage_load64(sbuf, lane, reinterpret_cast<const char*>(p.aux1 + nb + hh * 16), (long long)p.aux1_stride * 4, mypix, zv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[hh * 16 + i] = (1.f - zv[i]) * hv[i] + zv[i] * tanh_fast(v[hh * 16 + i]);
            }
            if (p.out_f32) {
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb, p.out_f32_stride, mypix, v);
              stage_store_f32(sbuf, lane, p.out_f32 + p.out_f32_coff + nb + 16, p.out_f32_stride, mypix, v + 16);
            }
            if (p.out_hl) stage_store_split32(sbuf, lane, p.out_hl + p.out_hl_coff + nb, p.out_hl_plane, p.out_hl_stride, mypix, v);
          }
        }
        g_begin = nslab * 2;
      }
      // ---- plain path (ragged channel tails, unaligned outputs): 16-column groups, direct per-thread stores
#pragma unroll 1
      for (int g = g_begin; g < ngroups; ++g) {
        float v[16];
        __syncwarp();
        tmem_ld16(t_addr + (uint32_t)(g * 16), v);
        if (p.stackn) {                        // second half of the stacked accumulator: A_hi * W_lo
          float v2[16];
          tmem_ld16(t_addr + (uint32_t)(p.BN + g * 16), v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += v2[i];
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
          if (p.out_f32) store_f32x16(p.out_f32 + pix * p.out_f32_stride + p.out_f32_coff + nb, v, nvalid);
          if (p.out_hl) store_split16(p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb, p.out_hl_plane, v, nvalid);
        } else if (EPI == SCF_EPI_GRU_ZR) {    // cout = 2*Ch, Ch % 16 == 0: a group is entirely z or entirely r
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = sigmoid_fast(v[i]);
          if (nb < half) {                     // z gate -> fp32 (read back by the q convolution's epilogue)
            store_f32x16(p.out_f32 + pix * p.out_f32_stride + p.out_f32_coff + nb, v, 16);
          } else {                             // r gate -> r*h as split-bf16, the q convolution's first input segment
            float hv[16];
            load16(p.aux0 + pix * p.aux0_stride + (nb - half), hv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= hv[i];
            store_split16(p.out2_hl + pix * p.out2_hl_stride + (nb - half), p.out2_hl_plane, v, 16);
          }
        } else {                               // SCF_EPI_GRU_Q: h' = (1-z) h + z tanh(.)   (cout % 16 == 0)
          float hv[16], zv[16];
          load16(p.aux0 + pix * p.aux0_stride + nb, hv);
          load16(p.aux1 + pix * p.aux1_stride + nb, zv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (1.f - zv[i]) * hv[i] + zv[i] * tanh_fast(v[i]);
          if (p.out_f32) store_f32x16(p.out_f32 + pix * p.out_f32_stride + p.out_f32_coff + nb, v, 16);
          if (p.out_hl) store_split16(p.out_hl + pix * p.out_hl_stride + p.out_hl_coff + nb, p.out_hl_plane, v, 16);
        }
      }
      // this warp has read everything it needs from accumulator `acc`: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (it == 0 && threadIdx.x == 64) stamp(5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 32) stamp(6);
}

// ------------------------------------------------------------------ prep kernels
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, int O, int I, int taps,
                                      int cin_pad, int cout_pad, int o_off) {
  const long long total = (long long)O * I * taps;
  const long long plane = (long long)taps * cout_pad * cin_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const long long r = idx / I;
    const int o = (int)(r % O);
    const int tap = (int)(r / O);
    __nv_bfloat16 hi, lo;
    split_bf16(w[((long long)o * I + i) * taps + tap], hi, lo);
    const long long dst = ((long long)tap * cout_pad + o_off + o) * cin_pad + i;
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

// ------------------------------------------------------------------ host side
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

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr) {
  EncodeTiledFn fn = get_encode_fn();
  SCF_REQUIRE(fn != nullptr, SCF_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  elem_strides ? elem_strides : ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

  TcParams p = {};
  p.nseg = d.nseg;
  const int stride = d.stride == 2 ? 2 : 1;
  SCF_REQUIRE(d.stride == 0 || d.stride == 1 || d.stride == 2, SCF_ERR_ARG, "scf_conv2d_tc: stride must be 1 or 2");
  p.stride = stride;
  p.kh = d.kh; p.kw = d.kw; p.ph = d.kh / 2; p.pw = d.kw / 2;
  p.B = d.B; p.H = (d.H + 2 * p.ph - d.kh) / stride + 1; p.W = (d.W + 2 * p.pw - d.kw) / stride + 1;
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
  const int stage_bytes = 2 * (int)TC_A_PLANE + 2 * p.BN * 128;
  p.stages = (232448 - 1024 - TC_HEADER) / stage_bytes;
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  SCF_REQUIRE(p.stages >= 2, SCF_ERR_UNSUPPORTED, "scf_conv2d_tc: tile does not fit in shared memory");
  {
    const char* sv = getenv("SCFLOW_TC_STACKN");
    p.stackn = (p.BN <= 128 && (sv ? atoi(sv) != 0 : true)) ? 1 : 0;
  }
  p.acc_cols = p.stackn ? 2 * p.BN : p.BN;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.acc_cols) p.tmem_cols <<= 1;
  int ctas_per_sm = 1;
  {
    // Small tiles: keep two CTAs resident per SM (<= 113 KB of shared memory and <= 256 TMEM columns each) so that the two
    // CTAs' main loops interleave; larger tiles run one persistent CTA per SM with double-buffered accumulators.
    const char* ov = getenv("SCFLOW_TC_OCC2");
    const bool occ2 = ov ? atoi(ov) != 0 : true;
    if (occ2 && 2 * stage_bytes + 1024 + TC_HEADER <= 115712 && p.tmem_cols <= 256 && p.num_tiles >= 4 * num_sms) {
      p.stages = (115712 - 1024 - TC_HEADER) / stage_bytes;
      ctas_per_sm = 2;
    }
  }
  const int smem = 1024 + TC_HEADER + p.stages * stage_bytes;
  p.bias = d.bias; p.scale = d.scale; p.epi = d.epi; p.act = d.act;
  p.out_f32 = d.out_f32; p.out_f32_stride = d.out_f32_stride; p.out_f32_coff = d.out_f32_coff;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(d.out_hl); p.out_hl_plane = d.out_hl_plane; p.out_hl_stride = d.out_hl_stride;
  p.out_hl_coff = d.out_hl_coff;
  p.aux0 = d.aux0; p.aux0_stride = d.aux0_stride; p.aux1 = d.aux1; p.aux1_stride = d.aux1_stride;
  p.out2_hl = reinterpret_cast<__nv_bfloat16*>(d.out2_hl); p.out2_hl_plane = d.out2_hl_plane; p.out2_hl_stride = d.out2_hl_stride;
  {
    const char* dt = getenv("SCFLOW_TC_DBG_TIMES");       // address of a device buffer, hex (timing experiments only)
    p.dbg_times = dt ? reinterpret_cast<long long*>(strtoull(dt, nullptr, 16)) : nullptr;
  }
  CUtensorMap tmA[3], tmW;
  int wcoff = 0;
  for (int s = 0; s < 3; ++s) {
    if (s >= d.nseg) { tmA[s] = tmA[0]; p.seg_chunks[s] = 0; p.seg_wcoff[s] = 0; continue; }
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
    cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)(p.TW * stride), (cuuint32_t)(p.TH * stride), (cuuint32_t)p.TB, 2};
    cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
    SCF_TRY(encode_map(&tmA[s], base, 5, dims, str, box, estr));
    p.seg_chunks[s] = cdiv(sg.nch, TC_BK);
    p.seg_wcoff[s] = wcoff;
    wcoff += sg.nch;
  }
  SCF_REQUIRE(wcoff <= d.cin_pad, SCF_ERR_ARG, "scf_conv2d_tc: segments carry %d channels but the packed weight has cin_pad %d", wcoff, d.cin_pad);
  {
    const int third = d.w_batched ? d.B : p.num_taps;
    cuuint64_t dims[4] = {(cuuint64_t)d.cin_pad, (cuuint64_t)d.cout_pad, (cuuint64_t)third, 2};
    cuuint64_t str[3] = {(cuuint64_t)d.cin_pad * 2, (cuuint64_t)d.cout_pad * d.cin_pad * 2,
                         (cuuint64_t)third * d.cout_pad * d.cin_pad * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.BN, 1, 2};
    SCF_REQUIRE(reinterpret_cast<uintptr_t>(d.w) % 16 == 0, SCF_ERR_ALIGN, "scf_conv2d_tc: packed weight must be 16B aligned");
    SCF_TRY(encode_map(&tmW, d.w, 4, dims, str, box));
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
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
  KernelFn kernel = nullptr;
  if (d.epi == SCF_EPI_GRU_ZR) kernel = conv_tc_kernel<SCF_EPI_GRU_ZR, SCF_ACT_SIGMOID>;
  else if (d.epi == SCF_EPI_GRU_Q) kernel = conv_tc_kernel<SCF_EPI_GRU_Q, SCF_ACT_TANH>;
  else if (d.act == SCF_ACT_NONE) kernel = conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_NONE>;
  else if (d.act == SCF_ACT_RELU) kernel = conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_RELU>;
  else if (d.act == SCF_ACT_SIGMOID) kernel = conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_SIGMOID>;
  else if (d.act == SCF_ACT_TANH) kernel = conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_TANH>;
  SCF_REQUIRE(kernel != nullptr, SCF_ERR_ARG, "scf_conv2d_tc: bad epilogue / activation");
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    KernelFn all[6] = {conv_tc_kernel<SCF_EPI_GRU_ZR, SCF_ACT_SIGMOID>, conv_tc_kernel<SCF_EPI_GRU_Q, SCF_ACT_TANH>,
                       conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_NONE>, conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_RELU>,
                       conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_SIGMOID>, conv_tc_kernel<SCF_EPI_ACT, SCF_ACT_TANH>};
    for (KernelFn f : all) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) attr_err = e;
    }
  });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(conv_tc_kernel): %s", cudaGetErrorString(attr_err));
  const int max_ctas = num_sms * ctas_per_sm;
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  kernel<<<grid, TC_THREADS, smem, st>>>(tmA[0], tmA[1], tmA[2], tmW, p);
  return check_launch("conv_tc_kernel");
}

}  // namespace scf

extern "C" {

int scf_conv2d_tc(const scf_tc_conv_desc* d, void* stream) {
  SCF_REQUIRE(d != nullptr, SCF_ERR_ARG, "scf_conv2d_tc: null descriptor");
  return scf::conv2d_tc(*d, (cudaStream_t)stream);
}

int scf_pack_conv_weight_tc(const float* w_oihw, void* packed, int O, int I, int kh, int kw, int cin_pad, int cout_pad,
                            int o_off, void* stream) {
  SCF_REQUIRE(w_oihw && packed && O > 0 && I > 0 && kh > 0 && kw > 0, SCF_ERR_ARG, "scf_pack_conv_weight_tc: bad args");
  SCF_REQUIRE(cin_pad >= I && cout_pad >= o_off + O, SCF_ERR_ARG, "scf_pack_conv_weight_tc: padding smaller than the weight");
  const long long total = (long long)O * I * kh * kw;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  scf::pack_weight_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(packed), O, I,
                                                                      kh * kw, cin_pad, cout_pad, o_off);
  return scf::check_launch("pack_weight_tc_kernel");
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
  scf::split_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, src_stride, src_coff, reinterpret_cast<__nv_bfloat16*>(dst_hl),
                                                                  plane_stride, dst_stride, dst_coff, npix, nch);
  return scf::check_launch("split_copy_kernel");
}

}  // extern "C"
