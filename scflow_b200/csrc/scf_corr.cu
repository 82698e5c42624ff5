// Correlation pyramid lookup (HBM/L2-bound gather) and the fp32 reference build of the pyramid.
//
// Index semantics follow models/utils/corr_lookup.py:102-136 exactly:
//   centre c = (x + flow_x, y + flow_y) / 2^level ; tap (a,b) samples (c_x + a - r, c_y + b - r)  [x-major window]
//   g = c*2/(W_l-1) - 1 ; i = ((g+1)*0.5)*(W_l-1)        -- replayed with round-to-nearest intrinsics, no FMA
//   bilinear taps floor(i), floor(i)+1 ; out-of-range taps contribute 0 (padding_mode='zeros').
#include "scf_common.cuh"
#include <stdlib.h>
#include "scf_tc.cuh"

namespace scf {

int conv2d_f32(const scf_conv_desc& d, cudaStream_t st);

constexpr int kMaxLevels = 6;

struct LookupParams {
  const float* lvl[kMaxLevels];
  int hl[kMaxLevels], wl[kMaxLevels];
  int num_levels, radius;
  const float* flow8;
  const float* mask;
  float* out;
  int out_stride, out_coff;
  __nv_bfloat16* out_hl;      // when set: split-bf16 planes instead of fp32 (pad channels up to out_stride zeroed)
  long long out_hl_plane;
  int H8, W8;
  long long nq;
};

// un-normalised sample coordinate of one axis, bit-exact replay of the reference's fp32 sequence
__device__ __forceinline__ float lookup_coord(float centre, int off, int size) {
  const float p = __fadd_rn(centre, (float)off);
  const float den = (float)(size - 1 > 1 ? size - 1 : 1);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

__global__ void __launch_bounds__(256) corr_lookup_kernel(const LookupParams p) {
  const int lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= p.nq) return;
  const int P = p.H8 * p.W8;
  const int pix = (int)(q % P);
  const int y = pix / p.W8, x = pix - y * p.W8;
  const float2 f = *reinterpret_cast<const float2*>(p.flow8 + q * 2);
  const float gx0 = __fadd_rn((float)x, f.x), gy0 = __fadd_rn((float)y, f.y);
  const float mval = p.mask ? p.mask[q] : 1.f;
  const int r = p.radius, k = 2 * r + 1, kk = k * k;
  float* outq = p.out ? p.out + q * p.out_stride + p.out_coff : nullptr;
  __nv_bfloat16* outh = p.out_hl ? p.out_hl + q * p.out_stride : nullptr;
  float inv = 1.f;
  for (int l = 0; l < p.num_levels; ++l, inv *= 0.5f) {
    const int hl = p.hl[l], wl = p.wl[l];
    const float* vol = p.lvl[l] + q * (long long)(hl * wl);
    const float cx = __fmul_rn(gx0, inv), cy = __fmul_rn(gy0, inv);   // division by 2^l is exact
    for (int tap = lane; tap < kk; tap += 32) {
      const int a = tap / k, b = tap - a * k;
      const float ix = lookup_coord(cx, a - r, wl);
      const float iy = lookup_coord(cy, b - r, hl);
      const float x0f = floorf(ix), y0f = floorf(iy);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const float wx1 = ix - x0f, wx0 = (x0f + 1.f) - ix;
      const float wy1 = iy - y0f, wy0 = (y0f + 1.f) - iy;
      const bool xin0 = x0 >= 0 && x0 < wl, xin1 = x0 + 1 >= 0 && x0 + 1 < wl;
      const bool yin0 = y0 >= 0 && y0 < hl, yin1 = y0 + 1 >= 0 && y0 + 1 < hl;
      float acc = 0.f;
      if (yin0) {
        const float* row = vol + y0 * wl;
        if (xin0) acc += __ldg(row + x0) * (wx0 * wy0);
        if (xin1) acc += __ldg(row + x0 + 1) * (wx1 * wy0);
      }
      if (yin1) {
        const float* row = vol + (y0 + 1) * wl;
        if (xin0) acc += __ldg(row + x0) * (wx0 * wy1);
        if (xin1) acc += __ldg(row + x0 + 1) * (wx1 * wy1);
      }
      acc *= mval;
      if (outq) outq[l * kk + tap] = acc;
      if (outh) {
        __nv_bfloat16 hi, lo;
        tc::split_bf16(acc, hi, lo);
        outh[l * kk + tap] = hi;
        outh[p.out_hl_plane + l * kk + tap] = lo;
      }
    }
  }
  if (outh) {
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (int c = p.num_levels * kk + lane; c < p.out_stride; c += 32) { outh[c] = zero; outh[p.out_hl_plane + c] = zero; }
  }
}

// Specialised lookup for the shipped configuration (4 levels, radius 4): one warp per query.
//  * the un-normalised coordinate of a tap depends on ONE window index per axis, so per level lanes 0-8 replay the
//    reference's fp32 sequence for the 9 x-offsets and lanes 9-17 for the 9 y-offsets (18 instead of 162 divisions);
//    taps fetch (floor, frac weights) by shuffle - neighbour indices stay bit-identical to the generic kernel;
//  * all 12 (level, round) tap groups are unrolled and their 48 gathers issued before any is consumed, so a query
//    pays one L2 latency instead of twelve.
__global__ void __launch_bounds__(256) corr_lookup_l4r4_kernel(const LookupParams p) {
  constexpr int R = 4, K = 9, KK = 81, L = 4, ROUNDS = 3;
  const int lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= p.nq) return;
  const int P = p.H8 * p.W8;
  const int pix = (int)(q % P);
  const int y = pix / p.W8, x = pix - y * p.W8;
  const float2 f = *reinterpret_cast<const float2*>(p.flow8 + q * 2);
  const float gx0 = __fadd_rn((float)x, f.x), gy0 = __fadd_rn((float)y, f.y);
  const float mval = p.mask ? p.mask[q] : 1.f;
  // ---- per-axis coordinates: lane a (0..8) -> x axis offset a-R ; lane 9+b -> y axis offset b-R
  int i0[L];
  float w0[L], w1[L];
  float inv = 1.f;
#pragma unroll
  for (int l = 0; l < L; ++l, inv *= 0.5f) {
    const bool isx = lane < K;
    const int off = (isx ? lane : lane - K) - R;
    const float c = __fmul_rn(isx ? gx0 : gy0, inv);
    const float ic = lookup_coord(c, off, isx ? p.wl[l] : p.hl[l]);
    const float fl = floorf(ic);
    i0[l] = (int)fl;
    w1[l] = ic - fl;
    w0[l] = (fl + 1.f) - ic;
  }
  // ---- gather: issue every load first
  float v[L][ROUNDS][4], wgt[L][ROUNDS][4];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int hl = p.hl[l], wl = p.wl[l];
    const float* vol = p.lvl[l] + q * (long long)(hl * wl);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int tap = r * 32 + lane;
      const int tc = tap < KK ? tap : KK - 1;          // lanes past the window replay the last tap (result discarded)
      const int a = tc / K, b = tc - a * K;
      const int x0 = __shfl_sync(0xffffffffu, i0[l], a), y0 = __shfl_sync(0xffffffffu, i0[l], K + b);
      const float wx0 = __shfl_sync(0xffffffffu, w0[l], a), wx1 = __shfl_sync(0xffffffffu, w1[l], a);
      const float wy0 = __shfl_sync(0xffffffffu, w0[l], K + b), wy1 = __shfl_sync(0xffffffffu, w1[l], K + b);
      const bool xin0 = x0 >= 0 && x0 < wl, xin1 = x0 + 1 >= 0 && x0 + 1 < wl;
      const bool yin0 = y0 >= 0 && y0 < hl, yin1 = y0 + 1 >= 0 && y0 + 1 < hl;
      const float* r0 = vol + y0 * wl + x0;
      v[l][r][0] = (yin0 && xin0) ? __ldg(r0) : 0.f;
      v[l][r][1] = (yin0 && xin1) ? __ldg(r0 + 1) : 0.f;
      v[l][r][2] = (yin1 && xin0) ? __ldg(r0 + wl) : 0.f;
      v[l][r][3] = (yin1 && xin1) ? __ldg(r0 + wl + 1) : 0.f;
      wgt[l][r][0] = wx0 * wy0; wgt[l][r][1] = wx1 * wy0; wgt[l][r][2] = wx0 * wy1; wgt[l][r][3] = wx1 * wy1;
    }
  }
  float* outq = p.out ? p.out + q * p.out_stride + p.out_coff : nullptr;
  __nv_bfloat16* outh = p.out_hl ? p.out_hl + q * p.out_stride : nullptr;
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int tap = r * 32 + lane;
      if (tap >= KK) continue;
      // same accumulation order as the generic kernel / ATen: nw, ne, sw, se, skipping out-of-range taps
      float acc = 0.f;
      acc += v[l][r][0] * wgt[l][r][0];
      acc += v[l][r][1] * wgt[l][r][1];
      acc += v[l][r][2] * wgt[l][r][2];
      acc += v[l][r][3] * wgt[l][r][3];
      acc *= mval;
      if (outq) outq[l * KK + tap] = acc;
      if (outh) {
        __nv_bfloat16 hi, lo;
        tc::split_bf16(acc, hi, lo);
        outh[l * KK + tap] = hi;
        outh[p.out_hl_plane + l * KK + tap] = lo;
      }
    }
  }
  if (outh) {
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (int c = L * KK + lane; c < p.out_stride; c += 32) { outh[c] = zero; outh[p.out_hl_plane + c] = zero; }
  }
}

// Same lookup, with each level's window region staged in shared memory first.  The 81 taps of a level touch a region of
// at most 11 x 11 texels; gathering them straight from global memory costs 4 scattered loads per tap (the LSU, not HBM, is
// the limit: ~48 warp-wide gathers per query).  Here a warp copies the bounding box of every level with row-contiguous
// loads (zeros outside the map = the reference's zeros padding), then the taps read shared memory.  Values, weights and
// accumulation order are unchanged, so the output is bit-identical to the kernel above.
constexpr int LK_RS = 12, LK_ROWS = 12;      // staged region: up to 12 x 12 texels per level
__global__ void __launch_bounds__(256) corr_lookup_l4r4_smem_kernel(const LookupParams p) {
  constexpr int R = 4, K = 9, KK = 81, L = 4, ROUNDS = 3;
  __shared__ float reg[8][L][LK_ROWS * LK_RS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  if (q >= p.nq) return;
  const int P = p.H8 * p.W8;
  const int pix = (int)(q % P);
  const int y = pix / p.W8, x = pix - y * p.W8;
  const float2 f = *reinterpret_cast<const float2*>(p.flow8 + q * 2);
  const float gx0 = __fadd_rn((float)x, f.x), gy0 = __fadd_rn((float)y, f.y);
  const float mval = p.mask ? p.mask[q] : 1.f;
  int i0[L], xlo[L], ylo[L];
  float w0[L], w1[L];
  float inv = 1.f;
#pragma unroll
  for (int l = 0; l < L; ++l, inv *= 0.5f) {
    const bool isx = lane < K;
    const int off = (isx ? lane : lane - K) - R;
    const float c = __fmul_rn(isx ? gx0 : gy0, inv);
    const float ic = lookup_coord(c, off, isx ? p.wl[l] : p.hl[l]);
    const float fl = floorf(ic);
    // clamp far-away coordinates (huge flows) so that the int conversion and the region arithmetic stay defined; a tap
    // that far outside the map reads zeros either way
    i0[l] = (int)fminf(fmaxf(fl, -65536.f), 65536.f);
    w1[l] = ic - fl;
    w0[l] = (fl + 1.f) - ic;
    // the per-axis floor coordinates increase with the window index: the region starts at index 0 of each axis
    xlo[l] = __shfl_sync(0xffffffffu, i0[l], 0);
    ylo[l] = __shfl_sync(0xffffffffu, i0[l], K);
  }
  // ---- stage the regions (coalesced along x), all loads issued before the first use
  float st[L][(LK_ROWS * LK_RS + 31) / 32];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int hl = p.hl[l], wl = p.wl[l];
    const float* vol = p.lvl[l] + q * (long long)(hl * wl);
#pragma unroll
    for (int j = 0; j < (LK_ROWS * LK_RS + 31) / 32; ++j) {
      const int idx = j * 32 + lane;
      const int r = idx / LK_RS, c = idx - r * LK_RS;
      const int yy = ylo[l] + r, xx = xlo[l] + c;
      st[l][j] = (idx < LK_ROWS * LK_RS && yy >= 0 && yy < hl && xx >= 0 && xx < wl) ? __ldg(vol + yy * wl + xx) : 0.f;
    }
  }
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < (LK_ROWS * LK_RS + 31) / 32; ++j) {
      const int idx = j * 32 + lane;
      if (idx < LK_ROWS * LK_RS) reg[wid][l][idx] = st[l][j];
    }
  __syncwarp();
  float* outq = p.out ? p.out + q * p.out_stride + p.out_coff : nullptr;
  __nv_bfloat16* outh = p.out_hl ? p.out_hl + q * p.out_stride : nullptr;
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int tap = r * 32 + lane;
      const int tc = tap < KK ? tap : KK - 1;
      const int a = tc / K, b = tc - a * K;
      const int x0 = __shfl_sync(0xffffffffu, i0[l], a), y0 = __shfl_sync(0xffffffffu, i0[l], K + b);
      const float wx0 = __shfl_sync(0xffffffffu, w0[l], a), wx1 = __shfl_sync(0xffffffffu, w1[l], a);
      const float wy0 = __shfl_sync(0xffffffffu, w0[l], K + b), wy1 = __shfl_sync(0xffffffffu, w1[l], K + b);
      if (tap >= KK) continue;
      const int rx = x0 - xlo[l], ry = y0 - ylo[l];
      float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
      if (rx >= 0 && rx + 1 < LK_RS && ry >= 0 && ry + 1 < LK_ROWS) {      // always true for finite coordinates
        const float* s0 = &reg[wid][l][ry * LK_RS + rx];
        v00 = s0[0]; v01 = s0[1]; v10 = s0[LK_RS]; v11 = s0[LK_RS + 1];
      }
      float acc = 0.f;
      acc += v00 * (wx0 * wy0);
      acc += v01 * (wx1 * wy0);
      acc += v10 * (wx0 * wy1);
      acc += v11 * (wx1 * wy1);
      acc *= mval;
      if (outq) outq[l * KK + tap] = acc;
      if (outh) {
        __nv_bfloat16 hi, lo;
        tc::split_bf16(acc, hi, lo);
        outh[l * KK + tap] = hi;
        outh[p.out_hl_plane + l * KK + tap] = lo;
      }
    }
  }
  if (outh) {
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (int c = L * KK + lane; c < p.out_stride; c += 32) { outh[c] = zero; outh[p.out_hl_plane + c] = zero; }
  }
}

// Third form of the same lookup ("lean"): identical arithmetic and staging idea as corr_lookup_l4r4_smem_kernel, but with the
// index work stripped down - the SASS of the kernel above is 55 % integer instructions (divisions by 12 and 9 per staged texel and
// per tap, 64-bit pixel decomposition, per-tap range checks) and the kernel is issue-bound.  Here: lanes stage a region as
// (row pair, 16 columns) so every address is base + immediate; the tap -> (a, b) decomposition is done once per lane for the
// three rounds; region-relative tap coordinates are clamped with one unsigned min instead of a four-way range branch; stores use
// one per-lane base pointer per plane.  Outputs are bit-identical to the two kernels above.
// TH, TW > 0: the query map size is a compile-time constant (32 x 32 = 256x256 crops), which turns every level size, row stride
// and range bound into an immediate; TH = TW = 0 reads them from the parameters.
template <int TH, int TW>
__global__ void __launch_bounds__(256) corr_lookup_l4r4_lean_kernel(const LookupParams p) {
  scf_pdl_enter();
  constexpr int R = 4, K = 9, KK = 81, L = 4, ROUNDS = 3, NIT = LK_ROWS / 2;
  __shared__ float reg[8][L][LK_ROWS * LK_RS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned q = blockIdx.x * (blockDim.x >> 5) + wid;
  if (q >= (unsigned)p.nq) return;
  const unsigned W8 = TW > 0 ? (unsigned)TW : (unsigned)p.W8, P = TH > 0 ? (unsigned)(TH * TW) : (unsigned)(p.H8 * p.W8);
  const unsigned pix = q % P;
  const unsigned y = pix / W8, x = pix - y * W8;
  const float2 f = *reinterpret_cast<const float2*>(p.flow8 + (size_t)q * 2);
  const float gx0 = __fadd_rn((float)x, f.x), gy0 = __fadd_rn((float)y, f.y);
  const float mval = p.mask ? p.mask[q] : 1.f;
  int i0[L], xlo[L], ylo[L];
  float w0[L], w1[L];
  {
    const bool isx = lane < K;
    const int off = (isx ? lane : lane - K) - R;
    float inv = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l, inv *= 0.5f) {
      const float c = __fmul_rn(isx ? gx0 : gy0, inv);
      const int wl = TW > 0 ? (TW >> l) : p.wl[l], hl = TH > 0 ? (TH >> l) : p.hl[l];
      const float ic = lookup_coord(c, off, isx ? wl : hl);
      const float fl = floorf(ic);
      i0[l] = (int)fminf(fmaxf(fl, -65536.f), 65536.f);
      w1[l] = ic - fl;
      w0[l] = (fl + 1.f) - ic;
      xlo[l] = __shfl_sync(0xffffffffu, i0[l], 0);
      ylo[l] = __shfl_sync(0xffffffffu, i0[l], K);
    }
  }
  // ---- stage the regions: lane = (row parity, column), six row pairs per level; all 24 loads issued before the first use
  const int rr = lane >> 4, cc = lane & 15;
  float st[L][NIT];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int wl = TW > 0 ? (TW >> l) : p.wl[l], hl = TH > 0 ? (TH >> l) : p.hl[l];
    const int xx = xlo[l] + cc, yy0 = ylo[l] + rr;
    const bool xin = cc < LK_RS && (unsigned)xx < (unsigned)wl;
    const float* src = p.lvl[l] + (size_t)q * (unsigned)(hl * wl) + (yy0 * wl + xx);
#pragma unroll
    for (int i = 0; i < NIT; ++i)
      st[l][i] = (xin && (unsigned)(yy0 + 2 * i) < (unsigned)hl) ? __ldg(src + 2 * i * wl) : 0.f;
  }
  if (cc < LK_RS) {
    float* dst = &reg[wid][0][rr * LK_RS + cc];
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int i = 0; i < NIT; ++i) dst[l * (LK_ROWS * LK_RS) + 2 * i * LK_RS] = st[l][i];
  }
  __syncwarp();
  // ---- taps: tap = round * 32 + lane -> (a, b), once for all levels
  int ta[ROUNDS], tb[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int tap = r * 32 + lane, tc = tap < KK ? tap : KK - 1;
    ta[r] = tc / K;
    tb[r] = K + tc - ta[r] * K;
  }
  float* outq = p.out ? p.out + (size_t)q * p.out_stride + p.out_coff + lane : nullptr;
  __nv_bfloat16* outh = p.out_hl ? p.out_hl + (size_t)q * p.out_stride + lane : nullptr;
  __nv_bfloat16* outl = outh + p.out_hl_plane;
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int x0 = __shfl_sync(0xffffffffu, i0[l], ta[r]), y0 = __shfl_sync(0xffffffffu, i0[l], tb[r]);
      const float wx0 = __shfl_sync(0xffffffffu, w0[l], ta[r]), wx1 = __shfl_sync(0xffffffffu, w1[l], ta[r]);
      const float wy0 = __shfl_sync(0xffffffffu, w0[l], tb[r]), wy1 = __shfl_sync(0xffffffffu, w1[l], tb[r]);
      // floor coordinates grow with the window index, so these are 0..10 (the min only guards the shared-memory access)
      const unsigned rx = min((unsigned)(x0 - xlo[l]), (unsigned)(LK_RS - 2)), ry = min((unsigned)(y0 - ylo[l]), (unsigned)(LK_ROWS - 2));
      const float* s0 = &reg[wid][l][ry * LK_RS + rx];
      const float v00 = s0[0], v01 = s0[1], v10 = s0[LK_RS], v11 = s0[LK_RS + 1];
      float acc = 0.f;
      acc += v00 * (wx0 * wy0);
      acc += v01 * (wx1 * wy0);
      acc += v10 * (wx0 * wy1);
      acc += v11 * (wx1 * wy1);
      acc *= mval;
      if (r < ROUNDS - 1 || lane < KK - 32 * (ROUNDS - 1)) {
        if (outq) outq[l * KK + r * 32] = acc;
        if (outh) {
          __nv_bfloat16 hi, lo;
          tc::split_bf16(acc, hi, lo);
          outh[l * KK + r * 32] = hi;
          outl[l * KK + r * 32] = lo;
        }
      }
    }
  }
  if (outh) {
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    for (int c = L * KK + lane; c < p.out_stride; c += 32) { outh[c - lane] = zero; outl[c - lane] = zero; }
  }
}

__global__ void corr_lookup_taps_kernel(int level, int radius, const float* __restrict__ flow8, int32_t* __restrict__ x0,
                                        int32_t* __restrict__ y0, int H8, int W8, long long nq) {
  const int k = 2 * radius + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nq * k) return;
  const long long q = idx / k;
  const int a = (int)(idx - q * k);
  const int pix = (int)(q % (H8 * W8));
  const int y = pix / W8, x = pix - y * W8;
  int hl = H8, wl = W8;
  float inv = 1.f;
  for (int l = 0; l < level; ++l) { hl >>= 1; wl >>= 1; inv *= 0.5f; }
  const float cx = __fmul_rn(__fadd_rn((float)x, flow8[q * 2]), inv);
  const float cy = __fmul_rn(__fadd_rn((float)y, flow8[q * 2 + 1]), inv);
  x0[idx] = (int)floorf(lookup_coord(cx, a - radius, wl));
  y0[idx] = (int)floorf(lookup_coord(cy, a - radius, hl));
}

// 2x2 mean pool with floor semantics (nn.AvgPool2d(2,2), raft_decoder.py:54-56): out[q][y][x] over in[q][2y..][2x..]
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int hi, int wi, int ho, int wo,
                                long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % wo);
    const long long r = idx / wo;
    const int yo = (int)(r % ho);
    const long long q = r / ho;
    const float* src = in + (q * hi + 2 * yo) * (long long)wi + 2 * xo;
    const float s = ((src[0] + src[1]) + src[wi]) + src[wi + 1];
    out[idx] = s * 0.25f;
  }
}

int avgpool2(const float* in, float* out, long long nq, int hi, int wi, cudaStream_t st) {
  const int ho = hi / 2, wo = wi / 2;
  const long long total = nq * ho * wo;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  avgpool2_kernel<<<blocks, 256, 0, st>>>(in, out, hi, wi, ho, wo, total);
  return check_launch("avgpool2_kernel");
}

// fp32 (CUDA-core) pyramid build: level0 = conv1x1 with per-sample "weights" feat_real ([C][P] is already the
// packed [k][n] layout), scaled by 1/sqrt(C); then successive floor pools. The tcgen05 build lives in scf_corr_tc.cu.
int corr_build_f32(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                   float* const* levels, void* scratch, cudaStream_t st) {
  const int P = H8 * W8;
  float* f1t = reinterpret_cast<float*>(scratch);  // [B,P,C]
  SCF_TRY(scf_nchw_to_nhwc(feat_render, f1t, B, C, H8, W8, C, 0, st));
  scf_conv_desc d = {};
  d.seg[0] = {f1t, C, 0, C};
  d.nseg = 1;
  d.B = B; d.Hi = d.Ho = H8; d.Wi = d.Wo = W8;
  d.kh = d.kw = 1; d.sh = d.sw = 1; d.ph = d.pw = 0;
  d.w = feat_real; d.w_batch_stride = (long long)C * P; d.ldw = P; d.cout = P;
  d.bias = nullptr; d.scale = 1.0f / sqrtf((float)C);
  d.epi = SCF_EPI_ACT; d.act = SCF_ACT_NONE;
  d.out = levels[0]; d.out_stride = P; d.out_coff = 0;
  SCF_TRY(conv2d_f32(d, st));
  int hl = H8, wl = W8;
  for (int l = 1; l < num_levels; ++l) {
    SCF_TRY(avgpool2(levels[l - 1], levels[l], (long long)B * P, hl, wl, st));
    hl /= 2; wl /= 2;
  }
  return 0;
}

}  // namespace scf

extern "C" {

size_t scf_corr_build_scratch_bytes(int B, int C, int H8, int W8) {
  // transposed feat_render fp32 [B,P,C]  (+ split-bf16 copies of both maps for the tensor-core build)
  return (size_t)B * H8 * W8 * C * 4 * 3 + 1024;
}

static int corr_lookup_impl(const float* const* h_levels, int num_levels, int radius, const float* flow8, const float* mask,
                            float* out, void* out_hl, long long plane, int out_stride, int out_coff, int B, int H8, int W8,
                            void* stream) {
  SCF_REQUIRE(h_levels && flow8 && (out || out_hl), SCF_ERR_ARG, "scf_corr_lookup: null pointer");
  SCF_REQUIRE(out_stride >= out_coff + num_levels * (2 * radius + 1) * (2 * radius + 1), SCF_ERR_ARG,
              "scf_corr_lookup: out_stride too small");
  SCF_REQUIRE(num_levels >= 1 && num_levels <= scf::kMaxLevels && radius >= 0 && radius <= 8, SCF_ERR_ARG,
              "scf_corr_lookup: num_levels 1..%d, radius 0..8", scf::kMaxLevels);
  SCF_REQUIRE(B > 0 && H8 > 0 && W8 > 0, SCF_ERR_ARG, "scf_corr_lookup: empty shape");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(flow8) % 8 == 0, SCF_ERR_ALIGN, "scf_corr_lookup: flow8 must be 8B aligned");
  scf::LookupParams p = {};
  int hl = H8, wl = W8;
  for (int l = 0; l < num_levels; ++l) {
    SCF_REQUIRE(h_levels[l] != nullptr && hl >= 1 && wl >= 1, SCF_ERR_ARG, "scf_corr_lookup: level %d missing/empty", l);
    p.lvl[l] = h_levels[l]; p.hl[l] = hl; p.wl[l] = wl;
    hl /= 2; wl /= 2;
  }
  p.num_levels = num_levels; p.radius = radius; p.flow8 = flow8; p.mask = mask; p.out = out;
  p.out_hl = reinterpret_cast<__nv_bfloat16*>(out_hl); p.out_hl_plane = plane;
  p.out_stride = out_stride; p.out_coff = out_coff; p.H8 = H8; p.W8 = W8; p.nq = (long long)B * H8 * W8;
  const int wpb = 8;
  if (num_levels == 4 && radius == 4) {
    // SCFLOW_LOOKUP_SMEM: 0 gathers from global memory, 1 shared-memory staged regions, 2 (default) the same with lean index work
    static const int staged = [] { const char* e = getenv("SCFLOW_LOOKUP_SMEM"); return e ? atoi(e) : 2; }();
    if (staged >= 2 && p.nq < (1ll << 31)) {
      if (H8 == 32 && W8 == 32) scf::launch_pdl(scf::corr_lookup_l4r4_lean_kernel<32, 32>, dim3(scf::cdiv(p.nq, wpb)), dim3(wpb * 32), 0, (cudaStream_t)stream, p);
      else scf::launch_pdl(scf::corr_lookup_l4r4_lean_kernel<0, 0>, dim3(scf::cdiv(p.nq, wpb)), dim3(wpb * 32), 0, (cudaStream_t)stream, p);
      return scf::check_launch("corr_lookup_l4r4_lean_kernel");
    }
    if (staged) {
      scf::corr_lookup_l4r4_smem_kernel<<<scf::cdiv(p.nq, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
      return scf::check_launch("corr_lookup_l4r4_smem_kernel");
    }
    scf::corr_lookup_l4r4_kernel<<<scf::cdiv(p.nq, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
    return scf::check_launch("corr_lookup_l4r4_kernel");
  }
  scf::corr_lookup_kernel<<<scf::cdiv(p.nq, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
  return scf::check_launch("corr_lookup_kernel");
}

int scf_corr_lookup(const float* const* h_levels, int num_levels, int radius, const float* flow8, const float* mask,
                    float* out, int out_stride, int out_coff, int B, int H8, int W8, void* stream) {
  return corr_lookup_impl(h_levels, num_levels, radius, flow8, mask, out, nullptr, 0, out_stride, out_coff, B, H8, W8, stream);
}

int scf_corr_lookup_split(const float* const* h_levels, int num_levels, int radius, const float* flow8, const float* mask,
                          void* out_hl, long long plane_stride, int out_stride, int B, int H8, int W8, void* stream) {
  return corr_lookup_impl(h_levels, num_levels, radius, flow8, mask, nullptr, out_hl, plane_stride, out_stride, 0, B, H8, W8,
                          stream);
}

int scf_corr_lookup_taps(int level, int radius, const float* flow8, int32_t* x0, int32_t* y0, int B, int H8, int W8,
                         void* stream) {
  SCF_REQUIRE(flow8 && x0 && y0 && level >= 0 && level < scf::kMaxLevels, SCF_ERR_ARG, "scf_corr_lookup_taps: bad args");
  const long long nq = (long long)B * H8 * W8;
  const long long total = nq * (2 * radius + 1);
  scf::corr_lookup_taps_kernel<<<scf::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(level, radius, flow8, x0, y0, H8,
                                                                                       W8, nq);
  return scf::check_launch("corr_lookup_taps_kernel");
}

}  // extern "C"
