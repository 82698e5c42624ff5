// CorrLookup fused with the motion encoder's first convolution (corr_lookup.py:102-136 + raft_decoder.py:152-155, corr_net[0]):
//     c1[q, :] = relu(W (256 x 324) * lookup(pyramid, flow8)[q, :] + b)
// One kernel: the 4-level 9x9 bilinear gather of a query never reaches HBM - it is written straight into shared memory as the
// split-bf16 A operand of the 1x1 convolution, which runs on tcgen05 in the same CTA (fp32 accumulation in TMEM), and only the
// 256-channel result (split-bf16, the next convolution's operand) is stored.  Saves the 1.34 MB / sample / iteration round trip
// of the 324-channel tensor and one launch per iteration.
//
// CTA = 128 consecutive queries (TMEM lanes).  Per pyramid level (81 taps, padded to 96 channels = 3 chunks of 32):
//   * 16 gather warps, 8 queries each: the warp replays the reference's fp32 coordinate sequence per axis (bit-exact neighbour
//     indices, same code as scf_corr.cu), copies the <= 12 x 12 texel region of the query's level map into shared memory with
//     cp.async (zero fill outside the map = zeros padding; four region buffers per warp: two queries are evaluated while the next two are in flight),
//     evaluates its 81 taps (same products, same order as the stand-alone lookup) and stores hi / lo bf16 into the level's A tile
//     (K-major, SWIZZLE_64B: round r of the taps = chunk r, lane = channel);
//   * the MMA warp contracts the level's A tile with the level's weight chunks (TMA, two 32 KB stages) while the gather warps fill
//     the other A buffer with the next level: 4 levels x 3 chunks x 2 k-steps x 3 products = 72 MMAs of M 128, N 256.
// Epilogue: bias + ReLU, split-bf16, [32 queries][32 channels] blocks through shared memory and TMA stores.
#include "scf_common.cuh"
#include "scf_tc.cuh"
#include <mutex>
#include <stdlib.h>

namespace scf {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapDataType dtype, CUtensorMapSwizzle swz);

constexpr int LC_TQ = 128, LC_R = 4, LC_K = 9, LC_KK = 81, LC_L = 4, LC_CHL = 96, LC_GW = 16, LC_COUT = 256;
constexpr int LC_RS = 12;                                           // staged region: 12 x 12 texels
constexpr uint32_t LC_A_PLANE = LC_TQ * 64, LC_A_CHUNK = 2 * LC_A_PLANE, LC_A_LEVEL = 3 * LC_A_CHUNK;      // 8 / 16 / 48 KB
constexpr uint32_t LC_B_PLANE = LC_COUT * 64, LC_B_STAGE = 2 * LC_B_PLANE;                                 // 16 / 32 KB
constexpr int LC_B_STAGES = 2;
constexpr int LC_DEPTH = 4;                                         // region buffers per gather warp: copies run 3 queries ahead
constexpr uint32_t LC_REG = LC_DEPTH * 640;                         // (576 B used of each)
constexpr int LC_SMEM = 1024 + 2048 + 2 * (int)LC_A_LEVEL + LC_B_STAGES * (int)LC_B_STAGE + LC_GW * (int)LC_REG;
static_assert(LC_SMEM <= 232448, "lookup_conv_kernel does not fit in shared memory");

struct LookupConvParams {
  const float* lvl[LC_L];
  int hl[LC_L], wl[LC_L];
  const float* flow8;       // [nq][2]
  const float* mask;        // optional [nq]
  const float* bias;        // [256]
  int H8, W8;
  int num_tiles;
};

// un-normalised sample coordinate of one axis, bit-exact replay of the reference's fp32 sequence (same as scf_corr.cu)
__device__ __forceinline__ float lc_coord(float centre, int off, int size) {
  const float p = __fadd_rn(centre, (float)off);
  const float den = (float)(size - 1 > 1 ? size - 1 : 1);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

__global__ void __launch_bounds__(64 + 32 * LC_GW, 1)
lookup_conv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const LookupConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_bfull = smem_base, bar_bempty = smem_base + 16, bar_afull = smem_base + 32, bar_aempty = smem_base + 48,
                 bar_tfull = smem_base + 64, tmem_slot = smem_base + 96;
  const uint32_t bias_s = smem_base + 1024;                       // 256 floats
  const uint32_t a0 = smem_base + 1024 + 2048;
  const uint32_t b0 = a0 + 2 * LC_A_LEVEL;
  const uint32_t reg0 = b0 + LC_B_STAGES * LC_B_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW); prefetch_tmap(&tmO);
    for (int s = 0; s < LC_B_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_afull + 8 * s, LC_GW); mbar_init(bar_aempty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256u);
  // bias -> shared memory (a weight: not written by the preceding kernel, so it may be read before griddepcontrol.wait)
  if (threadIdx.x >= 64 && threadIdx.x < 64 + LC_COUT)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * (threadIdx.x - 64)), "f"(p.bias ? __ldg(p.bias + threadIdx.x - 64) : 0.f) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ================= weights: per tile 4 levels x 3 chunks of [256 rows][32 ch] x 2 planes
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x)
        for (int c = 0; c < LC_L * 3; ++c) {
          mbar_wait(bar_bempty + 8 * stage, phase ^ 1u);
          mbar_arrive_expect_tx(bar_bfull + 8 * stage, LC_B_STAGE);
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(b0 + stage * LC_B_STAGE), "l"(reinterpret_cast<uint64_t>(&tmW)), "r"(bar_bfull + 8 * stage), "r"(c * 32), "r"(0), "r"(0)
                       : "memory");
          if (++stage == LC_B_STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer
      const uint32_t idesc = make_idesc_bf16(LC_TQ, LC_COUT);
      int stage = 0;
      uint32_t phase = 0, use = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        for (int l = 0; l < LC_L; ++l, ++use) {
          const uint32_t buf = use & 1u;
          mbar_wait(bar_afull + 8 * buf, (use >> 1) & 1u);
          tc_fence_after();
          for (int c = 0; c < 3; ++c) {
            mbar_wait(bar_bfull + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_addr = a0 + buf * LC_A_LEVEL + (uint32_t)c * LC_A_CHUNK, b_addr = b0 + stage * LC_B_STAGE;
            const uint64_t a_hi = make_smem_desc_sw64(a_addr, 512), a_lo = make_smem_desc_sw64(a_addr + LC_A_PLANE, 512);
            const uint64_t b_hi = make_smem_desc_sw64(b_addr, 512), b_lo = make_smem_desc_sw64(b_addr + LC_B_PLANE, 512);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t ko = (uint64_t)(k * 32 >> 4);
              umma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, (l > 0 || c > 0 || k > 0) ? 1u : 0u);
              umma_bf16(tmem_base, a_hi + ko, b_lo + ko, idesc, 1u);
              umma_bf16(tmem_base, a_lo + ko, b_hi + ko, idesc, 1u);
            }
            umma_commit(bar_bempty + 8 * stage);
            if (++stage == LC_B_STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_aempty + 8 * buf);
        }
        umma_commit(bar_tfull);
      }
    }
  } else {
    // ================= gather warps (also the epilogue)
    const int gw = warp - 2;                                      // 0..15
    const int P = p.H8 * p.W8;
    const uint32_t reg = reg0 + (uint32_t)gw * LC_REG;
    // per-lane constants: the (a, b) window offsets of this lane's taps in the three rounds, the texels it stages
    int ta[3], tb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int tap = r * 32 + lane, tc_ = tap < LC_KK ? tap : LC_KK - 1;
      ta[r] = tc_ / LC_K; tb[r] = tc_ - ta[r] * LC_K;
    }
    int sr[5], scx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int idx = k * 32 + lane;
      sr[k] = idx / LC_RS; scx[k] = idx - sr[k] * LC_RS;
    }
    const uint32_t swz_lane = ((uint32_t)lane & 7u) * 2u;
    uint32_t use = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const long long q0 = (long long)t * LC_TQ + gw * 8;          // this warp's first query
      // lane j < 8 holds query j's centre (x + flow_x, y + flow_y) and mask value: loaded once per tile, off the per-level loops
      float gx_l = 0.f, gy_l = 0.f, m_l = 1.f;
      if (lane < 8) {
        const long long q = q0 + lane;
        const int pix = (int)(q % P);
        const int y = pix / p.W8, x = pix - y * p.W8;
        const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow8) + q);
        gx_l = __fadd_rn((float)x, f.x);
        gy_l = __fadd_rn((float)y, f.y);
        if (p.mask) m_l = __ldg(p.mask + q);
      }
      for (int l = 0; l < LC_L; ++l, ++use) {
        const uint32_t buf = use & 1u;
        const int hl = p.hl[l], wl = p.wl[l];
        const float inv = 1.f / (float)(1 << l);                  // exact
        const uint32_t a_lvl = a0 + buf * LC_A_LEVEL;
        // per-axis coordinates of query j (lanes 0..8: x offsets, 9..17: y offsets) and its region copy, LC_DEPTH - 1 queries ahead
        int ci0[LC_DEPTH];
        float cw0[LC_DEPTH], cw1[LC_DEPTH];
        auto prepare = [&](int j, int slot) {
          const float gx0 = __shfl_sync(0xffffffffu, gx_l, j), gy0 = __shfl_sync(0xffffffffu, gy_l, j);
          const bool isx = lane < LC_K;
          const int off = (isx ? lane : lane - LC_K) - LC_R;
          const float c = __fmul_rn(isx ? gx0 : gy0, inv);
          const float ic = lc_coord(c, off, isx ? wl : hl);
          const float fl = floorf(ic);
          ci0[slot] = (int)fminf(fmaxf(fl, -65536.f), 65536.f);
          cw1[slot] = ic - fl;
          cw0[slot] = (fl + 1.f) - ic;
          const int xlo = __shfl_sync(0xffffffffu, ci0[slot], 0), ylo = __shfl_sync(0xffffffffu, ci0[slot], LC_K);
          const float* vol = p.lvl[l] + (q0 + j) * (long long)(hl * wl);
          const uint32_t dst = reg + (uint32_t)slot * 640u;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            if (k * 32 + lane < LC_RS * LC_RS) {
              const int yy = ylo + sr[k], xx = xlo + scx[k];
              const bool ok = yy >= 0 && yy < hl && xx >= 0 && xx < wl;
              const float* src = ok ? vol + yy * wl + xx : vol;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (uint32_t)(k * 32 + lane) * 4u), "l"(src), "r"(ok ? 4 : 0) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int j = 0; j < LC_DEPTH; ++j) prepare(j, j);          // all four region buffers in flight
        mbar_wait(bar_aempty + 8 * buf, ((use >> 1) & 1u) ^ 1u);  // the MMAs that read this A buffer two levels ago are done
        // Two queries per step: their (independent) shuffle -> shared-memory load -> FMA -> convert -> store chains interleave, which
        // is what keeps the issue slots busy with only four warps per scheduler.
        const float* regf = reinterpret_cast<const float*>(smem_raw) + ((reg - smem_u32(smem_raw)) >> 2);      // generic view of the region buffers
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          // queries j, j + 1 are complete when at most the two younger groups (j + 2, j + 3) are pending
          if (j + 3 < 8) asm volatile("cp.async.wait_group 2;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          float accs[2][3];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int slot = (j + u) % LC_DEPTH;
            const int i0 = ci0[slot];
            const float w0 = cw0[slot], w1 = cw1[slot];
            const int xlo = __shfl_sync(0xffffffffu, i0, 0), ylo = __shfl_sync(0xffffffffu, i0, LC_K);
            const float mval = __shfl_sync(0xffffffffu, m_l, j + u);
            const float* rs = regf + slot * 160;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int x0 = __shfl_sync(0xffffffffu, i0, ta[r]), y0 = __shfl_sync(0xffffffffu, i0, LC_K + tb[r]);
              const float wx0 = __shfl_sync(0xffffffffu, w0, ta[r]), wx1 = __shfl_sync(0xffffffffu, w1, ta[r]);
              const float wy0 = __shfl_sync(0xffffffffu, w0, LC_K + tb[r]), wy1 = __shfl_sync(0xffffffffu, w1, LC_K + tb[r]);
              float acc = 0.f;
              if (r * 32 + lane < LC_KK) {
                const int rx = x0 - xlo, ry = y0 - ylo;
                float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
                if (rx >= 0 && rx + 1 < LC_RS && ry >= 0 && ry + 1 < LC_RS) {      // always true for finite coordinates
                  const float* s0 = rs + ry * LC_RS + rx;
                  v00 = s0[0]; v01 = s0[1]; v10 = s0[LC_RS]; v11 = s0[LC_RS + 1];
                }
                // same products and accumulation order as the stand-alone lookup / ATen: nw, ne, sw, se
                acc += v00 * (wx0 * wy0);
                acc += v01 * (wx1 * wy0);
                acc += v10 * (wx0 * wy1);
                acc += v11 * (wx1 * wy1);
                acc *= mval;
              }
              accs[u][r] = acc;
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t row = (uint32_t)(gw * 8 + j + u);
            const uint32_t row_off = row * 64u + ((((uint32_t)lane >> 3) ^ ((row >> 1) & 3u)) << 4) + swz_lane;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const float acc = accs[u][r];
              const __nv_bfloat16 hi = __float2bfloat16_rn(acc);
              const __nv_bfloat16 lo = __float2bfloat16_rn(acc - __bfloat162float(hi));
              const uint32_t dsta = a_lvl + (uint32_t)r * LC_A_CHUNK + row_off;
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(dsta), "h"(__bfloat16_as_ushort(hi)) : "memory");
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(dsta + LC_A_PLANE), "h"(__bfloat16_as_ushort(lo)) : "memory");
            }
          }
          __syncwarp();                 // every lane has read the two region buffers: refill them (queries j + 4, j + 5)
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (j + u + LC_DEPTH < 8) prepare(j + u + LC_DEPTH, (j + u) % LC_DEPTH);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull + 8 * buf);
      }
      // ---------------- epilogue: bias + ReLU, split-bf16, TMA stores.  Warp -> TMEM lane quarter (warp & 3), 64 columns each
      {
        const int q = warp & 3, cg = gw >> 2;
        mbar_wait(bar_tfull, (uint32_t)it & 1u);
        tc_fence_after();
        const uint32_t stg = a0 + (uint32_t)gw * 4096u;           // staging inside the (idle) A buffers: [2 planes][32 rows][32 ch]
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 64);
        const uint32_t srow = (uint32_t)lane * 64u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          float v[32];
          __syncwarp();
          tmem_ld32(t_addr + (uint32_t)(s * 32), v);
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {                           // 8 channels = one 16 B unit per plane
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cidx = cg * 64 + s * 32 + j * 8 + 2 * e;
              float b0v, b1v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b0v) : "r"(bias_s + 4u * cidx));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b1v) : "r"(bias_s + 4u * cidx + 4u));
              const float x0 = fmaxf(v[j * 8 + 2 * e] + b0v, 0.f), x1 = fmaxf(v[j * 8 + 2 * e + 1] + b1v, 0.f);
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
              const float2 hf = __bfloat1622float2(h2);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
              hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
              lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            const uint32_t unit = (((uint32_t)j ^ (((uint32_t)lane >> 1) & 3u)) << 4);       // SWIZZLE_64B
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + srow + unit), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + 2048u + srow + unit), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(&tmO), "r"(stg), "r"(cg * 64 + s * 32), "r"(t * LC_TQ + q * 32), "r"(0) : "memory");
            bulk_commit_group();
          }
        }
        if (lane == 0) bulk_wait_group_read0();                   // the staging area is the next tile's A operand
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LC_GW) : "memory");   // every warp's staging has been read and TMEM drained
      }
    }
    if (lane == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256u);
}

// Measured on B200 (B = 32, profiles/r02_summary.md): 91 us per launch against 50 us (stand-alone lookup, 64 resident warps per
// SM, 82 % issue-slot utilisation) + 27 us (1x1 convolution) for the two-kernel form.  The gather is instruction-issue bound
// (~1300 warp instructions per query), not HBM-bound - its 43 MB output stays L2-resident between the two launches - and inside
// a 186 KB-shared-memory CTA only 16 gather warps fit on an SM, which halves the issue rate.  The decoder therefore uses this
// kernel only on request (SCFLOW_LOOKUP_FUSED=1); the C entry point scf_lookup_conv is always available.
bool lookup_conv_requested() {
  const char* e = getenv("SCFLOW_LOOKUP_FUSED");
  return e ? atoi(e) != 0 : false;
}

bool lookup_conv_ok(int num_levels, int radius, int B, int H8, int W8) {
  return  num_levels == LC_L && radius == LC_R && ((long long)B * H8 * W8) % LC_TQ == 0 && (H8 >> 3) >= 1 && (W8 >> 3) >= 1;
}

// w_lk: packed bf16 [2][256][4 * 96] (level l's 81 input channels at columns 96 l .., zero padded); out_hl: split-bf16 [2][nq][stride]
int lookup_conv_fused(const float* const* levels, const float* flow8, const float* mask, const void* w_lk, const float* bias,
                      void* out_hl, long long out_plane, int out_stride, int B, int H8, int W8, cudaStream_t st) {
  SCF_REQUIRE(levels && flow8 && w_lk && out_hl, SCF_ERR_ARG, "scf_lookup_conv: null pointer");
  SCF_REQUIRE(lookup_conv_ok(LC_L, LC_R, B, H8, W8), SCF_ERR_UNSUPPORTED, "scf_lookup_conv: needs 4 levels, radius 4 and B*H8*W8 %% 128 == 0");
  SCF_REQUIRE(out_stride >= LC_COUT && out_stride % 8 == 0 && (out_plane * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(out_hl) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(w_lk) % 16 == 0 && reinterpret_cast<uintptr_t>(flow8) % 8 == 0,
              SCF_ERR_ALIGN, "scf_lookup_conv: alignment");
  LookupConvParams p = {};
  int hl = H8, wl = W8;
  for (int l = 0; l < LC_L; ++l) {
    SCF_REQUIRE(levels[l] != nullptr, SCF_ERR_ARG, "scf_lookup_conv: level %d null", l);
    p.lvl[l] = levels[l]; p.hl[l] = hl; p.wl[l] = wl;
    hl /= 2; wl /= 2;
  }
  const long long nq = (long long)B * H8 * W8;
  p.flow8 = flow8; p.mask = mask; p.bias = bias; p.H8 = H8; p.W8 = W8; p.num_tiles = (int)(nq / LC_TQ);
  CUtensorMap tmW, tmO;
  {
    cuuint64_t dims[3] = {(cuuint64_t)LC_L * LC_CHL, LC_COUT, 2};
    cuuint64_t str[2] = {(cuuint64_t)LC_L * LC_CHL * 2, (cuuint64_t)LC_COUT * LC_L * LC_CHL * 2};
    cuuint32_t box[3] = {32, LC_COUT, 2};
    SCF_TRY(encode_map(&tmW, w_lk, 3, dims, str, box, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
    cuuint64_t od[3] = {(cuuint64_t)LC_COUT, (cuuint64_t)nq, 2};
    cuuint64_t os[2] = {(cuuint64_t)out_stride * 2, (cuuint64_t)out_plane * 2};
    cuuint32_t ob[3] = {32, 32, 2};
    SCF_TRY(encode_map(&tmO, out_hl, 3, od, os, ob, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    SCF_CUDA(cudaGetDevice(&dev));
    SCF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] { attr_err = cudaFuncSetAttribute(lookup_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LC_SMEM); });
  SCF_REQUIRE(attr_err == cudaSuccess, (int)attr_err, "cudaFuncSetAttribute(lookup_conv_kernel): %s", cudaGetErrorString(attr_err));
  static const bool pdl = [] { const char* e = getenv("SCFLOW_PDL"); return e ? atoi(e) != 0 : true; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_tiles < num_sms ? p.num_tiles : num_sms); cfg.blockDim = dim3(64 + 32 * LC_GW);
  cfg.dynamicSmemBytes = LC_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, lookup_conv_kernel, tmW, tmO, p);
  if (le != cudaSuccess) { cudaGetLastError(); set_error("lookup_conv_kernel launch: %s", cudaGetErrorString(le)); g_launches++; return (int)le; }
  return check_launch("lookup_conv_kernel");
}

int pack_conv_weight_tc_range(const float* w_oihw, void* packed, int O, int I_total, int i_begin, int i_count, int i_dst, int kh,
                              int kw, int cin_pad, int cout_pad, int o_off, cudaStream_t st);

}  // namespace scf

extern "C" {

size_t scf_lookup_conv_packed_bytes(void) { return (size_t)2 * scf::LC_COUT * scf::LC_L * scf::LC_CHL * 2; }

int scf_lookup_conv_pack(const float* w_oihw, void* packed, void* stream) {
  using namespace scf;
  SCF_REQUIRE(w_oihw && packed, SCF_ERR_ARG, "scf_lookup_conv_pack: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  SCF_CUDA(cudaMemsetAsync(packed, 0, scf_lookup_conv_packed_bytes(), st));
  for (int l = 0; l < LC_L; ++l)
    SCF_TRY(pack_conv_weight_tc_range(w_oihw, packed, LC_COUT, LC_L * LC_KK, l * LC_KK, LC_KK, l * LC_CHL, 1, 1, LC_L * LC_CHL, LC_COUT, 0, st));
  return 0;
}

int scf_lookup_conv(const float* const* h_levels, const float* flow8, const float* mask, const void* packed_w, const float* bias,
                    void* out_hl, long long out_plane_stride, int out_stride, int B, int H8, int W8, void* stream) {
  return scf::lookup_conv_fused(h_levels, flow8, mask, packed_w, bias, out_hl, out_plane_stride, out_stride, B, H8, W8, (cudaStream_t)stream);
}

}  // extern "C"
