// Whole-loop driver: SCFlowDecoder.forward (models/decoder/scflow_decoder.py:150-251) as one C call that enqueues
// every kernel of the iterative refinement on one stream.  No host synchronisation, no allocation: the caller
// hands in a packed-weight arena (scf_decoder_pack) and a workspace (scf_decoder_workspace_bytes), so the whole
// call is CUDA-graph capturable.
#include "scf_common.cuh"
#include <cuda_bf16.h>
#include <string.h>
#include <stdlib.h>

namespace scf {

int conv2d_f32(const scf_conv_desc& d, cudaStream_t st);
int conv2d_tc(const scf_tc_conv_desc& d, cudaStream_t st);
int avgpool2(const float* in, float* out, long long nq, int hi, int wi, cudaStream_t st);
int pack_conv_weight_tc_range(const float* w_oihw, void* packed, int O, int I_total, int i_begin, int i_count, int i_dst, int kh,
                              int kw, int cin_pad, int cout_pad, int o_off, cudaStream_t st);
int heads_predict(const void* hd_hl, long long plane, int stride, int hidden, const float* wf, int ldwf, const float* bf,
                  const float* wm, int ldwm, const float* bm, float* dflow, float* mask8, int B, int H, int W, cudaStream_t st);
int im2col_x_split(const float* in, int nchw, int cin, int kw, void* out_hl, long long plane, int N, int H, int Wi, int sx,
                   cudaStream_t st);
int pack_conv_weight_tc_foldx(const float* w_oihw, float* scratch, void* packed, int O, int C, int KH, int KW, int cin_pad,
                              int cout_pad, int o_off, cudaStream_t st);
int corr_build_dispatch(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                        float* const* levels, void* scratch, int precision, cudaStream_t st);
int gru_pass_fused(const scf_gru_pass_desc& d, cudaStream_t st);
bool lookup_conv_ok(int num_levels, int radius, int B, int H8, int W8);
bool lookup_conv_requested();
int group_norm_relu_partials(float* x, int nsplit, long long split_stride, const float* gamma, const float* beta, int B, int HW, int C,
                             int num_groups, float eps, void* out_hl, long long plane_stride, cudaStream_t stream);
int fc_tc(const float* x, int x_nsplit, long long x_split_stride, const float* x_bias, int x_relu, const void* w_packed, float* part,
          float* y, const float* bias, int relu, int B, int I, int O, int KR, cudaStream_t st);
int pose_project_partials(const float* part, int nsplit, long long split_stride, const float* bias, const float* rot_w, const float* rot_b,
                          const float* tr_w, const float* tr_b, const int64_t* label, float* d_rot, float* d_trs, int B, int I,
                          int rot_dim, int num_class, cudaStream_t st);
int pose_fc_tail(const float* x, const float* w0, const float* b0, int I0, int O0, const float* w1, const float* b1, int O1,
                 const float* rot_w, const float* rot_b, const float* tr_w, const float* tr_b, const int64_t* label, float* d_rot,
                 float* d_trs, int B, int rot_dim, int num_class, float* part0, float* part1, int ks0, int ks1, cudaStream_t st);
int pack_predict_tc(const float* wf_oihw, const float* wm_oihw, void* packed, int hidden, int cout_pad, cudaStream_t st);
int predict_gather(const float* d, int ld, const float* bf, const float* bm, float* dflow, float* mask8, int B, int H, int W,
                   cudaStream_t st);
int lookup_conv_fused(const float* const* levels, const float* flow8, const float* mask, const void* w_lk, const float* bias,
                      void* out_hl, long long out_plane, int out_stride, int B, int H8, int W8, cudaStream_t st);
int corr_build_presplit(void* scratch, int B, int C, int H8, int W8, int num_levels, float* const* levels, cudaStream_t st);

thread_local char g_err[512] = {0};
thread_local long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------ side stream for work off the critical path
// The pose regressor is a chain of small, latency-bound kernels that leaves most SMs idle; the x8 up-sampling of the flow /
// mask outputs of the same iteration does not depend on it, so it runs on a library-owned side stream (fork / join through
// events: capturable into the caller's CUDA graph, still no host synchronisation).
struct SideStream { cudaStream_t s; cudaEvent_t fork, join, join_a, join_b; int dev; bool ok; };
static thread_local SideStream g_side = {nullptr, nullptr, nullptr, nullptr, nullptr, -1, false};
// SCFLOW_DEC_OVERLAP bit mask: 1 = output up-sampling, 2 = flow branch of the motion encoder, 4 = mask encoder
static int overlap_mask() {
  static const int m = [] { const char* e = getenv("SCFLOW_DEC_OVERLAP"); return e ? atoi(e) : 7; }();
  return m;
}
static SideStream* side_stream() {
  if (!overlap_mask()) return nullptr;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (g_side.ok && g_side.dev == dev) return &g_side;
  SideStream n = {nullptr, nullptr, nullptr, nullptr, nullptr, dev, false};
  if (cudaStreamCreateWithFlags(&n.s, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&n.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n.join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n.join_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n.join_b, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  n.ok = true;
  g_side = n;          // (a previous device's stream is deliberately leaked: one process drives one GPU)
  return &g_side;
}

// ------------------------------------------------------------------ packed convolution table
enum PC {
  PC_CORR0, PC_CORR1, PC_FLOW0, PC_FLOW1, PC_OUT0, PC_ZR0, PC_Q0, PC_ZR1, PC_Q1,
  PC_CZR0, PC_CQ0, PC_CZR1, PC_CQ1,   // tensor-core only: the context-feature columns of the GRU convolutions (loop invariant)
  PC_HEADS, PC_FHP, PC_MHP,
  PC_DFE0, PC_DFE1, PC_ME0, PC_ME1, PC_PH0, PC_PH1, PC_PH2, PC_COUNT
};
struct PCInfo {
  int cin, kh, kw, cout;      // cout = total (merged) output channels
  int nsrc;                   // 1 or 2 reference convs merged along O
  int src_w[2], src_b[2];     // scf_decoder_weight indices (bias -1 = none)
  int src_cout[2];
  int ldw;
  size_t w_off, b_off;        // float offsets into the arena
  // tensor-core copy (precision 1): bf16 [2][taps][cout_pad][cin_pad] at byte offset tc_off (0 = layer stays on fp32 cores)
  int tc, cin_pad, cout_pad;
  size_t tc_off;
  int tc_only;                // no fp32 copy (exists for the tensor-core path only)
  int tc_nr, tc_src[2], tc_cnt[2];   // input-channel ranges of the source weight that the tensor-core copy holds, packed densely
  int foldx;                  // thin input: the tensor-core copy is the kh x 1 convolution over kx*cin + c channels (x-im2col)
  int tc_kh, tc_kw;           // kernel size the tensor-core convolution runs with
};

struct Arena {
  PCInfo pc[PC_COUNT];
  size_t gn_w[3], gn_b[3];
  size_t fc0_w, fc0_b, fc1_w, fc1_b, rot_w, rot_b, tr_w, tr_b;
  size_t fold_scratch;        // floats: staging of a folded thin-input weight while packing
  size_t pd_off;              // bytes: the two predict layers as one 1x1 convolution 512 -> 19 (tap-wise partial products), bf16 [2][32][512]
  size_t fc_off[2];           // bytes: fc0 / fc1 as split-bf16 [2][O][I] for the tcgen05 FC layers (0 = none)
  size_t lk_off;              // bytes: corr_net[0] repacked per pyramid level for the fused lookup + convolution (0 = none)
  size_t total_floats;
  size_t total_bytes;         // fp32 section + tensor-core section
  int nc, rot_rows, tr_rows;
};

static void build_arena(const scf_decoder_cfg& cfg, Arena& a) {
  const int L = cfg.num_levels, k = 2 * cfg.radius + 1;
  const int corr_ch = L * k * k;
  auto set = [&](int id, int cin, int kh, int kw, int w0, int b0, int c0, int w1 = -1, int b1 = -1, int c1 = 0) {
    PCInfo& p = a.pc[id];
    p.cin = cin; p.kh = kh; p.kw = kw; p.nsrc = w1 >= 0 ? 2 : 1;
    p.src_w[0] = w0; p.src_b[0] = b0; p.src_cout[0] = c0;
    p.src_w[1] = w1; p.src_b[1] = b1; p.src_cout[1] = c1;
    p.cout = c0 + c1;
    p.ldw = (p.cout + 3) / 4 * 4;
    p.tc_only = 0; p.tc_nr = 1; p.tc_src[0] = 0; p.tc_cnt[0] = cin; p.tc_src[1] = p.tc_cnt[1] = 0;
    p.foldx = 0; p.tc_kh = kh; p.tc_kw = kw;
  };
  set(PC_CORR0, corr_ch, 1, 1, SCF_W_CORR0_W, SCF_W_CORR0_B, 256);
  set(PC_CORR1, 256, 3, 3, SCF_W_CORR1_W, SCF_W_CORR1_B, 192);
  set(PC_FLOW0, 2, 7, 7, SCF_W_FLOW0_W, SCF_W_FLOW0_B, 128);
  set(PC_FLOW1, 128, 3, 3, SCF_W_FLOW1_W, SCF_W_FLOW1_B, 64);
  set(PC_OUT0, 256, 3, 3, SCF_W_OUT0_W, SCF_W_OUT0_B, 126);
  set(PC_ZR0, 384, 1, 5, SCF_W_GRU_Z0_W, SCF_W_GRU_Z0_B, 128, SCF_W_GRU_R0_W, SCF_W_GRU_R0_B, 128);
  set(PC_Q0, 384, 1, 5, SCF_W_GRU_Q0_W, SCF_W_GRU_Q0_B, 128);
  set(PC_ZR1, 384, 5, 1, SCF_W_GRU_Z1_W, SCF_W_GRU_Z1_B, 128, SCF_W_GRU_R1_W, SCF_W_GRU_R1_B, 128);
  set(PC_Q1, 384, 5, 1, SCF_W_GRU_Q1_W, SCF_W_GRU_Q1_B, 128);
  // precision 1: the GRU input is [h | cxt | motion]; the cxt columns never change inside the loop, so their contribution
  // (plus the bias) is evaluated once per forward (PC_C*) and the per-iteration convolutions keep [h | motion] only
  set(PC_CZR0, 384, 1, 5, SCF_W_GRU_Z0_W, SCF_W_GRU_Z0_B, 128, SCF_W_GRU_R0_W, SCF_W_GRU_R0_B, 128);
  set(PC_CQ0, 384, 1, 5, SCF_W_GRU_Q0_W, SCF_W_GRU_Q0_B, 128);
  set(PC_CZR1, 384, 5, 1, SCF_W_GRU_Z1_W, SCF_W_GRU_Z1_B, 128, SCF_W_GRU_R1_W, SCF_W_GRU_R1_B, 128);
  set(PC_CQ1, 384, 5, 1, SCF_W_GRU_Q1_W, SCF_W_GRU_Q1_B, 128);
  for (int id : {PC_CZR0, PC_CQ0, PC_CZR1, PC_CQ1}) { a.pc[id].tc_only = 1; a.pc[id].tc_src[0] = 128; a.pc[id].tc_cnt[0] = 128; }
  if (cfg.precision == 1)
    for (int id : {PC_ZR0, PC_Q0, PC_ZR1, PC_Q1}) {
      PCInfo& p = a.pc[id];
      p.tc_nr = 2; p.tc_src[0] = 0; p.tc_cnt[0] = 128; p.tc_src[1] = 256; p.tc_cnt[1] = 128;
    }
  set(PC_HEADS, 128, 3, 3, SCF_W_FH0_W, SCF_W_FH0_B, 256, SCF_W_MH0_W, SCF_W_MH0_B, 256);
  set(PC_FHP, 256, 3, 3, SCF_W_FHP_W, SCF_W_FHP_B, 2);
  set(PC_MHP, 256, 1, 1, SCF_W_MHP_W, SCF_W_MHP_B, 1);
  set(PC_DFE0, 2, 7, 7, SCF_W_DFE0_W, SCF_W_DFE0_B, 128);
  set(PC_DFE1, 128, 3, 3, SCF_W_DFE1_W, SCF_W_DFE1_B, 64);
  set(PC_ME0, 1, 3, 3, SCF_W_ME0_W, SCF_W_ME0_B, 64);
  set(PC_ME1, 64, 3, 3, SCF_W_ME1_W, SCF_W_ME1_B, 32);
  set(PC_PH0, 224, 3, 3, SCF_W_PH_C0_W, -1, 128);
  set(PC_PH1, 128, 3, 3, SCF_W_PH_C1_W, -1, 128);
  set(PC_PH2, 128, 3, 3, SCF_W_PH_C2_W, -1, 128);
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };   // 256B-aligned slots
  for (int i = 0; i < PC_COUNT; ++i) {
    PCInfo& p = a.pc[i];
    p.w_off = take(p.tc_only ? 0 : (size_t)p.kh * p.kw * p.cin * p.ldw);
    p.b_off = take(p.ldw);
  }
  for (int i = 0; i < 3; ++i) { a.gn_w[i] = take(128); a.gn_b[i] = take(128); }
  a.nc = cfg.num_class > 0 ? cfg.num_class : 1;
  a.rot_rows = cfg.rot_dim * a.nc;
  a.tr_rows = 3 * a.nc;
  a.fc0_w = take((size_t)1024 * 2048); a.fc0_b = take(1024);
  a.fc1_w = take((size_t)256 * 1024); a.fc1_b = take(256);
  a.rot_w = take((size_t)a.rot_rows * 256); a.rot_b = take(a.rot_rows);
  a.tr_w = take((size_t)a.tr_rows * 256); a.tr_b = take(a.tr_rows);
  a.fold_scratch = take((size_t)128 * 2 * 7 * 7);
  a.total_floats = off;
  // tensor-core section: every stride-1 convolution with >= 8 input channels
  size_t boff = (off * 4 + 1023) / 1024 * 1024;
  for (int i = 0; i < PC_COUNT; ++i) {
    PCInfo& p = a.pc[i];
    p.tc = (cfg.precision == 1 && p.cin >= 8) ? 1 : 0;
    p.cin_pad = (p.tc_cnt[0] + p.tc_cnt[1] + 7) / 8 * 8;
    if (cfg.precision == 1 && (i == PC_FLOW0 || i == PC_DFE0)) {     // 7x7 over 2 channels -> 7x1 over 14 (padded to 16)
      p.tc = 1; p.foldx = 1; p.tc_kh = p.kh; p.tc_kw = 1;
      p.cin_pad = (p.kw * p.cin + 7) / 8 * 8;
    }
    p.cout_pad = (p.cout + 15) / 16 * 16;
    p.tc_off = 0;
    if (p.tc) {
      p.tc_off = boff;
      boff += ((size_t)2 * p.tc_kh * p.tc_kw * p.cout_pad * p.cin_pad * 2 + 1023) / 1024 * 1024;
    }
  }
  a.pd_off = 0;
  if (cfg.precision == 1) {
    a.pd_off = boff;
    boff += ((size_t)2 * 32 * 512 * 2 + 1023) / 1024 * 1024;
  }
  a.fc_off[0] = a.fc_off[1] = 0;
  if (cfg.precision == 1 && cfg.pose_head) {
    a.fc_off[0] = boff; boff += (size_t)2 * 1024 * 2048 * 2;
    a.fc_off[1] = boff; boff += (size_t)2 * 256 * 1024 * 2;
  }
  a.lk_off = 0;
  if (cfg.precision == 1 && cfg.num_levels == 4 && cfg.radius == 4) {
    a.lk_off = boff;
    boff += (scf_lookup_conv_packed_bytes() + 1023) / 1024 * 1024;
  }
  a.total_bytes = boff;
}

// FC0 consumes the flattened NCHW map (index c*16 + pix, pose_head.py:203); our map is NHWC (pix*128 + c):
// permute the columns once at pack time.
__global__ void permute_fc0_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int C, int PIX) {
  const long long total = (long long)O * C * PIX;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    const int pix = (int)(r % PIX);
    const long long o = r / PIX;
    out[idx] = w[(o * C + c) * PIX + pix];
  }
}

__global__ void fill_kernel(float* p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void mul_mask_kernel(const float* __restrict__ flow8, const float* __restrict__ mask, float* __restrict__ out,
                                long long npix) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const float m = mask[i];
    out[2 * i] = flow8[2 * i] * m;
    out[2 * i + 1] = flow8[2 * i + 1] * m;
  }
}
__global__ void identity_delta_kernel(float* d_rot, float* d_trs, int B, int rot_dim) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < rot_dim; ++i) d_rot[b * rot_dim + i] = (i == 0 || i == 4) ? 1.f : 0.f;   // ortho6d identity
  d_trs[b * 3] = d_trs[b * 3 + 1] = d_trs[b * 3 + 2] = 0.f;
}

// ------------------------------------------------------------------ workspace
struct Workspace {
  size_t corr_scratch, lvl[8], pts4, flow8, flow8b, flowm, maskprev, corr, c1, cf, f1, h[2], cxt, motion, z, rh, hd, dflow,
      mask8, pd, df1, df2, mf1, mf2, p1, p2, p3, fc0, fc1;
  // precision 1: split-bf16 planes [2][B*P][C] (byte offsets) and their plane strides in elements
  size_t s_corr, s_c1, s_cf, s_f1, s_h[2], s_cxt, s_motion, s_rh, s_hd, s_df1, s_mf1, s_df2, s_mf2, s_p1, s_p2;
  size_t s_t7;                  // x-folded 2-channel flow map [2][B*P][16] feeding the 7x1 tensor-core form of the 7x7 flow encoders
  size_t pre_zr[2], pre_q[2];   // fp32 [B*P][256] / [B*P][128]: context contribution + bias of the GRU convolutions, per pass
  int corr_stride_s;
  int hl[8], wl[8];
  int corr_stride;
  size_t total_bytes;
};

static void build_workspace(const scf_decoder_cfg& cfg, int B, int H, int W, Workspace& w) {
  const int scale = 1 << (cfg.num_levels - 1);
  const int H8 = H / scale, W8 = W / scale;
  const size_t P = (size_t)H8 * W8, BP = (size_t)B * P;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  w.corr_scratch = take(scf_corr_build_scratch_bytes(B, 256, H8, W8));
  int hl = H8, wl = W8;
  for (int l = 0; l < cfg.num_levels; ++l) {
    w.hl[l] = hl; w.wl[l] = wl;
    w.lvl[l] = take(BP * (size_t)hl * wl * 4);
    hl /= 2; wl /= 2;
  }
  const int k = 2 * cfg.radius + 1;
  w.corr_stride = (cfg.num_levels * k * k + 3) / 4 * 4;
  w.pts4 = take((size_t)B * H * W * 16);
  w.flow8 = take(BP * 8); w.flow8b = take(BP * 8); w.flowm = take(BP * 8); w.maskprev = take(BP * 4);
  w.corr = take(BP * w.corr_stride * 4);
  w.c1 = take(BP * 256 * 4); w.cf = take(BP * 256 * 4); w.f1 = take(BP * 128 * 4);
  w.h[0] = take(BP * 128 * 4); w.h[1] = take(BP * 128 * 4);
  w.cxt = take(BP * 128 * 4); w.motion = take(BP * 128 * 4);
  w.z = take(BP * 128 * 4); w.rh = take(BP * 128 * 4);
  w.hd = take(BP * 512 * 4); w.dflow = take(BP * 8); w.mask8 = take(BP * 4); w.pd = take(BP * 32 * 4);
  w.df1 = take(BP * 128 * 4); w.df2 = take(BP * 64 * 4); w.mf1 = take(BP * 64 * 4); w.mf2 = take(BP * 32 * 4);
  // pose-head maps: up to 9 partial maps each (tap split of the stride-2 convolutions)
  w.p1 = take(9 * (BP / 4 * 128 * 4 + 1024)); w.p2 = take(9 * (BP / 16 * 128 * 4 + 1024)); w.p3 = take(9 * (BP / 64 * 128 * 4 + 1024));
  w.fc0 = take((size_t)16 * B * 1024 * 4); w.fc1 = take((size_t)16 * B * 256 * 4);     // up to 16 split-K partial maps
  w.corr_stride_s = (cfg.num_levels * k * k + 7) / 8 * 8;
  if (cfg.precision == 1) {
    auto split = [&](int ch) { return take(BP * ch * 2 * 2); };
    w.s_corr = split(w.corr_stride_s); w.s_c1 = split(256); w.s_cf = split(256); w.s_f1 = split(128);
    w.s_h[0] = split(128); w.s_h[1] = split(128); w.s_cxt = split(128); w.s_motion = split(128); w.s_rh = split(128);
    w.s_hd = split(512); w.s_df1 = split(128); w.s_mf1 = split(64); w.s_df2 = split(64); w.s_mf2 = split(32);
    w.s_p1 = take((BP / 4 + 64) * 128 * 2 * 2); w.s_p2 = take((BP / 16 + 64) * 128 * 2 * 2);
    for (int i = 0; i < 2; ++i) { w.pre_zr[i] = take(BP * 256 * 4); w.pre_q[i] = take(BP * 128 * 4); }
    w.s_t7 = split(16);
  }
  w.total_bytes = off;
}

static int check_cfg(const scf_decoder_cfg* cfg) {
  SCF_REQUIRE(cfg != nullptr, SCF_ERR_ARG, "scf_decoder: null cfg");
  SCF_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= 6, SCF_ERR_ARG, "scf_decoder: num_levels must be 1..6");
  SCF_REQUIRE(cfg->radius >= 0 && cfg->radius <= 8, SCF_ERR_ARG, "scf_decoder: radius must be 0..8");
  SCF_REQUIRE(cfg->rot_dim == 6, SCF_ERR_UNSUPPORTED, "scf_decoder: only rotation_mode='ortho6d' (rot_dim 6) is implemented");
  SCF_REQUIRE(cfg->precision == 0 || cfg->precision == 1, SCF_ERR_ARG, "scf_decoder: precision must be 0 or 1");
  return 0;
}

}  // namespace scf

using namespace scf;

extern "C" {

int scf_abi_version(void) { return SCF_ABI_VERSION; }
int scf_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(scf_conv_desc);
    case 1: return (int)sizeof(scf_tc_conv_desc);
    case 2: return (int)sizeof(scf_decoder_cfg);
    case 3: return (int)sizeof(scf_decoder_io);
    case 4: return (int)sizeof(scf_encoder_out);
    case 5: return (int)sizeof(scf_loss_desc);
    case 6: return (int)sizeof(scf_gru_pass_desc);
    default: return -1;
  }
}
const char* scf_last_error(void) { return scf::g_err; }

int scf_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

long long scf_launch_counter(void) { return scf::g_launches; }

size_t scf_decoder_packed_bytes(const scf_decoder_cfg* cfg) {
  if (check_cfg(cfg) != 0) return 0;
  Arena a;
  build_arena(*cfg, a);
  return a.total_bytes;
}

size_t scf_decoder_workspace_bytes(const scf_decoder_cfg* cfg, int B, int H, int W) {
  if (check_cfg(cfg) != 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  Workspace w;
  build_workspace(*cfg, B, H, W, w);
  return w.total_bytes;
}

int scf_decoder_workspace_slots(const scf_decoder_cfg* cfg, int B, int H, int W, size_t* slots4) {
  SCF_TRY(check_cfg(cfg));
  SCF_REQUIRE(slots4 && B > 0 && H > 0 && W > 0 && cfg->precision == 1, SCF_ERR_ARG, "scf_decoder_workspace_slots: bad args (precision 1 only)");
  Workspace w;
  build_workspace(*cfg, B, H, W, w);
  slots4[0] = w.corr_scratch; slots4[1] = w.s_h[0]; slots4[2] = w.h[0]; slots4[3] = w.s_cxt;
  return 0;
}

int scf_decoder_pack(const scf_decoder_cfg* cfg, const float* const* h_weights, void* packed, void* stream) {
  SCF_TRY(check_cfg(cfg));
  SCF_REQUIRE(h_weights && packed, SCF_ERR_ARG, "scf_decoder_pack: null pointer");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 256 == 0, SCF_ERR_ALIGN, "scf_decoder_pack: arena must be 256B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  Arena a;
  build_arena(*cfg, a);
  float* base = reinterpret_cast<float*>(packed);
  SCF_CUDA(cudaMemsetAsync(packed, 0, a.total_bytes, st));
  for (int i = 0; i < PC_COUNT; ++i) {
    const PCInfo& p = a.pc[i];
    if (!cfg->pose_head && i >= PC_PH0) continue;
    int o_off = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      SCF_REQUIRE(h_weights[p.src_w[s]] != nullptr, SCF_ERR_ARG, "scf_decoder_pack: weight %d is null", p.src_w[s]);
      if (!p.tc_only)
        SCF_TRY(scf_pack_conv_weight(h_weights[p.src_w[s]], base + p.w_off, p.src_cout[s], p.cin, p.kh, p.kw, p.ldw, o_off, st));
      if (p.tc && p.foldx) {
        SCF_TRY(pack_conv_weight_tc_foldx(h_weights[p.src_w[s]], base + a.fold_scratch, reinterpret_cast<char*>(packed) + p.tc_off,
                                          p.src_cout[s], p.cin, p.kh, p.kw, p.cin_pad, p.cout_pad, o_off, st));
      } else if (p.tc) {
        int i_dst = 0;
        for (int r = 0; r < p.tc_nr; ++r) {
          SCF_TRY(pack_conv_weight_tc_range(h_weights[p.src_w[s]], reinterpret_cast<char*>(packed) + p.tc_off, p.src_cout[s], p.cin,
                                            p.tc_src[r], p.tc_cnt[r], i_dst, p.kh, p.kw, p.cin_pad, p.cout_pad, o_off, st));
          i_dst += p.tc_cnt[r];
        }
      }
      if (p.src_b[s] >= 0) {
        SCF_REQUIRE(h_weights[p.src_b[s]] != nullptr, SCF_ERR_ARG, "scf_decoder_pack: bias %d is null", p.src_b[s]);
        SCF_CUDA(cudaMemcpyAsync(base + p.b_off + o_off, h_weights[p.src_b[s]], (size_t)p.src_cout[s] * 4,
                                 cudaMemcpyDeviceToDevice, st));
      }
      o_off += p.src_cout[s];
    }
  }
  if (a.pd_off) SCF_TRY(pack_predict_tc(h_weights[SCF_W_FHP_W], h_weights[SCF_W_MHP_W], reinterpret_cast<char*>(packed) + a.pd_off, 256, 32, st));
  if (a.lk_off) SCF_TRY(scf_lookup_conv_pack(h_weights[SCF_W_CORR0_W], reinterpret_cast<char*>(packed) + a.lk_off, st));
  if (cfg->pose_head) {
    const int gw[3] = {SCF_W_PH_G0_W, SCF_W_PH_G1_W, SCF_W_PH_G2_W}, gb[3] = {SCF_W_PH_G0_B, SCF_W_PH_G1_B, SCF_W_PH_G2_B};
    for (int i = 0; i < 3; ++i) {
      SCF_REQUIRE(h_weights[gw[i]] && h_weights[gb[i]], SCF_ERR_ARG, "scf_decoder_pack: GroupNorm params missing");
      SCF_CUDA(cudaMemcpyAsync(base + a.gn_w[i], h_weights[gw[i]], 128 * 4, cudaMemcpyDeviceToDevice, st));
      SCF_CUDA(cudaMemcpyAsync(base + a.gn_b[i], h_weights[gb[i]], 128 * 4, cudaMemcpyDeviceToDevice, st));
    }
    for (int i = SCF_W_PH_FC0_W; i <= SCF_W_PH_TR_B; ++i)
      SCF_REQUIRE(h_weights[i] != nullptr, SCF_ERR_ARG, "scf_decoder_pack: pose-head FC weight %d is null", i);
    permute_fc0_kernel<<<1024, 256, 0, st>>>(h_weights[SCF_W_PH_FC0_W], base + a.fc0_w, 1024, 128, 16);
    SCF_TRY(check_launch("permute_fc0_kernel"));
    SCF_CUDA(cudaMemcpyAsync(base + a.fc0_b, h_weights[SCF_W_PH_FC0_B], 1024 * 4, cudaMemcpyDeviceToDevice, st));
    SCF_CUDA(cudaMemcpyAsync(base + a.fc1_w, h_weights[SCF_W_PH_FC1_W], (size_t)256 * 1024 * 4, cudaMemcpyDeviceToDevice, st));
    SCF_CUDA(cudaMemcpyAsync(base + a.fc1_b, h_weights[SCF_W_PH_FC1_B], 256 * 4, cudaMemcpyDeviceToDevice, st));
    if (a.fc_off[0]) {
      // the tcgen05 FC layers read split-bf16 [2][O][I] copies (fc0 with its columns already permuted to the NHWC flatten)
      char* pb = reinterpret_cast<char*>(packed);
      SCF_TRY(scf_pack_conv_weight_tc(base + a.fc0_w, pb + a.fc_off[0], 1024, 2048, 1, 1, 2048, 1024, 0, st));
      SCF_TRY(scf_pack_conv_weight_tc(base + a.fc1_w, pb + a.fc_off[1], 256, 1024, 1, 1, 1024, 256, 0, st));
    }
    SCF_CUDA(cudaMemcpyAsync(base + a.rot_w, h_weights[SCF_W_PH_ROT_W], (size_t)a.rot_rows * 256 * 4, cudaMemcpyDeviceToDevice, st));
    SCF_CUDA(cudaMemcpyAsync(base + a.rot_b, h_weights[SCF_W_PH_ROT_B], (size_t)a.rot_rows * 4, cudaMemcpyDeviceToDevice, st));
    SCF_CUDA(cudaMemcpyAsync(base + a.tr_w, h_weights[SCF_W_PH_TR_W], (size_t)a.tr_rows * 256 * 4, cudaMemcpyDeviceToDevice, st));
    SCF_CUDA(cudaMemcpyAsync(base + a.tr_b, h_weights[SCF_W_PH_TR_B], (size_t)a.tr_rows * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int scf_corr_build(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                   float* const* h_levels, void* scratch, int precision, void* stream) {
  SCF_REQUIRE(feat_render && feat_real && h_levels && scratch, SCF_ERR_ARG, "scf_corr_build: null pointer");
  SCF_REQUIRE(B > 0 && C > 0 && H8 > 0 && W8 > 0 && num_levels >= 1 && num_levels <= 6, SCF_ERR_ARG, "scf_corr_build: bad shape");
  SCF_REQUIRE(C % 4 == 0 && (H8 * W8) % 4 == 0, SCF_ERR_UNSUPPORTED, "scf_corr_build: C and H8*W8 must be multiples of 4");
  for (int l = 0; l < num_levels; ++l) SCF_REQUIRE(h_levels[l] != nullptr, SCF_ERR_ARG, "scf_corr_build: level %d null", l);
  return corr_build_dispatch(feat_render, feat_real, B, C, H8, W8, num_levels, h_levels, scratch, precision,
                             (cudaStream_t)stream);
}

int scf_decoder_forward(const scf_decoder_cfg* cfg, const void* packed, const scf_decoder_io* io, int B, int H, int W,
                        int iters, void* workspace, size_t workspace_bytes, void* stream) {
  SCF_TRY(check_cfg(cfg));
  SCF_REQUIRE(packed && io && workspace, SCF_ERR_ARG, "scf_decoder_forward: null pointer");
  SCF_REQUIRE(B > 0 && iters > 0, SCF_ERR_ARG, "scf_decoder_forward: B and iters must be positive");
  const int scale = 1 << (cfg->num_levels - 1);
  SCF_REQUIRE(H % scale == 0 && W % scale == 0 && H >= 2 * scale && W >= 2 * scale, SCF_ERR_ARG,
              "scf_decoder_forward: H, W must be multiples of %d", scale);
  SCF_REQUIRE((io->native_inputs || (io->feat_render && io->feat_real && io->h_feat && io->cxt_feat)) && io->ref_rotation &&
                  io->ref_translation && io->depth && io->internel_k && io->init_flow,
              SCF_ERR_ARG, "scf_decoder_forward: null input");
  SCF_REQUIRE(io->flow_from_pose && io->flow_from_pred && io->rotation && io->translation && io->mask &&
                  io->delta_rotation && io->delta_translation,
              SCF_ERR_ARG, "scf_decoder_forward: null output");
  SCF_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0 && reinterpret_cast<uintptr_t>(packed) % 256 == 0,
              SCF_ERR_ALIGN, "scf_decoder_forward: workspace / packed arena must be 256B aligned");
  const int H8 = H / scale, W8 = W / scale, P = H8 * W8;
  if (cfg->pose_head) {
    SCF_REQUIRE(H8 == 32 && W8 == 32, SCF_ERR_UNSUPPORTED,
                "scf_decoder_forward: the pose head's FC expects a 32x32 feature map (pose_head.py:147-167); got %dx%d", H8, W8);
    SCF_REQUIRE(cfg->num_class <= 0 || io->label != nullptr, SCF_ERR_ARG, "scf_decoder_forward: label required");
  }
  Arena a;
  build_arena(*cfg, a);
  Workspace ws;
  build_workspace(*cfg, B, H, W, ws);
  SCF_REQUIRE(workspace_bytes >= ws.total_bytes, SCF_ERR_ARG, "scf_decoder_forward: workspace too small (%zu < %zu)",
              workspace_bytes, ws.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char* wsb = reinterpret_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(wsb + off); };
  const float* pw = reinterpret_cast<const float*>(packed);
  const long long HW = (long long)H * W;
  const size_t BP = (size_t)B * P;

  // ---- once per forward: pyramid, point map, layout conversion of h / context
  float* levels[8];
  for (int l = 0; l < cfg->num_levels; ++l) levels[l] = F(ws.lvl[l]);
  const bool native = io->native_inputs != 0;
  SCF_REQUIRE(!native || cfg->precision == 1, SCF_ERR_ARG, "scf_decoder_forward: native_inputs needs precision 1");
  if (native) SCF_TRY(corr_build_presplit(wsb + ws.corr_scratch, B, 256, H8, W8, cfg->num_levels, levels, st));
  else SCF_TRY(corr_build_dispatch(io->feat_render, io->feat_real, B, 256, H8, W8, cfg->num_levels, levels, wsb + ws.corr_scratch,
                                   cfg->precision, st));
  SCF_TRY(scf_unproject(io->depth, io->internel_k, io->ref_rotation, io->ref_translation, F(ws.pts4), B, H, W, st));
  const bool tcp = cfg->precision == 1;
  auto S = [&](size_t off) { return reinterpret_cast<void*>(wsb + off); };
  if (native) {
    // h (split + fp32) and the context map were written into their workspace slots by scf_encoder_forward_ex
  } else if (tcp) {
    SCF_TRY(scf_nchw_to_nhwc_split(io->h_feat, S(ws.s_h[0]), (long long)BP * 128, 128, 0, F(ws.h[0]), 128, B, 128, H8, W8, st));
    SCF_TRY(scf_nchw_to_nhwc_split(io->cxt_feat, S(ws.s_cxt), (long long)BP * 128, 128, 0, nullptr, 0, B, 128, H8, W8, st));
  } else {
    SCF_TRY(scf_nchw_to_nhwc(io->h_feat, F(ws.h[0]), B, 128, H8, W8, 128, 0, st));
    SCF_TRY(scf_nchw_to_nhwc(io->cxt_feat, F(ws.cxt), B, 128, H8, W8, 128, 0, st));
  }
  if (cfg->mask_corr || cfg->mask_flow) {
    fill_kernel<<<cdiv(BP, 256), 256, 0, st>>>(F(ws.maskprev), 1.f, (long long)BP);
    SCF_TRY(check_launch("fill_kernel"));
  }

  cudaStream_t lst = st;       // stream the conv / convtc helpers launch on (switched to the side stream for parallel branches)
  SideStream* const side_all = cfg->pose_head && tcp ? side_stream() : nullptr;
  const int ovl = side_all ? overlap_mask() : 0;
  auto conv = [&](int id, std::initializer_list<scf_conv_seg> segs, int Hi, int Wi, int Ho, int Wo, int stride, int act,
                  float* out, int out_stride, int out_coff, int epi = SCF_EPI_ACT, const float* aux0 = nullptr,
                  const float* aux1 = nullptr, float* out2 = nullptr, void* out_hl = nullptr, int out_hl_stride = 0) -> int {
    const PCInfo& p = a.pc[id];
    scf_conv_desc d = {};
    int n = 0;
    for (const scf_conv_seg& s : segs) d.seg[n++] = s;
    d.nseg = n;
    d.B = B; d.Hi = Hi; d.Wi = Wi; d.Ho = Ho; d.Wo = Wo;
    d.kh = p.kh; d.kw = p.kw; d.sh = d.sw = stride; d.ph = p.kh / 2; d.pw = p.kw / 2;
    d.w = pw + p.w_off; d.w_batch_stride = 0; d.ldw = p.ldw; d.cout = p.cout;
    d.bias = p.src_b[0] >= 0 ? pw + p.b_off : nullptr;
    d.scale = 1.f; d.epi = epi; d.act = act;
    d.out = out; d.out_stride = out_stride; d.out_coff = out_coff;
    d.aux0 = aux0; d.aux0_stride = 128; d.aux1 = aux1; d.aux1_stride = 128; d.out2 = out2; d.out2_stride = 128;
    d.out_hl = out_hl; d.out_hl_plane = (long long)BP * out_hl_stride; d.out_hl_stride = out_hl_stride; d.out_hl_coff = 0;
    return conv2d_f32(d, lst);
  };
  // tensor-core convolution on split-bf16 buffers: segs = {plane base, channels per pixel, first channel, channels}
  struct SSeg { void* ptr; int stride, coff, nch; };
  int tc_hin = H8, tc_win = W8, tc_stride = 1;      // geometry of the next convtc call (pose head overrides it)
  int pad_writable = 0;
  int tc_ksplit = 1;                                // tap split of the next convtc call (pose head)
  long long tc_split_stride = 0;
  auto convtc = [&](int id, std::initializer_list<SSeg> segs, int act, float* out_f32, int f32_stride, void* out_hl,
                    int hl_stride, int hl_coff, int epi = SCF_EPI_ACT, const float* aux0 = nullptr, const float* aux1 = nullptr,
                    void* out2_hl = nullptr, const float* pre = nullptr, int pre_stride = 0) -> int {
    const PCInfo& p = a.pc[id];
    scf_tc_conv_desc d = {};
    int n = 0;
    const long long in_pix = (long long)B * tc_hin * tc_win;
    for (const SSeg& sg : segs) { d.seg[n].ptr = sg.ptr; d.seg[n].plane_stride = in_pix * sg.stride; d.seg[n].stride = sg.stride;
                                  d.seg[n].coff = sg.coff; d.seg[n].nch = sg.nch; ++n; }
    d.nseg = n;
    d.B = B; d.H = tc_hin; d.W = tc_win; d.kh = p.tc_kh; d.kw = p.tc_kw; d.stride = tc_stride;
    d.w = reinterpret_cast<const char*>(packed) + p.tc_off; d.cin_pad = p.cin_pad; d.cout_pad = p.cout_pad; d.cout = p.cout;
    d.w_batched = 0;
    d.bias = p.src_b[0] >= 0 ? pw + p.b_off : nullptr;
    d.scale = 1.f; d.epi = epi; d.act = act;
    d.out_f32 = out_f32; d.out_f32_stride = f32_stride; d.out_f32_coff = 0;
    d.out_hl = out_hl; d.out_hl_plane = (long long)BP * hl_stride; d.out_hl_stride = hl_stride; d.out_hl_coff = hl_coff;
    d.aux0 = aux0; d.aux0_stride = 128; d.aux1 = aux1; d.aux1_stride = 128;
    d.out2_hl = out2_hl; d.out2_hl_plane = (long long)BP * 128; d.out2_hl_stride = 128;
    if (pre) { d.pre = pre; d.pre_stride = pre_stride; d.bias = nullptr; }   // the bias is part of the precomputed map
    d.out_pad_writable = pad_writable;
    d.ksplit = tc_ksplit; d.split_stride = tc_split_stride;
    return conv2d_tc(d, lst);
  };

  if (tcp) {
    // loop-invariant part of the SepConvGRU: W[:, cxt columns] * cxt + bias for z|r and q of both passes (raft_decoder.py:235-253)
    for (int pass = 0; pass < 2; ++pass) {
      SCF_TRY(convtc(pass == 0 ? PC_CZR0 : PC_CZR1, {{S(ws.s_cxt), 128, 0, 128}}, SCF_ACT_NONE, F(ws.pre_zr[pass]), 256, nullptr, 0, 0));
      SCF_TRY(convtc(pass == 0 ? PC_CQ0 : PC_CQ1, {{S(ws.s_cxt), 128, 0, 128}}, SCF_ACT_NONE, F(ws.pre_q[pass]), 128, nullptr, 0, 0));
    }
  }
  const size_t Btot = io->out_batch_total > 0 ? (size_t)io->out_batch_total : (size_t)B;
  const size_t boff = io->out_batch_total > 0 ? (size_t)io->out_batch_offset : 0;
  SCF_REQUIRE(io->out_batch_offset >= 0 && boff + B <= Btot, SCF_ERR_ARG, "scf_decoder_forward: output batch window out of range");
  const float* flow_full = io->init_flow;
  for (int it = 0; it < iters; ++it) {
    const size_t ob = (size_t)it * Btot + boff;          // first output sample of this call in iteration `it`
    float* flow_pose_k = io->flow_from_pose + ob * 2 * HW;
    float* flow_pred_k = io->flow_from_pred + ob * 2 * HW;
    float* mask_k = io->mask + ob * HW;
    float* rot_k = io->rotation + ob * 9;
    float* trs_k = io->translation + ob * 3;
    float* drot_k = io->delta_rotation + ob * cfg->rot_dim;
    float* dtrs_k = io->delta_translation + ob * 3;
    const float* rot_prev = it == 0 ? io->ref_rotation : io->rotation + (ob - Btot) * 9;
    const float* trs_prev = it == 0 ? io->ref_translation : io->translation + (ob - Btot) * 3;

    // flow8 = 1/8 * down8(flow)                                          (scflow_decoder.py:196-197)
    // iteration 0 resamples the caller's init_flow; later iterations receive flow8 from the previous iteration's re-projection
    // launch (two buffers: the side stream's x8 up-sampling of iteration `it` still reads flow8[it] while flow8[it+1] is written)
    float* const flow8 = F((it & 1) ? ws.flow8b : ws.flow8);
    float* const flow8_next = F((it & 1) ? ws.flow8 : ws.flow8b);
    if (it == 0)
      SCF_TRY(scf_resize_bilinear(flow_full, nullptr, 2 * HW, HW, W, 1, H, W, flow8, (long long)P * 2, 1,
                                  (long long)W8 * 2, 2, H8, W8, B, 2, 1.0f / scale, st));
    const float* menc_flow = flow8;
    if (cfg->mask_flow) {
      mul_mask_kernel<<<cdiv(BP, 256), 256, 0, st>>>(flow8, F(ws.maskprev), F(ws.flowm), (long long)BP);
      SCF_TRY(check_launch("mul_mask_kernel"));
      menc_flow = F(ws.flowm);
    }
    if (tcp) {
      // ---------------- tensor-core path: activations live as split-bf16 planes
      // flow branch of the motion encoder (flow_net) on the side stream, parallel to the lookup + corr_net branch
      if (ovl & 2) {
        SCF_CUDA(cudaEventRecord(side_all->fork, st));
        SCF_CUDA(cudaStreamWaitEvent(side_all->s, side_all->fork, 0));
        lst = side_all->s;
      }
      SCF_TRY(im2col_x_split(menc_flow, 0, 2, 7, S(ws.s_t7), (long long)BP * 16, B, H8, W8, 1, lst));
      SCF_TRY(convtc(PC_FLOW0, {{S(ws.s_t7), 16, 0, 16}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_f1), 128, 0));
      SCF_TRY(convtc(PC_FLOW1, {{S(ws.s_f1), 128, 0, 128}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_cf), 256, 192));
      if (ovl & 2) {
        SCF_CUDA(cudaEventRecord(side_all->join_a, side_all->s));
        lst = st;
      }
      if (a.lk_off && lookup_conv_requested() && lookup_conv_ok(cfg->num_levels, cfg->radius, B, H8, W8)) {
        // pyramid lookup + corr_net[0] (1x1, 324 -> 256, ReLU) in one kernel: the 324-channel tensor never reaches HBM
        SCF_TRY(lookup_conv_fused(levels, flow8, cfg->mask_corr ? F(ws.maskprev) : nullptr, reinterpret_cast<const char*>(packed) + a.lk_off,
                                  pw + a.pc[PC_CORR0].b_off, S(ws.s_c1), (long long)BP * 256, 256, B, H8, W8, st));
      } else {
        SCF_TRY(scf_corr_lookup_split(levels, cfg->num_levels, cfg->radius, flow8, cfg->mask_corr ? F(ws.maskprev) : nullptr,
                                      S(ws.s_corr), (long long)BP * ws.corr_stride_s, ws.corr_stride_s, B, H8, W8, st));
        SCF_TRY(convtc(PC_CORR0, {{S(ws.s_corr), ws.corr_stride_s, 0, ws.corr_stride_s}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_c1), 256, 0));
      }
      SCF_TRY(convtc(PC_CORR1, {{S(ws.s_c1), 256, 0, 256}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_cf), 256, 0));
      if (ovl & 2) SCF_CUDA(cudaStreamWaitEvent(st, side_all->join_a, 0));
      // motion features = [conv output (126) | flow (2)] (raft_decoder.py MotionEncoder: cat([out, flow])); the flow channels
      // are written after the convolution, which may clear its output's padding channels (out_pad_writable)
      pad_writable = 1;
      SCF_TRY(convtc(PC_OUT0, {{S(ws.s_cf), 256, 0, 256}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_motion), 128, 0));
      pad_writable = 0;
      SCF_TRY(scf_split_copy(menc_flow, 2, 0, S(ws.s_motion), (long long)BP * 128, 128, 126, (long long)BP, 2, lst));
      // SepConvGRU: one kernel per pass where the map allows whole-row / whole-column tiles (SCFLOW_GRU_FUSED=0: two kernels)
      static const bool gru_fused = [] { const char* e = getenv("SCFLOW_GRU_FUSED"); return e ? atoi(e) != 0 : true; }();
      for (int pass = 0; pass < 2; ++pass) {
        if (gru_fused && H8 == 32 && W8 == 32) {
          scf_gru_pass_desc g = {};
          g.h_hl = S(ws.s_h[pass]); g.h_plane = (long long)BP * 128; g.h_f32 = F(ws.h[pass]);
          g.m_hl = S(ws.s_motion); g.m_plane = (long long)BP * 128;
          g.w_zr = reinterpret_cast<const char*>(packed) + a.pc[pass == 0 ? PC_ZR0 : PC_ZR1].tc_off;
          g.w_q = reinterpret_cast<const char*>(packed) + a.pc[pass == 0 ? PC_Q0 : PC_Q1].tc_off;
          g.pre_zr = F(ws.pre_zr[pass]); g.pre_q = F(ws.pre_q[pass]);
          g.out_f32 = F(ws.h[pass ^ 1]); g.out_hl = S(ws.s_h[pass ^ 1]); g.out_plane = (long long)BP * 128;
          g.B = B; g.H = H8; g.W = W8; g.vertical = pass;
          SCF_TRY(gru_pass_fused(g, lst));
          continue;
        }
        SCF_TRY(convtc(pass == 0 ? PC_ZR0 : PC_ZR1, {{S(ws.s_h[pass]), 128, 0, 128}, {S(ws.s_motion), 128, 0, 128}},
                       SCF_ACT_SIGMOID, F(ws.z), 128, nullptr, 0, 0, SCF_EPI_GRU_ZR, F(ws.h[pass]), nullptr, S(ws.s_rh),
                       F(ws.pre_zr[pass]), 256));
        SCF_TRY(convtc(pass == 0 ? PC_Q0 : PC_Q1, {{S(ws.s_rh), 128, 0, 128}, {S(ws.s_motion), 128, 0, 128}},
                       SCF_ACT_TANH, F(ws.h[pass ^ 1]), 128, S(ws.s_h[pass ^ 1]), 128, 0, SCF_EPI_GRU_Q, F(ws.h[pass]), F(ws.z), nullptr,
                       F(ws.pre_q[pass]), 128));
      }
      SCF_TRY(convtc(PC_HEADS, {{S(ws.s_h[0]), 128, 0, 128}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_hd), 512, 0));
      // both predict layers (3x3 256->2, 1x1 256->1 + sigmoid) in one streaming fp32 kernel over the split hidden map
      static const bool predict_tc = [] { const char* e = getenv("SCFLOW_PREDICT_TC"); return e ? atoi(e) != 0 : false; }();
      static const bool predict_pd = [] { const char* e = getenv("SCFLOW_PREDICT_PD"); return e ? atoi(e) != 0 : true; }();
      if (predict_tc) {   // the two predict layers as tensor-core convolutions (N padded to 16; halo mode for the 3x3)
        SCF_TRY(convtc(PC_FHP, {{S(ws.s_hd), 512, 0, 256}}, SCF_ACT_NONE, F(ws.dflow), 2, nullptr, 0, 0));
        SCF_TRY(convtc(PC_MHP, {{S(ws.s_hd), 512, 256, 256}}, SCF_ACT_SIGMOID, F(ws.mask8), 1, nullptr, 0, 0));
      } else if (predict_pd && a.pd_off) {
        // both predict layers: one 1x1 tensor-core convolution 512 -> 19 tap-wise partial products + a 9-tap gather (scf_predict.cu)
        scf_tc_conv_desc d = {};
        d.seg[0].ptr = S(ws.s_hd); d.seg[0].plane_stride = (long long)BP * 512; d.seg[0].stride = 512; d.seg[0].coff = 0; d.seg[0].nch = 512;
        d.nseg = 1; d.B = B; d.H = H8; d.W = W8; d.kh = d.kw = 1; d.stride = 1;
        d.w = reinterpret_cast<const char*>(packed) + a.pd_off; d.cin_pad = 512; d.cout_pad = 32; d.cout = 32;
        d.scale = 1.f; d.epi = SCF_EPI_ACT; d.act = SCF_ACT_NONE;
        d.out_f32 = F(ws.pd); d.out_f32_stride = 32; d.out_f32_coff = 0;
        SCF_TRY(conv2d_tc(d, lst));
        SCF_TRY(predict_gather(F(ws.pd), 32, pw + a.pc[PC_FHP].b_off, pw + a.pc[PC_MHP].b_off, F(ws.dflow), F(ws.mask8), B, H8, W8, st));
      } else
      SCF_TRY(heads_predict(S(ws.s_hd), (long long)BP * 512, 512, 256, pw + a.pc[PC_FHP].w_off, a.pc[PC_FHP].ldw, pw + a.pc[PC_FHP].b_off,
                            pw + a.pc[PC_MHP].w_off, a.pc[PC_MHP].ldw, pw + a.pc[PC_MHP].b_off, F(ws.dflow), F(ws.mask8), B, H8, W8, st));
      if (cfg->pose_head) {
        // mask encoder on the side stream, parallel to the delta-flow encoder
        if (ovl & 4) {
          SCF_CUDA(cudaEventRecord(side_all->fork, st));
          SCF_CUDA(cudaStreamWaitEvent(side_all->s, side_all->fork, 0));
          lst = side_all->s;
        }
        SCF_TRY(conv(PC_ME0, {{F(ws.mask8), 1, 0, 1}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, nullptr, 64, 0, SCF_EPI_ACT, nullptr, nullptr,
                     nullptr, S(ws.s_mf1), 64));
        SCF_TRY(convtc(PC_ME1, {{S(ws.s_mf1), 64, 0, 64}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_mf2), 32, 0));
        if (ovl & 4) {
          SCF_CUDA(cudaEventRecord(side_all->join_b, side_all->s));
          lst = st;
        }
        SCF_TRY(im2col_x_split(F(ws.dflow), 0, 2, 7, S(ws.s_t7), (long long)BP * 16, B, H8, W8, 1, st));
        SCF_TRY(convtc(PC_DFE0, {{S(ws.s_t7), 16, 0, 16}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_df1), 128, 0));
        SCF_TRY(convtc(PC_DFE1, {{S(ws.s_df1), 128, 0, 128}}, SCF_ACT_RELU, nullptr, 0, S(ws.s_df2), 64, 0));
        if (ovl & 4) SCF_CUDA(cudaStreamWaitEvent(st, side_all->join_b, 0));
      }
    } else {
      // lookup                                                              (:198-201)
      SCF_TRY(scf_corr_lookup(levels, cfg->num_levels, cfg->radius, flow8, cfg->mask_corr ? F(ws.maskprev) : nullptr,
                              F(ws.corr), ws.corr_stride, 0, B, H8, W8, st));
      // motion encoder                                                      (raft_decoder.py:152-166)
      const int corr_ch = a.pc[PC_CORR0].cin;
      SCF_TRY(conv(PC_CORR0, {{F(ws.corr), ws.corr_stride, 0, corr_ch}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.c1), 256, 0));
      SCF_TRY(conv(PC_CORR1, {{F(ws.c1), 256, 0, 256}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.cf), 256, 0));
      SCF_TRY(conv(PC_FLOW0, {{menc_flow, 2, 0, 2}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.f1), 128, 0));
      SCF_TRY(conv(PC_FLOW1, {{F(ws.f1), 128, 0, 128}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.cf), 256, 192));
      SCF_TRY(conv(PC_OUT0, {{F(ws.cf), 256, 0, 256}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.motion), 128, 0));
      SCF_TRY(scf_resize_bilinear(menc_flow, nullptr, (long long)P * 2, 1, (long long)W8 * 2, 2, H8, W8, F(ws.motion) + 126,
                                  (long long)P * 128, 1, (long long)W8 * 128, 128, H8, W8, B, 2, 1.f, st));
      // SepConvGRU                                                          (raft_decoder.py:235-253)
      for (int pass = 0; pass < 2; ++pass) {
        float* hin = F(ws.h[pass]);
        float* hout = F(ws.h[pass ^ 1]);
        SCF_TRY(conv(pass == 0 ? PC_ZR0 : PC_ZR1, {{hin, 128, 0, 128}, {F(ws.cxt), 128, 0, 128}, {F(ws.motion), 128, 0, 128}},
                     H8, W8, H8, W8, 1, SCF_ACT_SIGMOID, F(ws.z), 128, 0, SCF_EPI_GRU_ZR, hin, nullptr, F(ws.rh)));
        SCF_TRY(conv(pass == 0 ? PC_Q0 : PC_Q1, {{F(ws.rh), 128, 0, 128}, {F(ws.cxt), 128, 0, 128}, {F(ws.motion), 128, 0, 128}},
                     H8, W8, H8, W8, 1, SCF_ACT_TANH, hout, 128, 0, SCF_EPI_GRU_Q, hin, F(ws.z), nullptr));
      }
      float* h = F(ws.h[0]);
      // flow / mask heads                                                   (scflow_decoder.py:210-213)
      SCF_TRY(conv(PC_HEADS, {{h, 128, 0, 128}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.hd), 512, 0));
      SCF_TRY(conv(PC_FHP, {{F(ws.hd), 512, 0, 256}}, H8, W8, H8, W8, 1, SCF_ACT_NONE, F(ws.dflow), 2, 0));
      SCF_TRY(conv(PC_MHP, {{F(ws.hd), 512, 256, 256}}, H8, W8, H8, W8, 1, SCF_ACT_SIGMOID, F(ws.mask8), 1, 0));
      if (cfg->pose_head) {
        // delta-flow / mask encoders                                        (:216-217)
        SCF_TRY(conv(PC_DFE0, {{F(ws.dflow), 2, 0, 2}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.df1), 128, 0));
        SCF_TRY(conv(PC_DFE1, {{F(ws.df1), 128, 0, 128}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.df2), 64, 0));
        SCF_TRY(conv(PC_ME0, {{F(ws.mask8), 1, 0, 1}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.mf1), 64, 0));
        SCF_TRY(conv(PC_ME1, {{F(ws.mf1), 64, 0, 64}}, H8, W8, H8, W8, 1, SCF_ACT_RELU, F(ws.mf2), 32, 0));
      }
    }
    // flow_pred = 8 * up8(flow8 + dflow) ; mask_up = up8(mask)            (:222-227)  - outputs only: off the critical path
    SideStream* side = (cfg->pose_head && (overlap_mask() & 1)) ? side_stream() : nullptr;
    cudaStream_t up_st = st;
    if (side) {
      SCF_CUDA(cudaEventRecord(side->fork, st));
      SCF_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
      up_st = side->s;
    }
    SCF_TRY(scf_resize_bilinear(flow8, F(ws.dflow), (long long)P * 2, 1, (long long)W8 * 2, 2, H8, W8, flow_pred_k, 2 * HW,
                                HW, W, 1, H, W, B, 2, (float)scale, up_st));
    SCF_TRY(scf_resize_bilinear(F(ws.mask8), nullptr, P, 0, W8, 1, H8, W8, mask_k, HW, 0, W, 1, H, W, B, 1, 1.f, up_st));
    if (side) SCF_CUDA(cudaEventRecord(side->join, side->s));
    float* h = F(ws.h[0]);
    if (cfg->pose_head) {
      // pose regressor                                                    (:218-219, pose_head.py:201-211)
      const int h1 = (H8 - 1) / 2 + 1, w1 = (W8 - 1) / 2 + 1, h2 = (h1 - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1,
                h3 = (h2 - 1) / 2 + 1, w3 = (w2 - 1) / 2 + 1;
      if (tcp) {
        // stride-2 convolutions on the tensor cores (TMA element strides), GroupNorm emits the next conv's split-bf16 input.
        // These maps have few pixel tiles (B=32: 64 / 16 / 4 tiles of 128 pixels for 148 SMs) and a long reduction (9 taps x
        // 128-224 channels), so the taps are split over CTAs and the GroupNorm adds the partial maps when it loads them.
        static const bool ph_split = [] { const char* e = getenv("SCFLOW_PH_SPLIT"); return e ? atoi(e) != 0 : true; }();
        auto pick_split = [&](int out_pix) {
          const int tiles = cdiv((long long)B * out_pix, 128);
          if (!ph_split) return 1;
          return tiles * 9 <= 160 ? 9 : (tiles * 3 <= 200 ? 3 : 1);
        };
        tc_stride = 2;
        tc_hin = H8; tc_win = W8;
        tc_ksplit = pick_split(h1 * w1); tc_split_stride = (long long)B * h1 * w1 * 128 + 256;
        SCF_TRY(convtc(PC_PH0, {{S(ws.s_h[0]), 128, 0, 128}, {S(ws.s_df2), 64, 0, 64}, {S(ws.s_mf2), 32, 0, 32}}, SCF_ACT_NONE,
                       F(ws.p1), 128, nullptr, 0, 0));
        SCF_TRY(group_norm_relu_partials(F(ws.p1), tc_ksplit, tc_split_stride, pw + a.gn_w[0], pw + a.gn_b[0], B, h1 * w1, 128, 32, 1e-5f,
                                         S(ws.s_p1), (long long)B * h1 * w1 * 128, st));
        tc_hin = h1; tc_win = w1;
        tc_ksplit = pick_split(h2 * w2); tc_split_stride = (long long)B * h2 * w2 * 128 + 256;
        SCF_TRY(convtc(PC_PH1, {{S(ws.s_p1), 128, 0, 128}}, SCF_ACT_NONE, F(ws.p2), 128, nullptr, 0, 0));
        SCF_TRY(group_norm_relu_partials(F(ws.p2), tc_ksplit, tc_split_stride, pw + a.gn_w[1], pw + a.gn_b[1], B, h2 * w2, 128, 32, 1e-5f,
                                         S(ws.s_p2), (long long)B * h2 * w2 * 128, st));
        tc_hin = h2; tc_win = w2;
        tc_ksplit = pick_split(h3 * w3); tc_split_stride = (long long)B * h3 * w3 * 128 + 256;
        SCF_TRY(convtc(PC_PH2, {{S(ws.s_p2), 128, 0, 128}}, SCF_ACT_NONE, F(ws.p3), 128, nullptr, 0, 0));
        SCF_TRY(group_norm_relu_partials(F(ws.p3), tc_ksplit, tc_split_stride, pw + a.gn_w[2], pw + a.gn_b[2], B, h3 * w3, 128, 32, 1e-5f,
                                         nullptr, 0, st));
        tc_ksplit = 1; tc_split_stride = 0;
        tc_stride = 1; tc_hin = H8; tc_win = W8;
      } else {
      SCF_TRY(conv(PC_PH0, {{h, 128, 0, 128}, {F(ws.df2), 64, 0, 64}, {F(ws.mf2), 32, 0, 32}}, H8, W8, h1, w1, 2, SCF_ACT_NONE,
                   F(ws.p1), 128, 0));
      SCF_TRY(scf_group_norm_relu(F(ws.p1), pw + a.gn_w[0], pw + a.gn_b[0], B, h1 * w1, 128, 32, 1e-5f, st));
      SCF_TRY(conv(PC_PH1, {{F(ws.p1), 128, 0, 128}}, h1, w1, h2, w2, 2, SCF_ACT_NONE, F(ws.p2), 128, 0));
      SCF_TRY(scf_group_norm_relu(F(ws.p2), pw + a.gn_w[1], pw + a.gn_b[1], B, h2 * w2, 128, 32, 1e-5f, st));
      SCF_TRY(conv(PC_PH2, {{F(ws.p2), 128, 0, 128}}, h2, w2, h3, w3, 2, SCF_ACT_NONE, F(ws.p3), 128, 0));
      SCF_TRY(scf_group_norm_relu(F(ws.p3), pw + a.gn_w[2], pw + a.gn_b[2], B, h3 * w3, 128, 32, 1e-5f, st));
      }
      // (measured, B = 32: the split-K form is 0.04 ms per step SLOWER - the FC kernels' fixed costs, not their weight stream,
      // dominate - so it is opt-in)
      static const bool fc_split = [] { const char* e = getenv("SCFLOW_FC_SPLIT"); return e ? atoi(e) != 0 : false; }();
      static const bool fc_tcore = [] { const char* e = getenv("SCFLOW_FC_TC"); return e ? atoi(e) != 0 : true; }();
      if (fc_tcore && tcp && a.fc_off[0] && B <= 32) {
        // small batch: both FC layers on tcgen05, split-K over 8 blocks per 128 output rows; each consumer sums the producer's
        // partial maps and applies its bias / ReLU while it forms its own operand
        const char* pb = reinterpret_cast<const char*>(packed);
        // SCFLOW_FC_REDUCE (default 1): each layer = one launch, the 8 K-range blocks of a row tile reduce in a cluster; 0: raw
        // split-K partial maps that the consumer sums while it forms its operand
        static const bool fc_reduce = [] { const char* e = getenv("SCFLOW_FC_REDUCE"); return e ? atoi(e) != 0 : true; }();
        if (fc_reduce) {
          SCF_TRY(fc_tc(F(ws.p3), 1, 0, nullptr, 0, pb + a.fc_off[0], nullptr, F(ws.fc0), pw + a.fc0_b, 1, B, 2048, 1024, 256, st));
          SCF_TRY(fc_tc(F(ws.fc0), 1, 0, nullptr, 0, pb + a.fc_off[1], nullptr, F(ws.fc1), pw + a.fc1_b, 1, B, 1024, 256, 128, st));
          SCF_TRY(scf_pose_project(F(ws.fc1), pw + a.rot_w, pw + a.rot_b, pw + a.tr_w, pw + a.tr_b, io->label, drot_k, dtrs_k, B,
                                   256, cfg->rot_dim, cfg->num_class, st));
        } else {
          SCF_TRY(fc_tc(F(ws.p3), 1, 0, nullptr, 0, pb + a.fc_off[0], F(ws.fc0), nullptr, nullptr, 0, B, 2048, 1024, 256, st));
          SCF_TRY(fc_tc(F(ws.fc0), 8, (long long)B * 1024, pw + a.fc0_b, 1, pb + a.fc_off[1], F(ws.fc1), nullptr, nullptr, 0, B, 1024, 256, 64, st));
          SCF_TRY(pose_project_partials(F(ws.fc1), 16, (long long)B * 256, pw + a.fc1_b, pw + a.rot_w, pw + a.rot_b, pw + a.tr_w, pw + a.tr_b,
                                        io->label, drot_k, dtrs_k, B, 256, cfg->rot_dim, cfg->num_class, st));
        }
      } else if (fc_split && B <= 32) {
        // small batch: the FC layers as split-K weight streams over 4x more blocks, bias / ReLU applied by the consumer
        SCF_TRY(pose_fc_tail(F(ws.p3), pw + a.fc0_w, pw + a.fc0_b, 2048, 1024, pw + a.fc1_w, pw + a.fc1_b, 256, pw + a.rot_w, pw + a.rot_b,
                             pw + a.tr_w, pw + a.tr_b, io->label, drot_k, dtrs_k, B, cfg->rot_dim, cfg->num_class, F(ws.fc0), F(ws.fc1), 4, 4, st));
      } else {
        SCF_TRY(scf_linear(F(ws.p3), pw + a.fc0_w, pw + a.fc0_b, F(ws.fc0), B, 2048, 1024, SCF_ACT_RELU, st));
        SCF_TRY(scf_linear(F(ws.fc0), pw + a.fc1_w, pw + a.fc1_b, F(ws.fc1), B, 1024, 256, SCF_ACT_RELU, st));
        SCF_TRY(scf_pose_project(F(ws.fc1), pw + a.rot_w, pw + a.rot_b, pw + a.tr_w, pw + a.tr_b, io->label, drot_k, dtrs_k, B,
                                 256, cfg->rot_dim, cfg->num_class, st));
      }
    } else {
      identity_delta_kernel<<<cdiv(B, 64), 64, 0, st>>>(drot_k, dtrs_k, B, cfg->rot_dim);
      SCF_TRY(check_launch("identity_delta_kernel"));
    }
    // pose update + pose-induced flow                                     (:230-243)
    SCF_TRY(scf_pose_update(drot_k, dtrs_k, rot_prev, trs_prev, rot_k, trs_k, B, st));
    if (it + 1 < iters)
      SCF_TRY(scf_reproject_down(F(ws.pts4), io->internel_k, rot_k, trs_k, io->invalid_flow_num, flow_pose_k, B, H, W, flow8_next, H8, W8, st));
    else
      SCF_TRY(scf_reproject(F(ws.pts4), io->internel_k, rot_k, trs_k, io->invalid_flow_num, flow_pose_k, B, H, W, st));
    flow_full = flow_pose_k;
    if (side) SCF_CUDA(cudaStreamWaitEvent(st, side->join, 0));    // flow8 / dflow / mask8 are rewritten by the next iteration
    if (cfg->mask_corr || cfg->mask_flow)
      SCF_CUDA(cudaMemcpyAsync(F(ws.maskprev), F(ws.mask8), BP * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (io->h_out) SCF_TRY(scf_nhwc_to_nchw(F(ws.h[0]), 128, 0, io->h_out, B, 128, H8, W8, st));
  return 0;
}

int scf_decoder_launch_count(const scf_decoder_cfg* cfg, int iters) {
  if (check_cfg(cfg) != 0) return -1;
  int per_iter = 1 /*lookup*/ + 6 /*menc*/ - (cfg->precision == 1 ? 1 : 0) /*merged predict layers*/ + (cfg->precision == 1 ? (cfg->pose_head ? 2 : 1) : 0) /*x-folding of the flow maps*/ + 4 /*gru*/ + 3 /*heads*/ + 2 /*up8*/ + 2 /*pose upd + reproject*/;
  per_iter += cfg->pose_head ? 4 + 6 + 3 : 1;
  per_iter += cfg->mask_flow ? 1 : 0;
  int once = (cfg->precision == 1 ? 3 : 2) /*layout change of the feature maps + level 0*/ + (cfg->num_levels - 1) + 1 /*unproject*/ +
             2 /*h, cxt*/ + (cfg->precision == 1 ? 4 : 0) /*GRU context terms*/;
  if (cfg->mask_corr || cfg->mask_flow) once += 1;
  return once + 1 /*down8 of init_flow*/ + per_iter * iters;
}

}  // extern "C"

namespace scf {
int corr_build_f32(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                   float* const* levels, void* scratch, cudaStream_t st);

// tcgen05 build: both feature maps are transposed to pixel-major split-bf16 ([2][B][P][C], K-major for both GEMM
// operands); level 0 = batched GEMM  f1[b] (P x C) * f2[b]^T (C x P) / sqrt(C)  on the tensor cores, written once as
// fp32; levels 1.. are the reference's successive floor 2x2 means.
// level 0 from pixel-major split-bf16 feature maps f1s (render) / f2s (real), both with hi->lo plane stride `plane`
bool corr_pyramid_fused_ok(int C, int H8, int W8, int num_levels);
int corr_pyramid_fused(const void* f1s, const void* f2s, long long plane, int B, int C, int H8, int W8, float* const* levels,
                       cudaStream_t st);

static int corr_gemm_and_pool(const void* f1s, const void* f2s, long long plane, int B, int C, int H8, int W8, int num_levels,
                              float* const* levels, cudaStream_t st) {
  // 32-wide maps (256x256 crops): volume and pooled levels from one kernel (scf_corr_fused.cu); other sizes: GEMM + pools
  if (corr_pyramid_fused_ok(C, H8, W8, num_levels)) return corr_pyramid_fused(f1s, f2s, plane, B, C, H8, W8, levels, st);
  const int P = H8 * W8;
  scf_tc_conv_desc d = {};
  d.seg[0].ptr = f1s; d.seg[0].plane_stride = plane; d.seg[0].stride = C; d.seg[0].coff = 0; d.seg[0].nch = C;
  d.nseg = 1;
  d.B = B; d.H = H8; d.W = W8; d.kh = d.kw = 1;
  d.w = f2s; d.cin_pad = C; d.cout_pad = P; d.cout = P; d.w_batched = 1; d.w_plane_stride = plane;
  d.bias = nullptr; d.scale = 1.0f / sqrtf((float)C); d.epi = SCF_EPI_ACT; d.act = SCF_ACT_NONE;
  d.out_f32 = levels[0]; d.out_f32_stride = P; d.out_f32_coff = 0;
  SCF_TRY(conv2d_tc(d, st));
  int hl = H8, wl = W8;
  for (int l = 1; l < num_levels; ++l) {
    SCF_TRY(avgpool2(levels[l - 1], levels[l], (long long)B * P, hl, wl, st));
    hl /= 2; wl /= 2;
  }
  return 0;
}

static int corr_build_tc(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                         float* const* levels, void* scratch, cudaStream_t st) {
  const int P = H8 * W8;
  SCF_REQUIRE(C % 8 == 0 && P % 16 == 0, SCF_ERR_UNSUPPORTED, "scf_corr_build(tc): C %% 8 and H8*W8 %% 16 required");
  char* sc = reinterpret_cast<char*>(scratch);
  const long long plane = (long long)B * P * C;
  void* f1s = sc;
  void* f2s = sc + (size_t)plane * 2 * 2;
  SCF_TRY(scf_nchw_to_nhwc_split(feat_render, f1s, plane, C, 0, nullptr, 0, B, C, H8, W8, st));
  SCF_TRY(scf_nchw_to_nhwc_split(feat_real, f2s, plane, C, 0, nullptr, 0, B, C, H8, W8, st));
  return corr_gemm_and_pool(f1s, f2s, plane, B, C, H8, W8, num_levels, levels, st);
}

// feature maps already in the scratch as ONE split tensor [2][2*B*P][C]: samples [0,B) real, [B,2B) render
int corr_build_presplit(void* scratch, int B, int C, int H8, int W8, int num_levels, float* const* levels, cudaStream_t st) {
  const int P = H8 * W8;
  SCF_REQUIRE(C % 8 == 0 && P % 16 == 0, SCF_ERR_UNSUPPORTED, "scf_corr_build(tc): C %% 8 and H8*W8 %% 16 required");
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(scratch);
  const long long plane = 2LL * B * P * C;
  return corr_gemm_and_pool(base + (long long)B * P * C, base, plane, B, C, H8, W8, num_levels, levels, st);
}

int corr_build_dispatch(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                        float* const* levels, void* scratch, int precision, cudaStream_t st) {
  if (precision == 1) return corr_build_tc(feat_render, feat_real, B, C, H8, W8, num_levels, levels, scratch, st);
  return corr_build_f32(feat_render, feat_real, B, C, H8, W8, num_levels, levels, scratch, st);
}
}  // namespace scf
