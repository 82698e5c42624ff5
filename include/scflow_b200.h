/*
 * scflow_b200.h - C ABI of libscflow_sm100a.so: the B200-native replacement for SCFlow's iterative
 * pose-refinement hot path (correlation-pyramid build -> pyramid lookup -> motion encoder -> SepConvGRU ->
 * flow/mask heads -> pose regressor -> pose update -> pose-induced-flow re-projection).
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   models/decoder/raft_decoder.py:35-58      CorrelationPyramid.forward      -> scf_corr_build
 *   models/utils/corr_lookup.py:102-136       CorrLookup.forward              -> scf_corr_lookup
 *   models/decoder/raft_decoder.py:152-166    MotionEncoder.forward           -> scf_conv2d (x5) / scf_decoder_forward
 *   models/decoder/raft_decoder.py:235-253    ConvGRU.forward                 -> scf_conv2d (GRU epilogues)
 *   models/decoder/raft_decoder.py:292-294    XHead.forward                   -> scf_conv2d
 *   models/head/pose_head.py:201-211          MultiClassPoseHead.forward      -> scf_group_norm_relu, scf_linear, scf_pose_project
 *   models/utils/pose.py:124-169              get_pose_from_delta_pose        -> scf_pose_update
 *   models/utils/pose.py:26-64                cal_3d_2d_corr / lift_2d_to_3d  -> scf_unproject
 *   models/utils/pose.py:66-88                get_flow_from_delta_pose_and_points -> scf_reproject
 *   models/decoder/scflow_decoder.py:196-197,222-227  F.interpolate x1/8, x8  -> scf_resize_bilinear
 *   models/decoder/scflow_decoder.py:150-251  SCFlowDecoder.forward (whole loop) -> scf_decoder_forward
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name starts with "h_" ; the caller (PyTorch) owns all
 *     memory. The library never allocates, frees or retains device memory.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), does no host synchronisation and
 *     has no data-dependent host control flow, so sequences of calls are CUDA-graph capturable.
 *   - Return value: 0 = ok, <0 = argument error (SCF_ERR_*), >0 = cudaError_t. scf_last_error() returns a
 *     thread-local message for the last non-zero return.
 *   - "NCHW" tensors are the reference's layout (contiguous fp32). "NHWC" tensors are the library's internal
 *     pixel-major activation layout: element (b, y, x, c) at ((b*H + y)*W + x)*stride + c.
 */
#ifndef SCFLOW_B200_H_
#define SCFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCF_ABI_VERSION 1

#define SCF_ERR_ARG (-1)        /* bad size / null pointer / unsupported combination */
#define SCF_ERR_ALIGN (-2)      /* pointer or stride not aligned as required */
#define SCF_ERR_UNSUPPORTED (-3)

/* activation codes */
#define SCF_ACT_NONE 0
#define SCF_ACT_RELU 1
#define SCF_ACT_SIGMOID 2
#define SCF_ACT_TANH 3

/* convolution epilogues (scf_conv2d) */
#define SCF_EPI_ACT 0     /* out = act(scale*acc + bias)                                                  */
#define SCF_EPI_GRU_ZR 1  /* cout = 2*Ch: n<Ch: out[n] = z = sigmoid(.) ; n>=Ch: out2[n-Ch] = sigmoid(.)*aux0[n-Ch] (r*h) */
#define SCF_EPI_GRU_Q 2   /* q = tanh(.) ; out[n] = (1-aux1[n])*aux0[n] + aux1[n]*q   (aux0 = h, aux1 = z)  */

int scf_abi_version(void);
/* sizeof() of a descriptor struct as this library was compiled, so that a binding can verify its mirror of the layout:
 * 0 scf_conv_desc, 1 scf_tc_conv_desc, 2 scf_decoder_cfg, 3 scf_decoder_io, 4 scf_encoder_out, 5 scf_loss_desc,
 * 6 scf_gru_pass_desc; -1 otherwise */
int scf_struct_size(int which);
const char* scf_last_error(void);
/* 1 if the tcgen05/TMA code paths are usable on the current device (compute capability 10.x), else 0 */
int scf_device_supported(void);
/* number of kernels launched so far by the calling thread through this library (bench.py's gpu_launches) */
long long scf_launch_counter(void);

/* ---------------------------------------------------------------- layout helpers ------------------------ */
/* NCHW fp32 [B,C,H,W] -> NHWC fp32 rows of `dst_stride` floats, written at channel offset dst_coff. */
int scf_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, int dst_stride, int dst_coff,
                     void* stream);
int scf_nhwc_to_nchw(const float* src, int src_stride, int src_coff, float* dst, int B, int C, int H, int W,
                     void* stream);
/* conv weight OIHW fp32 -> packed [kh*kw*I][ldw] (k = (ky*kw+kx)*I + i, n = o), ldw = round_up(O,4), zero padded.
 * o_off lets several convs be concatenated along O into one packed matrix of width ldw. */
int scf_pack_conv_weight(const float* w_oihw, float* packed, int O, int I, int kh, int kw, int ldw, int o_off,
                         void* stream);

/* ---------------------------------------------------------------- generic convolution -------------------- */
typedef struct scf_conv_seg {
  const float* ptr; /* NHWC activation buffer */
  int stride;       /* floats per pixel in that buffer */
  int coff;         /* first channel used */
  int nch;          /* number of channels taken (concatenated in order over the segments) */
} scf_conv_seg;

typedef struct scf_conv_desc {
  scf_conv_seg seg[3];
  int nseg;
  int B, Hi, Wi, Ho, Wo;
  int kh, kw, sh, sw, ph, pw;
  const float* w;            /* packed weight, see scf_pack_conv_weight */
  long long w_batch_stride;  /* floats between per-sample weight matrices (0 = shared weights) */
  int ldw, cout;
  const float* bias;         /* [cout] or NULL */
  float scale;               /* multiplies the accumulator before bias */
  int epi, act;
  float* out;                /* NHWC */
  int out_stride, out_coff;
  const float* aux0; int aux0_stride;
  const float* aux1; int aux1_stride;
  float* out2; int out2_stride;
  /* optional split-bf16 copy of `out` (hi plane at out_hl, lo plane at out_hl + out_hl_plane elements) for
   * consumption by the tensor-core convolutions; NULL = none */
  void* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff;
} scf_conv_desc;

/* fp32 CUDA-core implicit GEMM (exact fp32 accumulate); any kernel size / stride / channel count. */
int scf_conv2d(const scf_conv_desc* d, void* stream);

/* ---------------------------------------------------------------- tcgen05 convolution -------------------- */
/* Split-bf16 ("bf16x3") tensors: a value x is stored as two bf16 planes hi = bf16(x), lo = bf16(x - hi); a product
 * is evaluated as hi*hi + hi*lo + lo*hi on the 5th-gen tensor cores with fp32 accumulation in TMEM (error ~2^-16,
 * i.e. fp32-class for this network; a single bf16 pass misses the 1e-3 px EPE bar by 60x - BASELINE.md §5). */
typedef struct scf_tc_seg {
  const void* ptr;        /* bf16 hi plane, NHWC */
  long long plane_stride; /* elements from the hi plane to the lo plane */
  int stride, coff, nch;  /* elements per pixel, first channel, channels taken (multiples of 8) */
} scf_tc_seg;

typedef struct scf_tc_conv_desc {
  scf_tc_seg seg[3];
  int nseg;
  int B, H, W;             /* INPUT spatial size; output = (H + 2*(kh/2) - kh)/stride + 1 */
  int kh, kw;              /* odd; padding = kh/2, kw/2 */
  int stride;              /* 1 (0 is read as 1) or 2 */
  const void* w;           /* packed by scf_pack_conv_weight_tc: bf16 [2][taps][cout_pad][cin_pad] */
  int cin_pad, cout_pad, cout;
  int w_batched;           /* 1: the "tap" dimension of w indexes the sample (kh=kw=1; correlation build) */
  const float* bias; float scale; int epi, act;
  float* out_f32; int out_f32_stride, out_f32_coff;                     /* optional fp32 NHWC output */
  void* out_hl; long long out_hl_plane; int out_hl_stride, out_hl_coff; /* optional split-bf16 output */
  const float* aux0; int aux0_stride; const float* aux1; int aux1_stride;  /* EPI_ACT: aux0 = optional residual added before act */
  void* out2_hl; long long out2_hl_plane; int out2_hl_stride;           /* GRU_ZR: r*h as split-bf16 */
  /* --- optional extensions (zero = off) --- */
  const float* pre; int pre_stride;   /* GRU epilogues: fp32 NHWC map added to the accumulator before the gate non-linearity
                                       * (the loop-invariant context contribution of the GRU convolutions, computed once) */
  int stride_x, stride_y;             /* per-axis strides overriding `stride` when non-zero (1 or 2 each) */
  long long w_plane_stride;           /* elements from the hi to the lo plane of `w` (0 = tightly packed) */
  float* stats;                       /* EPI_ACT: per (pixel tile, epilogue warp group) partial sums of the fp32 output,
                                       * rows of [2][cout] floats (sum, then sum of squares), for InstanceNorm: 4 rows per
                                       * 128-pixel tile, or 2 rows per 256-pixel tile when the transposed tiling is used
                                       * (rows of one sample are contiguous; zero-initialised rows beyond the used ones add
                                       * nothing).  Needs one sample per tile and cout % 32 == 0 */
  int out_pad_writable;               /* non-zero: the outputs' padding channels [cout, round_up(cout, 8)) may be overwritten
                                       * (with act(0)); lets layers with cout % 8 != 0 leave through TMA stores, which clip
                                       * the channel axis at 16 B granularity */
  int ksplit; long long split_stride; /* ksplit > 1: split-K over the kernel taps for layers with few pixel tiles - split k contracts
                                       * taps [k*taps/ksplit, ...) and writes its PARTIAL sums to out_f32 + k*split_stride elements;
                                       * the caller adds the ksplit maps.  Plain linear layers only (no bias / activation / other
                                       * outputs), taps % ksplit == 0 */
  const void* aux0_hl; long long aux0_hl_plane; int aux0_hl_stride;
                                      /* EPI_ACT: the residual as split-bf16 NHWC planes (hi + lo, ~2^-17 relative) instead of the fp32
                                       * map aux0 - lets a residual block keep its activations in split form only.  Served by the
                                       * rolling-rows kernel (3x3, stride 1, 64 -> 64 channels, maps >= 96 pixels wide); other
                                       * layers refuse it (SCF_ERR_UNSUPPORTED) */
} scf_tc_conv_desc;
/* upper bound, in 128-pixel-tile units, of the pixel tiles scf_conv2d_tc may use for this output geometry whichever tiling it
 * chooses (size of the `stats` buffer = tiles*4*2*cout floats, zero-initialised), and the 128-pixel tiles per sample (0 if a
 * tile may span samples: then `stats` cannot be used) */
int scf_conv2d_tc_tiles(int B, int Hout, int Wout, int* tiles_per_sample);

int scf_conv2d_tc(const scf_tc_conv_desc* d, void* stream);
/* OIHW fp32 -> split-bf16 [2][kh*kw][cout_pad][cin_pad] (zero padded; caller memsets the buffer first).
 * Several convs can be merged along O with o_off. Bytes = 2*taps*cout_pad*cin_pad*2. */
int scf_pack_conv_weight_tc(const float* w_oihw, void* packed, int O, int I, int kh, int kw, int cin_pad, int cout_pad,
                            int o_off, void* stream);
/* NCHW fp32 -> split-bf16 NHWC planes (and, if dst_f32 != NULL, an fp32 NHWC copy with dst_f32_stride). */
int scf_nchw_to_nhwc_split(const float* src, void* dst_hl, long long plane_stride, int dst_stride, int dst_coff,
                           float* dst_f32, int dst_f32_stride, int B, int C, int H, int W, void* stream);
/* NHWC fp32 channels [src_coff, src_coff+nch) -> split-bf16 planes at dst_coff */
int scf_split_copy(const float* src, int src_stride, int src_coff, void* dst_hl, long long plane_stride, int dst_stride,
                   int dst_coff, long long npix, int nch, void* stream);

/* ---------------------------------------------------------------- SepConvGRU pass, one kernel ----------- */
/* ConvGRU.forward, one pass (models/decoder/raft_decoder.py:245-253): z, r gates, r*h, q and the state update of a 256-pixel
 * tile (whole rows for the 1x5 pass, whole columns for the 5x1 pass) on one SM - z and r accumulate side by side in TMEM, r*h
 * becomes the q convolution's operand in shared memory, h' leaves as fp32 + split-bf16.  The context columns of the GRU's input
 * are loop invariant: their contribution (+ bias) arrives as fp32 maps pre_zr / pre_q, the kernel contracts over [h | motion].
 * Needs a 32 x 32 map (256x256 crops at 1/8 resolution); other sizes keep the two-convolution form (scf_conv2d_tc). */
typedef struct scf_gru_pass_desc {
  const void* h_hl; long long h_plane;    /* state, split-bf16 NHWC [2][B*H*W][128] (lo plane h_plane elements after hi) */
  const float* h_f32;                     /* state, fp32 NHWC [B*H*W][128] */
  const void* m_hl; long long m_plane;    /* motion features, split-bf16 NHWC [2][B*H*W][128] */
  const void* w_zr;                       /* packed bf16 [2][5][256][256]: rows z (0..127) | r (128..255), columns [h | motion] */
  const void* w_q;                        /* packed bf16 [2][5][128][256]: columns [r*h | motion] */
  const float* pre_zr;                    /* fp32 [B*H*W][256] */
  const float* pre_q;                     /* fp32 [B*H*W][128] */
  float* out_f32;                         /* h' fp32 NHWC [B*H*W][128] (not in place) */
  void* out_hl; long long out_plane;      /* h' split-bf16 */
  int B, H, W;
  int vertical;                           /* 0: 1x5 pass (taps along x), 1: 5x1 pass (taps along y) */
} scf_gru_pass_desc;
int scf_gru_pass_fused(const scf_gru_pass_desc* d, void* stream);

/* ---------------------------------------------------------------- correlation pyramid -------------------- */
/* feat_render / feat_real: NCHW fp32 [B,C,H8,W8]. levels[l]: fp32 [B*H8*W8, Hl*Wl] (the reference's
 * [B*P,1,Hl,Wl]), Hl = floor(H_{l-1}/2). scratch: >= scf_corr_build_scratch_bytes(B,C,H8,W8) bytes. */
size_t scf_corr_build_scratch_bytes(int B, int C, int H8, int W8);
/* precision: 0 = fp32 CUDA-core build, 1 = tcgen05 split-bf16 (bf16x3) tensor-core build (fp32 accumulate). */
int scf_corr_build(const float* feat_render, const float* feat_real, int B, int C, int H8, int W8, int num_levels,
                   float* const* h_levels, void* scratch, int precision, void* stream);

/* flow8: NHWC fp32 [B,H8,W8,2] (x, y) at 1/8 resolution. out: NHWC [B,H8,W8,out_stride] written at out_coff,
 * num_levels*(2r+1)^2 channels, channel = level*(2r+1)^2 + a*(2r+1) + b sampling (x + a - r, y + b - r).
 * mask (optional, NHWC [B,H8,W8,1]) multiplies the result (decoder option mask_corr). */
int scf_corr_lookup(const float* const* h_levels, int num_levels, int radius, const float* flow8, const float* mask,
                    float* out, int out_stride, int out_coff, int B, int H8, int W8, void* stream);
/* same gather, written as split-bf16 planes (channels beyond num_levels*(2r+1)^2 up to out_stride are zeroed) */
int scf_corr_lookup_split(const float* const* h_levels, int num_levels, int radius, const float* flow8, const float* mask,
                          void* out_hl, long long plane_stride, int out_stride, int B, int H8, int W8, void* stream);
/* Lookup fused with the motion encoder's first convolution (corr_lookup.py:102-136 + raft_decoder.py:152-155):
 *   out[q, :] = relu(W[256 x 324] * lookup(levels, flow8)[q, :] + bias)   as split-bf16 NHWC [2][B*H8*W8][out_stride]
 * in ONE kernel (4 levels, radius 4; B*H8*W8 % 128 == 0): the gathered 324-vector goes straight to shared memory as the tcgen05
 * A operand and never reaches HBM.  packed_w: scf_lookup_conv_pack() of the OIHW [256, 324, 1, 1] weight (bf16 [2][256][4*96],
 * every level's 81 channels padded to 96).  Neighbour indices are those of scf_corr_lookup (bit-exact). */
size_t scf_lookup_conv_packed_bytes(void);
int scf_lookup_conv_pack(const float* w_oihw, void* packed, void* stream);
int scf_lookup_conv(const float* const* h_levels, const float* flow8, const float* mask, const void* packed_w, const float* bias,
                    void* out_hl, long long out_plane_stride, int out_stride, int B, int H8, int W8, void* stream);
/* debug/parity hook: integer neighbour indices of the lookup, bit-exact against the oracle.
 * x0,y0: int32 [B,H8,W8,2r+1] per level `level` (x0 indexed by a, y0 by b). */
int scf_corr_lookup_taps(int level, int radius, const float* flow8, int32_t* x0, int32_t* y0, int B, int H8, int W8,
                         void* stream);

/* ---------------------------------------------------------------- pose head pieces ----------------------- */
/* in-place GroupNorm(num_groups, eps) + ReLU on NHWC [B, HW, C] */
int scf_group_norm_relu(float* x, const float* gamma, const float* beta, int B, int HW, int C, int num_groups,
                        float eps, void* stream);
/* same (C=128, 32 groups only) plus an optional split-bf16 copy of the result for a following tensor-core conv */
int scf_group_norm_relu_split(float* x, const float* gamma, const float* beta, int B, int HW, int C, int num_groups,
                              float eps, void* out_hl, long long plane_stride, void* stream);
/* y[b, :] = act(W x[b, :] + bias), W row-major [O, I] */
int scf_linear(const float* x, const float* w, const float* bias, float* y, int B, int I, int O, int act,
               void* stream);
/* The same layer on tcgen05 for batches of at most 32 samples (split-K, split-bf16 products): block (o / 128, ks) contracts inputs
 * [ks*KR, (ks+1)*KR) and writes RAW partial sums part[ks][b][o] (fp32, I/KR maps of B*O floats); the consumer adds the maps, this
 * layer's bias and activation.  The layer's own input may itself be such a stack of partial maps: x = act(sum_s x[s*x_split_stride
 * + b*I + i] + x_bias[i]) with x_nsplit maps, x_relu != 0 for ReLU (x_nsplit = 1, x_bias = NULL, x_relu = 0: plain input).
 * w_packed: scf_pack_conv_weight_tc(w, ., O, I, 1, 1, I, O, 0) = split-bf16 [2][O][I].  KR: multiple of 64, <= 256, divides I.
 * y != NULL (needs I / KR == 8): the eight K-range blocks of a row tile run as one thread-block cluster and reduce their partial
 * tiles through distributed shared memory in fixed order: y[b][o] = act(W x + bias) (relu != 0: ReLU) is written directly and
 * `part` is not used. */
int scf_linear_tc(const float* x, int x_nsplit, long long x_split_stride, const float* x_bias, int x_relu, const void* w_packed,
                  float* part, float* y, const float* bias, int relu, int B, int I, int O, int KR, void* stream);
/* final pose projection with the reference's class selection: rows of class label[0] (device int64) only.
 * rot_w [rot_dim*num_class, I], tr_w [3*num_class, I]; outputs d_rot [B,rot_dim], d_trs [B,3]. num_class<=0: single-class head */
int scf_pose_project(const float* x, const float* rot_w, const float* rot_b, const float* tr_w, const float* tr_b,
                     const int64_t* label, float* d_rot, float* d_trs, int B, int I, int rot_dim, int num_class,
                     void* stream);
/* ortho6d delta rotation + 'exp' depth transform pose update (pose.py:124-169). rot [B,3,3], trs [B,3]. */
int scf_pose_update(const float* d_rot, const float* d_trs, const float* rot_in, const float* trs_in, float* rot_out,
                    float* trs_out, int B, void* stream);

/* ---------------------------------------------------------------- geometry ------------------------------- */
/* pts4: [B,H,W,4] = (X_obj, Y_obj, Z_obj, depth>0) per pixel */
int scf_unproject(const float* depth, const float* K, const float* rot, const float* trs, float* pts4, int B, int H,
                  int W, void* stream);
/* flow NCHW [B,2,H,W]: projected flow at depth>0, `invalid` elsewhere */
int scf_reproject(const float* pts4, const float* K, const float* rot, const float* trs, float invalid, float* flow,
                  int B, int H, int W, void* stream);
/* Same, and in the same launch the NEXT iteration's coarse flow  flow8 = (H8/H) * F.interpolate(flow, H8/H, bilinear,
 * align_corners=True)  (scflow_decoder.py:196-197) as NHWC [B,H8,W8,2]: each coarse pixel blends the pose-induced flow of its
 * four full-resolution taps, so the dense map is not read back and no separate resize is launched. */
int scf_reproject_down(const float* pts4, const float* K, const float* rot, const float* trs, float invalid, float* flow,
                       int B, int H, int W, float* flow8, int H8, int W8, void* stream);
/* bilinear, align_corners=True. src element (b,c,y,x) at b*s_b + c*s_c + y*s_y + x*s_x (floats); same for dst.
 * dst = scale * interp(src (+ add, same strides as src, optional)). */
int scf_resize_bilinear(const float* src, const float* add, long long s_b, long long s_c, long long s_y, long long s_x,
                        int Hi, int Wi, float* dst, long long d_b, long long d_c, long long d_y, long long d_x, int Ho,
                        int Wo, int B, int C, float scale, void* stream);

/* ---------------------------------------------------------------- RAFT encoder (feeds the loop) ----------- */
/* RAFTEncoder 'Basic' (models/encoder/raft_encoder.py:286-314): 7x7/2 stem, three residual stages (64,96,128; strides
 * 1,2,2; two BasicBlocks each, models/backbone/resnet.py:14-94), 1x1 to 256 channels at 1/8 resolution.
 * 16 convolution units in network order: 0 conv1; 1-4 res_layer1.{0,1}.{conv1,conv2}; 5,6 res_layer2.0.{conv1,conv2};
 * 7 res_layer2.0.downsample.0; 8,9 res_layer2.1.*; 10,11 res_layer3.0.*; 12 res_layer3.0.downsample.0; 13,14
 * res_layer3.1.*; 15 conv2.  h_weights holds 6 device pointers per unit: conv weight (OIHW), conv bias, and - for
 * norm = BN - the following BatchNorm's weight, bias, running_mean, running_var (NULL for IN / unit 15). */
#define SCF_ENC_UNITS 16
#define SCF_ENC_NORM_IN 0   /* InstanceNorm2d(affine=False), statistics computed per call */
#define SCF_ENC_NORM_BN 1   /* BatchNorm2d in eval mode, folded into the packed weights */
size_t scf_encoder_packed_bytes(void);
size_t scf_encoder_workspace_bytes(int N, int H, int W);
int scf_encoder_pack(int norm, const float* const* h_weights, void* packed, void* stream);
/* images: NCHW fp32 [N,3,H,W]; out_nchw: fp32 [N,256,H/8,W/8] */
int scf_encoder_forward(int norm, const void* packed, const float* images, int N, int H, int W, float* out_nchw,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Same network, but the final 1x1 convolution writes the consumer's layout directly (no NCHW round trip): pixel-major
 * split-bf16 planes and / or fp32 NHWC, with an optional activation per channel block.  split = 256: one block of 256
 * channels (feature encoder: hl0 = [2][N*P][256]); split = 128: channels [0,128) -> block 0, [128,256) -> block 1 (context
 * encoder: block 0 = tanh -> hidden state, block 1 = relu -> context; scflow_refiner.py:101-108). */
typedef struct scf_encoder_out {
  void* hl0; long long plane0; int stride0; float* f32_0; int f32_stride0; int act0;
  void* hl1; long long plane1; int stride1; float* f32_1; int f32_stride1; int act1;
  int split;
} scf_encoder_out;
int scf_encoder_forward_ex(int norm, const void* packed, const float* images, int N, int H, int W, const scf_encoder_out* out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- whole decoder loop --------------------- */
enum scf_decoder_weight {
  SCF_W_CORR0_W = 0, SCF_W_CORR0_B,   /* encoder.corr_net.0.conv   [256,324,1,1] */
  SCF_W_CORR1_W, SCF_W_CORR1_B,       /* encoder.corr_net.1.conv   [192,256,3,3] */
  SCF_W_FLOW0_W, SCF_W_FLOW0_B,       /* encoder.flow_net.0.conv   [128,2,7,7]   */
  SCF_W_FLOW1_W, SCF_W_FLOW1_B,       /* encoder.flow_net.1.conv   [64,128,3,3]  */
  SCF_W_OUT0_W, SCF_W_OUT0_B,         /* encoder.out_net.0.conv    [126,256,3,3] */
  SCF_W_GRU_Z0_W, SCF_W_GRU_Z0_B, SCF_W_GRU_R0_W, SCF_W_GRU_R0_B, SCF_W_GRU_Q0_W, SCF_W_GRU_Q0_B, /* [128,384,1,5] */
  SCF_W_GRU_Z1_W, SCF_W_GRU_Z1_B, SCF_W_GRU_R1_W, SCF_W_GRU_R1_B, SCF_W_GRU_Q1_W, SCF_W_GRU_Q1_B, /* [128,384,5,1] */
  SCF_W_FH0_W, SCF_W_FH0_B,           /* flow_pred.layers.0.conv   [256,128,3,3] */
  SCF_W_FHP_W, SCF_W_FHP_B,           /* flow_pred.predict_layer   [2,256,3,3]   */
  SCF_W_MH0_W, SCF_W_MH0_B,           /* mask_pred.layers.0.conv   [256,128,3,3] */
  SCF_W_MHP_W, SCF_W_MHP_B,           /* mask_pred.predict_layer   [1,256,1,1]   */
  SCF_W_DFE0_W, SCF_W_DFE0_B,         /* delta_flow_encoder.0.conv [128,2,7,7]   */
  SCF_W_DFE1_W, SCF_W_DFE1_B,         /* delta_flow_encoder.1.conv [64,128,3,3]  */
  SCF_W_ME0_W, SCF_W_ME0_B,           /* mask_encoder.0.conv       [64,1,3,3]    */
  SCF_W_ME1_W, SCF_W_ME1_B,           /* mask_encoder.1.conv       [32,64,3,3]   */
  SCF_W_PH_C0_W, SCF_W_PH_G0_W, SCF_W_PH_G0_B,  /* pose_pred.conv_layers.0.{conv.weight [128,224,3,3], gn.weight, gn.bias} */
  SCF_W_PH_C1_W, SCF_W_PH_G1_W, SCF_W_PH_G1_B,  /* [128,128,3,3] */
  SCF_W_PH_C2_W, SCF_W_PH_G2_W, SCF_W_PH_G2_B,
  SCF_W_PH_FC0_W, SCF_W_PH_FC0_B,     /* pose_pred.fc_layers.0.0   [1024,2048] */
  SCF_W_PH_FC1_W, SCF_W_PH_FC1_B,     /* pose_pred.fc_layers.1.0   [256,1024]  */
  SCF_W_PH_ROT_W, SCF_W_PH_ROT_B,     /* pose_pred.rotation_pred   [rot_dim*num_class,256] */
  SCF_W_PH_TR_W, SCF_W_PH_TR_B,       /* pose_pred.translation_pred [3*num_class,256] */
  SCF_W_COUNT
};

typedef struct scf_decoder_cfg {
  int num_levels;   /* 4 */
  int radius;       /* 4 */
  int num_class;    /* 21; <=0 = SingleClassPoseHead */
  int rot_dim;      /* 6 (ortho6d) */
  int mask_flow, mask_corr; /* decoder options (scflow_decoder.py:200-206) */
  int pose_head;    /* 1 = run the pose regressor; 0 = identity delta pose (config 3: stock head cannot run off 256x256) */
  int precision;    /* 0 = fp32 CUDA-core convolutions; 1 = tcgen05 split-bf16 (bf16x3, fp32 accumulate) */
} scf_decoder_cfg;

/* Bytes of the packed-weight arena and of the per-call workspace for a batch of B HxW crops. */
size_t scf_decoder_packed_bytes(const scf_decoder_cfg* cfg);
size_t scf_decoder_workspace_bytes(const scf_decoder_cfg* cfg, int B, int H, int W);
/* h_weights: host array of SCF_W_COUNT device pointers to the reference-layout fp32 parameters (OIHW / [O,I]). */
int scf_decoder_pack(const scf_decoder_cfg* cfg, const float* const* h_weights, void* packed, void* stream);

typedef struct scf_decoder_io {
  /* inputs (reference layouts, fp32 NCHW) */
  const float* feat_render; const float* feat_real;   /* [B,256,H/8,W/8] */
  const float* h_feat; const float* cxt_feat;         /* [B,128,H/8,W/8] */
  const float* ref_rotation; const float* ref_translation; /* [B,3,3], [B,3] */
  const float* depth;                                 /* [B,H,W] */
  const float* internel_k;                            /* [B,3,3] */
  const int64_t* label;                               /* [B] */
  const float* init_flow;                             /* [B,2,H,W] */
  float invalid_flow_num;
  /* outputs: stacked over iterations, [iters, ...] of the reference's 7 lists */
  float* flow_from_pose;   /* [iters,B,2,H,W] */
  float* flow_from_pred;   /* [iters,B,2,H,W] */
  float* rotation;         /* [iters,B,3,3] */
  float* translation;      /* [iters,B,3] */
  float* mask;             /* [iters,B,1,H,W] */
  float* delta_rotation;   /* [iters,B,rot_dim] */
  float* delta_translation;/* [iters,B,3] */
  float* h_out;            /* optional [B,128,H/8,W/8] NCHW final hidden state (may be NULL) */
  /* Sub-batch calls: when out_batch_total > 0 the seven outputs are [iters, out_batch_total, ...] tensors shared by several
   * calls and this call (B samples) writes samples out_batch_offset .. out_batch_offset + B - 1 of every iteration. The
   * Python module uses it to run two half batches on two streams (the tail of one chain overlaps the other's kernels).
   * `label` may then point at the FULL batch's labels: only label[0] is read (pose_head.py:201-211 class-selection quirk). */
  int out_batch_total, out_batch_offset;
  /* native_inputs = 1: feat_render / feat_real / h_feat / cxt_feat are ignored; the caller (scf_encoder_forward_ex) has already
   * written them in the loop's own layout into the workspace slots reported by scf_decoder_workspace_slots (precision 1). */
  int native_inputs;
} scf_decoder_io;
/* Byte offsets into the decoder workspace (precision 1) of the buffers an encoder can fill directly:
 *   slots[0] feature maps, split-bf16 [2][2*B*P][256]: samples [0,B) = feat_real, [B,2B) = feat_render (plane stride 2*B*P*256)
 *   slots[1] hidden state h, split-bf16 [2][B*P][128] ; slots[2] h, fp32 [B*P][128] ; slots[3] context, split-bf16 [2][B*P][128] */
int scf_decoder_workspace_slots(const scf_decoder_cfg* cfg, int B, int H, int W, size_t* slots4);

int scf_decoder_forward(const scf_decoder_cfg* cfg, const void* packed, const scf_decoder_io* io, int B, int H, int W,
                        int iters, void* workspace, size_t workspace_bytes, void* stream);
/* number of kernel launches one scf_decoder_forward call issues (for bench.py's gpu_launches) */
int scf_decoder_launch_count(const scf_decoder_cfg* cfg, int iters);

/* ---------------------------------------------------------------- training loss (forward) ---------------- */
/* models/utils/flow.py:6-26 filter_flow_by_mask, in place: flow [B,2,H,W] is set to `invalid` where it already is
 * (both components >= invalid) or where the bilinear sample (zeros padding, align_corners=False) of gt_mask [B,H,W] at
 * the flow target is < 0.9. */
int scf_filter_flow_by_mask(float* flow, const float* gt_mask, float invalid, int B, int H, int W, void* stream);

/* ---- RAFT baseline decoders: convex x8 up-sampling -------------------------------------------
 * models/decoder/raft_decoder.py:381-416 RAFTDecoder._upsample with a mask (num_levels = 4 => scale 8, radius = 4 => 3x3
 * grid) and raft_decoder_mask.py:143-162 upsample_mask: x [B,C,H,W] (C = 2 flow with mul = 8, C = 1 occlusion with mul = 1),
 * mask [B,576,H,W] (channel = k*64 + i*8 + j) -> out [B,C,8H,8W],
 * out[b,c,8h+i,8w+j] = sum_k softmax_k(mask) * mul*x[b,c,h+ky-1,w+kx-1] (zero padding).  fp32; agrees with the reference to
 * summation order (a few ulp). */
int scf_convex_upsample(const float* x, const float* mask, float* out, int B, int C, int H, int W, float mul, void* stream);

/* ---- input formatting next to the path ------------------------------------------------------
 * models/refiner/base_refiner.py:96-107 (BaseRefiner.format_data_test after the renderer call): images [B,H,W,cin>=3]
 * (renderer output, RGB(A) in [0,1]) -> out_images [B,3,H,W] = (rgb - mean) / std ; zbuf [B,H,W,zk] -> out_depth [B,H,W]
 * = zbuf[...,0], out_mask = (depth > 0).  mean3 / std3 are HOST arrays (already divided by 255 as the reference does).
 * Bit-exact with the reference's fp32 arithmetic. */
int scf_format_rendered(const float* images, int cin, const float* zbuf, int zk, const float* mean3, const float* std3,
                        float* out_images, float* out_depth, float* out_mask, int B, int H, int W, void* stream);

/* SCFlowRefiner.loss after get_pose (models/refiner/scflow_refiner.py:204-258) with the shipped loss configuration
 * (configs/refine_models/scflow.py:75-104): SequenceLoss(gamma) over RAFTLoss (flow), L1Loss (mask) and
 * DisentanglePointMatchingLoss (loss_type 'l1', disentangle_z=True, no scale factors). Forward only. */
typedef struct scf_loss_desc {
  const float* flow_pred;      /* [iters,B,2,H,W] sequence_flow_from_pred                                  */
  const float* mask_pred;      /* [iters,B,1,H,W] sequence_masks                                           */
  const float* rotation;       /* [iters,B,3,3]                                                            */
  const float* translation;    /* [iters,B,3]                                                              */
  const float* gt_flow;        /* [B,2,H,W] ground-truth flow, invalid pixels hold max_flow (already filtered) */
  const float* valid;          /* [B,H,W] rendered mask                                                    */
  const float* gt_rotation; const float* gt_translation;   /* [B,3,3], [B,3]                              */
  const int64_t* label;        /* [B]                                                                      */
  const float* points;         /* [num_class, max_points, 3] model points, zero padded                     */
  const int* num_points;       /* [num_class]                                                              */
  const unsigned char* symmetric; /* [num_class] 1 = nearest-neighbour matching (symmetry_types)           */
  const float* diameter;       /* [num_class] mesh_diameter                                                */
  int iters, B, H, W, num_class, max_points;
  float max_flow, gamma, w_flow, w_pose, w_mask, eps;     /* 400, 0.8, 0.1, 10, 10, 1e-10 in the shipped config */
  void* scratch; size_t scratch_bytes;                    /* >= scf_refiner_loss_scratch_bytes(iters, B), 256B aligned */
  float* out;                  /* [4 + 3*iters]: loss, loss_pose, loss_flow, loss_mask, then the per-iteration (weighted)
                                * pose, flow and mask losses - the reference's seq_*_loss_list */
} scf_loss_desc;
size_t scf_refiner_loss_scratch_bytes(int iters, int B);
int scf_refiner_loss(const scf_loss_desc* d, void* stream);

/* ---------------------------------------------------------------- training: optimizer step -------------- */
/* Global-norm gradient clipping + AdamW over flat fp32 buffers of n elements (n % 4 == 0, 16B aligned): what mmcv's
 * OptimizerHook(grad_clip=dict(max_norm=10.)) + torch.optim.AdamW do per step (configs/refine_models/scflow.py:117-125).
 * grads are first multiplied by grad_scale (1 / world size after an all-reduce SUM); the norm of the scaled gradient and the clip
 * coefficient are written to stats2[0..1]; max_norm <= 0 disables clipping.  scratch: >= min(4*SMs, n/1024) floats.
 * step = 1 for the first update (bias correction). */
int scf_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float max_norm, float grad_scale, float* scratch,
                   int scratch_floats, float* stats2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCFLOW_B200_H_ */
