/*
 * hcflow_b200 -- C ABI of the B200 (sm_100a) HCFlow flow-step engine.
 *
 * The reference (JingyunLiang/HCFlow) is pure Python on top of torch; it has no FFI of
 * its own.  The operators below are the device-side replacements for the torch calls its
 * hot path makes; each entry cites the reference lines it replaces (paths relative to
 * codes/models/modules/).  The host side (hcflow_b200/engine.py) binds them with ctypes
 * and drives them from drop-in HCFlowNet_SR / HCFlowNet_Rescaling modules.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated; `stream` is a cudaStream_t;
 *   - activations are NHWC ("pixel-major"): element (b,y,x,c) of a view lives at
 *     ptr[((b*H + y)*W + x)*ld + c]; a view is a channel slice of a wider buffer, so
 *     `ptr` already includes the channel offset and `ld` is the buffer's channel count
 *     (this is what makes torch.cat / Split / dense-concat zero-copy);
 *   - functions return 0 on success, a cudaError_t value (>0) if a launch failed, or a
 *     negative HCF_E* code for a rejected argument; hcf_last_error() gives the text;
 *   - nothing here synchronises the device or allocates memory, except the *_plan_*
 *     functions, which own small host/device descriptors.
 */
#ifndef HCFLOW_B200_H_
#define HCFLOW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCF_EINVAL (-1)   /* bad argument (shape / alignment / unsupported size) */
#define HCF_ENOTSUP (-2)  /* combination not implemented by this kernel */

#define HCF_ABI_VERSION 3

/* ---- introspection -------------------------------------------------------------- */
int hcf_abi_version(void);
const char* hcf_last_error(void);
/* number of kernel launches issued through this library since the last reset */
uint64_t hcf_launch_count(void);
void hcf_launch_count_reset(void);

/* ---- convolution ---------------------------------------------------------------- */
/* One input segment of a (virtually concatenated) conv input.  `up_shift` s means the
 * segment is stored at 1/2^s resolution and read with nearest-neighbour upsampling
 * (replaces F.interpolate(..., mode='nearest') + torch.cat, FlowNet_SR_x4.py:98,117,
 * FlowNet_SR_x8.py:109-113,132-137). */
typedef struct {
  const float* ptr;
  int32_t ld;
  int32_t C;
  int32_t up_shift;
  int32_t _pad;
} hcf_seg;

#define HCF_ACT_NONE 0
#define HCF_ACT_RELU 1
#define HCF_ACT_LRELU 2 /* negative slope 0.2 */

/* Stride-1 "same" convolution, ks in {1,3}, over the channel-concatenation of up to three
 * segments, with a fused epilogue:
 *     v = acc (+ pre); v = (v + bias[c]) * scale[c]; v = act(v);
 *     if (res1) v = v*alpha1 + res1;  if (res2) v = v*alpha2 + res2;
 *     out = v; if (out2) out2 = v;
 * Replaces nn.Conv2d / F.conv2d + the elementwise ops around it:
 *   Basic.py:49-53  (Conv2d + ActNorm: bias=actnorm.bias, scale=exp(actnorm.logs)) + ReLU :443-444
 *   Basic.py:70-72  (Conv2dZeros: bias, scale=exp(3*logs))
 *   Basic.py:349-356, 377-383 (dense convs + LeakyReLU 0.2, x5*0.2 + x), :394-398 (RRDB out*0.2 + x)
 *   ConditionalFlow.py:99-110 (conv_first, trunk_conv1 + skip)
 * Weights are pre-packed as w[tap][k][n]: tap = ky*ks+kx, k runs over the segments with
 * each segment zero-padded to a multiple of 8 channels (kpad = total), n < npad where
 * npad is 16, 32 or a multiple of 64 and >= cout.  bias / scale have npad entries. */
/* Optional FlowStep tail of a conv (tensor-core chains only): the conv is the last layer of a coupling sub-net
 * and its output h is consumed in the epilogue instead of being stored (FlowStep.py:55-64, reverse):
 *     z[n_pass + j] = z[n_pass + j] * exp(-0.318 * atan(2 * h[2j+1])) - h[2j]     (AffineCouplings.py:73-87)
 *     z = W^-1 z                                                                  (Permutations.py:103-108)
 *     z = z * an_scale - an_bias                                                  (ActNorms.py:66-69)
 * in place on the NHWC view z (C <= 24 channels, cout == 2 * (C - n_pass) <= 32).  Replaces a separate
 * hcf_step_inverse launch; lets the convs of consecutive FlowSteps run as one chain. */
typedef struct {
  float* z;
  int32_t z_ld;
  int32_t C;
  int32_t n_pass;
  int32_t z16_ld;
  const float* w;        /* W^-1 [C][C] row-major, or NULL (no permutation) */
  const float* an_scale; /* exp(-logs) [C] */
  const float* an_bias;  /* [C] */
  void* z16_hi;          /* optional: fp16 copy of the new z[:, :n_pass], row pitch z16_ld (fp16 chains) */
  void* z16_lo;          /* optional (with z16_hi): its lo plane, fp16((z - hi) * 2048): the next step's first conv runs split */
} hcf_conv_step;

typedef struct {
  int32_t B, H, W;
  int32_t nseg;
  hcf_seg seg[3];
  int32_t ks;
  int32_t kpad;
  int32_t cout;
  int32_t npad;
  const float* w;
  const float* bias;  /* may be NULL */
  const float* scale; /* may be NULL */
  int32_t act;
  int32_t out_ld;
  float* out;
  float* out2; /* may be NULL */
  int32_t out2_ld;
  int32_t res1_ld;
  const float* res1; /* may be NULL */
  const float* res2; /* may be NULL */
  int32_t res2_ld;
  float alpha1;
  float alpha2;
  int32_t pre_ld;
  const float* pre; /* may be NULL; tensor-core kernels only: fp32 NHWC view [B,H,W,>=cout] added to the
                       accumulator BEFORE bias / scale / activation (the part of a conv over an input that is
                       shared by several convs, computed once by another conv); excludes res1 / res2 */
  const hcf_conv_step* step; /* may be NULL; host pointer, read when the plan is created */
  float* raw2;     /* may be NULL; tensor-core kernels, cout == 64: accumulator columns [32, 64) are stored RAW (no
                      bias / scale / activation) as fp32 to raw2[..., c - 32] and only columns [0, 32) take the normal
                      epilogue into out: the conv computes, on the inputs they share, the partial sum of a later conv
                      next to its own output (that conv adds it through `pre`) */
  int32_t raw2_ld;
  int32_t _pad2;
} hcf_conv_args;

/* fp32 CUDA-core (FFMA) implementation: exact-fp32 parity mode and odd shapes. */
int hcf_conv_fp32(const hcf_conv_args* a, void* stream);

/* tcgen05 tensor-core implementation: ks in {1,3}, up to three segments without upsampling
 * (16-byte aligned views, ld % 4 == 0, any channel count: the K axis is walked in 32-channel
 * chunks per segment and TMA zero-fills beyond a segment's last channel), cout <= 128.
 * `wtc` is the UMMA-ready weight image produced by hcf_conv_tc_pack_weights from weights
 * whose input-channel axis is already padded per segment to multiples of 32 (kin);
 * passes = 1 (TF32) or 3 (3xTF32 split, ~fp32 accuracy).  A plan owns the TMA tensor maps
 * for one (args) tuple. */
typedef struct hcf_conv_tc_plan hcf_conv_tc_plan;
int hcf_conv_tc_supported(const hcf_conv_args* a);
/* bytes needed for the packed weight image of (kin channels, cout); the image depends on `passes` */
int64_t hcf_conv_tc_weight_bytes(int32_t kin, int32_t cout, int32_t ks, int32_t passes);
/* host-side packing: w_oihw [cout][kin][ks][ks] fp32 -> image (host pointers) */
int hcf_conv_tc_pack_weights(const float* w_oihw, int32_t kin, int32_t cout, int32_t ks, int32_t passes, float* image);
int hcf_conv_tc_plan_create(const hcf_conv_args* a, const float* wtc, int32_t passes,
                            hcf_conv_tc_plan** out);
/* A chain of n dependent convolutions on the same [B,H,W] grid and kernel size (e.g. the 211 convs
 * of one RRDB encoder level, ConditionalFlow.py:99-110) executed by ONE persistent launch: conv i
 * may read anything convs < i wrote; readiness is tracked per 3x3 tile neighbourhood through
 * `done_flags` (device int32[B*ceil(H/16)*ceil(W/8)], zeroed by the caller before every run).
 * `layer_passes[i]` in {1,3} selects TF32 / 3xTF32 per conv; wtc[i] must be packed with the same value. */
int hcf_conv_chain_create(const hcf_conv_args* args, const float* const* wtc, const int32_t* layer_passes,
                          int32_t n, int32_t* done_flags, hcf_conv_tc_plan** out);
/* The same chained launch on FP16 operands (kind::f16, 64 channels per 128-byte row: half the shared-memory
 * and L2 traffic per MAC of the TF32 path, and round-to-nearest instead of TF32 truncation).
 * Every fp32 activation buffer a chain conv reads or writes is shadowed by two fp16 planes of the same
 * geometry (same ld, same channel offsets): value = hi + lo / 2048, hi = fp16(value),
 * lo = fp16((value - hi) * 2048).  `shadows` lists them; views are matched to buffers by address range.
 * layer_passes[i] = 1: D = A_hi x W_hi (11-bit operands); 3: D = A_hi x W_hi + (A_hi x W_lo + A_lo x W_hi) / 2048
 * (~fp32 accuracy, needs the lo plane of its inputs); layer_split[i] (may be NULL; -1 = whole K) restricts the split
 * to the leading layer_split[i] padded input channels, the rest of K runs one pass (an RDB's conv5 only needs it on
 * the 64 residual-stream channels).  out_flags[i] says which representations conv i writes:
 * the fp32 view (residual sources and anything read outside the chain), the hi plane, the lo plane.
 * Inputs that no chain conv produced must be converted first (hcf_split16).  w16[i] comes from
 * hcf_conv_tc16_pack_weights with the same `passes`; the input-channel axis is padded per segment to
 * multiples of 64.  Returns HCF_ENOTSUP when a conv of the chain does not qualify (ld % 8, alignment).
 * Scheduling: work items (tile, layer) rotate over the persistent CTAs.  HCF_TC_WS=1 in the environment at plan creation
 * selects the weight-stationary schedule instead (csrc/conv_ws_kernel.cuh: a CTA owns up to three tiles per image group
 * and fetches every weight slab once per group pass) for 3x3 chains with B >= 2 whose accumulators fit its TMEM slots;
 * same results to fp32 rounding, measured slower on the encoder chains of configs[1], opt-in. */
typedef struct {
  const float* f32;  /* base of the fp32 buffer */
  int64_t bytes;     /* its size */
  void* hi;          /* fp16 planes, same element count */
  void* lo;
} hcf_shadow16;
#define HCF_OUT_F32 1
#define HCF_OUT_HI 2
#define HCF_OUT_LO 4
int hcf_conv_tc16_supported(const hcf_conv_args* a);
/* split_kin: number of LEADING (padded) input channels whose weights are packed as [hi ; lo] row blocks
 * (multiple of 64; 0 = one pass, kin = split everywhere) */
int64_t hcf_conv_tc16_weight_bytes(int32_t kin, int32_t cout, int32_t ks, int32_t split_kin);
int hcf_conv_tc16_pack_weights(const float* w_oihw, int32_t kin, int32_t cout, int32_t ks, int32_t split_kin, void* image);
/* seg16 (may be NULL): n x 3 explicit fp16 views for input segments whose fp32 geometry does not satisfy TMA's
 * 16-byte rules in fp16 (ld % 8, channel offset % 8): entries with hi != NULL replace the shadow lookup. */
typedef struct {
  const void* hi;
  const void* lo;
  int32_t ld;
  int32_t _pad;
} hcf_seg16;
int hcf_conv_chain16_create(const hcf_conv_args* args, const void* const* w16, const int32_t* layer_passes,
                            const int32_t* layer_split, const int32_t* out_flags, int32_t n, int32_t* done_flags, const hcf_shadow16* shadows,
                            int32_t n_shadows, const hcf_seg16* seg16, hcf_conv_tc_plan** out);
/* ---- fused FlowStep chains (csrc/flowstep_tc.cu): ONE work item = one whole FlowStep (FlowStep.py:40-64) with an FCN
 * coupling sub-net (Basic.py:426-447: conv3x3 -> ActNorm, ReLU -> conv1x1 -> ActNorm, ReLU -> conv3x3 * exp(3 logs)) on
 * one 16x8-pixel tile, the sub-net's hidden activations never leave the SM; n_steps consecutive FlowSteps on the same z
 * are one persistent cooperative launch.  Replaces, per step: 3 F.conv2d + 2 ActNorm + the coupling / invconv / ActNorm
 * elementwise ops of AffineCouplings.py:28-87, Permutations.py:94-108, ActNorms.py:45-94.  C in {6, 12, 21, 24},
 * n_pass = C / 2 channels condition the sub-net (hidden width 64), fp16 operands (split = 1: hi + lo on both operands). */
typedef struct {
  const void* w1;       /* conv1 over z1: image of hcf_flowstep_pack_w1 (device) */
  const void* w2;       /* conv2 1x1 64->64: hcf_conv_tc16_pack_weights(kin 64, cout 64, ks 1, split_kin 64 or 0) */
  const void* w3;       /* conv3 3x3 64->2(C - n_pass): hcf_conv_tc16_pack_weights(kin 64, cout, ks 3, split_kin 64 or 0) */
  const float* bias1;   /* device vectors: ActNorm of conv1 (bias, exp(logs)), of conv2, and conv3's bias / exp(3 logs); */
  const float* scale1;  /* >= 64 entries (conv3: >= ceil16(cout)) */
  const float* bias2;
  const float* scale2;
  const float* bias3;
  const float* scale3;
  const float* w;        /* per-pixel C x C matrix of THIS step (inverse: W^-1, forward: W), NULL = no permutation */
  const float* an_scale; /* ActNorm of THIS step: inverse exp(-logs), forward exp(logs) */
  const float* an_bias;
  const float* pre;      /* conditional steps: the W_u * u part of conv1, fp32 NHWC [B,H,W,pre_ld >= 64]; may be NULL */
  int32_t pre_ld;
  int32_t _pad;
} hcf_flowstep;

typedef struct {
  int32_t B, H, W;
  int32_t C, n_pass;
  int32_t n_steps;
  int32_t split;      /* 1: all three convs with hi + lo operands (x3 modes); 0: one fp16 pass */
  int32_t forward;    /* 0: inverse tails (coupling^-1, W^-1, ActNorm^-1 of each step);
                         1: forward tails (coupling of step s + log-det, then ActNorm + W of step s + 1); the caller
                            applies the FIRST step's ActNorm + W beforehand (hcf_step_forward_head) */
  float* z;           /* [B,H,W,z_ld] fp32, C channels, updated in place */
  int32_t z_ld;
  int32_t _pad;
  void* z16_a;        /* two staging buffers [B,H,W,32] fp16 (hi 16 | lo 16 channels per pixel), ping-pong between steps; */
  void* z16_b;        /* A must hold z[:, :n_pass] before the run (hcf_flowstep_stage_z1) */
  int32_t* done;      /* per-tile step counters, B * ceil(H/16) * ceil(W/8) int32, zeroed before every run */
  double* logdet;     /* forward: per-image log-det accumulators [B] (sum of the coupling log-scales is added); may be NULL */
  const hcf_flowstep* steps;
} hcf_flowstep_chain_args;

typedef struct hcf_flowstep_plan hcf_flowstep_plan;
int64_t hcf_flowstep_w1_bytes(void);
/* w: [64][n_pass][3][3] fp32 (host), the z1 input channels of the sub-net's first conv */
int hcf_flowstep_pack_w1(const float* w, int32_t n_pass, void* image);
int hcf_flowstep_stage_z1(const float* z, int32_t z_ld, int32_t n_pass, int64_t npix, void* z16, void* stream);
int hcf_flowstep_chain_create(const hcf_flowstep_chain_args* a, hcf_flowstep_plan** out);
/* re-gathers the per-step bias / scale / W / ActNorm tables after the arrays behind them were rewritten in place */
int hcf_flowstep_chain_refresh(hcf_flowstep_plan* p, void* stream);
int hcf_flowstep_chain_set_status(hcf_flowstep_plan* p, int32_t* status);
int hcf_flowstep_chain_run(const hcf_flowstep_plan* p, void* stream);
void hcf_flowstep_chain_destroy(hcf_flowstep_plan* p);

/* fp32 NHWC view (ld, C, npix pixels) -> hi / lo planes with row pitch dst_ld (lo may be NULL) */
int hcf_split16(const float* src, int32_t ld, int32_t C, int64_t npix, void* hi, void* lo, int32_t dst_ld, void* stream);
int32_t hcf_conv_tc_plan_layers(const hcf_conv_tc_plan* p);
/* a chained plan (n > 1) is launched COOPERATIVELY: its CTAs wait on each other's tiles, the driver guarantees that
 * the whole grid is co-resident (creation fails with HCF_ENOTSUP if it can not be) */
int hcf_conv_tc_run(const hcf_conv_tc_plan* p, void* stream);
/* sticky device status word of a plan (NULL = none), OR-ed by the kernels, read and cleared by the host */
#define HCF_STATUS_F16_OVERFLOW 1 /* an fp16 operand plane saturated at +-65504 (value out of the fp16 modes' range) */
#define HCF_STATUS_DEP_TIMEOUT 2  /* a dependency wait of a chained launch timed out (the kernel trapped) */
int hcf_conv_tc_plan_set_status(hcf_conv_tc_plan* p, int32_t* status);
/* a plan reads bias / scale through an inline per-layer table gathered at creation: call this after the arrays
 * behind hcf_conv_args.bias / .scale were rewritten in place (weight reload); copies are queued on `stream` */
int hcf_conv_tc_plan_refresh(hcf_conv_tc_plan* p, void* stream);
void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan* p);

/* ---- flow step ------------------------------------------------------------------ */
#define HCF_COUPLING_AFFINE 0       /* z[:, :n_pass] passes; rest gets affine from interleaved h */
#define HCF_COUPLING_SHIFT_FIRST3 1 /* z[:, :3] gets a pure shift h[0..2] (Affine3shift, LRvsothers=False) */

typedef struct {
  int32_t npix;        /* B*H*W */
  int32_t pix_per_img; /* H*W */
  float* z;            /* in place */
  int32_t z_ld;
  int32_t C;
  const float* h; /* coupling sub-net output: (shift,scale) interleaved, 2*(C-n_pass) ch; or 3 ch */
  int32_t h_ld;
  int32_t mode;
  int32_t n_pass;
  int32_t _pad;
  const float* w;        /* [C][C] row-major mixing matrix (W on forward, W^-1 on inverse); NULL = none */
  const float* an_scale; /* [C] exp(+logs) forward, exp(-logs) inverse */
  const float* an_bias;  /* [C] */
  double* logdet;        /* [B], forward only (may be NULL) */
} hcf_step_args;

/* Reverse FlowStep tail (FlowStep.py:53-64): z2 = z2*exp(-0.318*atan(2*scale)) - shift
 * (AffineCouplings.py:65-87, :148-160), z = W^-1 z (Permutations.py:72-74,105),
 * z = z*exp(-logs) - bias (ActNorms.py:90-93). */
int hcf_step_inverse(const hcf_step_args* a, void* stream);
/* Forward FlowStep head (FlowStep.py:40-47): z = (z + bias)*exp(logs) (ActNorms.py:84-88),
 * z = W z (Permutations.py:100).  The data-independent log-dets are added by the host. */
int hcf_step_forward_head(const hcf_step_args* a, void* stream);
/* Forward coupling (AffineCouplings.py:28-63, :117-139): z2 = (z2 + shift)*exp(ls);
 * logdet[b] += sum ls. */
int hcf_step_forward_coupling(const hcf_step_args* a, void* stream);

/* ---- prior (ConditionalFlow.py:44-96, Basic.py:75-100) -------------------------- */
typedef struct {
  int32_t B, H, W;
  int32_t Cz;
  const float* h; /* Conv2dZeros output, 2*Cz ch: mean = h[2c], second = h[2c+1] (cross split, thops.py:43-44) */
  int32_t h_ld;
  int32_t atan_logscale; /* 0: logs = second (SR);  1: logs = 0.318*atan(2*second) (Rescaling, :80,:89) */
  const float* eps_nchw; /* [B,Cz,H,W] noise already multiplied by eps_std; NULL = zeros */
  float* z;              /* NHWC view, Cz channels */
  int32_t z_ld;
  int32_t _pad;
  double* logdet;  /* [B] (logp) */
  float* out_nchw; /* [B,Cz,H,W] (standardize) */
} hcf_prior_args;
int hcf_prior_sample(const hcf_prior_args* a, void* stream);      /* z = mean + exp(logs)*eps */
int hcf_prior_logp(const hcf_prior_args* a, void* stream);        /* logdet[b] += log N(z; mean, exp(logs)) */
int hcf_prior_standardize(const hcf_prior_args* a, void* stream); /* out = (z-mean)*exp(-logs) */

/* ---- layout / plumbing ---------------------------------------------------------- */
typedef struct {
  int32_t B, C, H, W;
  const float* src;
  float* dst;
  int32_t ld;   /* NHWC side channel stride */
  int32_t post; /* nhwc_to_nchw: 0 raw, 1 clamp[0,1], 2 clamp + round to 1/255 (Basic.py:186-192) */
  const float* noise; /* nchw_to_nhwc: optional NCHW tensor added as src + noise*noise_scale
                         (HCFlowNet_SR_arch.py:52) */
  float noise_scale;
  int32_t _pad;
} hcf_layout_args;
int hcf_nchw_to_nhwc(const hcf_layout_args* a, void* stream);
int hcf_nhwc_to_nchw(const hcf_layout_args* a, void* stream);

typedef struct {
  int32_t B, C, H, W; /* C,H,W of the LOW-resolution (squeezed) side is (4C? no:) see below */
  const float* src;
  int32_t src_ld;
  int32_t dst_ld;
  float* dst;
} hcf_squeeze_args;
/* squeeze2d (Basic.py:127-140): src [B,2H,2W,C] -> dst [B,H,W,4C], dst[..., c*4+i*2+j] = src[2y+i, 2x+j, c].
 * B,C,H,W describe: C = channels of the high-res side, H,W = low-res size. */
int hcf_squeeze2d(const hcf_squeeze_args* a, void* stream);
int hcf_unsqueeze2d(const hcf_squeeze_args* a, void* stream); /* Basic.py:143-157, src low-res 4C -> dst high-res C */
/* Haar (Basic.py:470-487): forward src [B,2H,2W,C] -> dst [B,H,W,4C] with dst[..., k*C+c] = band k of channel c;
 * inverse the other way. */
int hcf_haar_forward(const hcf_squeeze_args* a, void* stream);
int hcf_haar_inverse(const hcf_squeeze_args* a, void* stream);

/* dst[..., :C] = src[..., :C] for two NHWC views of size [B,H,W] (Basic.py:489-499 Split / cat when a view
 * would break TMA alignment). */
int hcf_copy_view(const hcf_squeeze_args* a, void* stream);
/* 8-bit image edges: src uint8 [npix][3] as cv2 delivers it (BGR when swap_rb) -> the 3 leading channels of
 * an fp32 NHWC view in [0,1], RGB (codes/data/util.py:72-86, GTLQ_dataset.py:109-115); and back: clamp [0,1],
 * (x * 255).round() half-to-even, RGB -> BGR when swap_rb (codes/utils/util.py:790-816 tensor2img). */
int hcf_u8_hwc_to_nhwc(const uint8_t* src, float* dst, int32_t ld, int64_t npix, int32_t swap_rb, void* stream);
int hcf_nhwc_to_u8_hwc(const float* src, int32_t ld, uint8_t* dst, int64_t npix, int32_t swap_rb, void* stream);
/* nearest-neighbour upsampling by 2^shift (1..3) from the low-res view src to the high-res view dst (a->H, a->W =
 * high-res size; C % 4 == 0, 16-byte aligned views): materialises an up-sampled conv segment
 * (F.interpolate(mode='nearest'), FlowNet_SR_x4.py:98,117) so that the conv runs on the tensor cores. */
int hcf_upsample_nearest(const hcf_squeeze_args* a, int32_t shift, void* stream);

/* logdet[b] += sum_{c,h,w} log N(x; mean, exp(logs)) with a constant logs, all NCHW
 * (HCFlowNet_SR_arch.py:63 with logs = -6). n = C*H*W elements per image. */
int hcf_gauss_logp_const(const float* x, const float* mean, float logs, int32_t B, int32_t n,
                         double* logdet, void* stream);

/* ---- training path (csrc/grad_ops.cu): the backward halves of the hot-path ops, surfaced as torch.autograd.Function
 * extensions by hcflow_b200/autograd.py.  They replace what torch autograd records for the reference's modules in
 * HCFlow_SR_model.py:184-218 (optimize_parameters: nll.backward()): conv2d backward (Basic.py:14-72, 360-383), ActNorm
 * (ActNorms.py:45-94), the affine coupling (AffineCouplings.py:28-87), GaussianDiag.logp (Basic.py:79-93), the residual
 * scale-adds (Basic.py:383, 398), F.interpolate's adjoint (FlowNet_SR_x4.py:98) and Quant (Basic.py:186-198).
 * All tensors are dense fp32 NHWC ([npix, C] row-major) unless a row pitch is given.  dx of a conv is hcf_conv_fp32 on
 * dy with the flipped / transposed weights. */
/* dw[co][ci][ky][kx] = sum_pix dy[pix, co] * x[pix shifted by (ky, kx), ci]  (zero padding); dw is overwritten */
int hcf_conv_wgrad(const float* x, int32_t x_ld, const float* dy, int32_t dy_ld, int32_t B, int32_t H, int32_t W, int32_t Cin,
                   int32_t Cout, int32_t ks, float* dw, void* stream);
/* out[c] = sum over pixels of y[pix, c] (a conv's bias gradient) */
int hcf_channel_sum(const float* y, int32_t ld, int32_t C, int64_t npix, float* out, void* stream);
/* y = act((x + bias[c]) * scale[c]); bias / scale may be NULL; act = HCF_ACT_* */
int hcf_affine_act_fwd(const float* x, const float* bias, const float* scale, int32_t act, float* y, int64_t npix, int32_t C,
                       void* stream);
/* dx = dy * act'(.) * scale; dbias[c] = sum dx; dscale[c] = sum dy * act'(.) * (x + bias) (either may be NULL) */
int hcf_affine_act_bwd(const float* dy, const float* x, const float* bias, const float* scale, int32_t act, float* dx,
                       float* dbias, float* dscale, int64_t npix, int32_t C, void* stream);
/* affine coupling on the nc coupled channels, h = [shift0, scale0, shift1, ...] (2 nc channels):
 * forward out = (z2 + shift) * exp(ls), lsum[img] += sum ls (lsum may be NULL); inverse out = z2 * exp(-ls) - shift */
int hcf_coupling_fwd(const float* z2, const float* h, int32_t nc, int32_t inverse, float* out, double* lsum, int64_t npix,
                     int32_t pix_per_img, void* stream);
int hcf_coupling_bwd(const float* dout, const float* dlsum, const float* z2, const float* h, int32_t nc, int32_t inverse,
                     float* dz2, float* dh, int64_t npix, int32_t pix_per_img, void* stream);
/* out[img] += sum -0.5 (2 logs + (x - mean)^2 / exp(2 logs) + ln 2 pi); logs == NULL: the constant logs_const */
int hcf_gauss_logp_fwd(const float* x, const float* mean, const float* logs, float logs_const, int32_t B, int64_t per_img,
                       double* out, void* stream);
/* g[img] = d loss / d out[img] (fp32); dx / dmean / dlogs may be NULL */
int hcf_gauss_logp_bwd(const float* g, const float* x, const float* mean, const float* logs, float logs_const, int32_t B,
                       int64_t per_img, float* dx, float* dmean, float* dlogs, void* stream);
/* prior sample: g == NULL: out = mean + exp(logs) * eps;  g != NULL (backward): dlogs = g * exp(logs) * eps */
int hcf_gauss_sample(const float* mean, const float* logs, const float* eps, const float* g, float* out, float* dlogs, int64_t n,
                     void* stream);
/* y = alpha a + beta b (b may be NULL) */
int hcf_axpby(const float* a, float alpha, const float* b, float beta, float* y, int64_t n, void* stream);
/* y = round(clamp(x, 0, 1) * 255) / 255 */
int hcf_quantize8(const float* x, float* y, int64_t n, void* stream);
/* adjoint of nearest up-sampling by 2^shift: dst [B,H,W,C] = block sums of src [B, H << shift, W << shift, C] */
int hcf_downsample_sum(const float* src, float* dst, int32_t B, int32_t H, int32_t W, int32_t C, int32_t shift, void* stream);

/* ---- tiled inference and evaluation metrics on the device.
 * test_patchwise (codes/data/util.py:489-514): E[.., y0:y0+ph, x0:x0+pw] += patch, W += 1, E /= W.
 * patches: [n, C, ph, pw] fp32 (the module's NCHW output); y0 / x0: device int32 [n]; E: [C, H, W]; cnt: [H, W]. */
int hcf_tile_accumulate(const float* patches, int32_t n, int32_t C, int32_t ph, int32_t pw, const int32_t* y0,
                        const int32_t* x0, float* E, float* cnt, int32_t H, int32_t W, void* stream);
int hcf_tile_normalize(float* E, const float* cnt, int32_t C, int32_t H, int32_t W, void* stream);
/* calculate_psnr_ssim's raw sums (codes/utils/util.py:902-982, codes/data/util.py:209-230), fp64, [0, 255] scale:
 * out[0..C-1] squared-difference sums per channel, out[C..2C-1] SSIM-map sums per channel (11x11 Gaussian window `win`,
 * 121 doubles, 'valid' region), out[2C], out[2C+1] the same for the Y channel of a BGR image (C == 3).
 * a, b: HWC images, uint8 (is_f32 = 0) or float in [0, 1] (is_f32 = 1), cropped by `crop` pixels on every side. */
int hcf_image_metrics(const void* a, const void* b, int32_t is_f32, int32_t H, int32_t W, int32_t C, int32_t crop,
                      const double* win, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HCFLOW_B200_H_ */
