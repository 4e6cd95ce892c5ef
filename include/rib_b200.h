/* rib_b200 — C ABI of the B200-native pose-guided rendering hot path.
 *
 * The reference (azuxmioy/Render-In-Between) is pure Python/PyTorch and has no FFI; these entry
 * points are what a binding for its hot path would call.  Each one names the reference interface
 * it replaces (paths relative to Pose_Guided_Neural_Rendering/).
 *
 * Conventions: plain pointers and sizes only; every tensor pointer is a CUDA device pointer owned
 * by the caller unless stated otherwise; every call is asynchronous on `stream` (a cudaStream_t
 * passed as void*), allocates no device memory (except rib_generator_create) and returns 0 on
 * success or a negative code, with the message available from rib_last_error().  No CPU fallback.
 *
 * Environment switches read by the library (all optional; defaults are the measured-best settings):
 *   RIB_AUTOTUNE=0      no plan-time timing of candidate tilings (static heuristics + imported table only)
 *   RIB_XF=0|1|2        scope of the in-kernel instance-norm transform of the mask network (default 1)
 *   RIB_MERGE_LABEL=0   down_first and down_lbl.0 as two launches instead of one GEMM over the label
 *   RIB_SPADE2=0        the two SPADE layers of a shortcut res-block in separate N tiles
 *   RIB_PDL=0           plain stream order between kernels instead of programmatic dependent launch
 *   RIB_PAIRS=0         the auto-tuner does not try the CTA-pair (tcgen05 cta_group::2) forms
 *   RIB_SPADE256=0      128-column tiles for every SPADE layer (default: 256-column CTA-pair tiles at the low-resolution levels)
 *   RIB_SUBPIX_PPC=4    sub-pixel convs compute up to four output parities per CTA (measured slower, opt-in)
 *   RIB_GATHER=1        stride-2 convs over a normal-layout map gather their parity tiles with cp.async (measured slower, opt-in)
 */
#ifndef RIB_B200_H_
#define RIB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIB_ABI_VERSION 6

/* Message of the last failing call on this host thread ("" if none). */
const char* rib_last_error(void);
int rib_abi_version(void);
/* Total number of rib kernels launched by this process (all entry points). */
long long rib_kernel_launch_count(void);

/* ---- A1: pose rasterisation ------------------------------------------------------------------
 * Replaces HSMAutoDataset._generate_skeleton + _generate_pose_map + to_tensor_norm and the label
 * concatenation (datasets/HSM_auto_dataset.py:205-251, :73-75; utils/keypoint2img.py:36-173;
 * models/evaluator.py:222-229, :250).  Bit-exact.
 *   joints        device  f64 [B][19][3]  (x, y, confidence), model-pixel coordinates
 *   gauss_taps    HOST    f64 [41]        normalised Gaussian taps (sigma 5, radius 20)
 *   label         device  f32 [B][22][H][W]  = [skeleton RGB normalised to [-1,1] | 19 heat-maps]; may be NULL
 *   label_planar  device  16-bit [B][4][H][W][8]: the same label rounded to the generator's activation type
 *                 in its input layout (32 channels, 22..31 zero) (see rib_generator_bind); may be NULL
 *   workspace     device scratch of rib_rasterize_workspace_bytes(B, H, W) bytes, 16-byte aligned
 */
long long rib_rasterize_workspace_bytes(int B, int H, int W);
int rib_rasterize(const double* joints, int B, int H, int W, const double* gauss_taps, double skeleton_thres,
                  double foot_thres, float* label, void* label_planar, void* workspace, long long workspace_bytes,
                  void* stream);

/* ---- A3: flow-based bilinear resampling ------------------------------------------------------
 * No call site in the reference tree; semantics of imaginaire's resample():
 * out = grid_sample(src, identity + flow, bilinear, padding_mode='border', align_corners=True).
 *   src f32 [B][C][H][W], flow [B][2][H][W] (pixels; channel 0 = x) as f32, or as IEEE half when flow_fp16 != 0
 *   (converted exactly on load: halves the upload of a flow field), out f32 [B][C][H][W]
 *   *_bstride: elements between consecutive frames (0 = dense), so that every r-th frame of a clip can be
 *   addressed without a gather copy.
 */
int rib_warp(const float* src, const void* flow, int flow_fp16, float* out, int B, int C, int H, int W,
             long long src_bstride, long long flow_bstride, long long out_bstride, void* stream);

/* ---- A4: mask-blend composite ----------------------------------------------------------------
 * Replaces models/evaluator.py:256-258 (fuse = pred*mask + dain*(1-mask)) and, when out_u8 is not
 * NULL, utils/utils.py:122-147 tensor2images (uint8 HWC frame, truncating).
 *   img, dain f32 [B][3][H][W]; mask f32 [B][1][H][W]; out_f32 f32 [B][3][H][W] (may be NULL);
 *   out_u8 u8 [B][H][W][3] (may be NULL).  mask == NULL: out = img (key frames pass through,
 *   evaluator.py:240-244; dain is ignored).  img / out_f32 / out_u8 take a frame stride in elements
 *   (0 = dense): the batch of AR step s is written to frames s, s + r, s + 2r, ... of the clip.
 */
int rib_composite(const float* img, const float* mask, const float* dain, float* out_f32, uint8_t* out_u8, int B,
                  int H, int W, long long img_bstride, long long out_f32_bstride, long long out_u8_bstride,
                  void* stream);

/* ---- decoded frame -> network input ----------------------------------------------------------
 * Replaces dataset.to_tensor_norm (datasets/HSM_auto_dataset.py:73-75: transforms.ToTensor + Normalize(0.5, 0.5))
 * as applied to the key frames and the DAIN frames in models/evaluator.py:223-224:
 * out = (float32(v) / 255 - 0.5) / 0.5, bit-exact.  Lets a caller upload 8-bit frames (what a PNG decode
 * yields) instead of fp32 tensors.
 *   frames u8 [B][H][W][3]; out f32 [B][3][H][W]; *_bstride: elements between frames (0 = dense).
 *   out_u8 (may be NULL) u8 [B][H][W][3]: tensor2images(out) (utils/utils.py:122-147), the frame the evaluator saves
 *   for a key frame (evaluator.py:240-244, :265-266); it truncates, so it differs from `frames` on 63 of the 256 levels.
 */
int rib_frames_from_u8(const uint8_t* frames, float* out, uint8_t* out_u8, int B, int H, int W, long long in_bstride,
                       long long out_bstride, long long out_u8_bstride, void* stream);

/* Replaces the image half of the evaluator's A.Resize(height, width, interpolation=cv2.INTER_CUBIC)
 * (models/evaluator.py:18-26, applied to every key frame and DAIN frame at :218-220), i.e.
 * cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC) on a uint8 HWC image: OpenCV's separable 4-tap cubic
 * convolution (A = -0.75) in its floating-point form, bit-identical to oracle/resize_oracle.py, which is pinned to
 * cv2 4.13 (IPP) within one level on <= 0.05 % of the pixels.  Equal sizes copy the frame, like cv2.
 *   frames u8 [B][h][w][3] -> out u8 [B][H][W][3]; *_bstride: bytes between frames (0 = dense).
 */
int rib_resize_cubic_u8(const uint8_t* frames, uint8_t* out, int B, int h, int w, int H, int W, long long in_bstride,
                        long long out_bstride, void* stream);

/* ---- A2: generator ---------------------------------------------------------------------------
 * Replaces models.generator.Generator (models/generator.py:35-302), LabelEmbedder (:306-410) and
 * MaskGenerator (:415-510) in eval mode.
 */
typedef struct rib_gen_config {
  int label_nc, img_nc;          /* gen.input_label_nc, gen.input_image_nc */
  int nf, maxf, n_down, n_res;   /* gen.num_filters, max_num_filters, num_downsamples_img, derived num_res_blocks */
  int emb_nf, emb_max, emb_down; /* gen.embed.* */
  int mask_nf, mask_max, mask_down, mask_res; /* gen.mask.* */
} rib_gen_config;

/* One state-dict entry: the reference's key, a device fp32 pointer and the element count. */
typedef struct rib_tensor {
  const char* name;
  const float* data;
  long long numel;
} rib_tensor;

typedef struct rib_generator rib_generator;

/* Folds spectral norm (W / (u . W v), weight_norm.py:84-85), repacks every conv into its K-major
 * 16-bit GEMM operand and keeps the packed copy on the device.  `tensors` must hold the reference
 * state-dict (372 entries for configs/HSM.yaml; unused label_embedding.* / conv_mask.* are ignored).
 * Synchronises `stream` before returning. */
int rib_generator_create(const rib_gen_config* cfg, const rib_tensor* tensors, int n_tensors, void* stream,
                         rib_generator** out);
void rib_generator_destroy(rib_generator* g);

/* Bytes of caller-provided scratch needed by rib_generator_forward for a (B, H, W) batch. */
long long rib_generator_workspace_bytes(rib_generator* g, int B, int H, int W);

/* Builds the launch plan of a (B, H, W) batch on `workspace` (no kernel is launched) and returns the address,
 * inside the workspace, of the generator's label input in its native layout (16-bit [B][4][H][W][8]).  A caller
 * that rasterises on the GPU lets rib_rasterize write `label_planar` there and then passes label = NULL to
 * rib_generator_forward, which skips the fp32 -> 16-bit repack of the label (same values, one pass less). */
int rib_generator_bind(rib_generator* g, int B, int H, int W, void* workspace, long long workspace_bytes,
                       void** label_planar, void* stream);

/* Generator.forward(label, label_prev, img_fake, img_prev) -> (img_final, mask)
 * (models/generator.py:181-234; label_prev is dead in the reference and is not taken).
 *   label f32 [B][22][H][W] (or NULL after rib_generator_bind, see above); img_fake, img_prev f32 [B][3][H][W];
 *   H, W multiples of 16
 *   out_img f32 [B][3][H][W] in (-1,1); out_mask f32 [B][1][H][W] in (0,1)
 * The workspace must stay bound to this generator between calls with the same (B, H, W). */
int rib_generator_forward(rib_generator* g, int B, int H, int W, const float* label, const float* img_fake,
                          const float* img_prev, float* out_img, float* out_mask, void* workspace,
                          long long workspace_bytes, void* stream);

/* ---- upstream motion model ------------------------------------------
 * Replaces Human_Motion_Modelling/models/transformer.py:16-132 (`Transformer`, built by build_transformer :336-350 from
 * configs/config.yaml:77-94) in eval mode for the shipped options: pre_norm, leaky_relu, two_stage, no intermediate
 * outputs.  fp32 throughout. */
typedef struct rib_motion_config {
  int input_joints;     /* transformer.input_joints  (38: 19 joints x (x, y)) */
  int hidden_dim;       /* transformer.hidden_dim    (128) */
  int nheads;           /* transformer.nheads        (8; hidden_dim / nheads must be 16) */
  int dim_feedforward;  /* transformer.dim_feedforward (256) */
  int enc_layers, dec_layers;
} rib_motion_config;

typedef struct rib_motion rib_motion;

/* `tensors`: the reference module's state-dict (oracle/motion_oracle.state_spec lists the 188 keys of the shipped
 * configuration), device fp32.  The parameters are copied; the caller's tensors may be released afterwards. */
int rib_motion_create(const rib_motion_config* cfg, const rib_tensor* tensors, int n_tensors, void* stream, rib_motion** out);
void rib_motion_destroy(rib_motion* m);
long long rib_motion_workspace_bytes(rib_motion* m, int B, int L);

/* Transformer.forward(src, src_mask, src_pos, tgt, tgt_mask, tgt_pos, rate) -> (joints, reco)  (transformer.py:76-112;
 * with two_stage the decoder input is interpolate_embedding(reco, rate), `tgt` is not read and therefore not taken).
 *   src f32 [B][C][L] (C = input_joints, L = rate * n + 1 frames); src_mask / tgt_mask u8 [B][L], non-zero = the frame is
 *   hidden from attention as a key (key_padding_mask; may be NULL); src_pos / tgt_pos f32 [L][B][hidden_dim];
 *   joints, reco f32 [L][B][C].  Asynchronous on `stream`; workspace of rib_motion_workspace_bytes(B, L) bytes. */
int rib_motion_forward(rib_motion* m, int B, int L, const float* src, const uint8_t* src_mask, const float* src_pos,
                       const uint8_t* tgt_mask, const float* tgt_pos, int rate, float* joints, float* reco,
                       void* workspace, long long workspace_bytes, void* stream);

/* ---- measurement hooks (bench.py's roofline pass) ---------------------------------------------
 * While enabled, every launch of the implicit-GEMM convolution kernel is bracketed by CUDA events
 * on its stream; rib_profile_collect() synchronises them and returns the summed device time and
 * the number of launches since the last collect. */
void rib_profile_enable(int enable);
int rib_profile_collect(double* conv_ms, long long* conv_launches);
/* Same, and also copies the device time of each launch (in launch order) into per_launch_ms[0 .. cap). */
int rib_profile_collect_launches(double* conv_ms, long long* conv_launches, float* per_launch_ms, long long cap);

/* ---- bring-up / test hooks (not part of the drop-in surface) ----------------------------------
 * rib_debug_set_simt(1) replaces the tcgen05 main loop by a plain-FMA loop (same tiles and
 * epilogues); used by the tests to bisect tensor-core problems.  Never set by bench.py. */
void rib_debug_set_simt(int enable);
int rib_debug_get_simt(void);
/* Looks up an intermediate activation of the last forward plan by name.  16-bit chunk-planar storage
 * [B][ld/8][H][W][8] (ld = channels of the whole buffer the view is a slice of): element (n, c, y, x) at
 * ptr[n*ld*H*W + (c/8)*H*W*8 + (y*W + x)*8 + c%8].  Returns 0 if found. */
int rib_generator_debug_tensor(rib_generator* g, const char* name, const void** ptr, int* B, int* H, int* W, int* C,
                               int* ld);
/* Text description (one line per planned kernel launch, in launch order) of the plan built by the last
 * rib_generator_forward: layer name, tiling, algorithmic FLOPs.  Joined with ncu launch lists by tools/. */
int rib_generator_plan_text(rib_generator* g, char* buf, long long cap);
/* The launch plan rib_generator_forward would build for (B, H, W), WITHOUT a GPU: validates the configuration of every
 * layer for the shape (tile geometry, shared-memory budget, TMEM columns), applies the imported tuning table, and writes
 * the plan text (as rib_generator_plan_text; buf may be NULL) and the workspace size.  Nothing is allocated or launched. */
int rib_plan_dry_run(const rib_gen_config* cfg, int B, int H, int W, long long* ws_bytes, char* buf, long long cap);
/* The plan-time auto-tuner's log: one line per tuned launch shape of this process (candidate tilings with their
 * measured device times and the choice).  RIB_AUTOTUNE=0 in the environment disables the tuner. */
int rib_tune_log(char* buf, long long cap);
/* The tuning table of this process as text ("<launch-shape key>\t<mt>\t<policy>" per line) and its import, which
 * returns the number of entries read (>= 0).  A table imported before the first forward makes the tiling choices
 * identical across processes / ranks and skips the timing runs for the shapes it lists; the host mirror loads
 * render-in-between_b200/rib/tune_b200.txt (and $RIB_TUNE_FILE, which it also rewrites at exit). */
int rib_tune_export(char* buf, long long cap);
int rib_tune_import(const char* text);
/* 1 if activations are stored as IEEE fp16, 0 for bf16. */
int rib_act_is_fp16(void);
/* Stand-alone launch of the implicit-GEMM convolution for unit tests:
 *   x    16-bit chunk-planar [B][Cin/8][Hin][Win][8]   (Cin 16, 32 or a multiple of 64); for stride 2 the
 *        parity-planar layout [B][Cin/8][py][px][Hin/2][Win/2][8] that a stride-2 layer's producer writes
 *   w    f32 [Cout][Cin][k][k], bias f32 [Cout] (may be NULL), k in {1,3}, stride in {1,2}, pad k/2
 *   out  16-bit chunk-planar [B][Cout/8][Hout][Wout][8] (Cout 16/32/64 or a multiple of 128), act: 0 none, 1 leaky-relu 0.2
 *   stats [B][Cout][2] 64-bit slots (may be NULL; accumulated into): fixed-point integers, sum * 2^32 and sum of
 *         squares * 2^24 (integer atomics make the totals independent of CTA arrival order: bit-reproducible)
 *   scratch: device buffer of at least rib_conv_test_scratch_bytes() bytes */
long long rib_conv_test_scratch_bytes(int Cin, int Cout, int k);
int rib_conv_test(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                  int Cin, int Cout, int k, int stride, int act, void* scratch, void* stream);

/* Same with the two in-kernel fusions of the mask network:
 *   subpix = 1   conv3x3(nearest_x2(x)) in its sub-pixel form (k = 3, stride 1): `out` is the parity-planar
 *                [B][Cout/8][py][px][Hin][Win][8] map of the (2 Hin, 2 Win) result
 *   xf_stats     [B][Cin][2] 64-bit fixed-point slots as above (sum, sum of squares of x per image and channel), xf_w / xf_b f32 [Cin] (may be NULL):
 *                x is a RAW map and the kernel applies lrelu?(instance_norm_affine(x)) to its halo tiles in shared
 *                memory before the MMAs (xf_act = 1: LeakyReLU 0.2) */
int rib_conv_test_ex(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                     int Cin, int Cout, int k, int stride, int act, int subpix, const double* xf_stats, const float* xf_w,
                     const float* xf_b, int xf_act, void* scratch, void* stream);

/* Stand-alone launch of the generator's AvgPool2d(3, stride 2, padding 1) pass (generator.py:203-208: the down-sampling
 * between the SPADE blocks of the encoder) for unit tests:
 *   x     16-bit chunk-planar [B][C/8][H][W][8] (C a multiple of 8, H and W even)
 *   out   16-bit chunk-planar [B][C/8][H/2][W/2][8]: the fp32 window sums divided by 9 (zero padding counts), one rounding
 *   stats [B][C][2] 64-bit fixed-point slots as above (may be NULL; accumulated into): sum / sum of squares of the
 *         un-rounded pooled values, the instance-norm statistics of the next block */
int rib_avgpool_test(const void* x, void* out, double* stats, int B, int H, int W, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIB_B200_H_ */
