// Bandwidth-bound kernels of the rendering path (declarations).
#pragma once
#include "common.cuh"

namespace rib {

int device_sm_count(int dev);   // conv_gemm.cu


// All 16-bit activation maps are chunk-planar: [B][Ctot/8][H][W][8] (see conv_gemm.cuh).  A map (or a
// channel slice of one that starts on a multiple of 8) is passed as the pointer to its first plane
// plus `bstride`, the element distance between images (= Ctot/8 * H * W * 8).

// NCHW fp32 -> planar 16-bit.  Up to three sources are concatenated along channels and land in channels
// [c_off, c_off + sum C) of the destination; planes [plane0, plane0 + nplanes) are fully written
// (zero where no source channel maps), so padding channels need no separate memset.
struct PackSrc {
  const float* p;
  int C;
};
int launch_pack_nchw(const PackSrc* srcs, int nsrc, act_t* dst, long long dst_bstride, int c_off, int plane0,
                     int nplanes, int B, int H, int W, cudaStream_t s);

// Both image inputs of the generator in one pass (generator.py:197 and :232): img_fake, img_prev fp32 [B,3,H,W] ->
//   emb plane 0  = [fake(3), prev(3), 0, 0]          (cat([img_fake, img_prev]) for ref_embedding)
//   mask plane 0 = [prev(3), fake(3), 0, 0]          (cat([img_prev, img_fake, img_final]); img_final is written into
//                                                     channels 6..8 by conv_img's epilogue)
// Four pixels per thread (16-byte loads, 64-byte stores).  The all-zero second planes are not touched.
int launch_pack_images(const float* fake, const float* prev, act_t* emb, long long emb_bstride, act_t* mask,
                       long long mask_bstride, int B, int H, int W, cudaStream_t s);

// Instance-norm (affine) application for the C-N-A blocks of the mask network
// (conv.py:56-69 with order 'CNA'; residual.py:146-151 for the two-term form):
//   out = act(IN_a(a)) [+ IN_b(b) | + b]   with optional nearest x2 up-sampling on the store
struct InApplyParams {
  const act_t* a; long long a_bs; const double* astats; const float* aw; const float* ab;
  const act_t* b; long long b_bs; const double* bstats; const float* bw; const float* bb;  // b optional
  act_t* out; long long o_bs;
  int B, H, W, C;  // input spatial size
  int act;         // 0 none, 1 leaky-relu(0.2) on the first term
  int ups;         // 1: write each pixel to the 2x2 block of a (2H, 2W) output
  int in_parity;   // 1: `a` is a parity-planar (H, W) map (the output of a sub-pixel conv); the output is a normal map
  int out_parity;  // 1: write the output in parity-planar layout [plane][py][px][H/2][W/2][8] (feeds a stride-2 conv)
  float eps;
};
int launch_in_apply(const InApplyParams& p, cudaStream_t s);

// AvgPool2d(3, stride 2, pad 1, count_include_pad) (generator.py:127,208) + instance-norm statistics
// of the pooled map (sum, sum of squares per (n, c), fp64 atomics).
int launch_avgpool3s2(const act_t* src, long long src_bs, act_t* dst, long long dst_bs, double* stats, int B, int H,
                      int W, int C, cudaStream_t s);

// fuse = img * mask + dain * (1 - mask)   (evaluator.py:256-258); optional uint8 HWC frame
// (utils.py:137-142: clip(x*0.5+0.5, 0, 1)*255 truncated, float64 arithmetic).
// mask == nullptr: out = img (pure conversion of key frames).  *_bstride: elements between consecutive frames of
// img / out_f32 / out_u8 (0 = dense), so a batch can be read from / written into every r-th frame of a clip.
int launch_composite(const float* img, const float* mask, const float* dain, float* out_f32, uint8_t* out_u8, int B,
                     int H, int W, long long img_bstride, long long f32_bstride, long long u8_bstride, cudaStream_t s);

// out = bilinear sample of src at (x + flow_x, y + flow_y), border padding, align_corners=True.
int launch_resize_cubic_u8(const uint8_t* in, uint8_t* out, int B, int h, int w, int H, int W, long long in_bstride,
                           long long out_bstride, cudaStream_t s);
int launch_frames_from_u8(const uint8_t* in, float* out, uint8_t* out_u8, int B, int H, int W, long long in_bstride,
                          long long out_bstride, long long u8_bstride, cudaStream_t s);
int launch_warp(const float* src, const void* flow, int flow_fp16, float* out, int B, int C, int H, int W,
                long long src_bstride, long long flow_bstride, long long out_bstride, cudaStream_t s);

// sigma_inv[0] = 1 / (u . (W v)),  W = [Cout, K] fp32 (torch.nn.utils.spectral_norm, eval mode).
int launch_sn_sigma_inv(const float* w, const float* u, const float* v, int Cout, int K, float* sigma_inv,
                        double* scratch, cudaStream_t s);

// Repack a conv weight [Cout][Cin][taps] fp32 into the K-major 16-bit GEMM operand:
//   dst[row(co) * ktotal + koff + ((ci / bkc) * taps + tap) * bkc + ci % bkc] = w[co][ci][tap] * (sigma_inv ? *sigma_inv : 1)
// (bkc = channels per pipeline stage of the layer: all taps of one channel group are contiguous in K)
//   bias_dst[row(co)] (+)= bias[co] (+ 1 for SPADE gamma rows)
// row(co) = row_off + co for plain convs; for SPADE ([gamma(C) | beta(C)] -> per-tile [gamma(CT) | beta(CT)]):
//   c = co % C, half = co / C, row = row_off + (c / CT) * 2 * CT + half * CT + c % CT.
struct PackWeightParams {
  const float* w; const float* bias; const float* sigma_inv;
  int Cout, Cin, taps;
  act_t* dst; float* bias_dst;
  int ktotal, koff, bkc, row_off;
  int spade_C, spade_CT;  // 0 for plain convs
  int spade_nq, spade_q;  // SPADE outputs sharing an N tile (0/1: one) and this conv's index among them: with nq outputs a
                          // tile is [gamma_0|beta_0|...|gamma_{nq-1}|beta_{nq-1}] of CT channels each,
                          // row = row_off + (c / CT) * 2 * nq * CT + q * 2 * CT + half * CT + c % CT
  int bias_accumulate;    // add into bias_dst instead of overwriting (fused shortcut)
  // Sub-pixel form of "nearest x2 -> conv3x3" (taps == 9 in the source): emit the 2x2 kernel of output parity
  // (py, px) = (subpix_parity >> 1, subpix_parity & 1), K order (channel group, tap (a, b), channel) with 4 taps:
  //   w2[a][b] = sum of w[r][s] over r in R(py, a), s in R(px, b);  R(0,0)={0} R(0,1)={1,2} R(1,0)={0,1} R(1,1)={2}
  int subpix, subpix_parity;
};
int launch_pack_weight(const PackWeightParams& p, cudaStream_t s);

}  // namespace rib
