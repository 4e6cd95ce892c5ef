// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a) — declarations.
//
// Activations live in HBM as 16-bit "chunk-planar" maps  [B][C/8][H][W][8]  (8 channels = 16 bytes per
// pixel per plane).  One launch computes, for a batch of such maps,
//     D[pixel, n] = sum_{tap, c} X0[pixel + tap, c] * Wp[n, (c-group, tap, c)]  (+ sum_c X1[pixel, c] * Wp[n, K0 + c])
// A CTA is persistent: it walks super-tiles of MT x (16 rows x 8 columns) = MT x 128 output pixels
// (one tcgen05.mma M=128 per sub-tile), N = BN output columns.
//
// A operand: ONE halo tile per channel group, (8+2) x (16*MT+2) pixels x BKc channels, is brought into
// shared memory by TMA as [BKc/8][halo pixel][8] (no swizzle).  In the K-major "interleave" UMMA layout
// a row (pixel) is 16 bytes and an 8-row group is one image row of the tile, so the stride between
// 8-row groups (SBO) is the halo row pitch and the 3x3 taps are nine descriptors that differ only in
// their start address: every input byte crosses L2 -> SM once instead of nine times.  Zero padding is
// TMA out-of-bounds fill.  Stride-2 convolutions load four parity tiles (even/odd rows x even/odd columns) the same
// way from a parity-planar copy of their input.
// B operand: packed weights [Npad][Ktotal] (K-major, 32/64/128-byte swizzle), one [BN][BKc] sub-tile per
// (channel group, tap); kept resident in shared memory for the life of the CTA when they fit, streamed
// through their own ring (decoupled from the A ring: nine weight sub-tiles per halo tile) otherwise.
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the loads and MMAs of
// tile i+1.
//
// Epilogues (reference semantics each one replaces):
//   EPI_STORE  bias (+identity residual) -> instance-norm statistics -> activation -> 16-bit planar store
//              conv.py:56-69 (conv [+ nonlinearity]) and residual.py:146-151 (shortcut add)
//   EPI_SPADE  the GEMM produces [gamma|beta] = conv1x1(cond); the epilogue applies
//              lrelu?((x - mean) * rstd * (1 + gamma) + beta)   activation_norm.py:211-234
//              (x may be read through a nearest x2 up-sampling, generator.py:249)
//              sub-pixel form (ConvGemmParams::subpix) of "nearest x2 -> conv3x3" (generator.py:478-481 up_flow): output
//              pixel (2y+py, 2x+px) only sees the 2x2 low-resolution neighbourhood {y-1+py, y+py} x {x-1+px, x+px} with
//              the 3x3 weights summed per source pixel, so the up-sampled map is never written or read and the MACs
//              drop by 9/4
//   EPI_SPADE2 the same for the two SPADE layers of a res-block that modulate the same x over the same cond map
//              (conv_block_0 and conv_block_s): an N tile holds [gamma0|beta0|gamma1|beta1] of CT = BN/4 channels, so cond
//              and x are read once for both outputs and every MMA carries twice the columns
//   EPI_FINAL  bias -> tanh / sigmoid -> fp32 NCHW (+ optional 16-bit planar copy)  generator.py:228, :484-485
#pragma once
#include "common.cuh"

namespace rib {

enum { EPI_STORE = 0, EPI_SPADE = 1, EPI_FINAL = 2, EPI_SPADE2 = 3 };
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3 };

static constexpr int kTileW = 8;    // output pixels per tile row  (= one 8-row UMMA core-matrix group)
static constexpr int kTileH = 16;   // tile rows per M=128 sub-tile

// A chunk-planar activation map (or a channel slice of one): element (n, c, y, x) is at
//   p[n * bstride + (c / 8) * H * W * 8 + (y * W + x) * 8 + c % 8].
struct PlanarRef {
  act_t* p;
  long long bstride;  // elements between images = (channels of the whole buffer / 8) * H * W * 8
};

struct alignas(64) ConvGemmParams {
  CUtensorMap amap[4];  // stride 1: [0] = X0, [1] = X1 (optional 1x1 source); stride 2: parity views (py*2+px) of X0
  CUtensorMap bmap;     // packed weights, 3-D view (BKc, Npad, Ktotal/BKc)
  int B, H, W;          // output spatial size
  int tiles_x, tiles_y; // super-tiles per image
  int tw, th;           // pixels per sub-tile row / rows per M=128 sub-tile (8 x 16, or 32 x 4 for 1x1 convolutions)
  int MT;               // M=128 sub-tiles per super-tile (stacked vertically): 1 or 2
  int BN, n_tiles;
  int BKc;              // channels per group = per halo tile (16/32/64)
  int stages0, stages1; // channel groups of source 0 / source 1
  int ntaps, stride, halo;
  int subpix;           // 1: 3x3 conv on a nearest-x2 up-sampled map, folded into four 2x2 convs on the LOW-resolution map
                        //    (ntaps = 4; N tile -> output parity (py, px) = ntile / (n_tiles / 4); the output, twice the
                        //    size of H x W, is written parity-planar [plane][py][px][H][W][8])
  int ppc;              // sub-pixel conv: output parities computed per CTA (1, 2 or 4).  With ppc > 1 an N tile is
                        //    [parity par0 | par0 + 1 | ...] x (BN / ppc) output channels (= all channels of the layer): the
                        //    low-resolution halo tile is loaded once for ppc parities, each parity runs its own four taps
                        //    (N = BN / ppc per tcgen05.mma) into its own accumulator columns
  int s2_parity;        // stride 2: 1 = the input is a parity-planar map, 0 = strided parity views of a normal map
  int a_gather;         // stride 2 from a NORMAL map: the kernel's eight extra warps (the XF build) gather the four parity
                        //    tiles of every channel group with 16-byte cp.async copies instead of the strided TMA views, whose
                        //    16-byte elements make the TMA unit the bottleneck (emb_1: 9 400 clk per tile, profiles/r2)
  int a_ring, b_ring;   // ring slots (one channel group each); b_ring is unused
  int b_resident;       // 1: all weight sub-tiles of this CTA's N tile stay in shared memory
  int pair;             // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256): each CTA stages half of the weight rows
  int halo_w;           // pixels per halo-tile row
  uint32_t a_tile_bytes;  // one halo tile (one parity view for stride 2), padded to 128 bytes
  uint32_t a_slot_bytes;  // A bytes per ring slot (1 or 4 tiles)
  uint32_t g_slot_bytes;  // bytes per ring slot: halo tile(s) [+ the group's weight sub-tiles at b_off when streamed]
  uint32_t b_off;         // offset of the streamed weight sub-tiles inside a slot (1024-byte aligned)
  uint32_t b_tap_bytes;   // BN * BKc * 2: one weight sub-tile
  uint32_t a_tx_bytes;    // bytes TMA delivers per halo-tile slot
  uint32_t lbo, sbo;      // UMMA no-swizzle K-major descriptor strides (bytes)
  uint32_t idesc;
  int debug_simt;
  // raw views (bring-up mainloop) ------------------------------------------------------------
  PlanarRef src0; int Hin, Win;
  PlanarRef src1;
  const act_t* wpk; int ktotal;
  // epilogue ---------------------------------------------------------------------------------
  const float* bias;   // [Npad]
  int n_valid;         // real output columns (<= Npad)
  int act;
  // A-operand transform: source 0 holds RAW conv outputs; lrelu?(instance_norm_affine(x)) (the N-A of a C-N-A block,
  // conv.py:56-69) is applied to every halo tile in shared memory between its TMA load and the MMAs, so the
  // normalised map is never written to HBM.  Zero padding stays zero.  Null xf_stats = no transform.
  const double* xf_stats;   // [B][cin0][2] of source 0
  const float* xf_w;        // [cin0] affine weight / bias
  const float* xf_b;
  int xf_act;
  // EPI_STORE
  PlanarRef out;
  PlanarRef res; int has_res;
  int res_ups;                   // 1: `res` is a (H/2, W/2) map read through a nearest x2 up-sampling
  PlanarRef out2; int has_out2;  // optional second copy of the output in parity-planar layout (feeds a stride-2 conv)
  // Two convs over the same input merged along N (one N tile): columns [0, seg_cols) are the first conv (out, stats),
  // [seg_cols, n_valid) the second (out_b, stats_b).  Both multiples of 16; 0 = single output.
  int seg_cols;
  PlanarRef out_b; int out_b_parity;
  double* stats_b; int stats_b_ld;
  double* stats;       // [B][stats_ld][2] (sum, sum of squares; this launch's channel 0 first) or null
  int stats_ld;        // channels per image of the statistics buffer (>= n_valid: several producers may share one)
  int out_parity;      // 1: `out` is parity-planar [plane][py][px][H/2][W/2][8] (its only consumer is a stride-2 conv)
  // EPI_SPADE
  PlanarRef x; int Hx, Wx, ups;
  const double* xstats;  // [B][C][2]
  int C, CT;
  PlanarRef outq[2]; int actq[2];
  float eps;
  // EPI_FINAL
  float* out_f32;      // [B][n_valid][H][W]
  PlanarRef out_act; int out_act_coff; int has_out_act;
};

// Host helpers -------------------------------------------------------------------------------------
// Stride-1 view of a planar map: dims (W*8, H, C/8, B); box = (box_w*8, box_h, box_c/8, 1).
int make_tmap_act_s1(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int box_c,
                     int box_w, int box_h);
// Parity view (py, px) of a PARITY-PLANAR map for stride-2 convolutions.  A stride-2 consumer reads its input
// from the layout [B][C/8][py][px][H/2][W/2][8] (written by its producer, see ConvGemmParams::out2 and
// InApplyParams::out_parity), so that each of the four parity tiles is a dense box with long rows:
// dims (W/2*8, H/2, C/8, B), base = plane + (py*2+px) * (H/2*W/2*8).
int make_tmap_act_s2(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int py, int px,
                     int box_c, int box_w, int box_h);
// Weight view (BKc, Npad, Ktotal/BKc); box = (BKc, BN, taps).
int make_tmap_act_s2_strided(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int py,
                             int px, int box_c, int box_w, int box_h);
int make_tmap_w(CUtensorMap* m, const act_t* w, int K, int N, int bkc, int boxN, int taps);
// Channels per pipeline stage for a layer (also fixes the K ordering of the packed weights).
int choose_bkc(int cin0, int cin1, int taps, int BN, int stride);
// Tiling choices the plan-time auto-tuner may override (0 = the default heuristic):
//   mt      M=128 sub-tiles stacked per super-tile (1 or 2)
//   policy  1 = resident weights, ~100 KB CTAs (several per SM); 2 = resident weights, one CTA per SM, deep halo ring;
//           3 = weights streamed with the halo tiles; 4 = streamed, CTA pairs (cta_group::2): each CTA of a pair
//           stages half of the weight rows of every sub-tile and the leader issues M = 256 MMAs for both
//   ring    policy 2 only: cap of the halo ring (default 4, at most 8); deeper rings help one-CTA-per-SM layers whose
//           pipeline has the extra transform stage
struct ConvTune {
  int mt = 0, policy = 0, ring = 0;
};
// Fills the tiling / pipeline fields of p (everything except tensor maps and epilogue pointers).  Returns non-zero
// when the requested tuning does not fit the layer.
int conv_gemm_configure(ConvGemmParams* p, int B, int Hout, int Wout, int cin0, int cin1, int taps, int stride,
                        int BN, int n_pad, const ConvTune* tune = nullptr, int ppc = 1);
size_t conv_gemm_smem_bytes(const ConvGemmParams& p);
int launch_conv_gemm(const ConvGemmParams& p, int mode, cudaStream_t stream);
int device_sm_count(int dev);   // multiprocessors of device `dev` (cached)
long long conv_gemm_launch_count();
void conv_gemm_profile_enable(int on);
int conv_gemm_profile_collect(double* total_ms, long long* launches, float* per_launch_ms, long long cap);

}  // namespace rib
