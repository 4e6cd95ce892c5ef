// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a) — declarations.
//
// One launch computes, for a batch of NHWC 16-bit activation maps,
//     D[pixel, n] = sum_{tap, c} X0[pixel + tap, c] * Wp[n, (tap, c)]  (+ sum_c X1[pixel, c] * Wp[n, K0 + c])
// with M = 128 output pixels (a TH x TW spatial tile of one image) per CTA, N = BN output columns,
// K stepped as (tap, BK-channel chunk).  Zero padding of the 3x3 window comes from TMA
// out-of-bounds fill; stride-2 convolutions read four parity views of the input.
//
// Epilogues (reference semantics each one replaces):
//   EPI_STORE  bias (+identity residual) -> instance-norm statistics -> activation -> 16-bit NHWC store
//              conv.py:56-69 (conv [+ nonlinearity]) and residual.py:146-151 (shortcut add)
//   EPI_SPADE  the GEMM produces [gamma|beta] = conv1x1(cond); the epilogue applies
//              lrelu?((x - mean) * rstd * (1 + gamma) + beta)   activation_norm.py:211-234
//              (x may be read through a nearest x2 up-sampling, generator.py:249)
//   EPI_FINAL  bias -> tanh / sigmoid -> fp32 NCHW (+ optional 16-bit NHWC copy)  generator.py:228, :484-485
#pragma once
#include "common.cuh"

namespace rib {

enum { EPI_STORE = 0, EPI_SPADE = 1, EPI_FINAL = 2 };
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3 };

struct alignas(64) ConvGemmParams {
  CUtensorMap amap[4];  // stride 1: [0] = X0, [1] = X1 (optional); stride 2: parity views (py*2+px) of X0
  CUtensorMap bmap;     // packed weights [Npad][Ktotal], K-major
  int B, H, W;          // output spatial size
  int TW, TH, tiles_x, tiles_y;
  int BK, BN, stages, n_tiles;
  int ntaps, stride, cchunks0, cchunks1;
  uint32_t idesc;
  int debug_simt;
  // raw views (bring-up mainloop) ------------------------------------------------------------
  const act_t* src0; int ld0; int Hin, Win;
  const act_t* src1; int ld1;
  const act_t* wpk; int ktotal;
  // epilogue ---------------------------------------------------------------------------------
  const float* bias;   // [Npad]
  int n_valid;         // real output columns (<= Npad)
  int act;
  // EPI_STORE
  act_t* out; int ldo;
  const act_t* res; int ldr;
  double* stats;       // [B][n_valid][2] (sum, sum of squares) or null
  // EPI_SPADE
  const act_t* x; int ldx, Hx, Wx, ups;
  const double* xstats;  // [B][C][2]
  int C, CT;
  act_t* outq[2]; int ldq[2]; int actq[2];
  float eps;
  // EPI_FINAL
  float* out_f32;      // [B][n_valid][H][W]
  act_t* out_act; int ld_act;
};

// Host helpers -------------------------------------------------------------------------------------
// 4-D activation view (C, W, H, B) with explicit byte strides; box = (boxC, boxW, boxH, 1).
int make_tmap_act(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, size_t strideW, size_t strideH,
                  size_t strideB, int boxC, int boxW, int boxH);
// 2-D weight view (K, N); box = (boxK, boxN).
int make_tmap_w(CUtensorMap* m, const act_t* w, int K, int N, int boxK, int boxN);
// Chooses the spatial tile (TW x TH = 128) that wastes the fewest pixels.
void choose_tile(int H, int W, int* TW, int* TH);
size_t conv_gemm_smem_bytes(const ConvGemmParams& p);
int launch_conv_gemm(const ConvGemmParams& p, int mode, cudaStream_t stream);
long long conv_gemm_launch_count();
void conv_gemm_profile_enable(int on);
int conv_gemm_profile_collect(double* total_ms, long long* launches);

}  // namespace rib
