// Pose rasterisation (bit-exact restatement of the reference's CPU rasteriser, on the GPU).
//
//   heat-maps  HSMAutoDataset._generate_pose_map   PGNR/datasets/HSM_auto_dataset.py:205-236
//              (scipy.ndimage.gaussian_filter sigma=5, mode='reflect', then map / map.max(), fp64 -> fp32)
//   skeleton   HSMAutoDataset._generate_skeleton   PGNR/datasets/HSM_auto_dataset.py:238-251
//              connect_keypoints / interpPoints / drawEdge / setColor  PGNR/utils/keypoint2img.py:36-148
//   to-tensor  ToTensor + Normalize(0.5, 0.5)      PGNR/datasets/HSM_auto_dataset.py:73-75
//
// Output: fp32 label [B, 22, H, W] = [skeleton(3) | heat-maps(19)] (models/evaluator.py:250).
// All fp64 arithmetic uses explicit round-to-nearest intrinsics so that no FMA contraction can
// change a rounding with respect to numpy / scipy's C loops.
#include "raster.cuh"

namespace rib {

static constexpr int kRadius = 20;  // int(4.0 * 5 + 0.5)
static constexpr int kTaps = 41;

struct GaussTable {
  double w[kTaps];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {  // ndimage 'reflect': d c b a | a b c d | d c b a
  const int period = 2 * n;
  int m = i % period;
  if (m < 0) m += period;
  return m >= n ? period - 1 - m : m;
}

// One block per (joint, frame): the response of a unit impulse is non-zero only in the 41x41 window.
__global__ void __launch_bounds__(256) heatmap_kernel(const double* __restrict__ joints, GaussTable tab, double thres,
                                                      float* __restrict__ label, int H, int W, int njoints) {
  const int j = blockIdx.x, b = blockIdx.y;
  const double* jp = joints + ((size_t)b * njoints + j) * 3;
  const double x = jp[0], y = jp[1], c = jp[2];
  if (!(x >= 0.0 && y >= 0.0 && c > thres && x < (double)W && y < (double)H)) return;  // map stays all-zero
  const int iy = (int)y, ix = (int)x;
  __shared__ double s_ky[kTaps];
  __shared__ double s_max[8];
  // axis-0 pass: ky[o] for o in [iy-20, iy+20]; centre tap first, then pairs from the outside in.
  if (threadIdx.x < kTaps) {
    const int o = iy - kRadius + threadIdx.x;
    double tmp = 0.0;
    if (o >= 0 && o < H) {
      tmp = (o == iy) ? tab.w[kRadius] : 0.0;
      for (int ii = 0; ii < kRadius; ++ii) {
        const double a = reflect_idx(o - kRadius + ii, H) == iy ? 1.0 : 0.0;
        const double bb = reflect_idx(o + kRadius - ii, H) == iy ? 1.0 : 0.0;
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, bb), tab.w[ii]));
      }
    }
    s_ky[threadIdx.x] = tmp;
  }
  __syncthreads();
  // axis-1 pass over the window; each thread keeps up to 7 pixels in registers.
  double g[7];
  double lmax = 0.0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    g[k] = -1.0;
    const int idx = threadIdx.x + k * 256;
    if (idx >= kTaps * kTaps) continue;
    const int wy = idx / kTaps, wx = idx - wy * kTaps;
    const int oy = iy - kRadius + wy, ox = ix - kRadius + wx;
    if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
    const double ky = s_ky[wy];
    double tmp = __dmul_rn(ox == ix ? ky : 0.0, tab.w[kRadius]);
    for (int ii = 0; ii < kRadius; ++ii) {
      const bool ha = reflect_idx(ox - kRadius + ii, W) == ix;
      const bool hb = reflect_idx(ox + kRadius - ii, W) == ix;
      if (ha || hb) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(ha ? ky : 0.0, hb ? ky : 0.0), tab.w[ii]));
    }
    g[k] = tmp;
    lmax = fmax(lmax, tmp);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = lmax;
  __syncthreads();
  double gmax = s_max[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) gmax = fmax(gmax, s_max[k]);
  float* out = label + ((size_t)b * (3 + njoints) + 3 + j) * (size_t)H * W;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (g[k] < 0.0) continue;
    const int idx = threadIdx.x + k * 256;
    const int wy = idx / kTaps, wx = idx - wy * kTaps;
    const int oy = iy - kRadius + wy, ox = ix - kRadius + wx;
    out[(size_t)oy * W + ox] = __double2float_rn(__ddiv_rn(g[k], gmax));
  }
}

// ---------------------------------------------------------------------------------------------
// Skeleton.  The reference draws 18 limbs in order; each limb is 64 shifted copies of an integer
// curve followed by 193 two-point end-cap stamps.  A stamp either paints its pixels (if none of
// them was ever touched) or averages all of them with the colour, so stamps must be replayed in
// order.  One CTA replays the whole frame with a visited bitmap of the full image in shared
// memory and keeps the pixel values of its own band of rows; several CTAs (bands) per frame.
// ---------------------------------------------------------------------------------------------
__constant__ int c_edges[18][2] = {{0, 1}, {1, 8}, {1, 2}, {2, 3}, {3, 4}, {1, 5}, {5, 6}, {6, 7}, {8, 9},
                                   {9, 10}, {10, 11}, {8, 12}, {12, 13}, {13, 14}, {4, 18}, {7, 17}, {11, 16}, {14, 15}};
__constant__ unsigned char c_colors[18][3] = {{153, 0, 51}, {153, 0, 0}, {153, 51, 0}, {153, 102, 0}, {153, 153, 0},
                                              {102, 153, 0}, {51, 153, 0}, {0, 153, 0}, {0, 153, 51}, {0, 153, 102},
                                              {0, 153, 153}, {0, 102, 153}, {0, 51, 153}, {0, 0, 153}, {208, 208, 0},
                                              {0, 208, 0}, {0, 208, 208}, {0, 0, 208}};

__device__ __forceinline__ uint32_t avg_color(uint32_t old, uint32_t col) {
  // per channel (v + c) >> 1 on packed 0x00BBGGRR; channels are <= 255 so (v + c) fits in 9 bits
  const uint32_t r = (((old & 0xffu) + (col & 0xffu)) >> 1);
  const uint32_t g = ((((old >> 8) & 0xffu) + ((col >> 8) & 0xffu)) >> 1);
  const uint32_t b = ((((old >> 16) & 0xffu) + ((col >> 16) & 0xffu)) >> 1);
  return r | (g << 8) | (b << 16);
}

__global__ void __launch_bounds__(256) skeleton_kernel(const double* __restrict__ joints, double thres,
                                                       double foot_thres, float* __restrict__ label, int H, int W,
                                                       int njoints, int band_rows) {
  extern __shared__ uint32_t s_mem[];
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * band_rows;
  const int rows = min(band_rows, H - row0);
  const int nwords = (H * W + 31) / 32;
  uint32_t* s_bits = s_mem;                   // visited bitmap, whole image
  uint32_t* s_val = s_bits + nwords;          // packed RGB of rows [row0, row0+rows)
  int* s_px = reinterpret_cast<int*>(s_val + (size_t)band_rows * W);
  int* s_py = s_px + max(H, W);
  __shared__ double s_pts[19][2];
  __shared__ float s_lut[256];

  for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_bits[i] = 0u;
  for (int i = threadIdx.x; i < rows * W; i += blockDim.x) s_val[i] = 0u;
  // ToTensor: float32(v) / 255, Normalize: (t - 0.5) / 0.5, both in fp32
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)i, 255.0f), 0.5f), 0.5f);
  if (threadIdx.x < njoints) {  // extract_valid_keypoints (keypoint2img.py:114-130)
    const double* jp = joints + ((size_t)b * njoints + threadIdx.x) * 3;
    const int i = threadIdx.x;
    const double thr = (i >= 8 && i <= 16) ? foot_thres : thres;
    const double x = jp[0], y = jp[1], c = jp[2];
    const bool ok = x >= 0.0 && y >= 0.0 && c > thr && x < (double)W && y < (double)H;
    s_pts[i][0] = ok ? x : 0.0;
    s_pts[i][1] = ok ? y : 0.0;
  }
  __syncthreads();

  for (int e = 0; e < 18; ++e) {
    const double xa = s_pts[c_edges[e][0]][0], ya = s_pts[c_edges[e][0]][1];
    const double xb = s_pts[c_edges[e][1]][0], yb = s_pts[c_edges[e][1]][1];
    if (xa == 0.0 || xb == 0.0) continue;  // `0 not in x` (keypoint2img.py:144); uniform across the block
    // interpPoints: major axis = the one with the larger extent (keypoint2img.py:67-70)
    const bool swap = fabs(xa - xb) < fabs(ya - yb);
    double m0 = swap ? ya : xa, m1 = swap ? yb : xb;
    const double n0 = swap ? xa : ya, n1 = swap ? xb : yb;
    if (m0 == m1) continue;                       // zero-length: int(x1 - x0) == 0 -> empty curve
    const double slope = __ddiv_rn(__dsub_rn(n1, n0), __dsub_rn(m1, m0));
    const double icpt = __dsub_rn(n0, __dmul_rn(slope, m0));
    if (m0 > m1) {
      const double t = m0;
      m0 = m1;
      m1 = t;
    }
    const int npts = (int)__dsub_rn(m1, m0);
    if (npts <= 0) continue;
    const double start = (double)(int)m0, stop = (double)(int)m1;
    const double step = npts > 1 ? __ddiv_rn(__dsub_rn(stop, start), (double)(npts - 1)) : 0.0;
    for (int k = threadIdx.x; k < npts; k += blockDim.x) {  // numpy.linspace: k * step + start, last = stop
      double cm = __dadd_rn(__dmul_rn((double)k, step), start);
      if (npts > 1 && k == npts - 1) cm = stop;
      const double cn = __dadd_rn(__dmul_rn(slope, cm), icpt);
      const int im = (int)cm, in_ = (int)cn;  // astype(int): truncation toward zero
      s_px[k] = swap ? in_ : im;
      s_py[k] = swap ? im : in_;
    }
    const uint32_t col = (uint32_t)c_colors[e][0] | ((uint32_t)c_colors[e][1] << 8) | ((uint32_t)c_colors[e][2] << 16);
    __syncthreads();
    // ---- body: 64 shifted copies of the curve (drawEdge, keypoint2img.py:51-55) ----
    for (int i = -4; i < 4; ++i) {
      for (int j = -4; j < 4; ++j) {
        uint32_t old[4];
        int touched = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int k = threadIdx.x + r * 256;
          old[r] = 0u;
          if (k < npts) {
            const int yy = min(max(s_py[k] + i, 0), H - 1), xx = min(max(s_px[k] + j, 0), W - 1);
            const int p = yy * W + xx;
            touched |= (s_bits[p >> 5] >> (p & 31)) & 1u;
            if (yy >= row0 && yy < row0 + rows) old[r] = s_val[(yy - row0) * W + xx];
          }
        }
        const int any = __syncthreads_or(touched);  // setColor's `(im[yy, xx] == 0).all()` over the whole stamp
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int k = threadIdx.x + r * 256;
          if (k < npts) {
            const int yy = min(max(s_py[k] + i, 0), H - 1), xx = min(max(s_px[k] + j, 0), W - 1);
            const int p = yy * W + xx;
            atomicOr(&s_bits[p >> 5], 1u << (p & 31));
            if (yy >= row0 && yy < row0 + rows) s_val[(yy - row0) * W + xx] = any ? avg_color(old[r], col) : col;
          }
        }
        __syncthreads();
      }
    }
    // ---- end caps: 193 two-point stamps (keypoint2img.py:59-64), replayed by one thread ----
    if (threadIdx.x == 0) {
      const int ex[2] = {s_px[0], s_px[npts - 1]}, ey[2] = {s_py[0], s_py[npts - 1]};
      for (int i = -12; i < 12; ++i) {
        for (int j = -12; j < 12; ++j) {
          if (i * i + j * j >= 64) continue;
          int p[2], yy[2], xx[2];
          uint32_t old[2];
          int any = 0;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            yy[t] = min(max(ey[t] + i, 0), H - 1);
            xx[t] = min(max(ex[t] + j, 0), W - 1);
            p[t] = yy[t] * W + xx[t];
            any |= (s_bits[p[t] >> 5] >> (p[t] & 31)) & 1u;
            old[t] = (yy[t] >= row0 && yy[t] < row0 + rows) ? s_val[(yy[t] - row0) * W + xx[t]] : 0u;
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            s_bits[p[t] >> 5] |= 1u << (p[t] & 31);
            if (yy[t] >= row0 && yy[t] < row0 + rows)
              s_val[(yy[t] - row0) * W + xx[t]] = any ? avg_color(old[t], col) : col;
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- write the band: uint8 -> normalised fp32, channels 0..2 of the label ----
  float* out = label + (size_t)b * (3 + njoints) * (size_t)H * W;
  for (int c = 0; c < 3; ++c) {
    float* oc = out + (size_t)c * H * W + (size_t)row0 * W;
    for (int i = threadIdx.x; i < rows * W; i += blockDim.x) oc[i] = s_lut[(s_val[i] >> (8 * c)) & 0xffu];
  }
}

int launch_rasterize(const double* joints_dev, int B, int H, int W, const double* wtab41_host, double skeleton_thres,
                     double foot_thres, float* label, cudaStream_t stream) {
  const int njoints = 19;
  RIB_REQUIRE(H >= 1 && W >= 1 && B >= 1, "rasterize: bad shape");
  RIB_REQUIRE(H <= 1024 && W <= 1024, "rasterize: images larger than 1024 px are not supported");
  GaussTable tab;
  for (int i = 0; i < kTaps; ++i) tab.w[i] = wtab41_host[i];
  // heat-map channels are zero outside the 41x41 windows
  RIB_CHECK_CUDA(cudaMemsetAsync(label, 0, (size_t)B * (3 + njoints) * H * W * sizeof(float), stream));
  heatmap_kernel<<<dim3(njoints, B), 256, 0, stream>>>(joints_dev, tab, skeleton_thres, label, H, W, njoints);
  RIB_CHECK_CUDA(cudaGetLastError());

  const size_t bitmap_bytes = (size_t)((H * W + 31) / 32) * 4;
  const size_t fixed = bitmap_bytes + (size_t)2 * (H > W ? H : W) * 4;
  const size_t budget = 200 * 1024;
  RIB_REQUIRE(fixed + (size_t)W * 4 <= budget, "rasterize: image does not fit the shared-memory plan");
  int band_rows = (int)((budget - fixed) / ((size_t)W * 4));
  if (band_rows > H) band_rows = H;
  const int bands = ceil_div(H, band_rows);
  band_rows = ceil_div(H, bands);
  const size_t smem = fixed + (size_t)band_rows * W * 4;
  static size_t attr_smem = 0;  // the kernel also has ~1.4 KB of static shared memory
  if (smem > attr_smem) {
    RIB_CHECK_CUDA(cudaFuncSetAttribute(skeleton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  skeleton_kernel<<<dim3(bands, B), 256, smem, stream>>>(joints_dev, skeleton_thres, foot_thres, label, H, W, njoints,
                                                         band_rows);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rib
