// Pose rasterisation (bit-exact restatement of the reference's CPU rasteriser, on the GPU).
//
//   heat-maps  HSMAutoDataset._generate_pose_map   PGNR/datasets/HSM_auto_dataset.py:205-236
//              (scipy.ndimage.gaussian_filter sigma=5, mode='reflect', then map / map.max(), fp64 -> fp32)
//   skeleton   HSMAutoDataset._generate_skeleton   PGNR/datasets/HSM_auto_dataset.py:238-251
//              connect_keypoints / interpPoints / drawEdge / setColor  PGNR/utils/keypoint2img.py:36-148
//   to-tensor  ToTensor + Normalize(0.5, 0.5)      PGNR/datasets/HSM_auto_dataset.py:73-75
//
// Output: fp32 label [B, 22, H, W] = [skeleton(3) | heat-maps(19)] (models/evaluator.py:250) and / or the
// same label as the generator's 16-bit chunk-planar input [B][32/8][H][W][8] (channels 22..31 zero).
// All fp64 arithmetic uses explicit round-to-nearest intrinsics so that no FMA contraction can
// change a rounding with respect to numpy / scipy's C loops.
//
// Skeleton algorithm.  The reference draws 18 limbs in order; each limb is 64 shifted copies ("stamps") of
// an integer curve followed by 193 two-point end-cap stamps.  setColor paints a stamp's pixels with the
// colour if NONE of them was touched before, otherwise every pixel of the stamp becomes (v + c) >> 1.
// Because a touched pixel stays touched, a pixel's final value is a function of the ORDERED list of stamps
// that touch it and of one global bit per stamp ("did this stamp hit anything older?"), which only matters
// for the pixel's first stamp.  So instead of replaying stamps serially, three data-parallel passes:
//   prep   per frame: the integer curve of every limb as a lookup table f_e[major coord] -> minor coord
//   mark   per pixel: enumerate the stamps touching it in order; every stamp after the first one is, by
//          definition, a stamp that hit an older pixel -> set its flag.  The pixel's colour only depends on the flag
//          of its FIRST stamp, so both outcomes (first stamp painted as colour / as colour >> 1, then the averaging
//          chain) are computed here and parked with the first stamp's id in one 64-bit word per pixel
//   paint  per pixel: pick the outcome by the first stamp's flag; heat-map channels; all stores
// A stamp (i, j) of a limb touches pixel (y, x) iff f[x - j] == y - i (roles of x / y swapped for steep limbs):
// 8 table look-ups per limb give the 64-bit set of body stamps, whose bit order is the stamp order.
#include "raster.cuh"

namespace rib {

static constexpr int kRadius = 20;  // int(4.0 * 5 + 0.5)
static constexpr int kTaps = 41;
static constexpr int kJoints = 19;
static constexpr int kEdges = 18;
static constexpr int kWin = kTaps * kTaps;
static constexpr int kBodyKeys = 64;            // (i + 4) * 8 + (j + 4)
static constexpr int kCapKeys = 24 * 24;        // 64 + (i + 12) * 24 + (j + 12)
static constexpr int kKeys = kBodyKeys + kCapKeys;
static constexpr int kMaxDim = 1024;

struct GaussTable {
  double w[kTaps];
};

struct EdgeMeta {
  int valid, swap;      // swap: the major axis is y (keypoint2img.py:67-70)
  int mlo, mhi;         // first / last major coordinate of the curve
  int ex0, ey0, ex1, ey1;
  int bx0, bx1, by0, by1;  // bounding box of everything the limb can touch (curve +- 12, clipped later)
  float a, b;           // minor ~= a * major + b (culling only)
};

struct JointMeta {
  int valid, ix, iy, pad;
};

// Per-frame scratch (device): everything `paint` needs.
struct FrameScratch {
  short f[kEdges][kMaxDim];
  EdgeMeta edge[kEdges];
  JointMeta joint[kJoints];
  unsigned char flag[kEdges][kKeys];         // end-cap stamps (keys >= kBodyKeys): "hit an older pixel"
  unsigned long long bodyflag[kEdges];       // the same flag of the 64 body stamps of a limb, bit = key
  float win[kJoints][kWin];
};

// Workspace = B FrameScratch records, then one 64-bit word per pixel in which `mark` leaves what `paint` needs
// (see raster_mark_kernel).
static size_t scratch_bytes(int B) { return ((size_t)B * sizeof(FrameScratch) + 255) & ~(size_t)255; }
long long raster_workspace_bytes(int B, int H, int W) {
  return (long long)(scratch_bytes(B) + (size_t)B * H * W * sizeof(unsigned long long));
}

__device__ __forceinline__ int reflect_idx(int i, int n) {  // ndimage 'reflect': d c b a | a b c d | d c b a
  const int period = 2 * n;
  int m = i % period;
  if (m < 0) m += period;
  return m >= n ? period - 1 - m : m;
}

__constant__ int c_edges[kEdges][2] = {{0, 1}, {1, 8}, {1, 2}, {2, 3}, {3, 4}, {1, 5}, {5, 6}, {6, 7}, {8, 9},
                                       {9, 10}, {10, 11}, {8, 12}, {12, 13}, {13, 14}, {4, 18}, {7, 17}, {11, 16}, {14, 15}};
__constant__ unsigned char c_colors[kEdges][3] = {{153, 0, 51}, {153, 0, 0}, {153, 51, 0}, {153, 102, 0}, {153, 153, 0},
                                                  {102, 153, 0}, {51, 153, 0}, {0, 153, 0}, {0, 153, 51}, {0, 153, 102},
                                                  {0, 153, 153}, {0, 102, 153}, {0, 51, 153}, {0, 0, 153}, {208, 208, 0},
                                                  {0, 208, 0}, {0, 208, 208}, {0, 0, 208}};

// ---------------------------------------------------------------------------------------------
// prep: grid (19 + 18, B).  blockIdx.x < 19: the 41x41 response window of joint blockIdx.x (already divided
// by its maximum); blockIdx.x >= 19: the table of limb blockIdx.x - 19.
// ---------------------------------------------------------------------------------------------
__device__ void heat_window(const double* __restrict__ jp, const GaussTable& tab, double thres, FrameScratch* fs, int j,
                            int H, int W) {
  const double x = jp[0], y = jp[1], c = jp[2];
  const bool ok = x >= 0.0 && y >= 0.0 && c > thres && x < (double)W && y < (double)H;
  if (threadIdx.x == 0) {
    JointMeta m;
    m.valid = ok ? 1 : 0;
    m.ix = ok ? (int)x : 0;
    m.iy = ok ? (int)y : 0;
    m.pad = 0;
    fs->joint[j] = m;
  }
  if (!ok) return;  // map stays all-zero
  const int iy = (int)y, ix = (int)x;
  __shared__ double s_ky[kTaps];
  __shared__ unsigned long long s_hit[kTaps];
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
    // which taps of the axis-1 pass see the joint's column from window column threadIdx.x (the pattern does not depend on
    // the row): bits 0-19 first operand of tap ii, bits 20-39 second operand, bit 40 the centre tap
    const int ox = ix - kRadius + threadIdx.x;
    unsigned long long hm = ox == ix ? 1ull << 40 : 0ull;
    if (ox >= 0 && ox < W) {
      for (int ii = 0; ii < kRadius; ++ii) {
        if (reflect_idx(ox - kRadius + ii, W) == ix) hm |= 1ull << ii;
        if (reflect_idx(ox + kRadius - ii, W) == ix) hm |= 1ull << (20 + ii);
      }
    }
    s_hit[threadIdx.x] = hm;
  }
  __syncthreads();
  // axis-1 pass over the window; each thread keeps up to 7 pixels in registers.
  double g[7];
  double lmax = 0.0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    g[k] = -1.0;
    const int idx = threadIdx.x + k * 256;
    if (idx >= kWin) continue;
    const int wy = idx / kTaps, wx = idx - wy * kTaps;
    const int oy = iy - kRadius + wy, ox = ix - kRadius + wx;
    if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
    const double ky = s_ky[wy];
    const unsigned long long hm = s_hit[wx];
    double tmp = __dmul_rn((hm >> 40) & 1ull ? ky : 0.0, tab.w[kRadius]);
    uint32_t taps = ((uint32_t)hm | (uint32_t)(hm >> 20)) & 0xfffffu;   // taps with a hit, visited in ascending order
    while (taps) {                                                       // (a tap without a hit adds an exact zero)
      const int ii = __ffs((int)taps) - 1;
      taps &= taps - 1u;
      const bool ha = (hm >> ii) & 1ull, hb = (hm >> (20 + ii)) & 1ull;
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(ha ? ky : 0.0, hb ? ky : 0.0), tab.w[ii]));
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
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (g[k] < 0.0) continue;
    fs->win[j][threadIdx.x + k * 256] = __double2float_rn(__ddiv_rn(g[k], gmax));
  }
}

__device__ void limb_tables(const double* __restrict__ joints, double thres, double foot_thres, FrameScratch* fs, int H,
                            int W, int e) {
  __shared__ double s_pts[kJoints][2];
  for (int i = threadIdx.x; i < kKeys; i += blockDim.x) fs->flag[e][i] = 0;
  if (threadIdx.x == 0) fs->bodyflag[e] = 0ull;
  if (threadIdx.x < kJoints) {  // extract_valid_keypoints (keypoint2img.py:114-130)
    const double* jp = joints + threadIdx.x * 3;
    const int i = threadIdx.x;
    const double thr = (i >= 8 && i <= 16) ? foot_thres : thres;
    const double x = jp[0], y = jp[1], c = jp[2];
    const bool ok = x >= 0.0 && y >= 0.0 && c > thr && x < (double)W && y < (double)H;
    s_pts[i][0] = ok ? x : 0.0;
    s_pts[i][1] = ok ? y : 0.0;
  }
  __syncthreads();
  {
    EdgeMeta m;
    memset(&m, 0, sizeof(m));
    const double xa = s_pts[c_edges[e][0]][0], ya = s_pts[c_edges[e][0]][1];
    const double xb = s_pts[c_edges[e][1]][0], yb = s_pts[c_edges[e][1]][1];
    bool draw = !(xa == 0.0 || xb == 0.0);  // `0 not in x` (keypoint2img.py:144); uniform across the block
    // interpPoints: major axis = the one with the larger extent (keypoint2img.py:67-70)
    const bool swap = fabs(xa - xb) < fabs(ya - yb);
    double m0 = swap ? ya : xa, m1 = swap ? yb : xb;
    const double n0 = swap ? xa : ya, n1 = swap ? xb : yb;
    if (m0 == m1) draw = false;  // zero-length: int(x1 - x0) == 0 -> empty curve
    double slope = 0.0, icpt = 0.0;
    int npts = 0;
    double start = 0.0, stop = 0.0, step = 0.0;
    if (draw) {
      slope = __ddiv_rn(__dsub_rn(n1, n0), __dsub_rn(m1, m0));
      icpt = __dsub_rn(n0, __dmul_rn(slope, m0));
      if (m0 > m1) {
        const double t = m0;
        m0 = m1;
        m1 = t;
      }
      npts = (int)__dsub_rn(m1, m0);
      if (npts <= 0) draw = false;
      start = (double)(int)m0;
      stop = (double)(int)m1;
      step = npts > 1 ? __ddiv_rn(__dsub_rn(stop, start), (double)(npts - 1)) : 0.0;
    }
    if (!draw) {
      if (threadIdx.x == 0) fs->edge[e] = m;
      return;
    }
    const int dim = swap ? H : W;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) fs->f[e][i] = -1;
    __syncthreads();
    for (int k = threadIdx.x; k < npts; k += blockDim.x) {  // numpy.linspace: k * step + start, last = stop
      double cm = __dadd_rn(__dmul_rn((double)k, step), start);
      if (npts > 1 && k == npts - 1) cm = stop;
      const double cn = __dadd_rn(__dmul_rn(slope, cm), icpt);
      const int im = (int)cm, in_ = (int)cn;  // astype(int): truncation toward zero
      if (im >= 0 && im < dim) fs->f[e][im] = (short)in_;
    }
    if (threadIdx.x == 0) {
      const int mlo = (int)start, mhi = (int)stop;
      const int nlo = (int)__dadd_rn(__dmul_rn(slope, start), icpt);
      double cl = stop;
      if (npts == 1) cl = start;
      const int nhi = (int)__dadd_rn(__dmul_rn(slope, cl), icpt);
      const int mlast = npts == 1 ? mlo : mhi;
      m.valid = 1;
      m.swap = swap ? 1 : 0;
      m.mlo = mlo;
      m.mhi = mlast;
      m.ex0 = swap ? nlo : mlo;
      m.ey0 = swap ? mlo : nlo;
      m.ex1 = swap ? nhi : mlast;
      m.ey1 = swap ? mlast : nhi;
      const int nmin = min(nlo, nhi), nmax = max(nlo, nhi);
      m.bx0 = (swap ? nmin : mlo) - 12;
      m.bx1 = (swap ? nmax : mlast) + 12;
      m.by0 = (swap ? mlo : nmin) - 12;
      m.by1 = (swap ? mlast : nmax) + 12;
      m.a = (float)slope;
      m.b = (float)icpt;
      fs->edge[e] = m;
    }
  }
}

__global__ void __launch_bounds__(256) raster_prep_kernel(const double* __restrict__ joints, GaussTable tab, double thres,
                                                          double foot_thres, FrameScratch* __restrict__ scratch, int H,
                                                          int W) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const int b = blockIdx.y;
  FrameScratch* fs = scratch + b;
  const double* jf = joints + (size_t)b * kJoints * 3;
  if (blockIdx.x < kJoints) heat_window(jf + blockIdx.x * 3, tab, thres, fs, blockIdx.x, H, W);
  else limb_tables(jf, thres, foot_thres, fs, H, W, blockIdx.x - kJoints);
}

// ---------------------------------------------------------------------------------------------
// Stamp enumeration for one pixel and one limb, in the reference's drawing order.
// ---------------------------------------------------------------------------------------------
// Source coordinates s with clamp(s + shift, 0, n - 1) == v (np.clip in drawEdge, keypoint2img.py:52-53).
__device__ __forceinline__ void src_range(int v, int shift, int n, int* lo, int* hi) {
  if (v > 0 && v < n - 1) {
    *lo = *hi = v - shift;
  } else if (v == 0) {
    *lo = -(1 << 20);
    *hi = -shift;
  } else {
    *lo = n - 1 - shift;
    *hi = 1 << 20;
  }
}

// Hands the 64-bit set of body stamps of the limb that touch (y, x) (bit = key = stamp order) to `body`, then the end-cap
// stamps, which come after all body stamps of the limb: an interior pixel `visit`s its (at most two) stamps in order; a
// clamped pixel hands their number and smallest key to `caps` and `flag`s every stamp but the pixel's first one.
// Body-stamp shifts s in [-4, 3] with clamp(c + s, 0, n - 1) == v (lo > hi: none).
__device__ __forceinline__ void shift_range(int v, int c, int n, int* lo, int* hi) {
  if (v > 0 && v < n - 1) {
    *lo = max(v - c, -4);
    *hi = min(v - c, 3);
  } else if (v == 0) {
    *lo = -4;
    *hi = min(3, -c);
  } else {
    *lo = max(-4, n - 1 - c);
    *hi = 3;
  }
}

template <typename Body, typename Caps, typename Flag, typename Visit>
__device__ __forceinline__ void visit_limb(const EdgeMeta& m, const short* __restrict__ f, int y, int x, int H, int W,
                                           Body&& body, Caps&& caps, Flag&& flag, Visit&& visit) {
  const bool interior = x > 0 && x < W - 1 && y > 0 && y < H - 1;
  unsigned long long mask = 0ull;
  if (interior) {
    const int pm = m.swap ? y : x, pn = m.swap ? x : y;
    const bool near_body = pm >= m.mlo - 4 && pm <= m.mhi + 3 && fabsf((float)pn - (m.a * (float)pm + m.b)) <= 11.0f;
    if (near_body) {
#pragma unroll
      for (int s = -4; s < 4; ++s) {  // shift along the major axis
        const int mm = pm - s;
        if (mm < m.mlo || mm > m.mhi) continue;
        const int nn = f[mm];
        if (nn < 0) continue;
        const int t = pn - nn;  // shift along the minor axis
        if ((unsigned)(t + 4) < 8u) {
          const int i = m.swap ? s : t, j = m.swap ? t : s;
          mask |= 1ull << ((i + 4) * 8 + (j + 4));
        }
      }
    }
  } else {
    // Clamped pixel (np.clip in drawEdge): a curve point (mm, f[mm]) reaches it through a RANGE of shifts along every
    // clamped axis, i.e. through a rectangle of stamps; the candidate points are the same eight as for an interior pixel.
    const int pm = m.swap ? y : x, pn = m.swap ? x : y;
    const int Dm = m.swap ? H : W, Dn = m.swap ? W : H;
    for (int s = -4; s < 4; ++s) {
      const int mm = pm - s;
      if (mm < m.mlo || mm > m.mhi || mm < 0 || mm >= Dm) continue;
      const int nn = f[mm];
      if (nn < 0) continue;
      int a0, a1, b0, b1;
      shift_range(pm, mm, Dm, &a0, &a1);
      shift_range(pn, nn, Dn, &b0, &b1);
      if (a0 > a1 || b0 > b1) continue;
      const int i0 = m.swap ? a0 : b0, i1 = m.swap ? a1 : b1, j0 = m.swap ? b0 : a0, j1 = m.swap ? b1 : a1;
      const unsigned long long rows = (~0ull >> (8 * (3 - i1))) & (~0ull << (8 * (i0 + 4)));
      const unsigned long long cols = (unsigned long long)(((1u << (j1 + 5)) - 1u) & ~((1u << (j0 + 4)) - 1u));
      mask |= rows & (cols * 0x0101010101010101ull);
    }
  }
  body(mask);
  // end caps (keypoint2img.py:59-64): stamps (i, j), i outer, both end points in one stamp
  if (interior) {
    int k0 = -1, k1 = -1;
    {
      const int i = y - m.ey0, j = x - m.ex0;
      if (i >= -12 && i < 12 && j >= -12 && j < 12 && i * i + j * j < 64) k0 = kBodyKeys + (i + 12) * 24 + (j + 12);
    }
    {
      const int i = y - m.ey1, j = x - m.ex1;
      if (i >= -12 && i < 12 && j >= -12 && j < 12 && i * i + j * j < 64) k1 = kBodyKeys + (i + 12) * 24 + (j + 12);
    }
    if (k0 >= 0 && k1 >= 0) {
      if (k0 == k1) {
        visit(k0);
      } else {
        visit(min(k0, k1));
        visit(max(k0, k1));
      }
    } else if (k0 >= 0) {
      visit(k0);
    } else if (k1 >= 0) {
      visit(k1);
    }
  } else {
    // clamped pixel: the stamps (i, j) of end point t that land on (y, x) form a rectangle (a range where the
    // coordinate is clamped, a single value otherwise); enumerate the hull of the two rectangles in stamp order
    int ilo[2], ihi[2], jlo[2], jhi[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int ey = t ? m.ey1 : m.ey0, ex = t ? m.ex1 : m.ex0;
      src_range(y, ey, H, &ilo[t], &ihi[t]);  // i with clamp(ey + i) == y
      src_range(x, ex, W, &jlo[t], &jhi[t]);
      ilo[t] = max(ilo[t], -12), ihi[t] = min(ihi[t], 11);
      jlo[t] = max(jlo[t], -12), jhi[t] = min(jhi[t], 11);
    }
    // Both passes walk rectangle 0, then the part of rectangle 1 outside of it (each stamp once; the order only matters
    // for the smallest key, which is the first one in stamp order).  Pass 0 counts and finds that key, `caps` applies
    // the count in closed form and says whether the first stamp is this pixel's very first one, pass 1 flags the others.
    int n = 0, kmin = 1 << 30;
    int skip = -1;
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < 2; ++t)
        for (int i = ilo[t]; i <= ihi[t]; ++i)
          for (int j = jlo[t]; j <= jhi[t]; ++j) {
            if (i * i + j * j >= 64) continue;
            if (t == 1 && i >= ilo[0] && i <= ihi[0] && j >= jlo[0] && j <= jhi[0]) continue;
            const int key = kBodyKeys + (i + 12) * 24 + (j + 12);
            if (pass == 0) {
              ++n;
              kmin = min(kmin, key);
            } else if (key != skip) {
              flag(key);
            }
          }
      if (pass == 0) {
        if (n == 0) break;
        skip = caps(n, kmin) ? kmin : -1;
      }
    }
  }
}

__device__ __forceinline__ uint32_t avg_color(uint32_t old, uint32_t col) {
  // per channel (v + c) >> 1 on packed 0x00BBGGRR; channels are <= 255 so (v + c) fits in 9 bits
  const uint32_t r = (((old & 0xffu) + (col & 0xffu)) >> 1);
  const uint32_t g = ((((old >> 8) & 0xffu) + ((col >> 8) & 0xffu)) >> 1);
  const uint32_t b = ((((old >> 16) & 0xffu) + ((col >> 16) & 0xffu)) >> 1);
  return r | (g << 8) | (b << 16);
}

// m more stamps of the same colour on a touched pixel: ((v + c) >> 1 applied m times) == (v + (2^m - 1) c) >> m per
// channel, because floor(floor(a / 2) + c) / 2) == floor((a + 2 c) / 4) for integers; m is split into steps of at most 16
// (255 * 65535 + 255 fits in 32 bits).
__device__ __forceinline__ uint32_t avg_color_n(uint32_t old, uint32_t col, int m) {
  while (m > 0) {
    const int t = m < 16 ? m : 16;
    const uint32_t k = (1u << t) - 1u;
    const uint32_t r = ((old & 0xffu) + k * (col & 0xffu)) >> t;
    const uint32_t g = (((old >> 8) & 0xffu) + k * ((col >> 8) & 0xffu)) >> t;
    const uint32_t b = (((old >> 16) & 0xffu) + k * ((col >> 16) & 0xffu)) >> t;
    old = r | (g << 8) | (b << 16);
    m -= t;
  }
  return old;
}

// Tile of 8 rows x 128 columns per CTA (256 threads x 4 consecutive pixels).
static constexpr int kTileRows = 8, kTileCols = 128, kPxPerThread = 4;

struct TileEdges {
  int n;
  int idx[kEdges];
};

__device__ __forceinline__ void tile_edges(const FrameScratch* fs, int y0, int x0, int H, int W, TileEdges* te,
                                           EdgeMeta* s_meta) {
  if (threadIdx.x < 32) {  // one lane per limb, compacted with a ballot (limb order is preserved)
    const int e = threadIdx.x;
    bool keep = false;
    EdgeMeta m;
    if (e < kEdges) {
      m = fs->edge[e];
      if (m.valid) {
        // everything a limb touches lies in its bounding box clipped to the image
        const int bx0 = max(m.bx0, 0), bx1 = min(m.bx1, W - 1), by0 = max(m.by0, 0), by1 = min(m.by1, H - 1);
        keep = !(bx1 < x0 || bx0 >= x0 + kTileCols || by1 < y0 || by0 >= y0 + kTileRows);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = __popc(bal & ((1u << e) - 1u));
      s_meta[pos] = m;
      te->idx[pos] = e;
    }
    if (e == 0) te->n = __popc(bal);
  }
  __syncthreads();
}

// Conservative x-interval [lo, hi] of row y outside of which limb m touches no pixel (lo > hi: nothing on this row).
// Body: a stamp (i, j) in [-4, 3]^2 of curve point (mm, f[mm]) with f[mm] within ~2 of a*mm + b and |a| <= 1, i.e. the
// same band |minor - (a*major + b)| <= 11 that visit_limb tests, widened to 12; end caps: |i|, |j| <= 12 around the two
// end points.  Rows whose y is clamped (first / last row collect everything drawn beyond them) fall back to the
// bounding box; sources beyond the left / right border are clamped onto the border pixel by clamping the interval.
__device__ __forceinline__ void row_interval(const EdgeMeta& m, int y, int H, int W, int* lo, int* hi) {
  int l = 1 << 20, h = -(1 << 20);
  if (y < m.by0 || y > m.by1) {
    // nothing
  } else if (y <= 0 || y >= H - 1) {
    l = m.bx0;
    h = m.bx1;
  } else {
    if (abs(y - m.ey0) <= 12) l = min(l, m.ex0 - 12), h = max(h, m.ex0 + 12);
    if (abs(y - m.ey1) <= 12) l = min(l, m.ex1 - 12), h = max(h, m.ex1 + 12);
    const float fy = (float)y;
    if (m.swap) {  // major axis y: x ~ a*y + b
      if (y >= m.mlo - 4 && y <= m.mhi + 3) {
        const float c = m.a * fy + m.b;
        l = min(l, (int)floorf(c - 12.f));
        h = max(h, (int)ceilf(c + 12.f));
      }
    } else {       // major axis x: y ~ a*x + b
      int bl = m.mlo - 4, bh = m.mhi + 3;
      if (fabsf(m.a) > 1e-3f) {
        const float x0 = (fy - m.b - 12.f) / m.a, x1 = (fy - m.b + 12.f) / m.a;
        bl = max(bl, (int)floorf(fminf(x0, x1)) - 1);
        bh = min(bh, (int)ceilf(fmaxf(x0, x1)) + 1);
      } else if (fabsf(fy - (m.a * 0.5f * (float)(m.mlo + m.mhi) + m.b)) > 14.f) {
        bh = bl - 1;   // (the line moves by at most ~1 px over its whole length)
      }
      if (bl <= bh) l = min(l, bl), h = max(h, bh);
    }
  }
  if (l > h) {
    *lo = 1;
    *hi = 0;
  } else {
    *lo = min(max(l, 0), W - 1);
    *hi = min(max(h, 0), W - 1);
  }
}

// mark: every stamp that is not the first one of some pixel hit an older pixel.  Per pixel it leaves
//   bits 0-23 colour if the first stamp painted `col`, 24-47 colour if it painted col >> 1, 48-52 limb and 53-62 key
//   of the first stamp, bit 63 = touched at all
// in `cache` (only tiles that some limb can touch are written; paint skips the others the same way).
__global__ void __launch_bounds__(256) raster_mark_kernel(FrameScratch* __restrict__ scratch,
                                                          unsigned long long* __restrict__ cache, int H, int W) {
  __shared__ TileEdges te;
  __shared__ EdgeMeta s_meta[kEdges];
  pdl_wait();
  FrameScratch* fs = scratch + blockIdx.z;
  const int y0 = blockIdx.y * kTileRows, x0 = blockIdx.x * kTileCols;
  tile_edges(fs, y0, x0, H, W, &te, s_meta);
  if (te.n == 0) return;
  const int y = y0 + (threadIdx.x >> 5);
  if (y >= H) return;
  // One lane per limb works out the x-interval of this row outside of which the limb touches nothing (the whole warp
  // is on row y).  The warp then walks its 128 columns in four 32-pixel segments, one pixel per lane: a limb is only
  // enumerated in the segments its interval reaches, in drawing order (two shuffles per limb and segment).
  const int lane = threadIdx.x & 31;
  int my_lo = 1, my_hi = 0;
  if (lane < te.n) row_interval(s_meta[lane], y, H, W, &my_lo, &my_hi);
  unsigned long long* crow = cache + ((size_t)blockIdx.z * H + y) * W;
  for (int seg = 0; seg < kTileCols / 32; ++seg) {
    const int xs = x0 + seg * 32;
    if (xs >= W) break;
    const int x = xs + lane;
    int count = 0;
    uint32_t ca = 0u, cb = 0u, first = 0u;
    for (int k = 0; k < te.n; ++k) {
      const int lo = __shfl_sync(0xffffffffu, my_lo, k), hi = __shfl_sync(0xffffffffu, my_hi, k);
      if (xs + 31 < lo || xs > hi) continue;   // uniform
      const int e = te.idx[k];
      unsigned long long older = 0ull;         // body stamps of limb e that are not this pixel's first stamp
      if (x >= lo && x <= hi && x < W) {
        const EdgeMeta& m = s_meta[k];
        const uint32_t col = (uint32_t)c_colors[e][0] | ((uint32_t)c_colors[e][1] << 8) | ((uint32_t)c_colors[e][2] << 16);
        auto visit = [&](int key) {            // one end-cap stamp
          if (count++ == 0) {
            first = (uint32_t)e | ((uint32_t)key << 5);
            ca = col;
            cb = avg_color(0u, col);
          } else {
            fs->flag[e][key] = 1;
            ca = avg_color(ca, col);
            cb = avg_color(cb, col);
          }
        };
        // The body stamps come first, in key order; all of them carry the limb's colour, so the averaging chain of the
        // nb - (first ? 1 : 0) stamps that land on an already touched pixel has a closed form.
        auto apply_body = [&](unsigned long long mask) {
          const int nb = __popcll(mask);
          if (nb == 0) return;
          int more = nb;
          older = mask;
          if (count == 0) {
            const int bit = __ffsll((long long)mask) - 1;
            first = (uint32_t)e | ((uint32_t)bit << 5);
            ca = col;
            cb = avg_color(0u, col);
            older = mask & (mask - 1ull);
            more = nb - 1;
          }
          ca = avg_color_n(ca, col, more);
          cb = avg_color_n(cb, col, more);
          count += nb;
        };
        auto apply_caps = [&](int n, int kmin) {   // n end-cap stamps on a clamped pixel; true: kmin is the pixel's first stamp
          int more = n;
          const bool is_first = count == 0;
          if (is_first) {
            first = (uint32_t)e | ((uint32_t)kmin << 5);
            ca = col;
            cb = avg_color(0u, col);
            more = n - 1;
          }
          ca = avg_color_n(ca, col, more);
          cb = avg_color_n(cb, col, more);
          count += n;
          return is_first;
        };
        auto flag_cap = [&](int key) { fs->flag[e][key] = 1; };
        visit_limb(m, fs->f[e], y, x, H, W, apply_body, apply_caps, flag_cap, visit);
      }
      // one atomic per limb and segment: the union of the lanes' sets (most are known already)
      const uint32_t olo = __reduce_or_sync(0xffffffffu, (uint32_t)older);
      const uint32_t ohi = __reduce_or_sync(0xffffffffu, (uint32_t)(older >> 32));
      if (lane == 0) {
        const unsigned long long u = (unsigned long long)olo | ((unsigned long long)ohi << 32);
        if (u & ~fs->bodyflag[e]) atomicOr(&fs->bodyflag[e], u);
      }
    }
    if (x < W)
      crow[x] = count > 0 ? (unsigned long long)ca | ((unsigned long long)cb << 24) | ((unsigned long long)first << 48) | (1ull << 63)
                          : 0ull;
  }
}

// paint: all 22 channels of every pixel, as fp32 NCHW and / or the generator's 16-bit planar input.
__global__ void __launch_bounds__(256, 6) raster_paint_kernel(const FrameScratch* __restrict__ scratch,
                                                           const unsigned long long* __restrict__ cache,
                                                           float* __restrict__ label, act_t* __restrict__ planar,
                                                           int H, int W) {
  __shared__ TileEdges te;
  __shared__ EdgeMeta s_meta[kEdges];
  __shared__ JointMeta s_joint[kJoints];
  __shared__ float s_lut[256];
  __shared__ unsigned s_jmask;
  pdl_wait();
  const int b = blockIdx.z;
  const FrameScratch* fs = scratch + b;
  const int y0 = blockIdx.y * kTileRows, x0 = blockIdx.x * kTileCols;
  // ToTensor: float32(v) / 255, Normalize: (t - 0.5) / 0.5, both in fp32
  s_lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.0f), 0.5f), 0.5f);
  if (threadIdx.x < 32) {   // joints whose 41x41 window reaches this tile (bit j of s_jmask); the others are never tested
    bool in = false;
    if (threadIdx.x < kJoints) {
      const JointMeta jm = fs->joint[threadIdx.x];
      s_joint[threadIdx.x] = jm;
      in = jm.valid && jm.ix + kRadius >= x0 && jm.ix - kRadius < x0 + kTileCols && jm.iy + kRadius >= y0 &&
           jm.iy - kRadius < y0 + kTileRows;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (threadIdx.x == 0) s_jmask = bal;
  }
  tile_edges(const_cast<FrameScratch*>(fs), y0, x0, H, W, &te, s_meta);
  // Pixel p of a thread is column x0 + 32 p + lane: consecutive lanes own consecutive pixels, so every store instruction
  // of a warp writes one contiguous run (512 bytes of a planar plane, 128 bytes of an fp32 plane) in whole sectors.
  const int y = y0 + (threadIdx.x >> 5);
  const int xl = x0 + (threadIdx.x & 31);
  if (y >= H || xl >= W) return;
  const unsigned jmask = s_jmask;   // (written before the barrier inside tile_edges)
  uint32_t rgb[kPxPerThread];
#pragma unroll
  for (int p = 0; p < kPxPerThread; ++p) rgb[p] = 0u;
  if (te.n > 0) {  // same test as in mark: only then was the cache written
    const unsigned long long* cp = cache + ((size_t)b * H + y) * W + xl;
#pragma unroll
    for (int p = 0; p < kPxPerThread; ++p) {
      const unsigned long long word = xl + 32 * p < W ? cp[32 * p] : 0ull;
      if (!(word >> 63)) continue;
      const uint32_t first = (uint32_t)(word >> 48) & 0x7fffu;
      const uint32_t fe = first & 31u, fkey = first >> 5;
      const bool hit_older = fkey < (uint32_t)kBodyKeys ? ((fs->bodyflag[fe] >> fkey) & 1ull) != 0ull : fs->flag[fe][fkey] != 0;
      rgb[p] = (uint32_t)(hit_older ? word >> 24 : word) & 0xffffffu;
    }
  }
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)y * W + xl;
  float* out = label != nullptr ? label + (size_t)b * 22 * HW + pix : nullptr;
  act_t* pout = planar != nullptr ? planar + (size_t)b * 32 * HW + pix * 8 : nullptr;
  // channel c of the label: 0..2 skeleton, 3..21 heat-maps, 22..31 zero padding of the planar copy;
  // produced one 8-channel plane at a time to keep the register footprint small
#pragma unroll
  for (int pl = 0; pl < 4; ++pl) {
    float v[8][kPxPerThread];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = pl * 8 + k;
#pragma unroll
      for (int p = 0; p < kPxPerThread; ++p) {
        float h = 0.f;
        if (c < 3) {
          h = s_lut[(rgb[p] >> (8 * c)) & 0xffu];
        } else if (c < 22 && ((jmask >> (c - 3)) & 1u)) {   // (uniform over the block)
          const JointMeta jm = s_joint[c - 3];
          const int wy = y - jm.iy + kRadius, wx = xl + 32 * p - jm.ix + kRadius;
          if ((unsigned)wy < (unsigned)kTaps && (unsigned)wx < (unsigned)kTaps) h = fs->win[c - 3][wy * kTaps + wx];
        }
        v[k][p] = h;
      }
      if (out != nullptr && c < 22) {
#pragma unroll
        for (int p = 0; p < kPxPerThread; ++p)
          if (xl + 32 * p < W) out[(size_t)c * HW + 32 * p] = v[k][p];
      }
    }
    if (pout != nullptr) {  // [B][4][H][W][8]
#pragma unroll
      for (int p = 0; p < kPxPerThread; ++p) {
        if (xl + 32 * p >= W) break;
        *reinterpret_cast<uint4*>(pout + (size_t)pl * HW * 8 + (size_t)(32 * p) * 8) =
            make_uint4(pack2(v[0][p], v[1][p]), pack2(v[2][p], v[3][p]), pack2(v[4][p], v[5][p]), pack2(v[6][p], v[7][p]));
      }
    }
  }
}

int launch_rasterize(const double* joints_dev, int B, int H, int W, const double* wtab41_host, double skeleton_thres,
                     double foot_thres, float* label, act_t* label_planar, void* workspace, long long workspace_bytes,
                     cudaStream_t stream) {
  RIB_REQUIRE(H >= 2 && W >= 2 && B >= 1, "rasterize: bad shape");
  RIB_REQUIRE(H <= kMaxDim && W <= kMaxDim, "rasterize: images larger than 1024 px are not supported");
  RIB_REQUIRE(label != nullptr || label_planar != nullptr, "rasterize: no output requested");
  RIB_REQUIRE(workspace != nullptr && workspace_bytes >= raster_workspace_bytes(B, H, W), "rasterize: workspace too small");
  RIB_REQUIRE(((uintptr_t)workspace & 15) == 0, "rasterize: workspace must be 16-byte aligned");
  GaussTable tab;
  for (int i = 0; i < kTaps; ++i) tab.w[i] = wtab41_host[i];
  FrameScratch* fs = static_cast<FrameScratch*>(workspace);
  unsigned long long* cache = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + scratch_bytes(B));
  launch_pdl(raster_prep_kernel, dim3(kJoints + kEdges, B), dim3(256), 0, stream, joints_dev, tab, skeleton_thres, foot_thres,
             fs, H, W);
  RIB_CHECK_CUDA(cudaGetLastError());
  const dim3 grid(ceil_div(W, kTileCols), ceil_div(H, kTileRows), B);
  launch_pdl(raster_mark_kernel, grid, dim3(256), 0, stream, fs, cache, H, W);
  RIB_CHECK_CUDA(cudaGetLastError());
  launch_pdl(raster_paint_kernel, grid, dim3(256), 0, stream, static_cast<const FrameScratch*>(fs),
             static_cast<const unsigned long long*>(cache), label, label_planar, H, W);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rib
