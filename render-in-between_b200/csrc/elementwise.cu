// Bandwidth-bound kernels of the rendering path (sm_100a).  All are single-pass, 16-byte vectorised
// and sized so that consecutive threads touch consecutive addresses.
#include "elementwise.cuh"

namespace rib {

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> chunk-planar 16-bit.  One thread builds one 16-byte vector (8 channels of one pixel):
// reads are coalesced along x in every source plane, the store is one coalesced 16-byte piece.
// Up to three NCHW sources are concatenated along channels (torch.cat never materialises); channels
// of the destination planes outside [c_off, c_off + sum C) are written as zero.
// ---------------------------------------------------------------------------------------------
__global__ void pack_nchw_kernel(const PackSrc s0, const PackSrc s1, const PackSrc s2, act_t* __restrict__ dst,
                                 long long dst_bs, int c_off, int plane0, int nplanes, int HW, size_t total) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B * nplanes * HW
  if (i >= total) return;
  const int hw = (int)(i % HW);
  const size_t r = i / HW;
  const int pl = plane0 + (int)(r % nplanes);
  const size_t n = r / nplanes;
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int c = pl * 8 + 2 * k + h - c_off;
      float x = 0.f;
      if (c >= 0) {
        if (c < s0.C) x = __ldg(s0.p + (n * s0.C + c) * (size_t)HW + hw);
        else if ((c -= s0.C) < s1.C) x = __ldg(s1.p + (n * s1.C + c) * (size_t)HW + hw);
        else if ((c -= s1.C) < s2.C) x = __ldg(s2.p + (n * s2.C + c) * (size_t)HW + hw);
      }
      v[h] = x;
    }
    o[k] = pack2(v[0], v[1]);
  }
  *reinterpret_cast<uint4*>(dst + n * dst_bs + ((size_t)pl * HW + hw) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

int launch_pack_nchw(const PackSrc* srcs, int nsrc, act_t* dst, long long dst_bstride, int c_off, int plane0,
                     int nplanes, int B, int H, int W, cudaStream_t s) {
  RIB_REQUIRE(nsrc >= 1 && nsrc <= 3, "pack: 1..3 sources");
  PackSrc z = {nullptr, 0};
  const PackSrc s0 = srcs[0], s1 = nsrc > 1 ? srcs[1] : z, s2 = nsrc > 2 ? srcs[2] : z;
  const size_t total = (size_t)B * nplanes * H * W;
  const int threads = 256;
  launch_pdl(pack_nchw_kernel, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, s, s0, s1, s2, dst, dst_bstride, c_off,
                                                                                   plane0, nplanes, H * W, total);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void pack_images_kernel(const float* __restrict__ fake, const float* __restrict__ prev,
                                   act_t* __restrict__ emb, long long emb_bs, act_t* __restrict__ msk, long long msk_bs,
                                   int HW, size_t total) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B * HW / 4
  if (i >= total) return;
  const int q4 = HW >> 2;
  const size_t n = i / q4;
  const int pix = (int)(i - n * q4) * 4;
  float f[3][4], p[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(fake + (n * 3 + c) * (size_t)HW + pix));
    const float4 b = __ldg(reinterpret_cast<const float4*>(prev + (n * 3 + c) * (size_t)HW + pix));
    f[c][0] = a.x, f[c][1] = a.y, f[c][2] = a.z, f[c][3] = a.w;
    p[c][0] = b.x, p[c][1] = b.y, p[c][2] = b.z, p[c][3] = b.w;
  }
  uint4* eo = reinterpret_cast<uint4*>(emb + n * emb_bs + (size_t)pix * 8);
  uint4* mo = reinterpret_cast<uint4*>(msk + n * msk_bs + (size_t)pix * 8);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    eo[k] = make_uint4(pack2(f[0][k], f[1][k]), pack2(f[2][k], p[0][k]), pack2(p[1][k], p[2][k]), 0u);
    mo[k] = make_uint4(pack2(p[0][k], p[1][k]), pack2(p[2][k], f[0][k]), pack2(f[1][k], f[2][k]), 0u);
  }
}

int launch_pack_images(const float* fake, const float* prev, act_t* emb, long long emb_bstride, act_t* mask,
                       long long mask_bstride, int B, int H, int W, cudaStream_t s) {
  RIB_REQUIRE((H * W) % 4 == 0, "pack_images: H*W must be a multiple of 4");
  RIB_REQUIRE((((uintptr_t)fake | (uintptr_t)prev) & 15) == 0, "pack_images: inputs must be 16-byte aligned");
  const size_t total = (size_t)B * (H * W / 4);
  launch_pdl(pack_images_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, fake, prev, emb, emb_bstride, mask, mask_bstride,
                                                                     H * W, total);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Instance-norm application
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void in_coeffs(const double* stats, const float* w, const float* b, int n, int C, int c,
                                          double cnt, float eps, float* scale, float* shift) {
  const double s = stat_sum(stats + ((size_t)n * C + c) * 2), ss = stat_sumsq(stats + ((size_t)n * C + c) * 2);
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = w ? (double)w[c] : 1.0, be = b ? (double)b[c] : 0.0;
  *scale = (float)(g * rstd);
  *shift = (float)(be - mean * g * rstd);
}

__global__ void in_apply_kernel(const InApplyParams p) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  extern __shared__ float s_coef[];  // [4][C]: scale_a, shift_a, scale_b, shift_b
  const int n = blockIdx.y;
  const double cnt = (double)p.H * (double)p.W;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    in_coeffs(p.astats, p.aw, p.ab, n, p.C, c, cnt, p.eps, &s_coef[c], &s_coef[p.C + c]);
    if (p.b != nullptr && p.bstats != nullptr)
      in_coeffs(p.bstats, p.bw, p.bb, n, p.C, c, cnt, p.eps, &s_coef[2 * p.C + c], &s_coef[3 * p.C + c]);
  }
  __syncthreads();
  const size_t HW = (size_t)p.H * p.W;
  const size_t nvec = HW * (p.C / 8);
  const size_t step = (size_t)gridDim.x * blockDim.x;
  constexpr int U = 4;  // independent 16-byte vectors in flight per thread
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += U * step) {
    uint4 av[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * step;
      if (i < nvec) {
        av[u] = *reinterpret_cast<const uint4*>(p.a + (size_t)n * p.a_bs + i * 8);
        if (p.b != nullptr) bv[u] = *reinterpret_cast<const uint4*>(p.b + (size_t)n * p.b_bs + i * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * step;
      if (i >= nvec) break;
      const size_t pl = i / HW, pix = i - pl * HW;
      const int c0 = (int)pl * 8;
      const uint32_t au[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
      float v[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack2(au[k], v[2 * k], v[2 * k + 1]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = v[k] * s_coef[c0 + k] + s_coef[p.C + c0 + k];
        if (p.act) v[k] = lrelu02(v[k]);
      }
      if (p.b != nullptr) {
        const uint32_t bu[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
        float w[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack2(bu[k], w[2 * k], w[2 * k + 1]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[k] += p.bstats ? (w[k] * s_coef[2 * p.C + c0 + k] + s_coef[3 * p.C + c0 + k]) : w[k];
      }
      const uint4 o = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
      if (p.out_parity) {
        const int y = (int)(pix / p.W), x = (int)(pix - (size_t)y * p.W);
        const size_t q = (size_t)((y & 1) * 2 + (x & 1)) * (HW >> 2) + (size_t)(y >> 1) * (p.W >> 1) + (x >> 1);
        *reinterpret_cast<uint4*>(p.out + (size_t)n * p.o_bs + (pl * HW + q) * 8) = o;
      } else if (!p.ups) {
        *reinterpret_cast<uint4*>(p.out + (size_t)n * p.o_bs + i * 8) = o;
      } else {
        const int y = (int)(pix / p.W), x = (int)(pix - (size_t)y * p.W);
        const int W2 = 2 * p.W;
        act_t* ob = p.out + (size_t)n * p.o_bs + (pl * 4 * HW + (size_t)(2 * y) * W2 + 2 * x) * 8;
        *reinterpret_cast<uint4*>(ob) = o;
        *reinterpret_cast<uint4*>(ob + 8) = o;
        *reinterpret_cast<uint4*>(ob + (size_t)W2 * 8) = o;
        *reinterpret_cast<uint4*>(ob + (size_t)W2 * 8 + 8) = o;
      }
    }
  }
}

// Same operation for the parity-planar output of a sub-pixel conv: one thread reads the two x parities of an output
// pixel pair (two coalesced streams) and writes 32 contiguous bytes of the normal map.
__global__ void in_apply_unparity_kernel(const InApplyParams p) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  extern __shared__ float s_coef[];  // [2][C]: scale, shift
  const int n = blockIdx.y;
  const double cnt = (double)p.H * (double)p.W;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x)
    in_coeffs(p.astats, p.aw, p.ab, n, p.C, c, cnt, p.eps, &s_coef[c], &s_coef[p.C + c]);
  __syncthreads();
  const size_t HW = (size_t)p.H * p.W, Q = HW >> 2;
  const int Wh = p.W >> 1;
  const size_t npair = (size_t)(p.C / 8) * p.H * Wh;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  constexpr int U = 2;
  const act_t* ab = p.a + (size_t)n * p.a_bs;
  act_t* ob = p.out + (size_t)n * p.o_bs;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < npair; i0 += U * step) {
    uint4 v0[U], v1[U];
    size_t dst[U];
    int c0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * step;
      if (i < npair) {
        const int x2 = (int)(i % Wh);
        const size_t r = i / Wh;
        const int y = (int)(r % p.H);
        const size_t pl = r / p.H;
        const size_t src = pl * HW + (size_t)((y & 1) * 2) * Q + (size_t)(y >> 1) * Wh + x2;
        v0[u] = *reinterpret_cast<const uint4*>(ab + src * 8);
        v1[u] = *reinterpret_cast<const uint4*>(ab + (src + Q) * 8);
        dst[u] = (pl * HW + (size_t)y * p.W + 2 * x2) * 8;
        c0[u] = (int)pl * 8;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * step >= npair) break;
      const uint32_t au[8] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w, v1[u].x, v1[u].y, v1[u].z, v1[u].w};
      uint32_t o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a, b;
        unpack2(au[k], a, b);
        const int c = c0[u] + 2 * (k & 3);
        a = a * s_coef[c] + s_coef[p.C + c];
        b = b * s_coef[c + 1] + s_coef[p.C + c + 1];
        if (p.act) {
          a = lrelu02(a);
          b = lrelu02(b);
        }
        o[k] = pack2(a, b);
      }
      *reinterpret_cast<uint4*>(ob + dst[u]) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(ob + dst[u] + 8) = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
}

// SM count of the current device (grid sizing of the bandwidth kernels); 148 only if the query fails.
static unsigned sm_count_current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148u;
  const int n = device_sm_count(dev);
  return n > 0 ? (unsigned)n : 148u;
}

int launch_in_apply(const InApplyParams& p, cudaStream_t s) {
  RIB_REQUIRE(p.C % 8 == 0, "in_apply: channels must be a multiple of 8");
  if (p.in_parity) {
    RIB_REQUIRE(p.b == nullptr && !p.ups && !p.out_parity && p.H % 2 == 0 && p.W % 2 == 0,
                "in_apply: the parity-planar input form is single-term, same-size, normal output");
    const size_t npair = (size_t)p.H * (p.W / 2) * (p.C / 8);
    unsigned gx = (unsigned)((npair + 255) / 256);
    static const unsigned bps = getenv("RIB_INAPPLY_BPS") ? (unsigned)atoi(getenv("RIB_INAPPLY_BPS")) : 8u;
    unsigned want = (sm_count_current() * (bps ? bps : 8u) + (unsigned)p.B - 1u) / (unsigned)p.B;
    if (gx > want) gx = want;
    launch_pdl(in_apply_unparity_kernel, dim3(gx, (unsigned)p.B), dim3(256), 2 * p.C * sizeof(float), s, p);
    RIB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t nvec = (size_t)p.H * p.W * (p.C / 8);
  const int threads = 256;
  // every block first derives the normalisation coefficients of its image (fp64 divide + sqrt per channel),
  // so blocks are kept few and fat: about 8 blocks per SM over the whole batch (measured: a single resident wave of
  // 3-4 blocks per SM is slower, 348 -> 432 us per forward; so is requesting the first vectors before the prologue)
  unsigned gx = (unsigned)((nvec + threads - 1) / threads);
  static const unsigned bps = getenv("RIB_INAPPLY_BPS") ? (unsigned)atoi(getenv("RIB_INAPPLY_BPS")) : 8u;   // blocks per SM
  unsigned want = (sm_count_current() * (bps ? bps : 8u) + (unsigned)p.B - 1u) / (unsigned)p.B;
  if (want < 1u) want = 1u;
  if (gx > want) gx = want;
  dim3 grid(gx, (unsigned)p.B);
  launch_pdl(in_apply_kernel, grid, dim3(threads), 4 * p.C * sizeof(float), s, p);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// AvgPool 3x3 s2 p1 + statistics.  Grid = (row bands, channel plane, image).  The 2 RB + 1 input rows behind a band of RB
// output rows are contiguous in a plane ([H][W][8]), so one bulk copy (cp.async.bulk, completion on an mbarrier) stages
// them in shared memory: no register staging and no L1 pressure for the 2.25x overlap of the 3x3 windows, and the copies
// of the other resident blocks run under this block's arithmetic.  A thread then produces kPoolR consecutive output rows of
// one column; the horizontal sum of an odd input row is shared by the two output rows that touch it.  Every thread of a
// block works on the same 8 channels, so the statistics reduce with warp shuffles.
// ---------------------------------------------------------------------------------------------
static constexpr int kPoolR = 4;                       // output rows per thread
static constexpr int kPoolSmemTarget = 72 * 1024;      // bytes of staged rows per block (three blocks per SM)

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float div9(float a) {   // a / 9 (reciprocal, one Newton step on the remainder)
  const float r9 = 1.0f / 9.0f;
  const float q = a * r9;
  return fmaf(fmaf(-q, 9.0f, a), r9, q);
}

__global__ void __launch_bounds__(256) avgpool3s2_kernel(const act_t* __restrict__ src, long long src_bs,
                                                         act_t* __restrict__ dst, long long dst_bs,
                                                         double* __restrict__ stats, int H, int W, int C, int RB) {
  extern __shared__ __align__(128) uint8_t pool_smem[];   // [2 RB + 1][W] 16-byte vectors; row 0 = input row 2 oy0 - 1
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float s_red[8][16];  // one slot per warp: fixed-order (deterministic) block reduction
  const int pl = blockIdx.y, n = blockIdx.z;
  const int Ho = H / 2, Wo = W / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oy0 = blockIdx.x * RB;
  const int rb = min(RB, Ho - oy0);                       // output rows of this band
  const int iy_lo = 2 * oy0 - 1, row_lo = max(iy_lo, 0), row_hi = 2 * (oy0 + rb);   // input rows [row_lo, row_hi)
  uint4* tile = reinterpret_cast<uint4*>(pool_smem);
  const uint4* sp = reinterpret_cast<const uint4*>(src + (size_t)n * src_bs + (size_t)pl * H * W * 8);
  uint4* dp = reinterpret_cast<uint4*>(dst + (size_t)n * dst_bs + (size_t)pl * Ho * Wo * 8);
  const uint32_t bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (iy_lo < 0)   // the padding row above the image
    for (int i = threadIdx.x; i < W; i += blockDim.x) tile[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(row_hi - row_lo) * (uint32_t)W * 16u;
    mbar_arrive_expect_tx(bar, bytes);
    bulk_load_1d(smem_u32(tile + (size_t)(row_lo - iy_lo) * W), sp + (size_t)row_lo * W, bytes, bar);
  }
  mbar_wait(bar, 0);
  float t1[8], t2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) t1[k] = t2[k] = 0.f;
  const int nstrip = ((rb + kPoolR - 1) / kPoolR) * Wo;
  for (int s = threadIdx.x; s < nstrip; s += blockDim.x) {
    const int sub = s / Wo, ox = s - sub * Wo;
    const int o0 = sub * kPoolR;                         // first output row of the strip, relative to the band
    const uint4* col = tile + (size_t)(2 * o0) * W + 2 * ox;   // local row 2 o0 = input row 2 (oy0 + o0) - 1
    auto hsum = [&](int r, float* h) {   // input columns 2 ox - 1, 2 ox, 2 ox + 1 of local row 2 o0 + r, left to right
      const uint4* rp = col + (size_t)r * W;
      const uint4 vl = ox > 0 ? rp[-1] : make_uint4(0u, 0u, 0u, 0u), vc = rp[0], vr = rp[1];
      const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w}, cw[4] = {vc.x, vc.y, vc.z, vc.w}, rw[4] = {vr.x, vr.y, vr.z, vr.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float la, lb, ca, cb, ra, rb2;
        unpack2(lw[k], la, lb);
        unpack2(cw[k], ca, cb);
        unpack2(rw[k], ra, rb2);
        h[2 * k] = (la + ca) + ra;
        h[2 * k + 1] = (lb + cb) + rb2;
      }
    };
    float h0[8], h1[8], h2[8];
    hsum(0, h0);
#pragma unroll
    for (int o = 0; o < kPoolR; ++o) {
      if (o0 + o >= rb) break;
      hsum(2 * o + 1, h1);
      hsum(2 * o + 2, h2);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k] = div9((h0[k] + h1[k]) + h2[k]);
        h0[k] = h2[k];
        t1[k] += acc[k];
        t2[k] = fmaf(acc[k], acc[k], t2[k]);
      }
      dp[(size_t)(oy0 + o0 + o) * Wo + ox] =
          make_uint4(pack2(acc[0], acc[1]), pack2(acc[2], acc[3]), pack2(acc[4], acc[5]), pack2(acc[6], acc[7]));
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        t1[k] += __shfl_xor_sync(0xffffffffu, t1[k], off);
        t2[k] += __shfl_xor_sync(0xffffffffu, t2[k], off);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s_red[warp][k] = t1[k];
        s_red[warp][8 + k] = t2[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
      const int c = pl * 8 + (threadIdx.x & 7);
      stat_add(&stats[((size_t)n * C + c) * 2], threadIdx.x >> 3, t);
    }
  }
}

#ifdef RIB_POOL_V0
// The round-1 form (one output pixel per thread, nine 16-byte loads each), kept as an A/B build (tools/): 346 us per
// forward against 221 us for the band kernel.
__global__ void avgpool3s2_v0_kernel(const act_t* __restrict__ src, long long src_bs, act_t* __restrict__ dst,
                                  long long dst_bs, double* __restrict__ stats, int H, int W, int C) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  __shared__ float s_red[8][16];  // one slot per warp: fixed-order (deterministic) block reduction
  const int pl = blockIdx.y, n = blockIdx.z;
  const int Ho = H / 2, Wo = W / 2;
  const act_t* sp = src + (size_t)n * src_bs + (size_t)pl * H * W * 8;
  act_t* dp = dst + (size_t)n * dst_bs + (size_t)pl * Ho * Wo * 8;
  float t1[8], t2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) t1[k] = t2[k] = 0.f;
  const int npix = Ho * Wo;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const int oy = pix / Wo, ox = pix - oy * Wo;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = 2 * oy + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = 2 * ox + dx;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(sp + ((size_t)iy * W + ix) * 8);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a, b;
          unpack2(u[k], a, b);
          acc[2 * k] += a;
          acc[2 * k + 1] += b;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] = acc[k] / 9.0f;
      t1[k] += acc[k];
      t2[k] += acc[k] * acc[k];
    }
    *reinterpret_cast<uint4*>(dp + (size_t)pix * 8) =
        make_uint4(pack2(acc[0], acc[1]), pack2(acc[2], acc[3]), pack2(acc[4], acc[5]), pack2(acc[6], acc[7]));
  }
  if (stats != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        t1[k] += __shfl_xor_sync(0xffffffffu, t1[k], off);
        t2[k] += __shfl_xor_sync(0xffffffffu, t2[k], off);
      }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s_red[threadIdx.x >> 5][k] = t1[k];
        s_red[threadIdx.x >> 5][8 + k] = t2[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
      const int c = pl * 8 + (threadIdx.x & 7);
      stat_add(&stats[((size_t)n * C + c) * 2], threadIdx.x >> 3, t);
    }
  }
}

static int launch_avgpool3s2_v0(const act_t* src, long long src_bs, act_t* dst, long long dst_bs, double* stats, int B, int H,
                      int W, int C, cudaStream_t s) {
  RIB_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool: bad shape");
  const int npix = (H / 2) * (W / 2);
  const int threads = 256;
  unsigned gx = (unsigned)((npix + threads * 4 - 1) / (threads * 4));
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)(C / 8), (unsigned)B);
  launch_pdl(avgpool3s2_v0_kernel, grid, dim3(threads), 0, s, src, src_bs, dst, dst_bs, stats, H, W, C);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif

int launch_avgpool3s2(const act_t* src, long long src_bs, act_t* dst, long long dst_bs, double* stats, int B, int H,
                      int W, int C, cudaStream_t s) {
#ifdef RIB_POOL_V0
  return launch_avgpool3s2_v0(src, src_bs, dst, dst_bs, stats, B, H, W, C, s);
#endif
  RIB_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool: bad shape");
  RIB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && src_bs % 8 == 0 && dst_bs % 8 == 0,
              "avgpool: maps must be 16-byte aligned");
  const int Ho = H / 2;
  const size_t row_bytes = (size_t)W * 16;
  // output rows per band: a multiple of kPoolR whose 2 RB + 1 input rows stay near the shared-memory target
  int RB = (int)((kPoolSmemTarget / row_bytes - 1) / 2) / kPoolR * kPoolR;
  if (RB < kPoolR) RB = kPoolR;
  const int ho_up = (Ho + kPoolR - 1) / kPoolR * kPoolR;
  if (RB > ho_up) RB = ho_up;
  const size_t smem = (size_t)(2 * RB + 1) * row_bytes;
  RIB_REQUIRE(smem <= 200 * 1024 && smem < (1u << 20), "avgpool: rows wider than 1408 pixels are not supported");
  if (smem > 48 * 1024)
    RIB_CHECK_CUDA(cudaFuncSetAttribute((const void*)avgpool3s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)((Ho + RB - 1) / RB), (unsigned)(C / 8), (unsigned)B);
  launch_pdl(avgpool3s2_kernel, grid, dim3(256), smem, s, src, src_bs, dst, dst_bs, stats, H, W, C, RB);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Composite
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t to_u8(float x) {
  double v = (double)x * 0.5 + 0.5;  // tensor2images multiplies a float32 array by float64 std/mean arrays
  v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
  return (uint8_t)(v * 255.0);
}

__global__ void composite_kernel(const float* __restrict__ img, const float* __restrict__ mask,
                                 const float* __restrict__ dain, float* __restrict__ out_f32,
                                 uint8_t* __restrict__ out_u8, int HW4, int HW, size_t total, long long img_bs,
                                 long long f32_bs, long long u8_bs) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B * HW/4
  if (i >= total) return;
  const size_t n = i / HW4, q = i - n * HW4;
  float mm[4] = {1.f, 1.f, 1.f, 1.f};
  if (mask != nullptr) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mask + n * HW) + q);
    mm[0] = m.x, mm[1] = m.y, mm[2] = m.z, mm[3] = m.w;
  }
  float res[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t off = (n * 3 + c) * (size_t)HW;
    const float4 a = __ldg(reinterpret_cast<const float4*>(img + n * img_bs + (size_t)c * HW) + q);
    const float aa[4] = {a.x, a.y, a.z, a.w};
    if (mask != nullptr) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dain + off) + q);
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)  // separate roundings, as the reference's three elementwise ops
        res[c][k] = __fadd_rn(__fmul_rn(aa[k], mm[k]), __fmul_rn(dd[k], __fsub_rn(1.0f, mm[k])));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) res[c][k] = aa[k];
    }
    if (out_f32 != nullptr)
      reinterpret_cast<float4*>(out_f32 + n * f32_bs + (size_t)c * HW)[q] =
          make_float4(res[c][0], res[c][1], res[c][2], res[c][3]);
  }
  if (out_u8 != nullptr) {
    uint8_t px[12];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) px[3 * k + c] = to_u8(res[c][k]);
    uint32_t* o = reinterpret_cast<uint32_t*>(out_u8 + n * u8_bs + q * 12);
    o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
  }
}

int launch_composite(const float* img, const float* mask, const float* dain, float* out_f32, uint8_t* out_u8, int B,
                     int H, int W, long long img_bstride, long long f32_bstride, long long u8_bstride, cudaStream_t s) {
  RIB_REQUIRE((H * W) % 4 == 0, "composite: H*W must be a multiple of 4");
  const int HW = H * W;
  if (img_bstride == 0) img_bstride = 3LL * HW;
  if (f32_bstride == 0) f32_bstride = 3LL * HW;
  if (u8_bstride == 0) u8_bstride = 3LL * HW;
  RIB_REQUIRE(img_bstride % 4 == 0 && f32_bstride % 4 == 0 && u8_bstride % 4 == 0, "composite: strides must be multiples of 4");
  const size_t total = (size_t)B * (HW / 4);
  const int threads = 256;
  launch_pdl(composite_kernel, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, s,
             img, mask, dain, out_f32, out_u8, HW / 4, HW, total, img_bstride, f32_bstride, u8_bstride);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Decoded frame -> network input: uint8 HWC -> fp32 CHW in [-1, 1]
// (transforms.ToTensor + Normalize(0.5, 0.5): float32(v) / 255, then (t - 0.5) / 0.5, separate roundings;
//  HSM_auto_dataset.py:73-75, used by evaluator.py:223-224).  Four pixels per thread.
// ---------------------------------------------------------------------------------------------
// out_u8 (optional): the frame tensor2images(out) would give (utils/utils.py:122-147) - what the evaluator saves for a key
// frame (evaluator.py:240-244, :265-266); it truncates, so it is NOT the input byte for 63 of the 256 levels.
__global__ void __launch_bounds__(256) frames_from_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                             uint8_t* __restrict__ out_u8, int HW4, int HW, size_t total,
                                                             long long in_bs, long long out_bs, long long u8_bs) {
  // Both results are functions of the input byte alone: one table entry per thread (the exact arithmetic, two IEEE
  // divisions and the float64 truncation of tensor2images), then the pixels are look-ups.
  __shared__ float s_norm[256];
  __shared__ uint8_t s_save[256];
  {
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.0f), 0.5f), 0.5f);
    s_norm[threadIdx.x] = v;
    s_save[threadIdx.x] = to_u8(v);
  }
  __syncthreads();
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B * HW/4
  if (i >= total) return;
  const size_t n = i / HW4, q = i - n * HW4;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in + n * in_bs + q * 12);
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  const uint8_t px[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                          (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                          (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
#pragma unroll
  for (int c = 0; c < 3; ++c)
    reinterpret_cast<float4*>(out + n * out_bs + (size_t)c * HW)[q] =
        make_float4(s_norm[px[c]], s_norm[px[3 + c]], s_norm[px[6 + c]], s_norm[px[9 + c]]);
  if (out_u8 != nullptr) {
    uint32_t qx[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) qx[k] = s_save[px[k]];
    uint32_t* o = reinterpret_cast<uint32_t*>(out_u8 + n * u8_bs + q * 12);
    o[0] = qx[0] | (qx[1] << 8) | (qx[2] << 16) | (qx[3] << 24);
    o[1] = qx[4] | (qx[5] << 8) | (qx[6] << 16) | (qx[7] << 24);
    o[2] = qx[8] | (qx[9] << 8) | (qx[10] << 16) | (qx[11] << 24);
  }
}

int launch_frames_from_u8(const uint8_t* in, float* out, uint8_t* out_u8, int B, int H, int W, long long in_bstride,
                          long long out_bstride, long long u8_bstride, cudaStream_t s) {
  RIB_REQUIRE((H * W) % 4 == 0, "frames_from_u8: H*W must be a multiple of 4");
  const int HW = H * W;
  if (in_bstride == 0) in_bstride = 3LL * HW;
  if (out_bstride == 0) out_bstride = 3LL * HW;
  if (u8_bstride == 0) u8_bstride = 3LL * HW;
  RIB_REQUIRE(in_bstride % 4 == 0 && out_bstride % 4 == 0 && u8_bstride % 4 == 0 && ((uintptr_t)in & 3) == 0 &&
                  ((uintptr_t)out & 15) == 0 && ((uintptr_t)out_u8 & 3) == 0,
              "frames_from_u8: strides / pointers must be 4-element aligned");
  const size_t total = (size_t)B * (HW / 4);
  const int threads = 256;
  launch_pdl(frames_from_u8_kernel, dim3((unsigned)((total + threads - 1) / threads)), dim3(threads), 0, s, in, out, out_u8,
             HW / 4, HW, total, in_bstride, out_bstride, u8_bstride);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Bicubic resize of decoded uint8 frames: cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC) as the evaluator
// applies it through A.Resize to every key frame and DAIN frame (models/evaluator.py:18-26, :218-220).
// OpenCV's published algorithm in its floating-point form (what the IPP-backed cv2 computes): separable 4-tap cubic
// convolution, A = -0.75, source coordinate (d + 0.5) * scale - 0.5 rounded to float32, coefficients in float32,
// taps clamped to the image, sums in double, round-half-even, saturate.  Every operation is written with explicit
// roundings (no FMA contraction) in the order of oracle/resize_oracle.py, so the result equals that restatement bit
// for bit; the oracle itself is pinned to cv2 within one level on <= 0.05 % of the pixels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_taps(int d, double scale, int src, int* idx, double* coef) {
  const float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  const float fl = floorf(f);
  const int s = (int)fl;
  const float x = __fsub_rn(f, fl);
  const float a = -0.75f;
  const float x1 = __fadd_rn(x, 1.0f);
  const float c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(a, x1), __fmul_rn(5.0f, a)), x1), __fmul_rn(8.0f, a)), x1),
                             __fmul_rn(4.0f, a));
  const float a2 = __fadd_rn(a, 2.0f), a3 = __fadd_rn(a, 3.0f);
  const float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(a2, x), a3), x), x), 1.0f);
  const float y = __fsub_rn(1.0f, x);
  const float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(a2, y), a3), y), y), 1.0f);
  const float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c0), c1), c2);
  coef[0] = (double)c0, coef[1] = (double)c1, coef[2] = (double)c2, coef[3] = (double)c3;
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(s - 1 + k, 0), src - 1);
}

__global__ void resize_cubic_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int w, int H,
                                       int W, long long in_bs, long long out_bs) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  if (ox >= W) return;
  int xi[4], yi[4];
  double xc[4], yc[4];
  cubic_taps(ox, (double)w / (double)W, w, xi, xc);
  cubic_taps(oy, (double)h / (double)H, h, yi, yc);
  const uint8_t* src = in + (size_t)n * in_bs;
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint8_t* row = src + (size_t)yi[r] * w * 3;
    double hsum[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* px = row + (size_t)xi[k] * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double prod = __dmul_rn((double)__ldg(px + c), xc[k]);
        hsum[c] = k == 0 ? prod : __dadd_rn(hsum[c], prod);   // ((p0 + p1) + p2) + p3, as numpy's reduction over the tap axis
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double prod = __dmul_rn(hsum[c], yc[r]);
      acc[c] = r == 0 ? prod : __dadd_rn(acc[c], prod);
    }
  }
  uint8_t* o = out + (size_t)n * out_bs + ((size_t)oy * W + ox) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max(__double2int_rn(acc[c]), 0), 255);   // rint: half to even
}

__global__ void copy_frames_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t frame_bytes,
                                      long long in_bs, long long out_bs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < frame_bytes) out[(size_t)blockIdx.y * out_bs + i] = in[(size_t)blockIdx.y * in_bs + i];
}

int launch_resize_cubic_u8(const uint8_t* in, uint8_t* out, int B, int h, int w, int H, int W, long long in_bstride,
                           long long out_bstride, cudaStream_t s) {
  RIB_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && B <= 65535 && H <= 65535, "resize: bad shape");
  if (in_bstride == 0) in_bstride = 3LL * h * w;
  if (out_bstride == 0) out_bstride = 3LL * H * W;
  if (h == H && w == W) {   // cv2.resize returns the image unchanged
    const size_t fb = (size_t)3 * h * w;
    copy_frames_u8_kernel<<<dim3((unsigned)((fb + 255) / 256), (unsigned)B), 256, 0, s>>>(in, out, fb, in_bstride, out_bstride);
  } else {
    resize_cubic_u8_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)B), 128, 0, s>>>(in, out, h, w, H, W,
                                                                                                         in_bstride, out_bstride);
  }
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Flow warp (bilinear, border, align_corners=True)
// ---------------------------------------------------------------------------------------------
__global__ void warp_kernel(const float* __restrict__ src, const float* __restrict__ flow, float* __restrict__ out,
                            int C, int H, int W, float sx, float sy, size_t total, long long src_bs, long long flow_bs,
                            long long out_bs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*H*W
  if (i >= total) return;
  const int HW = H * W;
  const size_t n = i / HW;
  const int hw = (int)(i - n * HW);
  const int y = hw / W, x = hw - y * W;
  const float fx = __ldg(flow + n * flow_bs + hw);
  const float fy = __ldg(flow + n * flow_bs + (size_t)HW + hw);
  // same float sequence as building the normalised grid and un-normalising it in grid_sample (the division by 2 is
  // written as an exact multiplication by 0.5)
  const float gx = __fsub_rn(__fmul_rn(__fadd_rn((float)x, fx), sx), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(__fadd_rn((float)y, fy), sy), 1.0f);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
  ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float tx = ix - x0f, ty = iy - y0f;
  const float wnw = (1.f - tx) * (1.f - ty), wne = tx * (1.f - ty), wsw = (1.f - tx) * ty, wse = tx * ty;
  for (int c = 0; c < C; ++c) {
    const float* s = src + n * src_bs + (size_t)c * HW;
    float acc = __ldg(s + y0 * W + x0) * wnw;
    if (x1 < W) acc += __ldg(s + y0 * W + x1) * wne;
    if (y1 < H) acc += __ldg(s + y1 * W + x0) * wsw;
    if (x1 < W && y1 < H) acc += __ldg(s + y1 * W + x1) * wse;
    out[n * out_bs + (size_t)c * HW + hw] = acc;
  }
}

// Four consecutive pixels of a row per thread (W % 4 == 0): 16-byte flow loads and output stores, 16 gathers in flight.
// Per pixel the float sequence is the one of warp_kernel.
// FH: the flow is stored as IEEE half (exactly converted to fp32 on load; halves the upload of a flow field).
template <int CT, bool FH>   // CT > 0: channel count known at compile time (all gathers of a thread are issued together)
__global__ void warp4_kernel(const float* __restrict__ src, const void* __restrict__ flow_v, float* __restrict__ out,
                             int Crt, int H, int W, float sx, float sy, size_t total4, long long src_bs, long long flow_bs,
                             long long out_bs) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*H*W/4
  if (i >= total4) return;
  const int HW = H * W, HW4 = HW >> 2;
  const size_t n = i / HW4;
  const int hw = (int)(i - n * HW4) * 4;
  const int y = hw / W, xb = hw - y * W;
  float fxs[4], fys[4];
  if (FH) {
    const __half* flow = static_cast<const __half*>(flow_v);
    const uint2 hx = __ldg(reinterpret_cast<const uint2*>(flow + n * flow_bs + hw));
    const uint2 hy = __ldg(reinterpret_cast<const uint2*>(flow + n * flow_bs + (size_t)HW + hw));
    const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&hx.x)), x23 = __half22float2(*reinterpret_cast<const __half2*>(&hx.y));
    const float2 y01 = __half22float2(*reinterpret_cast<const __half2*>(&hy.x)), y23 = __half22float2(*reinterpret_cast<const __half2*>(&hy.y));
    fxs[0] = x01.x, fxs[1] = x01.y, fxs[2] = x23.x, fxs[3] = x23.y;
    fys[0] = y01.x, fys[1] = y01.y, fys[2] = y23.x, fys[3] = y23.y;
  } else {
    const float* flow = static_cast<const float*>(flow_v);
    const float4 fx4 = __ldg(reinterpret_cast<const float4*>(flow + n * flow_bs + hw));
    const float4 fy4 = __ldg(reinterpret_cast<const float4*>(flow + n * flow_bs + (size_t)HW + hw));
    fxs[0] = fx4.x, fxs[1] = fx4.y, fxs[2] = fx4.z, fxs[3] = fx4.w;
    fys[0] = fy4.x, fys[1] = fy4.y, fys[2] = fy4.z, fys[3] = fy4.w;
  }
  int o00[4], dx1[4], dy1[4];
  float wnw[4], wne[4], wsw[4], wse[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float gx = __fsub_rn(__fmul_rn(__fadd_rn((float)(xb + p), fxs[p]), sx), 1.0f);
    const float gy = __fsub_rn(__fmul_rn(__fadd_rn((float)y, fys[p]), sy), 1.0f);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float tx = ix - x0f, ty = iy - y0f;
    wnw[p] = (1.f - tx) * (1.f - ty), wne[p] = tx * (1.f - ty), wsw[p] = (1.f - tx) * ty, wse[p] = tx * ty;
    o00[p] = y0 * W + x0;
    dx1[p] = x0 + 1 < W ? 1 : -1;   // -1: neighbour outside the image (its weight is an exact 0 then; term skipped)
    dy1[p] = y0 + 1 < H ? W : -1;
  }
  const int C = CT > 0 ? CT : Crt;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* s = src + n * src_bs + (size_t)c * HW;
    float acc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float a = __ldg(s + o00[p]) * wnw[p];
      if (dx1[p] > 0) a += __ldg(s + o00[p] + 1) * wne[p];
      if (dy1[p] > 0) a += __ldg(s + o00[p] + W) * wsw[p];
      if (dx1[p] > 0 && dy1[p] > 0) a += __ldg(s + o00[p] + W + 1) * wse[p];
      acc[p] = a;
    }
    *reinterpret_cast<float4*>(out + n * out_bs + (size_t)c * HW + hw) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}


// One pixel per lane, four pixels (32 apart) per thread: the four neighbour gathers of a warp touch one or two cache lines
// each (consecutive lanes read consecutive source pixels displaced by a smooth flow), where the four-consecutive-pixels
// form above spreads every gather instruction over five lines and is bound by the L1 wavefront rate.  Needs
// H * W % 128 == 0 (a group of 128 pixels never straddles two frames).  Same float sequence per pixel as warp_kernel.
template <int CT, bool FH, int PX>
__global__ void __launch_bounds__(256) warp_px_kernel(const float* __restrict__ src, const void* __restrict__ flow_v,
                                                      float* __restrict__ out, int Crt, int H, int W, float sx, float sy,
                                                      size_t ngroups, long long src_bs, long long flow_bs,
                                                      long long out_bs) {
  pdl_wait();   // programmatic dependent launch: nothing above depends on the previous kernel
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // group of 32 PX pixels (one warp)
  if (g >= ngroups) return;
  const int lane = threadIdx.x & 31;
  const int HW = H * W, gpf = HW / (32 * PX);
  const size_t n = g / gpf;
  const int hw0 = (int)(g - n * gpf) * (32 * PX) + lane;
  float fxs[PX], fys[PX];
  if (FH) {
    const __half* flow = static_cast<const __half*>(flow_v) + n * flow_bs;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      fxs[p] = __half2float(flow[hw0 + 32 * p]);
      fys[p] = __half2float(flow[(size_t)HW + hw0 + 32 * p]);
    }
  } else {
    const float* flow = static_cast<const float*>(flow_v) + n * flow_bs;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      fxs[p] = __ldg(flow + hw0 + 32 * p);
      fys[p] = __ldg(flow + (size_t)HW + hw0 + 32 * p);
    }
  }
  int o00[PX], dx1[PX], dy1[PX];
  float wnw[PX], wne[PX], wsw[PX], wse[PX];
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int hw = hw0 + 32 * p;
    const int y = hw / W, x = hw - y * W;
    const float gx = __fsub_rn(__fmul_rn(__fadd_rn((float)x, fxs[p]), sx), 1.0f);
    const float gy = __fsub_rn(__fmul_rn(__fadd_rn((float)y, fys[p]), sy), 1.0f);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float tx = ix - x0f, ty = iy - y0f;
    wnw[p] = (1.f - tx) * (1.f - ty), wne[p] = tx * (1.f - ty), wsw[p] = (1.f - tx) * ty, wse[p] = tx * ty;
    o00[p] = y0 * W + x0;
    dx1[p] = x0 + 1 < W ? 1 : -1;   // -1: neighbour outside the image (its weight is an exact 0 then; term skipped)
    dy1[p] = y0 + 1 < H ? W : -1;
  }
  const int C = CT > 0 ? CT : Crt;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* s = src + n * src_bs + (size_t)c * HW;
    float acc[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float a = __ldg(s + o00[p]) * wnw[p];
      if (dx1[p] > 0) a += __ldg(s + o00[p] + 1) * wne[p];
      if (dy1[p] > 0) a += __ldg(s + o00[p] + W) * wsw[p];
      if (dx1[p] > 0 && dy1[p] > 0) a += __ldg(s + o00[p] + W + 1) * wse[p];
      acc[p] = a;
    }
    float* o = out + n * out_bs + (size_t)c * HW + hw0;
#pragma unroll
    for (int p = 0; p < PX; ++p) o[32 * p] = acc[p];
  }
}

int launch_warp(const float* src, const void* flow_v, int flow_fp16, float* out, int B, int C, int H, int W,
                long long src_bstride, long long flow_bstride, long long out_bstride, cudaStream_t s) {
  const float* flow = static_cast<const float*>(flow_v);
  const size_t total = (size_t)B * H * W;
  if (src_bstride == 0) src_bstride = (long long)C * H * W;
  if (flow_bstride == 0) flow_bstride = 2LL * H * W;
  if (out_bstride == 0) out_bstride = (long long)C * H * W;
  const int threads = 256;
  const float sx = (float)(2.0 / (double)(W > 1 ? W - 1 : 1)), sy = (float)(2.0 / (double)(H > 1 ? H - 1 : 1));
  RIB_REQUIRE(!flow_fp16 || (W % 4 == 0 && flow_bstride % 4 == 0 && out_bstride % 4 == 0 &&
                             ((uintptr_t)flow_v & 7) == 0 && ((uintptr_t)out & 15) == 0),
              "warp: half-precision flows need W % 4 == 0 and aligned frames");
  // pixels per thread: 2 (default), 1, 4.  Measured at the bench shape (tools/bw_bench.py, profiles/r3e_bw_bench.txt):
  // 90 us with 2 (40 registers, twice the resident warps) against 105 / 108 us with 4 / 1.
  static const int px_env = getenv("RIB_WARP_PX") ? atoi(getenv("RIB_WARP_PX")) : 2;
  if (((size_t)H * W) % 128 == 0) {
    const int PX = (px_env == 1 || px_env == 4) ? px_env : 2;
    const size_t ngroups = total / (32 * PX);
    const dim3 grid((unsigned)((ngroups * 32 + threads - 1) / threads)), block(threads);
#define RIB_WARP_LAUNCH(CT, FH, PXV)                                                                                  \
  launch_pdl(warp_px_kernel<CT, FH, PXV>, grid, block, 0, s, src, flow_v, out, C, H, W, sx, sy, ngroups, src_bstride, \
             flow_bstride, out_bstride)
#define RIB_WARP_PICK(PXV)                                     \
  do {                                                         \
    if (flow_fp16) {                                           \
      if (C == 3) RIB_WARP_LAUNCH(3, true, PXV);               \
      else RIB_WARP_LAUNCH(0, true, PXV);                      \
    } else {                                                   \
      if (C == 3) RIB_WARP_LAUNCH(3, false, PXV);              \
      else RIB_WARP_LAUNCH(0, false, PXV);                     \
    }                                                          \
  } while (0)
    if (PX == 1) RIB_WARP_PICK(1);
    else if (PX == 2) RIB_WARP_PICK(2);
    else RIB_WARP_PICK(4);
#undef RIB_WARP_PICK
#undef RIB_WARP_LAUNCH
    RIB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (W % 4 == 0 && flow_bstride % 4 == 0 && out_bstride % 4 == 0 &&
      (((uintptr_t)flow_v & (flow_fp16 ? 7 : 15)) | ((uintptr_t)out & 15)) == 0) {
    const size_t total4 = total / 4;
    const dim3 grid((unsigned)((total4 + threads - 1) / threads)), block(threads);
    if (flow_fp16) {
      if (C == 3) launch_pdl(warp4_kernel<3, true>, grid, block, 0, s, src, flow_v, out, C, H, W, sx, sy, total4, src_bstride, flow_bstride, out_bstride);
      else launch_pdl(warp4_kernel<0, true>, grid, block, 0, s, src, flow_v, out, C, H, W, sx, sy, total4, src_bstride, flow_bstride, out_bstride);
    } else {
      if (C == 3) launch_pdl(warp4_kernel<3, false>, grid, block, 0, s, src, flow_v, out, C, H, W, sx, sy, total4, src_bstride, flow_bstride, out_bstride);
      else launch_pdl(warp4_kernel<0, false>, grid, block, 0, s, src, flow_v, out, C, H, W, sx, sy, total4, src_bstride, flow_bstride, out_bstride);
    }
    RIB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  warp_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(src, flow, out, C, H, W, sx, sy, total,
                                                                              src_bstride, flow_bstride, out_bstride);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Weight folding / packing (runs once at model creation)
// ---------------------------------------------------------------------------------------------
// sigma = u^T W v (weight_norm.py:84-85 -> torch spectral_norm, eval mode).  One block per output row computes
// u[r] * dot(W[r, :], v) in fp64 (fixed-order tree), a second single-block pass adds the rows in a fixed order.
__global__ void sn_sigma_rows_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                     const float* __restrict__ v, int K, double* __restrict__ partial) {
  __shared__ double s_part[256];
  const int r = blockIdx.x;
  const float* wr = w + (size_t)r * K;
  double acc = 0.0;
  for (int c = threadIdx.x; c < K; c += blockDim.x) acc += (double)wr[c] * (double)v[c];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[r] = (double)u[r] * s_part[0];
}

__global__ void sn_sigma_final_kernel(const double* __restrict__ partial, int Cout, float* sigma_inv) {
  __shared__ double s_part[256];
  double acc = 0.0;
  for (int r = threadIdx.x; r < Cout; r += blockDim.x) acc += partial[r];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) sigma_inv[0] = (float)(1.0 / s_part[0]);
}

// `scratch`: at least Cout doubles, re-used by consecutive calls on the same stream.
int launch_sn_sigma_inv(const float* w, const float* u, const float* v, int Cout, int K, float* sigma_inv,
                        double* scratch, cudaStream_t s) {
  RIB_REQUIRE(scratch != nullptr && Cout >= 1, "sn_sigma_inv: bad arguments");
  sn_sigma_rows_kernel<<<Cout, 256, 0, s>>>(w, u, v, K, scratch);
  RIB_CHECK_CUDA(cudaGetLastError());
  sn_sigma_final_kernel<<<1, 256, 0, s>>>(scratch, Cout, sigma_inv);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__device__ __forceinline__ int pack_row(const PackWeightParams& p, int co) {
  if (p.spade_C == 0) return p.row_off + co;
  const int half = co / p.spade_C, c = co - half * p.spade_C;
  const int nq = p.spade_nq > 1 ? p.spade_nq : 1;
  return p.row_off + (c / p.spade_CT) * 2 * nq * p.spade_CT + p.spade_q * 2 * p.spade_CT + half * p.spade_CT + c % p.spade_CT;
}

__global__ void pack_weight_subpix_kernel(const PackWeightParams p) {
  const float scale = p.sigma_inv ? p.sigma_inv[0] : 1.f;
  const int py = p.subpix_parity >> 1, px = p.subpix_parity & 1;
  const size_t total = (size_t)p.Cout * p.Cin * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i & 3), a = tap >> 1, b = tap & 1;
    const size_t r = i >> 2;
    const int ci = (int)(r % p.Cin), co = (int)(r / p.Cin);
    // source rows / columns of the 3x3 kernel that land on low-resolution neighbour a / b for this output parity
    const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    const float* w9 = p.w + r * 9;
    float acc = 0.f;
    for (int rr = r0; rr <= r1; ++rr)
      for (int ss = s0; ss <= s1; ++ss) acc += p.sigma_inv ? w9[rr * 3 + ss] * scale : w9[rr * 3 + ss];
    p.dst[(size_t)pack_row(p, co) * p.ktotal + p.koff + ((ci / p.bkc) * 4 + tap) * p.bkc + ci % p.bkc] = f2act(acc);
  }
  if (p.bias_dst != nullptr)
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < p.Cout; co += gridDim.x * blockDim.x)
      p.bias_dst[pack_row(p, co)] = p.bias ? p.bias[co] : 0.f;
}

__global__ void pack_weight_kernel(const PackWeightParams p) {
  const float scale = p.sigma_inv ? p.sigma_inv[0] : 1.f;
  const size_t total = (size_t)p.Cout * p.Cin * p.taps;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % p.taps);
    const size_t r = i / p.taps;
    const int ci = (int)(r % p.Cin), co = (int)(r / p.Cin);
    // the reference divides the fp32 weight by sigma (W / sigma), then the conv runs on that weight
    const float wv = p.sigma_inv ? p.w[i] * scale : p.w[i];
    p.dst[(size_t)pack_row(p, co) * p.ktotal + p.koff + ((ci / p.bkc) * p.taps + tap) * p.bkc + ci % p.bkc] = f2act(wv);
  }
  if (p.bias_dst != nullptr) {
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < p.Cout; co += gridDim.x * blockDim.x) {
      float b = p.bias ? p.bias[co] : 0.f;
      if (p.spade_C != 0 && co < p.spade_C) b += 1.0f;  // (1 + gamma)
      const int row = pack_row(p, co);
      if (p.bias_accumulate) p.bias_dst[row] += b;
      else p.bias_dst[row] = b;
    }
  }
}

int launch_pack_weight(const PackWeightParams& p, cudaStream_t s) {
  const size_t total = (size_t)p.Cout * p.Cin * p.taps;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 2048u) blocks = 2048u;
  if (p.subpix) {
    RIB_REQUIRE(p.taps == 9 && p.spade_C == 0 && !p.bias_accumulate, "pack: the sub-pixel form needs a plain 3x3 conv");
    pack_weight_subpix_kernel<<<blocks, 256, 0, s>>>(p);
  } else {
    pack_weight_kernel<<<blocks, 256, 0, s>>>(p);
  }
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rib
