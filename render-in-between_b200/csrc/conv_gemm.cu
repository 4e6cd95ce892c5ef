// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).  See conv_gemm.cuh.
//
// CTA = 64 + 128 * kEpiGroups threads, persistent over super-tiles: warp 0 = TMA producer (one lane), warp 1 =
// TMEM allocator + tcgen05.mma issuer (one lane), then kEpiGroups epilogue groups of four warps (TMEM lane
// quarter = warp_idx & 3); group g drains every kEpiGroups-th tile of the CTA.  Two groups (one per accumulator
// buffer) were measured: no gain for the one-CTA-per-SM layers (they are MMA / TMA bound) and a loss for the
// small layers (fewer CTAs per SM), so one group is built.
// Kernels built with the A-operand transform (XF, see ConvGemmParams::xf_*) carry eight more warps that rewrite every
// halo tile in place between its TMA load and the MMAs (instance-norm affine + LeakyReLU of the producer's raw output).
// Barriers: full/empty per ring slot (TMA <-> MMA; XF: TMA -> transform via full, transform -> MMA via ready),
// tmem_full/tmem_empty per accumulator buffer (MMA <-> epilogue), one barrier for the resident weights.
#include "conv_gemm.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace rib {

// Epilogue groups (4 warps each), one per TMEM accumulator buffer.  -DRIB_EPI_GROUPS=2 builds the two-group form (group g
// drains the tiles of accumulator buffer g) for the A/B runs; -DRIB_EPI2_MINB_WIDE=2 then keeps two such CTAs per SM for
// the wide tiles as well (96 registers per thread).
#ifndef RIB_EPI_GROUPS
#define RIB_EPI_GROUPS 1
#endif
#ifndef RIB_EPI2_MINB_WIDE
#define RIB_EPI2_MINB_WIDE 1
#endif
// Resident CTAs per SM the narrow-tile kernels (BN <= 32) are compiled for: 3 caps them at 96 registers per thread (with
// spills in the epilogue), 2 lets them use 168.
// Measured (profiles/r2g): two CTAs per SM without spills beat three with them by 0.5 ms per forward
// (mask.up.2 579 -> 326 us, mask.down_img.0 429 -> 251 us).
#ifndef RIB_SMALL_MINB
#define RIB_SMALL_MINB 2
#endif
static constexpr int kEpiGroups = RIB_EPI_GROUPS;
static constexpr int kThreads = 64 + 128 * kEpiGroups;
// Accumulator buffers in TMEM per CTA.  Two everywhere in the shipped build; -DRIB_ACC4_MAXBN=16|32 builds the narrow
// tiles with four (the MMA lane may then run three tiles ahead of the epilogue) for the A/B runs of tools/gpu_ab.sh.
#ifndef RIB_ACC4_MAXBN
#define RIB_ACC4_MAXBN 0
#endif
__host__ __device__ constexpr int acc_bufs(int BN) { return BN <= RIB_ACC4_MAXBN ? 4 : 2; }
#ifndef RIB_XF_THREADS
#define RIB_XF_THREADS 256
#endif
static constexpr int kXfThreads = RIB_XF_THREADS;          // transform warps of the XF kernels (after the epilogue warps)

// Geometry of tap t of a stage: which halo tile of the slot it reads and the pixel offset of its
// top-left corner inside that tile.
struct TapGeom {
  int tile, poff;
};
__device__ __forceinline__ TapGeom tap_geom(const ConvGemmParams& p, bool src1, int t, int par) {
  TapGeom g;
  g.tile = 0;
  if (p.ntaps == 4) {  // sub-pixel conv: tap (a, b) of output parity (py, px) reads low-res pixel (y - 1 + py + a, x - 1 + px + b)
    g.poff = ((par >> 1) + (t >> 1)) * p.halo_w + (par & 1) + (t & 1);
    return g;
  }
  if (src1) {  // 1x1 second source, loaded with the same halo box: centre pixel
    g.poff = p.halo ? p.halo_w + 1 : 0;
    return g;
  }
  if (p.ntaps == 1) {
    g.poff = p.halo ? p.halo_w + 1 : 0;
    return g;
  }
  const int r = t / 3, s = t - 3 * r;
  if (p.stride == 1) {
    g.poff = r * p.halo_w + s;
  } else {  // input row 2*oy + r - 1 = 2*(oy + dy) + py, halo origin at (oy0 - 1, ox0 - 1)
    g.tile = (r != 1 ? 2 : 0) + (s != 1 ? 1 : 0);
    g.poff = (r == 0 ? 0 : 1) * p.halo_w + (s == 0 ? 0 : 1);
  }
  return g;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_LRELU: return lrelu02(v);
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// Sum each of 16 per-lane values over the 32 lanes of a warp with a transposing butterfly
// (31 shuffles instead of 80).  On return lane l holds the total of column (l >> 1) & 15.
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float keep = b4 ? v[i + 8] : v[i], send = b4 ? v[i] : v[i + 8];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float keep = b3 ? v[i + 4] : v[i], send = b3 ? v[i] : v[i + 4];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float keep = b2 ? v[i + 2] : v[i], send = b2 ? v[i] : v[i + 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    float keep = b1 ? v[1] : v[0], send = b1 ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

__device__ __forceinline__ const act_t* planar_at(const PlanarRef& r, int n, int c, int H, int W, int y, int x) {
  return r.p + (size_t)n * r.bstride + ((size_t)(c >> 3) * H * W + (size_t)y * W + x) * 8 + (c & 7);
}

// Bring-up mainloop: plain loads and FMAs for one pixel x 16 columns (selected only through
// rib_debug_set_simt(); used to bisect tcgen05/TMA problems, never in the measured path).
__device__ void simt_chunk(const ConvGemmParams& p, int n, int oy, int ox, int col0, float* acc, int par = 0) {
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  const int cin0 = p.stages0 * p.BKc, cin1 = p.stages1 * p.BKc;
  for (int ci = 0; ci < cin0; ++ci) {
    const int grp = ci / p.BKc, cc = ci - grp * p.BKc;
    for (int t = 0; t < p.ntaps; ++t) {
      int r = p.ntaps == 9 ? t / 3 : 1, s = p.ntaps == 9 ? t - 3 * (t / 3) : 1;
      if (p.ntaps == 4) {
        r = (par >> 1) + (t >> 1);
        s = (par & 1) + (t & 1);
      }
      const int iy = oy * p.stride + r - 1, ix = ox * p.stride + s - 1;
      if (iy < 0 || ix < 0 || iy >= p.Hin || ix >= p.Win) continue;
      // a stride-2 input is parity-planar: [plane][py][px][H/2][W/2][8]
      const act_t* ap = (p.stride == 2 && p.s2_parity)
                            ? p.src0.p + (size_t)n * p.src0.bstride +
                                  ((size_t)(ci >> 3) * p.Hin * p.Win + (size_t)((iy & 1) * 2 + (ix & 1)) * (p.Hin / 2) * (p.Win / 2) +
                                   (size_t)(iy >> 1) * (p.Win / 2) + (ix >> 1)) * 8 + (ci & 7)
                            : planar_at(p.src0, n, ci, p.Hin, p.Win, iy, ix);
      float av = act2f(*ap);
      if (p.xf_stats != nullptr) {  // A-operand transform, same arithmetic as the transform warps (result rounded to 16 bits)
        const double cnt = (double)p.Hin * (double)p.Win;
        const double s1 = stat_sum(p.xf_stats + ((size_t)n * cin0 + ci) * 2), s2 = stat_sumsq(p.xf_stats + ((size_t)n * cin0 + ci) * 2);
        const double mean = s1 / cnt;
        double var = s2 / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const double rstd = 1.0 / sqrt(var + (double)p.eps);
        const double g = p.xf_w ? (double)p.xf_w[ci] : 1.0, be = p.xf_b ? (double)p.xf_b[ci] : 0.0;
        av = fmaf(av, (float)(g * rstd), (float)(be - mean * g * rstd));
        if (p.xf_act) av = fmaxf(av, 0.2f * av);
        av = act2f(f2act(av));
      }
      const int k = (grp * p.ntaps + t) * p.BKc + cc;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] += av * act2f(p.wpk[(size_t)(col0 + c) * p.ktotal + k]);
    }
  }
  for (int ci = 0; ci < cin1; ++ci) {
    const float av = act2f(*planar_at(p.src1, n, ci, p.H, p.W, oy, ox));
    const int k = cin0 * p.ntaps + ci;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] += av * act2f(p.wpk[(size_t)(col0 + c) * p.ktotal + k]);
  }
}

// named barrier of one epilogue group (128 threads)
__device__ __forceinline__ void epi_bar(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

// Per-warp staging buffer for the instance-norm statistics: 32 rows (pixels) x 16 columns fp32 with a
// 20-float row pitch (conflict-free 16-byte row writes and scalar column reads).
static constexpr int kStatPitch = 20;
static constexpr int kStatWarpFloats = 32 * kStatPitch;

// State of the MMA issuer that does not change during a launch.
struct IssueCtx {
  uint32_t a_hi, b_hi, idesc;
  uint32_t b_tap16;          // 16-byte units between consecutive weight sub-tiles (taps)
  uint32_t bn;               // TMEM columns per sub-tile
};

// All tcgen05.mma of one channel group (NT taps x MT sub-tiles x KK k-steps), fully unrolled and free of waits:
// the halo tile and all weight sub-tiles of the group are covered by ONE full barrier.  Per MMA the issuing lane
// executes two adds and the instruction itself.
//   a_lo      low descriptor word of the halo tile in this ring slot (tap offset not yet added)
//   b_lo      low descriptor word of the group's first weight sub-tile
//   tap       per-tap A offsets (16-byte units);  mk[m * KK + kk] = m * m_step + kk * a_kk_step
template <int NT, int KK, int MT, bool PAIR>
__device__ __forceinline__ void issue_group(const IssueCtx& c, uint32_t a_lo, const uint32_t* tap, const uint32_t* mk,
                                            uint32_t b_lo, uint32_t d0, uint32_t acc_first) {
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const uint32_t a_t = a_lo + tap[t];
    const uint32_t b_t = b_lo + (uint32_t)t * c.b_tap16;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        const uint32_t acc = (t == 0 && kk == 0) ? acc_first : 1u;
        const uint64_t ad = ((uint64_t)c.a_hi << 32) | (uint64_t)(a_t + mk[m * KK + kk]);
        const uint64_t bd = ((uint64_t)c.b_hi << 32) | (uint64_t)(b_t + 2u * (uint32_t)kk);
        if (PAIR) umma_f16_pair(d0 + (uint32_t)m * c.bn, ad, bd, c.idesc, acc);
        else umma_f16(d0 + (uint32_t)m * c.bn, ad, bd, c.idesc, acc);
      }
    }
  }
}

// Sub-pixel conv with several output parities per CTA: parity q runs its four taps (their A offsets come from the
// table in shared memory: tapq[q * 4 + t]) against rows [q * bnp, (q + 1) * bnp) of every weight sub-tile and
// accumulates into columns [q * bnp, (q + 1) * bnp) of the tile.  The halo tile is read from HBM / L2 once for all of them.
template <int KK, int MT>
__device__ __forceinline__ void issue_group_subpix(int ppc, const IssueCtx& c, uint32_t a_lo, uint32_t tapq_addr,
                                                   const uint32_t* mk, uint32_t b_lo, uint32_t d0, uint32_t acc_first,
                                                   uint32_t bq16, uint32_t bnp) {
#pragma unroll 1
  for (int q = 0; q < ppc; ++q) {
    uint32_t tap[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) tap[t] = ld_shared_u32(tapq_addr + (uint32_t)(q * 4 + t) * 4u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t a_t = a_lo + tap[t];
      const uint32_t b_t = b_lo + (uint32_t)t * c.b_tap16 + (uint32_t)q * bq16;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const uint32_t acc = (t == 0 && kk == 0) ? acc_first : 1u;
          umma_f16(d0 + (uint32_t)m * c.bn + (uint32_t)q * bnp, ((uint64_t)c.a_hi << 32) | (uint64_t)(a_t + mk[m * KK + kk]),
                   ((uint64_t)c.b_hi << 32) | (uint64_t)(b_t + 2u * (uint32_t)kk), c.idesc, acc);
        }
      }
    }
  }
}

template <int KK, int MT, bool PAIR>
__device__ __forceinline__ void issue_group_nt(int nt, const IssueCtx& c, uint32_t a_lo, const uint32_t* tap,
                                               const uint32_t* mk, uint32_t b_lo, uint32_t d0, uint32_t acc_first) {
  if (nt == 9) issue_group<9, KK, MT, PAIR>(c, a_lo, tap, mk, b_lo, d0, acc_first);
  else if (nt == 4) issue_group<4, KK, MT, PAIR>(c, a_lo, tap, mk, b_lo, d0, acc_first);
  else issue_group<1, KK, MT, PAIR>(c, a_lo, tap, mk, b_lo, d0, acc_first);
}

// ---------------------------------------------------------------------------------------------
// Specialised EPI_STORE epilogue of one M = 128 sub-tile.  The generic code in the kernel decides everything per 16-column
// chunk at run time (statistics? residual? activation? second output? merged outputs? ragged tile?) and executes about
// 110 instructions per chunk for a plain bias + LeakyReLU + store layer, 45 of them branches, predicated-off loads and
// selects (profiles/r2b source capture of emb_0).  The layers of the generator only use six combinations, so the kernel
// picks one of these instantiations once per launch; ALLVALID (decided per sub-tile with one vote) drops the per-lane
// masking of ragged tiles.  The arithmetic, its order and the statistics are those of the generic path.
// ---------------------------------------------------------------------------------------------
enum { EF_STATS = 1, EF_RES = 2, EF_LRELU = 4, EF_OUT2 = 8, EF_SEG = 16 };

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

#ifndef RIB_REGSTATS_MAXBN
#define RIB_REGSTATS_MAXBN 16
#endif

struct EpiFastArgs {
  uint32_t trow;          // tensor-memory address of this thread's lane quarter, column 0 of the sub-tile
  act_t* obase;           // first output: plane 0 of this thread's pixel; planes are `oplane` elements apart
  size_t oplane;
  act_t* obase_b;         // EF_SEG: second output (chunks >= seg_chunk), planes HW8 apart
  int seg_chunk, nch_eff;
  act_t* obase2;          // EF_OUT2: parity-planar copy, planes HW8 apart
  size_t HW8;
  const act_t* rbase;     // EF_RES: residual, planes rHW8 apart
  size_t rHW8;
  float* sbuf;            // EF_STATS: this warp's transposition buffer
  int lane;
};

template <int BN, int EF, bool ALLVALID>
__device__ __forceinline__ void epi_store_fast(const EpiFastArgs& a, bool valid, float (&acc1)[BN / 16], float (&acc2)[BN / 16],
                                               float* ps1, float* ps2) {
  constexpr int NCH = BN / 16;
  constexpr bool STATS = (EF & EF_STATS) != 0, RES = (EF & EF_RES) != 0, LRELU = (EF & EF_LRELU) != 0;
  constexpr bool OUT2 = (EF & EF_OUT2) != 0, SEG = (EF & EF_SEG) != 0;
  constexpr bool kReg = STATS && BN <= RIB_REGSTATS_MAXBN;
  const bool ok = ALLVALID || valid;
  const int st_col = a.lane & 15, st_half = a.lane >> 4;
  // The accumulator columns of chunk j + 1 are fetched while chunk j is processed (measured, profiles/r2e: fetching two
  // or four chunks ahead is slower: 13.75 -> 13.85 / 14.5 ms per forward).  The bias is already in the accumulator: the
  // MMA lane opens every tile with a K = 16 MMA of a ones tile against the bias rows (see the kernel), so there is
  // neither a shared-memory load nor an add per value here.
  uint32_t r[2][16];
  tmem_ld16_issue(a.trow, r[0]);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    if (SEG && j >= a.nch_eff) break;   // merged convs: the padding columns are skipped
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (RES && ok) {
      r0 = *reinterpret_cast<const uint4*>(a.rbase + (size_t)(2 * j) * a.rHW8);
      r1 = *reinterpret_cast<const uint4*>(a.rbase + (size_t)(2 * j + 1) * a.rHW8);
    }
    tmem_ld16_wait(r[j & 1]);
    if (j + 1 < NCH && (!SEG || j + 1 < a.nch_eff)) tmem_ld16_issue(a.trow + (uint32_t)((j + 1) * 16), r[(j + 1) & 1]);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[j & 1][c]);
    if (RES) {
      const uint32_t ru[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float x0, x1;
        unpack2(ru[c], x0, x1);
        v[2 * c] += x0;
        v[2 * c + 1] += x1;
      }
    }
    if (STATS) {
      if (kReg) {   // running sums of this thread's pixel row (combined over the lanes when the image changes)
        if (ok) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            ps1[kReg ? j * 16 + c : 0] += v[c];
            ps2[kReg ? j * 16 + c : 0] = fmaf(v[c], v[c], ps2[kReg ? j * 16 + c : 0]);
          }
        }
      } else {      // transpose through shared memory: rows = pixels of this warp, then per-lane column sums
        float4* srow = reinterpret_cast<float4*>(a.sbuf + a.lane * kStatPitch);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          srow[c4] = ok ? make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int row = (i >> 2) * 8 + st_half * 4 + (i & 3);
          const float xv = a.sbuf[row * kStatPitch + st_col];
          s1 += xv;
          s2 = fmaf(xv, xv, s2);
        }
        acc1[j] += s1;
        acc2[j] += s2;
        __syncwarp();
      }
    }
    if (LRELU) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c], 0.2f * v[c]);
    }
    uint32_t o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = pack2(v[2 * c], v[2 * c + 1]);
    if (ok) {
      act_t* ob = a.obase + (size_t)(2 * j) * a.oplane;
      size_t pl = a.oplane;
      if (SEG && j >= a.seg_chunk) {
        ob = a.obase_b + (size_t)(2 * (j - a.seg_chunk)) * a.HW8;
        pl = a.HW8;
      }
      *reinterpret_cast<uint4*>(ob) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(ob + pl) = make_uint4(o[4], o[5], o[6], o[7]);
      if (OUT2) {
        *reinterpret_cast<uint4*>(a.obase2 + (size_t)(2 * j) * a.HW8) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(a.obase2 + (size_t)(2 * j + 1) * a.HW8) = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

template <int BN, int EF>
__device__ __forceinline__ void epi_store_fast_v(const EpiFastArgs& a, bool valid, float (&acc1)[BN / 16], float (&acc2)[BN / 16],
                                                 float* ps1, float* ps2) {
  if (__all_sync(0xffffffffu, valid)) epi_store_fast<BN, EF, true>(a, valid, acc1, acc2, ps1, ps2);
  else epi_store_fast<BN, EF, false>(a, valid, acc1, acc2, ps1, ps2);
}

// PAIR: the CTA is one half of a CTA pair (cluster of two on the SMs of one TPC).  The pair computes two adjacent
// super-tiles of the same N tile with tcgen05.mma.cta_group::2 (M = 256: 128 pixels from each CTA's shared memory, the
// N = BN weight rows split half / half between the two CTAs' shared memory), issued by the leader (cluster rank 0).
// Each CTA therefore stages only half of every weight sub-tile: half the L2 -> SM weight traffic and half the weight
// reads from shared memory per MMA.  Barriers: both producers arrive (with their byte counts) on the LEADER's full
// barrier; the leader's tcgen05.commit is multicast to the empty / accumulator-full barriers of both CTAs; the
// epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
template <int MODE, int BN, bool SIMT, bool XF, bool PAIR>
__global__ void __launch_bounds__(kThreads + (XF ? kXfThreads : 0), (SIMT || XF || PAIR) ? 1 : (kEpiGroups == 1 ? (BN <= 32 ? RIB_SMALL_MINB : 2) : (BN <= 32 ? 2 : RIB_EPI2_MINB_WIDE))) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  constexpr int NCH = BN / 16;                         // 16-column chunks of the N tile
  constexpr bool kSpade = MODE == EPI_SPADE || MODE == EPI_SPADE2;
  constexpr int NQ = MODE == EPI_SPADE2 ? 2 : 1;       // SPADE outputs per tile ([gamma | beta] pairs side by side)
  constexpr int CT = BN / (2 * NQ);                    // SPADE: channels per tile
  const int G = p.stages0 + p.stages1;                 // channel groups (halo tiles) per output tile
  const int n_bt = p.stages0 * p.ntaps + p.stages1;    // weight sub-tiles per output tile
  // ring of group slots: [halo tile(s) | (streamed mode) the group's weight sub-tiles]; then the resident weights
  uint8_t* sA = smem;
  uint8_t* sB = sA + (((size_t)p.a_ring * p.g_slot_bytes + 1023) & ~(size_t)1023);  // swizzled tiles: 1024-byte aligned
  uint8_t* sStat = sB + (size_t)(p.b_resident ? n_bt : 0) * p.b_tap_bytes;
  const bool want_stats = MODE == EPI_STORE && p.stats != nullptr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + (want_stats ? 4 * kEpiGroups * kStatWarpFloats * 4 : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_ring;
  uint64_t* a_ready = a_empty + p.a_ring;        // XF only: halo tile transformed, MMA may read it
  constexpr int NB = acc_bufs(BN);                                       // accumulator buffers
  uint64_t* tmem_full_bar = a_ready + (XF ? p.a_ring : 0);  // [NB]
  uint64_t* tmem_empty_bar = tmem_full_bar + NB;  // [NB]
  uint64_t* bres_bar = tmem_empty_bar + NB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 15) & ~(uintptr_t)15);  // [BN], 16-byte aligned
  float* s_aux = s_bias + BN;                              // per epilogue group: STORE [4 warps][2*BN] stats; SPADE [2*CT] rstd, -mean*rstd
  uint32_t* s_tapoff = reinterpret_cast<uint32_t*>(s_aux + kEpiGroups * 8 * BN);  // [16] A start offset (16-byte units) of tap t (sub-pixel conv: [parity q][tap]); [16]: 1x1 second source
  // Bias through the tensor core: a K = 16 MMA of a "ones" A tile (1.0 in k = 0 and k = 1 of every row; one 128-byte core
  // matrix per 8-channel plane, addressed by all sixteen row groups: SBO = 0) against a bias B tile (row n: k = 0 the 16-bit
  // rounding of bias[n], k = 1 the 16-bit rounding of the remainder) opens the accumulation of every sub-tile, so the
  // epilogues neither load nor add the bias (4 shared loads + 16 adds per 16 columns per thread before).
  uint8_t* s_ones = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_tapoff + 20) + 127) & ~(uintptr_t)127);  // 256 B
  uint8_t* s_btile = s_ones + 256;                         // [BN / 8][2 k-halves][8 rows][16 B] un-swizzled K-major, BN * 32 bytes
  float* s_xf = reinterpret_cast<float*>(s_btile + BN * 32);   // XF: [2][cin0] scale, shift of the current image

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;   // 0 = the pair's leader
  const int cta_lin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // pair index / CTA index
  const int n_ctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tstride = PAIR ? 2 : 1;                            // super-tiles advanced per iteration
  const int ntile = cta_lin % p.n_tiles;
  // sub-pixel conv: an N tile covers ppc consecutive output parities (from `par`) x channels [ncol0, ncol0 + BN / ppc)
  const int ppc = p.subpix ? p.ppc : 1;
  const int bnp = BN / ppc;                                                      // columns per parity
  const int tiles_per_pg = p.subpix ? p.n_tiles / (4 / ppc) : p.n_tiles;
  const int pgrp = ntile / tiles_per_pg;
  const int par = pgrp * ppc;
  const int ncol0 = (ntile - pgrp * tiles_per_pg) * bnp;
  const int cta_m = cta_lin / p.n_tiles;
  const int cta_groups = n_ctas / p.n_tiles;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const long long total_tiles = (long long)tiles_per_img * p.B;
  // contiguous range of super-tiles: neighbouring tiles share halo rows in L2 and an image's statistics
  // are flushed by few CTAs
  // (a pair walks pairs of adjacent super-tiles: 2 i + rank.  With an odd tile count the last pair's second tile lies
  //  behind the last image: its TMA boxes are out of bounds (zero fill, bytes still counted), its MMAs run on zeros and
  //  its epilogue stores nothing)
  const long long units = PAIR ? (total_tiles + 1) >> 1 : total_tiles;
  const int t_begin = (int)(units * cta_m / cta_groups) * tstride + (int)cta_rank;
  const int t_end = (int)(units * (cta_m + 1) / cta_groups) * tstride;
  const int acc_cols = p.MT * BN;  // TMEM columns of one accumulator buffer
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < NB * acc_cols) tmem_cols <<= 1;

  if (!SIMT) {
    if (warp == 0 && lane == 0) {
      prefetch_tmap(&p.amap[0]);
      prefetch_tmap(&p.bmap);
      for (int i = 0; i < p.a_ring; ++i) {
        mbar_init(smem_u32(&a_full[i]), PAIR ? 2 : 1);       // pair: one arrival (+ bytes) per CTA, on the leader's barrier
        mbar_init(smem_u32(&a_empty[i]), 1);
        if (XF) mbar_init(smem_u32(&a_ready[i]), kXfThreads / 32);
      }
      for (int i = 0; i < NB; ++i) {
        mbar_init(smem_u32(&tmem_full_bar[i]), 1);
        mbar_init(smem_u32(&tmem_empty_bar[i]), PAIR ? 8 : 4);   // pair: the epilogue warps of both CTAs
      }
      mbar_init(smem_u32(bres_bar), PAIR ? 2 : 1);   // pair: both CTAs' weight halves are counted on the leader's barrier
      fence_barrier_init();
    }
    if (warp == 1) {
      if (PAIR) {
        tmem_alloc_pair(smem_u32(tmem_ptr), tmem_cols);
        tmem_relinquish_pair();
      } else {
        tmem_alloc(smem_u32(tmem_ptr), tmem_cols);
        tmem_relinquish();
      }
    }
  }
  if (warp >= 2 && warp < 6) {
    const int e = threadIdx.x - 64;
    if (e < 17) {
      // sub-pixel conv: entry q * 4 + t = tap t of output parity par + q; otherwise entry t = tap t, entry 16 = 1x1 source
      const TapGeom tg = p.subpix ? tap_geom(p, false, e & 3, par + (e >> 2)) : tap_geom(p, e == 16, e == 16 ? 0 : (e < 9 ? e : 0), 0);
      s_tapoff[e] = (uint32_t)tg.tile * (p.a_tile_bytes >> 4) + (uint32_t)tg.poff;
    }
    // (sub-pixel convs with several parities per CTA issue N = BN / ppc MMAs per parity and keep the epilogue add)
    const bool bias_mma = !SIMT && ppc == 1;
    for (int c = e; c < BN; c += 128) s_bias[c] = bias_mma ? 0.f : p.bias[ntile * BN + c];
    if (bias_mma) {
      // ones tile: 2 core matrices x 8 rows x 8 elements; rows of plane 0 are {1, 1, 0, 0, 0, 0, 0, 0}
      reinterpret_cast<act_t*>(s_ones)[e] = (e < 64 && (e & 7) < 2) ? f2act(1.0f) : f2act(0.0f);
      const int nb = PAIR ? BN / 2 : BN, n0 = ntile * BN + (PAIR ? (int)cta_rank * (BN / 2) : 0);
      for (int nrow = e; nrow < nb; nrow += 128) {
        const float b = p.bias[n0 + nrow];
        const act_t hi = f2act(b), lo = f2act(b - act2f(hi));
        uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) << 16);
        uint4* row = reinterpret_cast<uint4*>(s_btile + (nrow >> 3) * 256 + (nrow & 7) * 16);
        row[0] = make_uint4(w0, 0u, 0u, 0u);   // k = 0..7
        row[8] = make_uint4(0u, 0u, 0u, 0u);   // k = 8..15 (the second core matrix, 128 bytes on)
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's shared-memory reads
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = SIMT ? 0u : *tmem_ptr;
  // Programmatic dependent launch: the next conv kernel of the stream may be scheduled from here on (its CTAs take an SM
  // as soon as one of ours leaves and run their prologue - barrier init, tensor-memory allocation, bias tile, resident
  // weights - while our other CTAs finish); its own pdl_wait() holds all its activation traffic back until we are done.
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // One elected lane runs the whole role: inside `if (elect_one())` the compiler knows that a single lane is
    // active, so addresses and coordinates move to uniform registers without per-lane loops.
    if (!SIMT && elect_one()) {
      const int ntaps = p.ntaps, stages0 = p.stages0, a_ring = p.a_ring, BKc = p.BKc;
      const int halo = p.halo, stride = p.stride, MT = p.MT, tiles_x = p.tiles_x;
      const bool b_resident = p.b_resident != 0;
      const uint32_t g_slot_bytes = p.g_slot_bytes, a_tile_bytes = p.a_tile_bytes, a_tx_bytes = p.a_tx_bytes;
      const uint32_t b_tap_bytes = p.b_tap_bytes, b_off = p.b_off;
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      // pair: completions are signalled on the leader's full barriers (shared::cluster addresses)
      const uint32_t a_full0 = PAIR ? mapa_shared(smem_u32(a_full), 0u) : smem_u32(a_full), a_empty0 = smem_u32(a_empty);
      const int planes_per_group = BKc >> 3;
      const int n_col0 = ntile * BN + (PAIR ? (int)cta_rank * (BN / 2) : 0);   // pair: this CTA stages half of the weight rows
      if (b_resident) {  // all weight sub-tiles of this N tile (pair: this CTA's half of their rows), once
        const uint32_t bb = PAIR ? mapa_shared(smem_u32(bres_bar), 0u) : smem_u32(bres_bar);
        if (PAIR) mbar_arrive_expect_tx_cluster(bb, (uint32_t)n_bt * b_tap_bytes);
        else mbar_arrive_expect_tx(bb, (uint32_t)n_bt * b_tap_bytes);
        for (int i = 0; i < n_bt; ++i) {
          if (PAIR) tma_load_2d_pair(sB_addr + (uint32_t)i * b_tap_bytes, &p.bmap, bb, i * BKc, n_col0);
          else tma_load_2d(sB_addr + (uint32_t)i * b_tap_bytes, &p.bmap, bb, i * BKc, n_col0);
        }
      }
      pdl_wait();   // the weights above are constants; the activations below are the previous kernel's output
      int a_slot = 0;
      uint32_t a_phase = 0;
      // tile coordinates are advanced incrementally (no divisions in the loop)
      int n = t_begin / tiles_per_img;
      int rem0 = t_begin - n * tiles_per_img;
      int tile_y = rem0 / tiles_x, tile_x = rem0 - tile_y * tiles_x;
      const int tiles_y = p.tiles_y;
      // (gather mode: the halo tiles are brought in by the extra warps, the weights are resident: nothing left to do)
      for (int mt = (XF && p.a_gather) ? t_end : t_begin; mt < t_end; mt += tstride) {
        const int oy0 = tile_y * p.th * MT, ox0 = tile_x * p.tw;
        for (int g = 0; g < G; ++g) {
          mbar_wait(a_empty0 + 8u * a_slot, a_phase ^ 1u);
          const uint32_t fb = a_full0 + 8u * a_slot;
          const bool src1 = g >= stages0;
          const uint32_t a_dst = sA_addr + (uint32_t)a_slot * g_slot_bytes;
          const int cg = (src1 ? g - stages0 : g) * planes_per_group;  // first 8-channel plane of the group
          const int nt = src1 ? 1 : ntaps;
          // one barrier covers the halo tile and (streamed mode) every weight sub-tile of the group
          if (PAIR) mbar_arrive_expect_tx_cluster(fb, a_tx_bytes + (b_resident ? 0u : (uint32_t)nt * b_tap_bytes));
          else mbar_arrive_expect_tx(fb, a_tx_bytes + (b_resident ? 0u : (uint32_t)nt * b_tap_bytes));
          if (stride == 1) {
            if (PAIR) tma_load_4d_pair(a_dst, &p.amap[src1 ? 1 : 0], fb, (ox0 - halo) * 8, oy0 - halo, cg, n);
            else tma_load_4d(a_dst, &p.amap[src1 ? 1 : 0], fb, (ox0 - halo) * 8, oy0 - halo, cg, n);
          } else {
            if (p.s2_parity) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (PAIR) tma_load_4d_pair(a_dst + q * a_tile_bytes, &p.amap[q], fb, (ox0 - 1) * 8, oy0 - 1, cg, n);
                else tma_load_4d(a_dst + q * a_tile_bytes, &p.amap[q], fb, (ox0 - 1) * 8, oy0 - 1, cg, n);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (PAIR) tma_load_5d_pair(a_dst + q * a_tile_bytes, &p.amap[q], fb, 0, ox0 - 1, oy0 - 1, cg, n);
                else tma_load_5d(a_dst + q * a_tile_bytes, &p.amap[q], fb, 0, ox0 - 1, oy0 - 1, cg, n);
              }
            }
          }
          if (!b_resident) {
            const int i0 = src1 ? stages0 * ntaps + (g - stages0) : g * ntaps;
#pragma unroll 1
            for (int t = 0; t < nt; ++t) {
              if (PAIR) tma_load_2d_pair(a_dst + b_off + (uint32_t)t * b_tap_bytes, &p.bmap, fb, (i0 + t) * BKc, n_col0);
              else tma_load_2d(a_dst + b_off + (uint32_t)t * b_tap_bytes, &p.bmap, fb, (i0 + t) * BKc, n_col0);
            }
          }
          if (++a_slot == a_ring) {
            a_slot = 0;
            a_phase ^= 1u;
          }
        }
        for (int k = 0; k < tstride; ++k) {
          if (++tile_x == tiles_x) {
            tile_x = 0;
            if (++tile_y == tiles_y) {
              tile_y = 0;
              ++n;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One elected lane; the MMAs of a channel group are fully unrolled (issue_group).
    if (!SIMT && (!PAIR || cta_rank == 0u) && elect_one()) {   // pair: only the leader issues (for both CTAs)
      const int ntaps = p.ntaps, stages0 = p.stages0, a_ring = p.a_ring, MT = p.MT;
      const bool b_resident = p.b_resident != 0;
      const uint32_t g_slot16 = p.g_slot_bytes >> 4, b_off16 = p.b_off >> 4;
      const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
      const uint32_t a_full0 = XF ? smem_u32(a_ready) : smem_u32(a_full), a_empty0 = smem_u32(a_empty);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      const uint32_t b_row_bytes = (uint32_t)p.BKc * 2u;
      const int kk_steps = p.BKc >> 4;
      // descriptor halves that never change (see make_nosw_desc / make_kmajor_desc); start addresses (>> 4) are
      // below 2^14, so they are simply added to the low word
      const uint32_t a_lo_c = ((p.lbo >> 4) & 0x3fffu) << 16;
      const uint32_t b_lo_c = 1u << 16;
      const uint32_t b_layout = b_row_bytes == 128 ? 2u : (b_row_bytes == 64 ? 4u : 6u);
      IssueCtx c;
      c.a_hi = ((p.sbo >> 4) & 0x3fffu) | (1u << 14);
      c.b_hi = (((8u * b_row_bytes) >> 4) & 0x3fffu) | (1u << 14) | (b_layout << 29);
      c.idesc = p.idesc;
      c.b_tap16 = p.b_tap_bytes >> 4;
      c.bn = BN;
      const uint32_t a_kk_step = (2u * p.lbo) >> 4;             // K = 16 is two 8-channel planes
      const uint32_t m_step = (uint32_t)(p.th * p.halo_w);    // 16-byte pixels between stacked sub-tiles
      uint32_t tap[9], mk[8];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap[t] = s_tapoff[t];
      const uint32_t tap1 = s_tapoff[16];
      const uint32_t tapq_addr = smem_u32(s_tapoff);
      const uint32_t bq16 = ((uint32_t)bnp * b_row_bytes) >> 4;   // 16-byte units between the parity row blocks of a weight sub-tile
#pragma unroll
      for (int i = 0; i < 8; ++i) mk[i] = (uint32_t)(i / kk_steps) * m_step + (uint32_t)(i % kk_steps) * a_kk_step;
      if (b_resident) {
        mbar_wait(smem_u32(bres_bar), 0u);
        tc_fence_after();
      }
      const bool bias_mma = ppc == 1;
      // un-swizzled K-major descriptors (make_nosw_desc): ones tile LBO = 128, SBO = 0; bias tile LBO = 128, SBO = 256
      const uint64_t ones_desc = ((uint64_t)(1u << 14) << 32) | (uint64_t)(((smem_u32(s_ones) >> 4) & 0x3fffu) | ((128u >> 4) << 16));
      const uint64_t bias_desc = ((uint64_t)((256u >> 4) | (1u << 14)) << 32) |
                                 (uint64_t)(((smem_u32(s_btile) >> 4) & 0x3fffu) | ((128u >> 4) << 16));
      int a_slot = 0;
      uint32_t a_phase = 0;
      int it = 0;
      for (int mt = t_begin; mt < t_end; mt += tstride, ++it) {
        const int buf = (int)((unsigned)it % (unsigned)NB);
        const uint32_t use = (uint32_t)it / (uint32_t)NB;
        mbar_wait(tempty0 + 8u * buf, (use & 1u) ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * acc_cols);
        if (bias_mma) {   // D = ones x bias^T: every row of the tile starts at bias[n] (hi + lo parts)
          for (int m = 0; m < MT; ++m) {
            if (PAIR) umma_f16_pair(d0 + (uint32_t)m * c.bn, ones_desc, bias_desc, c.idesc, 0u);
            else umma_f16(d0 + (uint32_t)m * c.bn, ones_desc, bias_desc, c.idesc, 0u);
          }
        }
        for (int g = 0; g < G; ++g) {
          mbar_wait(a_full0 + 8u * a_slot, a_phase);
          tc_fence_after();
          const bool src1 = g >= stages0;
          const int nt = src1 ? 1 : ntaps;
          const int i0 = src1 ? stages0 * ntaps + (g - stages0) : g * ntaps;
          const uint32_t slot16 = sA16 + (uint32_t)a_slot * g_slot16;
          const uint32_t a_lo = a_lo_c + slot16;
          const uint32_t b_lo = b_lo_c + (b_resident ? sB16 + (uint32_t)i0 * c.b_tap16 : slot16 + b_off16);
          const uint32_t acc_first = (g != 0 || bias_mma) ? 1u : 0u;
          const uint32_t* tp = src1 ? &tap1 : tap;
          if (!PAIR && ppc > 1) {
            if (MT == 1) {
              if (kk_steps == 1) issue_group_subpix<1, 1>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
              else if (kk_steps == 2) issue_group_subpix<2, 1>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
              else issue_group_subpix<4, 1>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
            } else {
              if (kk_steps == 1) issue_group_subpix<1, 2>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
              else if (kk_steps == 2) issue_group_subpix<2, 2>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
              else issue_group_subpix<4, 2>(ppc, c, a_lo, tapq_addr, mk, b_lo, d0, acc_first, bq16, (uint32_t)bnp);
            }
          } else if (MT == 1) {
            if (kk_steps == 1) issue_group_nt<1, 1, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
            else if (kk_steps == 2) issue_group_nt<2, 1, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
            else issue_group_nt<4, 1, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
          } else {
            if (kk_steps == 1) issue_group_nt<1, 2, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
            else if (kk_steps == 2) issue_group_nt<2, 2, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
            else issue_group_nt<4, 2, PAIR>(nt, c, a_lo, tp, mk, b_lo, d0, acc_first);
          }
          // frees the slot (halo tile + streamed weights) once the MMAs have read it (pair: in both CTAs)
          if (PAIR) umma_commit_pair(a_empty0 + 8u * a_slot, (uint16_t)3);
          else umma_commit(a_empty0 + 8u * a_slot);
          if (++a_slot == a_ring) {
            a_slot = 0;
            a_phase ^= 1u;
          }
        }
        if (PAIR) umma_commit_pair(tfull0 + 8u * buf, (uint16_t)3);
        else umma_commit(tfull0 + 8u * buf);
      }
    }
    __syncwarp();
  } else if (XF && warp >= 2 + 4 * kEpiGroups) {
    // ===================== A-operand transform =====================
    // Four warps rewrite every halo tile in place: x -> lrelu?(x * scale[c] + shift[c]) for in-image pixels (padding
    // stays zero), one 8-channel plane per warp at a time with its 16 coefficients in registers.
    pdl_wait();
    if (!SIMT && p.a_gather) {
      // ----- stride-2 gather: parity tile (py, px) of a channel group = input pixels (2 Y + py, 2 X + px) of a NORMAL
      // planar map.  Every thread owns a fixed set of 16-byte vectors of the slot (same for every tile), copies them with
      // cp.async (zero fill outside the image = the conv's zero padding) one ring slot ahead of the one it publishes.
      const int tid = (int)threadIdx.x - kThreads;
      const int stages0 = p.stages0, a_ring = p.a_ring, MT = p.MT, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
      const int planes = p.BKc >> 3, hw = p.halo_w, npix = (int)(p.lbo >> 4);
      const int per_tile = planes * npix, nvec = 4 * per_tile;
      const size_t plane_el = (size_t)p.Hin * p.Win * 8;
      constexpr int kMaxVec = 12;
      uint32_t doff[kMaxVec];   // byte offset inside the slot
      int dyx[kMaxVec];         // (2 hy + py) << 16 | (2 hx + px): input offset from the pixel 2 * (tile origin - 1)
      int plv[kMaxVec];         // 8-channel plane inside the group, -1 = no vector
#pragma unroll
      for (int k = 0; k < kMaxVec; ++k) {
        const int v = tid + k * kXfThreads;
        plv[k] = -1;
        doff[k] = 0u;
        dyx[k] = 0;
        if (v < nvec) {
          const int q = v / per_tile, r = v - q * per_tile, pl = r / npix, pix = r - pl * npix;
          const int hy = pix / hw, hx = pix - hy * hw;
          plv[k] = pl;
          doff[k] = (uint32_t)q * p.a_tile_bytes + (uint32_t)pl * p.lbo + (uint32_t)pix * 16u;
          dyx[k] = ((2 * hy + (q >> 1)) << 16) | (2 * hx + (q & 1));
        }
      }
      const uint32_t sA_addr = smem_u32(sA), a_empty0 = smem_u32(a_empty), a_ready0 = smem_u32(a_ready);
      int a_slot = 0, pend_slot = -1;
      uint32_t a_phase = 0;
      int n = t_begin / tiles_per_img;
      int rem0 = t_begin - n * tiles_per_img;
      int tile_y = rem0 / tiles_x, tile_x = rem0 - tile_y * tiles_x;
      for (int mt = t_begin; mt < t_end; ++mt) {
        const int iy0 = 2 * (tile_y * p.th * MT - 1), ix0 = 2 * (tile_x * p.tw - 1);
        for (int g = 0; g < stages0; ++g) {
          mbar_wait(a_empty0 + 8u * a_slot, a_phase ^ 1u);
          const act_t* src_g = p.src0.p + (size_t)n * p.src0.bstride + (size_t)(g * planes) * plane_el;
          const uint32_t dst0 = sA_addr + (uint32_t)a_slot * p.g_slot_bytes;
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k) {
            if (plv[k] >= 0) {
              const int iy = iy0 + (dyx[k] >> 16), ix = ix0 + (dyx[k] & 0xffff);
              const bool ok = (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win;
              const act_t* src = src_g + (size_t)plv[k] * plane_el + ((size_t)(ok ? iy : 0) * p.Win + (ok ? ix : 0)) * 8;
              cp_async_16(dst0 + doff[k], src, ok ? 16u : 0u);
            }
          }
          cp_async_commit();
          if (pend_slot >= 0) {   // the previous slot's copies have landed: publish it to the MMA lane
            cp_async_wait_group<1>();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready0 + 8u * pend_slot);
          }
          pend_slot = a_slot;
          if (++a_slot == a_ring) {
            a_slot = 0;
            a_phase ^= 1u;
          }
        }
        if (++tile_x == tiles_x) {
          tile_x = 0;
          if (++tile_y == tiles_y) {
            tile_y = 0;
            ++n;
          }
        }
      }
      if (pend_slot >= 0) {
        cp_async_wait_group<0>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready0 + 8u * pend_slot);
      }
    } else if (!SIMT) {
      const int xw = warp - (2 + 4 * kEpiGroups);
      const int stages0 = p.stages0, a_ring = p.a_ring, BKc = p.BKc, MT = p.MT, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
      const int planes = BKc >> 3, ntile_a = p.stride == 2 ? 4 : 1;
      const int hw = p.halo_w, npix = (int)(p.lbo >> 4), halo = p.halo, hh = npix / hw;
      const int S = planes * ntile_a >= 8 ? 1 : (planes * ntile_a >= 4 ? 2 : 4), s_shift = S == 1 ? 0 : (S == 2 ? 1 : 2);
      const int nitems = planes * ntile_a * S, chunk = (npix + S - 1) / S;
      const float inv_hw = 1.0f / (float)hw;
      const int Hs = p.stride == 2 ? p.Hin >> 1 : p.Hin, Ws = p.stride == 2 ? p.Win >> 1 : p.Win;  // extent a tile indexes
      const int org = p.stride == 2 ? 1 : halo;
      const int cin0 = stages0 * BKc;
      const double cnt = (double)p.Hin * (double)p.Win;
      const bool lrelu = p.xf_act != 0;
      int a_slot = 0;
      uint32_t a_phase = 0;
      int n = t_begin / tiles_per_img;
      int rem0 = t_begin - n * tiles_per_img;
      int tile_y = rem0 / tiles_x, tile_x = rem0 - tile_y * tiles_x;
      int cur_n = -1;
      for (int mt = t_begin; mt < t_end; ++mt) {
        if (n != cur_n) {  // uniform over the four warps
          asm volatile("bar.sync 8, %0;" ::"n"(kXfThreads) : "memory");   // nobody still reads the previous image's coefficients
          for (int c = threadIdx.x - (kThreads); c < cin0; c += kXfThreads) {
            const double s1 = stat_sum(p.xf_stats + ((size_t)n * cin0 + c) * 2), s2 = stat_sumsq(p.xf_stats + ((size_t)n * cin0 + c) * 2);
            const double mean = s1 / cnt;
            double var = s2 / cnt - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            const double rstd = 1.0 / sqrt(var + (double)p.eps);
            const double g = p.xf_w ? (double)p.xf_w[c] : 1.0, be = p.xf_b ? (double)p.xf_b[c] : 0.0;
            s_xf[c] = (float)(g * rstd);
            s_xf[cin0 + c] = (float)(be - mean * g * rstd);
          }
          asm volatile("bar.sync 8, %0;" ::"n"(kXfThreads) : "memory");
          cur_n = n;
        }
        const int oy0 = tile_y * p.th * MT - org, ox0 = tile_x * p.tw - org;   // image coordinates of halo pixel (0, 0)
        for (int g = 0; g < stages0; ++g) {
          mbar_wait(smem_u32(&a_full[a_slot]), a_phase);
          uint8_t* slot = sA + (size_t)a_slot * p.g_slot_bytes;
          // work items: (parity tile, plane, pixel range); S ranges per plane so that all warps have work
          const bool interior = oy0 >= 0 && ox0 >= 0 && oy0 + hh <= Hs && ox0 + hw <= Ws;
          for (int it = xw; it < nitems; it += kXfThreads / 32) {
            const int pi = it >> s_shift, sidx = it & (S - 1);
            const int q = pi / planes, pl = pi - q * planes;
            const float* cs = s_xf + g * BKc + pl * 8;
            const float4 sc0 = *reinterpret_cast<const float4*>(cs), sc1 = *reinterpret_cast<const float4*>(cs + 4);
            const float4 sh0 = *reinterpret_cast<const float4*>(cs + cin0), sh1 = *reinterpret_cast<const float4*>(cs + cin0 + 4);
            const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
            const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
            uint4* tile = reinterpret_cast<uint4*>(slot + (size_t)q * p.a_tile_bytes + (size_t)pl * p.lbo);
            const int p_lo = sidx * chunk, p_hi = min(npix, p_lo + chunk);
            auto xform = [&](uint4 v) {
              const uint32_t u[4] = {v.x, v.y, v.z, v.w};
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float a, b;
                unpack2(u[k], a, b);
                a = fmaf(a, sc[2 * k], sh[2 * k]);
                b = fmaf(b, sc[2 * k + 1], sh[2 * k + 1]);
                if (lrelu) {
                  a = fmaxf(a, 0.2f * a);
                  b = fmaxf(b, 0.2f * b);
                }
                o[k] = pack2(a, b);
              }
              return make_uint4(o[0], o[1], o[2], o[3]);
            };
            if (interior) {  // four independent vectors in flight per lane
              for (int base = p_lo + lane; base < p_hi; base += 128) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (base + 32 * u < p_hi) v[u] = tile[base + 32 * u];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (base + 32 * u < p_hi) tile[base + 32 * u] = xform(v[u]);
              }
            } else {
              for (int pix = p_lo + lane; pix < p_hi; pix += 32) {
                const int hy = (int)(((float)pix + 0.5f) * inv_hw), hx = pix - hy * hw;
                const int iy = oy0 + hy, ix = ox0 + hx;
                if ((unsigned)iy >= (unsigned)Hs || (unsigned)ix >= (unsigned)Ws) continue;   // zero padding stays zero
                tile[pix] = xform(tile[pix]);
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&a_ready[a_slot]));
          if (++a_slot == a_ring) {
            a_slot = 0;
            a_phase ^= 1u;
          }
        }
        if (++tile_x == tiles_x) {
          tile_x = 0;
          if (++tile_y == tiles_y) {
            tile_y = 0;
            ++n;
          }
        }
      }
    }
  } else {
    // ===================== Epilogue =====================
    pdl_wait();   // statistics, x, residuals and every store below belong to the stream-ordered part of the kernel
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;                     // epilogue group = accumulator buffer it drains
    const int e = (threadIdx.x - 64) & 127;             // thread index inside the group
    const int prow = q * 32 + lane;           // row of the M=128 sub-tile = TMEM lane
    const int ty = p.tw == 8 ? prow >> 3 : prow >> 5, tx = prow & (p.tw - 1);  // 16 x 8 (or 4 x 32) pixel tile
    const size_t HW8 = (size_t)p.H * p.W * 8;
    const size_t oplane = p.subpix ? 4 * HW8 : HW8;   // EPI_STORE: elements between output planes
    const float slope = p.act == ACT_LRELU ? 0.2f : 1.0f;  // lrelu(v) = max(v, 0.2 v); identity = max(v, v)
    float* sbuf = reinterpret_cast<float*>(sStat) + (warp - 2) * kStatWarpFloats;
    float* s_auxg = s_aux + eg * 8 * BN;                // this group's scratch (statistics slots / SPADE coefficients)
    // column sums: lane l owns column (l & 15) of every chunk for half of the 32 rows
    const int st_col = lane & 15, st_half = lane >> 4;
    float acc1[NCH], acc2[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc1[j] = acc2[j] = 0.f;
    // The narrowest tiles (BN = 16) are bound by the epilogue's instruction latency, not by the tensor pipe: there every
    // thread keeps the running sums of its own pixel row in registers (two FMAs per value, no shuffle, no shared
    // memory) and the 32 lanes are combined only when the image changes.
    constexpr bool kRegStats = MODE == EPI_STORE && BN <= RIB_REGSTATS_MAXBN;
    constexpr int kRegN = kRegStats ? BN : 1;
    float ps1[kRegN], ps2[kRegN];
#pragma unroll
    for (int c = 0; c < kRegN; ++c) ps1[c] = ps2[c] = 0.f;
    int cur_n = -1;
    int it = 0;

    // Deterministic: every warp parks its column sums in its own slot, then one thread per column adds the
    // four slots in a fixed order and issues the fixed-point integer atomics (order-independent, see stat_add).
    auto flush_stats = [&](int n_img) {
      float* slot = s_auxg + ((warp - 2) & 3) * 2 * BN;
      if (kRegStats) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float t1 = warp_colsum16(ps1 + (kRegStats ? j * 16 : 0), lane);  // lanes 2c, 2c+1 hold column c
          const float t2 = warp_colsum16(ps2 + (kRegStats ? j * 16 : 0), lane);
          if ((lane & 1) == 0) {
            slot[j * 16 + (lane >> 1)] = t1;
            slot[BN + j * 16 + (lane >> 1)] = t2;
          }
        }
#pragma unroll
        for (int c = 0; c < kRegN; ++c) ps1[c] = ps2[c] = 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float t1 = acc1[j] + __shfl_xor_sync(0xffffffffu, acc1[j], 16);
          const float t2 = acc2[j] + __shfl_xor_sync(0xffffffffu, acc2[j], 16);
          if (lane < 16) {
            slot[j * 16 + lane] = t1;
            slot[BN + j * 16 + lane] = t2;
          }
          acc1[j] = acc2[j] = 0.f;
        }
      }
      epi_bar(eg);
      for (int c = e; c < BN; c += 128) {
        const int col = ncol0 + (ppc > 1 ? c % bnp : c);   // several parities of a sub-pixel tile feed the same channel
        if (col < p.n_valid) {
          const float t1 = ((s_auxg[c] + s_auxg[2 * BN + c]) + s_auxg[4 * BN + c]) + s_auxg[6 * BN + c];
          const float t2 = ((s_auxg[BN + c] + s_auxg[3 * BN + c]) + s_auxg[5 * BN + c]) + s_auxg[7 * BN + c];
          double* st = (p.seg_cols && col >= p.seg_cols)
                           ? p.stats_b + ((size_t)n_img * p.stats_b_ld + (col - p.seg_cols)) * 2
                           : p.stats + ((size_t)n_img * p.stats_ld + col) * 2;
          stat_add(st, 0, t1);
          stat_add(st, 1, t2);
        }
      }
      epi_bar(eg);
    };

    // EPI_STORE: which specialised epilogue serves this launch (-1: the generic code below)
    int epi_fast = -1;
    if (MODE == EPI_STORE && !SIMT && ppc == 1 && (p.act == ACT_NONE || p.act == ACT_LRELU)) {
      const int ef = (want_stats ? EF_STATS : 0) | (p.has_res ? EF_RES : 0) | (p.act == ACT_LRELU ? EF_LRELU : 0) |
                     (p.has_out2 ? EF_OUT2 : 0) | (p.seg_cols ? EF_SEG : 0);
      if (ef == 0 || ef == EF_LRELU || ef == (EF_LRELU | EF_OUT2) || ef == EF_STATS || ef == (EF_STATS | EF_RES) ||
          ef == (EF_STATS | EF_SEG))
        epi_fast = ef;
    }
#ifdef RIB_NO_FAST_EPI
    epi_fast = -1;
#endif

    // group eg takes tiles t_begin + eg, t_begin + eg + 2, ...; tile coordinates advance incrementally
    const int t_first = t_begin + eg * tstride;
    int n = t_first / tiles_per_img;
    int tile_y = (t_first - n * tiles_per_img) / p.tiles_x;
    int tile_x = (t_first - n * tiles_per_img) - tile_y * p.tiles_x;
    auto advance_tile = [&]() {
      if (++tile_x == p.tiles_x) {
        tile_x = 0;
        if (++tile_y == p.tiles_y) {
          tile_y = 0;
          ++n;
        }
      }
    };
    it = eg;
    for (int mt = t_first; mt < t_end; mt += kEpiGroups * tstride, it += kEpiGroups) {
      const int buf = (int)((unsigned)it % (unsigned)NB);
      const uint32_t use = (uint32_t)it / (uint32_t)NB;
      const int oy0 = tile_y * p.th * p.MT, ox0 = tile_x * p.tw;

      if (n != cur_n && n < p.B) {  // uniform over the 128 epilogue threads (n == B: the dummy tile of an odd pair)
        if (MODE == EPI_STORE && want_stats && cur_n >= 0) flush_stats(cur_n);
        if (kSpade) {
          const int tiles_per_q = p.C / CT;
          const int c0 = (ntile % tiles_per_q) * CT;
          const double cnt = (double)p.Hx * (double)p.Wx;
          epi_bar(eg);
          for (int c = e; c < CT; c += 128) {
            const double s = stat_sum(p.xstats + ((size_t)n * p.C + c0 + c) * 2);
            const double ss = stat_sumsq(p.xstats + ((size_t)n * p.C + c0 + c) * 2);
            const double mean = s / cnt;
            double var = ss / cnt - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            const double rstd = 1.0 / sqrt(var + (double)p.eps);
            s_auxg[c] = (float)rstd;
            s_auxg[CT + c] = (float)(-mean * rstd);
          }
          epi_bar(eg);
        }
        cur_n = n;
      }

      // EPI_SPADE: the x values of a sub-tile are fetched one sub-tile ahead (the first one before waiting for
      // the accumulator), so that their DRAM latency overlaps the MMAs / the previous sub-tile's arithmetic
      // (tiles with more than 64 channels - the 256-column pair tiles - fetch x per 16-channel chunk, one chunk ahead:
      //  they only serve the low-resolution levels, whose x maps are L2-resident)
      constexpr bool kXChunked = kSpade && CT / 16 > 4;
      constexpr int NXC = (kSpade && !kXChunked) ? CT / 16 : 1;
      uint4 xcur[NXC][2], xnext[NXC][2];
      auto load_x = [&](int m, uint4 (*dst)[2]) {
        if (kXChunked) return;
        const int oy = oy0 + m * p.th + ty, ox = ox0 + tx;
        const bool valid = (oy < p.H) && (ox < p.W) && (n < p.B);
        const int tiles_per_q = p.C / CT;
        const int c0 = (ntile % tiles_per_q) * CT;
        const int sy = p.ups ? (oy >> 1) : oy, sx = p.ups ? (ox >> 1) : ox;
        const size_t xHW8 = (size_t)p.Hx * p.Wx * 8;
        const act_t* xrow = p.x.p + (size_t)n * p.x.bstride + (size_t)(c0 >> 3) * xHW8 + ((size_t)sy * p.Wx + sx) * 8;
#pragma unroll
        for (int j = 0; j < NXC; ++j) {
          dst[j][0] = dst[j][1] = make_uint4(0, 0, 0, 0);
          if (valid) {
            dst[j][0] = *reinterpret_cast<const uint4*>(xrow + (size_t)(2 * j) * xHW8);
            dst[j][1] = *reinterpret_cast<const uint4*>(xrow + (size_t)(2 * j + 1) * xHW8);
          }
        }
      };
      if (kSpade) load_x(0, xcur);
      if (!SIMT) {
        mbar_wait(smem_u32(&tmem_full_bar[buf]), use & 1u);
        tc_fence_after();
      }
      for (int m = 0; m < p.MT; ++m) {
        const int oy = oy0 + m * p.th + ty, ox = ox0 + tx;
        const bool valid = (oy < p.H) && (ox < p.W) && (n < p.B);
        const size_t pix8 = ((size_t)oy * p.W + ox) * 8;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + m * BN);

        if (MODE == EPI_STORE) {
          // (sub-pixel conv: planes of the (2H, 2W) output are 4 HW8 apart and parity `par` is a dense H x W image)
          act_t* obase = p.subpix ? p.out.p + (size_t)n * p.out.bstride + ((size_t)(ncol0 >> 3) * 4 + par) * HW8 + pix8
                                  : p.out.p + (size_t)n * p.out.bstride + (size_t)((ntile * BN) >> 3) * HW8 + pix8;
          if (p.out_parity)
            obase = p.out.p + (size_t)n * p.out.bstride + (size_t)((ntile * BN) >> 3) * HW8 +
                    (size_t)((oy & 1) * 2 + (ox & 1)) * (HW8 >> 2) + ((size_t)(oy >> 1) * (p.W >> 1) + (ox >> 1)) * 8;
          // parity-planar copy for a stride-2 consumer: [plane][py][px][H/2][W/2][8]
          // merged convs: base of the second output for this pixel (chunks at and after seg_cols go there)
          act_t* obase_b = nullptr;
          if (p.seg_cols)
            obase_b = p.out_b.p + (size_t)n * p.out_b.bstride +
                      (p.out_b_parity ? (size_t)((oy & 1) * 2 + (ox & 1)) * (HW8 >> 2) + ((size_t)(oy >> 1) * (p.W >> 1) + (ox >> 1)) * 8
                                      : pix8);
          const int nch_eff = p.seg_cols ? (p.n_valid + 15) >> 4 : NCH;   // merged convs: the padding columns are skipped
          act_t* obase2 = nullptr;
          if (p.has_out2)
            obase2 = p.out2.p + (size_t)n * p.out2.bstride + (size_t)((ntile * BN) >> 3) * HW8 +
                     (size_t)((oy & 1) * 2 + (ox & 1)) * (HW8 >> 2) + ((size_t)(oy >> 1) * (p.W >> 1) + (ox >> 1)) * 8;
          // identity shortcut (residual.py:146-151); res_ups: the shortcut is the nearest x2 up-sampling of a half-size map
          // (generator.py:248-249 followed by a block whose input and output widths are equal)
          const size_t rHW8 = p.res_ups ? HW8 >> 2 : HW8;
          const act_t* rbase = p.has_res ? p.res.p + (size_t)n * p.res.bstride + (size_t)((ntile * BN) >> 3) * rHW8 +
                                               (p.res_ups ? ((size_t)(oy >> 1) * (p.W >> 1) + (ox >> 1)) * 8 : pix8)
                                         : nullptr;
          if (epi_fast >= 0) {
            EpiFastArgs fa;
            fa.trow = trow;
            fa.obase = obase;
            fa.oplane = oplane;
            fa.obase_b = obase_b;
            fa.seg_chunk = p.seg_cols >> 4;
            fa.nch_eff = nch_eff;
            fa.obase2 = obase2;
            fa.HW8 = HW8;
            fa.rbase = rbase;
            fa.rHW8 = rHW8;
            fa.sbuf = sbuf;
            fa.lane = lane;
            float* q1 = ps1;
            float* q2 = ps2;
            switch (epi_fast) {
              case 0: epi_store_fast_v<BN, 0>(fa, valid, acc1, acc2, q1, q2); break;
              case EF_LRELU: epi_store_fast_v<BN, EF_LRELU>(fa, valid, acc1, acc2, q1, q2); break;
              case EF_LRELU | EF_OUT2: epi_store_fast_v<BN, EF_LRELU | EF_OUT2>(fa, valid, acc1, acc2, q1, q2); break;
              case EF_STATS: epi_store_fast_v<BN, EF_STATS>(fa, valid, acc1, acc2, q1, q2); break;
              case EF_STATS | EF_RES: epi_store_fast_v<BN, EF_STATS | EF_RES>(fa, valid, acc1, acc2, q1, q2); break;
              default: epi_store_fast_v<BN, EF_STATS | EF_SEG>(fa, valid, acc1, acc2, q1, q2); break;
            }
            continue;   // next sub-tile
          }
          uint32_t r[2][16];
          if (!SIMT) tmem_ld16_issue(trow, r[0]);
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            if (j >= nch_eff) break;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (p.has_res && valid) {
              r0 = *reinterpret_cast<const uint4*>(rbase + (size_t)(2 * j) * rHW8);
              r1 = *reinterpret_cast<const uint4*>(rbase + (size_t)(2 * j + 1) * rHW8);
            }
            float v[16];
            if (SIMT) {
              simt_chunk(p, n, oy, ox, ntile * BN + j * 16, v, par + (ppc > 1 ? (j * 16) / bnp : 0));
            } else {
              tmem_ld16_wait(r[j & 1]);
              if (j + 1 < nch_eff) tmem_ld16_issue(trow + (uint32_t)((j + 1) * 16), r[(j + 1) & 1]);
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[j & 1][c]);
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j * 16 + c4 * 4);
              v[c4 * 4 + 0] += b4.x;
              v[c4 * 4 + 1] += b4.y;
              v[c4 * 4 + 2] += b4.z;
              v[c4 * 4 + 3] += b4.w;
            }
            if (p.has_res) {
              const uint32_t ru[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float a, b;
                unpack2(ru[c], a, b);
                v[2 * c] += a;
                v[2 * c + 1] += b;
              }
            }
            if (want_stats && kRegStats) {
              if (valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  ps1[kRegStats ? j * 16 + c : 0] += v[c];
                  ps2[kRegStats ? j * 16 + c : 0] = fmaf(v[c], v[c], ps2[kRegStats ? j * 16 + c : 0]);
                }
              }
            } else if (want_stats) {
              // transpose through shared memory: rows = pixels of this warp, then per-lane column sums
              float4* srow = reinterpret_cast<float4*>(sbuf + lane * kStatPitch);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4)
                srow[c4] = valid ? make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3])
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
              __syncwarp();
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int row = (i >> 2) * 8 + st_half * 4 + (i & 3);
                const float xv = sbuf[row * kStatPitch + st_col];
                s1 += xv;
                s2 = fmaf(xv, xv, s2);
              }
              acc1[j] += s1;
              acc2[j] += s2;
              __syncwarp();
            }
            if (valid) {
              uint32_t o[8];
              if (p.act == ACT_LRELU) {  // uniform branch
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c], slope * v[c]);
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = pack2(v[2 * c], v[2 * c + 1]);
              const bool segb = p.seg_cols && j * 16 >= p.seg_cols;   // uniform
              act_t* ob = segb ? obase_b + (size_t)(2 * j - (p.seg_cols >> 3)) * HW8 : obase + (size_t)(2 * j) * oplane;
              if (ppc > 1) {   // chunk j = channels [16 jc, 16 jc + 16) of output parity par + q
                const int q = (j * 16) / bnp, jc = j - q * (bnp >> 4);
                ob = obase + (size_t)q * HW8 + (size_t)(2 * jc) * oplane;
              }
              *reinterpret_cast<uint4*>(ob) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(ob + (segb ? HW8 : oplane)) = make_uint4(o[4], o[5], o[6], o[7]);
              if (p.has_out2) {
                *reinterpret_cast<uint4*>(obase2 + (size_t)(2 * j) * HW8) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(obase2 + (size_t)(2 * j + 1) * HW8) = make_uint4(o[4], o[5], o[6], o[7]);
              }
            }
          }
        } else if (kSpade) {
          constexpr int NCS = CT / 16;
          const int tiles_per_q = p.C / CT;
          const int q0 = NQ == 1 ? ntile / tiles_per_q : 0;              // first output of this tile
          const int c0 = (ntile - q0 * tiles_per_q) * CT;
          uint32_t rg[16], rb[16];
          if (m + 1 < p.MT) load_x(m + 1, xnext);
          // chunked form: this pixel's x row, planes 2 j and 2 j + 1 are fetched one chunk ahead
          const size_t xHW8c = (size_t)p.Hx * p.Wx * 8;
          const act_t* xrowc = nullptr;
          constexpr int kXD = 4;   // chunks of x in flight (a ring of registers): L2 latency >> one chunk's arithmetic
          uint4 xring[kXChunked ? kXD : 1][2];
          if (kXChunked) {
            const int sy = p.ups ? (oy >> 1) : oy, sx = p.ups ? (ox >> 1) : ox;
            xrowc = p.x.p + (size_t)(valid ? n : 0) * p.x.bstride + (size_t)(c0 >> 3) * xHW8c +
                    ((size_t)(valid ? sy : 0) * p.Wx + (valid ? sx : 0)) * 8;
#pragma unroll
            for (int k = 0; k < kXD; ++k) {
              xring[kXChunked ? k : 0][0] = xring[kXChunked ? k : 0][1] = make_uint4(0, 0, 0, 0);
              if (valid && k < NCS) {
                xring[kXChunked ? k : 0][0] = *reinterpret_cast<const uint4*>(xrowc + (size_t)(2 * k) * xHW8c);
                xring[kXChunked ? k : 0][1] = *reinterpret_cast<const uint4*>(xrowc + (size_t)(2 * k + 1) * xHW8c);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < NCS; ++j) {
            uint4 x0, x1;
            if (kXChunked) {
              x0 = xring[kXChunked ? j % kXD : 0][0];
              x1 = xring[kXChunked ? j % kXD : 0][1];
              if (j + kXD < NCS && valid) {   // refill the slot with the chunk kXD ahead
                xring[kXChunked ? j % kXD : 0][0] = *reinterpret_cast<const uint4*>(xrowc + (size_t)(2 * (j + kXD)) * xHW8c);
                xring[kXChunked ? j % kXD : 0][1] = *reinterpret_cast<const uint4*>(xrowc + (size_t)(2 * (j + kXD) + 1) * xHW8c);
              }
            } else {
              x0 = xcur[kXChunked ? 0 : j][0];
              x1 = xcur[kXChunked ? 0 : j][1];
            }
            const uint32_t xu[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            float xn[16];   // (x - mean) * rstd, shared by the outputs of the tile
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int cb = j * 16 + c4 * 4;
              const float4 rs = *reinterpret_cast<const float4*>(s_auxg + cb);        // rstd
              const float4 ms = *reinterpret_cast<const float4*>(s_auxg + CT + cb);   // -mean * rstd
              float xa, xb, xc, xd;
              unpack2(xu[c4 * 2], xa, xb);
              unpack2(xu[c4 * 2 + 1], xc, xd);
              xn[c4 * 4 + 0] = fmaf(xa, rs.x, ms.x);
              xn[c4 * 4 + 1] = fmaf(xb, rs.y, ms.y);
              xn[c4 * 4 + 2] = fmaf(xc, rs.z, ms.z);
              xn[c4 * 4 + 3] = fmaf(xd, rs.w, ms.w);
            }
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
              const int qq = q0 + qi;
              const int colg = qi * 2 * CT + j * 16, colb = colg + CT;       // gamma / beta columns inside the tile
              float g[16], b[16];
              if (SIMT) {
                simt_chunk(p, n, oy, ox, ntile * BN + colg, g);
                simt_chunk(p, n, oy, ox, ntile * BN + colb, b);
              } else {
                tmem_ld16_issue(trow + (uint32_t)colg, rg);
                tmem_ld16_issue(trow + (uint32_t)colb, rb);
                tmem_ld16_wait(rg);
                tmem_ld16_wait(rb);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  g[c] = __uint_as_float(rg[c]);
                  b[c] = __uint_as_float(rb[c]);
                }
              }
              if (valid) {
                float y[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                  if (SIMT) {   // bring-up loop: the biases (gamma's includes the "+ 1") are added here
                    const float4 bg = *reinterpret_cast<const float4*>(s_bias + colg + c4 * 4);
                    const float4 bb = *reinterpret_cast<const float4*>(s_bias + colb + c4 * 4);
                    y[c4 * 4 + 0] = fmaf(xn[c4 * 4 + 0], g[c4 * 4 + 0] + bg.x, b[c4 * 4 + 0] + bb.x);
                    y[c4 * 4 + 1] = fmaf(xn[c4 * 4 + 1], g[c4 * 4 + 1] + bg.y, b[c4 * 4 + 1] + bb.y);
                    y[c4 * 4 + 2] = fmaf(xn[c4 * 4 + 2], g[c4 * 4 + 2] + bg.z, b[c4 * 4 + 2] + bb.z);
                    y[c4 * 4 + 3] = fmaf(xn[c4 * 4 + 3], g[c4 * 4 + 3] + bg.w, b[c4 * 4 + 3] + bb.w);
                  } else {      // tcgen05 path: 1 + gamma and beta arrive complete from the bias MMA
                    y[c4 * 4 + 0] = fmaf(xn[c4 * 4 + 0], g[c4 * 4 + 0], b[c4 * 4 + 0]);
                    y[c4 * 4 + 1] = fmaf(xn[c4 * 4 + 1], g[c4 * 4 + 1], b[c4 * 4 + 1]);
                    y[c4 * 4 + 2] = fmaf(xn[c4 * 4 + 2], g[c4 * 4 + 2], b[c4 * 4 + 2]);
                    y[c4 * 4 + 3] = fmaf(xn[c4 * 4 + 3], g[c4 * 4 + 3], b[c4 * 4 + 3]);
                  }
                }
                const float sl = p.actq[qq] == ACT_LRELU ? 0.2f : 1.0f;
                act_t* orow = p.outq[qq].p + (size_t)n * p.outq[qq].bstride + (size_t)(c0 >> 3) * HW8 + pix8;
                uint32_t o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = pack2(fmaxf(y[2 * c], sl * y[2 * c]), fmaxf(y[2 * c + 1], sl * y[2 * c + 1]));
                *reinterpret_cast<uint4*>(orow + (size_t)(2 * j) * HW8) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(orow + (size_t)(2 * j + 1) * HW8) = make_uint4(o[4], o[5], o[6], o[7]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < NXC; ++j) {
            xcur[j][0] = xnext[j][0];
            xcur[j][1] = xnext[j][1];
          }
        } else {  // EPI_FINAL (BN == 16, at most 4 real output channels)
          float v[16];
          if (SIMT) simt_chunk(p, n, oy, ox, 0, v);
          else tmem_ld16(trow, v);
          if (valid) {
            // (kept small on purpose: the 16-way unrolled form of this block thrashed the instruction cache)
            float y[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = SIMT ? v[c] + s_bias[c] : v[c];   // (tcgen05 path: bias MMA)
            if (p.act == ACT_TANH) {
#pragma unroll
              for (int c = 0; c < 4; ++c) y[c] = tanhf(y[c]);
            } else if (p.act == ACT_SIGMOID) {
#pragma unroll
              for (int c = 0; c < 4; ++c) y[c] = 1.f / (1.f + expf(-y[c]));
            } else if (p.act == ACT_LRELU) {
#pragma unroll
              for (int c = 0; c < 4; ++c) y[c] = lrelu02(y[c]);
            }
            float* of = p.out_f32 + ((size_t)n * p.n_valid * p.H + oy) * p.W + ox;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < p.n_valid) {
                of[(size_t)c * p.H * p.W] = y[c];
                if (p.has_out_act) {
                  const int cc = p.out_act_coff + c;
                  p.out_act.p[(size_t)n * p.out_act.bstride + (size_t)(cc >> 3) * HW8 + pix8 + (cc & 7)] = f2act(y[c]);
                }
              }
            }
          }
        }
      }
      if (!SIMT) {  // hand the accumulator buffer back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[buf]), 0u));   // the leader's barrier
          else mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
        }
      }
      for (int k = 0; k < kEpiGroups * tstride; ++k) advance_tile();
    }
    if (MODE == EPI_STORE && want_stats && cur_n >= 0) flush_stats(cur_n);
    tc_fence_before();
  }

  if (PAIR) cluster_sync_all();   // neither CTA leaves (or frees its tensor memory) while the peer may still touch it
  else __syncthreads();
  if (!SIMT && warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int make_tmap_act_s1(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int box_c,
                     int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  RIB_REQUIRE(C % 8 == 0 && box_c % 8 == 0 && box_w * 8 <= 256 && box_h <= 256, "bad activation box");
  RIB_REQUIRE(((uintptr_t)base & 15) == 0, "activation view must be 16-byte aligned");
  const cuuint64_t es = sizeof(act_t);
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 8 * es, (cuuint64_t)H * W * 8 * es, (cuuint64_t)bstride * es};
  cuuint32_t box[4] = {(cuuint32_t)(box_w * 8), (cuuint32_t)box_h, (cuuint32_t)(box_c / 8), 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 4, const_cast<act_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
  return 0;
}

int make_tmap_act_s2(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int py, int px,
                     int box_c, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  RIB_REQUIRE(C % 8 == 0 && box_c % 8 == 0 && W % 2 == 0 && H % 2 == 0, "bad stride-2 activation view");
  RIB_REQUIRE(box_w * 8 <= 256 && box_h <= 256, "bad stride-2 activation box");
  const cuuint64_t es = sizeof(act_t);
  const int Hh = H / 2, Wh = W / 2;
  const act_t* b0 = base + (size_t)(py * 2 + px) * Hh * Wh * 8;
  RIB_REQUIRE(((uintptr_t)b0 & 15) == 0, "activation view must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)Wh * 8, (cuuint64_t)Hh, (cuuint64_t)(C / 8), (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Wh * 8 * es, (cuuint64_t)H * W * 8 * es, (cuuint64_t)bstride * es};
  cuuint32_t box[4] = {(cuuint32_t)(box_w * 8), (cuuint32_t)box_h, (cuuint32_t)(box_c / 8), 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 4, const_cast<act_t*>(b0), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(stride-2 activation) failed: " + std::to_string((int)r));
  return 0;
}

// Parity view (py, px) of a NORMAL planar map (strided gather, 16-byte pieces): dims (8, W/2, H/2, C/8, B).  Used when
// the producer also has stride-1 consumers and a second, parity-planar copy of its output would cost more than it saves.
int make_tmap_act_s2_strided(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, long long bstride, int py,
                             int px, int box_c, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  RIB_REQUIRE(C % 8 == 0 && box_c % 8 == 0 && W % 2 == 0 && H % 2 == 0, "bad stride-2 activation view");
  const cuuint64_t es = sizeof(act_t);
  const act_t* b0 = base + ((size_t)py * W + px) * 8;
  cuuint64_t dims[5] = {8, (cuuint64_t)(W / 2), (cuuint64_t)(H / 2), (cuuint64_t)(C / 8), (cuuint64_t)B};
  cuuint64_t strides[4] = {16 * es, (cuuint64_t)2 * W * 8 * es, (cuuint64_t)H * W * 8 * es, (cuuint64_t)bstride * es};
  cuuint32_t box[5] = {8u, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)(box_c / 8), 1u};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 5, const_cast<act_t*>(b0), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(strided stride-2 activation) failed: " + std::to_string((int)r));
  return 0;
}

int make_tmap_w(CUtensorMap* m, const act_t* w, int K, int N, int bkc, int boxN, int taps) {
  (void)taps;
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bkc, (cuuint32_t)boxN};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 2, const_cast<act_t*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(bkc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
  return 0;
}

static constexpr size_t kStatStageBytes = (size_t)4 * kEpiGroups * kStatWarpFloats * 4;
static constexpr size_t kSmallResident = (size_t)72 * 1024;   // weights of an N tile that leave room for several CTAs per SM
static constexpr size_t kBigResident = (size_t)148 * 1024;    // ... that still fit beside two halo-tile slots (one CTA per SM)

// Channels per group.  Also fixes the K order of the packed weights, so it may only depend on the layer shape.
// RIB_STREAM_BKC: channels per ring slot of a stride-1 layer with streamed weights (32: two ~95 KB slots; 16: four ~48 KB
// slots, i.e. finer-grained refills of the same shared memory).
#ifndef RIB_STREAM_BKC
#define RIB_STREAM_BKC 32
#endif
int choose_bkc(int cin0, int cin1, int taps, int BN, int stride) {
  const size_t b_all = ((size_t)cin0 * taps + cin1) * BN * 2;
  int cap;
  if (BN == 256) cap = 64;   // 256-column pair tiles (1x1 SPADE layers): each CTA keeps its 128 weight rows resident
  else if (b_all <= kSmallResident) cap = stride == 2 ? 16 : 64;       // stride 2 keeps four parity tiles per slot
  else if (b_all <= kBigResident) cap = stride == 2 ? 16 : 64;    // big resident weights: small halo slots
  else cap = stride == 2 ? 16 : RIB_STREAM_BKC;                   // streamed: a slot holds the halo tile(s) AND 9 weight sub-tiles
  int bk = cin0 < cap ? cin0 : cap;
  if (cin1 > 0 && cin1 < bk) bk = cin1;
  return bk;
}

int conv_gemm_configure(ConvGemmParams* p, int B, int Hout, int Wout, int cin0, int cin1, int taps, int stride,
                        int BN, int n_pad, const ConvTune* tune, int ppc) {
  RIB_REQUIRE(taps == 1 || taps == 9 || taps == 4, "conv_gemm: 1x1, 3x3 or sub-pixel 2x2 only");
  RIB_REQUIRE(ppc == 1 || ppc == 2 || ppc == 4, "conv_gemm: parities per CTA must be 1, 2 or 4");
  RIB_REQUIRE(taps == 4 || ppc == 1, "conv_gemm: parities per CTA only apply to the sub-pixel conv");
  RIB_REQUIRE(taps != 4 || (stride == 1 && cin1 == 0 && ((n_pad / BN) * ppc) % 4 == 0 && (BN / ppc) % 16 == 0 &&
                            (ppc == 1 || n_pad == 4 * (BN / ppc))),
              "conv_gemm: bad sub-pixel conv");
  RIB_REQUIRE(stride == 1 || (stride == 2 && taps == 9 && cin1 == 0), "conv_gemm: stride 2 needs a plain 3x3");
  RIB_REQUIRE(BN >= 16 && BN <= 256 && (BN & (BN - 1)) == 0 && n_pad % BN == 0, "conv_gemm: bad BN");
  RIB_REQUIRE(BN != 256 || (taps == 1 && cin1 == 0 && stride == 1), "conv_gemm: 256-column tiles are for 1x1 layers (CTA pairs)");
  const int bkc = choose_bkc(cin0, cin1, taps, BN, stride);
  RIB_REQUIRE((bkc == 16 || bkc == 32 || bkc == 64) && cin0 % bkc == 0 && cin1 % bkc == 0,
              "conv_gemm: channel counts must be 16, 32 or multiples of 64");
  p->B = B;
  p->H = Hout;
  p->W = Wout;
  p->BN = BN;
  p->n_tiles = n_pad / BN;
  p->BKc = bkc;
  p->stages0 = cin0 / bkc;
  p->stages1 = cin1 / bkc;
  p->ntaps = taps;
  p->stride = stride;
  p->halo = taps == 1 ? 0 : 1;
  p->subpix = taps == 4 ? 1 : 0;
  p->ppc = ppc;
  p->b_tap_bytes = (uint32_t)(BN * bkc * 2);
  const int n_bt = p->stages0 * taps + p->stages1;
  const size_t b_all = (size_t)n_bt * p->b_tap_bytes;
  const int G = p->stages0 + p->stages1;
  const int ntile_a = stride == 2 ? 4 : 1;
  // Sub-tile shape.  A 3x3 tile is 16 rows x 8 pixels (every tile row is one 8-row UMMA group, the halo row pitch is
  // the group stride).  A 1x1 tile has no halo, so its 8-pixel groups may also sit side by side: 4 rows x 32 pixels
  // quarters the number of rows TMA has to fetch per tile (512-byte instead of 128-byte rows) and makes the
  // epilogue's x loads / stores 512-byte contiguous per warp.
  p->tw = (taps == 1 && Wout % 32 == 0) ? 32 : kTileW;
  p->th = 128 / p->tw;
  // geometry of one halo-tile slot for a given number of stacked sub-tiles
  auto set_geometry = [&](int MT) {
    p->MT = MT;
    const int hw = stride == 1 ? p->tw + 2 * p->halo : p->tw + 1;
    const int hh = stride == 1 ? p->th * MT + 2 * p->halo : p->th * MT + 1;
    p->halo_w = hw;
    p->lbo = (uint32_t)(hw * hh * 16);
    // stride between 8-pixel groups: the next tile row (3x3), or the next 8 pixels of the same row (32-wide 1x1 tile,
    // whose rows are contiguous in shared memory)
    p->sbo = p->tw == kTileW ? (uint32_t)(hw * 16) : 128u;
    const uint32_t tile_raw = (uint32_t)(hw * hh * 16 * (bkc / 8));
    p->a_tile_bytes = (tile_raw + 127u) & ~127u;
    p->a_slot_bytes = ((uint32_t)ntile_a * p->a_tile_bytes + 127u) & ~127u;
    p->a_tx_bytes = (uint32_t)ntile_a * tile_raw;
    p->tiles_x = ceil_div(Wout, p->tw);
    p->tiles_y = ceil_div(Hout, p->th * MT);
    p->b_off = 0;
    p->g_slot_bytes = p->a_slot_bytes;
  };
  const size_t kSmemMax = (size_t)227 * 1024, kOverhead = 12 * 1024;  // barriers, bias, statistics slots, alignment
  const bool can_mt2 = Hout >= 2 * p->th;
  p->b_ring = 1;  // unused
  // Residency policy: 1 = weights resident, CTA kept near 100 KB so that several fit on an SM (bandwidth-bound layers);
  // 2 = weights resident, one CTA per SM with the rest of shared memory as halo ring; 3 = weights streamed with the
  // halo tiles.  Defaults by weight size; a ConvTune (the plan-time auto-tuner, generator.cu) may override policy and MT.
  int policy = b_all <= kSmallResident ? 1 : (b_all <= kBigResident ? 2 : 3);
  if (tune != nullptr && tune->policy != 0) policy = tune->policy;
  if (tune == nullptr) {   // kernel tests select a policy for the stand-alone conv (read per call, never set by the product)
    const char* force = getenv("RIB_TEST_POLICY");
    if (force != nullptr && atoi(force) >= 1 && atoi(force) <= 5) policy = atoi(force);
  }
  if (BN == 256) policy = 5;   // only exists as a CTA pair with resident half-tiles (two accumulators of 256 columns)
  RIB_REQUIRE(policy >= 1 && policy <= 5 &&
                  (policy == 5 ? (b_all > kSmallResident && b_all / 2 <= kBigResident)
                               : (policy >= 3 ? b_all > kSmallResident : b_all <= kBigResident)),
              "conv_gemm: residency policy does not apply to this layer");
  p->pair = policy >= 4 ? 1 : 0;
  if (policy >= 4) {
    RIB_REQUIRE((BN == 128 || BN == 256) && taps != 4 && (cin1 == 0 || stride == 1), "conv_gemm: CTA pairs need a plain layer with BN = 128 / 256");
    p->b_tap_bytes = (uint32_t)((BN / 2) * bkc * 2);   // each CTA of the pair stages half of the weight rows
  }
  int mt = (policy == 1 || BN == 256) ? 1 : (can_mt2 ? 2 : 1);
  const bool mt_forced = tune != nullptr && tune->mt != 0;
  if (mt_forced) mt = tune->mt;
  RIB_REQUIRE(mt == 1 || (mt == 2 && can_mt2 && BN != 256), "conv_gemm: cannot stack two sub-tiles here");
  if (policy == 1) {
    p->b_resident = 1;
    set_geometry(mt);
    const size_t budget = (size_t)100 * 1024 - b_all - kStatStageBytes;
    int ring = (int)(budget / p->a_slot_bytes);
    const int want = G <= 2 ? 4 : 2 * G;
    if (ring > want) ring = want;
    if (ring > 8) ring = 8;
    if (ring < 2) ring = 2;
    p->a_ring = ring;
    RIB_REQUIRE(b_all + kStatStageBytes + kOverhead + (size_t)ring * p->a_slot_bytes <= kSmemMax,
                "conv_gemm: halo ring does not fit beside the resident weights");
  } else if (policy == 2 || policy == 5) {
    // (policy 5: CTA pairs with RESIDENT weights - each CTA keeps its half of the N tile's rows, so layers whose whole
    //  weight tile leaves room for two or three halo slots get a ring of up to eight)
    p->b_resident = 1;
    const size_t fixed = (policy == 5 ? b_all / 2 : b_all) + kStatStageBytes + kOverhead;
    set_geometry(mt);
    if (!mt_forced && fixed + 2 * (size_t)p->a_slot_bytes > kSmemMax) set_geometry(1);
    RIB_REQUIRE(fixed + 2 * (size_t)p->a_slot_bytes <= kSmemMax, "conv_gemm: resident weights do not fit");
    int ring = (int)((kSmemMax - fixed) / p->a_slot_bytes);
    const int ring_cap = (tune != nullptr && tune->ring >= 2 && tune->ring <= 8) ? tune->ring : (policy == 5 ? 8 : 4);
    p->a_ring = ring > ring_cap ? ring_cap : ring;
  } else {
    // a ring slot holds the halo tile(s) of a channel group AND that group's weight sub-tiles (one barrier round trip
    // per group); two stacked sub-tiles halve the weight traffic per pixel (measured: also for stride 2, whose four
    // parity tiles otherwise make the layer L2-bound).  CTA pairs (policy 4): same slots with half-height weight tiles.
    p->b_resident = 0;
    const size_t budget = kSmemMax - kStatStageBytes - kOverhead;
    for (;;) {
      set_geometry(mt);
      p->b_off = (p->a_slot_bytes + 1023u) & ~1023u;
      p->g_slot_bytes = p->b_off + (uint32_t)taps * p->b_tap_bytes;
      if (budget / p->g_slot_bytes >= 2 || mt == 1 || mt_forced) break;
      mt = 1;
    }
    const int ring = (int)(budget / p->g_slot_bytes);
    RIB_REQUIRE(ring >= 2, "conv_gemm: streamed group slots do not fit");
    const int ring_cap = (tune != nullptr && tune->ring >= 2 && tune->ring <= 8) ? tune->ring : 4;
    p->a_ring = ring > ring_cap ? ring_cap : ring;
  }
  p->idesc = make_idesc_f16(p->pair ? 256 : 128, BN / ppc);
  return 0;
}

size_t conv_gemm_smem_bytes(const ConvGemmParams& p) {
  const int n_bt = p.stages0 * p.ntaps + p.stages1;
  size_t tiles = (((size_t)p.a_ring * p.g_slot_bytes + 1023) & ~(size_t)1023) +
                 (size_t)(p.b_resident ? n_bt : 0) * p.b_tap_bytes;
  size_t stat = p.stats != nullptr ? kStatStageBytes : 0;
  const bool xf = p.xf_stats != nullptr || p.a_gather;
  size_t bars = (size_t)((xf ? 3 : 2) * p.a_ring + 2 * acc_bufs(p.BN) + 1) * 8 + 32;
  size_t scratch = (size_t)p.BN * 4 * (1 + 8 * kEpiGroups) + 64 + 128 + (xf ? (size_t)2 * p.stages0 * p.BKc * 4 : 0) +
                   128 + 256 + (size_t)p.BN * 32;   // ones tile + bias tile of the bias MMA (128-byte aligned)
  return 1024 + tiles + stat + bars + scratch;
}

static std::atomic<long long> g_launches{0};
long long conv_gemm_launch_count() { return g_launches.load(); }

// Optional per-launch CUDA-event timing of the implicit-GEMM kernel (bench.py's roofline pass).
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static std::mutex g_prof_mutex;
void conv_gemm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  g_profile = on != 0;
}
int conv_gemm_profile_collect(double* total_ms, long long* launches, float* per_launch_ms, long long cap) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  double ms = 0.0;
  long long i = 0;
  for (auto& ev : g_prof_events) {
    RIB_CHECK_CUDA(cudaEventSynchronize(ev.second));
    float t = 0.f;
    RIB_CHECK_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
    ms += t;
    if (per_launch_ms != nullptr && i < cap) per_launch_ms[i] = t;
    ++i;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *total_ms = ms;
  *launches = (long long)g_prof_events.size();
  g_prof_events.clear();
  return 0;
}

// Resident CTAs per SM for a given dynamic shared-memory size: limited by shared memory (227 KB usable,
// 1 KB reserved per CTA), registers (64 K per SM) and TMEM columns (512 per SM).
static int ctas_per_sm(const void* fn, size_t smem, int tmem_cols, int threads, int* occ) {
  cudaFuncAttributes fa;
  RIB_CHECK_CUDA(cudaFuncGetAttributes(&fa, fn));
  const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * threads;
  int o = (int)((size_t)227 * 1024 / (smem + 1024));
  if (regs_per_cta > 0 && o > 65536 / regs_per_cta) o = 65536 / regs_per_cta;
  if (o > 512 / tmem_cols) o = 512 / tmem_cols;
  if (o > 4) o = 4;
  if (o < 1) o = 1;
  *occ = o;
  return 0;
}

typedef void (*ConvKernel)(const ConvGemmParams);

// SM count of a device (queried once per device; the persistent grids are sized from it).
int device_sm_count(int dev) {
  static std::mutex mu;
  static std::vector<int> cache;
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0) return 0;
  if ((int)cache.size() <= dev) cache.resize(dev + 1, 0);
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache[dev] = n;
  }
  return cache[dev];
}

template <bool SIMT>
static ConvKernel pick_kernel(int mode, int BN, bool xf, bool pair) {
  if (pair) {   // CTA pairs: 128-column tiles of the heavy 3x3 layers and of the low-resolution SPADE layers
    if (SIMT || xf) return nullptr;
    if (BN == 256) {   // 256-column tiles of the low-resolution SPADE layers: [gamma | beta] of 128 (SPADE2: 64 + 64) channels
      if (mode == EPI_SPADE) return conv_gemm_kernel<EPI_SPADE, 256, false, false, true>;
      if (mode == EPI_SPADE2) return conv_gemm_kernel<EPI_SPADE2, 256, false, false, true>;
      return nullptr;
    }
    if (BN != 128) return nullptr;
    if (mode == EPI_STORE) return conv_gemm_kernel<EPI_STORE, 128, false, false, true>;
    if (mode == EPI_SPADE) return conv_gemm_kernel<EPI_SPADE, 128, false, false, true>;
    if (mode == EPI_SPADE2) return conv_gemm_kernel<EPI_SPADE2, 128, false, false, true>;
    return nullptr;
  }
  if (xf) {  // the A-operand transform is built for the layers that use it: plain-store convs with 64 / 128 columns
    if (mode != EPI_STORE) return nullptr;
    switch (BN) {
      case 64: return conv_gemm_kernel<EPI_STORE, 64, SIMT, true, false>;
      case 128: return conv_gemm_kernel<EPI_STORE, 128, SIMT, true, false>;
    }
    return nullptr;
  }
  if (mode == EPI_STORE) {
    switch (BN) {
      case 16: return conv_gemm_kernel<EPI_STORE, 16, SIMT, false, false>;
      case 32: return conv_gemm_kernel<EPI_STORE, 32, SIMT, false, false>;
      case 64: return conv_gemm_kernel<EPI_STORE, 64, SIMT, false, false>;
      case 128: return conv_gemm_kernel<EPI_STORE, 128, SIMT, false, false>;
    }
  } else if (mode == EPI_SPADE) {
    switch (BN) {
      case 32: return conv_gemm_kernel<EPI_SPADE, 32, SIMT, false, false>;
      case 64: return conv_gemm_kernel<EPI_SPADE, 64, SIMT, false, false>;
      case 128: return conv_gemm_kernel<EPI_SPADE, 128, SIMT, false, false>;
      case 256: if (SIMT) return conv_gemm_kernel<EPI_SPADE, 256, true, false, false>; break;   // (tcgen05: pairs only)
    }
  } else if (mode == EPI_SPADE2) {
    switch (BN) {
      case 64: return conv_gemm_kernel<EPI_SPADE2, 64, SIMT, false, false>;
      case 128: return conv_gemm_kernel<EPI_SPADE2, 128, SIMT, false, false>;
      case 256: if (SIMT) return conv_gemm_kernel<EPI_SPADE2, 256, true, false, false>; break;
    }
  } else if (mode == EPI_FINAL && BN == 16) {
    return conv_gemm_kernel<EPI_FINAL, 16, SIMT, false, false>;
  }
  return nullptr;
}

int launch_conv_gemm(const ConvGemmParams& p, int mode, cudaStream_t stream) {
  RIB_REQUIRE(p.BKc == 16 || p.BKc == 32 || p.BKc == 64, "conv_gemm: BKc must be 16/32/64");
  RIB_REQUIRE(p.a_ring >= 2 && p.a_ring <= 8, "conv_gemm: bad ring depth");
  RIB_REQUIRE(p.MT == 1 || p.MT == 2, "conv_gemm: MT must be 1 or 2");
  RIB_REQUIRE(acc_bufs(p.BN) * p.MT * p.BN <= 512, "conv_gemm: accumulators exceed TMEM");
  RIB_REQUIRE(p.n_tiles >= 1, "conv_gemm: no N tiles");
  RIB_REQUIRE(mode != EPI_FINAL || p.n_valid <= 4, "conv_gemm: EPI_FINAL writes at most 4 channels");
  RIB_REQUIRE(p.seg_cols == 0 || (mode == EPI_STORE && p.n_tiles == 1 && p.seg_cols % 16 == 0 && p.n_valid % 16 == 0 &&
                                  p.seg_cols < p.n_valid && !p.has_res && !p.has_out2 && !p.subpix && !p.out_parity &&
                                  p.BN > RIB_REGSTATS_MAXBN),
              "conv_gemm: bad merged-conv launch");
  RIB_REQUIRE(mode != EPI_SPADE || (p.BN == 2 * p.CT && p.CT % 16 == 0 && p.C % p.CT == 0),
              "conv_gemm: EPI_SPADE needs BN == 2*CT");
  RIB_REQUIRE(mode != EPI_SPADE2 || (p.BN == 4 * p.CT && p.CT % 16 == 0 && p.C % p.CT == 0 && p.n_tiles == p.C / p.CT),
              "conv_gemm: EPI_SPADE2 needs BN == 4*CT and one tile per CT channels");
  const bool xf = p.xf_stats != nullptr || p.a_gather;
  RIB_REQUIRE(!xf || (p.stages1 == 0 && p.subpix == 0), "conv_gemm: the A-operand transform needs a single source");
  RIB_REQUIRE(!p.a_gather || (p.xf_stats == nullptr && p.stride == 2 && !p.s2_parity && p.b_resident && !p.pair &&
                              4 * (p.BKc / 8) * (int)(p.lbo >> 4) <= 12 * kXfThreads),
              "conv_gemm: the stride-2 gather needs a normal-layout input, resident weights and at most 12 vectors per thread");
  RIB_REQUIRE(!p.out_parity || (p.H % 2 == 0 && p.W % 2 == 0 && !p.subpix), "conv_gemm: bad parity-planar output");
  RIB_REQUIRE(!p.res_ups || (p.has_res && p.H % 2 == 0 && p.W % 2 == 0 && !p.subpix && !p.out_parity),
              "conv_gemm: bad up-sampled residual");
  // (the bring-up FMA loop has no pair form: a paired layer falls back to single CTAs with the same shared-memory plan)
  const bool pair = p.pair != 0 && !p.debug_simt;
  RIB_REQUIRE(!p.pair || ((p.stride == 1 || p.stages1 == 0) && !p.subpix && !xf),
              "conv_gemm: CTA pairs: no sub-pixel form, no transform, stride 2 only with a single source");
  ConvKernel fn = p.debug_simt ? pick_kernel<true>(mode, p.BN, xf, false) : pick_kernel<false>(mode, p.BN, xf, pair);
  RIB_REQUIRE(fn != nullptr, "conv_gemm: no kernel for this (epilogue, BN)");
  const size_t smem = conv_gemm_smem_bytes(p);
  RIB_REQUIRE(smem <= 227 * 1024, "conv_gemm: shared memory budget exceeded");
  int dev = 0;
  RIB_CHECK_CUDA(cudaGetDevice(&dev));
  {
    // the opt-in is per device (context), not per process: key the cache by (device, kernel)
    static std::mutex mu;
    static std::vector<std::pair<int, const void*>> done;
    std::lock_guard<std::mutex> lk(mu);
    bool seen = false;
    for (const auto& f : done) seen = seen || (f.first == dev && f.second == (const void*)fn);
    if (!seen) {
      RIB_CHECK_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      done.push_back({dev, (const void*)fn});
    }
  }
  // persistent grid: as many CTAs as fit on the chip, a multiple of n_tiles
  int tmem_cols = 32;
  while (tmem_cols < acc_bufs(p.BN) * p.MT * p.BN) tmem_cols <<= 1;
  int occ = 1;
  const int threads = kThreads + (xf ? kXfThreads : 0);
  int rc = ctas_per_sm((const void*)fn, smem, tmem_cols, threads, &occ);
  if (rc) return rc;
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.B;
  const int n_sms = device_sm_count(dev);
  RIB_REQUIRE(n_sms > 0, "conv_gemm: cannot query the SM count");
  // pair: `groups` counts CTA pairs per N tile, each pair walks pairs of super-tiles
  long long groups = pair ? ((long long)(n_sms / 2) * occ) / p.n_tiles : ((long long)n_sms * occ) / p.n_tiles;
  if (groups < 1) groups = 1;
  if (groups > (pair ? (m_tiles + 1) / 2 : m_tiles)) groups = pair ? (m_tiles + 1) / 2 : m_tiles;
  dim3 grid((unsigned)(groups * p.n_tiles * (pair ? 2 : 1)));
  dim3 block(threads);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_profile) {
    RIB_CHECK_CUDA(cudaEventCreate(&ev0));
    RIB_CHECK_CUDA(cudaEventCreate(&ev1));
    RIB_CHECK_CUDA(cudaEventRecord(ev0, stream));
  }
  {
    // RIB_PDL=0: plain stream order between the conv launches (A/B runs)
    static const bool pdl_on = !(getenv("RIB_PDL") != nullptr && atoi(getenv("RIB_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pair) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl_on && !p.debug_simt) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    RIB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, p));
  }
  RIB_CHECK_CUDA(cudaGetLastError());
  if (g_profile) {
    RIB_CHECK_CUDA(cudaEventRecord(ev1, stream));
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    g_prof_events.push_back({ev0, ev1});
  }
  g_launches.fetch_add(1);
  return 0;
}

}  // namespace rib
