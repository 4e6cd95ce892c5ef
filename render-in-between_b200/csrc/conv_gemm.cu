// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).  See conv_gemm.cuh.
//
// CTA = 192 threads, persistent over super-tiles: warp 0 = TMA producer (one lane), warp 1 = TMEM
// allocator + tcgen05.mma issuer (one lane), warps 2..5 = epilogue (TMEM lane quarter = warp_idx & 3).
// Barriers: full/empty per ring slot (TMA <-> MMA), tmem_full/tmem_empty per accumulator buffer
// (MMA <-> epilogue), one barrier for the resident weights.
#include "conv_gemm.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace rib {

static constexpr int kThreads = 192;
static constexpr int kNumSms = 148;

// Geometry of tap t of a stage: which halo tile of the slot it reads and the pixel offset of its
// top-left corner inside that tile.
struct TapGeom {
  int tile, poff;
};
__device__ __forceinline__ TapGeom tap_geom(const ConvGemmParams& p, bool src1, int t) {
  TapGeom g;
  g.tile = 0;
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
__device__ void simt_chunk(const ConvGemmParams& p, int n, int oy, int ox, int col0, float* acc) {
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  const int cin0 = p.stages0 * p.BKc, cin1 = p.stages1 * p.BKc;
  for (int ci = 0; ci < cin0; ++ci) {
    const int grp = ci / p.BKc, cc = ci - grp * p.BKc;
    for (int t = 0; t < p.ntaps; ++t) {
      const int r = p.ntaps == 9 ? t / 3 : 1, s = p.ntaps == 9 ? t - 3 * (t / 3) : 1;
      const int iy = oy * p.stride + r - 1, ix = ox * p.stride + s - 1;
      if (iy < 0 || ix < 0 || iy >= p.Hin || ix >= p.Win) continue;
      const float av = act2f(*planar_at(p.src0, n, ci, p.Hin, p.Win, iy, ix));
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

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int S = p.stages0 + p.stages1;  // pipeline stages per tile
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.ring * p.a_slot_bytes;
  const size_t b_bytes = (size_t)(p.b_resident ? S : p.ring) * p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + b_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.ring;
  uint64_t* tmem_full_bar = bars + 2 * p.ring;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * p.ring + 2;  // [2]
  uint64_t* bres_bar = bars + 2 * p.ring + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * p.ring + 5);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);  // [BN]
  float* s_aux = s_bias + p.BN;                            // STORE: [2*BN] stats; SPADE: [2*CT] mean, rstd

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntile = blockIdx.x % p.n_tiles;
  const int cta_m = blockIdx.x / p.n_tiles;
  const int cta_m_stride = gridDim.x / p.n_tiles;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_img * p.B;
  const int acc_cols = p.MT * p.BN;  // TMEM columns of one accumulator buffer
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;
  const bool tc = !p.debug_simt;

  if (warp == 0 && lane == 0 && tc) {
    prefetch_tmap(&p.amap[0]);
    prefetch_tmap(&p.bmap);
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);
    }
    mbar_init(smem_u32(bres_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1 && tc) {
    tmem_alloc(smem_u32(tmem_ptr), tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int e = threadIdx.x - 64;
    for (int c = e; c < p.BN; c += 128) s_bias[c] = p.bias[ntile * p.BN + c];
    if (MODE == EPI_STORE) {
      for (int c = e; c < 2 * p.BN; c += 128) s_aux[c] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tc ? *tmem_ptr : 0u;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && tc) {
      if (p.b_resident) {  // all weight stages of this N tile, once
        const uint32_t bb = smem_u32(bres_bar);
        mbar_arrive_expect_tx(bb, (uint32_t)((p.stages0 * p.ntaps + p.stages1) * p.b_tap_bytes));
        for (int s = 0; s < S; ++s) {
          const int nt = s < p.stages0 ? p.ntaps : 1;
          const int k0 = s < p.stages0 ? s * p.ntaps * p.BKc : (p.stages0 * p.ntaps + (s - p.stages0)) * p.BKc;
          for (int t = 0; t < nt; ++t)
            tma_load_2d(smem_u32(sB + (size_t)s * p.b_stage_bytes + (size_t)t * p.b_tap_bytes), &p.bmap, bb,
                        k0 + t * p.BKc, ntile * p.BN);
        }
      }
      int slot = 0;
      uint32_t phase = 0;
      for (int mt = cta_m; mt < total_tiles; mt += cta_m_stride) {
        const int n = mt / tiles_per_img;
        const int rem = mt - n * tiles_per_img;
        const int tile_y = rem / p.tiles_x, tile_x = rem - tile_y * p.tiles_x;
        const int oy0 = tile_y * kTileH * p.MT, ox0 = tile_x * kTileW;
        for (int s = 0; s < S; ++s) {
          mbar_wait(smem_u32(&empty_bar[slot]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[slot]);
          const bool src1 = s >= p.stages0;
          const int nt = src1 ? 1 : p.ntaps;
          mbar_arrive_expect_tx(fb, p.a_tx_bytes + (p.b_resident ? 0u : (uint32_t)nt * p.b_tap_bytes));
          const uint32_t a_dst = smem_u32(sA + (size_t)slot * p.a_slot_bytes);
          const int cg = (src1 ? s - p.stages0 : s) * (p.BKc >> 3);  // first 8-channel plane of the group
          if (p.stride == 1) {
            tma_load_4d(a_dst, &p.amap[src1 ? 1 : 0], fb, (ox0 - p.halo) * 8, oy0 - p.halo, cg, n);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_5d(a_dst + q * p.a_tile_bytes, &p.amap[q], fb, 0, ox0 - 1, oy0 - 1, cg, n);
          }
          if (!p.b_resident) {
            const int k0 = src1 ? (p.stages0 * p.ntaps + (s - p.stages0)) * p.BKc : s * p.ntaps * p.BKc;
            for (int t = 0; t < nt; ++t)
              tma_load_2d(smem_u32(sB + (size_t)slot * p.b_stage_bytes + (size_t)t * p.b_tap_bytes), &p.bmap, fb,
                          k0 + t * p.BKc, ntile * p.BN);
          }
          if (++slot == p.ring) {
            slot = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && tc) {
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t b_row_bytes = (uint32_t)p.BKc * 2u;
      const int kk_steps = p.BKc >> 4;
      if (p.b_resident) {
        mbar_wait(smem_u32(bres_bar), 0u);
        tc_fence_after();
      }
      int it = 0;
      for (int mt = cta_m; mt < total_tiles; mt += cta_m_stride, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&tmem_empty_bar[buf]), (use & 1u) ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * acc_cols);
        for (int s = 0; s < S; ++s) {
          mbar_wait(smem_u32(&full_bar[slot]), phase);
          tc_fence_after();
          const bool src1 = s >= p.stages0;
          const int nt = src1 ? 1 : p.ntaps;
          const uint32_t a_base = smem_u32(sA + (size_t)slot * p.a_slot_bytes);
          const uint32_t b_base = smem_u32(sB + (size_t)(p.b_resident ? s : slot) * p.b_stage_bytes);
          for (int t = 0; t < nt; ++t) {
            const TapGeom g = tap_geom(p, src1, t);
            const uint64_t bdesc = make_kmajor_desc(b_base + (uint32_t)t * p.b_tap_bytes, b_row_bytes);
            for (int m = 0; m < p.MT; ++m) {
              const uint32_t a_addr =
                  a_base + (uint32_t)g.tile * p.a_tile_bytes + (uint32_t)(g.poff + m * kTileH * p.halo_w) * 16u;
              for (int kk = 0; kk < kk_steps; ++kk) {
                const uint64_t adesc = make_nosw_desc(a_addr + (uint32_t)(2 * kk) * p.lbo, p.lbo, p.sbo);
                umma_f16(d0 + (uint32_t)(m * p.BN), adesc, bdesc + (uint64_t)(2 * kk), p.idesc,
                         (uint32_t)((s | t | kk) != 0));
              }
            }
          }
          umma_commit(smem_u32(&empty_bar[slot]));  // frees this ring slot once the MMAs have read it
          if (++slot == p.ring) {
            slot = 0;
            phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&tmem_full_bar[buf]));
      }
    }
  } else {
    // ===================== Epilogue =====================
    const int q = warp & 3;
    const int e = threadIdx.x - 64;
    const int prow = q * 32 + lane;             // row of the M=128 sub-tile = TMEM lane
    const int ty = prow >> 3, tx = prow & 7;    // 16 x 8 pixel tile
    const size_t HW8 = (size_t)p.H * p.W * 8;
    int cur_n = -1;
    int it = 0;
    for (int mt = cta_m; mt < total_tiles; mt += cta_m_stride, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int tile_y = rem / p.tiles_x, tile_x = rem - tile_y * p.tiles_x;
      const int oy0 = tile_y * kTileH * p.MT, ox0 = tile_x * kTileW;

      if (n != cur_n) {  // uniform over the 128 epilogue threads
        if (MODE == EPI_STORE && p.stats != nullptr && cur_n >= 0) {
          epi_bar();
          for (int c = e; c < p.BN; c += 128) {
            const int col = ntile * p.BN + c;
            if (col < p.n_valid) {
              atomicAdd(&p.stats[((size_t)cur_n * p.n_valid + col) * 2 + 0], (double)s_aux[c]);
              atomicAdd(&p.stats[((size_t)cur_n * p.n_valid + col) * 2 + 1], (double)s_aux[p.BN + c]);
            }
            s_aux[c] = 0.f;
            s_aux[p.BN + c] = 0.f;
          }
          epi_bar();
        }
        if (MODE == EPI_SPADE) {
          const int tiles_per_q = p.C / p.CT;
          const int c0 = (ntile % tiles_per_q) * p.CT;
          const double cnt = (double)p.Hx * (double)p.Wx;
          epi_bar();
          for (int c = e; c < p.CT; c += 128) {
            const double s = p.xstats[((size_t)n * p.C + c0 + c) * 2 + 0];
            const double ss = p.xstats[((size_t)n * p.C + c0 + c) * 2 + 1];
            const double mean = s / cnt;
            double var = ss / cnt - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            s_aux[c] = (float)mean;
            s_aux[p.CT + c] = (float)(1.0 / sqrt(var + (double)p.eps));
          }
          epi_bar();
        }
        cur_n = n;
      }

      if (tc) {
        mbar_wait(smem_u32(&tmem_full_bar[buf]), use & 1u);
        tc_fence_after();
      }
      for (int m = 0; m < p.MT; ++m) {
        const int oy = oy0 + m * kTileH + ty, ox = ox0 + tx;
        const bool valid = (oy < p.H) && (ox < p.W);
        const size_t pix8 = ((size_t)oy * p.W + ox) * 8;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + m * p.BN);

        if (MODE == EPI_STORE) {
          const int nchunks = p.BN / 16;
          for (int j = 0; j < nchunks; ++j) {
            float v[16];
            const int col0 = ntile * p.BN + j * 16;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (p.has_res && valid) {
              const act_t* rp = p.res.p + (size_t)n * p.res.bstride + (size_t)(col0 >> 3) * HW8 + pix8;
              r0 = *reinterpret_cast<const uint4*>(rp);
              r1 = *reinterpret_cast<const uint4*>(rp + HW8);
            }
            if (!tc) simt_chunk(p, n, oy, ox, col0, v);
            else tmem_ld16(trow + (uint32_t)(j * 16), v);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] += s_bias[j * 16 + c];
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
            if (p.stats != nullptr) {
              float s1[16], s2[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                s1[c] = valid ? v[c] : 0.f;
                s2[c] = valid ? v[c] * v[c] : 0.f;
              }
              const float t1 = warp_colsum16(s1, lane);
              const float t2 = warp_colsum16(s2, lane);
              if ((lane & 1) == 0) {
                const int c = (lane >> 1) & 15;
                atomicAdd(&s_aux[j * 16 + c], t1);
                atomicAdd(&s_aux[p.BN + j * 16 + c], t2);
              }
            }
            if (valid && col0 < p.n_valid) {
              uint32_t o[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) o[c] = pack2(apply_act(v[2 * c], p.act), apply_act(v[2 * c + 1], p.act));
              act_t* op = p.out.p + (size_t)n * p.out.bstride + (size_t)(col0 >> 3) * HW8 + pix8;
              *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(op + HW8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        } else if (MODE == EPI_SPADE) {
          const int tiles_per_q = p.C / p.CT;
          const int qq = ntile / tiles_per_q;
          const int c0 = (ntile - qq * tiles_per_q) * p.CT;
          const int sy = p.ups ? (oy >> 1) : oy, sx = p.ups ? (ox >> 1) : ox;
          const size_t xHW8 = (size_t)p.Hx * p.Wx * 8;
          const act_t* xrow = p.x.p + (size_t)n * p.x.bstride + (size_t)(c0 >> 3) * xHW8 + ((size_t)sy * p.Wx + sx) * 8;
          act_t* orow = p.outq[qq].p + (size_t)n * p.outq[qq].bstride + (size_t)(c0 >> 3) * HW8 + pix8;
          const int actq = p.actq[qq];
          const int nchunks = p.CT / 16;
          for (int j = 0; j < nchunks; ++j) {
            float g[16], b[16];
            uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
            if (valid) {
              x0 = *reinterpret_cast<const uint4*>(xrow + (size_t)(2 * j) * xHW8);
              x1 = *reinterpret_cast<const uint4*>(xrow + (size_t)(2 * j + 1) * xHW8);
            }
            if (!tc) {
              simt_chunk(p, n, oy, ox, ntile * p.BN + j * 16, g);
              simt_chunk(p, n, oy, ox, ntile * p.BN + p.CT + j * 16, b);
            } else {
              tmem_ld16(trow + (uint32_t)(j * 16), g);
              tmem_ld16(trow + (uint32_t)(p.CT + j * 16), b);
            }
            if (valid) {
              const uint32_t xu[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
              uint32_t o[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float xa, xb;
                unpack2(xu[c], xa, xb);
                const int ca = j * 16 + 2 * c, cb = ca + 1;
                // s_bias of the gamma half already holds (bias + 1)
                float ya = (xa - s_aux[ca]) * s_aux[p.CT + ca] * (g[2 * c] + s_bias[ca]) + (b[2 * c] + s_bias[p.CT + ca]);
                float yb = (xb - s_aux[cb]) * s_aux[p.CT + cb] * (g[2 * c + 1] + s_bias[cb]) + (b[2 * c + 1] + s_bias[p.CT + cb]);
                if (actq == ACT_LRELU) {
                  ya = lrelu02(ya);
                  yb = lrelu02(yb);
                }
                o[c] = pack2(ya, yb);
              }
              *reinterpret_cast<uint4*>(orow + (size_t)(2 * j) * HW8) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(orow + (size_t)(2 * j + 1) * HW8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        } else {  // EPI_FINAL (BN == 16)
          float v[16];
          if (!tc) simt_chunk(p, n, oy, ox, 0, v);
          else tmem_ld16(trow, v);
          if (valid) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c < p.n_valid) {
                const float y = apply_act(v[c] + s_bias[c], p.act);
                p.out_f32[(((size_t)n * p.n_valid + c) * p.H + oy) * p.W + ox] = y;
                if (p.has_out_act) {
                  const int cc = p.out_act_coff + c;
                  p.out_act.p[(size_t)n * p.out_act.bstride + (size_t)(cc >> 3) * HW8 + pix8 + (cc & 7)] = f2act(y);
                }
              }
            }
          }
        }
      }
      if (tc) {  // hand the accumulator buffer back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
      }
    }
    if (MODE == EPI_STORE && p.stats != nullptr && cur_n >= 0) {
      epi_bar();
      for (int c = e; c < p.BN; c += 128) {
        const int col = ntile * p.BN + c;
        if (col < p.n_valid) {
          atomicAdd(&p.stats[((size_t)cur_n * p.n_valid + col) * 2 + 0], (double)s_aux[c]);
          atomicAdd(&p.stats[((size_t)cur_n * p.n_valid + col) * 2 + 1], (double)s_aux[p.BN + c]);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1 && tc) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
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
  const cuuint64_t es = sizeof(act_t);
  const act_t* b0 = base + ((size_t)py * W + px) * 8;
  cuuint64_t dims[5] = {8, (cuuint64_t)(W / 2), (cuuint64_t)(H / 2), (cuuint64_t)(C / 8), (cuuint64_t)B};
  cuuint64_t strides[4] = {16 * es, (cuuint64_t)2 * W * 8 * es, (cuuint64_t)H * W * 8 * es, (cuuint64_t)bstride * es};
  cuuint32_t box[5] = {8u, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)(box_c / 8), 1u};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 5, const_cast<act_t*>(b0), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(stride-2 activation) failed: " + std::to_string((int)r));
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

int choose_bkc(int cin0, int cin1, int taps, int BN) {
  int bk = cin0 < 64 ? cin0 : 64;
  if (cin1 > 0 && cin1 < bk) bk = cin1;
  while (bk > 16 && (size_t)taps * BN * bk * 2 > 40 * 1024) bk /= 2;
  return bk;
}

int conv_gemm_configure(ConvGemmParams* p, int B, int Hout, int Wout, int cin0, int cin1, int taps, int stride,
                        int BN, int n_pad) {
  RIB_REQUIRE(taps == 1 || taps == 9, "conv_gemm: 1x1 or 3x3 only");
  RIB_REQUIRE(stride == 1 || (stride == 2 && taps == 9 && cin1 == 0), "conv_gemm: stride 2 needs a plain 3x3");
  RIB_REQUIRE(BN >= 16 && BN <= 128 && (BN & (BN - 1)) == 0 && n_pad % BN == 0, "conv_gemm: bad BN");
  const int bkc = choose_bkc(cin0, cin1, taps, BN);
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
  p->halo = taps == 9 ? 1 : 0;
  p->b_tap_bytes = (uint32_t)(BN * bkc * 2);
  p->b_stage_bytes = (uint32_t)taps * p->b_tap_bytes;
  const int S = p->stages0 + p->stages1;
  const size_t b_all = (size_t)S * p->b_stage_bytes;
  p->b_resident = b_all <= 72 * 1024 ? 1 : 0;
  p->MT = (!p->b_resident && Hout >= 2 * kTileH) ? 2 : 1;
  const int hw = stride == 1 ? kTileW + 2 * p->halo : kTileW + 1;
  const int hh = stride == 1 ? kTileH * p->MT + 2 * p->halo : kTileH * p->MT + 1;
  p->halo_w = hw;
  p->lbo = (uint32_t)(hw * hh * 16);
  p->sbo = (uint32_t)(hw * 16);
  const uint32_t tile_raw = (uint32_t)(hw * hh * 16 * (bkc / 8));
  p->a_tile_bytes = (tile_raw + 127u) & ~127u;
  const int ntile_a = stride == 2 ? 4 : 1;
  p->a_slot_bytes = ((uint32_t)ntile_a * p->a_tile_bytes + 1023u) & ~1023u;
  p->a_tx_bytes = (uint32_t)ntile_a * tile_raw;
  p->tiles_x = ceil_div(Wout, kTileW);
  p->tiles_y = ceil_div(Hout, kTileH * p->MT);
  // ring depth: resident-weight (bandwidth-bound) layers keep the CTA small so that two fit on an SM
  const size_t per_slot = p->a_slot_bytes + (p->b_resident ? 0 : p->b_stage_bytes);
  const size_t budget = p->b_resident ? (size_t)100 * 1024 - b_all : (size_t)200 * 1024;
  int ring = (int)(budget / per_slot);
  const int want = p->b_resident ? (S <= 2 ? 4 : 2 * S) : 8;
  if (ring > want) ring = want;
  if (ring > 8) ring = 8;
  if (ring < 2) ring = 2;
  p->ring = ring;
  p->idesc = make_idesc_f16(128, BN);
  return 0;
}

size_t conv_gemm_smem_bytes(const ConvGemmParams& p) {
  const int S = p.stages0 + p.stages1;
  size_t tiles = (size_t)p.ring * p.a_slot_bytes + (size_t)(p.b_resident ? S : p.ring) * p.b_stage_bytes;
  size_t bars = (size_t)(2 * p.ring + 5) * 8 + 16;
  size_t scratch = (size_t)p.BN * 4 * 3 + (size_t)(p.CT > 0 ? p.CT : 0) * 8 + 64;
  return 1024 + tiles + bars + scratch;
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
int conv_gemm_profile_collect(double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  double ms = 0.0;
  for (auto& ev : g_prof_events) {
    RIB_CHECK_CUDA(cudaEventSynchronize(ev.second));
    float t = 0.f;
    RIB_CHECK_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
    ms += t;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *total_ms = ms;
  *launches = (long long)g_prof_events.size();
  g_prof_events.clear();
  return 0;
}

template <int MODE>
static int occupancy_for(size_t smem, int* occ) {
  int n = 0;
  RIB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_gemm_kernel<MODE>, kThreads, smem));
  *occ = n;
  return 0;
}

int launch_conv_gemm(const ConvGemmParams& p, int mode, cudaStream_t stream) {
  RIB_REQUIRE(p.BKc == 16 || p.BKc == 32 || p.BKc == 64, "conv_gemm: BKc must be 16/32/64");
  RIB_REQUIRE(p.BN >= 16 && p.BN <= 128 && (p.BN & (p.BN - 1)) == 0, "conv_gemm: BN must be a power of two in [16,128]");
  RIB_REQUIRE(p.ring >= 2 && p.ring <= 8, "conv_gemm: bad ring depth");
  RIB_REQUIRE(p.MT == 1 || p.MT == 2, "conv_gemm: MT must be 1 or 2");
  RIB_REQUIRE(2 * p.MT * p.BN <= 512, "conv_gemm: accumulators exceed TMEM");
  RIB_REQUIRE(p.n_tiles >= 1, "conv_gemm: no N tiles");
  RIB_REQUIRE(mode != EPI_FINAL || p.BN == 16, "conv_gemm: EPI_FINAL needs BN == 16");
  RIB_REQUIRE(mode != EPI_SPADE || (p.BN == 2 * p.CT && p.CT % 16 == 0 && p.C % p.CT == 0),
              "conv_gemm: EPI_SPADE needs BN == 2*CT");
  const size_t smem = conv_gemm_smem_bytes(p);
  RIB_REQUIRE(smem <= 227 * 1024, "conv_gemm: shared memory budget exceeded");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    cudaError_t e;
    e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
    e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_SPADE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
    e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) attr_err = e;
  });
  RIB_CHECK_CUDA(attr_err);
  // persistent grid: as many CTAs as fit on the chip (shared memory and TMEM columns), a multiple of n_tiles
  int occ = 1;
  int rc = mode == EPI_STORE ? occupancy_for<EPI_STORE>(smem, &occ)
                             : (mode == EPI_SPADE ? occupancy_for<EPI_SPADE>(smem, &occ) : occupancy_for<EPI_FINAL>(smem, &occ));
  if (rc) return rc;
  int tmem_cols = 32;
  while (tmem_cols < 2 * p.MT * p.BN) tmem_cols <<= 1;
  if (occ > 512 / tmem_cols) occ = 512 / tmem_cols;
  if (occ > 4) occ = 4;
  if (occ < 1) occ = 1;
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.B;
  long long groups = ((long long)kNumSms * occ) / p.n_tiles;
  if (groups < 1) groups = 1;
  if (groups > m_tiles) groups = m_tiles;
  dim3 grid((unsigned)(groups * p.n_tiles));
  dim3 block(kThreads);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_profile) {
    RIB_CHECK_CUDA(cudaEventCreate(&ev0));
    RIB_CHECK_CUDA(cudaEventCreate(&ev1));
    RIB_CHECK_CUDA(cudaEventRecord(ev0, stream));
  }
  if (mode == EPI_STORE) conv_gemm_kernel<EPI_STORE><<<grid, block, smem, stream>>>(p);
  else if (mode == EPI_SPADE) conv_gemm_kernel<EPI_SPADE><<<grid, block, smem, stream>>>(p);
  else conv_gemm_kernel<EPI_FINAL><<<grid, block, smem, stream>>>(p);
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
