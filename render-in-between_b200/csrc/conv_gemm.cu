// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).  See conv_gemm.cuh.
//
// CTA = 192 threads: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + tcgen05.mma issuer
// (one lane), warps 2..5 = epilogue (TMEM lane quarter = warp_idx & 3).  A multi-stage smem ring is
// guarded by full/empty mbarriers; the accumulator [128 x BN] fp32 lives in TMEM.
#include "conv_gemm.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace rib {

static constexpr int kThreads = 192;
static constexpr int kTileM = 128;

struct KStep {
  int map, dx, dy, c0, r, s, src;
};

__device__ __forceinline__ KStep decode_kstep(const ConvGemmParams& p, int ks) {
  KStep k;
  const int main_steps = p.ntaps * p.cchunks0;
  if (ks < main_steps) {
    const int tap = ks / p.cchunks0;
    k.c0 = (ks - tap * p.cchunks0) * p.BK;
    k.r = p.ntaps == 9 ? tap / 3 : 1;
    k.s = p.ntaps == 9 ? tap - 3 * (tap / 3) : 1;
    k.src = 0;
    if (p.stride == 1) {
      k.map = 0;
      k.dx = k.s - 1;
      k.dy = k.r - 1;
    } else {  // input row 2*oy + r - 1 = 2*(oy + dy) + py
      k.map = (k.r != 1 ? 2 : 0) + (k.s != 1 ? 1 : 0);
      k.dx = k.s == 0 ? -1 : 0;
      k.dy = k.r == 0 ? -1 : 0;
    }
  } else {
    k.map = 1;
    k.dx = k.dy = 0;
    k.r = k.s = 1;
    k.src = 1;
    k.c0 = (ks - main_steps) * p.BK;
  }
  return k;
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
// (16 shuffles instead of 80).  On return lane l holds the total of column (l >> 1) & 15.
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

// Bring-up mainloop: plain loads and FMAs for one pixel row x 16 columns (selected only through
// rib_debug_set_simt(); used to bisect tcgen05/TMA problems, never in the measured path).
__device__ void simt_chunk(const ConvGemmParams& p, int n, int oy, int ox, int col0, float* acc) {
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  const int total = p.ntaps * p.cchunks0 + p.cchunks1;
  for (int ks = 0; ks < total; ++ks) {
    KStep k = decode_kstep(p, ks);
    int iy, ix;
    const act_t* src;
    int ld;
    if (k.src == 0) {
      iy = oy * p.stride + k.r - 1;
      ix = ox * p.stride + k.s - 1;
      src = p.src0;
      ld = p.ld0;
    } else {
      iy = oy;
      ix = ox;
      src = p.src1;
      ld = p.ld1;
    }
    if (iy < 0 || ix < 0 || iy >= p.Hin || ix >= p.Win) continue;
    const act_t* a = src + ((size_t)(n * p.Hin + iy) * p.Win + ix) * ld + k.c0;
    for (int kk = 0; kk < p.BK; ++kk) {
      float av = act2f(a[kk]);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        acc[c] += av * act2f(p.wpk[(size_t)(col0 + c) * p.ktotal + ks * p.BK + kk]);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int A_BYTES = kTileM * p.BK * 2;
  const int B_BYTES = p.BN * p.BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.stages * B_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);  // [BN]
  float* s_aux = s_bias + p.BN;                            // STORE: [2*BN] stats; SPADE: [2*CT] mean, rstd

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntile = blockIdx.x;
  const int tile_y = blockIdx.y / p.tiles_x;
  const int tile_x = blockIdx.y - tile_y * p.tiles_x;
  const int n = blockIdx.z;
  const int oy0 = tile_y * p.TH, ox0 = tile_x * p.TW;
  const int total_ksteps = p.ntaps * p.cchunks0 + p.cchunks1;
  const uint32_t tmem_cols = p.BN < 32 ? 32u : (uint32_t)p.BN;

  if (warp == 0 && lane == 0 && !p.debug_simt) {
    prefetch_tmap(&p.amap[0]);
    prefetch_tmap(&p.bmap);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1 && !p.debug_simt) {
    tmem_alloc(smem_u32(tmem_ptr), tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int e = threadIdx.x - 64;
    for (int c = e; c < p.BN; c += 128) s_bias[c] = p.bias[ntile * p.BN + c];
    if (MODE == EPI_STORE) {
      for (int c = e; c < 2 * p.BN; c += 128) s_aux[c] = 0.f;
    }
    if (MODE == EPI_SPADE) {
      const int tiles_per_q = p.C / p.CT;
      const int c0 = (ntile % tiles_per_q) * p.CT;
      const double cnt = (double)p.Hx * (double)p.Wx;
      for (int c = e; c < p.CT; c += 128) {
        const double s = p.xstats[((size_t)n * p.C + c0 + c) * 2 + 0];
        const double ss = p.xstats[((size_t)n * p.C + c0 + c) * 2 + 1];
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        s_aux[c] = (float)mean;
        s_aux[p.CT + c] = (float)(1.0 / sqrt(var + (double)p.eps));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = p.debug_simt ? 0u : *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && !p.debug_simt) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < total_ksteps; ++ks) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, (uint32_t)(A_BYTES + B_BYTES));
        const KStep k = decode_kstep(p, ks);
        tma_load_4d(smem_u32(sA + stage * A_BYTES), &p.amap[k.map], fb, k.c0, ox0 + k.dx, oy0 + k.dy, n);
        tma_load_2d(smem_u32(sB + stage * B_BYTES), &p.bmap, fb, ks * p.BK, ntile * p.BN);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && !p.debug_simt) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t row_bytes = (uint32_t)p.BK * 2u;
      const int kk_steps = p.BK / 16;
      for (int ks = 0; ks < total_ksteps; ++ks) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint64_t adesc = make_kmajor_desc(smem_u32(sA + stage * A_BYTES), row_bytes);
        const uint64_t bdesc = make_kmajor_desc(smem_u32(sB + stage * B_BYTES), row_bytes);
        for (int kk = 0; kk < kk_steps; ++kk)
          umma_f16(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), p.idesc,
                   (uint32_t)((ks | kk) != 0));
        umma_commit(smem_u32(&empty_bar[stage]));  // frees this smem stage once the MMAs have read it
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(smem_u32(tmem_full_bar));
    }
  } else {
    // ===================== Epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty = m / p.TW, tx = m - ty * p.TW;
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool valid = (oy < p.H) && (ox < p.W);
    const size_t pix = ((size_t)n * p.H + oy) * p.W + ox;
    if (!p.debug_simt) {
      mbar_wait(smem_u32(tmem_full_bar), 0u);
      tc_fence_after();
    }
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

    if (MODE == EPI_STORE) {
      const int nchunks = p.BN / 16;
      for (int j = 0; j < nchunks; ++j) {
        float v[16];
        const int col0 = ntile * p.BN + j * 16;
        if (p.debug_simt) simt_chunk(p, n, oy, ox, col0, v);
        else tmem_ld16(trow + (uint32_t)(j * 16), v);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] += s_bias[j * 16 + c];
        if (p.res != nullptr && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.ldr + col0);
          uint4 r0 = rp[0], r1 = rp[1];
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
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.ldo + col0);
          op[0] = make_uint4(o[0], o[1], o[2], o[3]);
          op[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      if (p.stats != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int e = threadIdx.x - 64;
        for (int c = e; c < p.BN; c += 128) {
          const int col = ntile * p.BN + c;
          if (col < p.n_valid) {
            atomicAdd(&p.stats[((size_t)n * p.n_valid + col) * 2 + 0], (double)s_aux[c]);
            atomicAdd(&p.stats[((size_t)n * p.n_valid + col) * 2 + 1], (double)s_aux[p.BN + c]);
          }
        }
      }
    } else if (MODE == EPI_SPADE) {
      const int tiles_per_q = p.C / p.CT;
      const int qq = ntile / tiles_per_q;
      const int c0 = (ntile - qq * tiles_per_q) * p.CT;
      const int sy = p.ups ? (oy >> 1) : oy, sx = p.ups ? (ox >> 1) : ox;
      const act_t* xrow = p.x + (((size_t)n * p.Hx + sy) * p.Wx + sx) * p.ldx + c0;
      act_t* orow = p.outq[qq] + pix * p.ldq[qq] + c0;
      const int actq = p.actq[qq];
      const int nchunks = p.CT / 16;
      for (int j = 0; j < nchunks; ++j) {
        float g[16], b[16];
        if (p.debug_simt) {
          simt_chunk(p, n, oy, ox, ntile * p.BN + j * 16, g);
          simt_chunk(p, n, oy, ox, ntile * p.BN + p.CT + j * 16, b);
        } else {
          tmem_ld16(trow + (uint32_t)(j * 16), g);
          tmem_ld16(trow + (uint32_t)(p.CT + j * 16), b);
        }
        if (valid) {
          const uint4* xp = reinterpret_cast<const uint4*>(xrow + j * 16);
          uint4 x0 = xp[0], x1 = xp[1];
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
          uint4* op = reinterpret_cast<uint4*>(orow + j * 16);
          op[0] = make_uint4(o[0], o[1], o[2], o[3]);
          op[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    } else {  // EPI_FINAL (BN == 16)
      float v[16];
      if (p.debug_simt) simt_chunk(p, n, oy, ox, 0, v);
      else tmem_ld16(trow, v);
      if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < p.n_valid) {
            const float y = apply_act(v[c] + s_bias[c], p.act);
            p.out_f32[(((size_t)n * p.n_valid + c) * p.H + oy) * p.W + ox] = y;
            if (p.out_act != nullptr) p.out_act[pix * p.ld_act + c] = f2act(y);
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1 && !p.debug_simt) {
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

int make_tmap_act(CUtensorMap* m, const act_t* base, int C, int W, int H, int B, size_t strideW, size_t strideH,
                  size_t strideB, int boxC, int boxW, int boxH) {
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  RIB_REQUIRE(boxC == 16 || boxC == 32 || boxC == 64, "activation K chunk must be 16, 32 or 64 channels");
  RIB_REQUIRE(((uintptr_t)base & 15) == 0 && strideW % 16 == 0, "activation view must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)strideW, (cuuint64_t)strideH, (cuuint64_t)strideB};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 4, const_cast<act_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxC * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
  return 0;
}

int make_tmap_w(CUtensorMap* m, const act_t* w, int K, int N, int boxK, int boxN) {
  EncodeTiledFn fn = get_encode_fn();
  RIB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxN};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, RIB_TMAP_DTYPE, 2, const_cast<act_t*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxK * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RIB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
  return 0;
}

void choose_tile(int H, int W, int* TW, int* TH) {
  long best = -1;
  int bw = 16, bh = 8;
  const int cand[6][2] = {{16, 8}, {32, 4}, {8, 16}, {64, 2}, {128, 1}, {4, 32}};
  for (int i = 0; i < 6; ++i) {
    const int tw = cand[i][0], th = cand[i][1];
    const long area = (long)ceil_div(W, tw) * tw * (long)ceil_div(H, th) * th;
    if (best < 0 || area < best) {
      best = area;
      bw = tw;
      bh = th;
    }
  }
  *TW = bw;
  *TH = bh;
}

size_t conv_gemm_smem_bytes(const ConvGemmParams& p) {
  size_t tiles = (size_t)p.stages * ((size_t)kTileM * p.BK * 2 + (size_t)p.BN * p.BK * 2);
  size_t bars = (size_t)(2 * p.stages + 1) * 8 + 16;
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

int launch_conv_gemm(const ConvGemmParams& p, int mode, cudaStream_t stream) {
  RIB_REQUIRE(p.TW * p.TH == kTileM, "conv_gemm: spatial tile must hold 128 pixels");
  RIB_REQUIRE(p.BK == 16 || p.BK == 32 || p.BK == 64, "conv_gemm: BK must be 16/32/64");
  RIB_REQUIRE(p.BN >= 16 && p.BN <= 256 && (p.BN & (p.BN - 1)) == 0,
              "conv_gemm: BN must be a power of two in [16,256]");
  RIB_REQUIRE(p.stages >= 1 && p.stages <= 16, "conv_gemm: bad stage count");
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
  dim3 grid((unsigned)p.n_tiles, (unsigned)(p.tiles_x * p.tiles_y), (unsigned)p.B);
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
