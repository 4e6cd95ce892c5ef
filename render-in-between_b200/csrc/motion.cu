// Motion Transformer of Human_Motion_Modelling on the GPU.
//
//   Transformer.forward / encode / decode / interpolate_embedding    Human_Motion_Modelling/models/transformer.py:57-132
//   TransformerEncoderLayer.forward_pre / DecoderLayer.forward_pre   Human_Motion_Modelling/models/transformer.py:243-254, :316-337
//   final LayerNorms of TransformerEncoder / TransformerDecoder       Human_Motion_Modelling/models/transformer.py:134-196
// for the shipped configuration (configs/config.yaml:77-94): pre-norm, leaky_relu, two_stage, eval mode.
//
// The model is tiny (d_model 128, 8 heads of 16, feed-forward 256, 6 + 6 layers, at most 321 tokens): ~1.3 GFLOP per
// sequence.  Everything stays in fp32 on the CUDA cores (the parity target is the reference's fp32 output, and at these
// sizes a forward is bound by the ~90 dependent launches, not by arithmetic): three kernels,
//   linear   Y = act(LN?(X) (+ pos for the first n_pos columns) . W^T + b) (+ residual): 32 x 64 output tile per block, both
//            operand tiles resident in shared memory over the whole K (all loads of a block issued up front), 4 x 4
//            outputs per thread; the LayerNorm of the pre-norm blocks and the "+ pos" of q / k are applied to the A tile
//   mha      one block per (16 or 64 queries, head, sequence): the head's K and V rows in shared memory, a warp per four
//            queries at a time (scores, masks as -inf, softmax, P.V; recursive-halving shuffle reduction)
//   interp   interpolate_embedding, the exact float sequence of the reference
// launched back to back with programmatic dependent launch on the caller's stream.
#include "motion.cuh"

#include <map>
#include <string>
#include <vector>

namespace rib {

namespace {

constexpr int kDH = 16;          // head dimension (hidden_dim / nheads of the shipped configuration)
constexpr int kLinRows = 32, kLinCols = 64, kLinMaxK = 256;
constexpr int kMhaWarps = 4;     // warps per attention block; a warp serves QW queries, one after the other

struct LinParams {
  const float* X;                // element (b, r, k) at X[b * xb + r * xr + k * xk]
  long long xb, xr, xk;
  const float* W;                // TRANSPOSED weight: element (k, n) at W[k * ldw + n] (made once by motion_create from the
  int ldw;                       //   [N][K] nn.Linear weight; ldw a multiple of 4, rows zero-padded), 16-byte aligned
  const float* bias;             // [N]
  const float* ln_g;             // LayerNorm over K applied to every row of X first (null: none)
  const float* ln_b;
  const float* pos;              // element (b, r, k) at pos[b * pb + r * pr + k]: added to the (normalised) row for the
  long long pb, pr;              //   output columns < n_pos (q / k projections take x + pos, v takes x)
  int n_pos;
  const float* R;                // residual element (b, r, n) at R[b * rb + r * rr + n * rn] (null: none)
  long long rb, rr, rn;
  float* Y;                      // element (b, r, n) at Y[b * yb + r * yr + n]
  long long yb, yr;
  int L, N, K, act;              // act 1: leaky_relu(0.01)
};

// Shared memory of the linear kernel (dynamic): As[kLinRows][Kp + 4] (the A tile, row-major) and Wt[Kp][kLinCols + 4] (the
// weight tile, k-major), Kp = K rounded up to 4.  The whole K extent of both is resident and arrives by 16-byte cp.async
// copies that are all in flight at once (one memory latency per block, not one per K chunk; the weights, constants, are
// requested before the wait on the previous kernel); the product loop then runs without barriers, 4 x 4 outputs per thread,
// four k per step with 16-byte shared-memory loads of both operands.
constexpr int kLinWP = kLinCols + 4;
static size_t linear_smem_bytes(int K) {
  const int Kp = (K + 3) & ~3;
  return ((size_t)kLinRows * (Kp + 4) + (size_t)Kp * kLinWP) * sizeof(float);
}

__global__ void __launch_bounds__(128) motion_linear_kernel(const LinParams p) {
  extern __shared__ __align__(16) float lin_sm[];
  const int K = p.K, Kp = (K + 3) & ~3, AP = Kp + 4;
  float* As = lin_sm;                          // [kLinRows][AP]
  float* Wt = lin_sm + (size_t)kLinRows * AP;  // [Kp][kLinWP]
  const int b = blockIdx.z, r0 = blockIdx.x * kLinRows, n0 = blockIdx.y * kLinCols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {   // weight tile: rows k < K of the transposed weight, 16 chunks of 16 bytes each (zero fill behind ldw and for k >= K)
    const uint32_t wt0 = smem_u32(Wt);
    for (int i = tid; i < Kp * (kLinCols / 4); i += 128) {
      const int k = i >> 4, c = (i & 15) * 4;
      const bool ok = k < K && n0 + c < ((p.N + 3) & ~3);   // (rows are padded to a multiple of 4 columns)
      cp_async_16(wt0 + (uint32_t)(k * kLinWP + c) * 4u, ok ? p.W + (size_t)k * p.ldw + n0 + c : p.W, ok ? 16u : 0u);
    }
    cp_async_commit();
  }
  pdl_wait();   // programmatic dependent launch: everything below reads the previous kernel's output
  const bool a_async = p.xk == 1 && (K & 3) == 0 && (p.xr & 3) == 0 && (p.xb & 3) == 0 && ((uintptr_t)p.X & 15) == 0;
  if (a_async) {
    const uint32_t as0 = smem_u32(As);
    const int cpr = K >> 2;   // 16-byte chunks per row
    for (int i = tid; i < kLinRows * cpr; i += 128) {
      const int r = i / cpr, c = (i - r * cpr) * 4;
      const bool ok = r0 + r < p.L;
      cp_async_16(as0 + (uint32_t)(r * AP + c) * 4u, ok ? p.X + (size_t)b * p.xb + (size_t)(r0 + r) * p.xr + c : p.X, ok ? 16u : 0u);
    }
  } else if (p.xk == 1) {
    for (int i = tid; i < kLinRows * Kp; i += 128) {
      const int r = i / Kp, k = i - r * Kp;
      As[r * AP + k] = (r0 + r < p.L && k < K) ? p.X[(size_t)b * p.xb + (size_t)(r0 + r) * p.xr + k] : 0.f;
    }
  } else {   // [B][C][L] input: contiguous along the rows
    for (int i = tid; i < kLinRows * Kp; i += 128) {
      const int k = i / kLinRows, r = i - k * kLinRows;
      As[r * AP + k] = (r0 + r < p.L && k < K) ? p.X[(size_t)b * p.xb + (size_t)(r0 + r) * p.xr + (size_t)k * p.xk] : 0.f;
    }
  }
  cp_async_commit();
  cp_async_wait_group<0>();
  __syncthreads();
  if (p.ln_g != nullptr) {   // nn.LayerNorm(K), eps 1e-5: mean, biased variance (two passes), affine
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {
      float* row = As + r * AP;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += row[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / (float)K;
      float v = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float d = row[k] - mean;
        v = fmaf(d, d, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float rstd = rsqrtf(v / (float)K + 1e-5f);
      for (int k = lane; k < K; k += 32) row[k] = (row[k] - mean) * rstd * p.ln_g[k] + p.ln_b[k];
    }
    __syncthreads();
  }
  if (p.pos != nullptr && n0 < p.n_pos) {
    for (int i = tid; i < kLinRows * K; i += 128) {
      const int r = i / K, k = i - r * K;
      if (r0 + r < p.L) As[r * AP + k] += p.pos[(size_t)b * p.pb + (size_t)(r0 + r) * p.pr + k];
    }
    __syncthreads();
  }
  const int tx = tid & 15, ty = tid >> 4;   // columns n0 + 4 tx .. + 3, rows 4 ty .. + 3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* arow = As + (4 * ty) * AP;
  const float* wcol = Wt + 4 * tx;
#pragma unroll 2
  for (int k = 0; k < Kp; k += 4) {
    float4 a[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * AP + k);
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(wcol + (k + q) * kLinWP);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // ascending k: one summation order per output
        acc[i][0] = fmaf(av[q], w[q].x, acc[i][0]);
        acc[i][1] = fmaf(av[q], w[q].y, acc[i][1]);
        acc[i][2] = fmaf(av[q], w[q].z, acc[i][2]);
        acc[i][3] = fmaf(av[q], w[q].w, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 4 * ty + i;
    if (r >= p.L) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tx + j;
      if (n >= p.N) continue;
      float v = acc[i][j] + p.bias[n];
      if (p.act == 1) v = v > 0.f ? v : 0.01f * v;
      if (p.R != nullptr) v += p.R[(size_t)b * p.rb + (size_t)r * p.rr + (size_t)n * p.rn];
      p.Y[(size_t)b * p.yb + (size_t)r * p.yr + n] = v;
    }
  }
}

// [N][K] nn.Linear weight -> transposed, zero-padded [K][ld] copy (once, at model creation)
__global__ void motion_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int N, int K, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * ld) return;
  const int k = i / ld, n = i - k * ld;
  wt[i] = n < N ? w[(size_t)n * K + k] : 0.f;
}

// nn.LayerNorm over the last dimension of [B][L][E] (the encoder's final norm: its output is the decoder's memory).
__global__ void __launch_bounds__(256) motion_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                               const float* __restrict__ be, float* __restrict__ y,
                                                               int rows, int E) {
  pdl_wait();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + (size_t)r * E;
  float s = 0.f;
  for (int k = lane; k < E; k += 32) s += xr[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)E;
  float v = 0.f;
  for (int k = lane; k < E; k += 32) {
    const float d = xr[k] - mean;
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)E + 1e-5f);
  for (int k = lane; k < E; k += 32) y[(size_t)r * E + k] = (xr[k] - mean) * rstd * g[k] + be[k];
}

// nn.MultiheadAttention core (after the in-projection, before the out-projection): q scaled by sqrt(1 / head_dim) first,
// scores + masks (-inf), softmax over the keys, P . V.  Q / K / V / O rows are `ld*` floats apart (slices of the packed
// projection buffer); head h owns columns [16 h, 16 h + 16).
struct MhaParams {
  const float *Q, *K, *V;
  long long qb, kb, vb;      // floats between sequences
  int ldq, ldk, ldv;
  float* O;
  long long ob;
  int ldo;
  const uint8_t* kpm;        // [B][Lk], non-zero = the key is ignored (null: none)
  int Lq, Lk, eye;           // eye: query i may not attend to key i (Transformer.encode's mask)
};

// A warp serves its QW queries four at a time: a lane owns every 32nd key, reads the key's K row once (four 16-byte
// shared-memory loads) for the four dot products and its V row once for the four weighted sums, so the 64 FMAs of a step
// cost 4 loads instead of 64.  The 4 x 16 partial outputs of the 32 lanes are combined by recursive halving (a lane sends
// the half of its values the partner is responsible for: 62 shuffles instead of 320) and every lane ends with two adjacent
// output values of one query.
constexpr int kMhaQT = 4;            // queries per step
constexpr int kMhaKP = kDH + 4;      // K / V row pitch in floats: 16-byte aligned rows, conflict-free 16-byte loads

template <int QW>   // queries per warp: 4 keeps a single sequence spread over the chip, 16 amortises the K / V staging of a batch
__global__ void __launch_bounds__(128) motion_mha_kernel(const MhaParams p) {
  extern __shared__ __align__(16) float sm[];
  const int Lk = p.Lk, lkp = (Lk + 31) & ~31;
  float* Ks = sm;                                  // [Lk][kMhaKP]
  float* Vs = Ks + (size_t)Lk * kMhaKP;
  float* S = Vs + (size_t)Lk * kMhaKP;             // [4 warps][kMhaQT][lkp]
  float* qs = S + (size_t)kMhaWarps * kMhaQT * lkp;   // [4 warps][kMhaQT][16]
  pdl_wait();
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (kMhaWarps * QW);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < Lk * (kDH / 4); i += 128) {
    const int j = i >> 2, c = (i & 3) * 4;
    *reinterpret_cast<float4*>(&Ks[j * kMhaKP + c]) =
        *reinterpret_cast<const float4*>(&p.K[(size_t)b * p.kb + (size_t)j * p.ldk + h * kDH + c]);
    *reinterpret_cast<float4*>(&Vs[j * kMhaKP + c]) =
        *reinterpret_cast<const float4*>(&p.V[(size_t)b * p.vb + (size_t)j * p.ldv + h * kDH + c]);
  }
  __syncthreads();
  const uint8_t* kpm = p.kpm != nullptr ? p.kpm + (size_t)b * Lk : nullptr;
  float* Sw = S + (size_t)warp * kMhaQT * lkp;
  float* qw = qs + warp * kMhaQT * kDH;
  for (int qb = q0 + warp * QW; qb < q0 + warp * QW + QW && qb < p.Lq; qb += kMhaQT) {
    // the four queries of this step (clamped behind the sequence: computed, not stored), scaled by sqrt(1 / 16)
    for (int i = lane; i < kMhaQT * kDH; i += 32) {
      const int qi = min(qb + (i >> 4), p.Lq - 1);
      qw[i] = p.Q[(size_t)b * p.qb + (size_t)qi * p.ldq + h * kDH + (i & 15)] * 0.25f;
    }
    __syncwarp();
    float q[kMhaQT][kDH];
#pragma unroll
    for (int g = 0; g < kMhaQT; ++g)
#pragma unroll
      for (int d = 0; d < kDH; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&qw[g * kDH + d]);
        q[g][d] = t.x, q[g][d + 1] = t.y, q[g][d + 2] = t.z, q[g][d + 3] = t.w;
      }
    float mx[kMhaQT];
#pragma unroll
    for (int g = 0; g < kMhaQT; ++g) mx[g] = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      float kr[kDH];
#pragma unroll
      for (int d = 0; d < kDH; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ks[j * kMhaKP + d]);
        kr[d] = t.x, kr[d + 1] = t.y, kr[d + 2] = t.z, kr[d + 3] = t.w;
      }
      const bool hidden = kpm != nullptr && kpm[j];
#pragma unroll
      for (int g = 0; g < kMhaQT; ++g) {
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < kDH; ++d) sc = fmaf(q[g][d], kr[d], sc);
        if (hidden || (p.eye && j == qb + g)) sc = -INFINITY;
        Sw[g * lkp + j] = sc;
        mx[g] = fmaxf(mx[g], sc);
      }
    }
#pragma unroll
    for (int g = 0; g < kMhaQT; ++g)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], o));
    float sum[kMhaQT], acc[kMhaQT * kDH];   // acc[g * 16 + d]
#pragma unroll
    for (int g = 0; g < kMhaQT; ++g) sum[g] = 0.f;
#pragma unroll
    for (int i = 0; i < kMhaQT * kDH; ++i) acc[i] = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      float vr[kDH];
#pragma unroll
      for (int d = 0; d < kDH; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Vs[j * kMhaKP + d]);
        vr[d] = t.x, vr[d + 1] = t.y, vr[d + 2] = t.z, vr[d + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < kMhaQT; ++g) {
        const float e = expf(Sw[g * lkp + j] - mx[g]);   // (all keys masked: -inf - -inf = NaN, as in torch)
        sum[g] += e;
#pragma unroll
        for (int d = 0; d < kDH; ++d) acc[g * kDH + d] = fmaf(e, vr[d], acc[g * kDH + d]);
      }
    }
#pragma unroll
    for (int g = 0; g < kMhaQT; ++g)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], o);
    // recursive halving over the 64 partial outputs: after the step with partner distance o, a lane keeps the half of its
    // values whose index bit matches its lane bit; the final lane l owns indices 2 l and 2 l + 1 (query l / 8)
#pragma unroll
    for (int o = 16, n = kMhaQT * kDH / 2; o > 0; o >>= 1, n >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float send = up ? acc[i] : acc[i + n];
        const float keep = up ? acc[i + n] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    const int g = lane >> 3, qi = qb + g;
    float sg = sum[0];
#pragma unroll
    for (int t = 1; t < kMhaQT; ++t)
      if (g == t) sg = sum[t];
    if (qi < p.Lq && qi < q0 + warp * QW + QW)
      *reinterpret_cast<float2*>(&p.O[(size_t)b * p.ob + (size_t)qi * p.ldo + h * kDH + 2 * (lane & 7)]) =
          make_float2(acc[0] / sg, acc[1] / sg);
    __syncwarp();
  }
}

// Transformer.interpolate_embedding (transformer.py:57-73) on reco [L][B][C] -> interp [B][L][C]:
//   prev / rate * (rate - i % rate) + next / rate * (i % rate), separately rounded operations as in torch.
__global__ void motion_interp_kernel(const float* __restrict__ reco, float* __restrict__ interp, int B, int L, int C,
                                     int rate) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L * C) return;
  const int c = i % C, l = (i / C) % L, b = i / (C * L);
  const int chunk = l / rate, rem = l - chunk * rate;
  const int lp = chunk * rate, ln = (l == L - 1) ? L - 1 : (chunk + 1) * rate;
  const float prev = reco[((size_t)lp * B + b) * C + c], next = reco[((size_t)ln * B + b) * C + c];
  const float fr = (float)rate;
  interp[i] = __fadd_rn(__fmul_rn(__fdiv_rn(prev, fr), (float)(rate - rem)), __fmul_rn(__fdiv_rn(next, fr), (float)rem));
}

struct WT {               // transposed weight [K][ld]
  const float* p;
  int ld;
};
struct AttnW {
  WT in_w, out_w;
  const float *in_b, *out_b;
};
struct LayerW {
  AttnW self, cross;
  WT l1w, l2w;
  const float *l1b, *l2b, *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
};

}  // namespace

struct MotionModel {
  rib_motion_config cfg;
  float* blob = nullptr;   // one device buffer with every parameter
  WT in_w, out_w;
  const float *in_b, *out_b, *enc_ng, *enc_nb, *dec_ng, *dec_nb;
  std::vector<LayerW> enc, dec;
};

int motion_create(const rib_motion_config* cfg, const rib_tensor* tensors, int n_tensors, cudaStream_t stream,
                  MotionModel** out) {
  RIB_REQUIRE(cfg && tensors && out, "motion_create: null argument");
  const int J = cfg->input_joints, E = cfg->hidden_dim, FF = cfg->dim_feedforward;
  RIB_REQUIRE(J > 0 && J <= kLinMaxK && E > 0 && E <= kLinMaxK && FF > 0 && FF <= kLinMaxK,
              "motion_create: dimensions above 256 are not supported");
  RIB_REQUIRE(cfg->nheads > 0 && E == cfg->nheads * kDH, "motion_create: head dimension must be 16");
  RIB_REQUIRE(E % kLinCols == 0, "motion_create: hidden_dim must be a multiple of 64");
  RIB_REQUIRE(cfg->enc_layers >= 0 && cfg->dec_layers >= 0, "motion_create: bad layer count");
  std::map<std::string, std::pair<const float*, long long>> src;
  size_t total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    RIB_REQUIRE(tensors[i].name && tensors[i].data && tensors[i].numel > 0, "motion_create: bad tensor entry");
    src[tensors[i].name] = {tensors[i].data, tensors[i].numel};
    total += (((size_t)tensors[i].numel + 3) & ~(size_t)3) + 4 * (size_t)kLinMaxK;   // room for the row padding of a transposed copy
  }
  MotionModel* m = new MotionModel();
  m->cfg = *cfg;
  if (cudaMalloc(&m->blob, total * sizeof(float)) != cudaSuccess) {
    delete m;
    set_error("motion_create: cudaMalloc failed");
    return -2;
  }
  size_t off = 0;
  bool ok = true;
  std::string missing;
  auto take = [&](const std::string& key, long long numel) -> const float* {
    auto it = src.find(key);
    if (it == src.end() || it->second.second != numel) {
      ok = false;
      missing = key;
      return nullptr;
    }
    float* dst = m->blob + off;
    off += ((size_t)numel + 3) & ~(size_t)3;
    if (cudaMemcpyAsync(dst, it->second.first, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
      ok = false;
    return dst;
  };
  // a 2-D nn.Linear weight [N][K]: kept as its transposed, row-padded copy [K][ld] (what the linear kernel stages)
  auto take_w = [&](const std::string& key, int N, int K) -> WT {
    WT r = {nullptr, (N + 3) & ~3};
    auto it = src.find(key);
    if (it == src.end() || it->second.second != (long long)N * K) {
      ok = false;
      missing = key;
      return r;
    }
    off = (off + 3) & ~(size_t)3;   // 16-byte aligned rows
    float* dst = m->blob + off;
    off += (size_t)K * r.ld;
    motion_transpose_kernel<<<(unsigned)ceil_div(K * r.ld, 256), 256, 0, stream>>>(it->second.first, dst, N, K, r.ld);
    if (cudaGetLastError() != cudaSuccess) ok = false;
    r.p = dst;
    return r;
  };
  auto attn = [&](const std::string& p) {
    AttnW a;
    a.in_w = take_w(p + ".in_proj_weight", 3 * E, E);
    a.in_b = take(p + ".in_proj_bias", 3LL * E);
    a.out_w = take_w(p + ".out_proj.weight", E, E);
    a.out_b = take(p + ".out_proj.bias", E);
    return a;
  };
  auto layer = [&](const std::string& p, bool dec) {
    LayerW l = {};
    l.self = attn(p + ".self_attn");
    if (dec) l.cross = attn(p + ".multihead_attn");
    l.l1w = take_w(p + ".linear1.weight", FF, E);
    l.l1b = take(p + ".linear1.bias", FF);
    l.l2w = take_w(p + ".linear2.weight", E, FF);
    l.l2b = take(p + ".linear2.bias", E);
    l.n1g = take(p + ".norm1.weight", E);
    l.n1b = take(p + ".norm1.bias", E);
    l.n2g = take(p + ".norm2.weight", E);
    l.n2b = take(p + ".norm2.bias", E);
    if (dec) {
      l.n3g = take(p + ".norm3.weight", E);
      l.n3b = take(p + ".norm3.bias", E);
    }
    return l;
  };
  m->in_w = take_w("input_embed.weight", E, J);
  m->in_b = take("input_embed.bias", E);
  for (int i = 0; i < cfg->enc_layers; ++i) m->enc.push_back(layer("encoder.layers." + std::to_string(i), false));
  m->enc_ng = take("encoder.norm.weight", E);
  m->enc_nb = take("encoder.norm.bias", E);
  for (int i = 0; i < cfg->dec_layers; ++i) m->dec.push_back(layer("decoder.layers." + std::to_string(i), true));
  m->dec_ng = take("decoder.norm.weight", E);
  m->dec_nb = take("decoder.norm.bias", E);
  m->out_w = take_w("joints_embed.weight", J, E);
  m->out_b = take("joints_embed.bias", J);
  if (ok && cudaStreamSynchronize(stream) != cudaSuccess) ok = false;
  if (!ok) {
    cudaFree(m->blob);
    delete m;
    set_error(missing.empty() ? std::string("motion_create: copying the parameters failed")
                              : "motion_create: state-dict entry missing or of the wrong size: " + missing);
    return -1;
  }
  *out = m;
  return 0;
}

void motion_destroy(MotionModel* m) {
  if (m == nullptr) return;
  cudaFree(m->blob);
  delete m;
}

// workspace: x, y, att, mem [B L E]; qkv [B L 3E]; ffh [B L FF]; interp [B L J]
long long motion_workspace_bytes(const MotionModel* m, int B, int L) {
  const long long E = m->cfg.hidden_dim, FF = m->cfg.dim_feedforward, J = m->cfg.input_joints;
  return (long long)B * L * (7 * E + FF + ((J + 3) & ~3LL)) * (long long)sizeof(float) + 256;
}

namespace {

int run_linear(const LinParams& p, int B, cudaStream_t s) {
  RIB_REQUIRE(p.K <= kLinMaxK, "motion: inner dimension above 256");
  RIB_REQUIRE(p.pos == nullptr || p.n_pos >= p.N || p.n_pos % kLinCols == 0, "motion: n_pos must fall on a column tile");
  const size_t smem = linear_smem_bytes(p.K);
  if (smem > 48 * 1024)
    RIB_CHECK_CUDA(cudaFuncSetAttribute((const void*)motion_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)linear_smem_bytes(kLinMaxK)));
  const dim3 grid((unsigned)ceil_div(p.L, kLinRows), (unsigned)ceil_div(p.N, kLinCols), (unsigned)B);
  launch_pdl(motion_linear_kernel, grid, dim3(128), smem, s, p);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int run_mha(const MhaParams& p, int B, int H, cudaStream_t s) {
  const int lkp = (p.Lk + 31) & ~31;
  const size_t smem = ((size_t)2 * p.Lk * kMhaKP + (size_t)kMhaWarps * kMhaQT * (lkp + kDH)) * sizeof(float);
  RIB_REQUIRE(((uintptr_t)p.K & 15) == 0 && ((uintptr_t)p.V & 15) == 0 && ((uintptr_t)p.O & 7) == 0 && p.ldk % 4 == 0 &&
                  p.ldv % 4 == 0 && p.kb % 4 == 0 && p.vb % 4 == 0 && p.ldo % 2 == 0 && p.ob % 2 == 0,
              "motion: attention operands must be 16-byte aligned");
  RIB_REQUIRE(smem <= 200 * 1024, "motion: sequences longer than ~1100 frames are not supported");
  const bool wide = (long long)B * H * ceil_div(p.Lq, kMhaWarps * 4) > 2048;   // enough blocks: 64 queries per block
  if (smem > 48 * 1024) {
    RIB_CHECK_CUDA(cudaFuncSetAttribute((const void*)motion_mha_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RIB_CHECK_CUDA(cudaFuncSetAttribute((const void*)motion_mha_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (wide) launch_pdl(motion_mha_kernel<16>, dim3((unsigned)ceil_div(p.Lq, kMhaWarps * 16), (unsigned)H, (unsigned)B), dim3(128), smem, s, p);
  else launch_pdl(motion_mha_kernel<4>, dim3((unsigned)ceil_div(p.Lq, kMhaWarps * 4), (unsigned)H, (unsigned)B), dim3(128), smem, s, p);
  RIB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int motion_forward(MotionModel* m, int B, int L, const float* src, const uint8_t* src_mask, const float* src_pos,
                   const uint8_t* tgt_mask, const float* tgt_pos, int rate, float* joints, float* reco, void* workspace,
                   long long workspace_bytes, cudaStream_t stream) {
  RIB_REQUIRE(m && src && src_pos && tgt_pos && joints && reco && workspace, "motion_forward: null argument");
  RIB_REQUIRE(B >= 1 && L >= 2 && rate >= 1, "motion_forward: bad shape");
  RIB_REQUIRE((L - 1) % rate == 0, "motion_forward: the sequence length must be rate * n + 1 (transformer.py:57-73)");
  RIB_REQUIRE(workspace_bytes >= motion_workspace_bytes(m, B, L), "motion_forward: workspace too small");
  RIB_REQUIRE(((uintptr_t)workspace & 15) == 0, "motion_forward: workspace must be 16-byte aligned");
  const int E = m->cfg.hidden_dim, FF = m->cfg.dim_feedforward, J = m->cfg.input_joints, H = m->cfg.nheads;
  const long long BL = (long long)B * L;
  float* x = static_cast<float*>(workspace);
  float* y = x + BL * E;
  float* att = y + BL * E;
  float* mem = att + BL * E;
  float* qkv = mem + BL * E;
  float* ffh = qkv + BL * 3 * E;
  float* interp = ffh + BL * FF;
  const long long sE = (long long)L * E, s3E = (long long)L * 3 * E, sFF = (long long)L * FF, sJ = (long long)L * J;

  auto lin = [&](const float* X, long long xb, long long xr, long long xk, int K, WT W, const float* bias, int N, float* Y,
                 long long yb, long long yr) {
    LinParams p = {};
    p.X = X, p.xb = xb, p.xr = xr, p.xk = xk;
    p.W = W.p, p.ldw = W.ld, p.bias = bias, p.N = N, p.K = K;
    p.Y = Y, p.yb = yb, p.yr = yr;
    p.L = L;
    return p;
  };
  auto with_ln = [](LinParams p, const float* g, const float* b) {
    p.ln_g = g, p.ln_b = b;
    return p;
  };
  auto with_pos = [&](LinParams p, const float* pos, int n_pos) {   // pos is [L][B][E]
    p.pos = pos, p.pb = E, p.pr = (long long)B * E, p.n_pos = n_pos;
    return p;
  };
  auto with_res = [](LinParams p, const float* R, long long rb, long long rr, long long rn) {
    p.R = R, p.rb = rb, p.rr = rr, p.rn = rn;
    return p;
  };
  auto mha = [&](const float* Q, const float* K, const float* V, const uint8_t* kpm, int eye) {
    MhaParams p = {};
    p.Q = Q, p.K = K, p.V = V;
    p.qb = p.kb = p.vb = s3E;
    p.ldq = p.ldk = p.ldv = 3 * E;
    p.O = att, p.ob = sE, p.ldo = E;
    p.kpm = kpm, p.Lq = L, p.Lk = L, p.eye = eye;
    return run_mha(p, B, H, stream);
  };
  int rc;
#define RIB_MOTION_RUN(expr) \
  if ((rc = (expr)) != 0) return rc

  // x = input_embed(src^T)            src is [B][J][L]
  RIB_MOTION_RUN(run_linear(lin(src, (long long)J * L, 1, L, J, m->in_w, m->in_b, E, x, sE, E), B, stream));
  for (const LayerW& l : m->enc) {
    // q, k = (norm1(x) + pos) Wq, Wk; v = norm1(x) Wv
    RIB_MOTION_RUN(run_linear(with_pos(with_ln(lin(x, sE, E, 1, E, l.self.in_w, l.self.in_b, 3 * E, qkv, s3E, 3 * E), l.n1g, l.n1b),
                                       src_pos, 2 * E), B, stream));
    RIB_MOTION_RUN(mha(qkv, qkv + E, qkv + 2 * E, src_mask, 1));
    RIB_MOTION_RUN(run_linear(with_res(lin(att, sE, E, 1, E, l.self.out_w, l.self.out_b, E, x, sE, E), x, sE, E, 1), B, stream));
    LinParams f1 = with_ln(lin(x, sE, E, 1, E, l.l1w, l.l1b, FF, ffh, sFF, FF), l.n2g, l.n2b);
    f1.act = 1;
    RIB_MOTION_RUN(run_linear(f1, B, stream));
    RIB_MOTION_RUN(run_linear(with_res(lin(ffh, sFF, FF, 1, FF, l.l2w, l.l2b, E, x, sE, E), x, sE, E, 1), B, stream));
  }
  {   // memory = encoder.norm(x)
    launch_pdl(motion_layernorm_kernel, dim3((unsigned)ceil_div((int)BL, 8)), dim3(256), 0, stream, (const float*)x, m->enc_ng,
               m->enc_nb, mem, (int)BL, E);
    RIB_CHECK_CUDA(cudaGetLastError());
  }
  // reco[L][B][J] = joints_embed(memory) + src^T
  RIB_MOTION_RUN(run_linear(with_res(lin(mem, sE, E, 1, E, m->out_w, m->out_b, J, reco, J, (long long)B * J), src, (long long)J * L, 1, L),
                            B, stream));
  {
    const int total = B * L * J;
    launch_pdl(motion_interp_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, stream, (const float*)reco, interp, B, L, J,
               rate);
    RIB_CHECK_CUDA(cudaGetLastError());
  }
  // y = input_embed(interp)
  RIB_MOTION_RUN(run_linear(lin(interp, sJ, J, 1, J, m->in_w, m->in_b, E, y, sE, E), B, stream));
  for (const LayerW& l : m->dec) {
    RIB_MOTION_RUN(run_linear(with_pos(with_ln(lin(y, sE, E, 1, E, l.self.in_w, l.self.in_b, 3 * E, qkv, s3E, 3 * E), l.n1g, l.n1b),
                                       tgt_pos, 2 * E), B, stream));
    RIB_MOTION_RUN(mha(qkv, qkv + E, qkv + 2 * E, tgt_mask, 0));
    RIB_MOTION_RUN(run_linear(with_res(lin(att, sE, E, 1, E, l.self.out_w, l.self.out_b, E, y, sE, E), y, sE, E, 1), B, stream));
    // cross attention: q = (norm2(y) + tgt_pos) Wq; k = (memory + src_pos) Wk; v = memory Wv
    RIB_MOTION_RUN(run_linear(with_pos(with_ln(lin(y, sE, E, 1, E, l.cross.in_w, l.cross.in_b, E, qkv, s3E, 3 * E), l.n2g, l.n2b),
                                       tgt_pos, E), B, stream));
    RIB_MOTION_RUN(run_linear(with_pos(lin(mem, sE, E, 1, E, WT{l.cross.in_w.p + E, l.cross.in_w.ld}, l.cross.in_b + E, 2 * E, qkv + E, s3E, 3 * E),
                                       src_pos, E), B, stream));
    RIB_MOTION_RUN(mha(qkv, qkv + E, qkv + 2 * E, src_mask, 0));
    RIB_MOTION_RUN(run_linear(with_res(lin(att, sE, E, 1, E, l.cross.out_w, l.cross.out_b, E, y, sE, E), y, sE, E, 1), B, stream));
    LinParams f1 = with_ln(lin(y, sE, E, 1, E, l.l1w, l.l1b, FF, ffh, sFF, FF), l.n3g, l.n3b);
    f1.act = 1;
    RIB_MOTION_RUN(run_linear(f1, B, stream));
    RIB_MOTION_RUN(run_linear(with_res(lin(ffh, sFF, FF, 1, FF, l.l2w, l.l2b, E, y, sE, E), y, sE, E, 1), B, stream));
  }
  // joints[L][B][J] = joints_embed(decoder.norm(y)) + interp
  RIB_MOTION_RUN(run_linear(with_res(with_ln(lin(y, sE, E, 1, E, m->out_w, m->out_b, J, joints, J, (long long)B * J), m->dec_ng, m->dec_nb),
                                     interp, sJ, J, 1), B, stream));
#undef RIB_MOTION_RUN
  return 0;
}

}  // namespace rib
