// Generator forward plan: weight folding/packing at creation, then one static list of kernel
// launches per (B, H, W).  Restates the dataflow of PGNR/models/generator.py:181-234 (Generator),
// :360-387 (LabelEmbedder, arch 'encoder'), :493-510 (MaskGenerator) with these fusions
// (all parity-neutral):
//   * spectral norm folded once (weight_norm.py:84-85)
//   * the 34 same-size F.interpolate calls of SPADE dropped (activation_norm.py:224)
//   * per SPADE res-block: the gamma/beta 1x1 convs of conv_block_0 and conv_block_s share one
//     GEMM over the same cond map, the instance-norm + modulation + LeakyReLU are its epilogue
//   * conv_block_1 (3x3) and the learned 1x1 shortcut accumulate into one TMEM tile (K = 9*hid + cin)
//   * instance-norm statistics come out of the producing conv's epilogue (fp64 atomics)
//   * nearest x2 up-sampling is a gather in the consumer (main branch) or the producer's store (mask net)
//   * torch.cat never materialises: producers write channel slices
//   * activations are chunk-planar 16-bit maps [B][C/8][H][W][8] (conv_gemm.cuh) so that a 3x3 conv reads one
//     halo tile per channel group instead of nine shifted tiles
#include "generator.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "conv_gemm.cuh"
#include "elementwise.cuh"

namespace rib {

int g_debug_simt = 0;

namespace {

struct View {  // chunk-planar 16-bit activation [B][Ctot/8][H][W][8] (or a channel slice of one)
  act_t* p = nullptr;  // first plane of the slice, image 0
  int B = 0, H = 0, W = 0, C = 0, Ctot = 0;
  bool parity = false;  // parity-planar layout [plane][py][px][H/2][W/2][8] (input of a stride-2 conv)
  long long bstride() const { return (long long)Ctot * H * W; }
  PlanarRef ref() const { return PlanarRef{p, bstride()}; }
};

struct GemmLayer {
  std::string name;
  int n_pad = 0, n_valid = 0, ktotal = 0;
  int cin0 = 0, taps = 0, cin1 = 0;  // padded input channels of source 0, its taps, source 1 (1x1) channels
  int BN = 0, bkc = 0, stride = 1;   // N tile, channels per group (fixes the packed K order), conv stride
  int ppc = 1;                       // sub-pixel conv: output parities per CTA (ConvGemmParams::ppc)
  act_t* w = nullptr;
  float* bias = nullptr;
  size_t w_off = 0, b_off = 0;
};

enum OpKind { OP_MEMSET, OP_PACK, OP_PACK_IMAGES, OP_GEMM, OP_IN_APPLY, OP_POOL };
enum ExtPtr { EXT_NONE = 0, EXT_LABEL, EXT_FAKE, EXT_PREV, EXT_OUT_IMG, EXT_OUT_MASK };

struct Op {
  OpKind kind;
  std::string name;
  // memset
  void* ms_ptr = nullptr;
  size_t ms_bytes = 0;
  // pack
  int ext = EXT_NONE;
  int pk_nsrc = 0, pk_ext[3] = {0, 0, 0}, pk_C[3] = {0, 0, 0}, pk_nplanes = 0;
  long long pk_bs = 0;
  act_t* pk_dst = nullptr;
  // pack_images
  act_t* pi_emb = nullptr;
  act_t* pi_mask = nullptr;
  long long pi_emb_bs = 0, pi_mask_bs = 0;
  // gemm
  ConvGemmParams g;
  int mode = 0;
  const GemmLayer* layer = nullptr;  // for re-tiling (auto-tuner)
  View in0, in1;
  bool has_in1 = false;
  // in_apply
  InApplyParams ia;
  // pool
  const act_t* pl_src = nullptr;
  act_t* pl_dst = nullptr;
  double* pl_stats = nullptr;
  long long pl_sbs = 0, pl_dbs = 0;
  int pl_H = 0, pl_W = 0, pl_C = 0;
};

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  void* take(size_t bytes) {
    off = align_up(off, 1024);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

}  // namespace

struct Generator {
  rib_gen_config cfg;
  std::map<std::string, GemmLayer> layers;
  std::unordered_map<std::string, const float*> in_affine;  // IN weight / bias device pointers (caller-owned fp32)
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  float* cat_w = nullptr;  // IN affine weight / bias of cat([down_lbl, down_img]) in one table (RIB_XF=2)
  float* cat_b = nullptr;
  // cached plan
  int pB = 0, pH = 0, pW = 0;
  void* pws = nullptr;
  size_t ws_bytes_needed = 0;
  std::vector<Op> ops;
  std::map<std::string, View> debug_views;
  bool plan_simt = false;
  // planes that are all-zero padding for the life of the plan: cleared by the first forward after a plan (re)build
  struct ZeroPlane {
    void* ptr;
    size_t pitch, width;
    int rows;
  };
  std::vector<ZeroPlane> zero_once;
  bool zero_pending = false;
};

namespace {

int nfilt(const rib_gen_config& c, int i) { return std::min(c.maxf, c.nf << i); }
int mask_nfilt(const rib_gen_config& c, int i) { return std::min(c.mask_max, c.mask_nf << i); }
int emb_ch(const rib_gen_config& c, int i) { return std::min(c.emb_max, c.emb_nf << i); }

struct BlockDef {
  std::string name;
  int cin, hid, cout, lvl;
  bool shortcut;
};

std::vector<BlockDef> res_blocks(const rib_gen_config& c) {
  std::vector<BlockDef> v;
  for (int i = 0; i <= c.n_down; ++i) {
    int ci = nfilt(c, i), co = nfilt(c, i + 1);
    v.push_back({"down_" + std::to_string(i), ci, std::min(ci, co), co, std::min(c.emb_down, i), ci != co});
  }
  int ch = nfilt(c, c.n_down + 1);
  for (int i = 0; i < c.n_res; ++i)
    v.push_back({"res_" + std::to_string(i), ch, ch, ch, std::min(c.emb_down, c.n_down + 1), false});
  for (int i = c.n_down; i >= 0; --i) {
    int ci = nfilt(c, i + 1), co = nfilt(c, i);
    v.push_back({"up_" + std::to_string(i), ci, std::min(ci, co), co, std::min(i, c.emb_down), ci != co});
  }
  return v;
}

// ---- model creation -------------------------------------------------------------------------
struct Source {
  std::unordered_map<std::string, std::pair<const float*, long long>> t;
  const float* get(const std::string& k, long long numel, int* err) const {
    auto it = t.find(k);
    if (it == t.end()) {
      set_error("state-dict entry missing: " + k);
      *err = -3;
      return nullptr;
    }
    if (it->second.second != numel) {
      set_error("state-dict entry has wrong size: " + k);
      *err = -3;
      return nullptr;
    }
    return it->second.first;
  }
};

struct PackJob {  // one reference conv placed inside a GEMM layer
  std::string layer, prefix;
  bool sn;
  int cout, cin, taps, koff, cin_pad /* padded input channels (informational) */, row_off, spade_C, spade_CT;
  bool bias_acc;
  int subpix = 0, parity = 0;  // sub-pixel form of a conv that follows nearest x2 (see PackWeightParams)
  int spade_nq = 1, spade_q = 0;  // SPADE outputs sharing an N tile / this conv's index among them
};

// Channels per SPADE tile: BN = 2 * nq * CT <= 128.  RIB_SPADE2=0 keeps the two shortcut-block outputs in separate tiles.
bool spade_pairs() {
  static const bool on = !(getenv("RIB_SPADE2") != nullptr && atoi(getenv("RIB_SPADE2")) == 0);
  return on;
}
// Low-resolution levels (cond maps with >= 256 channels, i.e. K >= 256): 256-column tiles run by CTA pairs, so that a
// cond tile is read once for twice the columns and each CTA keeps only its 128 weight rows (RIB_SPADE256=0: 128 columns).
int spade_ct(int C, int nq_tile = 1, int cond = 0) {
  static const bool wide_on = !(getenv("RIB_SPADE256") != nullptr && atoi(getenv("RIB_SPADE256")) == 0);
  // (only the paired-output tiles [gamma0|beta0|gamma_s|beta_s] x 64 channels: measured, profiles/r2o, the single-output
  //  form with 128 channels per tile loses 16 us per layer to its chunk-wise x loads while the paired form gains 13-35)
  if (wide_on && nq_tile == 2 && cond >= 256 && C >= 64 && C % 64 == 0) return 64;
  static const bool wide1_on = getenv("RIB_SPADE256_SINGLE") != nullptr && atoi(getenv("RIB_SPADE256_SINGLE")) == 1;
  if (wide_on && wide1_on && nq_tile == 1 && cond >= 256 && C >= 128 && C % 128 == 0) return 128;
  return std::min(C, nq_tile == 2 ? 32 : 64);
}

// BN = 0 selects the plain-store default min(n_pad, 128).
void add_layer(Generator* G, const std::string& name, int n_valid, int n_pad, int cin0_pad, int taps, int cin1,
               int BN = 0, int stride = 1, int ppc = 1) {
  GemmLayer L;
  L.ppc = ppc;
  L.BN = BN ? BN : std::min(n_pad, 128);
  L.stride = stride;
  L.bkc = choose_bkc(cin0_pad, cin1, taps, L.BN, stride);
  L.name = name;
  L.n_valid = n_valid;
  L.n_pad = n_pad;
  L.cin0 = cin0_pad;
  L.taps = taps;
  L.cin1 = cin1;
  L.ktotal = cin0_pad * taps + cin1;
  G->layers[name] = L;
}

}  // namespace

// The GEMM layers of the generator and the reference convs packed into each (no device memory involved: also used by
// the CPU dry run of the launch plan).
static void define_layers(Generator* G, std::vector<PackJob>& jobs) {
  const rib_gen_config& c = G->cfg;
  const int lab_pad = 32, emb_in_pad = 16, mask_in_pad = 16;

  // ref_embedding
  add_layer(G, "emb_0", emb_ch(c, 0), emb_ch(c, 0), emb_in_pad, 9, 0);
  jobs.push_back({"emb_0", "ref_embedding.conv_first.layers.conv", true, emb_ch(c, 0), 2 * c.img_nc, 9, 0, emb_in_pad, 0, 0, 0, false});
  for (int i = 0; i < c.emb_down; ++i) {
    std::string ln = "emb_" + std::to_string(i + 1);
    add_layer(G, ln, emb_ch(c, i + 1), emb_ch(c, i + 1), emb_ch(c, i), 9, 0, 0, 2);
    jobs.push_back({ln, "ref_embedding.down_" + std::to_string(i) + ".layers.conv", true, emb_ch(c, i + 1), emb_ch(c, i), 9, 0, emb_ch(c, i), 0, 0, 0, false});
  }
  // down_first and the mask net's down_lbl.0 are both 3x3 convs over the label: one GEMM with their output channels side
  // by side (N = nf + mask_nf) reads the label once and doubles the columns per tcgen05.mma, which is what bounds these
  // narrow full-resolution layers.  RIB_MERGE_LABEL=0 keeps them apart.
  static const bool merge_env = !(getenv("RIB_MERGE_LABEL") != nullptr && atoi(getenv("RIB_MERGE_LABEL")) == 0);
  const bool merge_label = merge_env && c.nf + c.mask_nf <= 128;
  if (merge_label) {
    int bn = 16;
    while (bn < c.nf + c.mask_nf) bn <<= 1;
    add_layer(G, "label3x3", c.nf + c.mask_nf, bn, lab_pad, 9, 0, bn);
    jobs.push_back({"label3x3", "down_first.layers.conv", false, c.nf, c.label_nc, 9, 0, lab_pad, 0, 0, 0, false});
    jobs.push_back({"label3x3", "flow_network_temp.down_lbl.0.layers.conv", true, c.mask_nf, c.label_nc, 9, 0, lab_pad, c.nf, 0, 0, false});
  } else {
    add_layer(G, "down_first", c.nf, c.nf, lab_pad, 9, 0);
    jobs.push_back({"down_first", "down_first.layers.conv", false, c.nf, c.label_nc, 9, 0, lab_pad, 0, 0, 0, false});
  }
  // SPADE res-blocks
  for (const BlockDef& b : res_blocks(c)) {
    const int cond = emb_ch(c, b.lvl);
    const int nq = b.shortcut ? 2 : 1;
    // conv_block_0 and conv_block_s modulate the same x over the same cond map: their [gamma|beta] columns share N
    // tiles (EPI_SPADE2), so cond and x are read once for both outputs
    const int nqt = (nq == 2 && spade_pairs()) ? 2 : 1;
    const int ctA = spade_ct(b.cin, nqt, cond);
    add_layer(G, b.name + ".spadeA", nq * 2 * b.cin, nq * 2 * b.cin, cond, 1, 0, 2 * nqt * ctA);
    jobs.push_back({b.name + ".spadeA", b.name + ".conv_block_0.layers.norm.mlps.0.0.layers.conv", false, 2 * b.cin, cond, 1, 0, cond, 0, b.cin, ctA, false, 0, 0, nqt, 0});
    if (b.shortcut)
      jobs.push_back({b.name + ".spadeA", b.name + ".conv_block_s.layers.norm.mlps.0.0.layers.conv", false, 2 * b.cin, cond, 1, 0, cond, nqt == 2 ? 0 : 2 * b.cin, b.cin, ctA, false, 0, 0, nqt, nqt == 2 ? 1 : 0});
    add_layer(G, b.name + ".conv0", b.hid, b.hid, b.cin, 9, 0);
    jobs.push_back({b.name + ".conv0", b.name + ".conv_block_0.layers.conv", true, b.hid, b.cin, 9, 0, b.cin, 0, 0, 0, false});
    add_layer(G, b.name + ".spadeB", 2 * b.hid, 2 * b.hid, cond, 1, 0, 2 * spade_ct(b.hid, 1, cond));
    jobs.push_back({b.name + ".spadeB", b.name + ".conv_block_1.layers.norm.mlps.0.0.layers.conv", false, 2 * b.hid, cond, 1, 0, cond, 0, b.hid, spade_ct(b.hid, 1, cond), false});
    add_layer(G, b.name + ".conv1", b.cout, b.cout, b.hid, 9, b.shortcut ? b.cin : 0);
    jobs.push_back({b.name + ".conv1", b.name + ".conv_block_1.layers.conv", true, b.cout, b.hid, 9, 0, b.hid, 0, 0, 0, false});
    if (b.shortcut)
      jobs.push_back({b.name + ".conv1", b.name + ".conv_block_s.layers.conv", true, b.cout, b.cin, 1, 9 * b.hid, b.cin, 0, 0, 0, true});
  }
  add_layer(G, "conv_img", c.img_nc, 16, c.nf, 9, 0);
  jobs.push_back({"conv_img", "conv_img.layers.conv", false, c.img_nc, c.nf, 9, 0, c.nf, 0, 0, 0, false});
  // mask network
  const std::string f = "flow_network_temp.";
  for (int br = 0; br < 2; ++br) {
    const std::string bn = br == 0 ? "down_lbl" : "down_img";
    const int cin = br == 0 ? c.label_nc : 3 * c.img_nc, cpad = br == 0 ? lab_pad : mask_in_pad;
    if (!(br == 0 && merge_label)) {
      add_layer(G, "mask." + bn + ".0", c.mask_nf, c.mask_nf, cpad, 9, 0);
      jobs.push_back({"mask." + bn + ".0", f + bn + ".0.layers.conv", true, c.mask_nf, cin, 9, 0, cpad, 0, 0, 0, false});
    }
    for (int i = 0; i < c.mask_down; ++i) {
      std::string ln = "mask." + bn + "." + std::to_string(i + 1);
#ifdef RIB_MASKDOWN_BN64
      // A/B build: 64-column tiles for the 128-channel level, whose weights (72 KB) then stay resident beside a deep ring of
      // halo slots (the stride-2 slots of 16 channels carry ~0.6k clocks of MMAs each, far less than a TMA round trip)
      add_layer(G, ln, mask_nfilt(c, i + 1), mask_nfilt(c, i + 1), mask_nfilt(c, i), 9, 0, mask_nfilt(c, i + 1) == 128 ? 64 : 0, 2);
#else
      add_layer(G, ln, mask_nfilt(c, i + 1), mask_nfilt(c, i + 1), mask_nfilt(c, i), 9, 0, 0, 2);
#endif
      jobs.push_back({ln, f + bn + "." + std::to_string(i + 1) + ".layers.conv", true, mask_nfilt(c, i + 1), mask_nfilt(c, i), 9, 0, mask_nfilt(c, i), 0, 0, 0, false});
    }
  }
  const int mch = mask_nfilt(c, c.mask_down);
  for (int i = 0; i < c.mask_res; ++i) {
    const int ci = i == 0 ? 2 * mch : mch;
    const std::string rn = "mask.res." + std::to_string(i), rp = f + "res_flow." + std::to_string(i);
    add_layer(G, rn + ".conv0", mch, mch, ci, 9, 0);
    jobs.push_back({rn + ".conv0", rp + ".conv_block_0.layers.conv", true, mch, ci, 9, 0, ci, 0, 0, 0, false});
    add_layer(G, rn + ".conv1", mch, mch, mch, 9, 0);
    jobs.push_back({rn + ".conv1", rp + ".conv_block_1.layers.conv", true, mch, mch, 9, 0, mch, 0, 0, 0, false});
    if (i == 0) {
      add_layer(G, rn + ".convs", mch, mch, ci, 1, 0);
      jobs.push_back({rn + ".convs", rp + ".conv_block_s.layers.conv", true, mch, ci, 1, 0, ci, 0, 0, 0, false});
    }
  }
  for (int k = 0; k < c.mask_down; ++k) {
    const int i = c.mask_down - 1 - k;
    const std::string ln = "mask.up." + std::to_string(k);
    // nn.Upsample(2) -> conv3x3 (generator.py:478-481) runs as four 2x2 convs on the low-resolution map, one per output
    // parity: N = 4 x Cout, K = 4 x Cin
    const int co = mask_nfilt(c, i), ci = mask_nfilt(c, i + 1);
    // Narrow layers (Cout < 128) can compute 128 / Cout output parities per CTA from one halo load.
    // Measured (profiles/r2b): with today's epilogue the wider tiles lose more CTAs per SM (fewer epilogue warps) than the
    // shared halo load saves (mask.up.1 191 -> 259 us, up.2 432 -> 463 us), so the form is opt-in: RIB_SUBPIX_PPC=4.
    static const int ppc_env = getenv("RIB_SUBPIX_PPC") != nullptr ? atoi(getenv("RIB_SUBPIX_PPC")) : 1;
    const int ppc = (ppc_env > 1 && co < 128 && 128 % co == 0 && co % 16 == 0) ? std::min(std::min(4, ppc_env), 128 / co) : 1;
    add_layer(G, ln, co, 4 * co, ci, 4, 0, std::min(co * ppc, 128), 1, ppc);
    for (int par = 0; par < 4; ++par)
      jobs.push_back({ln, f + "up_flow." + std::to_string(2 * k + 1) + ".layers.conv", true, co, ci, 9, 0, ci, par * co, 0, 0, false, 1, par});
  }
  add_layer(G, "mask.conv_mask", 1, 16, c.mask_nf, 9, 0);
  jobs.push_back({"mask.conv_mask", f + "conv_mask.0.layers.conv", false, 1, c.mask_nf, 9, 0, c.mask_nf, 0, 0, 0, false});
}

int generator_create(const rib_gen_config* cfg, const rib_tensor* tensors, int n, cudaStream_t stream,
                     Generator** out) {
  RIB_REQUIRE(cfg && tensors && out, "generator_create: null argument");
  const rib_gen_config& c = *cfg;
  RIB_REQUIRE(c.nf % 16 == 0 && c.emb_nf % 16 == 0 && c.mask_nf % 16 == 0, "channel counts must be multiples of 16");
  RIB_REQUIRE(c.label_nc <= 32 && c.img_nc == 3, "unsupported input channel counts");
  Source src;
  for (int i = 0; i < n; ++i) src.t[tensors[i].name] = {tensors[i].data, tensors[i].numel};

  Generator* G = new Generator();
  G->cfg = c;
  std::vector<PackJob> jobs;
  define_layers(G, jobs);
  const std::string f = "flow_network_temp.";
  const int mch = mask_nfilt(c, c.mask_down);

  // arena: packed weights + biases + one sigma scalar per job
  size_t off = 0;
  for (auto& kv : G->layers) {
    GemmLayer& L = kv.second;
    off = align_up(off, 256);
    L.w_off = off;
    off += (size_t)L.n_pad * L.ktotal * sizeof(act_t);
    off = align_up(off, 256);
    L.b_off = off;
    off += (size_t)L.n_pad * sizeof(float);
  }
  off = align_up(off, 256);
  const size_t sigma_off = off;
  off += jobs.size() * sizeof(float);
  off = align_up(off, 256);
  const size_t cat_off = off;
  off += (size_t)4 * mask_nfilt(c, c.mask_down) * sizeof(float);
  off = align_up(off, 256);
  const size_t sn_scratch_off = off;   // per-row partial sums of the spectral-norm sigma (re-used by every conv)
  int max_cout = 1;
  for (const PackJob& pj : jobs) max_cout = std::max(max_cout, pj.cout);
  off += (size_t)max_cout * sizeof(double);
  G->arena_bytes = off;
  cudaError_t ce = cudaMalloc(&G->arena, G->arena_bytes);
  if (ce != cudaSuccess) {
    set_error(std::string("cudaMalloc(weights arena): ") + cudaGetErrorString(ce));
    delete G;
    return -2;
  }
  int rc = 0;
  auto fail = [&](int code) {
    cudaFree(G->arena);
    delete G;
    return code;
  };
  if (cudaMemsetAsync(G->arena, 0, G->arena_bytes, stream) != cudaSuccess) return fail(-2);
  for (auto& kv : G->layers) {
    kv.second.w = reinterpret_cast<act_t*>(G->arena + kv.second.w_off);
    kv.second.bias = reinterpret_cast<float*>(G->arena + kv.second.b_off);
  }
  float* sigmas = reinterpret_cast<float*>(G->arena + sigma_off);
  for (size_t j = 0; j < jobs.size(); ++j) {
    const PackJob& pj = jobs[j];
    GemmLayer& L = G->layers[pj.layer];
    int err = 0;
    const long long wn = (long long)pj.cout * pj.cin * pj.taps;
    const float* w = src.get(pj.prefix + (pj.sn ? ".weight_orig" : ".weight"), wn, &err);
    const float* bias = src.get(pj.prefix + ".bias", pj.cout, &err);
    const float* sig = nullptr;
    if (pj.sn && !err) {
      const float* u = src.get(pj.prefix + ".weight_u", pj.cout, &err);
      const float* v = src.get(pj.prefix + ".weight_v", (long long)pj.cin * pj.taps, &err);
      if (!err) {
        rc = launch_sn_sigma_inv(w, u, v, pj.cout, pj.cin * pj.taps, sigmas + j,
                                 reinterpret_cast<double*>(G->arena + sn_scratch_off), stream);
        if (rc) return fail(rc);
        sig = sigmas + j;
      }
    }
    if (err) return fail(err);
    PackWeightParams pp;
    pp.w = w;
    pp.bias = bias;
    pp.sigma_inv = sig;
    pp.Cout = pj.cout;
    pp.Cin = pj.cin;
    pp.taps = pj.taps;
    pp.dst = L.w;
    pp.bias_dst = L.bias;
    pp.ktotal = L.ktotal;
    pp.koff = pj.koff;
    pp.bkc = L.bkc;
    pp.row_off = pj.row_off;
    pp.spade_C = pj.spade_C;
    pp.spade_CT = pj.spade_CT;
    pp.bias_accumulate = pj.bias_acc ? 1 : 0;
    pp.subpix = pj.subpix;
    pp.subpix_parity = pj.parity;
    pp.spade_nq = pj.spade_nq;
    pp.spade_q = pj.spade_q;
    rc = launch_pack_weight(pp, stream);
    if (rc) return fail(rc);
  }
  // instance-norm affine parameters stay in the caller's fp32 tensors
  {
    int err = 0;
    auto reg = [&](const std::string& prefix, int ch) {
      G->in_affine[prefix + ".weight"] = src.get(prefix + ".weight", ch, &err);
      G->in_affine[prefix + ".bias"] = src.get(prefix + ".bias", ch, &err);
    };
    for (int br = 0; br < 2; ++br) {
      const std::string bn = br == 0 ? "down_lbl" : "down_img";
      for (int i = 0; i <= c.mask_down; ++i) reg(f + bn + "." + std::to_string(i) + ".layers.norm", mask_nfilt(c, i));
    }
    for (int i = 0; i < c.mask_res; ++i) {
      reg(f + "res_flow." + std::to_string(i) + ".conv_block_0.layers.norm", mch);
      reg(f + "res_flow." + std::to_string(i) + ".conv_block_1.layers.norm", mch);
      if (i == 0) reg(f + "res_flow.0.conv_block_s.layers.norm", mch);
    }
    for (int k = 0; k < c.mask_down; ++k)
      reg(f + "up_flow." + std::to_string(2 * k + 1) + ".layers.norm", mask_nfilt(c, c.mask_down - 1 - k));
    if (err) return fail(err);
    G->cat_w = reinterpret_cast<float*>(G->arena + cat_off);
    G->cat_b = G->cat_w + 2 * mch;
    for (int br = 0; br < 2; ++br) {
      const std::string pre = f + (br == 0 ? "down_lbl." : "down_img.") + std::to_string(c.mask_down) + ".layers.norm";
      if (cudaMemcpyAsync(G->cat_w + br * mch, G->in_affine.at(pre + ".weight"), mch * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
          cudaMemcpyAsync(G->cat_b + br * mch, G->in_affine.at(pre + ".bias"), mch * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return fail(-2);
    }
  }
  ce = cudaStreamSynchronize(stream);
  if (ce != cudaSuccess) {
    set_error(std::string("generator_create: ") + cudaGetErrorString(ce));
    return fail(-2);
  }
  *out = G;
  return 0;
}

void generator_destroy(Generator* G) {
  if (!G) return;
  if (G->arena) cudaFree(G->arena);
  delete G;
}

// ---- plan construction ----------------------------------------------------------------------
namespace {

struct PlanBuilder {
  Generator* G;
  Bump ws;
  int B;
  bool real;  // false: size query only
  std::vector<Op> ops;
  std::map<std::string, View> views;
  std::vector<std::pair<double*, size_t>> stats_bufs;
  float* final_scratch = nullptr;  // stand-in for the caller's fp32 outputs while the auto-tuner times EPI_FINAL launches
  int rc = 0;

  PlanBuilder(Generator* g, void* base, int b) : G(g), ws(base), B(b), real(base != nullptr) {}
  // CPU dry run: views get (never dereferenced) non-null addresses so that every "is this pointer set" decision of the
  // plan, and with it the tuning-table key of each launch, is the one of a real plan; no tensor maps are built.
  PlanBuilder(Generator* g, int b, bool /*dry*/) : G(g), ws(reinterpret_cast<void*>((uintptr_t)1 << 30)), B(b), real(false) {}

  View alloc(const std::string& name, int H, int W, int C) {
    View v;
    v.B = B;
    v.H = H;
    v.W = W;
    v.C = C;
    v.Ctot = C;
    v.p = static_cast<act_t*>(ws.take((size_t)B * H * W * C * sizeof(act_t)));
    views[name] = v;
    return v;
  }
  static View slice(const View& v, int c_off, int C) {
    View s = v;
    s.p = v.p ? v.p + (size_t)(c_off / 8) * v.H * v.W * 8 : nullptr;  // c_off is a multiple of 8
    s.C = C;
    return s;
  }
  double* alloc_stats(int C) {  // carved from one arena that is zeroed by a single memset
    double* p = static_cast<double*>(ws.take((size_t)B * C * 2 * sizeof(double)));
    stats_bufs.push_back({p, (size_t)B * C * 2 * sizeof(double)});
    return p;
  }

  // Fills the geometry / tensor maps of one implicit-GEMM launch.
  ConvGemmParams gemm_common(const GemmLayer& L, const View& in0, const View* in1, int stride, int Hout, int Wout,
                             int BN) {
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    if (in0.C != L.cin0 || (in1 ? in1->C : 0) != L.cin1 || BN != L.BN || stride != L.stride || in0.C != in0.Ctot ||
        (in1 && in1->C != in1->Ctot) || (in0.parity && stride != 2) || (in1 && in1->parity)) {
      set_error("plan: layer/view channel mismatch in " + L.name);
      rc = -4;
      return p;
    }
    int r = conv_gemm_configure(&p, B, Hout, Wout, L.cin0, L.cin1, L.taps, stride, BN, L.n_pad, nullptr, L.ppc);
    if (r || p.BKc != L.bkc) {
      if (!r) set_error("plan: K ordering mismatch in " + L.name);
      rc = r ? r : -4;
      return p;
    }
    p.debug_simt = g_debug_simt;
    p.s2_parity = in0.parity ? 1 : 0;
    p.src0 = in0.ref();
    p.Hin = in0.H;
    p.Win = in0.W;
    if (in1) p.src1 = in1->ref();
    p.wpk = L.w;
    p.ktotal = L.ktotal;
    p.bias = L.bias;
    p.n_valid = L.n_valid;
    p.stats_ld = L.n_valid;
    p.eps = 1e-5f;
    if (real && rc == 0) rc = make_maps(&p, L, in0, in1);
    last_layer = &L;
    last_in0 = in0;
    last_has_in1 = in1 != nullptr;
    if (in1) last_in1 = *in1;
    return p;
  }

  // inputs of the most recent gemm_common() call (copied into the Op by push_gemm / conv_final)
  const GemmLayer* last_layer = nullptr;
  View last_in0, last_in1;
  bool last_has_in1 = false;

  // Tensor maps for the tiling currently in p (they depend on the halo box, i.e. on MT).
  int make_maps(ConvGemmParams* p, const GemmLayer& L, const View& in0, const View* in1) {
    int r = 0;
    const int halo_h = (int)(p->lbo / 16u) / p->halo_w;
    if (p->stride == 1) {
      r = make_tmap_act_s1(&p->amap[0], in0.p, in0.C, in0.W, in0.H, B, in0.bstride(), p->BKc, p->halo_w, halo_h);
      if (!r && in1)
        r = make_tmap_act_s1(&p->amap[1], in1->p, in1->C, in1->W, in1->H, B, in1->bstride(), p->BKc, p->halo_w, halo_h);
    } else {
      for (int py = 0; py < 2 && !r; ++py)
        for (int px = 0; px < 2 && !r; ++px)
          r = in0.parity ? make_tmap_act_s2(&p->amap[py * 2 + px], in0.p, in0.C, in0.W, in0.H, B, in0.bstride(), py, px,
                                            p->BKc, p->halo_w, halo_h)
                         : make_tmap_act_s2_strided(&p->amap[py * 2 + px], in0.p, in0.C, in0.W, in0.H, B, in0.bstride(),
                                                    py, px, p->BKc, p->halo_w, halo_h);
    }
    if (!r) r = make_tmap_w(&p->bmap, L.w, L.ktotal, L.n_pad, p->BKc, p->pair ? p->BN / 2 : p->BN, L.taps);
    return r;
  }

  // The same launch with another tiling: configure() rewrites only the geometry fields, the epilogue stays.
  int retile(const Op& op, const ConvTune& t, ConvGemmParams* out) {
    ConvGemmParams p = op.g;
    const GemmLayer& L = *op.layer;
    int r = conv_gemm_configure(&p, B, p.H, p.W, L.cin0, L.cin1, L.taps, L.stride, p.BN, L.n_pad, &t, L.ppc);
    if (r || p.BKc != L.bkc) return r ? r : -4;
    if (p.a_gather && (!p.b_resident || p.pair)) p.a_gather = 0;   // streamed weights / CTA pairs: the strided TMA views
    if (real) r = make_maps(&p, L, op.in0, op.has_in1 ? &op.in1 : nullptr);   // (the CPU dry run has no tensors to map)
    if (r) return r;
    *out = p;
    return 0;
  }

  void push_gemm(const ConvGemmParams& p, int mode, const std::string& name) {
    Op op;
    op.kind = OP_GEMM;
    op.name = name;
    op.g = p;
    op.mode = mode;
    op.layer = last_layer;
    op.in0 = last_in0;
    op.in1 = last_in1;
    op.has_in1 = last_has_in1;
    ops.push_back(op);
  }

  // The consumer-side form of the N-A of a C-N-A block: in0 holds the producer's RAW output and the conv kernel
  // normalises (+ LeakyReLU) every halo tile in shared memory (ConvGemmParams::xf_*).
  struct Xf {
    const double* stats = nullptr;
    const float* w = nullptr;
    const float* b = nullptr;
    int act = 0;
  };

  // conv (+bias, optional residual/second source) -> 16-bit planar store, optional statistics.  out.parity: the output
  // is written parity-planar (only a stride-2 conv reads it).  stats_ld: channels per image of the statistics buffer
  // when this layer's statistics are a slice of a wider one (0 = its own).
  void conv_store(const std::string& lname, const View& in0, const View* in1, int stride, const View& out,
                  double* stats, int act, const View* res, const View* out2 = nullptr, const Xf* xf = nullptr,
                  int stats_ld = 0, bool res_ups = false) {
    const GemmLayer& L = G->layers.at(lname);
    ConvGemmParams p = gemm_common(L, in0, in1, stride, out.H, out.W, L.BN);
    p.out = out.ref();
    p.out_parity = out.parity ? 1 : 0;
    if (stats_ld) p.stats_ld = stats_ld;
    if (xf) {
      p.xf_stats = xf->stats;
      p.xf_w = xf->w;
      p.xf_b = xf->b;
      p.xf_act = xf->act;
    }
    p.has_out2 = out2 ? 1 : 0;
    if (out2) p.out2 = out2->ref();
    // stride 2 from a normal-layout map (emb_1 reads cond_0, which its stride-1 consumers need in normal layout): the
    // kernel's extra warps gather the parity tiles with cp.async instead of 16-byte strided TMA elements
    // (opt-in: measured slower than the TMA views, 684 vs 555 us for emb_1 - profiles/r2k - the layer is bound by the
    //  bytes in flight, not by the TMA element rate)
    static const bool gather_on = getenv("RIB_GATHER") != nullptr && atoi(getenv("RIB_GATHER")) == 1;
    if (gather_on && stride == 2 && !in0.parity && !xf && p.b_resident && !p.pair && (L.BN == 64 || L.BN == 128)) p.a_gather = 1;
    p.stats = stats;
    p.act = act;
    p.has_res = res ? 1 : 0;
    if (res) {
      p.res = res->ref();
      p.res_ups = res_ups ? 1 : 0;
      const int rh = res_ups ? out.H / 2 : out.H, rw = res_ups ? out.W / 2 : out.W;
      if (res->H != rh || res->W != rw || res->C != out.C || res->parity) {
        set_error("plan: residual shape mismatch in " + lname);
        rc = -4;
      }
    }
    push_gemm(p, EPI_STORE, lname);
  }

  // Two convs over in0 merged along N: columns [0, a.C) -> a (+ stats_a), the rest -> b (+ stats_b); no activation.
  void conv_store_merged(const std::string& lname, const View& in0, const View& a, double* stats_a, const View& b,
                         double* stats_b) {
    const GemmLayer& L = G->layers.at(lname);
    if (a.C + b.C != L.n_valid || a.H != b.H || a.W != b.W || a.parity) {
      set_error("plan: bad merged conv " + lname);
      rc = -4;
    }
    ConvGemmParams p = gemm_common(L, in0, nullptr, 1, a.H, a.W, L.BN);
    p.out = a.ref();
    p.stats = stats_a;
    p.stats_ld = a.C;
    p.seg_cols = a.C;
    p.out_b = b.ref();
    p.out_b_parity = b.parity ? 1 : 0;
    p.stats_b = stats_b;
    p.stats_b_ld = b.C;
    p.act = ACT_NONE;
    push_gemm(p, EPI_STORE, lname);
  }

  // conv3x3(nearest_x2(in0)) + bias as a sub-pixel conv: `out` is the parity-planar (2H, 2W) raw map
  void conv_subpix(const std::string& lname, const View& in0, const View& out, double* stats) {
    const GemmLayer& L = G->layers.at(lname);
    if (out.H != 2 * in0.H || out.W != 2 * in0.W || !out.parity || out.C != L.n_valid) {
      set_error("plan: bad sub-pixel conv output in " + lname);
      rc = -4;
    }
    ConvGemmParams p = gemm_common(L, in0, nullptr, 1, in0.H, in0.W, L.BN);
    p.out = out.ref();
    p.stats = stats;
    p.act = ACT_NONE;
    push_gemm(p, EPI_STORE, lname);
  }

  // [gamma|beta] = conv1x1(cond); out_q = act_q((x - mean) * rstd * (1 + gamma) + beta)
  void spade(const std::string& lname, const View& cond, const View& x, const double* xstats, bool ups, int nq,
             const View* outs, const int* acts) {
    const GemmLayer& L = G->layers.at(lname);
    const int nqt = (nq == 2 && spade_pairs()) ? 2 : 1;
    const int CT = spade_ct(x.C, nqt, cond.C);
    ConvGemmParams p = gemm_common(L, cond, nullptr, 1, cond.H, cond.W, 2 * nqt * CT);
    p.x = x.ref();
    p.Hx = x.H;
    p.Wx = x.W;
    p.ups = ups ? 1 : 0;
    p.xstats = xstats;
    p.C = x.C;
    p.CT = CT;
    for (int q = 0; q < nq; ++q) {
      p.outq[q] = outs[q].ref();
      p.actq[q] = acts[q];
    }
    push_gemm(p, nqt == 2 ? EPI_SPADE2 : EPI_SPADE, lname);
  }

  void conv_final(const std::string& lname, const View& in0, int act, int ext, const View* copy, int copy_coff) {
    const GemmLayer& L = G->layers.at(lname);
    ConvGemmParams p = gemm_common(L, in0, nullptr, 1, in0.H, in0.W, 16);
    p.act = act;
    p.has_out_act = copy ? 1 : 0;
    if (copy) p.out_act = copy->ref();
    p.out_act_coff = copy_coff;
    Op op;
    op.kind = OP_GEMM;
    op.name = lname;
    op.g = p;
    op.mode = EPI_FINAL;
    op.ext = ext;
    op.layer = last_layer;
    op.in0 = last_in0;
    op.has_in1 = false;
    ops.push_back(op);
  }

  void in_apply(const View& a, const double* astats, const std::string& aprefix, const View* b, const double* bstats,
                const std::string& bprefix, const View& out, int act, bool ups) {
    if (out.parity && ups) {
      set_error("plan: in_apply cannot up-sample into a parity-planar map");
      rc = -4;
    }
    Op op;
    op.kind = OP_IN_APPLY;
    InApplyParams& p = op.ia;
    memset(&p, 0, sizeof(p));
    p.a = a.p;
    p.a_bs = a.bstride();
    p.astats = astats;
    p.aw = G->in_affine.at(aprefix + ".weight");
    p.ab = G->in_affine.at(aprefix + ".bias");
    if (b) {
      p.b = b->p;
      p.b_bs = b->bstride();
      p.bstats = bstats;
      if (bstats) {
        p.bw = G->in_affine.at(bprefix + ".weight");
        p.bb = G->in_affine.at(bprefix + ".bias");
      }
    }
    p.out = out.p;
    p.o_bs = out.bstride();
    p.B = B;
    p.H = a.H;
    p.W = a.W;
    p.C = a.C;
    p.act = act;
    p.ups = ups ? 1 : 0;
    p.in_parity = a.parity ? 1 : 0;
    p.out_parity = out.parity ? 1 : 0;
    p.eps = 1e-5f;
    ops.push_back(op);
  }

  void pool(const View& src, const View& dst, double* stats) {
    Op op;
    op.kind = OP_POOL;
    op.pl_src = src.p;
    op.pl_sbs = src.bstride();
    op.pl_dst = dst.p;
    op.pl_dbs = dst.bstride();
    op.pl_stats = stats;
    op.pl_H = src.H;
    op.pl_W = src.W;
    op.pl_C = src.C;
    ops.push_back(op);
  }

  // dst channels [0, sum C) <- cat(sources); every plane of dst is written (zero padding included)
  void pack(int nsrc, const int* exts, const int* Cs, const View& dst) {
    Op op;
    op.kind = OP_PACK;
    op.pk_nsrc = nsrc;
    for (int i = 0; i < nsrc; ++i) {
      op.pk_ext[i] = exts[i];
      op.pk_C[i] = Cs[i];
    }
    op.pk_dst = dst.p;
    op.pk_bs = dst.bstride();
    op.pk_nplanes = dst.Ctot / 8;
    ops.push_back(op);
  }
};

}  // namespace

// ---- plan-time auto-tuner ---------------------------------------------------------------------
// Which tiling is fastest for a layer (stacked sub-tiles or not, several ~100 KB CTAs per SM or one big one, weights
// resident or streamed) depends on how its halo ring, TMEM footprint and epilogue balance out; the measurements in
// profiles/r1f_autotune_candidates.txt show no simple rule.  So the first plan for a given launch shape times the
// candidates on the real buffers (CUDA events, best of 3) and every later plan of this process re-uses the choice.
// RIB_AUTOTUNE=0 keeps the default heuristics of conv_gemm_configure().
namespace {

std::mutex g_tune_mu;
std::map<std::string, ConvTune> g_tune_cache;
std::string g_tune_log;

bool autotune_enabled() {
  static const bool on = !(getenv("RIB_AUTOTUNE") != nullptr && atoi(getenv("RIB_AUTOTUNE")) == 0);
  return on;
}

std::string tune_key(const Op& op) {
  const ConvGemmParams& p = op.g;
  const GemmLayer& L = *op.layer;
  char buf[256];
  snprintf(buf, sizeof(buf), "m%d B%d H%d W%d c%d+%d t%d s%d BN%d N%d r%d o%d st%d u%d q%d par%d", op.mode, p.B, p.H, p.W,
           L.cin0, L.cin1, L.taps, L.stride, p.BN, L.n_pad, p.has_res + 2 * p.res_ups, p.has_out2, p.stats != nullptr, p.ups,
           (op.mode == EPI_SPADE || op.mode == EPI_SPADE2) ? p.n_tiles * p.BN / (2 * p.C) : 0, p.s2_parity + 2 * (p.xf_stats != nullptr) + 4 * p.out_parity + 8 * p.a_gather);
  return buf;
}

float time_gemm(const ConvGemmParams& p, int mode, cudaStream_t s) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.f;
  float best = -1.f;
  if (launch_conv_gemm(p, mode, s) == 0) {  // warm-up (also sets the kernel's shared-memory attribute)
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0, s);
      if (launch_conv_gemm(p, mode, s) != 0) break;
      cudaEventRecord(e1, s);
      if (cudaEventSynchronize(e1) != cudaSuccess) break;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (best < 0.f || ms < best) best = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

bool same_tiling(const ConvGemmParams& a, const ConvGemmParams& b) {
  return a.MT == b.MT && a.a_ring == b.a_ring && a.g_slot_bytes == b.g_slot_bytes && a.b_resident == b.b_resident &&
         a.pair == b.pair;
}

void autotune(PlanBuilder& pb, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  for (Op& op : pb.ops) {
    if (op.kind != OP_GEMM || op.layer == nullptr) continue;
    if (op.mode == EPI_FINAL) op.g.out_f32 = pb.final_scratch;
    const std::string key = tune_key(op);
    auto it = g_tune_cache.find(key);
    if (it != g_tune_cache.end()) {
      if (it->second.mt != 0 || it->second.policy != 0 || it->second.ring != 0) {
        ConvGemmParams q;
        if (pb.retile(op, it->second, &q) == 0) op.g = q;
      }
      continue;
    }
    const float t_def = time_gemm(op.g, op.mode, stream);
    ConvTune best_tune;  // zeros = keep the default
    ConvGemmParams best_p = op.g;
    float best_t = t_def;
    std::vector<ConvGemmParams> tried{op.g};
    char line[160];
    snprintf(line, sizeof(line), "%-20s default MT%d ring%d res%d %.1f us |", op.name.c_str(), op.g.MT, op.g.a_ring,
             op.g.b_resident, t_def * 1e3f);
    (void)0;
    std::string log = line;
    if (t_def > 0.f) {
      static const bool pairs_on = !(getenv("RIB_PAIRS") != nullptr && atoi(getenv("RIB_PAIRS")) == 0);
      for (int cand = 0; cand < (pairs_on ? 12 : 8); ++cand) {
        {
          // candidates 0-5: policies 1-3 x MT 1/2; 6-7: one CTA per SM with a ring of up to 8 slots; 8-9: CTA pairs with
          // streamed weights; 10-11: CTA pairs with resident half-tiles of the weights
          const int policy = cand < 6 ? cand / 2 + 1 : (cand < 8 ? 2 : (cand < 10 ? 4 : 5)), mt = cand % 2 + 1;
          ConvTune t;
          t.mt = mt;
          t.policy = policy;
          t.ring = (cand < 6 || cand >= 8) ? 0 : 8;
          ConvGemmParams q;
          if (pb.retile(op, t, &q) != 0) continue;
          bool dup = false;
          for (const ConvGemmParams& o : tried) dup = dup || same_tiling(o, q);
          if (dup) continue;
          tried.push_back(q);
          const float tq = time_gemm(q, op.mode, stream);
          snprintf(line, sizeof(line), " p%d MT%d ring%d %.1f", policy, mt, q.a_ring, tq * 1e3f);
          log += line;
          if (tq > 0.f && tq < best_t) {
            best_t = tq;
            best_tune = t;
            best_p = q;
          }
        }
      }
    }
    if (best_t >= 0.97f * t_def) {  // not worth leaving the default (and keeps the choice stable run to run)
      best_tune = ConvTune();
      best_p = op.g;
    }
    snprintf(line, sizeof(line), " -> p%d MT%d r%d", best_tune.policy, best_tune.mt, best_tune.ring);
    g_tune_log += log + line + "\n";
    g_tune_cache[key] = best_tune;
    op.g = best_p;
  }
  for (Op& op : pb.ops)
    if (op.kind == OP_GEMM && op.mode == EPI_FINAL) op.g.out_f32 = nullptr;   // bound to the caller's tensors per forward
  set_error("");  // candidates that did not fit left messages behind
}

}  // namespace

// Tuning table as text, one line per launch shape: "<key>\t<mt>\t<policy>".  Importing a table (before the first
// plan) makes the choices reproducible across processes and ranks and skips the timing runs for known shapes.
int generator_tune_export(char* buf, long long cap) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  std::string out;
  for (const auto& kv : g_tune_cache)
    out += kv.first + "\t" + std::to_string(kv.second.mt) + "\t" + std::to_string(kv.second.policy) + "\t" +
           std::to_string(kv.second.ring) + "\n";
  if ((long long)out.size() + 1 > cap) return -1;
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int generator_tune_import(const char* text) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  int n = 0;
  const char* p = text;
  while (*p) {
    const char* e = strchr(p, '\n');
    std::string line = e ? std::string(p, e - p) : std::string(p);
    p = e ? e + 1 : p + line.size();
    const size_t t1 = line.find('\t'), t2 = t1 == std::string::npos ? t1 : line.find('\t', t1 + 1);
    if (t2 == std::string::npos) continue;
    ConvTune t;
    t.mt = atoi(line.substr(t1 + 1, t2 - t1 - 1).c_str());
    t.policy = atoi(line.substr(t2 + 1).c_str());           // (stops at the next tab)
    const size_t t3 = line.find('\t', t2 + 1);
    t.ring = t3 == std::string::npos ? 0 : atoi(line.substr(t3 + 1).c_str());
    if (t.mt < 0 || t.mt > 2 || t.policy < 0 || t.policy > 5 || t.ring < 0 || t.ring > 8) continue;
    g_tune_cache[line.substr(0, t1)] = t;
    ++n;
  }
  return n;
}

// CPU dry run: the tilings of the imported tuning table without any timing (cache hits only).
static void apply_tune_table(PlanBuilder& pb) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  for (Op& op : pb.ops) {
    if (op.kind != OP_GEMM || op.layer == nullptr) continue;
    auto it = g_tune_cache.find(tune_key(op));
    if (it == g_tune_cache.end() || (it->second.mt == 0 && it->second.policy == 0 && it->second.ring == 0)) continue;
    ConvGemmParams q;
    if (pb.retile(op, it->second, &q) == 0) op.g = q;
  }
  set_error("");
}

int generator_tune_log(char* buf, long long cap) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  if ((long long)g_tune_log.size() + 1 > cap) return -1;
  memcpy(buf, g_tune_log.c_str(), g_tune_log.size() + 1);
  return 0;
}

static int build_plan(Generator* G, int B, int H, int W, void* wsbase, size_t* bytes_out, cudaStream_t stream,
                      std::vector<Op>* dry_ops = nullptr) {
  const rib_gen_config& c = G->cfg;
  RIB_REQUIRE(B >= 1 && H >= 16 && W >= 16, "forward: bad batch or image size");
  const int max_down = std::max(std::max(c.n_down, c.emb_down), c.mask_down);
  RIB_REQUIRE(H % (1 << max_down) == 0 && W % (1 << max_down) == 0,
              "forward: H and W must be multiples of " + std::to_string(1 << max_down));
  PlanBuilder pb = (wsbase == nullptr && dry_ops != nullptr) ? PlanBuilder(G, B, true) : PlanBuilder(G, wsbase, B);
  const std::string f = "flow_network_temp.";

  // statistics arena first (zeroed by one memset per forward)
  Op ms;
  ms.kind = OP_MEMSET;
  pb.ops.push_back(ms);
  const size_t stats_begin = align_up(pb.ws.off, 1024);

  // -- statistics buffers are allocated lazily below; remember the arena start --
  std::vector<BlockDef> blocks = res_blocks(c);

  // Input staging (padding channels are zeroed once, when the plan is built)
  View lab = pb.alloc("label", H, W, 32);
  View emb_in = pb.alloc("emb_in", H, W, 16);
  View mask_in = pb.alloc("mask_in", H, W, 16);
  // All statistics live in one contiguous region.
  const size_t stats_region_begin = align_up(pb.ws.off, 1024);
  (void)stats_begin;
  std::vector<double*> blk_xstats;
  // main-branch statistics: x0, per block: hidden, out (or pooled)
  double* st_x0 = pb.alloc_stats(c.nf);
  struct BlkStats {
    double* hid;
    double* out;
    double* pooled;
  };
  std::vector<BlkStats> bst(blocks.size());
  for (size_t i = 0; i < blocks.size(); ++i) {
    bst[i].hid = pb.alloc_stats(blocks[i].hid);
    bst[i].out = pb.alloc_stats(blocks[i].cout);
    bst[i].pooled = pb.alloc_stats(blocks[i].cout);
  }
  // mask-network statistics
  std::map<std::string, double*> mst;
  for (int br = 0; br < 2; ++br) {
    const std::string bn = br == 0 ? "down_lbl" : "down_img";
    for (int i = 0; i <= c.mask_down; ++i) mst[bn + std::to_string(i)] = pb.alloc_stats(mask_nfilt(c, i));
  }
  const int mch = mask_nfilt(c, c.mask_down);
  double* st_cat = pb.alloc_stats(2 * mch);   // statistics of both branches' last level, in cat order (RIB_XF=2)
  for (int i = 0; i < c.mask_res; ++i) {
    mst["res" + std::to_string(i) + "c0"] = pb.alloc_stats(mch);
    mst["res" + std::to_string(i) + "c1"] = pb.alloc_stats(mch);
    if (i == 0) mst["res0cs"] = pb.alloc_stats(mch);
  }
  for (int k = 0; k < c.mask_down; ++k) mst["up" + std::to_string(k)] = pb.alloc_stats(mask_nfilt(c, c.mask_down - 1 - k));
  const size_t stats_region_end = pb.ws.off;
  pb.ops[0].ms_ptr = wsbase ? static_cast<uint8_t*>(wsbase) + stats_region_begin : nullptr;
  pb.ops[0].ms_bytes = stats_region_end - stats_region_begin;

  // -- inputs --
  {
    const int e_lab[1] = {EXT_LABEL}, c_lab[1] = {c.label_nc};
    pb.pack(1, e_lab, c_lab, lab);
    // cat([img_fake, img_prev]) (generator.py:197) and cat([img_prev, img_fake, img_final]) (:232; img_final comes from
    // conv_img's epilogue) in one pass over the two images; channels 8..15 of both are padding (mask_in's channel 8 is
    // img_final's third channel) and are zeroed once per plan
    Op op;
    op.kind = OP_PACK_IMAGES;
    op.pi_emb = emb_in.p;
    op.pi_emb_bs = emb_in.bstride();
    op.pi_mask = mask_in.p;
    op.pi_mask_bs = mask_in.bstride();
    pb.ops.push_back(op);
  }

  // -- ref_embedding: 5 convs + LeakyReLU (generator.py:360-377) --
  std::vector<View> cond;
  {
    View prev = emb_in;
    for (int i = 0; i <= c.emb_down; ++i) {
      const int h = H >> i, w = W >> i;
      View o = pb.alloc("cond_" + std::to_string(i), h, w, emb_ch(c, i));
      // cond_1..3 are also written as a parity-planar copy for the next (stride-2) embedder conv: its four parity
      // tiles become dense TMA boxes.  cond_0 is not: emb_0 is bound by its epilogue and the copy would add 1 GB of
      // stores, more than emb_1 saves (measured), so emb_1 gathers strided parity views of cond_0.
      View o2;
      const bool dual = i >= 1 && i < c.emb_down;
      if (dual) {
        o2 = pb.alloc("cond_" + std::to_string(i) + ".s2", h, w, emb_ch(c, i));
        o2.parity = true;
      }
      pb.conv_store("emb_" + std::to_string(i), prev, nullptr, i == 0 ? 1 : 2, o, nullptr, ACT_LRELU, nullptr,
                    dual ? &o2 : nullptr);
      prev = dual ? o2 : o;
      cond.push_back(o);
    }
  }
  // -- main branch --
  View x = pb.alloc("down_first", H, W, c.nf);
  const bool merge_label = G->layers.count("label3x3") != 0;
  static const int xf_mode = getenv("RIB_XF") ? atoi(getenv("RIB_XF")) : 1;
  View lbl0_raw;   // merged launch: the mask net's down_lbl.0 raw output comes out of the same GEMM
  if (merge_label) {
    lbl0_raw = pb.alloc("mask.down_lbl.0.raw", H, W, c.mask_nf);
    if (c.mask_down > 0 && xf_mode >= 1) lbl0_raw.parity = true;
    pb.conv_store_merged("label3x3", lab, x, st_x0, lbl0_raw, mst["down_lbl0"]);
  } else {
    pb.conv_store("down_first", lab, nullptr, 1, x, st_x0, ACT_NONE, nullptr);
  }
  const double* xst = st_x0;
  bool x_ups = false;
  int lvl_h = H, lvl_w = W;  // resolution at which the current block runs
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    const BlockDef& b = blocks[bi];
    const bool is_down = b.name.compare(0, 5, "down_") == 0;
    const bool is_up = b.name.compare(0, 3, "up_") == 0;
    const int idx = std::stoi(b.name.substr(b.name.find('_') + 1));
    const View& cd = cond[b.lvl];
    if (cd.H != lvl_h || cd.W != lvl_w) {
      set_error("plan: cond map resolution mismatch at " + b.name);
      return -4;
    }
    View a0 = pb.alloc(b.name + ".a0", lvl_h, lvl_w, b.cin);
    View as;
    View outsA[2] = {a0, a0};
    int actsA[2] = {ACT_LRELU, ACT_NONE};
    if (b.shortcut) {
      as = pb.alloc(b.name + ".as", lvl_h, lvl_w, b.cin);
      outsA[1] = as;
    }
    pb.spade(b.name + ".spadeA", cd, x, xst, x_ups, b.shortcut ? 2 : 1, outsA, actsA);
    View hbuf = pb.alloc(b.name + ".h", lvl_h, lvl_w, b.hid);
    pb.conv_store(b.name + ".conv0", a0, nullptr, 1, hbuf, bst[bi].hid, ACT_NONE, nullptr);
    View a1 = pb.alloc(b.name + ".a1", lvl_h, lvl_w, b.hid);
    View outsB[1] = {a1};
    int actsB[1] = {ACT_LRELU};
    pb.spade(b.name + ".spadeB", cd, hbuf, bst[bi].hid, false, 1, outsB, actsB);
    View o = pb.alloc(b.name, lvl_h, lvl_w, b.cout);
    const bool last = (bi + 1 == blocks.size());
    const bool pooled_next = is_down && idx != c.n_down;
    double* ost = (last || pooled_next) ? nullptr : bst[bi].out;
    // up_0 feeds conv_img through a LeakyReLU (conv_img order 'AC', generator.py:114-116)
    pb.conv_store(b.name + ".conv1", a1, b.shortcut ? &as : nullptr, 1, o, ost, last ? ACT_LRELU : ACT_NONE,
                  b.shortcut ? nullptr : &x, nullptr, nullptr, 0, !b.shortcut && x_ups);
    if (pooled_next) {  // AvgPool2d(3, 2, 1)  generator.py:207-208
      lvl_h /= 2;
      lvl_w /= 2;
      View pl = pb.alloc(b.name + ".pool", lvl_h, lvl_w, b.cout);
      pb.pool(o, pl, bst[bi].pooled);
      x = pl;
      xst = bst[bi].pooled;
      x_ups = false;
    } else {
      x = o;
      xst = bst[bi].out;
      x_ups = false;
      if (is_up && idx != 0) {  // nearest x2 (generator.py:248-249): gathered by the next block's SPADE epilogue
        x_ups = true;
        lvl_h *= 2;
        lvl_w *= 2;
      }
    }
  }
  // -- final image: tanh(conv_img(lrelu(x)))  generator.py:228 --
  pb.conv_final("conv_img", x, ACT_TANH, EXT_OUT_IMG, &mask_in, 2 * c.img_nc);

  // -- mask network (generator.py:493-510) --
  // down_lbl / down_img (C-N-A, stride 2 from level 1 on): a level's conv stores its RAW output parity-planar (+ its
  // statistics) and the next level's conv applies the instance norm + LeakyReLU to its halo tiles in shared memory
  // (Xf), which removes a read + write of every full-resolution map.  Measured (profiles/r1h_xf_layers.txt): the
  // transform costs an MMA-bound layer 15-25 % (tcgen05 already uses the whole shared-memory bandwidth), so the last
  // level and the res_flow blocks keep the separate in_apply pass.
  // RIB_XF = 0: separate in_apply everywhere; 1 (default): transform in the stride-2 down convs; 2: also in the
  // res_flow convs (the statistics of both branches' last level then share one buffer in cat order).
  View cat = pb.alloc("mask.cat", H >> c.mask_down, W >> c.mask_down, 2 * mch);
  for (int br = 0; br < 2; ++br) {
    const std::string bn = br == 0 ? "down_lbl" : "down_img";
    View cur = br == 0 ? lab : mask_in;
    PlanBuilder::Xf xf_prev;
    for (int i = 0; i <= c.mask_down; ++i) {
      const int h = H >> i, w = W >> i, ch = mask_nfilt(c, i);
      const std::string nm = "mask." + bn + "." + std::to_string(i);
      const std::string pre = f + bn + "." + std::to_string(i) + ".layers.norm";
      const bool lastl = i == c.mask_down;
      const bool raw_in_cat = lastl && xf_mode >= 2;
      const bool merged = merge_label && br == 0 && i == 0;   // already computed with down_first
      View raw = merged ? lbl0_raw : (raw_in_cat ? PlanBuilder::slice(cat, br * mch, mch) : pb.alloc(nm + ".raw", h, w, ch));
      if (!merged && !lastl && xf_mode >= 1) raw.parity = true;  // consumed only by the next (stride-2) conv
      double* st = raw_in_cat ? st_cat + (size_t)br * mch * 2 : mst[bn + std::to_string(i)];
      if (!merged)
        pb.conv_store(nm, cur, nullptr, i == 0 ? 1 : 2, raw, st, ACT_NONE, nullptr, nullptr,
                      (i == 0 || xf_mode < 1) ? nullptr : &xf_prev, raw_in_cat ? 2 * mch : 0);
      xf_prev.stats = st;
      xf_prev.w = G->in_affine.at(pre + ".weight");
      xf_prev.b = G->in_affine.at(pre + ".bias");
      xf_prev.act = 1;
      cur = raw;
      if (xf_mode < 1 || (lastl && xf_mode < 2)) {
        View o = lastl ? PlanBuilder::slice(cat, br * mch, mch) : pb.alloc(nm, h, w, ch);
        if (!lastl) o.parity = true;
        pb.in_apply(raw, st, pre, nullptr, nullptr, "", o, 1, false);
        cur = o;
      }
    }
  }
  PlanBuilder::Xf xf_cat;
  xf_cat.stats = st_cat;
  xf_cat.w = G->cat_w;
  xf_cat.b = G->cat_b;
  xf_cat.act = 1;
  const PlanBuilder::Xf* xfc = xf_mode >= 2 ? &xf_cat : nullptr;
  View r = cat;
  const int rh = H >> c.mask_down, rw = W >> c.mask_down;
  for (int i = 0; i < c.mask_res; ++i) {
    const std::string rn = "mask.res." + std::to_string(i), rp = f + "res_flow." + std::to_string(i);
    double* st0 = mst["res" + std::to_string(i) + "c0"];
    View raw0 = pb.alloc(rn + ".raw0", rh, rw, mch);
    pb.conv_store(rn + ".conv0", r, nullptr, 1, raw0, st0, ACT_NONE, nullptr, nullptr, i == 0 ? xfc : nullptr);
    View raw1 = pb.alloc(rn + ".raw1", rh, rw, mch);
    if (xf_mode >= 2) {
      PlanBuilder::Xf xf0;
      xf0.stats = st0;
      xf0.w = G->in_affine.at(rp + ".conv_block_0.layers.norm.weight");
      xf0.b = G->in_affine.at(rp + ".conv_block_0.layers.norm.bias");
      xf0.act = 1;
      pb.conv_store(rn + ".conv1", raw0, nullptr, 1, raw1, mst["res" + std::to_string(i) + "c1"], ACT_NONE, nullptr, nullptr, &xf0);
    } else {
      View t0 = pb.alloc(rn + ".t0", rh, rw, mch);
      pb.in_apply(raw0, st0, rp + ".conv_block_0.layers.norm", nullptr, nullptr, "", t0, 1, false);
      pb.conv_store(rn + ".conv1", t0, nullptr, 1, raw1, mst["res" + std::to_string(i) + "c1"], ACT_NONE, nullptr);
    }
    View o = pb.alloc(rn, rh, rw, mch);   // (the last block's x2 up-sampling is absorbed by the sub-pixel conv)
    if (i == 0) {
      View raws = pb.alloc(rn + ".raws", rh, rw, mch);
      pb.conv_store(rn + ".convs", r, nullptr, 1, raws, mst["res0cs"], ACT_NONE, nullptr, nullptr, xfc);
      pb.in_apply(raw1, mst["res0c1"], rp + ".conv_block_1.layers.norm", &raws, mst["res0cs"],
                  rp + ".conv_block_s.layers.norm", o, 0, false);
    } else {
      pb.in_apply(raw1, mst["res" + std::to_string(i) + "c1"], rp + ".conv_block_1.layers.norm", &r, nullptr, "", o, 0, false);
    }
    r = o;
  }
  for (int k = 0; k < c.mask_down; ++k) {
    const int i = c.mask_down - 1 - k;
    const int h = H >> i, w = W >> i, ch = mask_nfilt(c, i);
    const std::string nm = "mask.up." + std::to_string(k);
    View raw = pb.alloc(nm + ".raw", h, w, ch);
    raw.parity = true;
    pb.conv_subpix(nm, r, raw, mst["up" + std::to_string(k)]);
    View o = pb.alloc(nm, h, w, ch);
    pb.in_apply(raw, mst["up" + std::to_string(k)], f + "up_flow." + std::to_string(2 * k + 1) + ".layers.norm", nullptr, nullptr,
                "", o, 1, false);
    r = o;
  }
  pb.conv_final("mask.conv_mask", r, ACT_SIGMOID, EXT_OUT_MASK, nullptr, 0);
  pb.final_scratch = static_cast<float*>(pb.ws.take((size_t)B * 4 * H * W * sizeof(float)));

  if (pb.rc) return pb.rc;
  *bytes_out = align_up(pb.ws.off, 1024);
  if (dry_ops != nullptr && wsbase == nullptr) {
    if (autotune_enabled()) apply_tune_table(pb);
    dry_ops->swap(pb.ops);
  }
  if (wsbase) {
    if (!g_debug_simt && autotune_enabled()) autotune(pb, stream);
    G->ops.swap(pb.ops);
    G->zero_once.clear();
    for (const View* v : {&emb_in, &mask_in})
      G->zero_once.push_back({v->p + (size_t)H * W * 8, (size_t)v->bstride() * sizeof(act_t), (size_t)H * W * 8 * sizeof(act_t), B});
    G->zero_pending = true;
    G->debug_views.swap(pb.views);
    G->pB = B;
    G->pH = H;
    G->pW = W;
    G->pws = wsbase;
    G->ws_bytes_needed = *bytes_out;
    G->plan_simt = g_debug_simt != 0;
  }
  return 0;
}

long long generator_workspace_bytes(Generator* G, int B, int H, int W) {
  size_t bytes = 0;
  int rc = build_plan(G, B, H, W, nullptr, &bytes, nullptr);
  if (rc) return rc;
  return (long long)bytes;
}

static std::atomic<long long> g_misc_launches{0};
long long misc_launch_count() { return g_misc_launches.load(); }
void count_misc_launch(int n) { g_misc_launches.fetch_add(n); }

// Builds (or re-uses) the launch plan for (B, H, W) on this workspace.
static int ensure_plan(Generator* G, int B, int H, int W, void* ws, long long ws_bytes, cudaStream_t stream,
                       bool* rebuilt = nullptr) {
  RIB_REQUIRE(G && ws, "forward: null argument");
  RIB_REQUIRE(((uintptr_t)ws & 1023) == 0, "forward: workspace must be 1024-byte aligned");
  if (rebuilt) *rebuilt = false;
  if (G->pB != B || G->pH != H || G->pW != W || G->pws != ws || G->plan_simt != (g_debug_simt != 0)) {
    if (rebuilt) *rebuilt = true;
    size_t need = 0;
    int rc = build_plan(G, B, H, W, nullptr, &need, nullptr);
    if (rc) return rc;
    RIB_REQUIRE((long long)need <= ws_bytes, "forward: workspace too small");
    rc = build_plan(G, B, H, W, ws, &need, stream);
    if (rc) return rc;
  }
  RIB_REQUIRE((long long)G->ws_bytes_needed <= ws_bytes, "forward: workspace too small");
  return 0;
}

int generator_bind(Generator* G, int B, int H, int W, void* ws, long long ws_bytes, void** label_planar,
                   cudaStream_t stream) {
  // (the auto-tuner's timing launches write into the workspace: they run on the caller's stream)
  int rc = ensure_plan(G, B, H, W, ws, ws_bytes, stream);
  if (rc) return rc;
  if (label_planar) *label_planar = G->debug_views.at("label").p;
  return 0;
}

int generator_forward(Generator* G, int B, int H, int W, const float* label, const float* img_fake,
                      const float* img_prev, float* out_img, float* out_mask, void* ws, long long ws_bytes,
                      cudaStream_t stream) {
  RIB_REQUIRE(G && img_fake && img_prev && out_img && out_mask && ws, "forward: null argument");
  bool rebuilt = false;
  int prc = ensure_plan(G, B, H, W, ws, ws_bytes, stream, &rebuilt);
  if (prc) return prc;
  // label == NULL means "the caller rasterised into the buffer rib_generator_bind returned": that buffer only exists
  // if the plan for this exact (B, H, W, workspace) was already in place
  RIB_REQUIRE(label != nullptr || !rebuilt,
              "forward: label is NULL but no plan was bound to this (B, H, W, workspace); call rib_generator_bind first");
  if (G->zero_pending) {
    for (const Generator::ZeroPlane& z : G->zero_once)
      RIB_CHECK_CUDA(cudaMemset2DAsync(z.ptr, z.pitch, 0, z.width, (size_t)z.rows, stream));
    G->zero_pending = false;
  }
  for (const Op& op : G->ops) {
    int rc = 0;
    switch (op.kind) {
      case OP_PACK_IMAGES:
        rc = launch_pack_images(img_fake, img_prev, op.pi_emb, op.pi_emb_bs, op.pi_mask, op.pi_mask_bs, B, H, W, stream);
        count_misc_launch(1);
        break;
      case OP_MEMSET:
        RIB_CHECK_CUDA(cudaMemsetAsync(op.ms_ptr, 0, op.ms_bytes, stream));
        break;
      case OP_PACK: {
        if (op.pk_ext[0] == EXT_LABEL && label == nullptr) break;  // the caller filled the planar label buffer
        PackSrc srcs[3];
        for (int i = 0; i < op.pk_nsrc; ++i) {
          srcs[i].p = op.pk_ext[i] == EXT_LABEL ? label : (op.pk_ext[i] == EXT_FAKE ? img_fake : img_prev);
          srcs[i].C = op.pk_C[i];
        }
        rc = launch_pack_nchw(srcs, op.pk_nsrc, op.pk_dst, op.pk_bs, 0, 0, op.pk_nplanes, B, H, W, stream);
        count_misc_launch(1);
        break;
      }
      case OP_GEMM: {
        if (op.mode == EPI_FINAL) {
          ConvGemmParams p = op.g;
          p.out_f32 = op.ext == EXT_OUT_IMG ? out_img : out_mask;
          rc = launch_conv_gemm(p, op.mode, stream);
        } else {
          rc = launch_conv_gemm(op.g, op.mode, stream);
        }
        break;
      }
      case OP_IN_APPLY:
        rc = launch_in_apply(op.ia, stream);
        count_misc_launch(1);
        break;
      case OP_POOL:
        rc = launch_avgpool3s2(op.pl_src, op.pl_sbs, op.pl_dst, op.pl_dbs, op.pl_stats, B, op.pl_H, op.pl_W, op.pl_C,
                               stream);
        count_misc_launch(1);
        break;
    }
    if (rc) return rc;
  }
  return 0;
}

// One text line per planned launch: kind, layer, tiling and the algorithmic work, in launch order.
static int plan_text_of(const std::vector<Op>& ops, char* buf, long long cap);

int generator_plan_text(Generator* G, char* buf, long long cap) { return plan_text_of(G->ops, buf, cap); }

// Launch plan for (B, H, W) without a GPU: layer table, tiling (default heuristics + imported tuning table), shared-memory
// and workspace sizes.  Validates every layer's configuration for the shape; nothing is launched or allocated.
int generator_plan_dry_run(const rib_gen_config* cfg, int B, int H, int W, long long* ws_bytes, char* buf, long long cap) {
  RIB_REQUIRE(cfg != nullptr, "plan_dry_run: null config");
  const rib_gen_config& c = *cfg;
  RIB_REQUIRE(c.nf % 16 == 0 && c.emb_nf % 16 == 0 && c.mask_nf % 16 == 0, "channel counts must be multiples of 16");
  RIB_REQUIRE(c.label_nc <= 32 && c.img_nc == 3, "unsupported input channel counts");
  Generator G;
  G.cfg = c;
  std::vector<PackJob> jobs;
  define_layers(&G, jobs);
  {  // the instance-norm affine tables of the mask network: names only
    const std::string f = "flow_network_temp.";
    auto reg = [&](const std::string& prefix) {
      G.in_affine[prefix + ".weight"] = nullptr;
      G.in_affine[prefix + ".bias"] = nullptr;
    };
    for (int br = 0; br < 2; ++br)
      for (int i = 0; i <= c.mask_down; ++i)
        reg(f + (br == 0 ? "down_lbl." : "down_img.") + std::to_string(i) + ".layers.norm");
    for (int i = 0; i < c.mask_res; ++i) {
      reg(f + "res_flow." + std::to_string(i) + ".conv_block_0.layers.norm");
      reg(f + "res_flow." + std::to_string(i) + ".conv_block_1.layers.norm");
      if (i == 0) reg(f + "res_flow.0.conv_block_s.layers.norm");
    }
    for (int k = 0; k < c.mask_down; ++k) reg(f + "up_flow." + std::to_string(2 * k + 1) + ".layers.norm");
  }
  size_t bytes = 0;
  std::vector<Op> ops;
  int rc = build_plan(&G, B, H, W, nullptr, &bytes, nullptr, &ops);
  if (rc) return rc;
  if (ws_bytes) *ws_bytes = (long long)bytes;
  if (buf == nullptr) return 0;
  return plan_text_of(ops, buf, cap);
}

static int plan_text_of(const std::vector<Op>& ops, char* buf, long long cap) {
  std::string out;
  char line[512];
  for (const Op& op : ops) {
    switch (op.kind) {
      case OP_MEMSET:
        snprintf(line, sizeof(line), "memset bytes=%zu\n", op.ms_bytes);
        break;
      case OP_PACK:
        snprintf(line, sizeof(line), "pack planes=%d\n", op.pk_nplanes);
        break;
      case OP_PACK_IMAGES:
        snprintf(line, sizeof(line), "pack_images\n");
        break;
      case OP_GEMM: {
        const ConvGemmParams& p = op.g;
        const long long K = (long long)p.stages0 * p.BKc * p.ntaps + (long long)p.stages1 * p.BKc;
        const long long N = (long long)p.n_tiles * p.BN;
        const long long M = (long long)p.B * p.H * p.W;
        snprintf(line, sizeof(line),
                 "gemm %s mode=%d B=%d H=%d W=%d N=%lld nvalid=%d K=%lld BN=%d BKc=%d MT=%d taps=%d stride=%d "
                 "cin0=%d cin1=%d bres=%d aring=%d bring=%d smem=%zu flops=%.6g pair=%d ppc=%d\n",
                 op.name.c_str(), op.mode, p.B, p.H, p.W, N, p.n_valid, K, p.BN, p.BKc, p.MT, p.ntaps, p.stride,
                 p.stages0 * p.BKc, p.stages1 * p.BKc, p.b_resident, p.a_ring, p.b_ring, conv_gemm_smem_bytes(p),
                 2.0 * (double)M * (double)N * (double)K, p.pair, p.subpix ? p.ppc : 1);
        break;
      }
      case OP_IN_APPLY:
        snprintf(line, sizeof(line), "in_apply B=%d H=%d W=%d C=%d ups=%d two=%d\n", op.ia.B, op.ia.H, op.ia.W, op.ia.C,
                 op.ia.ups, op.ia.b != nullptr);
        break;
      case OP_POOL:
        snprintf(line, sizeof(line), "pool H=%d W=%d C=%d\n", op.pl_H, op.pl_W, op.pl_C);
        break;
    }
    out += line;
  }
  if ((long long)out.size() + 1 > cap) {
    set_error("plan text does not fit the buffer");
    return -1;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int generator_debug_tensor(Generator* G, const char* name, const void** ptr, int* B, int* H, int* W, int* C,
                           int* ld) {
  auto it = G->debug_views.find(name);
  if (it == G->debug_views.end()) {
    set_error(std::string("no such plan tensor: ") + name);
    return -1;
  }
  *ptr = it->second.p;
  *B = it->second.B;
  *H = it->second.H;
  *W = it->second.W;
  *C = it->second.C;
  *ld = it->second.Ctot;
  return 0;
}

// ---- stand-alone conv for unit tests ----------------------------------------------------------
long long conv_test_scratch_bytes(int Cin, int Cout, int k) {
  // (sized for the sub-pixel form as well: 4 parities x 4 taps)
  const size_t kk = (size_t)std::max(k * k, 16);
  return (long long)(align_up((size_t)Cout * Cin * kk * sizeof(act_t), 256) + align_up((size_t)Cout * 4 * 4, 256) + 1024);
}

int conv_test(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
              int Cin, int Cout, int k, int stride, int act, void* scratch, cudaStream_t stream) {
  return conv_test_ex(x, w, bias, out, stats, B, Hin, Win, Cin, Cout, k, stride, act, 0, nullptr, nullptr, nullptr, 0,
                      scratch, stream);
}

int conv_test_ex(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                 int Cin, int Cout, int k, int stride, int act, int subpix, const double* xf_stats, const float* xf_w,
                 const float* xf_b, int xf_act, void* scratch, cudaStream_t stream) {
  RIB_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "conv_test: channels must be multiples of 16");
  RIB_REQUIRE((k == 1 || k == 3) && (stride == 1 || stride == 2), "conv_test: unsupported kernel/stride");
  RIB_REQUIRE(Cin <= 64 || Cin % 64 == 0, "conv_test: Cin must be 16, 32, 64 or a multiple of 64");
  RIB_REQUIRE(!subpix || (k == 3 && stride == 1 && xf_stats == nullptr && act == 0), "conv_test: bad sub-pixel conv");
  Generator G;  // a throw-away holder for one layer
  GemmLayer L;
  L.name = "test";
  L.n_valid = Cout;
  L.n_pad = subpix ? 4 * Cout : Cout;
  L.cin0 = Cin;
  L.taps = subpix ? 4 : k * k;
  L.cin1 = 0;
  L.ktotal = Cin * L.taps;
  L.BN = std::min(Cout, 128);
  if (subpix && Cout < 128 && 128 % Cout == 0 && getenv("RIB_TEST_PPC") != nullptr && atoi(getenv("RIB_TEST_PPC")) > 1) {
    L.ppc = std::min(std::min(4, atoi(getenv("RIB_TEST_PPC"))), 128 / Cout);   // several output parities per CTA (kernel tests)
    L.BN = Cout * L.ppc;
  }
  L.stride = stride;
  L.bkc = choose_bkc(Cin, 0, L.taps, L.BN, stride);
  uint8_t* sp = static_cast<uint8_t*>(scratch);
  sp = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)sp, 256));
  L.w = reinterpret_cast<act_t*>(sp);
  L.bias = reinterpret_cast<float*>(sp + align_up((size_t)L.n_pad * L.ktotal * sizeof(act_t), 256));
  G.layers["test"] = L;
  RIB_CHECK_CUDA(cudaMemsetAsync(L.bias, 0, (size_t)L.n_pad * 4, stream));
  for (int par = 0; par < (subpix ? 4 : 1); ++par) {
    PackWeightParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.w = w;
    pp.bias = bias;
    pp.Cout = Cout;
    pp.Cin = Cin;
    pp.taps = k * k;
    pp.dst = L.w;
    pp.bias_dst = L.bias;
    pp.ktotal = L.ktotal;
    pp.bkc = L.bkc;
    pp.row_off = par * Cout;
    pp.subpix = subpix ? 1 : 0;
    pp.subpix_parity = par;
    int rc = launch_pack_weight(pp, stream);
    if (rc) return rc;
  }
  PlanBuilder pb(&G, scratch, B);  // non-null base => builds real tensor maps; no allocation is made
  View in;
  in.p = static_cast<act_t*>(const_cast<void*>(x));
  in.B = B;
  in.H = Hin;
  in.W = Win;
  in.C = in.Ctot = Cin;
  // (RIB_TEST_S2_NORMAL=1, kernel tests only: the stride-2 input is given in the normal layout, as emb_1 reads cond_0)
  in.parity = stride == 2 && !(getenv("RIB_TEST_S2_NORMAL") != nullptr && atoi(getenv("RIB_TEST_S2_NORMAL")) == 1);
  View o;
  o.p = static_cast<act_t*>(out);
  o.B = B;
  o.H = subpix ? 2 * Hin : Hin / stride;
  o.W = subpix ? 2 * Win : Win / stride;
  o.C = o.Ctot = Cout;
  if (subpix) {
    o.parity = true;
    pb.conv_subpix("test", in, o, stats);
  } else {
    PlanBuilder::Xf xf;
    xf.stats = xf_stats;
    xf.w = xf_w;
    xf.b = xf_b;
    xf.act = xf_act;
    pb.conv_store("test", in, nullptr, stride, o, stats, act, nullptr, nullptr, xf_stats ? &xf : nullptr);
  }
  if (pb.rc) return pb.rc;
  return launch_conv_gemm(pb.ops[0].g, EPI_STORE, stream);
}

}  // namespace rib
