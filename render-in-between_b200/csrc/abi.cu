// extern "C" surface of librib_b200.so (see include/rib_b200.h).  Nothing throws across this boundary.
#include <string>

#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "generator.cuh"
#include "motion.cuh"
#include "raster.cuh"

namespace rib {
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
const char* last_error() { return t_error.c_str(); }
}  // namespace rib

using namespace rib;

#define RIB_GUARD_BEGIN try {
#define RIB_GUARD_END                                      \
  }                                                        \
  catch (const std::exception& e) {                        \
    rib::set_error(std::string("exception: ") + e.what()); \
    return -9;                                             \
  }                                                        \
  catch (...) {                                            \
    rib::set_error("unknown exception");                   \
    return -9;                                             \
  }

extern "C" {

const char* rib_last_error(void) { return rib::last_error(); }
int rib_abi_version(void) { return RIB_ABI_VERSION; }
long long rib_kernel_launch_count(void) { return conv_gemm_launch_count() + misc_launch_count(); }

long long rib_rasterize_workspace_bytes(int B, int H, int W) {
  return B > 0 && H > 0 && W > 0 ? raster_workspace_bytes(B, H, W) : -1;
}

int rib_rasterize(const double* joints, int B, int H, int W, const double* gauss_taps, double skeleton_thres,
                  double foot_thres, float* label, void* label_planar, void* workspace, long long workspace_bytes,
                  void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(joints && gauss_taps && (label || label_planar) && workspace, "rib_rasterize: null argument");
  int rc = launch_rasterize(joints, B, H, W, gauss_taps, skeleton_thres, foot_thres, label,
                            static_cast<act_t*>(label_planar), workspace, workspace_bytes, (cudaStream_t)stream);
  if (!rc) count_misc_launch(3);
  return rc;
  RIB_GUARD_END
}

int rib_warp(const float* src, const void* flow, int flow_fp16, float* out, int B, int C, int H, int W,
             long long src_bstride, long long flow_bstride, long long out_bstride, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(src && flow && out && B > 0 && C > 0 && H > 0 && W > 0, "rib_warp: bad argument");
  int rc = launch_warp(src, flow, flow_fp16, out, B, C, H, W, src_bstride, flow_bstride, out_bstride, (cudaStream_t)stream);
  if (!rc) count_misc_launch(1);
  return rc;
  RIB_GUARD_END
}

int rib_composite(const float* img, const float* mask, const float* dain, float* out_f32, uint8_t* out_u8, int B,
                  int H, int W, long long img_bstride, long long out_f32_bstride, long long out_u8_bstride,
                  void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(img && (mask == nullptr || dain != nullptr) && (out_f32 || out_u8) && B > 0 && H > 0 && W > 0,
              "rib_composite: bad argument");
  int rc = launch_composite(img, mask, dain, out_f32, out_u8, B, H, W, img_bstride, out_f32_bstride, out_u8_bstride,
                            (cudaStream_t)stream);
  if (!rc) count_misc_launch(1);
  return rc;
  RIB_GUARD_END
}

int rib_resize_cubic_u8(const uint8_t* frames, uint8_t* out, int B, int h, int w, int H, int W, long long in_bstride,
                        long long out_bstride, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(frames && out, "rib_resize_cubic_u8: null argument");
  int rc = launch_resize_cubic_u8(frames, out, B, h, w, H, W, in_bstride, out_bstride, (cudaStream_t)stream);
  if (!rc) count_misc_launch(1);
  return rc;
  RIB_GUARD_END
}

int rib_frames_from_u8(const uint8_t* frames, float* out, uint8_t* out_u8, int B, int H, int W, long long in_bstride,
                       long long out_bstride, long long out_u8_bstride, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(frames && out && B > 0 && H > 0 && W > 0, "rib_frames_from_u8: bad argument");
  int rc = launch_frames_from_u8(frames, out, out_u8, B, H, W, in_bstride, out_bstride, out_u8_bstride, (cudaStream_t)stream);
  if (!rc) count_misc_launch(1);
  return rc;
  RIB_GUARD_END
}

int rib_generator_create(const rib_gen_config* cfg, const rib_tensor* tensors, int n_tensors, void* stream,
                         rib_generator** out) {
  RIB_GUARD_BEGIN
  return generator_create(cfg, tensors, n_tensors, (cudaStream_t)stream, reinterpret_cast<Generator**>(out));
  RIB_GUARD_END
}

void rib_generator_destroy(rib_generator* g) { generator_destroy(reinterpret_cast<Generator*>(g)); }

long long rib_generator_workspace_bytes(rib_generator* g, int B, int H, int W) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(g, "rib_generator_workspace_bytes: null generator");
  return generator_workspace_bytes(reinterpret_cast<Generator*>(g), B, H, W);
  RIB_GUARD_END
}

int rib_generator_bind(rib_generator* g, int B, int H, int W, void* workspace, long long workspace_bytes,
                       void** label_planar, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(g && workspace, "rib_generator_bind: null argument");
  return generator_bind(reinterpret_cast<Generator*>(g), B, H, W, workspace, workspace_bytes, label_planar,
                        (cudaStream_t)stream);
  RIB_GUARD_END
}

int rib_generator_forward(rib_generator* g, int B, int H, int W, const float* label, const float* img_fake,
                          const float* img_prev, float* out_img, float* out_mask, void* workspace,
                          long long workspace_bytes, void* stream) {
  RIB_GUARD_BEGIN
  return generator_forward(reinterpret_cast<Generator*>(g), B, H, W, label, img_fake, img_prev, out_img, out_mask,
                           workspace, workspace_bytes, (cudaStream_t)stream);
  RIB_GUARD_END
}

int rib_motion_create(const rib_motion_config* cfg, const rib_tensor* tensors, int n_tensors, void* stream, rib_motion** out) {
  RIB_GUARD_BEGIN
  return motion_create(cfg, tensors, n_tensors, (cudaStream_t)stream, reinterpret_cast<MotionModel**>(out));
  RIB_GUARD_END
}

void rib_motion_destroy(rib_motion* m) { motion_destroy(reinterpret_cast<MotionModel*>(m)); }

long long rib_motion_workspace_bytes(rib_motion* m, int B, int L) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(m && B > 0 && L > 0, "rib_motion_workspace_bytes: bad argument");
  return motion_workspace_bytes(reinterpret_cast<MotionModel*>(m), B, L);
  RIB_GUARD_END
}

int rib_motion_forward(rib_motion* m, int B, int L, const float* src, const uint8_t* src_mask, const float* src_pos,
                       const uint8_t* tgt_mask, const float* tgt_pos, int rate, float* joints, float* reco,
                       void* workspace, long long workspace_bytes, void* stream) {
  RIB_GUARD_BEGIN
  return motion_forward(reinterpret_cast<MotionModel*>(m), B, L, src, src_mask, src_pos, tgt_mask, tgt_pos, rate, joints, reco,
                        workspace, workspace_bytes, (cudaStream_t)stream);
  RIB_GUARD_END
}

void rib_profile_enable(int enable) { conv_gemm_profile_enable(enable); }
int rib_profile_collect(double* conv_ms, long long* conv_launches) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(conv_ms && conv_launches, "rib_profile_collect: null argument");
  return conv_gemm_profile_collect(conv_ms, conv_launches, nullptr, 0);
  RIB_GUARD_END
}
int rib_profile_collect_launches(double* conv_ms, long long* conv_launches, float* per_launch_ms, long long cap) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(conv_ms && conv_launches && per_launch_ms && cap > 0, "rib_profile_collect_launches: bad argument");
  return conv_gemm_profile_collect(conv_ms, conv_launches, per_launch_ms, cap);
  RIB_GUARD_END
}

void rib_debug_set_simt(int enable) { rib::g_debug_simt = enable ? 1 : 0; }
int rib_debug_get_simt(void) { return rib::g_debug_simt; }

int rib_generator_debug_tensor(rib_generator* g, const char* name, const void** ptr, int* B, int* H, int* W, int* C,
                               int* ld) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(g && name && ptr && B && H && W && C && ld, "rib_generator_debug_tensor: null argument");
  return generator_debug_tensor(reinterpret_cast<Generator*>(g), name, ptr, B, H, W, C, ld);
  RIB_GUARD_END
}

int rib_generator_plan_text(rib_generator* g, char* buf, long long cap) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(g && buf && cap > 0, "rib_generator_plan_text: bad argument");
  return generator_plan_text(reinterpret_cast<Generator*>(g), buf, cap);
  RIB_GUARD_END
}

int rib_plan_dry_run(const rib_gen_config* cfg, int B, int H, int W, long long* ws_bytes, char* buf, long long cap) {
  RIB_GUARD_BEGIN
  return generator_plan_dry_run(cfg, B, H, W, ws_bytes, buf, cap);
  RIB_GUARD_END
}

int rib_tune_log(char* buf, long long cap) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(buf && cap > 0, "rib_tune_log: bad argument");
  RIB_REQUIRE(generator_tune_log(buf, cap) == 0, "rib_tune_log: buffer too small");
  return 0;
  RIB_GUARD_END
}

int rib_tune_export(char* buf, long long cap) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(buf && cap > 0, "rib_tune_export: bad argument");
  RIB_REQUIRE(generator_tune_export(buf, cap) == 0, "rib_tune_export: buffer too small");
  return 0;
  RIB_GUARD_END
}

int rib_tune_import(const char* text) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(text != nullptr, "rib_tune_import: null argument");
  return generator_tune_import(text);
  RIB_GUARD_END
}

int rib_act_is_fp16(void) {
#ifdef RIB_ACT_FP16
  return 1;
#else
  return 0;
#endif
}

long long rib_conv_test_scratch_bytes(int Cin, int Cout, int k) { return conv_test_scratch_bytes(Cin, Cout, k); }

int rib_conv_test(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                  int Cin, int Cout, int k, int stride, int act, void* scratch, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(x && w && out && scratch, "rib_conv_test: null argument");
  return conv_test(x, w, bias, out, stats, B, Hin, Win, Cin, Cout, k, stride, act, scratch, (cudaStream_t)stream);
  RIB_GUARD_END
}

int rib_conv_test_ex(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                     int Cin, int Cout, int k, int stride, int act, int subpix, const double* xf_stats, const float* xf_w,
                     const float* xf_b, int xf_act, void* scratch, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(x && w && out && scratch, "rib_conv_test_ex: null argument");
  return conv_test_ex(x, w, bias, out, stats, B, Hin, Win, Cin, Cout, k, stride, act, subpix, xf_stats, xf_w, xf_b, xf_act,
                      scratch, (cudaStream_t)stream);
  RIB_GUARD_END
}

int rib_avgpool_test(const void* x, void* out, double* stats, int B, int H, int W, int C, void* stream) {
  RIB_GUARD_BEGIN
  RIB_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0, "rib_avgpool_test: bad argument");
  return launch_avgpool3s2(static_cast<const act_t*>(x), (long long)C * H * W, static_cast<act_t*>(out),
                           (long long)C * (H / 2) * (W / 2), stats, B, H, W, C, (cudaStream_t)stream);
  RIB_GUARD_END
}

}  // extern "C"
