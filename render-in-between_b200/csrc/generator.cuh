// Generator plan / forward (declarations).  See generator.cu.
#pragma once
#include "common.cuh"
#include "../../include/rib_b200.h"

namespace rib {

struct Generator;
extern int g_debug_simt;

int generator_create(const rib_gen_config* cfg, const rib_tensor* tensors, int n, cudaStream_t stream, Generator** out);
void generator_destroy(Generator* g);
long long generator_workspace_bytes(Generator* g, int B, int H, int W);
int generator_forward(Generator* g, int B, int H, int W, const float* label, const float* img_fake,
                      const float* img_prev, float* out_img, float* out_mask, void* ws, long long ws_bytes,
                      cudaStream_t stream);
int generator_bind(Generator* g, int B, int H, int W, void* ws, long long ws_bytes, void** label_planar,
                   cudaStream_t stream);
int generator_debug_tensor(Generator* g, const char* name, const void** ptr, int* B, int* H, int* W, int* C, int* ld);
int generator_plan_text(Generator* g, char* buf, long long cap);
int generator_plan_dry_run(const rib_gen_config* cfg, int B, int H, int W, long long* ws_bytes, char* buf, long long cap);
int generator_tune_log(char* buf, long long cap);
int generator_tune_export(char* buf, long long cap);
int generator_tune_import(const char* text);
long long conv_test_scratch_bytes(int Cin, int Cout, int k);
int conv_test(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
              int Cin, int Cout, int k, int stride, int act, void* scratch, cudaStream_t stream);
int conv_test_ex(const void* x, const float* w, const float* bias, void* out, double* stats, int B, int Hin, int Win,
                 int Cin, int Cout, int k, int stride, int act, int subpix, const double* xf_stats, const float* xf_w,
                 const float* xf_b, int xf_act, void* scratch, cudaStream_t stream);
long long misc_launch_count();
void count_misc_launch(int n);

}  // namespace rib
