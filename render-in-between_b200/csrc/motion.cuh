// Upstream motion model: the encoder / decoder Transformer of Human_Motion_Modelling that
// interpolates 2-D joint sequences, on the GPU in the renderer's process (no JSON round trip between the two stages).
#pragma once
#include "common.cuh"
#include "../../include/rib_b200.h"

namespace rib {

struct MotionModel;

int motion_create(const rib_motion_config* cfg, const rib_tensor* tensors, int n_tensors, cudaStream_t stream,
                  MotionModel** out);
void motion_destroy(MotionModel* m);
long long motion_workspace_bytes(const MotionModel* m, int B, int L);
int motion_forward(MotionModel* m, int B, int L, const float* src, const uint8_t* src_mask, const float* src_pos,
                   const uint8_t* tgt_mask, const float* tgt_pos, int rate, float* joints, float* reco, void* workspace,
                   long long workspace_bytes, cudaStream_t stream);

}  // namespace rib
