// Pose rasterisation kernels (declarations).  See raster.cu.
#pragma once
#include "common.cuh"

namespace rib {

// Bytes of device scratch launch_rasterize needs for B frames (limb tables, stamp flags, heat-map windows).
long long raster_workspace_bytes(int B, int H, int W);

// joints_dev: device [B][19][3] float64 (x, y, confidence) in model-pixel coordinates.
// wtab41_host: host [41] float64 normalised Gaussian taps (scipy _gaussian_kernel1d(sigma=5, radius=20)).
// label: device [B][22][H][W] float32, fully overwritten (may be null).
// label_planar: device 16-bit chunk-planar [B][4][H][W][8] (the generator's padded 32-channel label input,
//               channels 22..31 zero), fully overwritten (may be null).  At least one output is required.
int launch_rasterize(const double* joints_dev, int B, int H, int W, const double* wtab41_host, double skeleton_thres,
                     double foot_thres, float* label, act_t* label_planar, void* workspace, long long workspace_bytes,
                     cudaStream_t stream);

}  // namespace rib
