// Pose rasterisation kernels (declarations).  See raster.cu.
#pragma once
#include "common.cuh"

namespace rib {

// joints_dev: device [B][19][3] float64 (x, y, confidence) in model-pixel coordinates.
// wtab41_host: host [41] float64 normalised Gaussian taps (scipy _gaussian_kernel1d(sigma=5, radius=20)).
// label: device [B][22][H][W] float32, fully overwritten.
int launch_rasterize(const double* joints_dev, int B, int H, int W, const double* wtab41_host, double skeleton_thres,
                     double foot_thres, float* label, cudaStream_t stream);

}  // namespace rib
