// TMA halo-box throughput probe (sm_100a): how fast can persistent CTAs stream a chunk-planar 16-bit activation map
// [B][C/8][H][W][8] into shared memory with the conv kernel's 4-D boxes, as a function of the box shape?
// No MMA, no epilogue: one elected lane issues the loads into a ring, another waits for each slot and frees it.
// Output: GB/s of useful (non-halo) input bytes and of bytes actually moved, per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I render-in-between_b200/csrc -o tools/probe/tma_probe tools/probe/tma_probe.cu -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace rib;

struct Variant {
  int tile_w, tile_h;  // output pixels covered by one box (the box adds a 1-pixel halo on every side)
  int planes;          // 8-channel planes per box
  int ring;            // slots per CTA
  int ctas_per_sm;
};

struct Params {
  CUtensorMap map;
  int B, H, W, C8;
  int tile_w, tile_h, planes, ring;
  int tiles_x, tiles_y;
  uint32_t slot_bytes, tx_bytes;
};

__global__ void __launch_bounds__(64) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int groups = p.C8 / p.planes;
  const long long per_img = (long long)p.tiles_x * p.tiles_y;
  const long long total = per_img * p.B;
  const long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;
  if (warp == 0) {
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / per_img);
        const int r = (int)(t - n * per_img);
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(smem_u32(&empty[slot]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full[slot]);
          mbar_arrive_expect_tx(fb, p.tx_bytes);
          tma_load_4d(smem_u32(smem) + slot * p.slot_bytes, &p.map, fb, (tx * p.tile_w - 1) * 8, ty * p.tile_h - 1,
                      g * p.planes, n);
          if (++slot == p.ring) slot = 0, phase ^= 1u;
        }
      }
    }
  } else {
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      for (long long t = t0; t < t1; ++t)
        for (int g = 0; g < groups; ++g) {
          mbar_wait(smem_u32(&full[slot]), phase);
          mbar_arrive(smem_u32(&empty[slot]));
          if (++slot == p.ring) slot = 0, phase ^= 1u;
        }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 32, H = 512, W = 512;
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int Cmax = 64;
  const size_t bytes = (size_t)B * Cmax * H * W * 2;
  void* buf = nullptr;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  printf("attr: %s\n", cudaGetErrorString(cudaGetLastError()));
  printf("%-34s %10s %12s %12s\n", "variant", "us", "useful GB/s", "moved GB/s");
  const int chans[] = {16, 32, 64};
  const Variant vs[] = {
      {8, 16, 2, 4, 3},  {8, 32, 2, 4, 3},  {16, 16, 2, 4, 3}, {16, 32, 2, 4, 3}, {32, 16, 2, 4, 3}, {32, 32, 2, 2, 3},
      {8, 16, 4, 4, 3},  {8, 32, 4, 4, 3},  {16, 16, 4, 4, 3}, {16, 32, 4, 3, 3}, {32, 16, 4, 3, 3},
      {8, 16, 2, 8, 3},  {8, 16, 2, 4, 4},  {8, 16, 2, 4, 2},  {8, 16, 2, 4, 1},  {8, 32, 2, 8, 2},
  };
  for (int C : chans)
    for (const Variant& v : vs) {
      if (v.planes * 8 > C) continue;
      Params p;
      p.B = B, p.H = H, p.W = W, p.C8 = C / 8;
      p.tile_w = v.tile_w, p.tile_h = v.tile_h, p.planes = v.planes, p.ring = v.ring;
      p.tiles_x = W / v.tile_w, p.tiles_y = H / v.tile_h;
      const int bw = v.tile_w + 2, bh = v.tile_h + 2;
      p.tx_bytes = (uint32_t)(bw * bh * 16 * v.planes);
      p.slot_bytes = (p.tx_bytes + 1023u) & ~1023u;
      const size_t smem = (size_t)p.slot_bytes * v.ring + 2048;
      if (smem * v.ctas_per_sm > 220 * 1024) continue;
      cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C * H * W * 2};
      cuuint32_t box[4] = {(cuuint32_t)(bw * 8), (cuuint32_t)bh, (cuuint32_t)v.planes, 1u};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        continue;
      }
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      const int grid = 148 * v.ctas_per_sm;
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<<<grid, 64, smem>>>(p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      cudaError_t ce = cudaGetLastError();
      if (ce != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(ce));
        return 1;
      }
      const double useful = (double)B * C * H * W * 2;
      const double moved = (double)p.tiles_x * p.tiles_y * B * (C / 8 / v.planes) * p.tx_bytes;
      char name[96];
      snprintf(name, sizeof(name), "C%d tile %2dx%2d planes %d ring %d cta/sm %d", C, v.tile_w, v.tile_h, v.planes, v.ring,
               v.ctas_per_sm);
      printf("%-34s %10.1f %12.1f %12.1f\n", name, best * 1e3, useful / best / 1e6, moved / best / 1e6);
    }
  return 0;
}
