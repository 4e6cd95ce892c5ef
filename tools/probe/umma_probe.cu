// UMMA shared-memory descriptor probe (sm_100a): for a set of A-operand layouts (swizzle mode, 8-row-group stride,
// start-row shift, base_offset policy) checks whether tcgen05.mma reads the rows we expect and measures
// cycles per MMA (M=128, K=16).  Data is written the way TMA writes a swizzled box: the 16-byte chunk index of
// every row is XORed with address bits [7:9] (128B), [7:8] (64B) or [7] (32B) of the ABSOLUTE shared address.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I render-in-between_b200/csrc -o tools/probe/umma_probe tools/probe/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace rib;

struct Variant {
  int swz;        // 0, 32, 64, 128 : swizzle bytes (= row bytes of the K-major tile; 0: interleaved 16-byte rows)
  int group_rows; // rows (pixels) between consecutive 8-row groups (halo_w)
  int shift;      // start row
  int bo_policy;  // 0: base_offset = 0, 1: (start >> 7) & 7
  int n;          // MMA N
  int lbo_rows;   // no-swizzle only: rows between the two 8-element K pieces
};

__device__ __forceinline__ uint32_t swz_chunk(uint32_t addr, int swz) {  // physical address of a 16-byte chunk
  if (swz == 128) return addr ^ (((addr >> 7) & 7u) << 4);
  if (swz == 64) return addr ^ (((addr >> 7) & 3u) << 4);
  if (swz == 32) return addr ^ (((addr >> 7) & 1u) << 4);
  return addr;
}

__global__ void __launch_bounds__(128) probe(Variant v, int iters, int* out_err, long long* out_cyc, float* out_dump) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowb = v.swz ? v.swz : 16;
  const int nrows = v.shift + 16 * v.group_rows + 8;
  // A: logical row p, element k (k < rowb/2 ... for no-swizzle: piece kk at p*16 + kk*lbo)
  const uint32_t a_off = 0, b_off = 96 * 1024;
  // value of A[p][k] = p + k/64  (bf16-exact for p < 256? keep p small: use (p % 61) + k * 0.015625)
  for (int i = tid; i < nrows * 16; i += 128) {
    const int p = i / 16, k = i % 16;
    const float val = (float)(p % 61) + (float)k * 0.015625f;
    uint32_t addr;
    if (v.swz) addr = swz_chunk(sbase + a_off + p * rowb + (k / 8) * 16, v.swz) + (k % 8) * 2;
    else addr = sbase + a_off + p * 16 + (k / 8) * (v.lbo_rows * 16) + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + (addr - sbase)) = __float2bfloat16_rn(val);
  }
  // B: N rows x 16 k, identity on k (B[n][k] = n == k), canonical SWIZZLE_32B tile (row = 32 bytes), n <= 16 useful
  for (int i = tid; i < v.n * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    const uint32_t addr = swz_chunk(sbase + b_off + n * 32 + (k / 8) * 16, 32) + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + (addr - sbase)) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 256);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // descriptor
  const uint32_t start = sbase + a_off + v.shift * rowb;
  uint64_t adesc;
  if (v.swz) {
    adesc = make_kmajor_desc(start, rowb);
    adesc &= ~((uint64_t)0x3fff << 32);
    adesc |= (uint64_t)(((uint32_t)(v.group_rows * rowb) >> 4) & 0x3fffu) << 32;
    if (v.bo_policy) adesc |= (uint64_t)((start >> 7) & 7u) << 49;
  } else {
    adesc = make_nosw_desc(start, v.lbo_rows * 16, v.group_rows * 16);
  }
  const uint64_t bdesc = make_kmajor_desc(sbase + b_off, 32);
  uint32_t idesc = 0;
  idesc |= 1u << 4; idesc |= 1u << 7; idesc |= 1u << 10; idesc |= (uint32_t)(v.n >> 3) << 17; idesc |= (uint32_t)(128 >> 4) << 24;
  if (tid == 0) {
    umma_f16(tmem, adesc, bdesc, idesc, 0);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
  }
  __syncthreads();
  tc_fence_after();
  // check D[m][n] == A[shift + (m/8)*group_rows + m%8][n]
  {
    float d[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), d);
    const int m = tid;
    const int p = v.shift + (m >> 3) * v.group_rows + (m & 7);
    int err = 0;
    for (int n = 0; n < 16; ++n) {
      const float want = __bfloat162float(__float2bfloat16_rn((float)(p % 61) + (float)n * 0.015625f));
      if (d[n] != want) ++err;
      if (out_dump && m < 24 && n < 2) out_dump[m * 2 + n] = d[n];
    }
    if (err) atomicAdd(out_err, err);
  }
  tc_fence_before();
  __syncthreads();
  // timing: `iters` back-to-back MMAs
  if (tid == 0) {
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_f16(tmem + 32, adesc, bdesc, idesc, i != 0);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 1);
    const long long t1 = clock64();
    *out_cyc = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  std::vector<Variant> vs;
  // swizzled, canonical (group = 8 rows) and halo strides, with shifts
  for (int swz : {128, 64, 32}) {
    for (int gr : {8, 16, 24, 10}) {
      for (int sh : {0, 1, 2, 3, 17, 18}) {
        for (int bo : {0, 1}) vs.push_back({swz, gr, sh, bo, 16, 0});
      }
    }
  }
  for (int gr : {8, 10}) for (int sh : {0, 1, 11}) vs.push_back({0, gr, sh, 0, 16, 8 * 0 + 180});
  vs.push_back({0, 8, 0, 0, 16, 128});  // canonical no-swizzle: pieces 128 rows apart
  int* d_err; long long* d_cyc; float* d_dump;
  cudaMalloc(&d_err, 4); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_dump, 48 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("swz group shift bo  N | errors  cyc/MMA\n");
  for (const Variant& v : vs) {
    for (int n : {16, 128}) {
      Variant w = v; w.n = n;
      if (n != 16 && !(v.shift == 0 || v.shift == 1)) continue;
      cudaMemset(d_err, 0, 4);
      const int iters = 512;
      probe<<<1, 128, 200 * 1024>>>(w, iters, d_err, d_cyc, d_dump);
      cudaError_t e = cudaDeviceSynchronize();
      int err; long long cyc;
      cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("%3d %5d %5d %2d %3d | %6d %8.1f %s\n", w.swz, w.group_rows, w.shift, w.bo_policy, w.n, n == 16 ? err : -1,
             (double)cyc / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
