// Stand-alone probe of the PTX building blocks lbm_stream.cuh uses (mbarrier + cp.async.bulk.tensor),
// one tiny kernel per feature, so that a failure on the GPU box names the feature.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#include "../../hpc-lattice-boltzmann_b200/csrc/lbm_stream.cuh"

using namespace lbm;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__global__ void k_barrier_only(int* out)
{
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(b, 32); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  mbar_arrive(b);
  mbar_wait(b, 0);
  if (threadIdx.x == 0) out[0] = 1;
}

__device__ int probe_wait(uint32_t bar, uint32_t parity)
{
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return 0;
    if (clock64() - t0 > 400000000LL) return 1;
  }
}

template <int MODE>   // 0: 2D float, 1: 3D float, 2: 2D u8
__global__ void k_tma(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int z)
{
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * 512);
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) { mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(b, MODE == 2 ? 4 * 128 : 4 * 512);
    if (MODE == 0) tma_load_2d(smem_u32(sm), &tm, x, y, b);
    if (MODE == 1) tma_load_3d(smem_u32(sm), &tm, x, y, z, b);
    if (MODE == 2) tma_load_2d(smem_u32(sm), &tm, x, y, b);
  }
  if (probe_wait(b, 0)) { if (threadIdx.x == 0) out[0] = -12345.0f; return; }
  if (MODE == 2) { for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) out[i] = (float)sm[i]; }
  else { for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i]; }
}

int main()
{
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaFree(0));
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int nx = 256, ny = 40, nz = 18;
  std::vector<float> h((size_t)nx * ny * nz);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
  float *d, *out;
  CHECK(cudaMalloc(&d, h.size() * 4));
  CHECK(cudaMalloc(&out, 4 * 128 * 4));
  CHECK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> res(4 * 128);
  int* flag;
  CHECK(cudaMalloc(&flag, 4));

  printf("[1] mbarrier only\n");
  k_barrier_only<<<1, 32>>>(flag);
  CHECK(cudaDeviceSynchronize());
  printf("  ok\n");

  printf("[2] 2D float box 128x4 at (8, 3)\n");
  {
    CUtensorMap tm;
    cuuint64_t dims[2] = {nx, (cuuint64_t)ny * nz}; cuuint64_t str[1] = {nx * 4};
    cuuint32_t box[2] = {128, 4}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode -> %d\n", (int)r);
    k_tma<0><<<1, 128, 4 * 512 + 64>>>(tm, out, 8, 3, 0);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost));
    printf("  got %g %g %g (want %g %g %g)\n", res[0], res[1], res[128], h[3 * nx + 8], h[3 * nx + 9], h[4 * nx + 8]);
  }
  {
    CUtensorMap tm;
    cuuint64_t dims[2] = {nx, (cuuint64_t)ny}; cuuint64_t str[1] = {nx * 4};
    cuuint32_t box[2] = {128, 4}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("[2b] encode -> %d\n", (int)r);
    const int xs[] = {8, -8, -4, 200, 8, 4}, ys[] = {38, 3, 3, 3, -1, 3};
    for (int t = 0; t < 6; t++) {
      k_tma<0><<<1, 128, 4 * 512 + 64>>>(tm, out, xs[t], ys[t], 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  2D at (%d,%d): %s\n", xs[t], ys[t], cudaGetErrorString(e)); return 1; }
      CHECK(cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost));
      printf("  2D at (%d,%d): got %g %g %g %g\n", xs[t], ys[t], res[0], res[8], res[128 + 8], res[3 * 128 + 127]);
    }
  }
  printf("[3a] 3D float box 128x4x1 at (8, 2, 7)\n");
  {
    CUtensorMap tm;
    cuuint64_t dims[3] = {nx, ny, nz}; cuuint64_t str[2] = {nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {128, 4, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode -> %d\n", (int)r);
    k_tma<1><<<1, 128, 4 * 512 + 64>>>(tm, out, 8, 2, 7);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  sync -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  printf("[3] 3D float box 128x4x1 at (-5, 2, 7)\n");
  {
    CUtensorMap tm;
    cuuint64_t dims[3] = {nx, ny, nz}; cuuint64_t str[2] = {nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {128, 4, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode -> %d\n", (int)r);
    k_tma<1><<<1, 128, 4 * 512 + 64>>>(tm, out, -4, 2, 7);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost));
    const size_t o = (size_t)7 * nx * ny + 2 * nx;
    printf("  got %g %g %g (want 0 %g %g)\n", res[0], res[4], res[128 + 5], h[o + 0], h[o + nx + 1]);
  }
  printf("[4] 2D u8 box 128x4 at (16, 1)\n");
  {
    std::vector<uint8_t> hb((size_t)nx * ny);
    for (size_t i = 0; i < hb.size(); i++) hb[i] = (uint8_t)(i * 7);
    uint8_t* db;
    CHECK(cudaMalloc(&db, hb.size()));
    CHECK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[2] = {nx, ny}; cuuint64_t str[1] = {nx};
    cuuint32_t box[2] = {128, 4}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, db, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode -> %d\n", (int)r);
    k_tma<2><<<1, 128, 4 * 512 + 64>>>(tm, out, 16, 1, 0);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost));
    printf("  got %g %g (want %d %d)\n", res[0], res[129], hb[nx + 16], hb[2 * nx + 17]);
  }
  printf("probe done\n");
  return 0;
}
