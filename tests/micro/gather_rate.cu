// gather_rate.cu — microbenchmark: how fast can one SM gather 128-byte lines (16 doubles of random rows of a
// [rows][LD] fp64 matrix) into a swizzled shared-memory tile?  (a) cp.async 16 B x 8 lanes per line (what the
// CD kernels do), (b) TMA tile copies, one 16 x 1 box per line, (c) TMA tile::gather4, four lines per
// instruction.  Decides whether TMA staging is worth building into the CD kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_rate gather_rate.cu && ./gather_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int LD = 128;          // doubles per row
constexpr int ROWS_PER_TILE = 128;
constexpr int WARPS = 8;         // warps per CTA, each with its own tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4_2d(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// MODE 0: cp.async; 1: TMA tile boxes; 2: TMA gather4
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1)
gather_kernel(const double* __restrict__ Y, const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map4,
              const int* __restrict__ ids, int tiles_per_warp, double* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* tile = smem + (size_t)warp * ROWS_PER_TILE * 128;
  __shared__ uint64_t bars[WARPS];
  if (lane == 0) mbar_init(&bars[warp], 1);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  const int gw = blockIdx.x * WARPS + warp;
  double acc = 0.0;
  uint32_t parity = 0;
  for (int t = 0; t < tiles_per_warp; t++) {
    const int* my = ids + ((size_t)gw * tiles_per_warp + t) * ROWS_PER_TILE;
    const int fb = t & 7;
    if (MODE == 0) {
      const int c = lane & 7;
      for (int r = lane >> 3; r < ROWS_PER_TILE; r += 4) {
        const double* src = Y + (size_t)my[r] * LD + fb * 16 + c * 2;
        cp_async16(tile + r * 128 + (((c ^ r) & 7) << 4), src);
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
      __syncwarp();
    } else if (MODE == 1) {
      if (lane == 0) mbar_expect_tx(&bars[warp], ROWS_PER_TILE * 128);
      __syncwarp();
      for (int r = lane; r < ROWS_PER_TILE; r += 32) tma_tile_2d(tile + r * 128, &map1, fb * 16, my[r], &bars[warp]);
      mbar_wait(&bars[warp], parity);
      parity ^= 1;
    } else {
      if (lane == 0) mbar_expect_tx(&bars[warp], ROWS_PER_TILE * 128);
      __syncwarp();
      for (int g = lane; g < ROWS_PER_TILE / 4; g += 32)
        tma_gather4_2d(tile + g * 512, &map4, fb * 16, my[4 * g], my[4 * g + 1], my[4 * g + 2], my[4 * g + 3], &bars[warp]);
      mbar_wait(&bars[warp], parity);
      parity ^= 1;
    }
    // touch the tile so the copies cannot be dropped: one double per lane and row group
    for (int r = lane; r < ROWS_PER_TILE; r += 32) acc += *reinterpret_cast<double*>(tile + r * 128 + 8);
    __syncwarp();
  }
  if (acc == 1.2345) out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 4 * 1000 * 1000;   // 4 GB of rows: far larger than L2
  double* Y; cudaMalloc(&Y, (size_t)rows * LD * 8); cudaMemset(Y, 0, (size_t)rows * LD * 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, tiles_per_warp = 64;
  const size_t nids = (size_t)sms * WARPS * tiles_per_warp * ROWS_PER_TILE;
  std::vector<int> h(nids);
  srand(1);
  for (auto& v : h) v = (int)(((unsigned)rand() * 2654435761u) % (unsigned)rows);
  int* ids; cudaMalloc(&ids, nids * 4); cudaMemcpy(ids, h.data(), nids * 4, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, 8);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map1, map4;
  cuuint64_t gdim[2] = {LD, (cuuint64_t)rows}, gstr[1] = {LD * 8};
  cuuint32_t box1[2] = {16, 1}, estr[2] = {1, 1};
  CUresult r1 = encode(&map1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, Y, gdim, gstr, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r4 = encode(&map4, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, Y, gdim, gstr, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r1, (int)r4);
  const size_t smem = (size_t)WARPS * ROWS_PER_TILE * 128 + 1024;
  cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const double bytes = (double)nids * 128;
  for (int mode = 0; mode < 3; mode++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) gather_kernel<0><<<sms, WARPS * 32, smem>>>(Y, map1, map4, ids, tiles_per_warp, out);
      if (mode == 1) gather_kernel<1><<<sms, WARPS * 32, smem>>>(Y, map1, map4, ids, tiles_per_warp, out);
      if (mode == 2) gather_kernel<2><<<sms, WARPS * 32, smem>>>(Y, map1, map4, ids, tiles_per_warp, out);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("mode %d (%s) rep %d: %s  %.3f ms  %.1f GB/s gathered lines\n", mode,
             mode == 0 ? "cp.async 16B x 8" : (mode == 1 ? "TMA tile 16x1 boxes" : "TMA tile::gather4"), rep,
             cudaGetErrorString(err), ms, bytes / ms / 1e6);
      if (err != cudaSuccess) { cudaGetLastError(); break; }
    }
  }
  return 0;
}
