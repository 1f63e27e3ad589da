// Micro-benchmark: what does ONE global->shared copy cost an SM, by size and by path?  Every warp of every SM streams
// rows of SIZE bytes (each row from a different 4 KB-aligned place of a 1 GB buffer, i.e. from HBM) into its own
// shared-memory buffers, DEPTH rows in flight, and only waits for them -- no compute.
//   a) cp.async.bulk (1D bulk copy, mbarrier completion): what dw_bulk_kernel stages its rows with
//   b) 16-byte cp.async (LDGSTS), SIZE / 512 warp-wide instructions per row: what dw_mma_kernel uses
//   c) TMA tensor load of a [SIZE/2 x 8 rows] box (rows 4 KB apart): what dw_rows_kernel uses (8 rows per instruction)
// Prints rows per microsecond per SM and the aggregate GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_load_cost bulk_load_cost.cu -lcuda && ./bulk_load_cost
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

constexpr int kWarps = 8, kDepth = 2, kRowStride = 4096;   // bytes between rows in global memory

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) load_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* src,
                                                           int size, int rows_per_warp, long long n_rows_total, int pitch, int dst_off, int src_off) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[kWarps][kDepth];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_buf = MODE == 2 ? 8 * size : pitch;
  uint8_t* buf = smem + size_t(warp) * kDepth * per_buf + dst_off;
  if (lane < kDepth) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp][lane])) : "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long long gw = (static_cast<long long>(blockIdx.x) * kWarps + warp);
  const int step = MODE == 2 ? 8 : 1;     // rows per issue
  auto issue = [&](int i) {               // i = issue index of this warp
    const long long row = (gw * rows_per_warp + static_cast<long long>(i) * step) % n_rows_total;
    const int b = i % kDepth;
    if (MODE == 0) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][b])), "r"(size) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf + b * per_buf)), "l"(src + row * kRowStride + src_off), "r"(size), "r"(smem_u32(&bars[warp][b])) : "memory");
      }
    } else if (MODE == 1) {
      for (int o = lane * 16; o < size; o += 512)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + b * per_buf + o)), "l"(src + row * kRowStride + o) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][b])), "r"(8 * size) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(buf + b * per_buf)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bars[warp][b])), "r"(0), "r"(int(row)) : "memory");
      }
    }
  };
  const int n_issue = rows_per_warp / step;
  for (int i = 0; i < kDepth && i < n_issue; ++i) issue(i);
  for (int i = 0; i < n_issue; ++i) {
    const int b = i % kDepth;
    if (MODE == 1) {
      if (i + 1 < n_issue) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      while (!try_wait(&bars[warp][b], (i / kDepth) & 1)) {}
    }
    __syncwarp();
    if (i + kDepth < n_issue) issue(i + kDepth);
  }
}

template <int MODE>
static void run(const char* name, const CUtensorMap& tm, const uint8_t* src, int size, long long n_rows_total, int pitch = 0, int dst_off = 0, int src_off = 0) {
  if (pitch == 0) pitch = size;
  const int rows_per_warp = 512;
  const int ctas = 148 * 4;
  const size_t smem = size_t(kWarps) * kDepth * (MODE == 2 ? 8 * size : pitch) + 256;
  cudaFuncSetAttribute(load_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  load_kernel<MODE><<<ctas, kWarps * 32, smem>>>(tm, src, size, rows_per_warp, n_rows_total, pitch, dst_off, src_off);
  cudaEventRecord(e0);
  load_kernel<MODE><<<ctas, kWarps * 32, smem>>>(tm, src, size, rows_per_warp, n_rows_total, pitch, dst_off, src_off);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double rows = double(ctas) * kWarps * rows_per_warp;
  printf("%-44s %5d B/row: %7.1f us, %6.2f rows/us/SM (%5.0f clk/row/SM at 1.9 GHz), %6.0f GB/s  %s\n", name, size, ms * 1e3,
         rows / (ms * 1e3) / 148, 1.9e3 * 148 * (ms * 1e3) / rows, rows * size / (ms * 1e-3) / 1e9,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  const long long n_rows_total = 262144;      // x 4 KB = 1 GB
  uint8_t* src;
  cudaMalloc(&src, n_rows_total * kRowStride);
  cudaMemset(src, 1, n_rows_total * kRowStride);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Fn encode = reinterpret_cast<Fn>(p);
  for (int size : {128, 208, 512, 1504}) {
    CUtensorMap tm;
    const int elems = ((size / 2) + 7) / 8 * 8;           // 16-bit elements per row piece (16-byte multiple)
    const cuuint64_t dims[2] = {cuuint64_t(kRowStride / 2), cuuint64_t(n_rows_total)};
    const cuuint64_t strides[1] = {cuuint64_t(kRowStride)};
    const cuuint32_t box[2] = {cuuint32_t(elems > 256 ? 256 : elems), 8}, ones[2] = {1, 1};
    encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int sz = elems * 2;
    run<0>("cp.async.bulk, one copy per row", tm, src, sz, n_rows_total);
    run<1>("16-byte cp.async (LDGSTS), warp-wide", tm, src, sz, n_rows_total);
    if (elems <= 256) run<2>("TMA tensor load, box = row piece x 8 rows", tm, src, sz, n_rows_total);
  }
  {
    CUtensorMap tm{};
    // destination / source alignment of a 1504-byte bulk copy (dw_bulk_kernel: rows 2848 B apart, copy lands 96 B into its buffer)
    run<0>("cp.async.bulk 1504 B, dst pitch 1536 (128 B aligned)", tm, src, 1504, n_rows_total, 1536, 0, 0);
    run<0>("cp.async.bulk 1504 B, dst pitch 1536, dst + 96 B", tm, src, 1504, n_rows_total, 1664, 96, 0);
    run<0>("cp.async.bulk 1504 B, dst pitch 2848, dst + 96 B", tm, src, 1504, n_rows_total, 2848, 96, 0);
    run<0>("cp.async.bulk 1536 B, dst pitch 1536", tm, src, 1536, n_rows_total, 1536, 0, 0);
    run<0>("cp.async.bulk 1024 B, dst pitch 1024", tm, src, 1024, n_rows_total, 1024, 0, 0);
    run<0>("cp.async.bulk 2048 B, dst pitch 2048", tm, src, 2048, n_rows_total, 2048, 0, 0);
    run<0>("cp.async.bulk 1504 B, src + 16 B", tm, src, 1504, n_rows_total, 1536, 0, 16);
  }
  return 0;
}
