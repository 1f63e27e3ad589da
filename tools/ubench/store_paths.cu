// Micro-benchmark: how fast can one SM push a [128 channels x 256 steps] bf16 output tile (64 KB, rows of 512 B in
// a [C][T] tensor) to global memory, by path?  All 148 SMs store disjoint tiles, data comes from shared memory.
//   a) TMA tensor store, 128B-swizzled boxes of 64 steps x 32 channels (4 KB; what conv_gemm.cu does)
//   b) cp.async.bulk (1D) of one row piece per lane: 128 / 256 / 512 bytes
//   c) st.global.v4: a warp writes one 512-byte row piece per instruction (fully coalesced)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_paths store_paths.cu -lcuda && ./store_paths
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kC = 2048, kT = 768;   // one "batch element": [kC][kT] bf16; tiles of 128 ch x 256 steps

template <int MODE, int PIECE>
__global__ void __launch_bounds__(256, 1) store_kernel(const __grid_constant__ CUtensorMap tm, unsigned short* out,
                                                      int tiles_per_cta, unsigned long long* ns) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 64 KB tile image
  for (int i = threadIdx.x; i < 65536 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int it = 0; it < tiles_per_cta; ++it) {
    const int tile = blockIdx.x * tiles_per_cta + it;           // (b, m_tile, t_tile)
    const int t_tile = tile % 3, m_tile = (tile / 3) % 16, b = tile / 48;
    const int ch0 = m_tile * 128 + (warp & 3) * 32, t0c = t_tile * 256 + (warp >> 2) * 128;  // warp: 32 ch x 128 steps
    if (MODE == 0) {
      if (lane == 0) {
        for (int c = 0; c < 2; ++c) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(smem + warp * 8192 + c * 4096)), "r"(t0c + c * 64),
                         "r"(ch0), "r"(b) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else if (MODE == 1) {
      // each lane: its channel row, 256 bytes (128 steps) in PIECE-byte bulk copies
      unsigned short* g = out + (static_cast<long long>(b) * kC + ch0 + lane) * kT + t0c;
      for (int o = 0; o < 256; o += PIECE) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(g + o / 2), "r"(smem_u32(smem + warp * 8192 + lane * 256 + o)), "r"(PIECE) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
      // warp-coalesced: instruction i writes 512 contiguous bytes: rows (2 i, 2 i + 1) x 256 bytes
      for (int i = 0; i < 16; ++i) {
        const int row = 2 * i + (lane >> 4);
        unsigned short* g = out + (static_cast<long long>(b) * kC + ch0 + row) * kT + t0c + (lane & 15) * 8;
        *reinterpret_cast<uint4*>(g) = reinterpret_cast<const uint4*>(smem + warp * 8192 + row * 256)[lane & 15];
      }
    }
    __syncwarp();
  }
  if (MODE != 2 && (MODE == 1 || lane == 0)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  unsigned long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) ns[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int PIECE>
void run(const char* name, const CUtensorMap& tm, unsigned short* out, unsigned long long* ns, int B) {
  auto kern = store_kernel<MODE, PIECE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int tiles = B * 48, ctas = 148, per = tiles / ctas;
  for (int rep = 0; rep < 2; ++rep) kern<<<ctas, 256, 65536>>>(tm, out, per, ns);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, ns, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-44s %s  %.2f us per 64 KB tile per SM = %.1f GB/s per SM, %.2f TB/s chip\n", name, cudaGetErrorString(e),
         mx / per / 1e3, 65536.0 / (mx / per), 65536.0 * per * ctas / mx / 1e3);
}

int main() {
  const int B = 74;   // 74 x 48 tiles = 24 per SM; 74 x 2048 x 768 x 2 B = 233 MB
  unsigned short* out; unsigned long long* ns;
  cudaMalloc(&out, size_t(B) * kC * kT * 2); cudaMalloc(&ns, 148 * 8);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t dims[3] = {kT, kC, cuuint64_t(B)}; const cuuint64_t strides[2] = {kT * 2, cuuint64_t(kC) * kT * 2};
  const cuuint32_t box[3] = {64, 32, 1}, ones[3] = {1, 1, 1};
  ((EncodeFn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  run<0, 0>("TMA tensor store, boxes 64 x 32 (128 B rows)", tm, out, ns, B);
  run<1, 128>("cp.async.bulk 1D, 128 B per copy", tm, out, ns, B);
  run<1, 256>("cp.async.bulk 1D, 256 B per copy", tm, out, ns, B);
  run<2, 0>("st.global.v4, 512 B per warp instruction", tm, out, ns, B);
  return 0;
}
