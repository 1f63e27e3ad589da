// Micro-benchmark behind the LSTM exchange design: how long does one "step" of an all-to-all h exchange take
// inside a 16-CTA cluster?  Every CTA owns 32 columns of a [64 rows x 512] bf16 tile that each of the 16 CTAs
// needs in full: per step each of 128 threads (8 warps x 16 lanes) stores 32 B into all 16 CTAs' shared memory
// (st.shared::cluster), then the cluster synchronises (barrier.cluster release/acquire).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_exchange dsmem_exchange.cu && ./dsmem_exchange
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
  uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

template <int CS, int MODE>
__global__ void __launch_bounds__(256, 1) exchange_kernel(int steps, unsigned long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];   // [64 rows][512] bf16 = 64 KB
  const uint32_t rank = ctarank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool active = lane < 16;
  const int row = (warp & 3) * 16 + lane;             // 64 rows over 4 lane quadrants
  const int half = warp >> 2;                         // which 16 of this CTA's 32 columns
  const uint32_t base = smem_u32(smem);
  // K-major 128B-swizzled tile: k-block kb = 64 columns = 128 B per row, 8-row groups 1024 B apart
  const int col0 = rank * (512 / CS) + half * (256 / CS);       // first column this thread writes
  const int kb = col0 / 64, chunk = (col0 % 64) / 8;            // 16-byte chunk inside the 128-byte row
  const uint32_t off = kb * (64 * 128) + (row >> 3) * 1024 + (row & 7) * 128;
  cluster_sync();
  unsigned long long t0 = 0;
  uint32_t acc = 0;
  for (int s = 0; s < steps; ++s) {
    if (s == 8 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (active) {
      uint4 v = make_uint4(s, row, rank, acc);
      constexpr int NCH = (512 / CS) / 2 / 8;         // 16-byte chunks per thread
#pragma unroll
      for (int d = 0; d < CS; ++d) {
        const uint32_t dst = mapa(base + off, (rank + d) % CS);
        if (MODE == 0 || d == 0) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) st_cluster_v4(dst + (((chunk + c) ^ (row & 7)) << 4), v);
        }
      }
    }
    cluster_sync();
    acc += reinterpret_cast<uint32_t*>(smem)[(threadIdx.x * 67 + s) & 16383];   // consume something
  }
  if (threadIdx.x == 0) {
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[blockIdx.x] = (t1 - t0);
  }
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int CS, int MODE>
void run(const char* name, int clusters) {
  auto kern = exchange_kernel<CS, MODE>;
  const int smem = 224 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (CS > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS * clusters); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int active = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&active, kern, &cfg);
  printf("%-28s cluster %2d smem 224KB: max active clusters %d (%s)\n", name, CS, active, cudaGetErrorString(e));
  if (e != cudaSuccess || active < clusters) { cudaGetLastError(); return; }
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 8 * CS * clusters); cudaMalloc(&sink, 4 * 256 * CS * clusters);
  const int steps = 1008;
  e = cudaLaunchKernelEx(&cfg, kern, steps, out, sink);
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e != cudaSuccess || e2 != cudaSuccess) { printf("  launch failed: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
  unsigned long long h[256];
  cudaMemcpy(h, out, 8 * CS * clusters, cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < CS * clusters; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("  %d clusters: %.0f ns per step (%.1f KB received per CTA per step)\n", clusters, mx / (steps - 8),
         MODE == 0 ? 64.0 : 64.0 / CS);
  cudaFree(out); cudaFree(sink);
}

int main() {
  run<16, 0>("all-to-all push", 8);
  run<16, 0>("all-to-all push", 1);
  run<16, 1>("barrier only (local store)", 8);
  run<8, 0>("all-to-all push", 16);
  run<8, 1>("barrier only (local store)", 16);
  run<4, 0>("all-to-all push", 32);
  return 0;
}
