// Micro-benchmark: per-SM write throughput to global memory (L2) with STG.128 vs TMA bulk stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu && ./store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void stg_kernel(float4* out, int bytes_per_cta, int reps, long long* cycles) {
  float4* base = out + (size_t)blockIdx.x * (bytes_per_cta / 16);
  const int n = bytes_per_cta / 16;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (int i = threadIdx.x; i < n; i += blockDim.x) base[i] = make_float4(r, i, 0.f, 1.f);
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// rows of 128 B written by 8 lanes each, 4 rows per warp instruction, row pitch `pitch` bytes (GEMM epilogue pattern)
__global__ void stg_rows_kernel(char* out, int rows_per_cta, int pitch, int reps, long long* cycles) {
  char* base = out + (size_t)blockIdx.x * rows_per_cta * pitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (int row = warp * 4 + (lane >> 3); row < rows_per_cta; row += nw * 4)
      *reinterpret_cast<float4*>(base + (size_t)row * pitch + (lane & 7) * 16) = make_float4(r, row, 0.f, 1.f);
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void bulk_kernel(char* out, int bytes_per_cta, int chunk, int reps, long long* cycles) {
  extern __shared__ __align__(128) char sm[];
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    char* base = out + (size_t)blockIdx.x * bytes_per_cta;
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    for (int r = 0; r < reps; ++r)
      for (int off = 0; off < bytes_per_cta; off += chunk) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(s), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t total = 512u << 20;
  char* out; cudaMalloc(&out, total);
  long long* cyc; cudaMallocManaged(&cyc, 148 * sizeof(long long));
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char* name, int ctas, size_t bytes) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long mx = 0; for (int i = 0; i < ctas; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    printf("%-44s ctas %3d: %8lld cyc  %6.1f B/cyc/SM\n", name, ctas, mx, (double)bytes / mx);
  };
  for (int ctas : {8, 148}) {
    for (int kb : {128, 1024}) {
      const int bytes = kb << 10, reps = 4;
      char nm[96];
      for (int w = 0; w < 2; ++w) stg_kernel<<<ctas, 256, 0>>>((float4*)out, bytes, reps, cyc);
      snprintf(nm, 96, "STG.128 contiguous %d KB x%d, 256 thr", kb, reps); report(nm, ctas, (size_t)bytes * reps);
      for (int w = 0; w < 2; ++w) stg_kernel<<<ctas, 1024, 0>>>((float4*)out, bytes, reps, cyc);
      snprintf(nm, 96, "STG.128 contiguous %d KB x%d, 1024 thr", kb, reps); report(nm, ctas, (size_t)bytes * reps);
      for (int pitch : {128, 1024, 3072}) {
        const int rows = bytes / 128;
        if ((size_t)ctas * rows * pitch > total) continue;
        for (int w = 0; w < 2; ++w) stg_rows_kernel<<<ctas, 256, 0>>>(out, rows, pitch, reps, cyc);
        snprintf(nm, 96, "STG.128 rows of 128 B pitch %d, %d KB x%d", pitch, kb, reps); report(nm, ctas, (size_t)bytes * reps);
      }
      for (int chunk : {4096, 16384, 32768}) {
        cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
        for (int w = 0; w < 2; ++w) bulk_kernel<<<ctas, 128, chunk>>>(out, bytes, chunk, reps, cyc);
        snprintf(nm, 96, "bulk S2G chunk %d, %d KB x%d", chunk, kb, reps); report(nm, ctas, (size_t)bytes * reps);
      }
    }
  }
  return 0;
}
