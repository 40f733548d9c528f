// Does a small-footprint ("thin") streaming kernel share an SM with a persistent CTA that holds most of the register
// file and shared memory?  hog<>: 148 x 192 threads, ~REGS live registers, SMEM bytes of dynamic shared memory, spins
// for a fixed time.  stream_copy: 148 x 256 threads, <= 64 registers.  Prints the copy's duration alone and next to
// each hog variant, and how many copy blocks started while the hog was resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/coresidency_probe tools/coresidency_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int NLIVE>
__global__ void __launch_bounds__(192, 1) hog(unsigned long long ns, float* out, unsigned long long* t_end) {
  extern __shared__ float sm[];
  float acc[NLIVE];
#pragma unroll
  for (int i = 0; i < NLIVE; ++i) acc[i] = threadIdx.x * 0.001f + i;
  const unsigned long long t0 = gtime();
  sm[threadIdx.x] = acc[0];
  while (gtime() - t0 < ns) {
#pragma unroll
    for (int i = 0; i < NLIVE; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
  }
  float s = sm[threadIdx.x];
#pragma unroll
  for (int i = 0; i < NLIVE; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) t_end[blockIdx.x] = gtime();
}

__global__ void __launch_bounds__(256, 4) stream_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n,
                                                     unsigned long long* t_start) {
  if (threadIdx.x == 0) t_start[blockIdx.x] = gtime();
  const size_t stride = (size_t)gridDim.x * 256 * 8;
  for (size_t i = (size_t)blockIdx.x * 256 * 8 + threadIdx.x; i < n; i += stride) {
    uint4 r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i + u * 256 < n) r[u] = __ldcs(src + i + u * 256);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i + u * 256 < n) dst[i + u * 256] = r[u];
  }
}

template <typename K>
static void run_case(const char* name, K hog_kernel, size_t smem, int copy_grid, int copy_carveout, const uint4* src,
                     uint4* dst, size_t n, float* d_out, unsigned long long* d_tend, unsigned long long* d_tstart) {
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  CK(cudaFuncSetAttribute(hog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaFuncSetAttribute(stream_copy, cudaFuncAttributePreferredSharedMemoryCarveout, copy_carveout));
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
  // copy alone
  stream_copy<<<copy_grid, 256, 0, s2>>>(src, dst, n, d_tstart);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0, s2);
  stream_copy<<<copy_grid, 256, 0, s2>>>(src, dst, n, d_tstart);
  cudaEventRecord(e1, s2);
  CK(cudaDeviceSynchronize());
  float alone = 0;
  cudaEventElapsedTime(&alone, e0, e1);
  // hog (3 ms) first, copy right behind it on the other stream
  const unsigned long long hog_ns = 3000000ull;
  cudaEventRecord(e2, s1);
  hog_kernel<<<148, 192, smem, s1>>>(hog_ns, d_out, d_tend);
  cudaEventRecord(e3, s1);
  cudaEventRecord(e0, s2);
  stream_copy<<<copy_grid, 256, 0, s2>>>(src, dst, n, d_tstart);
  cudaEventRecord(e1, s2);
  CK(cudaDeviceSynchronize());
  float both_copy = 0, hog_ms = 0;
  cudaEventElapsedTime(&both_copy, e0, e1);
  cudaEventElapsedTime(&hog_ms, e2, e3);
  static unsigned long long h_tend[148], h_tstart[4096];
  CK(cudaMemcpy(h_tend, d_tend, sizeof(unsigned long long) * 148, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_tstart, d_tstart, sizeof(unsigned long long) * copy_grid, cudaMemcpyDeviceToHost));
  unsigned long long first_end = ~0ull;
  for (int i = 0; i < 148; ++i) first_end = h_tend[i] < first_end ? h_tend[i] : first_end;
  int early = 0;
  for (int i = 0; i < copy_grid; ++i) early += h_tstart[i] < first_end;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, hog_kernel);
  printf("%-34s hog regs %3d smem %6zu | copy grid %4d carveout %3d: alone %.3f ms, behind the hog %.3f ms (hog %.3f ms); "
         "%d / %d copy blocks started while every hog CTA was still resident\n",
         name, fa.numRegs, smem, copy_grid, copy_carveout, alone, both_copy, hog_ms, early, copy_grid);
  cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

int main() {
  const size_t n = (size_t)64 << 20;  // 64 Mi uint4 = 1 GiB each way
  uint4 *src, *dst;
  float* d_out;
  unsigned long long *d_tend, *d_tstart;
  CK(cudaMalloc(&src, n * 16)); CK(cudaMalloc(&dst, n * 16));
  CK(cudaMemset(src, 1, n * 16));
  CK(cudaMalloc(&d_out, 148 * 192 * 4)); CK(cudaMalloc(&d_tend, 148 * 8)); CK(cudaMalloc(&d_tstart, 4096 * 8));
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, stream_copy);
  printf("stream_copy: %d registers\n", fa.numRegs);
  const int MAXS = cudaSharedmemCarveoutMaxShared;
  run_case("small hog (32 live, 16 KB)", hog<32>, 16 << 10, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("register hog (200 live, 16 KB)", hog<200>, 16 << 10, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("smem hog (32 live, 170 KB)", hog<32>, 170496 + 128, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("smem hog (32 live, 221 KB)", hog<32>, 225920, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("conv-like (200 live, 170 KB)", hog<200>, 170496 + 128, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("conv-like (200 live, 221 KB)", hog<200>, 225920, 148, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  run_case("conv-like, default carveout", hog<200>, 170496 + 128, 148, -1, src, dst, n, d_out, d_tend, d_tstart);
  run_case("conv-like, copy grid 148*8", hog<200>, 170496 + 128, 148 * 8, MAXS, src, dst, n, d_out, d_tend, d_tstart);
  return 0;
}
