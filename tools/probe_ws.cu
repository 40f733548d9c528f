// Probe: tcgen05.mma.ws (weight-stationary: the B operand stays in a collector buffer) for the conv's MMA shapes.
// The dz-folded conv issues, per (K chunk, tap), one MMA per input z-plane with the SAME weight block and a different
// activation plane; at N = 96 the regular MMA is bound by the shared-memory operand port (A 32 + B 24 cycles vs 48
// of math).  Questions: (1) does .ws with N = 96 / 192 compute the same D as the regular MMA (same TMEM layout)?
// (2) what does an MMA cost when B is reused from the collector?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_ws tools/probe_ws.cu
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../body-and-organ-analysis_b200/csrc/ptx.cuh"

using namespace boa;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

// mode 0: regular MMA.  1: .ws, fill every time (no reuse).  2: .ws, fill / use ... / lastuse over groups of `group`.
__device__ __forceinline__ void mma_variant(int mode, int pos, int group, uint32_t d, uint64_t a, uint64_t b,
                                            uint32_t idesc, uint32_t acc) {
  if (mode == 0) {
    umma_f16(d, a, b, idesc, acc);
    return;
  }
  const int which = mode == 1 ? 3 : (pos == 0 ? 0 : (pos == group - 1 ? 2 : 1));
  if (which == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else if (which == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else if (which == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
                 "r"(acc) : "memory");
  }
}

// smem: A region 64 KB (random fp16), B region 16 KB.  Each MMA j of a group reads A at a different (16-byte shifted,
// like a conv tap) start and writes its own D column block, all with the same B.
__global__ void __launch_bounds__(128, 1)
ws_probe(int N, int mode, int group, int iters, int check, float* __restrict__ out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 80 * 1024 / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    reinterpret_cast<__half*>(smem)[i] = __float2half((float)((int)((h >> 20) & 31) - 16) * 0.03125f);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const int nd = 512 / N >= 2 ? 2 : 1;  // distinct D blocks (first MMA of a block overwrites)
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
    const uint64_t ad0 = umma_desc(a0, 17280, 160);
    const uint64_t bd0 = umma_desc(b0, N * 16, 128);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += group) {
      for (int j = 0; j < group; ++j) {
        const uint64_t ad = ad0 + (uint64_t)(j * 180);  // next z-plane of a 10 x 18 halo tile
        const uint32_t dc = tbase + (uint32_t)((j % nd) * N);
        const uint32_t acc = (it == 0 && j < nd) ? 0u : 1u;
        if (elect_one()) mma_variant(mode, j, group, dc, ad, bd0, idesc, acc);
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (check && blockIdx.x == 0) {  // dump D: [128 rows][nd * N columns]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < nd * N; c += 32) {
      uint32_t v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int k = 0; k < 32; ++k) out[(size_t)(warp * 32 + lane) * 512 + c + k] = __uint_as_float(v[k]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  const int smem = 80 * 1024;
  CK(cudaFuncSetAttribute(ws_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float* dOut;
  long long* dCyc;
  CK(cudaMalloc(&dOut, 3 * 128 * 512 * 4));
  CK(cudaMalloc(&dCyc, nsm * 8));
  std::vector<float> ref(128 * 512), got(128 * 512);
  std::vector<long long> hc(nsm);
  const char* names[] = {"regular", ".ws no collector hint", ".ws fill/use/lastuse"};
  printf("tcgen05.mma.ws probe: M=128, K=16, fp16, groups of 8 MMAs sharing B (A = 8 z-planes of a halo tile)\n");
  printf("%5s %26s %10s %12s\n", "N", "variant", "cyc/mma", "vs regular");
  for (int N : {64, 96, 128, 192}) {
    for (int mode = 0; mode < 3; ++mode) {
      // correctness: 16 MMAs (2 groups), dump D
      CK(cudaMemset(dOut, 0, 128 * 512 * 4));
      ws_probe<<<1, 128, smem>>>(N, mode, 8, 16, 1, dOut, dCyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%5d %26s  FAILED: %s\n", N, names[mode], cudaGetErrorString(e));
        return 1;
      }
      CK(cudaMemcpy(mode == 0 ? ref.data() : got.data(), dOut, 128 * 512 * 4, cudaMemcpyDeviceToHost));
      size_t bad = 0;
      double maxd = 0;
      if (mode > 0)
        for (size_t i = 0; i < ref.size(); ++i) {
          const double d = fabs((double)ref[i] - (double)got[i]);
          maxd = d > maxd ? d : maxd;
          bad += d > 1e-3;
        }
      // rate
      for (int rep = 0; rep < 2; ++rep) {
        ws_probe<<<nsm, 128, smem>>>(N, mode, 8, 4000, 0, dOut, dCyc);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(hc.data(), dCyc, nsm * 8, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (auto c : hc) avg += c;
      avg /= nsm * 4000.0;
      if (mode == 0) printf("%5d %26s %10.2f %12s  (|D| max %.3f)\n", N, names[mode], avg, "-", [&] { double m = 0; for (float x : ref) m = fabs(x) > m ? fabs(x) : m; return m; }());
      else printf("%5d %26s %10.2f %12s  (%zu elements differ, max diff %.2e)\n", N, names[mode], avg, bad ? "MISMATCH" : "same D", bad, maxd);
    }
  }
  return 0;
}
