// Hardware probe for the assumptions the conv kernels rest on (run once on a B200, output kept in profiles/):
//  T1  TMA 5-D box load of a C8 tensor with negative start coordinates -> zero-filled halo in smem
//  T2  tcgen05.mma with SWIZZLE_NONE K-major descriptors addressing tap-shifted views of that halo tile
//      (start addresses that are only 16-byte aligned, SBO = 160 B) == a 3x3x3 conv, checked against the host
//  T3  MMA issue rate vs N (is N=32 limited by shared-memory operand bandwidth?)
//  T4  TMEM -> register read rate
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../body-and-organ-analysis_b200/csrc/ptx.cuh"
#include "../body-and-organ-analysis_b200/csrc/tmap.cuh"

using namespace boa;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int W = 8, H = 16, D = 4, CIN = 32, COUT = 32, CG = CIN / 8;
constexpr int XB = W + 2, YB = H + 2, ZB = D + 2;
constexpr int PLANE = XB * YB * ZB;          // positions per channel group
constexpr int A_BYTES = CG * PLANE * 16;     // 69120
constexpr int B_TILE = 2 * COUT * 16;        // one (tap, k16) B block: [kchunk][n][8] = 1024 B
constexpr int B_BYTES = 27 * (CIN / 16) * B_TILE;

__global__ void __launch_bounds__(192, 1)
probe_conv(const __grid_constant__ CUtensorMap tmapX, const __half* __restrict__ Bpacked, float* __restrict__ out,
           uint4* __restrict__ smem_dump) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmapX);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(&bars[0], A_BYTES + B_BYTES);
    tma_load_5d(sA, &tmapX, &bars[0], 0, -1, -1, -1, 0);
    bulk_load(sB, Bpacked, B_BYTES, &bars[0]);
  } else if (warp == 1 && lane == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, COUT);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int z = 0; z < D; ++z) {
      uint32_t acc = 0;
      for (int tap = 0; tap < 27; ++tap) {
        const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
        const int pos = ((z + dz) * YB + dy) * XB + dx;
        for (int k = 0; k < CIN / 16; ++k) {
          uint64_t ad = umma_desc(a0 + (2 * k) * PLANE * 16 + pos * 16, PLANE * 16, XB * 16);
          uint64_t bd = umma_desc(b0 + (tap * (CIN / 16) + k) * B_TILE, COUT * 16, 128);
          umma_f16(tbase + z * COUT, ad, bd, idesc, acc);
          acc = 1;
        }
      }
    }
    umma_commit(&bars[1]);
  } else if (warp >= 2) {
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const int q = warp & 3;
    for (int z = 0; z < D; ++z) {
      uint32_t v[32];
      tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + z * COUT, v);
      tmem_ld_wait();
      const int row = q * 32 + lane;
      for (int c = 0; c < 32; ++c) out[(z * 128 + row) * COUT + c] = __uint_as_float(v[c]);
    }
    // dump the TMA-written halo tile
    for (int i = threadIdx.x - 64; i < A_BYTES / 16; i += 128) smem_dump[i] = reinterpret_cast<uint4*>(sA)[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 128);
}

// ---------------------------------------------------------------- T3: MMA issue rate
__global__ void __launch_bounds__(128, 1)
mma_rate(int N, int iters, int a_off, int sbo, int lbo, int rotate_d, int rotate_a, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // fill smem with small fp16 values
  for (int i = threadIdx.x; i < 160 * 1024 / 2; i += blockDim.x)
    reinterpret_cast<__half*>(smem)[i] = __float2half(((i * 37) % 17 - 8) * 0.01f);
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
  if (threadIdx.x < 32) {
    // whole warp runs the loop (warp-uniform descriptor arithmetic stays in uniform registers);
    // one elected lane issues.
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(smem) + a_off;
    const uint32_t b0 = smem_u32(smem) + 96 * 1024;  // B region
    const int nd = rotate_d ? (512 / N) : 1;
    const uint64_t ad0 = umma_desc(a0, lbo, sbo);
    const uint64_t bd0 = umma_desc(b0, N * 16, 128);
    const uint32_t astep = rotate_a ? 7 : 0;  // in 16-byte units, added to the descriptor's address field
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint64_t ad = ad0 + (uint64_t)(j * astep);
        const uint64_t bd = bd0 + (uint64_t)((j & 3) * 512);
        const uint32_t dc = tbase + (nd > 1 ? (j & (nd - 1)) * N : 0);
        if (elect_one()) umma_f16(dc, ad, bd, idesc, 1);
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
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------- T4: TMEM read rate
__global__ void __launch_bounds__(128, 1) tmem_read_rate(int reps, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const int q = (threadIdx.x >> 5) & 3;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int c = 0; c < 512; c += 64) {
      uint32_t v[32], w[32];
      tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + c, v);
      tmem_ld32(tbase + ((uint32_t)(q * 32) << 16) + c + 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i]) * 1e-30f + __uint_as_float(w[i]) * 1e-30f;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d smem/block optin %zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);

  // ---------------- T1 + T2
  srand(7);
  std::vector<__half> hX((size_t)CG * D * H * W * 8);
  std::vector<float> fX(hX.size());
  for (size_t i = 0; i < hX.size(); ++i) {
    float v = ((rand() % 2001) - 1000) / 1000.f;
    hX[i] = __float2half(v);
    fX[i] = __half2float(hX[i]);
  }
  std::vector<float> fWt((size_t)COUT * CIN * 27);
  std::vector<__half> hB((size_t)B_BYTES / 2);
  for (size_t i = 0; i < fWt.size(); ++i) {
    float v = ((rand() % 2001) - 1000) / 4000.f;
    fWt[i] = __half2float(__float2half(v));
  }
  // pack B: [tap][k16][kchunk][n][8]
  for (int tap = 0; tap < 27; ++tap)
    for (int k = 0; k < CIN / 16; ++k)
      for (int kc = 0; kc < 2; ++kc)
        for (int n = 0; n < COUT; ++n)
          for (int e = 0; e < 8; ++e) {
            int ci = k * 16 + kc * 8 + e;
            hB[(size_t)((tap * (CIN / 16) + k) * B_TILE) / 2 + (kc * COUT + n) * 8 + e] =
                __float2half(fWt[((size_t)n * CIN + ci) * 27 + tap]);
          }
  __half *dX, *dB;
  float* dOut;
  uint4* dDump;
  CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dB, B_BYTES));
  CK(cudaMalloc(&dOut, D * 128 * COUT * 4));
  CK(cudaMalloc(&dDump, A_BYTES));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), B_BYTES, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0xff, D * 128 * COUT * 4));
  CUtensorMap tm;
  if (make_c8_tmap(&tm, dX, CG, D, H, W, XB, YB, ZB, CG)) return 3;
  const int smem_bytes = A_BYTES + B_BYTES + 64;
  CK(cudaFuncSetAttribute(probe_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_conv<<<1, 192, smem_bytes>>>(tm, dB, dOut, dDump);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hOut(D * 128 * COUT);
  std::vector<__half> hDump(A_BYTES / 2);
  CK(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hDump.data(), dDump, A_BYTES, cudaMemcpyDeviceToHost));
  // T1 check
  {
    size_t bad = 0;
    for (int cg = 0; cg < CG; ++cg)
      for (int z = 0; z < ZB; ++z)
        for (int y = 0; y < YB; ++y)
          for (int x = 0; x < XB; ++x)
            for (int e = 0; e < 8; ++e) {
              int gz = z - 1, gy = y - 1, gx = x - 1;
              float want = 0.f;
              if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
                want = fX[((((size_t)cg * D + gz) * H + gy) * W + gx) * 8 + e];
              float got = __half2float(hDump[((((size_t)cg * ZB + z) * YB + y) * XB + x) * 8 + e]);
              if (want != got) {
                if (bad < 5) printf("  T1 mismatch cg%d z%d y%d x%d e%d want %f got %f\n", cg, z, y, x, e, want, got);
                ++bad;
              }
            }
    printf("T1 TMA 5D halo load with OOB zero fill: %s (%zu mismatches)\n", bad ? "FAIL" : "PASS", bad);
  }
  // T2 check
  {
    double maxerr = 0, maxref = 0;
    size_t bad = 0;
    for (int z = 0; z < D; ++z)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int co = 0; co < COUT; ++co) {
            double acc = 0;
            for (int tap = 0; tap < 27; ++tap) {
              int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
              int gz = z + dz - 1, gy = y + dy - 1, gx = x + dx - 1;
              if (gz < 0 || gz >= D || gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
              for (int ci = 0; ci < CIN; ++ci)
                acc += (double)fX[((((size_t)(ci / 8) * D + gz) * H + gy) * W + gx) * 8 + (ci % 8)] *
                       fWt[((size_t)co * CIN + ci) * 27 + tap];
            }
            int row = y * 8 + x;
            float got = hOut[(z * 128 + row) * COUT + co];
            double err = fabs(got - acc);
            if (err > maxerr) maxerr = err;
            if (fabs(acc) > maxref) maxref = fabs(acc);
            if (!(err < 2e-3)) {
              if (bad < 5) printf("  T2 mismatch z%d y%d x%d co%d want %f got %f\n", z, y, x, co, acc, got);
              ++bad;
            }
          }
    printf("T2 tcgen05 tap-shifted no-swizzle conv: %s (max abs err %.3e, max |ref| %.3f, %zu bad)\n",
           bad ? "FAIL" : "PASS", maxerr, maxref, bad);
  }

  // ---------------- T3
  {
    const int nsm = prop.multiProcessorCount;
    long long* dCyc;
    CK(cudaMalloc(&dCyc, nsm * 8));
    std::vector<long long> hc(nsm);
    const int smem3 = 160 * 1024;
    CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    const int iters = 4000;
    int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    printf("T3 MMA issue rate (M=128, K=16, fp16, SS, %d iters, all %d SMs busy). floor = N/2 cyc\n", iters, nsm);
    printf("   %4s %26s %10s %10s %8s\n", "N", "variant", "cyc/mma", "floor", "frac");
    for (int N : Ns) {
      struct V { const char* name; int a_off, sbo, lbo, rot_d, rot_a; } vs[] = {
          {"aligned sbo128 same-D", 0, 128, 8192, 0, 0},
          {"aligned sbo128 rot-D", 0, 128, 8192, 1, 0},
          {"taps sbo160 rot-D", 16, 160, 17280, 1, 1},
          {"taps sbo128 rot-D", 16, 128, 17280, 1, 1},
      };
      for (auto& v : vs) {
        for (int rep = 0; rep < 2; ++rep) {
          mma_rate<<<nsm, 128, smem3>>>(N, iters, v.a_off, v.sbo, v.lbo, v.rot_d, v.rot_a, dCyc);
          CK(cudaGetLastError());
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(hc.data(), dCyc, nsm * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto c : hc) avg += c;
        avg /= nsm;
        double per = avg / iters;
        printf("   %4d %26s %10.2f %10.1f %8.3f\n", N, v.name, per, N / 2.0, (N / 2.0) / per);
      }
    }
    CK(cudaFree(dCyc));
  }
  // ---------------- T4
  {
    const int nsm = prop.multiProcessorCount;
    long long* dCyc;
    float* dSink;
    CK(cudaMalloc(&dCyc, nsm * 8));
    CK(cudaMalloc(&dSink, 1024));
    std::vector<long long> hc(nsm);
    const int reps = 200;
    for (int rep = 0; rep < 2; ++rep) {
      tmem_read_rate<<<nsm, 128>>>(reps, dCyc, dSink);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(hc.data(), dCyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto c : hc) avg += c;
    avg /= nsm;
    printf("T4 TMEM read: %.1f cycles per 128x512 fp32 tile (4 warps) = %.1f B/cyc/SM\n", avg / reps,
           128.0 * 512 * 4 / (avg / reps));
  }
  printf("probe done\n");
  return 0;
}
