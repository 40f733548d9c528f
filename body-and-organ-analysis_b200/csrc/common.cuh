// Shared host/device helpers for libboa_b200.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/boa_b200.h"

namespace boa {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define BOA_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      boa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return BOA_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define BOA_CHECK_LAUNCH()                                                                      \
  do {                                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                        \
    if (e_ != cudaSuccess) {                                                                    \
      boa::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return BOA_ERR_CUDA;                                                                      \
    }                                                                                           \
    boa::count_launch();                                                                        \
  } while (0)

#define BOA_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      boa::set_error(__VA_ARGS__);    \
      return BOA_ERR_ARG;             \
    }                                 \
  } while (0)

constexpr int MAX_DYN_SMEM = 232448;  // 227 KB: the per-CTA opt-in maximum on sm_100

// Kernels can only share an SM with the persistent tcgen05 conv CTAs (227 KB of dynamic shared memory) when they ask
// for the same L1 / shared-memory split: the HBM-bound passes of one lane overlap the convolutions of the other only
// with the carve-out preference set to "all shared".
template <typename K>
inline void prefer_max_smem_carveout(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
#define BOA_CARVEOUT_ONCE(kernel)            \
  do {                                       \
    static bool done_ = false;               \
    if (!done_) {                            \
      boa::prefer_max_smem_carveout(kernel); \
      done_ = true;                          \
    }                                        \
  } while (0)

// Co-resident ("thin") launch shapes for the HBM-bound passes of a forward (opt-in with BOA_B200_THIN=1): see
// norm_lrelu_thin_kernel in net_simt.cu.  Measured (profiles/r01_overlap.txt): the passes do co-run with the conv
// CTAs, but both sides then share HBM / L2 and the sum barely moves, while the thin shapes are slower on their own.
bool thin_passes();

// Number of SMs of the current device (cached) - grids are sized in multiples of it.
int sm_count();

inline int grid_for(size_t work_items, int threads, int blocks_per_sm = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace boa
