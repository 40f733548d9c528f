// Device side of the multi-GPU exchange: after the NCCL send/recv of the overlapping logits slabs (one exchange per
// model) the owner adds what it received to its own contribution in fixed rank order.
#include "common.cuh"

namespace boa {
__global__ void __launch_bounds__(256) add_slab_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  const size_t nvec = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + i);
    a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __fadd_rn(dst[i], src[i]);
}
}  // namespace boa

extern "C" int boa_add_slab(float* d_dst, const float* d_src, size_t n, void* stream) {
  BOA_REQUIRE(d_dst && d_src, "boa_add_slab: null pointer");
  BOA_REQUIRE(((reinterpret_cast<uintptr_t>(d_dst) | reinterpret_cast<uintptr_t>(d_src)) & 15) == 0,
              "boa_add_slab: pointers must be 16-byte aligned");
  if (n == 0) return BOA_OK;
  boa::add_slab_kernel<<<boa::grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dst, d_src, n);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}
