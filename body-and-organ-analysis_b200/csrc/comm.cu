// Device side of the multi-GPU exchange (dist.py).
//   * Peer-memory path (default on one box): every rank accumulates its patches into a PRIVATE buffer that lives in
//     memory allocated here and exported with CUDA IPC; the owner of a dim-0 slab maps its peers' buffers and reads
//     their parts of its slab directly over NVLink inside the fused reduce + finalize kernel
//     (boa_reduce_finalize_peers, passes.cu) - no send / recv, no staging copies, no summed-logits tensor.
//   * NCCL path (gloo tests, no peer access): grouped send / recv of the pieces, then the owner adds what it received
//     to its own contribution in fixed rank order (boa_add_slab*).
// The reference has no multi-GPU inference path; nothing here replaces reference code.
#include <string.h>
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

// dst[c][i] += src[c][i] for c < C, i < n, with independent channel strides (elements): one launch adds a whole
// [C, len, Y, X] piece into a slab of a different length.
namespace boa {
__global__ void __launch_bounds__(256) add_slab_strided_kernel(float* __restrict__ dst, size_t dst_cstride,
                                                               const float* __restrict__ src, size_t src_cstride, int C,
                                                               size_t n, int vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int c = 0; c < C; ++c) {
    float* d = dst + (size_t)c * dst_cstride;
    const float* s = src + (size_t)c * src_cstride;
    if (vec) {
      const size_t nvec = n / 4;
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 a = reinterpret_cast<float4*>(d)[i];
        const float4 b = __ldg(reinterpret_cast<const float4*>(s) + i);
        a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        reinterpret_cast<float4*>(d)[i] = a;
      }
      for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        d[i] = __fadd_rn(d[i], s[i]);
    } else {
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = __fadd_rn(d[i], s[i]);
    }
  }
}
}  // namespace boa

extern "C" int boa_add_slab_strided(float* d_dst, size_t dst_cstride, const float* d_src, size_t src_cstride, int C,
                                    size_t n, void* stream) {
  BOA_REQUIRE(d_dst && d_src && C > 0, "boa_add_slab_strided: bad argument");
  if (n == 0) return BOA_OK;
  const bool vec = ((reinterpret_cast<uintptr_t>(d_dst) | reinterpret_cast<uintptr_t>(d_src)) & 15) == 0 &&
                   dst_cstride % 4 == 0 && src_cstride % 4 == 0;
  boa::add_slab_strided_kernel<<<boa::grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_dst, dst_cstride, d_src, src_cstride, C, n, vec ? 1 : 0);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// ---- peer memory: buffers allocated here can be mapped by the other ranks of the box (CUDA IPC)
extern "C" int boa_comm_alloc(size_t bytes, void** d_ptr, unsigned char* ipc_handle_out) {
  BOA_REQUIRE(d_ptr && ipc_handle_out && bytes > 0, "boa_comm_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == BOA_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  BOA_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    boa::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return BOA_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *d_ptr = p;
  return BOA_OK;
}

extern "C" int boa_comm_open(const unsigned char* ipc_handle, void** d_ptr) {
  BOA_REQUIRE(ipc_handle && d_ptr, "boa_comm_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  BOA_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return BOA_OK;
}

extern "C" int boa_comm_close(void* d_ptr) {
  if (d_ptr) BOA_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return BOA_OK;
}

extern "C" int boa_comm_free(void* d_ptr) {
  if (d_ptr) BOA_CUDA(cudaFree(d_ptr));
  return BOA_OK;
}

extern "C" int boa_comm_zero(void* d_ptr, size_t bytes, void* stream) {
  BOA_REQUIRE(d_ptr, "boa_comm_zero: null");
  BOA_CUDA(cudaMemsetAsync(d_ptr, 0, bytes, static_cast<cudaStream_t>(stream)));
  return BOA_OK;
}

extern "C" int boa_add_slab(float* d_dst, const float* d_src, size_t n, void* stream) {
  BOA_REQUIRE(d_dst && d_src, "boa_add_slab: null pointer");
  if (n == 0) return BOA_OK;
  if (((reinterpret_cast<uintptr_t>(d_dst) | reinterpret_cast<uintptr_t>(d_src)) & 15) != 0)  // unaligned slab offsets
    return boa_add_slab_strided(d_dst, n, d_src, n, 1, n, stream);
  boa::add_slab_kernel<<<boa::grid_for(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dst, d_src, n);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}
