// Host-side TMA descriptor construction without linking libcuda: the driver entry point is resolved at run time
// through the CUDA runtime, so the shared library still loads on a CPU-only box.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace boa {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Activation tensors are "C8": [n_groups][D][H][W][8] fp16, n_groups = batch * C/8 (channel groups of 8 are the
// outermost dimension, the 8 channels of a group are the innermost 16 bytes). Out-of-bounds coordinates are
// zero-filled, which is the conv's zero padding.
// merged (default): the TMA view is 4-D (8*W, H, D, n_groups) with boxes (8*bx, by, bz, bg) - channel and x are
// contiguous in memory, so one box row is bx*16 contiguous bytes.  The shared-memory image is byte-identical to the
// 5-D view (8, W, H, D, n_groups) / box (8, bx, by, bz, bg), whose 16-byte innermost rows make the TMA unit issue one
// request per voxel (measured: the stride-2 convs were bound by exactly that rate).
inline bool c8_tmap_merged() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOA_B200_TMAP5D");
    v = (e && atoi(e) != 0) ? 0 : 1;
  }
  return v == 1;
}

inline int make_c8_tmap(CUtensorMap* out, const void* base, int n_groups, int D, int H, int W, int bx, int by,
                        int bz, int bg) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -1;
  CUresult r;
  if (c8_tmap_merged()) {
    cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)n_groups};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)W * H * 16, (cuuint64_t)W * H * D * 16};
    cuuint32_t box[4] = {(cuuint32_t)bx * 8, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bg};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)n_groups};
    cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)W * H * 16, (cuuint64_t)W * H * D * 16};
    cuuint32_t box[5] = {8, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bg};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d (groups %d D %d H %d W %d box %d %d %d %d)\n", (int)r,
            n_groups, D, H, W, bx, by, bz, bg);
    return -2;
  }
  return 0;
}

}  // namespace boa
