// Fused input normalisation of the tcgen05 conv kernels.
//
// A conv's input is the RAW fp16 output of its producer(s); InstanceNorm3d (per (patch, channel) scale / shift, final
// once the producer kernel has finished) + LeakyReLU are applied to the operand tile in shared memory, between the TMA
// write and the MMAs, by a dedicated warpgroup - so the standalone normalise pass (read + write of every activation
// tensor: 18 % of the GPU time and 10 of 26 GB of DRAM traffic per batch in round 1) does not exist.
//   reference op: dynamic_network_architectures ConvDropoutNormReLU = Conv3d -> InstanceNorm3d(affine) -> LeakyReLU,
//   kwargs from _external/nnunetv2/utilities/plans_handling/plans_handler.py:72-82.
// Same fp32 operations as norm_lrelu_kernel (fma, LeakyReLU, round to fp16), so the fused and the unfused schedules
// are bit-identical (tests/test_gpu_network.py).  Positions outside the volume were zero-filled by TMA and
// must stay zero (the conv pads the ACTIVATED tensor), so only the in-volume part of the halo box is touched.
// Channel groups that are final already (the transposed-conv half of a decoder concat) are skipped.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace boa {

struct InXform {
  const float* scale = nullptr;  // [B][channels]; nullptr: the input needs no transform
  const float* shift = nullptr;
  int channels = 0;              // row length of scale / shift = channels of the input view
  int ident_groups = 0;          // leading channel groups of the input view that are final (left untouched)
  float slope = 0.01f;
};

// y = lrelu(fma(x, a, s)) for the 8 channels of one voxel.  ONE definition for every kernel that normalises (the
// tensor-core kernels' transform warps, the SIMT kernels, the head, the standalone pass of the unfused schedule), so
// that all schedules are bit-identical.  slope < 1: lrelu(z) = max(z, z * slope).
__device__ __forceinline__ uint4 xform8(const uint4& raw, const float (&a)[8], const float (&sh)[8], float slope) {
  uint4 o;
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
  __half2* r = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __half22float2(h[e]);
    f.x = fmaf(f.x, a[2 * e], sh[2 * e]);
    f.y = fmaf(f.y, a[2 * e + 1], sh[2 * e + 1]);
    f.x = fmaxf(f.x, __fmul_rn(f.x, slope));
    f.y = fmaxf(f.y, __fmul_rn(f.y, slope));
    r[e] = __floats2half2_rn(f.x, f.y);
  }
  return o;
}

// One operand stage: 2 channel groups x [bz][BY][BX] positions x 16 bytes, in place.  Called by nt threads (tid).
// [zlo,zhi) x [ylo,yhi) x [xlo,xhi): in-volume part of the box.  sc / sh: scale / shift of the 16 channels of this
// K chunk for this batch item.  skip: bit g set = group g is final (or does not exist).
// Work items are (z phase of zs = 2 or 4, in-plane position; at most MAXP planes per item): a thread's (y, x) is fixed per item, so the bounds test and the
// index arithmetic happen once per item and the planes of an item are loaded together before they are transformed.
template <int BX, int BY, int MAXP>
__device__ __forceinline__ void xform_stage(uint8_t* sa, int bz, int zlo, int zhi, int ylo, int yhi, int xlo, int xhi,
                                            const float* __restrict__ sc, const float* __restrict__ sh, int skip,
                                            float slope, int tid, int nt, int zs) {
  constexpr int SL = BX * BY;
  const int per_group = bz * SL;
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    if ((skip >> g) & 1) continue;
    float a[8], s[8];
    {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(sc + 8 * g)), a1 = __ldg(reinterpret_cast<const float4*>(sc + 8 * g) + 1);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(sh + 8 * g)), s1 = __ldg(reinterpret_cast<const float4*>(sh + 8 * g) + 1);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    }
    uint4* t = reinterpret_cast<uint4*>(sa) + g * per_group;
#pragma unroll 1
    for (int it = tid; it < zs * SL; it += nt) {
      const int zp = it / SL, r = it - zp * SL, y = r / BX, x = r - y * BX;
      if (y < ylo || y >= yhi || x < xlo || x >= xhi) continue;
      // planes zlo <= z < zhi with z % zs == zp (zs is a power of two)
      const int z = zlo + ((zp - zlo) & (zs - 1));
      uint4 v[MAXP];
#pragma unroll
      for (int k = 0; k < MAXP; ++k)
        if (z + k * zs < zhi) v[k] = t[(z + k * zs) * SL + r];
#pragma unroll
      for (int k = 0; k < MAXP; ++k)
        if (z + k * zs < zhi) t[(z + k * zs) * SL + r] = xform8(v[k], a, s, slope);
    }
  }
}

}  // namespace boa
