// Shared epilogue of the tcgen05 conv kernels: TMEM -> registers (+bias) -> fp16 C8 store, with the InstanceNorm
// sum / sum^2 of the fp32 accumulators.  Software pipelined: the TMEM load of z-slot s+1 is in flight while slot s is
// converted and stored.  Statistics are kept as per-lane running fp64 sums across the tiles a persistent CTA handles
// and flushed with one atomic per (batch item, channel) run instead of one per tile (same-address fp64 atomics from
// every tile of every CTA serialise in L2).
#pragma once
#include "ptx.cuh"

namespace boa {

struct RunningStats {  // lane l of the warp holds the sums of channel key*32 + l
  double s1 = 0.0, s2 = 0.0;
  int key = -1;        // (b * Cout + first channel of the strip) of the run being accumulated
};

__device__ __forceinline__ void stats_flush(RunningStats& r, double* __restrict__ stats, int lane) {
  if (r.key >= 0 && stats) {
    double* st = stats + ((size_t)r.key + lane) * 2;
    atomicAdd(st, r.s1);
    atomicAdd(st + 1, r.s2);
  }
  r.s1 = 0.0; r.s2 = 0.0; r.key = -1;
}

// Optional second destination of a conv output: the space-to-depth copy [B][phases][G][D/sz][H/sy][W/sx][8] that the
// next stage's strided conv reads (strides 1 or 2 per axis; phase = ((z%sz)*sy + y%sy)*sx + x%sx).  `base` is this
// thread's voxel (x, y) of the strip's first channel group at z = 0; nullptr = no copy.
struct S2dDst {
  uint4* base = nullptr;
  size_t zpar_stride = 0;   // z odd (sz == 2): + sy * sx phases
  size_t zhalf_stride = 0;  // per z / sz
  size_t gstride = 0;       // per channel group
  int zshift = 0;           // sz - 1
  __device__ __forceinline__ uint4* at(int z) const {
    return base + (size_t)(z & zshift) * zpar_stride + (size_t)(z >> zshift) * zhalf_stride;
  }
};

__device__ __forceinline__ S2dDst s2d_dst(__half* s2d, const int (&st)[3], int b, int groups, int g0, int D, int H, int W,
                                          int y, int x) {
  S2dDst d;
  if (!s2d) return d;
  const int sz = st[0], sy = st[1], sx = st[2];
  const int Hh = H / sy, Wh = W / sx;
  const size_t hv = (size_t)(D / sz) * Hh * Wh;
  const int ph_xy = (y % sy) * sx + (x % sx);
  d.base = reinterpret_cast<uint4*>(s2d) + ((size_t)(b * sz * sy * sx + ph_xy) * groups + g0) * hv +
           (size_t)(y / sy) * Wh + (x / sx);
  d.zpar_stride = (size_t)sy * sx * groups * hv;
  d.zhalf_stride = (size_t)Hh * Wh;
  d.gstride = hv;
  d.zshift = sz - 1;
  return d;
}

__device__ __forceinline__ void epi_store_slot(const uint32_t (&v)[32], const float (&bs)[32], float (&s1)[32],
                                               float (&s2)[32], uint4* __restrict__ dst, size_t gstride,
                                               uint4* __restrict__ dst2 = nullptr, size_t gstride2 = 0) {
  float f[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    f[c] = __uint_as_float(v[c]) + bs[c];
    s1[c] += f[c];
    s2[c] = fmaf(f[c], f[c], s2[c]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
    dst[j * gstride] = o;
    if (dst2) dst2[j * gstride2] = o;
  }
}

// One 32-channel strip of one tile for one warp.  taddr: TMEM address (lane quadrant | first column of slot 0 of this
// strip); slot_cols: column distance between consecutive z-slots; dst: this thread's voxel of plane z0 in channel
// group (first channel of the strip)/8; zstride = H*W, gstride = D*H*W (uint4 units).
__device__ __forceinline__ void conv_epilogue_strip(uint32_t taddr, int slot_cols, int zt, const float* __restrict__ bias32,
                                                    bool rowvalid, int z0, int D, uint4* __restrict__ dst,
                                                    size_t zstride, size_t gstride, int lane, int key,
                                                    RunningStats& run, double* __restrict__ stats,
                                                    const S2dDst s2d = S2dDst()) {
  float bs[32], s1[32], s2[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { bs[c] = __ldg(bias32 + c); s1[c] = 0.f; s2[c] = 0.f; }
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll 1
  for (int slot = 0; slot < zt; slot += 2) {
    tmem_ld_wait();
    if (slot + 1 < zt) tmem_ld32(taddr + (uint32_t)((slot + 1) * slot_cols), vb);
    if (rowvalid && z0 + slot < D)
      epi_store_slot(va, bs, s1, s2, dst + (size_t)slot * zstride, gstride, s2d.base ? s2d.at(z0 + slot) : nullptr,
                     s2d.gstride);
    if (slot + 1 < zt) {
      tmem_ld_wait();
      if (slot + 2 < zt) tmem_ld32(taddr + (uint32_t)((slot + 2) * slot_cols), va);
      if (rowvalid && z0 + slot + 1 < D)
        epi_store_slot(vb, bs, s1, s2, dst + (size_t)(slot + 1) * zstride, gstride,
                       s2d.base ? s2d.at(z0 + slot + 1) : nullptr, s2d.gstride);
    }
  }
  // transpose-reduce over the 32 lanes: afterwards lane l holds the warp total of channel l of the strip
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send1 = upper ? s1[k] : s1[k + off];
      const float send2 = upper ? s2[k] : s2[k + off];
      const float r1 = __shfl_xor_sync(0xffffffffu, send1, off);
      const float r2 = __shfl_xor_sync(0xffffffffu, send2, off);
      s1[k] = (upper ? s1[k + off] : s1[k]) + r1;
      s2[k] = (upper ? s2[k + off] : s2[k]) + r2;
    }
  }
  if (run.key != key) {
    stats_flush(run, stats, lane);
    run.key = key;
  }
  run.s1 += (double)s1[0];
  run.s2 += (double)s2[0];
}

// ---- transposed conv (k = s = 2): GEMM column n = (((pz*2+py) * Cout/8 + cg) * 2 + px) * 8 + e.  The two x-phases
// of a channel group are neighbours on N, so a thread owns 32 contiguous output bytes (xo = 2x, 2x+1) per group and
// writes them with ONE 256-bit store (full 32-byte sectors; two 16-byte stores would be partial-sector writes).
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ void tconv_store_slot(const uint32_t (&v)[32], const float (&bs)[32], uint4* __restrict__ d0,
                                                 uint4* __restrict__ d1) {
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    uint4 o[2];
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      __half2* h = reinterpret_cast<__half2*>(&o[px]);
      const int c0 = 16 * jj + 8 * px;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __floats2half2_rn(__uint_as_float(v[c0 + 2 * e]) + bs[c0 + 2 * e],
                                 __uint_as_float(v[c0 + 2 * e + 1]) + bs[c0 + 2 * e + 1]);
    }
    st_global_256(jj ? d1 : d0, o[0], o[1]);
  }
}

// One 32-column strip (two channel groups x two x-phases of one (pz,py) phase) of all z-slots of a tile.
// d0 / d1: this thread's output voxel pair (zo of slot 0, yo, 2x) in the two channel groups; zstride2 = 2*Ho*Wo.
__device__ __forceinline__ void tconv_epilogue_strip(uint32_t taddr, int slot_cols, int zt, const float (&bs)[32],
                                                     bool rowvalid, int z0, int D, uint4* __restrict__ d0,
                                                     uint4* __restrict__ d1, size_t zstride2) {
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll 1
  for (int slot = 0; slot < zt; slot += 2) {
    tmem_ld_wait();
    if (slot + 1 < zt) tmem_ld32(taddr + (uint32_t)((slot + 1) * slot_cols), vb);
    if (rowvalid && z0 + slot < D) tconv_store_slot(va, bs, d0 + (size_t)slot * zstride2, d1 + (size_t)slot * zstride2);
    if (slot + 1 < zt) {
      tmem_ld_wait();
      if (slot + 2 < zt) tmem_ld32(taddr + (uint32_t)((slot + 2) * slot_cols), va);
      if (rowvalid && z0 + slot + 1 < D)
        tconv_store_slot(vb, bs, d0 + (size_t)(slot + 1) * zstride2, d1 + (size_t)(slot + 1) * zstride2);
    }
  }
}

}  // namespace boa
