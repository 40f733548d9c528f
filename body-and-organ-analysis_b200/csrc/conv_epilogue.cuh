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

__device__ __forceinline__ void epi_store_slot(const uint32_t (&v)[32], const float (&bs)[32], float (&s1)[32],
                                               float (&s2)[32], uint4* __restrict__ dst, size_t gstride) {
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
  }
}

// One 32-channel strip of one tile for one warp.  taddr: TMEM address (lane quadrant | first column of slot 0 of this
// strip); slot_cols: column distance between consecutive z-slots; dst: this thread's voxel of plane z0 in channel
// group (first channel of the strip)/8; zstride = H*W, gstride = D*H*W (uint4 units).
__device__ __forceinline__ void conv_epilogue_strip(uint32_t taddr, int slot_cols, int zt, const float* __restrict__ bias32,
                                                    bool rowvalid, int z0, int D, uint4* __restrict__ dst,
                                                    size_t zstride, size_t gstride, int lane, int key,
                                                    RunningStats& run, double* __restrict__ stats) {
  float bs[32], s1[32], s2[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { bs[c] = __ldg(bias32 + c); s1[c] = 0.f; s2[c] = 0.f; }
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll 1
  for (int slot = 0; slot < zt; slot += 2) {
    tmem_ld_wait();
    if (slot + 1 < zt) tmem_ld32(taddr + (uint32_t)((slot + 1) * slot_cols), vb);
    if (rowvalid && z0 + slot < D) epi_store_slot(va, bs, s1, s2, dst + (size_t)slot * zstride, gstride);
    if (slot + 1 < zt) {
      tmem_ld_wait();
      if (slot + 2 < zt) tmem_ld32(taddr + (uint32_t)((slot + 2) * slot_cols), va);
      if (rowvalid && z0 + slot + 1 < D) epi_store_slot(vb, bs, s1, s2, dst + (size_t)(slot + 1) * zstride, gstride);
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

// ---- 16-column strips: the same epilogue at half the register footprint, for kernels that share the register file
// with more warps (conv3_fold_ldnorm_kernel: 13 warps).  Slower per tile (twice the TMEM round trips), which only
// matters where the epilogue is exposed - not with double-buffered accumulators.  Two neighbouring strips share one
// RunningStats16: after a strip's reduction lanes 2c and 2c+1 both hold the total of its channel c; even lanes keep the
// even strip's totals, odd lanes the odd strip's.
struct RunningStats16 {  // lane l holds the sums of channel key + 16 * (l & 1) + (l >> 1)
  double s1 = 0.0, s2 = 0.0;
  int key = -1;          // (b * Cout + first channel of the EVEN strip) of the run being accumulated
};

__device__ __forceinline__ void stats_flush16(RunningStats16& r, double* __restrict__ stats, int lane) {
  if (r.key >= 0 && stats) {
    double* st = stats + ((size_t)r.key + 16 * (lane & 1) + (lane >> 1)) * 2;
    atomicAdd(st, r.s1);
    atomicAdd(st + 1, r.s2);
  }
  r.s1 = 0.0; r.s2 = 0.0; r.key = -1;
}

__device__ __forceinline__ void epi_store_slot16(const uint32_t (&v)[16], const float (&bs)[16], float (&s1)[16],
                                                 float (&s2)[16], uint4* __restrict__ dst, size_t gstride) {
  float f[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    f[c] = __uint_as_float(v[c]) + bs[c];
    s1[c] += f[c];
    s2[c] = fmaf(f[c], f[c], s2[c]);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
    dst[j * gstride] = o;
  }
}

// One 16-channel strip of one tile for one warp; `key` is the pair's key (first channel of the even strip), `odd` says
// which strip of the pair this is.  Other arguments as conv_epilogue_strip.
__device__ __forceinline__ void conv_epilogue_strip16(uint32_t taddr, int slot_cols, int zt, const float* __restrict__ bias16,
                                                      bool rowvalid, int z0, int D, uint4* __restrict__ dst,
                                                      size_t zstride, size_t gstride, int lane, int key, int odd,
                                                      RunningStats16& run, double* __restrict__ stats) {
  float bs[16], s1[16], s2[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { bs[c] = __ldg(bias16 + c); s1[c] = 0.f; s2[c] = 0.f; }
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
#pragma unroll 1
  for (int slot = 0; slot < zt; slot += 2) {
    tmem_ld_wait();
    if (slot + 1 < zt) tmem_ld16(taddr + (uint32_t)((slot + 1) * slot_cols), vb);
    if (rowvalid && z0 + slot < D) epi_store_slot16(va, bs, s1, s2, dst + (size_t)slot * zstride, gstride);
    if (slot + 1 < zt) {
      tmem_ld_wait();
      if (slot + 2 < zt) tmem_ld16(taddr + (uint32_t)((slot + 2) * slot_cols), va);
      if (rowvalid && z0 + slot + 1 < D) epi_store_slot16(vb, bs, s1, s2, dst + (size_t)(slot + 1) * zstride, gstride);
    }
  }
  // transpose-reduce over the 32 lanes: four halving exchanges (16 -> 1 values per lane) and a final pair add;
  // afterwards lanes 2c and 2c+1 both hold the warp total of channel c of the strip
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step;   // lane distance of the exchange
    const int n = 8 >> step;      // values kept
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const float send1 = upper ? s1[k] : s1[k + n];
      const float send2 = upper ? s2[k] : s2[k + n];
      const float r1 = __shfl_xor_sync(0xffffffffu, send1, off);
      const float r2 = __shfl_xor_sync(0xffffffffu, send2, off);
      s1[k] = (upper ? s1[k + n] : s1[k]) + r1;
      s2[k] = (upper ? s2[k + n] : s2[k]) + r2;
    }
  }
  s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
  s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
  if (run.key != key) {
    stats_flush16(run, stats, lane);
    run.key = key;
  }
  if ((lane & 1) == odd) {
    run.s1 += (double)s1[0];
    run.s2 += (double)s2[0];
  }
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
