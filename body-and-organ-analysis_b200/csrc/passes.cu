// HBM-bound passes of the BOA hot path: CT normalisation, Gaussian patch accumulation, normalise+argmax+LUT merge,
// HU tissue rules, per-slice / per-label integer reductions.  All are single-pass, 128-bit vectorised, grid sized
// in multiples of the SM count; floating-point ops that must match numpy bit-for-bit use explicit *_rn intrinsics
// (no FMA contraction).
#include "common.cuh"

namespace boa {

// ------------------------------------------------------------------------------------------------ CT normalise
// default_normalization_schemes.py:56-67 : clip, -= mean, /= max(std, 1e-8)  (fp32 throughout)
__device__ __forceinline__ float ct_norm1(float x, float lo, float hi, float mean, float std) {
  x = fminf(fmaxf(x, lo), hi);
  return __fdiv_rn(__fsub_rn(x, mean), std);
}

template <typename T>
__global__ void __launch_bounds__(256) ct_normalize_kernel(const T* __restrict__ in, size_t n, float lo, float hi,
                                                           float mean, float std, float* __restrict__ out) {
  const size_t nvec = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float v[8];
    if constexpr (sizeof(T) == 2) {
      uint4 raw = __ldg(reinterpret_cast<const uint4*>(in) + i);
      const short* s = reinterpret_cast<const short*>(&raw);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (float)s[k];
    } else {
      float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
      float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ct_norm1(v[k], lo, hi, mean, std);
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  // tail
  for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = ct_norm1((float)in[i], lo, hi, mean, std);
}

// ------------------------------------------------------------------------------------------------ patch accumulate
// predict_from_raw_data.py:609-613 : prediction *= gaussian ; predicted_logits[sl] += prediction
__global__ void __launch_bounds__(256)
accumulate_patch_kernel(const float* __restrict__ logits, int C, int p0, int p1, int p2, int o0, int o1, int o2,
                        const float* __restrict__ g, float* __restrict__ acc, int d1, int d2, size_t vol_voxels) {
  const size_t pv = (size_t)p0 * p1 * p2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pv; i += stride) {
    const int k = (int)(i % p2);
    const int j = (int)((i / p2) % p1);
    const int ii = (int)(i / ((size_t)p2 * p1));
    const size_t v = ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k);
    const float gw = __ldg(g + i);
    for (int c = 0; c < C; ++c) {
      const float pr = __fmul_rn(__ldg(logits + (size_t)c * pv + i), gw);
      float* a = acc + (size_t)c * vol_voxels + v;
      *a = __fadd_rn(*a, pr);
    }
  }
}

// predict_from_raw_data.py:614 : n_predictions[sl[1:]] += gaussian   (one launch per patch keeps the reference order)
__global__ void __launch_bounds__(256)
accumulate_weight_kernel(int p0, int p1, int p2, int o0, int o1, int o2, const float* __restrict__ g,
                         float* __restrict__ wacc, int d1, int d2) {
  const size_t pv = (size_t)p0 * p1 * p2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pv; i += stride) {
    const int k = (int)(i % p2);
    const int j = (int)((i / p2) % p1);
    const int ii = (int)(i / ((size_t)p2 * p1));
    const size_t v = ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k);
    wacc[v] = __fadd_rn(wacc[v], __ldg(g + i));
  }
}

// ------------------------------------------------------------------------------------------------ finalize + argmax
struct Lut256 {
  uint8_t v[256];
};

// Where the accumulated logits of a voxel come from.
// LocalSrc: one accumulator [C][V] in this GPU's memory.
struct LocalSrc {
  const float* acc;
  size_t V;
  template <int VEC>
  __device__ __forceinline__ void load(int c, size_t v0, float (&x)[VEC]) const {
    if constexpr (VEC == 4) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(acc + (size_t)c * V + v0));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
      x[0] = __ldcs(acc + (size_t)c * V + v0);
    }
  }
};
// PeerSrc (multi-GPU, dist.py): rank r holds the partial sums of ITS patches in a private buffer
// [C][zhi[r] - zlo[r]][Y][X] covering the slices its patches touch; the owner of a dim-0 slab reads every rank's part
// of its slab straight from that rank's memory (NVLink peer loads; base[r] is the IPC-mapped buffer) and adds them in
// RANK ORDER - the exchange, the reduction and the argmax are one kernel, the summed logits are never stored.
// v0 is relative to the owner's slab, which starts at absolute slice z0.  VEC == 4 needs plane % 4 == 0.
constexpr int MAX_PEERS = 8;
struct PeerSrc {
  const float* base[MAX_PEERS];
  int zlo[MAX_PEERS], zhi[MAX_PEERS];
  int n;
  int z0;
  size_t plane;  // Y * X
  template <int VEC>
  __device__ __forceinline__ void load(int c, size_t v0, float (&x)[VEC]) const {
    const int z = z0 + (int)(v0 / plane);
    const size_t rem = v0 - (size_t)(z - z0) * plane;
#pragma unroll
    for (int k = 0; k < VEC; ++k) x[k] = 0.f;
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r) {
      if (r < n && z >= zlo[r] && z < zhi[r]) {
        const float* p = base[r] + ((size_t)c * (zhi[r] - zlo[r]) + (z - zlo[r])) * plane + rem;
        if constexpr (VEC == 4) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
          x[0] = __fadd_rn(x[0], t.x); x[1] = __fadd_rn(x[1], t.y); x[2] = __fadd_rn(x[2], t.z); x[3] = __fadd_rn(x[3], t.w);
        } else {
          x[0] = __fadd_rn(x[0], __ldcs(p));
        }
      }
    }
  }
};

// ResampledSrc: the network ran on a grid that nnU-Net's preprocessing resampled the volume to; the reference divides
// the logits by the weight sum ON THAT GRID and resamples every channel back to the pre-resampling shape with order 1
// before the argmax (_external/nnunetv2/inference/export_prediction.py:25-38 ->
// resample_data_or_seg_to_shape(..., is_seg=False, order=1, order_z=0), default_resampling.py:117-203:
// skimage.transform.resize(order=1, mode="edge") = scipy zoom with grid_mode coordinates (o + 0.5) * n_in / n_out - 0.5,
// clamped; slice by slice with an order-0 pick along the anisotropic axis when separate_z).  Here the interpolation is
// part of the argmax pass: value(c, voxel) = sum_k weight_k * (acc[c][p_k] / w[p_k]) over the 4 / 8 grid neighbours.
// v0 indexes the OUTPUT grid; VEC == 1 only.
struct ResampledSrc {
  const float* acc;
  const float* w;
  int Zn, Yn, Xn, Zo, Yo, Xo;
  int separate_z;
  __device__ __forceinline__ static void axis(int o, int n_in, int n_out, int& i0, int& i1, double& t) {
    double cc = ((double)o + 0.5) * ((double)n_in / (double)n_out) - 0.5;
    cc = cc < 0.0 ? 0.0 : (cc > (double)(n_in - 1) ? (double)(n_in - 1) : cc);
    const double fl = floor(cc);
    i0 = (int)fl;
    i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
    t = cc - fl;
  }
  template <int VEC>
  __device__ __forceinline__ void load(int c, size_t v0, float (&x)[VEC]) const {
    static_assert(VEC == 1, "ResampledSrc is scalar");
    const int xo = (int)(v0 % Xo), yo = (int)((v0 / Xo) % Yo), zo = (int)(v0 / ((size_t)Xo * Yo));
    int x0, x1, y0, y1, z0, z1;
    double tx, ty, tz;
    axis(xo, Xn, Xo, x0, x1, tx);
    axis(yo, Yn, Yo, y0, y1, ty);
    if (separate_z) {  // map_coordinates(order=0): floor(scale * (k + 0.5) - 0.5 + 0.5)
      int k = (int)floor(((double)Zn / (double)Zo) * ((double)zo + 0.5));
      z0 = z1 = k < 0 ? 0 : (k >= Zn ? Zn - 1 : k);
      tz = 0.0;
    } else {
      axis(zo, Zn, Zo, z0, z1, tz);
    }
    const size_t Vn = (size_t)Zn * Yn * Xn;
    const float* a = acc + (size_t)c * Vn;
    auto at = [&](int z, int y, int xx) -> double {
      const size_t i = ((size_t)z * Yn + y) * Xn + xx;
      return (double)__fdiv_rn(__ldg(a + i), __ldg(w + i));
    };
    const double v00 = at(z0, y0, x0) * (1.0 - tx) + at(z0, y0, x1) * tx;
    const double v01 = at(z0, y1, x0) * (1.0 - tx) + at(z0, y1, x1) * tx;
    double v = v00 * (1.0 - ty) + v01 * ty;
    if (tz != 0.0) {
      const double v10 = at(z1, y0, x0) * (1.0 - tx) + at(z1, y0, x1) * tx;
      const double v11 = at(z1, y1, x0) * (1.0 - tx) + at(z1, y1, x1) * tx;
      v = v * (1.0 - tz) + (v10 * (1.0 - ty) + v11 * ty) * tz;
    }
    x[0] = (float)v;
  }
};

// predict_from_raw_data.py:620-625 (divide, isinf) ; label_handling.py:178 (argmax(0), first max wins) ;
// totalsegmentator/nnunet.py:553-556 (part label -> global label, non-zero overwrite)
template <int VEC, typename Src>
__device__ __forceinline__ void argmax_group(const Src& src, const float* __restrict__ wacc, int C, size_t v0,
                                             const Lut256& lut, int overwrite_nz, uint8_t* __restrict__ label,
                                             int& bad) {
  float w[VEC], best[VEC];
  int arg[VEC];
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(wacc + v0));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else {
    w[0] = wacc ? __ldg(wacc + v0) : 1.f;  // no weight sum: the source divides itself (ResampledSrc)
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) { best[k] = 0.f; arg[k] = 0; }
  for (int c = 0; c < C; ++c) {
    float x[VEC];
    src.template load<VEC>(c, v0, x);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float q = __fdiv_rn(x[k], w[k]);
      if (!isfinite(q)) ++bad;
      // numpy argmax: first maximum wins, NaN counts as maximum
      if (c == 0 || (q > best[k] && !(best[k] != best[k])) || (q != q && best[k] == best[k])) {
        best[k] = q;
        arg[k] = c;
      }
    }
  }
  uint8_t out[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) out[k] = lut.v[arg[k]];
  if (overwrite_nz) {
    if constexpr (VEC == 4) {
      uchar4 old = *reinterpret_cast<const uchar4*>(label + v0);
      const uint8_t o[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (out[k] == 0) out[k] = o[k];
    } else {
      if (out[0] == 0) out[0] = label[v0];
    }
  }
  if constexpr (VEC == 4) {
    *reinterpret_cast<uchar4*>(label + v0) = make_uchar4(out[0], out[1], out[2], out[3]);
  } else {
    label[v0] = out[0];
  }
}

// Fast path.  Dividing every channel by the weight sum costs ~20 instructions per (voxel, channel) and made the pass
// instruction bound at half of the HBM rate.  Division by one positive w is monotone, so the argmax of x / w is the
// argmax of x unless two channels are so close that their quotients may round to the same float (then numpy's "first
// maximum wins" looks at the quotients).  The loop therefore tracks max / second max / max |x| / NaN of the raw sums
// with the channel loads issued UNR at a time, and only a voxel group that is suspicious - near tie, tiny maximum
// (quotients may underflow to a tie), non-finite input, w == 0 or an overflowing quotient - is redone by the exact
// per-channel division of argmax_group<> (bit-identical result in every case, tests/test_gpu_passes.py).
template <int UNR, typename Src>
__global__ void __launch_bounds__(256)
finalize_argmax_kernel(const Src src, const float* __restrict__ wacc, int C, size_t V, Lut256 lut, int overwrite_nz,
                       uint8_t* __restrict__ label, int* __restrict__ nonfinite) {
  int bad = 0;
  const size_t nvec = V / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const size_t v0 = i * 4;
    const float4 w4 = __ldcs(reinterpret_cast<const float4*>(wacc + v0));
    const float w[4] = {w4.x, w4.y, w4.z, w4.w};
    float m1[4], m2[4], mabs[4];
    int arg[4];
    bool odd = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) { m1[k] = -INFINITY; m2[k] = -INFINITY; mabs[k] = 0.f; arg[k] = 0; }
    for (int c0 = 0; c0 < C; c0 += UNR) {
      float x4[UNR][4];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (c0 + u < C) src.template load<4>(c0 + u, v0, x4[u]);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (c0 + u < C) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x = x4[u][k];
            odd |= !(fabsf(x) <= 3.0e38f);  // NaN / inf (and the few finite values above: handled exactly below)
            mabs[k] = fmaxf(mabs[k], fabsf(x));
            m2[k] = fmaxf(m2[k], fminf(m1[k], x));
            if (x > m1[k]) arg[k] = c0 + u;
            m1[k] = fmaxf(m1[k], x);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float qa = mabs[k] / w[k];
      odd |= !(qa <= 3.0e38f);                                           // w == 0, NaN weight, overflowing quotient
      odd |= !(w[k] > 0.f);
      odd |= (C > 1) && !(m1[k] - m2[k] > 4.8e-7f * fabsf(m1[k]));       // near tie (2^-21 relative)
      odd |= fabsf(m1[k]) < 1e-30f;                                      // quotients may underflow into a tie
    }
    if (odd) {
      argmax_group<4>(src, wacc, C, v0, lut, overwrite_nz, label, bad);
      continue;
    }
    uint8_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = lut.v[arg[k]];
    if (overwrite_nz) {
      const uchar4 old = *reinterpret_cast<const uchar4*>(label + v0);
      const uint8_t o[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (out[k] == 0) out[k] = o[k];
    }
    *reinterpret_cast<uchar4*>(label + v0) = make_uchar4(out[0], out[1], out[2], out[3]);
  }
  for (size_t v = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += stride)
    argmax_group<1>(src, wacc, C, v, lut, overwrite_nz, label, bad);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nonfinite, bad);
}

// Scalar variant for channel planes that are not 16-byte aligned (V % 4 != 0: odd-sized volumes, or slabs of a sharded
// volume): same exact arithmetic, one voxel per thread (coalesced 4-byte accesses).
template <typename Src>
__global__ void __launch_bounds__(256)
finalize_argmax_scalar_kernel(const Src src, const float* __restrict__ wacc, int C, size_t V, Lut256 lut,
                              int overwrite_nz, uint8_t* __restrict__ label, int* __restrict__ nonfinite) {
  int bad = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += stride)
    argmax_group<1>(src, wacc, C, v, lut, overwrite_nz, label, bad);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nonfinite, bad);
}

// ------------------------------------------------------------------------------------------------ tissue rules
// tissue/definition.py:6-30 ; subclassification.py:38-53.  Bounds inclusive.
__device__ __forceinline__ uint8_t tissue_rule(float hu, uint8_t region) {
  const bool fat = hu >= -190.f && hu <= -30.f;
  switch (region) {
    case 2: return (hu >= -29.f && hu <= 150.f) ? 1 : (fat ? 5 : 0);  // MUSCLE / IMAT
    case 5: return (hu >= -1000.f && hu <= 3000.f) ? 2 : 0;           // BONE
    case 1: return fat ? 3 : 0;                                        // SAT
    case 3: return fat ? 4 : 0;                                        // VAT
    case 9: return fat ? 6 : 0;                                        // PAT
    case 7: return fat ? 7 : 0;                                        // EAT
    default: return 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
tissue_kernel(const T* __restrict__ ct, const uint8_t* __restrict__ regions, size_t n, uint8_t* __restrict__ out) {
  const size_t nvec = n / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(regions) + i);
    const uint8_t* r = reinterpret_cast<const uint8_t*>(&r4);
    float hu[16];
    if constexpr (sizeof(T) == 2) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(ct) + 2 * i);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(ct) + 2 * i + 1);
      const short* sa = reinterpret_cast<const short*>(&a);
      const short* sb = reinterpret_cast<const short*>(&b);
#pragma unroll
      for (int k = 0; k < 8; ++k) { hu[k] = (float)sa[k]; hu[8 + k] = (float)sb[k]; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(ct) + 4 * i + q);
        hu[4 * q] = f.x; hu[4 * q + 1] = f.y; hu[4 * q + 2] = f.z; hu[4 * q + 3] = f.w;
      }
    }
    uint4 o4;
    uint8_t* o = reinterpret_cast<uint8_t*>(&o4);
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = tissue_rule(hu[k], r[k]);
    reinterpret_cast<uint4*>(out)[i] = o4;
  }
  for (size_t v = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
    out[v] = tissue_rule((float)ct[v], regions[v]);
}

// ------------------------------------------------------------------------------------------------ per-slice stats
// builder.py:406-432 (per-slice per-tissue counts, with / without body_parts==TORSO), :284-305 (mean HU),
// :56-99 and commands.py:34-44 (slice presence == count > 0).
// One block handles a chunk of <= 65536 voxels of one slice: per-warp private histograms in shared memory
// (u32 counts, u32 sums of hu+32768 - cannot overflow within a chunk), then 64-bit global atomics.
constexpr int SLICE_CHUNK = 65536;
constexpr int SLICE_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(SLICE_THREADS)
slice_stats_kernel(const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask, int mask_value,
                   const T* __restrict__ ct, size_t slice_voxels, int n_labels, int chunks_per_slice,
                   unsigned long long* __restrict__ counts, long long* __restrict__ hu_sums) {
  extern __shared__ uint32_t sm[];
  const int nwarps = SLICE_THREADS / 32;
  uint32_t* cnt = sm;                       // [nwarps][n_labels]
  uint32_t* sum = sm + nwarps * n_labels;   // [nwarps][n_labels]
  for (int i = threadIdx.x; i < 2 * nwarps * n_labels; i += SLICE_THREADS) sm[i] = 0;
  __syncthreads();
  const int z = blockIdx.x / chunks_per_slice;
  const int chunk = blockIdx.x % chunks_per_slice;
  const size_t begin = (size_t)chunk * SLICE_CHUNK;
  const size_t end = min(begin + (size_t)SLICE_CHUNK, slice_voxels);
  const size_t base = (size_t)z * slice_voxels;
  const int warp = threadIdx.x >> 5;
  uint32_t* mycnt = cnt + warp * n_labels;
  uint32_t* mysum = sum + warp * n_labels;
  const bool want_sum = (hu_sums != nullptr);
  // 16-byte vector path needs (base+begin) 16-aligned for labels; fall back to scalar otherwise
  const bool aligned = ((base + begin) % 16 == 0) && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0) &&
                       (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0) &&
                       (!want_sum || (reinterpret_cast<uintptr_t>(ct) & 15) == 0);
  size_t v = begin;
  if (aligned) {
    const size_t nvec = (end - begin) / 16;
    for (size_t i = threadIdx.x; i < nvec; i += SLICE_THREADS) {
      const size_t g = base + begin + i * 16;
      const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(labels + g));
      const uint8_t* l = reinterpret_cast<const uint8_t*>(&l4);
      uint4 m4 = make_uint4(0, 0, 0, 0);
      if (mask) m4 = __ldg(reinterpret_cast<const uint4*>(mask + g));
      const uint8_t* m = reinterpret_cast<const uint8_t*>(&m4);
      int hu[16];
      if (want_sum) {
        if constexpr (sizeof(T) == 2) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(ct + g));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(ct + g) + 1);
          const short* sa = reinterpret_cast<const short*>(&a);
          const short* sb = reinterpret_cast<const short*>(&b);
#pragma unroll
          for (int k = 0; k < 8; ++k) { hu[k] = sa[k]; hu[8 + k] = sb[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) hu[k] = __float2int_rn((float)ct[g + k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int lab = l[k];
        if (lab < n_labels && (!mask || m[k] == mask_value)) {
          atomicAdd(&mycnt[lab], 1u);
          if (want_sum) atomicAdd(&mysum[lab], (uint32_t)(hu[k] + 32768));
        }
      }
    }
    v = begin + nvec * 16;
  }
  for (size_t i = v + threadIdx.x; i < end; i += SLICE_THREADS) {
    const size_t g = base + i;
    const int lab = labels[g];
    if (lab < n_labels && (!mask || mask[g] == mask_value)) {
      atomicAdd(&mycnt[lab], 1u);
      if (want_sum) {
        int h;
        if constexpr (sizeof(T) == 2) h = ct[g]; else h = __float2int_rn((float)ct[g]);
        atomicAdd(&mysum[lab], (uint32_t)(h + 32768));
      }
    }
  }
  __syncthreads();
  for (int lab = threadIdx.x; lab < n_labels; lab += SLICE_THREADS) {
    unsigned long long c = 0, s = 0;
    for (int w = 0; w < nwarps; ++w) { c += cnt[w * n_labels + lab]; s += sum[w * n_labels + lab]; }
    if (c) {
      if (counts) atomicAdd(&counts[(size_t)z * n_labels + lab], c);
      if (want_sum) {
        const long long signed_sum = (long long)s - 32768ll * (long long)c;
        atomicAdd(reinterpret_cast<unsigned long long*>(&hu_sums[(size_t)z * n_labels + lab]),
                  (unsigned long long)signed_sum);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ per-label HU hist
// measurements.py:74-123 : every statistic of a label's HU distribution is an exact function of its integer histogram.
template <typename T>
__global__ void __launch_bounds__(256)
label_hist_kernel(const T* __restrict__ ct, const uint8_t* __restrict__ labels, size_t n, int n_labels, int hu_min,
                  int n_bins, uint32_t* __restrict__ hist, uint32_t* __restrict__ out_of_range) {
  const size_t nvec = n / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t oor = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(labels) + i);
    if ((l4.x | l4.y | l4.z | l4.w) == 0) continue;  // all background: skip the CT read
    const uint8_t* l = reinterpret_cast<const uint8_t*>(&l4);
    int hu[16];
    if constexpr (sizeof(T) == 2) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(ct) + 2 * i);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(ct) + 2 * i + 1);
      const short* sa = reinterpret_cast<const short*>(&a);
      const short* sb = reinterpret_cast<const short*>(&b);
#pragma unroll
      for (int k = 0; k < 8; ++k) { hu[k] = sa[k]; hu[8 + k] = sb[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) hu[k] = __float2int_rn((float)ct[i * 16 + k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int lab = l[k];
      if (lab == 0 || lab >= n_labels) continue;
      const int bin = hu[k] - hu_min;
      if (bin >= 0 && bin < n_bins) atomicAdd(&hist[(size_t)lab * n_bins + bin], 1u);
      else ++oor;
    }
  }
  for (size_t v = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
    const int lab = labels[v];
    if (lab == 0 || lab >= n_labels) continue;
    int h;
    if constexpr (sizeof(T) == 2) h = ct[v]; else h = __float2int_rn((float)ct[v]);
    const int bin = h - hu_min;
    if (bin >= 0 && bin < n_bins) atomicAdd(&hist[(size_t)lab * n_bins + bin], 1u);
    else ++oor;
  }
  oor = __reduce_add_sync(0xffffffffu, oor);
  if ((threadIdx.x & 31) == 0 && oor) atomicAdd(out_of_range, oor);
}


// ------------------------------------------------------------------------------------------------ label-set mask
// compute/util.py:25-31 create_mask  +  measurements.py:29-39 (region minus fat: hu < lo OR hu > hi, strict)
//                                    /  measurements.py:134-141 (lung fat: lo <= hu <= hi, inclusive)
struct LabelSet {
  uint8_t in[256];
};
template <typename T>
__global__ void __launch_bounds__(256)
mask_window_kernel(const T* __restrict__ ct, const uint8_t* __restrict__ labels, size_t n, LabelSet set, float lo,
                   float hi, int inside, int use_window, uint8_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
    uint8_t m = set.in[labels[v]];
    if (m && use_window) {
      const float hu = (float)ct[v];
      const bool in = hu >= lo && hu <= hi;
      m = inside ? (in ? 1 : 0) : ((hu < lo || hu > hi) ? 1 : 0);
    }
    out[v] = m;
  }
}

// ------------------------------------------------------------------------------------------------ box erosion
// measurements.py:61-71 erode_region: footprint ones(6,6,6) padded at the end to 7^3 => window offsets -3..+2 per
// axis; skimage.binary_erosion treats voxels outside the image as foreground.  Separable: one pass per axis.
__global__ void __launch_bounds__(256)
erode_axis_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int d0, int d1, int d2, int axis,
                  int before, int after) {
  const size_t n = (size_t)d0 * d1 * d2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int dims[3] = {d0, d1, d2};
  const size_t strides[3] = {(size_t)d1 * d2, (size_t)d2, 1};
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
    const int c2 = (int)(v % d2), c1 = (int)((v / d2) % d1), c0 = (int)(v / ((size_t)d1 * d2));
    const int c = axis == 0 ? c0 : (axis == 1 ? c1 : c2);
    uint8_t m = 1;
    for (int o = -before; o <= after; ++o) {
      const int cc = c + o;
      if (cc < 0 || cc >= dims[axis]) continue;  // outside the image counts as foreground
      m &= in[v + (ptrdiff_t)o * (ptrdiff_t)strides[axis]] ? 1 : 0;
    }
    out[v] = m;
  }
}

}  // namespace boa

using namespace boa;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int boa_ct_normalize(const void* d_in, int in_dtype, size_t n, float lo, float hi, float mean, float std,
                                float* d_out, void* stream) {
  BOA_REQUIRE(d_in && d_out, "boa_ct_normalize: null pointer");
  BOA_REQUIRE(aligned16(d_in) && aligned16(d_out), "boa_ct_normalize: pointers must be 16-byte aligned");
  BOA_REQUIRE(in_dtype == BOA_DT_I16 || in_dtype == BOA_DT_F32, "boa_ct_normalize: bad dtype %d", in_dtype);
  if (n == 0) return BOA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std = fmaxf(std, 1e-8f);
  const int threads = 256, grid = grid_for(n / 8 + 1, threads);
  if (in_dtype == BOA_DT_I16)
    ct_normalize_kernel<short><<<grid, threads, 0, s>>>(static_cast<const short*>(d_in), n, lo, hi, mean, std, d_out);
  else
    ct_normalize_kernel<float><<<grid, threads, 0, s>>>(static_cast<const float*>(d_in), n, lo, hi, mean, std, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

static int check_patch_in_volume(const int32_t* patch, const int32_t* origin, const int32_t* vol) {
  for (int a = 0; a < 3; ++a)
    if (origin[a] < 0 || patch[a] <= 0 || origin[a] + patch[a] > vol[a]) {
      set_error("patch [%d,%d,%d]+[%d,%d,%d] outside volume [%d,%d,%d]", origin[0], origin[1], origin[2], patch[0],
                patch[1], patch[2], vol[0], vol[1], vol[2]);
      return BOA_ERR_ARG;
    }
  return BOA_OK;
}

extern "C" int boa_accumulate_patch(const float* d_logits, int C, const int32_t* patch, const int32_t* origin,
                                    const float* d_gaussian, float* d_logits_acc, const int32_t* vol_shape,
                                    void* stream) {
  BOA_REQUIRE(d_logits && patch && origin && d_gaussian && d_logits_acc && vol_shape, "boa_accumulate_patch: null");
  BOA_REQUIRE(C > 0, "boa_accumulate_patch: C must be positive");
  if (int r = check_patch_in_volume(patch, origin, vol_shape)) return r;
  const size_t pv = (size_t)patch[0] * patch[1] * patch[2];
  const size_t vv = (size_t)vol_shape[0] * vol_shape[1] * vol_shape[2];
  accumulate_patch_kernel<<<grid_for(pv, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_logits, C, patch[0], patch[1], patch[2], origin[0], origin[1], origin[2], d_gaussian, d_logits_acc,
      vol_shape[1], vol_shape[2], vv);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_accumulate_weights(const int32_t* h_origins, int n_patches, const int32_t* patch,
                                      const float* d_gaussian, float* d_weight_acc, const int32_t* vol_shape,
                                      void* stream) {
  BOA_REQUIRE(h_origins && patch && d_gaussian && d_weight_acc && vol_shape, "boa_accumulate_weights: null");
  const size_t pv = (size_t)patch[0] * patch[1] * patch[2];
  for (int p = 0; p < n_patches; ++p) {
    const int32_t* o = h_origins + 3 * p;
    if (int r = check_patch_in_volume(patch, o, vol_shape)) return r;
    accumulate_weight_kernel<<<grid_for(pv, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        patch[0], patch[1], patch[2], o[0], o[1], o[2], d_gaussian, d_weight_acc, vol_shape[1], vol_shape[2]);
    BOA_CHECK_LAUNCH();
  }
  return BOA_OK;
}

extern "C" int boa_finalize_argmax(const float* d_logits_acc, const float* d_weight_acc, int C, size_t V,
                                   const uint8_t* h_lut, int overwrite_nonzero_only, uint8_t* d_label_inout,
                                   int32_t* d_nonfinite, void* stream) {
  BOA_REQUIRE(d_logits_acc && d_weight_acc && h_lut && d_label_inout && d_nonfinite, "boa_finalize_argmax: null");
  BOA_REQUIRE(C > 0 && C <= 256, "boa_finalize_argmax: C=%d out of range", C);
  if (V == 0) return BOA_OK;
  Lut256 lut;
  for (int c = 0; c < 256; ++c) lut.v[c] = c < C ? h_lut[c] : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // float4 channel loads need every channel plane 16-byte aligned (V % 4 == 0) and the label map 4-byte aligned;
  // anything else (odd-sized volumes, slabs of a sharded volume) takes the scalar kernel
  const bool vec = V % 4 == 0 && aligned16(d_logits_acc) && aligned16(d_weight_acc) &&
                   (reinterpret_cast<uintptr_t>(d_label_inout) & 3) == 0;
  const LocalSrc src{d_logits_acc, V};
  if (vec)
    finalize_argmax_kernel<8, LocalSrc><<<grid_for(V / 4, 256, 8), 256, 0, s>>>(
        src, d_weight_acc, C, V, lut, overwrite_nonzero_only, d_label_inout, d_nonfinite);
  else
    finalize_argmax_scalar_kernel<LocalSrc><<<grid_for(V, 256, 8), 256, 0, s>>>(
        src, d_weight_acc, C, V, lut, overwrite_nonzero_only, d_label_inout, d_nonfinite);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// `predicted_logits /= n_predictions` + the isinf check of the logits-returning entry
// (predict_from_raw_data.py:620-625; the fold mean of predict_logits_from_preprocessed_data :494-500 as one divisor):
// acc[c][v] = acc[c][v] / (w[v] * folds), IEEE division, non-finite quotients counted.
__global__ void __launch_bounds__(256)
normalize_logits_kernel(float* __restrict__ acc, const float* __restrict__ w, int C, size_t V, float folds,
                        int32_t* __restrict__ nonfinite) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int bad = 0;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += stride) {
    const float d = __fmul_rn(w[v], folds);
    for (int c = 0; c < C; ++c) {
      const float q = __fdiv_rn(acc[(size_t)c * V + v], d);
      acc[(size_t)c * V + v] = q;
      bad |= !isfinite(q);
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAdd(nonfinite, 1);
}

extern "C" int boa_normalize_logits(float* d_logits_acc, const float* d_weight_acc, int C, size_t V, float folds,
                                    int32_t* d_nonfinite, void* stream) {
  BOA_REQUIRE(d_logits_acc && d_weight_acc && d_nonfinite, "boa_normalize_logits: null");
  BOA_REQUIRE(C > 0 && folds > 0.0f, "boa_normalize_logits: C=%d folds=%f", C, (double)folds);
  if (V == 0) return BOA_OK;
  normalize_logits_kernel<<<grid_for(V, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_logits_acc, d_weight_acc, C, V, folds, d_nonfinite);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// Finalize on a different grid than the network's (nnU-Net resampled the volume to the plan's spacing): ResampledSrc.
extern "C" int boa_finalize_argmax_resampled(const float* d_logits_acc, const float* d_weight_acc, int C,
                                             const int32_t* net_shape, const int32_t* out_shape, int separate_z,
                                             const uint8_t* h_lut, int overwrite_nonzero_only, uint8_t* d_label_inout,
                                             int32_t* d_nonfinite, void* stream) {
  BOA_REQUIRE(d_logits_acc && d_weight_acc && net_shape && out_shape && h_lut && d_label_inout && d_nonfinite,
              "boa_finalize_argmax_resampled: null");
  BOA_REQUIRE(C > 0 && C <= 256, "boa_finalize_argmax_resampled: C=%d out of range", C);
  for (int k = 0; k < 3; ++k)
    BOA_REQUIRE(net_shape[k] >= 1 && out_shape[k] >= 1, "boa_finalize_argmax_resampled: bad shape");
  ResampledSrc src{d_logits_acc, d_weight_acc, net_shape[0], net_shape[1], net_shape[2],
                   out_shape[0],  out_shape[1],  out_shape[2], separate_z ? 1 : 0};
  Lut256 lut;
  for (int c = 0; c < 256; ++c) lut.v[c] = c < C ? h_lut[c] : 0;
  const size_t V = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
  finalize_argmax_scalar_kernel<ResampledSrc><<<grid_for(V, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, nullptr, C, V, lut, overwrite_nonzero_only, d_label_inout, d_nonfinite);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// Multi-GPU: exchange + rank-ordered reduction + finalize of one dim-0 slab in ONE kernel over peer memory (PeerSrc).
extern "C" int boa_reduce_finalize_peers(const void* const* d_peer_bases, const int32_t* zlo, const int32_t* zhi,
                                         int n_ranks, int slab_lo, int slab_hi, int Y, int X,
                                         const float* d_weight_slab, int C, const uint8_t* h_lut,
                                         int overwrite_nonzero_only, uint8_t* d_label_slab, int32_t* d_nonfinite,
                                         void* stream) {
  BOA_REQUIRE(d_peer_bases && zlo && zhi && d_weight_slab && h_lut && d_label_slab && d_nonfinite,
              "boa_reduce_finalize_peers: null");
  BOA_REQUIRE(n_ranks >= 1 && n_ranks <= MAX_PEERS, "boa_reduce_finalize_peers: %d ranks (max %d)", n_ranks, MAX_PEERS);
  BOA_REQUIRE(C > 0 && C <= 256 && slab_hi >= slab_lo && Y > 0 && X > 0, "boa_reduce_finalize_peers: bad shape");
  if (slab_hi == slab_lo) return BOA_OK;
  PeerSrc src;
  src.n = n_ranks; src.z0 = slab_lo; src.plane = (size_t)Y * X;
  bool vec = src.plane % 4 == 0 && aligned16(d_weight_slab) && (reinterpret_cast<uintptr_t>(d_label_slab) & 3) == 0;
  for (int r = 0; r < MAX_PEERS; ++r) {
    src.base[r] = r < n_ranks ? static_cast<const float*>(d_peer_bases[r]) : nullptr;
    src.zlo[r] = r < n_ranks ? zlo[r] : 0;
    src.zhi[r] = r < n_ranks ? zhi[r] : 0;
    if (r < n_ranks && zhi[r] > zlo[r]) {
      BOA_REQUIRE(src.base[r], "boa_reduce_finalize_peers: rank %d has patches but no buffer", r);
      vec = vec && aligned16(src.base[r]);
    }
  }
  Lut256 lut;
  for (int c = 0; c < 256; ++c) lut.v[c] = c < C ? h_lut[c] : 0;
  const size_t V = (size_t)(slab_hi - slab_lo) * src.plane;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec)
    finalize_argmax_kernel<4, PeerSrc><<<grid_for(V / 4, 256, 8), 256, 0, s>>>(
        src, d_weight_slab, C, V, lut, overwrite_nonzero_only, d_label_slab, d_nonfinite);
  else
    finalize_argmax_scalar_kernel<PeerSrc><<<grid_for(V, 256, 8), 256, 0, s>>>(
        src, d_weight_slab, C, V, lut, overwrite_nonzero_only, d_label_slab, d_nonfinite);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_tissue_subclassify(const void* d_ct, int ct_dtype, const uint8_t* d_regions, size_t n,
                                      uint8_t* d_tissues, void* stream) {
  BOA_REQUIRE(d_ct && d_regions && d_tissues, "boa_tissue_subclassify: null pointer");
  BOA_REQUIRE(aligned16(d_ct) && aligned16(d_regions) && aligned16(d_tissues), "boa_tissue_subclassify: misaligned");
  BOA_REQUIRE(ct_dtype == BOA_DT_I16 || ct_dtype == BOA_DT_F32, "boa_tissue_subclassify: bad dtype %d", ct_dtype);
  if (n == 0) return BOA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n / 16 + 1, 256);
  if (ct_dtype == BOA_DT_I16)
    tissue_kernel<short><<<grid, 256, 0, s>>>(static_cast<const short*>(d_ct), d_regions, n, d_tissues);
  else
    tissue_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_ct), d_regions, n, d_tissues);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// ------------------------------------------------------------------------------------------ in-plane median
// scipy.ndimage.median_filter(image, size=[1, 3, 3]) of subclassify_tissues(median_filtering=True)
// (_external/body_composition_analysis/tissue/subclassification.py:20-36): 3x3 median inside every slice, boundary
// mode "reflect" (index -1 -> 0, n -> n-1).  Exact for integer HU: a 19-exchange median-of-9 network on int16.
namespace boa {
__device__ __forceinline__ void med_cswap(int& a, int& b) {
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  a = lo; b = hi;
}
__global__ void __launch_bounds__(256)
median3x3_kernel(const int16_t* __restrict__ in, int D, int H, int W, int16_t* __restrict__ out) {
  const size_t n = (size_t)D * H * W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int16_t* sl = in + (i - (size_t)y * W - x);
    const int xs[3] = {x > 0 ? x - 1 : 0, x, x + 1 < W ? x + 1 : W - 1};
    const int ys[3] = {y > 0 ? y - 1 : 0, y, y + 1 < H ? y + 1 : H - 1};
    int p[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) p[a * 3 + b] = sl[(size_t)ys[a] * W + xs[b]];
    // Paeth's median-of-9 network
    med_cswap(p[1], p[2]); med_cswap(p[4], p[5]); med_cswap(p[7], p[8]);
    med_cswap(p[0], p[1]); med_cswap(p[3], p[4]); med_cswap(p[6], p[7]);
    med_cswap(p[1], p[2]); med_cswap(p[4], p[5]); med_cswap(p[7], p[8]);
    med_cswap(p[0], p[3]); med_cswap(p[5], p[8]); med_cswap(p[4], p[7]);
    med_cswap(p[3], p[6]); med_cswap(p[1], p[4]); med_cswap(p[2], p[5]);
    med_cswap(p[4], p[7]); med_cswap(p[4], p[2]); med_cswap(p[6], p[4]);
    med_cswap(p[4], p[2]);
    out[i] = (int16_t)p[4];
  }
}
}  // namespace boa

extern "C" int boa_median3x3_slices(const int16_t* d_in, const int32_t* shape, int16_t* d_out, void* stream) {
  BOA_REQUIRE(d_in && d_out && shape && d_in != d_out, "boa_median3x3_slices: bad pointers (in place is not supported)");
  BOA_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "boa_median3x3_slices: bad shape");
  const size_t n = (size_t)shape[0] * shape[1] * shape[2];
  boa::median3x3_kernel<<<boa::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, shape[0], shape[1],
                                                                                           shape[2], d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_slice_label_stats(const uint8_t* d_labels, const uint8_t* d_mask, int mask_value, const void* d_ct,
                                     int ct_dtype, int Z, size_t slice_voxels, int n_labels, uint64_t* d_counts,
                                     int64_t* d_hu_sums, void* stream) {
  BOA_REQUIRE(d_labels, "boa_slice_label_stats: null labels");
  BOA_REQUIRE(d_counts || d_hu_sums, "boa_slice_label_stats: no output requested");
  BOA_REQUIRE(!d_hu_sums || d_ct, "boa_slice_label_stats: hu_sums needs ct");
  BOA_REQUIRE(n_labels > 0 && n_labels <= 256, "boa_slice_label_stats: n_labels=%d out of range", n_labels);
  BOA_REQUIRE(ct_dtype == BOA_DT_I16 || ct_dtype == BOA_DT_F32, "boa_slice_label_stats: bad dtype %d", ct_dtype);
  if (Z <= 0 || slice_voxels == 0) return BOA_OK;
  const int chunks = (int)((slice_voxels + SLICE_CHUNK - 1) / SLICE_CHUNK);
  const size_t smem = (size_t)2 * (SLICE_THREADS / 32) * n_labels * sizeof(uint32_t);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto c = reinterpret_cast<unsigned long long*>(d_counts);
  auto h = reinterpret_cast<long long*>(d_hu_sums);
  if (ct_dtype == BOA_DT_I16)
    slice_stats_kernel<short><<<Z * chunks, SLICE_THREADS, smem, s>>>(d_labels, d_mask, mask_value,
                                                                     static_cast<const short*>(d_ct), slice_voxels,
                                                                     n_labels, chunks, c, h);
  else
    slice_stats_kernel<float><<<Z * chunks, SLICE_THREADS, smem, s>>>(d_labels, d_mask, mask_value,
                                                                     static_cast<const float*>(d_ct), slice_voxels,
                                                                     n_labels, chunks, c, h);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_label_hu_hist(const void* d_ct, int ct_dtype, const uint8_t* d_labels, size_t n, int n_labels,
                                 int hu_min, int n_bins, uint32_t* d_hist, uint32_t* d_out_of_range, void* stream) {
  BOA_REQUIRE(d_ct && d_labels && d_hist && d_out_of_range, "boa_label_hu_hist: null pointer");
  BOA_REQUIRE(aligned16(d_ct) && aligned16(d_labels), "boa_label_hu_hist: misaligned");
  BOA_REQUIRE(n_labels > 0 && n_labels <= 256 && n_bins > 0, "boa_label_hu_hist: bad sizes");
  BOA_REQUIRE(ct_dtype == BOA_DT_I16 || ct_dtype == BOA_DT_F32, "boa_label_hu_hist: bad dtype %d", ct_dtype);
  if (n == 0) return BOA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n / 16 + 1, 256);
  if (ct_dtype == BOA_DT_I16)
    label_hist_kernel<short><<<grid, 256, 0, s>>>(static_cast<const short*>(d_ct), d_labels, n, n_labels, hu_min,
                                                  n_bins, d_hist, d_out_of_range);
  else
    label_hist_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_ct), d_labels, n, n_labels, hu_min,
                                                  n_bins, d_hist, d_out_of_range);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_mask_label_minus_window(const void* d_ct, int ct_dtype, const uint8_t* d_labels, size_t n,
                                           const uint8_t* h_label_set, int lo, int hi, int mode, uint8_t* d_mask,
                                           void* stream) {
  BOA_REQUIRE(d_labels && h_label_set && d_mask, "boa_mask_label_minus_window: null pointer");
  BOA_REQUIRE(mode >= 0 && mode <= 2, "boa_mask_label_minus_window: mode must be 0 (labels only), 1 (inside) or 2 (outside)");
  BOA_REQUIRE(mode == 0 || d_ct, "boa_mask_label_minus_window: window modes need the CT");
  BOA_REQUIRE(ct_dtype == BOA_DT_I16 || ct_dtype == BOA_DT_F32, "boa_mask_label_minus_window: bad dtype %d", ct_dtype);
  if (n == 0) return BOA_OK;
  LabelSet set;
  for (int i = 0; i < 256; ++i) set.in[i] = h_label_set[i] ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, 256);
  if (ct_dtype == BOA_DT_I16)
    mask_window_kernel<short><<<grid, 256, 0, s>>>(static_cast<const short*>(d_ct), d_labels, n, set, (float)lo,
                                                   (float)hi, mode == 1, mode != 0, d_mask);
  else
    mask_window_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_ct), d_labels, n, set, (float)lo,
                                                   (float)hi, mode == 1, mode != 0, d_mask);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_erode_box(const uint8_t* d_mask, const int32_t* shape, int before, int after, uint8_t* d_tmp,
                             uint8_t* d_out, void* stream) {
  BOA_REQUIRE(d_mask && shape && d_tmp && d_out, "boa_erode_box: null pointer");
  BOA_REQUIRE(before >= 0 && after >= 0, "boa_erode_box: negative window");
  BOA_REQUIRE(d_tmp != d_mask && d_tmp != d_out, "boa_erode_box: d_tmp must not alias the input or output");
  const size_t n = (size_t)shape[0] * shape[1] * shape[2];
  if (n == 0) return BOA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, 256);
  erode_axis_kernel<<<grid, 256, 0, s>>>(d_mask, d_out, shape[0], shape[1], shape[2], 2, before, after);
  BOA_CHECK_LAUNCH();
  erode_axis_kernel<<<grid, 256, 0, s>>>(d_out, d_tmp, shape[0], shape[1], shape[2], 1, before, after);
  BOA_CHECK_LAUNCH();
  erode_axis_kernel<<<grid, 256, 0, s>>>(d_tmp, d_out, shape[0], shape[1], shape[2], 0, before, after);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}
