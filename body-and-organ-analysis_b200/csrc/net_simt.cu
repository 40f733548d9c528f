// Elementwise network passes (patch extraction, InstanceNorm finalisation, normalise + LeakyReLU) and the SIMT
// reference kernels (direct conv / transposed conv / head).  The SIMT convs exist as an on-device cross-check for
// the tcgen05 kernels and as the implementation of ops that have not moved to tensor cores yet; they are selected
// explicitly, never as a silent fallback.
#include <stdlib.h>
#include <algorithm>
#include "net_kernels.cuh"

namespace boa {

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return raw;
}

// ------------------------------------------------------------------------------------------ patch extraction
// predict_from_raw_data.py:568-571 (producer thread cutting `data[sl]`): fp32 volume -> fp16 C8 with 16 channels,
// channel 0 = voxel value, channels 1..15 = 0 (the first conv runs on the tensor-core kernel with K padded to 16).
// Padding batch items (b >= n_valid) re-read patch 0 so that every launch does identical work.
__global__ void __launch_bounds__(256)
extract_patches_kernel(const FwdCall* __restrict__ call, int n, int p0, int p1, int p2, uint4* __restrict__ out) {
  const float* __restrict__ vol = call->vol;
  const int d1 = call->d1, d2 = call->d2, n_valid = call->n_valid;
  const size_t pv = (size_t)p0 * p1 * p2;
  const size_t total = (size_t)n * 2 * pv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t v = i % pv;
    const int g = (int)((i / pv) % 2);
    const int b = (int)(i / (2 * pv));
    uint4 o = make_uint4(0, 0, 0, 0);
    if (g == 0) {
      const int bb = b < n_valid ? b : 0;
      const int k = (int)(v % p2), j = (int)((v / p2) % p1), ii = (int)(v / ((size_t)p2 * p1));
      const int o0 = call->origins[bb][0], o1 = call->origins[bb][1], o2 = call->origins[bb][2];
      const float x = __ldg(vol + ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k));
      o.x = (uint32_t)__half_as_ushort(__float2half_rn(x));
    }
    out[i] = o;
  }
}

// Neighbour variant: C8 with 2 groups whose channels 0..8 are the in-plane neighbours of the voxel (channel 0 = the
// voxel, 1..4 = taps 0..3, 5..8 = taps 5..8 with tap = dy*3+dx), zero
// outside the patch (the conv's zero padding is relative to the patch).  src == nullptr: cut from the volume.
__global__ void __launch_bounds__(256)
extract_patches_nb9_kernel(const FwdCall* __restrict__ call, const float* __restrict__ src, int n, int p0, int p1,
                           int p2, uint4* __restrict__ out) {
  const size_t pv = (size_t)p0 * p1 * p2;
  const size_t total = (size_t)n * pv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float* vol = src ? src : call->vol;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t v = i % pv;
    const int b = (int)(i / pv);
    const int k = (int)(v % p2), j = (int)((v / p2) % p1), ii = (int)(v / ((size_t)p2 * p1));
    size_t base;
    size_t sy, sz;
    if (src) {
      base = (size_t)b * pv; sy = (size_t)p2; sz = (size_t)p2 * p1;
    } else {
      const int bb = b < call->n_valid ? b : 0;
      sy = (size_t)call->d2; sz = (size_t)call->d2 * call->d1;
      base = (size_t)call->origins[bb][0] * sz + (size_t)call->origins[bb][1] * sy + (size_t)call->origins[bb][2];
    }
    float f[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) f[t] = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const int t = c == 0 ? 4 : (c <= 4 ? c - 1 : c);  // channel 0 = the voxel itself (what the SIMT cross-check reads)
      const int jj = j + t / 3 - 1, kk = k + t % 3 - 1;
      if (jj >= 0 && jj < p1 && kk >= 0 && kk < p2) f[c] = __ldg(vol + base + (size_t)ii * sz + (size_t)jj * sy + kk);
    }
    float lo[8], hi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { lo[e] = f[e]; hi[e] = f[8 + e]; }
    out[((size_t)b * 2) * pv + v] = pack8(lo);
    out[((size_t)b * 2 + 1) * pv + v] = pack8(hi);
  }
}

// Plain variant for the dedicated first-layer kernel: fp16 [n][p0][p1][p2].
__global__ void __launch_bounds__(256)
extract_patches_plain_kernel(const FwdCall* __restrict__ call, int n, int p0, int p1, int p2,
                             __half* __restrict__ out) {
  const float* __restrict__ vol = call->vol;
  const int d1 = call->d1, d2 = call->d2, n_valid = call->n_valid;
  const size_t pv = (size_t)p0 * p1 * p2;
  const size_t total = (size_t)n * pv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t v = i % pv;
    const int b = (int)(i / pv);
    const int bb = b < n_valid ? b : 0;
    const int k = (int)(v % p2), j = (int)((v / p2) % p1), ii = (int)(v / ((size_t)p2 * p1));
    const int o0 = call->origins[bb][0], o1 = call->origins[bb][1], o2 = call->origins[bb][2];
    out[i] = __float2half_rn(__ldg(vol + ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k)));
  }
}

__global__ void __launch_bounds__(256)
pack_patches_plain_kernel(const float* __restrict__ in, size_t total, __half* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    out[i] = __float2half_rn(__ldg(in + i));
}

// ------------------------------------------------------------------------------------------ first layer (Cin = 1)
// Conv3d(1 -> COUT, 3x3x3, stride 1, pad 1) of the first encoder block.  K = 27 is far too thin for the tensor
// cores (it would run with K padded to 16 channels = 16x the work); as a direct convolution it is FP32-FMA work of
// 27*COUT per voxel with a 2-byte read and a 2*COUT-byte write.  One thread = one voxel, all COUT channels in
// registers; weights [27][COUT] broadcast from shared memory as float4; InstanceNorm sum / sum^2 reduced per block.
template <int COUT>
__global__ void __launch_bounds__(256)
conv_first_kernel(const __half* __restrict__ in, const float* __restrict__ w /*[COUT][27]*/,
                  const float* __restrict__ bias, uint4* __restrict__ out, double* __restrict__ stats, int D, int H,
                  int W) {
  constexpr int ZT = 4;  // voxels per thread along z: each weight vector read from shared memory feeds 4 FMAs
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ float sred[2][8][COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += 256) sw[i] = w[(i % COUT) * 27 + i / COUT];
  __syncthreads();
  const int zchunks = (D + ZT - 1) / ZT;
  const int b = blockIdx.z / zchunks, z0 = (blockIdx.z % zchunks) * ZT;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool valid = x < W && y < H;
  const size_t vox = (size_t)D * H * W;
  const __half* src = in + (size_t)b * vox;
  float acc[ZT][COUT];
#pragma unroll
  for (int j = 0; j < ZT; ++j)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[j][c] = 0.f;
  if (valid) {
#pragma unroll
    for (int t2 = 0; t2 < 9; ++t2) {
      const int yi = y + t2 / 3 - 1, xi = x + t2 % 3 - 1;
      const bool inplane = yi >= 0 && yi < H && xi >= 0 && xi < W;
      float v[ZT + 2];
#pragma unroll
      for (int k = 0; k < ZT + 2; ++k) {
        const int zi = z0 + k - 1;
        v[k] = (inplane && zi >= 0 && zi < D) ? __half2float(__ldg(src + ((size_t)zi * H + yi) * W + xi)) : 0.f;
      }
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const float4* wr = reinterpret_cast<const float4*>(sw + (dz * 9 + t2) * COUT);
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
          const float4 w4 = wr[q];
#pragma unroll
          for (int j = 0; j < ZT; ++j) {
            acc[j][4 * q] = fmaf(v[j + dz], w4.x, acc[j][4 * q]);
            acc[j][4 * q + 1] = fmaf(v[j + dz], w4.y, acc[j][4 * q + 1]);
            acc[j][4 * q + 2] = fmaf(v[j + dz], w4.z, acc[j][4 * q + 2]);
            acc[j][4 * q + 3] = fmaf(v[j + dz], w4.w, acc[j][4 * q + 3]);
          }
        }
      }
    }
  }
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
  if (valid) {
#pragma unroll
    for (int j = 0; j < ZT; ++j) {
      const int z = z0 + j;
      if (z < D) {
        uint4* dst = out + ((size_t)b * (COUT / 8)) * vox + ((size_t)z * H + y) * W + x;
#pragma unroll
        for (int g = 0; g < COUT / 8; ++g) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = acc[j][g * 8 + e] + __ldg(bias + g * 8 + e);
            s1[g * 8 + e] += f[e];
            s2[g * 8 + e] = fmaf(f[e], f[e], s2[g * 8 + e]);
          }
          dst[(size_t)g * vox] = pack8(f);
        }
      }
    }
  }
  // block reduction of sum / sum^2 per channel: warp shuffle, then the 8 warps through shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    float a1 = s1[c], a2 = s2[c];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (lane == 0) { sred[0][warp][c] = a1; sred[1][warp][c] = a2; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * COUT) {
    const int which = threadIdx.x / COUT, c = threadIdx.x % COUT;
    double t = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += (double)sred[which][wv][c];
    atomicAdd(&stats[((size_t)b * COUT + c) * 2 + which], t);
  }
}

// Generic: fp32 [n][cin][P] -> C8 with `groups` channel groups (channels >= cin zero).
__global__ void __launch_bounds__(256)
pack_patches_kernel(const float* __restrict__ in, int n, int cin, size_t pv, int groups, uint4* __restrict__ out) {
  const size_t total = (size_t)n * groups * pv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t v = i % pv;
    const int g = (int)((i / pv) % groups);
    const int b = (int)(i / ((size_t)groups * pv));
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      f[e] = c < cin ? __ldg(in + ((size_t)b * cin + c) * pv + v) : 0.f;
    }
    out[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------ InstanceNorm finalise
// nn.InstanceNorm3d(eps, affine=True): biased variance over D*H*W per (sample, channel) (plans_handler.py:72-76).
// scale = gamma * rstd ; shift = beta - mean * gamma * rstd   (fp64 internally)
// scale2 / shift2 (optional): a second copy at [b * stride2 + off2 + c] - the rows of a decoder concat's combined
// scale / shift table that belong to this (skip) producer.
__global__ void stats_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int B, int C, double n_vox, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift,
                                      float* __restrict__ scale2, float* __restrict__ shift2, int stride2, int off2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C;
  const double mean = stats[2 * i] / n_vox;
  double var = stats[2 * i + 1] / n_vox - mean * mean;
  if (var < 0) var = 0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma[c];
  const float sc = (float)(g * rstd), sh = (float)((double)beta[c] - mean * g * rstd);
  scale[i] = sc;
  shift[i] = sh;
  if (scale2) {
    const int j = (i / C) * stride2 + off2 + c;
    scale2[j] = sc;
    shift2[j] = sh;
  }
}

// ------------------------------------------------------------------------------------------ normalise + LeakyReLU
__global__ void __launch_bounds__(256)
norm_lrelu_kernel(const uint4* __restrict__ raw, int groups, int D, int H, int W, const float* __restrict__ scale,
                  const float* __restrict__ shift, float slope, uint4* __restrict__ dst, int dst_groups_total,
                  int dst_group_off, uint4* __restrict__ s2d) {
  // blockIdx.y = b * groups + g : the 8 scale/shift pairs are loaded once per thread
  const int bg = blockIdx.y, b = bg / groups, g = bg % groups;
  const size_t vox = (size_t)D * H * W;
  float a[8], sh[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + (size_t)bg * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(scale + (size_t)bg * 8) + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + (size_t)bg * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(shift + (size_t)bg * 8) + 1);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
  }
  const uint4* __restrict__ src = raw + (size_t)bg * vox;
  uint4* __restrict__ out = dst ? dst + ((size_t)b * dst_groups_total + dst_group_off + g) * vox : nullptr;
  uint4* __restrict__ out2 = s2d ? s2d + ((size_t)b * 8 * groups + g) * (vox >> 3) : nullptr;
  const size_t phase_stride = (size_t)groups * (vox >> 3);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < vox; v0 += U * stride) {
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v0 + u * stride < vox) r[u] = __ldcs(src + v0 + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = v0 + u * stride;
      if (v >= vox) break;
      const uint4 o = xform8(r[u], a, sh, slope);
      if (out) out[v] = o;
      if (out2) {
        const int x = (int)(v % W), y = (int)((v / W) % H), z = (int)(v / ((size_t)W * H));
        const int ph = ((z & 1) * 2 + (y & 1)) * 2 + (x & 1);
        const size_t hv = ((size_t)(z >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
        out2[(size_t)ph * phase_stride + hv] = o;
      }
    }
  }
}

// Co-resident ("thin") variant.  The persistent tcgen05 conv CTAs hold ~48 K of the SM's 64 K registers and almost all
// of its shared memory; a pass of the OTHER lane only overlaps them if one of its blocks fits into what is left
// (<= 16 K registers: 256 threads x 64) AND never occupies more than that, so that a conv CTA can always start next to
// it.  Hence: one persistent block per SM, 8 independent 16-byte loads in flight per thread (32 KB per SM) instead of
// occupancy to cover the HBM latency.  Same arithmetic as norm_lrelu_kernel.
constexpr int THIN_U = 8;
__global__ void __launch_bounds__(256, 4)
norm_lrelu_thin_kernel(const uint4* __restrict__ raw, int planes, int groups, int D, int H, int W,
                       const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                       uint4* __restrict__ dst, int dst_groups_total, int dst_group_off, uint4* __restrict__ s2d) {
  const int vox = D * H * W;  // voxels of one (batch item, channel group) plane: < 2^31
  constexpr int CHUNK = 256 * THIN_U;
  const int chunks = (vox + CHUNK - 1) / CHUNK;
  const int items = planes * chunks;
  __shared__ float ws[8][16];
  const int lane = threadIdx.x & 31;
  float* wsl = ws[threadIdx.x >> 5];
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int bg = item / chunks;
    const int vbase = (item - bg * chunks) * CHUNK + (int)threadIdx.x;
    const int nrem = vox - vbase;  // element u of this thread exists iff u * 256 < nrem
    const uint4* __restrict__ src = raw + (size_t)bg * vox + vbase;
    uint4 r[THIN_U];
#pragma unroll
    for (int u = 0; u < THIN_U; ++u)
      if (u * 256 < nrem) r[u] = __ldcs(src + u * 256);
    // the 8 scale / shift pairs of this plane live in a warp-private shared-memory slot (broadcast reads), not in 16
    // registers: the register budget goes to the loads in flight
    __syncwarp();
    if (lane < 16) wsl[lane] = lane < 8 ? __ldg(scale + (size_t)bg * 8 + lane) : __ldg(shift + (size_t)bg * 8 + lane - 8);
    __syncwarp();
    const int b = bg / groups, g = bg - b * groups;
    uint4* __restrict__ out = dst ? dst + ((size_t)b * dst_groups_total + dst_group_off + g) * vox + vbase : nullptr;
    uint4* __restrict__ out2 = s2d ? s2d + ((size_t)b * 8 * groups + g) * (vox >> 3) : nullptr;
    const int phase_stride = groups * (vox >> 3);
#pragma unroll
    for (int u = 0; u < THIN_U; ++u) {
      if (u * 256 < nrem) {
        float a[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] = wsl[e]; sh[e] = wsl[8 + e]; }
        const uint4 o = xform8(r[u], a, sh, slope);
        if (out) out[u * 256] = o;
        if (out2) {
          const int v = vbase + u * 256;
          const int x = v % W, y = (v / W) % H, z = v / (W * H);
          const int ph = ((z & 1) * 2 + (y & 1)) * 2 + (x & 1);
          const int hv = ((z >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
          out2[(size_t)ph * phase_stride + hv] = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ SIMT direct conv
// One thread = one output voxel x 8 output channels.  Accumulates InstanceNorm statistics with fp64 atomics.
struct Int3 { int z, y, x; };

__global__ void __launch_bounds__(128)
conv_simt_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int Di, int Hi, int Wi,
                 const float* __restrict__ w, const float* __restrict__ bias, int cin_w, int Cout, Int3 ks,
                 Int3 stride, uint4* __restrict__ out, int out_groups_total, int out_group_off, int Do, int Ho, int Wo,
                 double* __restrict__ stats, int B, InXform xf) {
  const size_t ovox = (size_t)Do * Ho * Wo;
  const int ogroups = Cout / 8;
  const size_t total = (size_t)B * ogroups * ovox;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < total;
  const size_t ii = active ? i : total - 1;
  const size_t v = ii % ovox;
  const int og = (int)((ii / ovox) % ogroups);
  const int b = (int)(ii / ((size_t)ogroups * ovox));
  const int xo = (int)(v % Wo), yo = (int)((v / Wo) % Ho), zo = (int)(v / ((size_t)Wo * Ho));
  const int k3 = ks.z * ks.y * ks.x;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  const size_t ivox = (size_t)Di * Hi * Wi;
  for (int tap = 0; tap < k3; ++tap) {
    const int dz = tap / (ks.y * ks.x), dy = (tap / ks.x) % ks.y, dx = tap % ks.x;
    const int zi = zo * stride.z + dz - ks.z / 2, yi = yo * stride.y + dy - ks.y / 2,
              xi = xo * stride.x + dx - ks.x / 2;
    if (zi < 0 || zi >= Di || yi < 0 || yi >= Hi || xi < 0 || xi >= Wi) continue;
    const size_t iv = ((size_t)zi * Hi + yi) * Wi + xi;
    for (int g = 0; g * 8 < cin_w; ++g) {
      float f[8];
      uint4 raw = __ldg(in + ((size_t)b * in_groups_total + in_group_off + g) * ivox + iv);
      if (xf.scale && g >= xf.ident_groups) {  // RAW producer output: its InstanceNorm affine + LeakyReLU on the fly
        float a[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a[e] = __ldg(xf.scale + (size_t)b * xf.channels + g * 8 + e);
          sh[e] = __ldg(xf.shift + (size_t)b * xf.channels + g * 8 + e);
        }
        raw = xform8(raw, a, sh, xf.slope);
      }
      unpack8(raw, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ci = g * 8 + e;
        if (ci >= cin_w) break;
#pragma unroll
        for (int o = 0; o < 8; ++o)
          acc[o] = fmaf(f[e], __ldg(w + ((size_t)(og * 8 + o) * cin_w + ci) * k3 + tap), acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] += bias[og * 8 + o];
  if (active) out[((size_t)b * out_groups_total + out_group_off + og) * ovox + v] = pack8(acc);
  if (stats) {
    // whole warp in the same (b, og)?  then shuffle-reduce, else per-thread atomics
    const unsigned key = (unsigned)(b * ogroups + og);
    const bool uniform = __all_sync(0xffffffffu, active && key == __shfl_sync(0xffffffffu, key, 0));
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      double s1 = active ? (double)acc[o] : 0.0, s2 = active ? (double)acc[o] * acc[o] : 0.0;
      if (uniform) {
        for (int off = 16; off; off >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, off);
          s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if ((threadIdx.x & 31) == 0) {
          atomicAdd(&stats[((size_t)b * Cout + og * 8 + o) * 2], s1);
          atomicAdd(&stats[((size_t)b * Cout + og * 8 + o) * 2 + 1], s2);
        }
      } else if (active) {
        atomicAdd(&stats[((size_t)b * Cout + og * 8 + o) * 2], s1);
        atomicAdd(&stats[((size_t)b * Cout + og * 8 + o) * 2 + 1], s2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ SIMT transposed conv
__global__ void __launch_bounds__(128)
tconv_simt_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int Di, int Hi, int Wi,
                  const float* __restrict__ w, const float* __restrict__ bias, int Cin, int Cout, Int3 st,
                  uint4* __restrict__ out, int out_groups_total, int out_group_off, int B, InXform xf) {
  const int Do = st.z * Di, Ho = st.y * Hi, Wo = st.x * Wi;
  const size_t ovox = (size_t)Do * Ho * Wo, ivox = (size_t)Di * Hi * Wi;
  const int ogroups = Cout / 8;
  const size_t total = (size_t)B * ogroups * ovox;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t v = i % ovox;
  const int og = (int)((i / ovox) % ogroups);
  const int b = (int)(i / ((size_t)ogroups * ovox));
  const int xo = (int)(v % Wo), yo = (int)((v / Wo) % Ho), zo = (int)(v / ((size_t)Wo * Ho));
  const int nph = st.z * st.y * st.x;
  const int ph = ((zo % st.z) * st.y + (yo % st.y)) * st.x + (xo % st.x);
  const size_t iv = ((size_t)(zo / st.z) * Hi + (yo / st.y)) * Wi + (xo / st.x);
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  for (int g = 0; g < Cin / 8; ++g) {
    float f[8];
    uint4 raw = __ldg(in + ((size_t)b * in_groups_total + in_group_off + g) * ivox + iv);
    if (xf.scale && g >= xf.ident_groups) {
      float a[8], sh[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] = __ldg(xf.scale + (size_t)b * xf.channels + g * 8 + e);
        sh[e] = __ldg(xf.shift + (size_t)b * xf.channels + g * 8 + e);
      }
      raw = xform8(raw, a, sh, xf.slope);
    }
    unpack8(raw, f);
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int o = 0; o < 8; ++o)
        acc[o] = fmaf(f[e], __ldg(w + ((size_t)(g * 8 + e) * Cout + og * 8 + o) * nph + ph), acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] += bias[og * 8 + o];
  out[((size_t)b * out_groups_total + out_group_off + og) * ovox + v] = pack8(acc);
}

// ------------------------------------------------------------------------------------------ head (+ accumulate)
// 1x1x1 segmentation head fused with `prediction *= gaussian; predicted_logits[sl] += prediction`
// (predict_from_raw_data.py:543,609-613).  HBM-bound: reads CIN fp16 and read-modify-writes C fp32 per voxel.
// One thread = one voxel; the RMW runs in batches of 8 classes (8 independent loads in flight per thread);
// weights [C][CIN] are broadcast from shared memory as float4.
constexpr int HEAD_MAX_CIN = 64;
constexpr int HEAD_MAX_C = 128;
template <int CIN>
__global__ void __launch_bounds__(256)
head_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int b, int D, int H, int W,
            const float* __restrict__ w, const float* __restrict__ bias, int C, float* __restrict__ logits_b,
            const FwdCall* __restrict__ call, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
            float slope) {
  extern __shared__ __align__(16) float sw[];  // [C][CIN] then [C] bias
  float* sb = sw + C * CIN;
  float* ssc = sb + C;        // [CIN] input scale / shift of batch item b (fused normalisation)
  float* ssh = ssc + CIN;
  for (int i = threadIdx.x; i < CIN * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
  if (in_scale)
    for (int i = threadIdx.x; i < CIN; i += blockDim.x) { ssc[i] = in_scale[(size_t)b * CIN + i]; ssh[i] = in_shift[(size_t)b * CIN + i]; }
  __syncthreads();
  float* __restrict__ acc = nullptr;
  const float* __restrict__ g = nullptr;
  int o0 = 0, o1 = 0, o2 = 0, d1 = 0, d2 = 0;
  size_t vol_voxels = 0;
  if (!logits_b) {
    if (b >= call->n_valid) return;
    acc = call->acc; g = call->gaussian;
    o0 = call->origins[b][0]; o1 = call->origins[b][1]; o2 = call->origins[b][2];
    d1 = call->d1; d2 = call->d2;
    vol_voxels = (size_t)call->d0 * d1 * d2;
  }
  const size_t vox = (size_t)D * H * W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += stride) {
    float x[CIN];
#pragma unroll
    for (int gi = 0; gi < CIN / 8; ++gi) {
      float f[8];
      unpack8(__ldg(in + ((size_t)b * in_groups_total + in_group_off + gi) * vox + v), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) x[gi * 8 + e] = f[e];
    }
    if (in_scale) {
      // the source is the RAW output of the last conv: apply its InstanceNorm affine + LeakyReLU and the fp16
      // rounding the standalone pass would have stored (same operations as norm_lrelu_kernel)
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float z = fmaf(x[c], ssc[c], ssh[c]);
        x[c] = __half2float(__float2half_rn(fmaxf(z, __fmul_rn(z, slope))));
      }
    }
    float* a = nullptr;
    float gw = 0.f;
    if (!logits_b) {
      const int k = (int)(v % W), j = (int)((v / W) % H), ii = (int)(v / ((size_t)W * H));
      a = acc + ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k);
      gw = __ldg(g + v);
    }
    // batches of up to 32 classes: issue every accumulator load first, hide their latency behind the dot products
    for (int c0 = 0; c0 < C; c0 += 32) {
      float old[32];
      if (!logits_b) {
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (c0 + u < C) old[u] = __ldcg(a + (size_t)(c0 + u) * vol_voxels);
      }
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if (c0 + u < C) {
          const float4* wr = reinterpret_cast<const float4*>(sw + (c0 + u) * CIN);
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int q = 0; q < CIN / 4; q += 2) {
            const float4 w4 = wr[q], w5 = wr[q + 1];
            s0 = fmaf(x[4 * q], w4.x, s0); s0 = fmaf(x[4 * q + 1], w4.y, s0);
            s0 = fmaf(x[4 * q + 2], w4.z, s0); s0 = fmaf(x[4 * q + 3], w4.w, s0);
            s1 = fmaf(x[4 * q + 4], w5.x, s1); s1 = fmaf(x[4 * q + 5], w5.y, s1);
            s1 = fmaf(x[4 * q + 6], w5.z, s1); s1 = fmaf(x[4 * q + 7], w5.w, s1);
          }
          const float sres = (s0 + s1) + sb[c0 + u];
          if (logits_b) logits_b[(size_t)(c0 + u) * vox + v] = sres;
          else __stcg(a + (size_t)(c0 + u) * vol_voxels, __fadd_rn(old[u], __fmul_rn(sres, gw)));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ head on mma.sync
// Same op as head_kernel for CIN = 32 and C <= 32 (every BOA network): the 32 x C product runs on the tensor cores
// through mma.sync.m16n8k16 (fp16 operands - the activations ARE fp16 and the weights are fp16-representable - fp32
// accumulation), which leaves the kernel with what it is bound by: the fp32 read-modify-write of the C logits planes.
// (tcgen05 is not an option here: the accumulators of a 1x1x1 conv are consumed once, and a TMEM round trip per 128
// voxels buys nothing for K = 32.)
// A warp step covers 32 consecutive voxels (two m16 tiles).  The MMA's K index is a free permutation of the input
// channels as long as both operands agree, so it is chosen such that thread (g = lane / 4, t = lane % 4) needs exactly
// the 8 channels of channel group t of its rows - one 16-byte load per row straight from the C8 tensor:
//   k-step s, k = 2t + j      <->  channel 8t + 4s + j        (j = 0, 1)
//   k-step s, k = 2t + 8 + j  <->  channel 8t + 4s + 2 + j
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// red.global.add.f32: round-to-nearest fp32 addition executed by the L2 (subnormal operands / results are flushed to
// zero - |logit * gaussian| would have to be below 1.2e-38, the Gaussian's minimum is 6e-8).
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}

// RED: the accumulation `logits[v] += pred * g` is issued as a reduction (red.global.add.f32, executed by the L2) instead
// of a load + add + store in the SM: the same IEEE fp32 addition on the same operands - every address is touched once
// per launch and launches are stream ordered, so the result is bit-identical - but no accumulator value travels to
// the SM and back, and the 8 * NT registers of loads in flight per thread are gone.
template <int NT, bool RED>  // n-tiles of 8 classes: C <= 8 * NT
__global__ void __launch_bounds__(128, RED ? 6 : 4)
head_mma_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int b, int D, int H, int W,
                const float* __restrict__ w, const float* __restrict__ bias, int C, float* __restrict__ logits_b,
                const FwdCall* __restrict__ call, const float* __restrict__ in_scale,
                const float* __restrict__ in_shift, float slope) {
  __shared__ uint2 sB[NT][2][32];  // B fragments: [n-tile][k-step][lane] = {b0 (k = 2t, 2t+1), b1 (k = 2t+8, 2t+9)}
  __shared__ float sbias[8 * NT];
  __shared__ float snorm[64];      // [32] scale, [32] shift of batch item b (fused normalisation of the input)
  // per-warp transpose buffer [class][voxel]: the MMA leaves a thread with 2 classes x 4 voxels per n-tile, the
  // accumulate wants thread = voxel so that every class plane is touched 128 contiguous bytes at a time.  Row stride
  // 36: bank = 4 * class + voxel, conflict-free for the fragment writes (class = c0 + 2t, voxel = v0 + g -> 8t + g)
  // and for the row reads (voxel = lane).
  __shared__ float sT[4][8 * NT][36];
  if (in_scale && threadIdx.x < 64)
    snorm[threadIdx.x] = threadIdx.x < 32 ? in_scale[(size_t)b * 32 + threadIdx.x] : in_shift[(size_t)b * 32 + threadIdx.x - 32];
  for (int i = threadIdx.x; i < NT * 64; i += blockDim.x) {
    const int ln = i & 31, ks = (i >> 5) & 1, nt = i >> 6;
    const int cls = nt * 8 + (ln >> 2), ch = 8 * (ln & 3) + 4 * ks;
    __half2 lo = __floats2half2_rn(0.f, 0.f), hi = lo;
    if (cls < C) {
      lo = __floats2half2_rn(w[cls * 32 + ch], w[cls * 32 + ch + 1]);
      hi = __floats2half2_rn(w[cls * 32 + ch + 2], w[cls * 32 + ch + 3]);
    }
    sB[nt][ks][ln] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  for (int i = threadIdx.x; i < 8 * NT; i += blockDim.x) sbias[i] = i < C ? bias[i] : 0.f;
  __syncthreads();
  float* __restrict__ acc = nullptr;
  const float* __restrict__ gauss = nullptr;
  int o0 = 0, o1 = 0, o2 = 0, d1 = 0, d2 = 0;
  size_t vol_voxels = 0;
  if (!logits_b) {
    if (b >= call->n_valid) return;
    acc = call->acc; gauss = call->gaussian;
    o0 = call->origins[b][0]; o1 = call->origins[b][1]; o2 = call->origins[b][2];
    d1 = call->d1; d2 = call->d2;
    vol_voxels = (size_t)call->d0 * d1 * d2;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int vox = D * H * W;
  const float* sc = snorm + 8 * t;  // scale / shift of this thread's 8 channels (broadcast reads)
  const float* sh = snorm + 32 + 8 * t;
  float (*tb)[36] = sT[warp];
  const uint4* __restrict__ src = in + ((size_t)b * in_groups_total + in_group_off + t) * vox;
  const size_t cstride = logits_b ? (size_t)vox : vol_voxels;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int v0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32; v0 < vox; v0 += warps * 32) {
    // ---- this thread as an accumulate lane: voxel v0 + lane, every class.  Issue its loads first.
    const int v = v0 + lane;
    const bool ok = v < vox;
    float* dst = nullptr;
    float gw = 0.f;
    float old[RED ? 1 : 8 * NT];
    if (ok) {
      if (logits_b) {
        dst = logits_b + v;
      } else {
        const int k = v % W, j = (v / W) % H, ii = v / (W * H);
        dst = acc + ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k);
        gw = __ldg(gauss + v);
        if constexpr (!RED) {
#pragma unroll
          for (int c = 0; c < 8 * NT; ++c)
            if (c < C) old[c] = __ldcg(dst + (size_t)c * cstride);
        }
      }
    }
    // ---- this thread as an MMA lane: rows v0 + g + 8 i, i = 0..3 (m-tile i / 2, upper half i & 1), channel group t
    uint32_t a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int vr = v0 + g + 8 * i;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (vr < vox) raw = __ldg(src + vr);
      if (in_scale) {
        // RAW output of the last conv: its InstanceNorm affine + LeakyReLU and the fp16 rounding of the standalone
        // pass (same operations as norm_lrelu_kernel)
        float a8[8], s8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { a8[e] = sc[e]; s8[e] = sh[e]; }
        raw = xform8(raw, a8, s8, slope);
      }
      a[i][0] = raw.x; a[i][1] = raw.y; a[i][2] = raw.z; a[i][3] = raw.w;
    }
    __syncwarp();  // the previous step's reads of the transpose buffer are done
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint2 bf = sB[nt][ks][lane];
          mma_m16n8k16(d, a[2 * m][2 * ks], a[2 * m + 1][2 * ks], a[2 * m][2 * ks + 1], a[2 * m + 1][2 * ks + 1], bf.x,
                       bf.y);
        }
        // d[0], d[1]: row 16 m + g, classes nt*8 + 2t, +1;  d[2], d[3]: row 16 m + g + 8
        tb[nt * 8 + 2 * t][16 * m + g] = d[0];
        tb[nt * 8 + 2 * t + 1][16 * m + g] = d[1];
        tb[nt * 8 + 2 * t][16 * m + g + 8] = d[2];
        tb[nt * 8 + 2 * t + 1][16 * m + g + 8] = d[3];
      }
    __syncwarp();
    if (ok) {
#pragma unroll
      for (int c = 0; c < 8 * NT; ++c) {
        if (c < C) {
          const float sres = tb[c][lane] + sbias[c];
          if (logits_b) dst[(size_t)c * cstride] = sres;
          else if constexpr (RED) red_add_f32(dst + (size_t)c * cstride, __fmul_rn(sres, gw));
          else __stcg(dst + (size_t)c * cstride, __fadd_rn(old[c], __fmul_rn(sres, gw)));
        }
      }
    }
  }
}

// ================================================================================================ launchers
int launch_pack_patches_nb9(const float* d_patches, int n, int p0, int p1, int p2, __half* d_out, cudaStream_t s) {
  const size_t total = (size_t)n * p0 * p1 * p2;
  BOA_CARVEOUT_ONCE(extract_patches_nb9_kernel);
  extract_patches_nb9_kernel<<<grid_for(total, 256), 256, 0, s>>>(nullptr, d_patches, n, p0, p1, p2,
                                                                  reinterpret_cast<uint4*>(d_out));
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_extract_patches(const FwdCall* d_call, int B, int p0, int p1, int p2, __half* d_out, int mode,
                           cudaStream_t s) {
  if (mode == 2) {
    const size_t total = (size_t)B * p0 * p1 * p2;
    BOA_CARVEOUT_ONCE(extract_patches_nb9_kernel);
    extract_patches_nb9_kernel<<<grid_for(total, 256, thin_passes() ? 2 : 8), 256, 0, s>>>(d_call, nullptr, B, p0, p1, p2,
                                                                    reinterpret_cast<uint4*>(d_out));
  } else if (mode == 1) {
    const size_t total = (size_t)B * p0 * p1 * p2;
    extract_patches_plain_kernel<<<grid_for(total, 256), 256, 0, s>>>(d_call, B, p0, p1, p2, d_out);
  } else {
    const size_t total = (size_t)B * 2 * p0 * p1 * p2;
    extract_patches_kernel<<<grid_for(total, 256), 256, 0, s>>>(d_call, B, p0, p1, p2, reinterpret_cast<uint4*>(d_out));
  }
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_pack_patches_plain(const float* d_patches, size_t total, __half* d_out, cudaStream_t s) {
  pack_patches_plain_kernel<<<grid_for(total, 256), 256, 0, s>>>(d_patches, total, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

static bool head_on_mma() {  // BOA_B200_HEAD_FMA=1 selects the FP32-FMA head (cross-check of the mma.sync head)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOA_B200_HEAD_FMA");
    v = (e && atoi(e) != 0) ? 0 : 1;
  }
  return v == 1;
}

bool conv_first_supported(int Cout) { return Cout == 32 || Cout == 16 || Cout == 8; }

int launch_conv_first(const __half* d_in, int B, const float* d_w, const float* d_bias, int Cout, __half* d_raw_out,
                      int D, int H, int W, double* d_stats, cudaStream_t s) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)(((D + 3) / 4) * B));
  uint4* out = reinterpret_cast<uint4*>(d_raw_out);
  if (Cout == 32) conv_first_kernel<32><<<grid, 256, 0, s>>>(d_in, d_w, d_bias, out, d_stats, D, H, W);
  else if (Cout == 16) conv_first_kernel<16><<<grid, 256, 0, s>>>(d_in, d_w, d_bias, out, d_stats, D, H, W);
  else if (Cout == 8) conv_first_kernel<8><<<grid, 256, 0, s>>>(d_in, d_w, d_bias, out, d_stats, D, H, W);
  else {
    set_error("conv_first: Cout=%d not instantiated", Cout);
    return BOA_ERR_UNSUPPORTED;
  }
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_pack_patches(const float* d_patches, int n, int cin, int p0, int p1, int p2, __half* d_out, int groups,
                        cudaStream_t s) {
  const size_t pv = (size_t)p0 * p1 * p2;
  pack_patches_kernel<<<grid_for((size_t)n * groups * pv, 256), 256, 0, s>>>(d_patches, n, cin, pv, groups,
                                                                            reinterpret_cast<uint4*>(d_out));
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_stats_finalize(const double* d_stats, const float* d_gamma, const float* d_beta, int B, int C,
                          double n_vox, float eps, float* d_scale, float* d_shift, cudaStream_t s, float* d_scale2,
                          float* d_shift2, int stride2, int off2) {
  BOA_CARVEOUT_ONCE(stats_finalize_kernel);
  stats_finalize_kernel<<<(B * C + 127) / 128, 128, 0, s>>>(d_stats, d_gamma, d_beta, B, C, n_vox, eps, d_scale,
                                                            d_shift, d_scale2, d_shift2, stride2, off2);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_norm_lrelu(const __half* d_raw, int B, int groups, int D, int H, int W, const float* d_scale,
                      const float* d_shift, float slope, const ActView& dst, __half* d_s2d, cudaStream_t s) {
  const size_t vox = (size_t)D * H * W;
  const int planes = B * groups;
  if (thin_passes()) {
    BOA_CARVEOUT_ONCE(norm_lrelu_thin_kernel);
    const size_t items = (size_t)planes * ((vox + 256 * THIN_U - 1) / (256 * THIN_U));
    const int grid = (int)std::min<size_t>(items, (size_t)sm_count());
    norm_lrelu_thin_kernel<<<grid, 256, 0, s>>>(
        reinterpret_cast<const uint4*>(d_raw), planes, groups, D, H, W, d_scale, d_shift, slope,
        reinterpret_cast<uint4*>(dst.base), dst.groups_total, dst.group_off, reinterpret_cast<uint4*>(d_s2d));
    BOA_CHECK_LAUNCH();
    return BOA_OK;
  }
  // ~8 resident blocks per SM in total, split over the (b, group) planes; 4 vectors per thread per iteration
  int bx = (int)((vox + 4 * 256 - 1) / (4 * 256));
  const int cap = (sm_count() * 8 + planes - 1) / planes;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  BOA_CARVEOUT_ONCE(norm_lrelu_kernel);
  norm_lrelu_kernel<<<dim3(bx, planes), 256, 0, s>>>(
      reinterpret_cast<const uint4*>(d_raw), groups, D, H, W, d_scale, d_shift, slope,
      reinterpret_cast<uint4*>(dst.base), dst.groups_total, dst.group_off, reinterpret_cast<uint4*>(d_s2d));
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_conv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int cin_w, int Cout,
                     const int* ks, const int* stride, const ActView& out, int Do, int Ho, int Wo, double* d_stats,
                     cudaStream_t s, const InXform& xf) {
  const size_t total = (size_t)B * (Cout / 8) * Do * Ho * Wo;
  conv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
      reinterpret_cast<const uint4*>(src.base), src.groups_total, src.group_off, src.D, src.H, src.W, d_w, d_bias,
      cin_w, Cout, Int3{ks[0], ks[1], ks[2]}, Int3{stride[0], stride[1], stride[2]},
      reinterpret_cast<uint4*>(out.base), out.groups_total, out.group_off, Do, Ho, Wo, d_stats, B, xf);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_tconv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int Cin, int Cout,
                      const int* stride, const ActView& dst, cudaStream_t s, const InXform& xf) {
  const size_t total = (size_t)B * (Cout / 8) * dst.voxels();
  tconv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
      reinterpret_cast<const uint4*>(src.base), src.groups_total, src.group_off, src.D, src.H, src.W, d_w, d_bias, Cin,
      Cout, Int3{stride[0], stride[1], stride[2]}, reinterpret_cast<uint4*>(dst.base), dst.groups_total,
      dst.group_off, B, xf);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_head(const ActView& src, int b, const float* d_w, const float* d_bias, int Cin, int C, float* d_logits_b,
                const FwdCall* d_call, const float* d_in_scale, const float* d_in_shift, float slope, cudaStream_t s) {
  if (Cin > HEAD_MAX_CIN || Cin % 8 || C > HEAD_MAX_C) {
    set_error("head: Cin=%d C=%d exceed the head kernel limits (%d, %d)", Cin, C, HEAD_MAX_CIN, HEAD_MAX_C);
    return BOA_ERR_UNSUPPORTED;
  }
  const size_t vox = src.voxels();
  if (Cin == 32 && C <= 32 && head_on_mma()) {
    const int nt = (C + 7) / 8;
    // BOA_B200_HEAD_RMW=1: load + add + store in the SM (four blocks of 128 threads per SM, up to 8 * NT + 5 loads in
    // flight per thread) instead of the L2 reduction (six blocks per SM)
    static const bool red = !(getenv("BOA_B200_HEAD_RMW") && atoi(getenv("BOA_B200_HEAD_RMW")) != 0);
    const int grid = (int)std::min<size_t>((vox + 127) / 128, (size_t)sm_count() * (red ? 6 : 4));
    const uint4* in4 = reinterpret_cast<const uint4*>(src.base);
#define BOA_HEAD_MMA(NT_)                                                                                            \
  if (red) {                                                                                                           \
    BOA_CARVEOUT_ONCE((head_mma_kernel<NT_, true>));                                                                   \
    head_mma_kernel<NT_, true><<<grid, 128, 0, s>>>(in4, src.groups_total, src.group_off, b, src.D, src.H, src.W, d_w, \
                                                    d_bias, C, d_logits_b, d_call, d_in_scale, d_in_shift, slope);     \
  } else {                                                                                                             \
    BOA_CARVEOUT_ONCE((head_mma_kernel<NT_, false>));                                                                  \
    head_mma_kernel<NT_, false><<<grid, 128, 0, s>>>(in4, src.groups_total, src.group_off, b, src.D, src.H, src.W, d_w, \
                                                     d_bias, C, d_logits_b, d_call, d_in_scale, d_in_shift, slope);    \
  }
    switch (nt) {
      case 1: BOA_HEAD_MMA(1); break;
      case 2: BOA_HEAD_MMA(2); break;
      case 3: BOA_HEAD_MMA(3); break;
      default: BOA_HEAD_MMA(4); break;
    }
#undef BOA_HEAD_MMA
    BOA_CHECK_LAUNCH();
    return BOA_OK;
  }
  const size_t smem = ((size_t)C * Cin + C + 2 * Cin) * sizeof(float);
  const uint4* in = reinterpret_cast<const uint4*>(src.base);
  // thin mode: 128 threads x <= 128 registers per SM (see norm_lrelu_thin_kernel), one persistent block per SM
  const bool thin = thin_passes();
  const int threads = thin ? 128 : 256;
  const int grid = thin ? (int)std::min<size_t>((vox + 127) / 128, (size_t)sm_count()) : grid_for(vox, 256, 8);
#define BOA_HEAD(CIN_)                                                                                              \
  BOA_CARVEOUT_ONCE(head_kernel<CIN_>);                                                                               \
  head_kernel<CIN_><<<grid, threads, smem, s>>>(in, src.groups_total, src.group_off, b, src.D, src.H, src.W, d_w, d_bias, \
                                            C, d_logits_b, d_call, d_in_scale, d_in_shift, slope)
  switch (Cin) {
    case 8: BOA_HEAD(8); break;
    case 16: BOA_HEAD(16); break;
    case 24: BOA_HEAD(24); break;
    case 32: BOA_HEAD(32); break;
    case 40: BOA_HEAD(40); break;
    case 48: BOA_HEAD(48); break;
    case 56: BOA_HEAD(56); break;
    case 64: BOA_HEAD(64); break;
    default:
      set_error("head: Cin=%d not instantiated", Cin);
      return BOA_ERR_UNSUPPORTED;
  }
#undef BOA_HEAD
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

}  // namespace boa
