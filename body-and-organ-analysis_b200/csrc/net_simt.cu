// Elementwise network passes (patch extraction, InstanceNorm finalisation, normalise + LeakyReLU) and the SIMT
// reference kernels (direct conv / transposed conv / head).  The SIMT convs exist as an on-device cross-check for
// the tcgen05 kernels and as the implementation of ops that have not moved to tensor cores yet; they are selected
// explicitly, never as a silent fallback.
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
__global__ void stats_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int B, int C, double n_vox, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C;
  const double mean = stats[2 * i] / n_vox;
  double var = stats[2 * i + 1] / n_vox - mean * mean;
  if (var < 0) var = 0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma[c];
  scale[i] = (float)(g * rstd);
  shift[i] = (float)((double)beta[c] - mean * g * rstd);
}

// ------------------------------------------------------------------------------------------ normalise + LeakyReLU
__global__ void __launch_bounds__(256)
norm_lrelu_kernel(const uint4* __restrict__ raw, int B, int groups, int D, int H, int W,
                  const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                  uint4* __restrict__ dst, int dst_groups_total, int dst_group_off, uint4* __restrict__ s2d) {
  const size_t vox = (size_t)D * H * W;
  const size_t total = (size_t)B * groups * vox;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t v = i % vox;
    const int g = (int)((i / vox) % groups);
    const int b = (int)(i / ((size_t)groups * vox));
    float f[8];
    unpack8(__ldg(raw + i), f);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + ((size_t)b * groups + g) * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(scale + ((size_t)b * groups + g) * 8) + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + ((size_t)b * groups + g) * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(shift + ((size_t)b * groups + g) * 8) + 1);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float z = __fadd_rn(__fmul_rn(f[e], a[e]), s[e]);
      f[e] = z > 0.f ? z : __fmul_rn(z, slope);
    }
    const uint4 o = pack8(f);
    if (dst) dst[((size_t)b * dst_groups_total + dst_group_off + g) * vox + v] = o;
    if (s2d) {
      const int x = (int)(v % W), y = (int)((v / W) % H), z = (int)(v / ((size_t)W * H));
      const int ph = ((z & 1) * 2 + (y & 1)) * 2 + (x & 1);
      const size_t hv = ((size_t)(z >> 1) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
      s2d[(((size_t)b * 8 + ph) * groups + g) * (vox >> 3) + hv] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------ SIMT direct conv
// One thread = one output voxel x 8 output channels.  Accumulates InstanceNorm statistics with fp64 atomics.
struct Int3 { int z, y, x; };

__global__ void __launch_bounds__(128)
conv_simt_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int Di, int Hi, int Wi,
                 const float* __restrict__ w, const float* __restrict__ bias, int cin_w, int Cout, Int3 ks,
                 Int3 stride, uint4* __restrict__ out, int Do, int Ho, int Wo, double* __restrict__ stats, int B) {
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
      unpack8(__ldg(in + ((size_t)b * in_groups_total + in_group_off + g) * ivox + iv), f);
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
  if (active) out[ii] = pack8(acc);
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
                  uint4* __restrict__ out, int out_groups_total, int out_group_off, int B) {
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
    unpack8(__ldg(in + ((size_t)b * in_groups_total + in_group_off + g) * ivox + iv), f);
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
// (predict_from_raw_data.py:543,609-613).  One thread = one voxel; weights [C][Cin] in shared memory, read as float4
// broadcasts.  HBM-bound: reads Cin fp16 + RMW of C fp32 per voxel.
constexpr int HEAD_MAX_CIN = 64;
constexpr int HEAD_MAX_C = 128;
__global__ void __launch_bounds__(256)
head_kernel(const uint4* __restrict__ in, int in_groups_total, int in_group_off, int b, int D, int H, int W,
            const float* __restrict__ w, const float* __restrict__ bias, int Cin, int C,
            float* __restrict__ logits_b, const FwdCall* __restrict__ call) {
  extern __shared__ float sw[];  // [C][Cin] then [C] bias
  float* sb = sw + C * Cin;
  for (int i = threadIdx.x; i < Cin * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
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
  const int ngroups = Cin / 8;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += stride) {
    float x[HEAD_MAX_CIN];
#pragma unroll
    for (int gi = 0; gi < HEAD_MAX_CIN / 8; ++gi) {
      if (gi < ngroups) {
        float f[8];
        unpack8(__ldg(in + ((size_t)b * in_groups_total + in_group_off + gi) * vox + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[gi * 8 + e] = f[e];
      }
    }
    size_t av = 0;
    float gw = 0.f;
    if (!logits_b) {
      const int k = (int)(v % W), j = (int)((v / W) % H), ii = (int)(v / ((size_t)W * H));
      av = ((size_t)(o0 + ii) * d1 + (o1 + j)) * d2 + (o2 + k);
      gw = __ldg(g + v);
    }
    for (int c = 0; c < C; ++c) {
      const float4* wr = reinterpret_cast<const float4*>(sw + c * Cin);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < HEAD_MAX_CIN / 4; ++q) {
        if (q * 4 < Cin) {
          const float4 w4 = wr[q];
          s = fmaf(x[4 * q], w4.x, s);
          s = fmaf(x[4 * q + 1], w4.y, s);
          s = fmaf(x[4 * q + 2], w4.z, s);
          s = fmaf(x[4 * q + 3], w4.w, s);
        }
      }
      s += sb[c];
      if (logits_b) {
        logits_b[(size_t)c * vox + v] = s;
      } else {
        float* a = acc + (size_t)c * vol_voxels + av;
        *a = __fadd_rn(*a, __fmul_rn(s, gw));
      }
    }
  }
}

// ================================================================================================ launchers
int launch_extract_patches(const FwdCall* d_call, int B, int p0, int p1, int p2, __half* d_out, cudaStream_t s) {
  const size_t total = (size_t)B * 2 * p0 * p1 * p2;
  extract_patches_kernel<<<grid_for(total, 256), 256, 0, s>>>(d_call, B, p0, p1, p2, reinterpret_cast<uint4*>(d_out));
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
                          double n_vox, float eps, float* d_scale, float* d_shift, cudaStream_t s) {
  stats_finalize_kernel<<<(B * C + 127) / 128, 128, 0, s>>>(d_stats, d_gamma, d_beta, B, C, n_vox, eps, d_scale,
                                                            d_shift);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_norm_lrelu(const __half* d_raw, int B, int groups, int D, int H, int W, const float* d_scale,
                      const float* d_shift, float slope, const ActView& dst, __half* d_s2d, cudaStream_t s) {
  const size_t total = (size_t)B * groups * D * H * W;
  norm_lrelu_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      reinterpret_cast<const uint4*>(d_raw), B, groups, D, H, W, d_scale, d_shift, slope,
      reinterpret_cast<uint4*>(dst.base), dst.groups_total, dst.group_off, reinterpret_cast<uint4*>(d_s2d));
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_conv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int cin_w, int Cout,
                     const int* ks, const int* stride, __half* d_raw_out, int Do, int Ho, int Wo, double* d_stats,
                     cudaStream_t s) {
  const size_t total = (size_t)B * (Cout / 8) * Do * Ho * Wo;
  conv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
      reinterpret_cast<const uint4*>(src.base), src.groups_total, src.group_off, src.D, src.H, src.W, d_w, d_bias,
      cin_w, Cout, Int3{ks[0], ks[1], ks[2]}, Int3{stride[0], stride[1], stride[2]},
      reinterpret_cast<uint4*>(d_raw_out), Do, Ho, Wo, d_stats, B);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_tconv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int Cin, int Cout,
                      const int* stride, const ActView& dst, cudaStream_t s) {
  const size_t total = (size_t)B * (Cout / 8) * dst.voxels();
  tconv_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
      reinterpret_cast<const uint4*>(src.base), src.groups_total, src.group_off, src.D, src.H, src.W, d_w, d_bias, Cin,
      Cout, Int3{stride[0], stride[1], stride[2]}, reinterpret_cast<uint4*>(dst.base), dst.groups_total,
      dst.group_off, B);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

int launch_head(const ActView& src, int b, const float* d_w, const float* d_bias, int Cin, int C, float* d_logits_b,
                const FwdCall* d_call, cudaStream_t s) {
  if (Cin > HEAD_MAX_CIN || Cin % 8 || C > HEAD_MAX_C) {
    set_error("head: Cin=%d C=%d exceed the head kernel limits (%d, %d)", Cin, C, HEAD_MAX_CIN, HEAD_MAX_C);
    return BOA_ERR_UNSUPPORTED;
  }
  const size_t vox = src.voxels();
  const size_t smem = ((size_t)C * Cin + C) * sizeof(float);
  head_kernel<<<grid_for(vox, 256, 4), 256, smem, s>>>(reinterpret_cast<const uint4*>(src.base), src.groups_total,
                                                       src.group_off, b, src.D, src.H, src.W, d_w, d_bias, Cin, C,
                                                       d_logits_b, d_call);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

}  // namespace boa
