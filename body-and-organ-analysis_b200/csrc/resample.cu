// Slice-thickness resampling for the body-composition networks (resample_only_thickness,
// _external/totalsegmentator/nnunet.py:457-475 and :685-687; change_spacing -> scipy.ndimage.zoom,
// _external/totalsegmentator/resampling.py:24-56,129-222).
//
// With zoom factors (z, 1, 1) scipy's order-3 zoom (mode="nearest") reduces to a 1-D interpolating cubic B-spline
// along z: the volume is edge-padded by 12 samples, pre-filtered with the single pole z1 = sqrt(3) - 2 (mirror
// initialisation on the padded line), and evaluated at o * (Zin - 1) / (Zout - 1) + 12 with the four cubic weights;
// at the integer in-plane coordinates the spline reproduces the samples.  Everything runs in fp64 like scipy; the
// result is truncated toward zero as `new_data.astype(np.int32)` does (resampling.py:213-214).
// Layout: [z][y][x]; one thread per (y, x) column, neighbouring threads read neighbouring x => coalesced.
#include "common.cuh"

namespace boa {

constexpr int NPAD = 12;

// The reference keeps the resampled CT as int32 (resampling.py:213-214); here it is int16, so a spline overshoot beyond
// the int16 range saturates instead of wrapping around in sign.
__device__ __forceinline__ int16_t trunc_sat_i16(double v) {
  const int t = (int)v;
  return (int16_t)(t < -32768 ? -32768 : (t > 32767 ? 32767 : t));
}

template <typename T>
__global__ void __launch_bounds__(256)
spline_prefilter_z_kernel(const T* __restrict__ in, int z_in, size_t plane, double* __restrict__ c) {
  const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= plane) return;
  const int n = z_in + 2 * NPAD;
  const double z1 = sqrt(3.0) - 2.0;
  const double gain = (1.0 - z1) * (1.0 - 1.0 / z1);
  auto sample = [&](int i) -> double {  // gain-scaled, edge-padded input
    int k = i - NPAD;
    k = k < 0 ? 0 : (k >= z_in ? z_in - 1 : k);
    return (double)in[(size_t)k * plane + col] * gain;
  };
  // causal initialisation, mirror boundary: c0 = sum_i z1^i * p[i]  (terms beyond i ~ 64 are below one ulp)
  double c0 = sample(0);
  {
    double zi = z1;
    const int horizon = n - 1 < 80 ? n - 1 : 80;
    for (int i = 1; i < horizon; ++i) {
      c0 += zi * sample(i);
      zi *= z1;
    }
  }
  double prev = c0;
  c[col] = prev;
  for (int i = 1; i < n; ++i) {
    prev = sample(i) + z1 * prev;
    c[(size_t)i * plane + col] = prev;
  }
  // anticausal initialisation + recursion, in place
  const double cn2 = c[(size_t)(n - 2) * plane + col];
  double next = (z1 / (z1 * z1 - 1.0)) * (z1 * cn2 + prev);
  c[(size_t)(n - 1) * plane + col] = next;
  for (int i = n - 2; i >= 0; --i) {
    next = z1 * (next - c[(size_t)i * plane + col]);
    c[(size_t)i * plane + col] = next;
  }
}

__global__ void __launch_bounds__(256)
spline_eval_z_kernel(const double* __restrict__ c, int z_in, size_t plane, int z_out, int16_t* __restrict__ out) {
  const size_t total = (size_t)z_out * plane;
  const double zoom = z_out > 1 ? (double)(z_in - 1) / (double)(z_out - 1) : 1.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i / plane);
    const size_t col = i % plane;
    const double cc = (double)o * zoom + (double)NPAD;
    const double fl = floor(cc);
    const double y = cc - fl, z = 1.0 - y;
    const double w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
    const double w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
    const double w0 = z * z * z / 6.0;
    const double w3 = 1.0 - w0 - w1 - w2;
    const size_t s = (size_t)((int)fl - 1);
    const double v = w0 * c[s * plane + col] + w1 * c[(s + 1) * plane + col] + w2 * c[(s + 2) * plane + col] +
                     w3 * c[(s + 3) * plane + col];
    out[i] = trunc_sat_i16(v);  // truncation toward zero, as astype(np.int32); saturated instead of wrapped
  }
}

__global__ void __launch_bounds__(256)
nearest_z_kernel(const uint8_t* __restrict__ in, int z_in, size_t plane, int z_out, uint8_t* __restrict__ out) {
  const size_t total = (size_t)z_out * plane;
  const double zoom = z_out > 1 ? (double)(z_in - 1) / (double)(z_out - 1) : 1.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i / plane);
    const size_t col = i % plane;
    int k = (int)floor((double)o * zoom + 0.5);  // scipy order-0: round half up of the input coordinate
    k = k < 0 ? 0 : (k >= z_in ? z_in - 1 : k);
    out[i] = in[(size_t)k * plane + col];
  }
}


// ------------------------------------------------------------------------------------------ 3-D resampling
// change_spacing(img, [1.5]*3, order=3) (_external/totalsegmentator/nnunet.py:466-470 -> resampling.py:129-222):
// scipy's order-3 zoom is a tensor product, so it is applied as three 1-D passes (prefilter + evaluation along one
// axis each, fp64 intermediates, the same edge padding / mirror initialisation per axis as above); only the last pass
// truncates to integers.  A volume is addressed as [outer][n][inner] around the axis being resampled; one thread owns
// one line (outer, inner), neighbouring threads neighbouring `inner` => coalesced for the z and y passes (the x pass,
// inner = 1, walks contiguous lines and runs last, on the smallest intermediate).
template <typename T>
__device__ __forceinline__ double load_as_double(const T* p, size_t i) { return (double)p[i]; }

template <typename T>
__global__ void __launch_bounds__(256)
spline_prefilter_axis_kernel(const T* __restrict__ in, size_t outer, int n_in, size_t inner, double* __restrict__ c) {
  const size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= outer * inner) return;
  const size_t o = line / inner, i0 = line - o * inner;
  const T* src = in + o * (size_t)n_in * inner + i0;
  const int n = n_in + 2 * NPAD;
  double* dst = c + o * (size_t)n * inner + i0;
  const double z1 = sqrt(3.0) - 2.0;
  const double gain = (1.0 - z1) * (1.0 - 1.0 / z1);
  auto sample = [&](int i) -> double {
    int k = i - NPAD;
    k = k < 0 ? 0 : (k >= n_in ? n_in - 1 : k);
    return load_as_double(src, (size_t)k * inner) * gain;
  };
  double c0 = sample(0);
  {
    double zi = z1;
    const int horizon = n - 1 < 80 ? n - 1 : 80;
    for (int i = 1; i < horizon; ++i) {
      c0 += zi * sample(i);
      zi *= z1;
    }
  }
  double prev = c0;
  dst[0] = prev;
  for (int i = 1; i < n; ++i) {
    prev = sample(i) + z1 * prev;
    dst[(size_t)i * inner] = prev;
  }
  const double cn2 = dst[(size_t)(n - 2) * inner];
  double next = (z1 / (z1 * z1 - 1.0)) * (z1 * cn2 + prev);
  dst[(size_t)(n - 1) * inner] = next;
  for (int i = n - 2; i >= 0; --i) {
    next = z1 * (next - dst[(size_t)i * inner]);
    dst[(size_t)i * inner] = next;
  }
}

// out_mode 0: fp64 (intermediate pass);  1: int16, truncated toward zero;  2: int32, truncated toward zero;
//          3: fp32.
// grid_mode 0: scipy.ndimage.zoom's default coordinates  in = out * (n_in - 1) / (n_out - 1);
//           1: grid_mode=True (what skimage.transform.resize passes)  in = (out + 0.5) * n_in / n_out - 0.5.
__global__ void __launch_bounds__(256)
spline_eval_axis_kernel(const double* __restrict__ c, size_t outer, int n_in, size_t inner, int n_out, void* __restrict__ out,
                        int out_mode, int grid_mode) {
  const size_t total = outer * (size_t)n_out * inner;
  const double zoom = grid_mode ? (double)n_in / (double)n_out
                                : (n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 1.0);
  const int n = n_in + 2 * NPAD;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t i0 = i % inner;
    const int k = (int)((i / inner) % (size_t)n_out);
    const size_t o = i / (inner * (size_t)n_out);
    const double cc = (grid_mode ? ((double)k + 0.5) * zoom - 0.5 : (double)k * zoom) + (double)NPAD;
    const double fl = floor(cc);
    const double y = cc - fl, z = 1.0 - y;
    const double w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
    const double w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
    const double w0 = z * z * z / 6.0;
    const double w3 = 1.0 - w0 - w1 - w2;
    const double* line = c + o * (size_t)n * inner + i0;
    const size_t s = (size_t)((int)fl - 1);
    const double v = w0 * line[s * inner] + w1 * line[(s + 1) * inner] + w2 * line[(s + 2) * inner] +
                     w3 * line[(s + 3) * inner];
    if (out_mode == 0) static_cast<double*>(out)[i] = v;
    else if (out_mode == 1) static_cast<int16_t*>(out)[i] = trunc_sat_i16(v);
    else if (out_mode == 2) static_cast<int32_t*>(out)[i] = (int32_t)v;
    else static_cast<float*>(out)[i] = (float)v;
  }
}

// Order-0 zoom of a label map in 3-D (change_spacing(..., target_shape=original, order=0), nnunet.py:685-687): every
// axis picks floor(o * (n_in - 1) / (n_out - 1) + 0.5), clamped.
__global__ void __launch_bounds__(256)
nearest_3d_kernel(const uint8_t* __restrict__ in, int zi, int yi, int xi, int zo, int yo, int xo,
                  uint8_t* __restrict__ out) {
  const size_t total = (size_t)zo * yo * xo;
  const double fz = zo > 1 ? (double)(zi - 1) / (double)(zo - 1) : 1.0;
  const double fy = yo > 1 ? (double)(yi - 1) / (double)(yo - 1) : 1.0;
  const double fx = xo > 1 ? (double)(xi - 1) / (double)(xo - 1) : 1.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x = (int)(i % xo), y = (int)((i / xo) % yo), z = (int)(i / ((size_t)xo * yo));
    int kz = (int)floor((double)z * fz + 0.5), ky = (int)floor((double)y * fy + 0.5), kx = (int)floor((double)x * fx + 0.5);
    kz = kz < 0 ? 0 : (kz >= zi ? zi - 1 : kz);
    ky = ky < 0 ? 0 : (ky >= yi ? yi - 1 : ky);
    kx = kx < 0 ? 0 : (kx >= xi ? xi - 1 : kx);
    out[i] = in[((size_t)kz * yi + ky) * xi + kx];
  }
}

}  // namespace boa

using namespace boa;

extern "C" int boa_resample_z_cubic(const void* d_in, int in_dtype, int z_in, size_t plane, int z_out,
                                    double* d_scratch, int16_t* d_out, void* stream) {
  BOA_REQUIRE(d_in && d_scratch && d_out, "boa_resample_z_cubic: null pointer");
  BOA_REQUIRE(z_in >= 2 && z_out >= 1 && plane > 0, "boa_resample_z_cubic: bad sizes (z_in=%d z_out=%d)", z_in, z_out);
  BOA_REQUIRE(in_dtype == BOA_DT_I16 || in_dtype == BOA_DT_F32, "boa_resample_z_cubic: bad dtype %d", in_dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((plane + 255) / 256);
  if (in_dtype == BOA_DT_I16)
    spline_prefilter_z_kernel<short><<<blocks, 256, 0, s>>>(static_cast<const short*>(d_in), z_in, plane, d_scratch);
  else
    spline_prefilter_z_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(d_in), z_in, plane, d_scratch);
  BOA_CHECK_LAUNCH();
  spline_eval_z_kernel<<<grid_for((size_t)z_out * plane, 256), 256, 0, s>>>(d_scratch, z_in, plane, z_out, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_resample_z_nearest_u8(const uint8_t* d_in, int z_in, size_t plane, int z_out, uint8_t* d_out,
                                         void* stream) {
  BOA_REQUIRE(d_in && d_out, "boa_resample_z_nearest_u8: null pointer");
  BOA_REQUIRE(z_in >= 1 && z_out >= 1 && plane > 0, "boa_resample_z_nearest_u8: bad sizes");
  nearest_z_kernel<<<grid_for((size_t)z_out * plane, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, z_in, plane, z_out, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_resample_axis_cubic(const void* d_in, int in_dtype, size_t outer, int n_in, size_t inner, int n_out,
                                       double* d_scratch, void* d_out, int out_mode, void* stream) {
  BOA_REQUIRE(d_in && d_scratch && d_out, "boa_resample_axis_cubic: null pointer");
  BOA_REQUIRE(n_in >= 2 && n_out >= 1 && outer > 0 && inner > 0, "boa_resample_axis_cubic: bad sizes (n_in=%d n_out=%d)",
              n_in, n_out);
  BOA_REQUIRE(in_dtype == BOA_DT_I16 || in_dtype == BOA_DT_F32 || in_dtype == BOA_DT_F64,
              "boa_resample_axis_cubic: bad input dtype %d", in_dtype);
  BOA_REQUIRE(out_mode >= 0 && out_mode <= 2, "boa_resample_axis_cubic: bad output mode %d", out_mode);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t lines = outer * inner;
  const unsigned blocks = (unsigned)((lines + 255) / 256);
  if (in_dtype == BOA_DT_I16)
    spline_prefilter_axis_kernel<short><<<blocks, 256, 0, s>>>(static_cast<const short*>(d_in), outer, n_in, inner, d_scratch);
  else if (in_dtype == BOA_DT_F32)
    spline_prefilter_axis_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(d_in), outer, n_in, inner, d_scratch);
  else
    spline_prefilter_axis_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(d_in), outer, n_in, inner, d_scratch);
  BOA_CHECK_LAUNCH();
  spline_eval_axis_kernel<<<grid_for(outer * (size_t)n_out * inner, 256), 256, 0, s>>>(d_scratch, outer, n_in, inner, n_out,
                                                                                    d_out, out_mode, 0);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// The same 1-D pass with grid_mode=True coordinates and an optional fp32 result (out_mode 3): one axis of
// skimage.transform.resize(order=3, mode="edge", anti_aliasing=False) = scipy.ndimage.zoom(..., mode="nearest",
// grid_mode=True), which nnU-Net's resample_data_or_seg calls for the image data
// (_external/nnunetv2/preprocessing/resampling/default_resampling.py:117-203).
extern "C" int boa_resample_axis_cubic_grid(const void* d_in, int in_dtype, size_t outer, int n_in, size_t inner,
                                            int n_out, double* d_scratch, void* d_out, int out_mode, void* stream) {
  BOA_REQUIRE(d_in && d_scratch && d_out, "boa_resample_axis_cubic_grid: null pointer");
  BOA_REQUIRE(n_in >= 2 && n_out >= 1 && outer > 0 && inner > 0, "boa_resample_axis_cubic_grid: bad sizes");
  BOA_REQUIRE(in_dtype == BOA_DT_F32 || in_dtype == BOA_DT_F64, "boa_resample_axis_cubic_grid: bad input dtype %d", in_dtype);
  BOA_REQUIRE(out_mode == 0 || out_mode == 3, "boa_resample_axis_cubic_grid: bad output mode %d", out_mode);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((outer * inner + 255) / 256);
  if (in_dtype == BOA_DT_F32)
    spline_prefilter_axis_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(d_in), outer, n_in, inner, d_scratch);
  else
    spline_prefilter_axis_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(d_in), outer, n_in, inner, d_scratch);
  BOA_CHECK_LAUNCH();
  spline_eval_axis_kernel<<<grid_for(outer * (size_t)n_out * inner, 256), 256, 0, s>>>(d_scratch, outer, n_in, inner, n_out,
                                                                                    d_out, out_mode, 1);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

// skimage.transform.resize(..., clip=True) clips its result to the value range of ITS input - one 2-D slice when nnU-Net
// resamples slice by slice (separate z), the whole volume otherwise.  d_ref: the input of the resize [n_slices][ref_n],
// d_data: its output [n_slices][data_n], clipped in place; d_minmax: 2 * n_slices ints of scratch.
namespace boa {
__device__ __forceinline__ int float_order(float f) {  // monotone map float -> int (for atomicMin / atomicMax)
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
slice_minmax_kernel(const float* __restrict__ ref, size_t ref_n, int chunks, int* __restrict__ minmax) {
  const int sl = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const size_t per = (ref_n + chunks - 1) / chunks;
  const size_t b = (size_t)ch * per, e = b + per < ref_n ? b + per : ref_n;
  float lo = INFINITY, hi = -INFINITY;
  for (size_t i = b + threadIdx.x; i < e; i += 256) {
    const float v = __ldg(ref + (size_t)sl * ref_n + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  for (int off = 16; off; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&minmax[2 * sl], float_order(lo));
    atomicMax(&minmax[2 * sl + 1], float_order(hi));
  }
}

__global__ void __launch_bounds__(256)
clip_slices_kernel(float* __restrict__ data, size_t data_n, int n_slices, const int* __restrict__ minmax) {
  const size_t total = (size_t)n_slices * data_n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int sl = (int)(i / data_n);
    const float lo = order_float(minmax[2 * sl]), hi = order_float(minmax[2 * sl + 1]);
    data[i] = fminf(fmaxf(data[i], lo), hi);
  }
}

__global__ void minmax_init_kernel(int* minmax, int n_slices) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slices) { minmax[2 * i] = 0x7fffffff; minmax[2 * i + 1] = (int)0x80000000; }
}

// Order-0 pick along dim 0 with grid coordinates: index = floor(scale * (k + 0.5)), scale = n_in / n_out - what
// map_coordinates(order=0, mode="nearest") does with the coordinates scale * (k + 0.5) - 0.5
// (default_resampling.py:176-192).
__global__ void __launch_bounds__(256)
nearest_z_grid_f32_kernel(const float* __restrict__ in, int z_in, size_t plane, int z_out, float* __restrict__ out) {
  const size_t total = (size_t)z_out * plane;
  const double scale = (double)z_in / (double)z_out;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i / plane);
    int k = (int)floor(scale * ((double)o + 0.5) - 0.5 + 0.5);
    k = k < 0 ? 0 : (k >= z_in ? z_in - 1 : k);
    out[i] = in[(size_t)k * plane + i % plane];
  }
}
}  // namespace boa

extern "C" int boa_clip_slices_f32(const float* d_ref, size_t ref_slice_voxels, float* d_data, size_t data_slice_voxels,
                                   int n_slices, int32_t* d_minmax, void* stream) {
  BOA_REQUIRE(d_ref && d_data && d_minmax && n_slices > 0 && ref_slice_voxels > 0 && data_slice_voxels > 0,
              "boa_clip_slices_f32: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  minmax_init_kernel<<<(n_slices + 255) / 256, 256, 0, s>>>(d_minmax, n_slices);
  BOA_CHECK_LAUNCH();
  int chunks = (int)((ref_slice_voxels + 65535) / 65536);
  const int cap = (sm_count() * 8 + n_slices - 1) / n_slices;
  chunks = chunks > cap ? cap : chunks;
  chunks = chunks < 1 ? 1 : chunks;
  slice_minmax_kernel<<<(unsigned)(n_slices * chunks), 256, 0, s>>>(d_ref, ref_slice_voxels, chunks, d_minmax);
  BOA_CHECK_LAUNCH();
  clip_slices_kernel<<<grid_for((size_t)n_slices * data_slice_voxels, 256), 256, 0, s>>>(d_data, data_slice_voxels, n_slices,
                                                                                        d_minmax);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_resample_z_nearest_grid_f32(const float* d_in, int z_in, size_t plane, int z_out, float* d_out,
                                               void* stream) {
  BOA_REQUIRE(d_in && d_out && z_in >= 1 && z_out >= 1 && plane > 0, "boa_resample_z_nearest_grid_f32: bad argument");
  nearest_z_grid_f32_kernel<<<grid_for((size_t)z_out * plane, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, z_in, plane, z_out, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_resample_nearest_u8(const uint8_t* d_in, const int32_t* in_shape, const int32_t* out_shape,
                                       uint8_t* d_out, void* stream) {
  BOA_REQUIRE(d_in && d_out && in_shape && out_shape, "boa_resample_nearest_u8: null pointer");
  for (int k = 0; k < 3; ++k)
    BOA_REQUIRE(in_shape[k] >= 1 && out_shape[k] >= 1, "boa_resample_nearest_u8: bad shape");
  const size_t total = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
  nearest_3d_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, in_shape[0], in_shape[1], in_shape[2], out_shape[0], out_shape[1], out_shape[2], d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}
