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
    out[i] = (int16_t)(int)v;  // truncation toward zero, as astype(np.int32)
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
