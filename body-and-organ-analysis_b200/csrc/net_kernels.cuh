// Declarations shared between the network orchestrator (net.cu), the elementwise / SIMT kernels (net_simt.cu) and
// the tcgen05 implicit-GEMM kernels (conv_mma.cu, conv_taps.cu).
//
// Activation layout "C8": [batch][C/8][D][H][W][8] fp16 - channel groups of 8 outermost, the 8 channels of a group
// innermost (16 bytes).  A TMA box over it lands in shared memory as the no-swizzle K-major UMMA operand layout
// directly (16-byte rows, see ptx.cuh umma_desc), tap shifts are plain address offsets, and a channel concat is two
// adjacent sub-tensors.
#pragma once
#include "common.cuh"
#include "conv_xform.cuh"

namespace boa {

constexpr int MAX_BATCH = 32;

struct ActView {  // a C8 tensor, possibly a channel-group slice of a wider buffer
  __half* base = nullptr;  // start of the whole buffer
  int groups_total = 0;    // channel groups of the whole buffer (per batch item)
  int group_off = 0;       // first group of this view
  int groups = 0;          // groups in this view
  int D = 0, H = 0, W = 0;
  size_t voxels() const { return (size_t)D * H * W; }
};

// Per-forward call parameters, kept in DEVICE memory so that the launch sequence of one forward is identical from
// call to call (CUDA-graph friendly): kernels that touch the volume read them from here.
struct FwdCall {
  const float* vol;       // normalised volume fp32 [d0][d1][d2]
  float* acc;             // logits accumulator fp32 [C][d0][d1][d2]
  const float* gaussian;  // fp32 [p0][p1][p2]
  int32_t d0, d1, d2;
  int32_t n_valid;        // patches of this batch that are real (the rest are padding, never accumulated)
  int32_t origins[MAX_BATCH][3];
};

// ---- elementwise / SIMT (net_simt.cu)
// predict_from_raw_data.py:568-571: cut `data[sl]` -> fp16 C8 with 2 groups (channel 0 = voxel, others 0).
// mode 0: 16-channel C8 tensor, channel 0 = voxel.  mode 1: plain fp16 [B][p0][p1][p2] (direct first-layer kernel).
// mode 2: 16-channel C8 tensor whose channels 0..8 are the 9 in-plane (dy,dx) neighbours of the voxel, zero outside
//         the PATCH (first layer on the tensor cores with the in-plane taps on K).
int launch_extract_patches(const FwdCall* d_call, int B, int p0, int p1, int p2, __half* d_out, int mode,
                           cudaStream_t s);
int launch_pack_patches_plain(const float* d_patches, size_t total, __half* d_out, cudaStream_t s);
int launch_pack_patches_nb9(const float* d_patches, int n, int p0, int p1, int p2, __half* d_out, cudaStream_t s);
// First encoder conv (Cin = 1, 3x3x3, stride 1): direct FP32 convolution from the plain fp16 patch.
bool conv_first_supported(int Cout);
int launch_conv_first(const __half* d_in, int B, const float* d_w, const float* d_bias, int Cout, __half* d_raw_out,
                      int D, int H, int W, double* d_stats, cudaStream_t s);
int launch_pack_patches(const float* d_patches, int n, int cin, int p0, int p1, int p2, __half* d_out_c8,
                        int groups, cudaStream_t s);
// d_scale2 / d_shift2 (optional): second copy at [b * stride2 + off2 + c] (rows of a concat's combined table).
int launch_stats_finalize(const double* d_stats, const float* d_gamma, const float* d_beta, int B, int C,
                          double n_vox, float eps, float* d_scale, float* d_shift, cudaStream_t s,
                          float* d_scale2 = nullptr, float* d_shift2 = nullptr, int stride2 = 0, int off2 = 0);
// y = lrelu(x * scale + shift) -> fp16, written to dst view and (optionally) to a space-to-depth copy
// [B][8 phases][groups][D/2][H/2][W/2][8] (phase = (z&1)*4 + (y&1)*2 + (x&1)) that turns the next stage's stride-2
// convolution into a stride-1 one.
int launch_norm_lrelu(const __half* d_raw, int B, int groups, int D, int H, int W, const float* d_scale,
                      const float* d_shift, float slope, const ActView& dst, __half* d_s2d, cudaStream_t s);
// SIMT direct convolution (anisotropic kernels / strides, and the on-device cross-check of the tensor-core kernels).
// w: fp32 [Cout][cin_w][kz][ky][kx], already rounded to fp16 precision.
// xf: fused normalisation of the input (the source holds RAW producer output), see conv_xform.cuh.
int launch_conv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int cin_w, int Cout,
                     const int* ks, const int* stride, const ActView& out, int Do, int Ho, int Wo, double* d_stats,
                     cudaStream_t s, const InXform& xf = InXform());
// ConvTranspose3d kernel = stride: w fp32 [Cin][Cout][sz][sy][sx]; writes fp16 into dst view.
int launch_tconv_simt(const ActView& src, int B, const float* d_w, const float* d_bias, int Cin, int Cout,
                      const int* stride, const ActView& dst, cudaStream_t s, const InXform& xf = InXform());
// 1x1x1 head. If d_logits_b != nullptr: write raw logits fp32 [C][P] of batch item b. Else accumulate logits*g into
// the volume accumulator at the origin of batch item b (skipped when b >= call->n_valid); one launch per patch keeps
// the reference's patch order (predict_from_raw_data.py:603-616).
// in_scale / in_shift != nullptr: src holds the RAW output of the last conv; its InstanceNorm affine + LeakyReLU are
// applied on the fly ([B][Cin] arrays).
int launch_head(const ActView& src, int b, const float* d_w, const float* d_bias, int Cin, int C, float* d_logits_b,
                const FwdCall* d_call, const float* d_in_scale, const float* d_in_shift, float slope, cudaStream_t s);

// Where a conv kernel writes and how it reads: the RAW output goes to `out` (a dense tensor or a channel-group slice
// of a decoder concat buffer) and optionally also to a space-to-depth copy for the next stage's stride-2 conv; `xf`
// describes the normalisation the kernel applies to its INPUT on the fly (conv_xform.cuh).
struct ConvIO {
  ActView out;
  __half* s2d = nullptr;
  int s2d_stride[3] = {2, 2, 2};  // strides (z, y, x) of the conv that will read the copy: phases = product
  InXform xf;
};

// ---- tcgen05 implicit GEMM, 3x3x3 stride 1 with the dz taps folded into N (conv_mma.cu)
struct ConvMmaPlan;
ConvMmaPlan* conv_mma_plan_create(const float* h_w /*[Cout][Cin_w][27] fp32*/, const float* h_bias, int cin_w,
                                  int cin_padded, int Cout, const ActView& src, int B, const ConvIO& io,
                                  double* d_stats, bool taps_on_k = false, int kz = 3);
void conv_mma_plan_destroy(ConvMmaPlan* p);
// nb: batch items to process (<= the B the plan was created with; the tensors keep their [B]-strided layout)
int conv_mma_launch(ConvMmaPlan* p, cudaStream_t s, int nb = -1);

// ---- tcgen05 implicit GEMM over an explicit tap list (conv_taps.cu): strided convs on the space-to-depth copy,
//      transposed convs (one tap, the output phases on N), stride-1 convs with any kernel in {1,3}^3.
struct ConvTapsPlan;
enum TapsKind { TAPS_CONV = 0, TAPS_TCONV = 2 };
struct TapsGeom {  // per axis (z, y, x)
  int ks[3];       // conv: kernel size 1 or 3 (in-plane sizes equal); transposed conv: unused (kernel = stride)
  int stride[3];   // 1 or 2 (in-plane strides equal; transposed conv: x stride 2)
};
// TAPS_CONV:  h_w [Cout][cin_w][kz*ky*kx]; src = activation view (all strides 1) or the space-to-depth view of it
//             ([phases * groups], dims divided by the strides); out = raw conv output at the source view's dims, stats.
// TAPS_TCONV: h_w [Cin][Cout][sz*sy*sx];   src = activation view; out = dst view at (sz, sy, sx) x the dims (bias only).
ConvTapsPlan* conv_taps_plan_create(TapsKind kind, const TapsGeom& geo, const float* h_w, const float* h_bias,
                                    int cin_w, int Cout, const ActView& src, int B, const ConvIO& io, double* d_stats);
void conv_taps_plan_destroy(ConvTapsPlan* p);
int conv_taps_launch(ConvTapsPlan* p, cudaStream_t s, int nb = -1);

}  // namespace boa
