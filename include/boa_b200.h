/* boa_b200.h - C ABI of libboa_b200.so: the B200-native drop-in for the BOA segmentation + body-composition hot path.
 *
 * The reference (UMEssen/Body-and-Organ-Analysis) is 100 % Python and has no FFI; its seams for this path are Python
 * callables.  Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference/body_organ_analysis).  The Python mirror of those interfaces (package boa_b200) binds this
 * library with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named d_* is a CUDA DEVICE pointer owned by the caller; h_* is a HOST pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls enqueue work and return, unless
 *     stated otherwise;
 *   - return value 0 = ok, negative = error; boa_last_error() returns a thread-local message;
 *   - nothing here falls back to a CPU implementation: without a CUDA device every compute entry fails with
 *     BOA_ERR_CUDA.
 *   - volumes are C-contiguous [d0][d1][d2] (nnU-Net's (x, y, z) naming of the array axes; with SimpleITK I/O that
 *     is (z, y, x) of the patient) - innermost axis contiguous.
 */
#ifndef BOA_B200_H
#define BOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOA_OK 0
#define BOA_ERR_ARG (-1)
#define BOA_ERR_CUDA (-2)
#define BOA_ERR_STATE (-3)
#define BOA_ERR_UNSUPPORTED (-4)

#define BOA_MAX_STAGES 8

/* dtype tags for CT volumes */
#define BOA_DT_I16 0
#define BOA_DT_F32 1
#define BOA_DT_F64 2

const char* boa_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int boa_abi_version(void);
/* Number of kernels this library launched since load (diagnostic: proves the native path ran). */
uint64_t boa_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Network: PlainConvUNet forward + Gaussian-weighted sliding-window accumulation.
 * Replaces  nnUNetPredictor.initialize_from_trained_model_folder            (_external/nnunetv2/inference/predict_from_raw_data.py:67-129)
 *           get_network_from_plans / PlainConvUNet(**arch_kwargs)            (_external/nnunetv2/utilities/get_network_from_plans.py:9-43,
 *                                                                             _external/nnunetv2/utilities/plans_handling/plans_handler.py:36-97)
 *           self.network(x) + `pred *= g; logits[sl] += pred; n[sl] += g`    (predict_from_raw_data.py:543,603-616)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct boa_arch {
  int32_t n_stages;                       /* encoder stages (decoder has n_stages-1 levels)            */
  int32_t in_channels;                    /* 1 for CT                                                  */
  int32_t num_classes;                    /* segmentation heads incl. background                       */
  int32_t features[BOA_MAX_STAGES];       /* min(base * 2^i, max)  (plans_handler.py:66-68)            */
  int32_t n_conv_enc[BOA_MAX_STAGES];
  int32_t n_conv_dec[BOA_MAX_STAGES];     /* n_stages-1 entries, decoder level order (deepest first)   */
  int32_t strides[BOA_MAX_STAGES][3];     /* pool_op_kernel_sizes; stage 0 is [1,1,1]                  */
  int32_t kernels[BOA_MAX_STAGES][3];     /* conv_kernel_sizes; only 3x3x3 (and 1x1x1) are implemented */
  int32_t patch[3];                       /* patch_size                                                */
  float eps;                              /* InstanceNorm eps (1e-5)                                   */
  float leaky_slope;                      /* LeakyReLU negative slope (0.01)                           */
} boa_arch;

typedef struct boa_net boa_net;

/* Create an empty network on CUDA device `device`; max_batch = patches processed per forward launch group. */
int boa_net_create(const boa_arch* arch, int device, int max_batch, boa_net** out);
/* Before boa_net_finalize: make `net` use the scratch memory (activations, statistics) of `donor`, a network of
 * identical geometry and batch on the same device (fold ensembles, the five part models of `total`).  Networks that
 * share a workspace must not run concurrently. */
int boa_net_share_workspace(boa_net* net, boa_net* donor);
/* Supply one state-dict tensor by its dynamic_network_architectures key name (e.g.
 * "encoder.stages.0.0.convs.0.conv.weight"), fp32 on the HOST, C-contiguous PyTorch layout. Unknown alias keys
 * (all_modules.*, decoder.encoder.*) are accepted and ignored. */
int boa_net_set_tensor(boa_net* net, const char* key, const float* h_data, const int64_t* shape, int ndim);
/* Pack weights into MMA operand layout, allocate activation buffers. Fails if a required tensor is missing. */
int boa_net_finalize(boa_net* net);
/* 0: tcgen05 tensor-core path (default, product). 1: SIMT reference kernels (debug cross-check only). */
int boa_net_set_mode(boa_net* net, int mode);
/* Forward n_patches patches cut from the normalised volume d_vol (fp32 [d0][d1][d2]) at the given origins (HOST
 * int32 [n][3]); adds logits * gaussian into d_logits_acc (fp32 [num_classes][d0][d1][d2]) in patch order, exactly
 * as predict_from_raw_data.py:609-613.  d_gaussian is fp32 [patch0][patch1][patch2]. */
int boa_net_forward_accumulate(boa_net* net, const float* d_vol, const int32_t* vol_shape, const int32_t* h_origins,
                               int n_patches, const float* d_gaussian, float* d_logits_acc, void* stream);
/* Debug/parity: forward ONE batch of patches already cut out, d_patches fp32 [n][Cin][p0][p1][p2], and write the raw
 * logits fp32 [n][num_classes][p0][p1][p2]. */
int boa_net_forward_logits(boa_net* net, const float* d_patches, int n_patches, float* d_logits, void* stream);
/* 1 (default): replay one captured CUDA graph per batch of patches; 0: plain kernel launches. */
int boa_net_set_graph(boa_net* net, int enable);
/* Launch schedule of one forward: for each step (conv / transposed conv) its kernel kind (0 = tcgen05 dz-folded,
 * 1 = tcgen05 tap list, 2 = SIMT conv, 3 = tcgen05 transposed conv, 4 = SIMT transposed conv), its algorithmic MACs
 * per patch and its state-dict prefix.  Returns the number of steps. */
int boa_net_describe(const boa_net* net, int cap, int32_t* kinds, double* macs, char* names, int name_stride);
/* Run one forward body on the current activations with a CUDA-event pair around every conv kernel and return the
 * per-step device milliseconds (profiling / roofline accounting only). */
int boa_net_time_layers(boa_net* net, int cap, float* ms, void* stream);
/* Algorithmic multiply-accumulates of one patch forward (for roofline accounting). */
int64_t boa_net_macs_per_patch(const boa_net* net);
/* Cumulative milliseconds spent in conv kernels since the last reset, measured with CUDA events on `stream` when
 * timing is enabled (bench only). */
int boa_net_enable_timing(boa_net* net, int enable);
int boa_net_read_timing(boa_net* net, double* ms_convs, double* ms_total, int64_t* n_conv_launches, int reset);
/* The same record split by kernel kind (0 dz-folded tcgen05 conv, 1 stride-2 tap-list conv, 2 SIMT conv, 3 transposed
 * tap-list conv, 4 SIMT transposed conv, 5 first-layer SIMT conv, 6 head + Gaussian accumulate - for kind 6 `flop`
 * holds algorithmic BYTES): milliseconds, algorithmic FLOP and launches of the kernels launched by forward_accumulate
 * calls while timing was enabled (single-lane schedule, exclusive event brackets) - the roofline of bench.py is
 * computed from these, inside a long step.  Bench only; the reference has no counterpart. */
int boa_net_read_timing_kinds(boa_net* net, int n_kinds, double* ms, double* flop, int64_t* launches, int reset);
void boa_net_destroy(boa_net* net);

/* ------------------------------------------------------------------------------------------------------------
 * Memory-bound passes
 * ------------------------------------------------------------------------------------------------------------ */

/* CTNormalization.run: clip to [lo, hi], subtract mean, divide by max(std, 1e-8), all in fp32.
 * (_external/nnunetv2/preprocessing/normalization/default_normalization_schemes.py:53-67) */
int boa_ct_normalize(const void* d_in, int in_dtype, size_t n, float lo, float hi, float mean, float std,
                     float* d_out, void* stream);

/* Standalone `pred *= g; logits[sl] += pred` for one patch whose logits are already on the device
 * (predict_from_raw_data.py:609-613): d_logits fp32 [C][p0][p1][p2]. */
int boa_accumulate_patch(const float* d_logits, int C, const int32_t* patch, const int32_t* origin,
                         const float* d_gaussian, float* d_logits_acc, const int32_t* vol_shape, void* stream);
/* n_predictions for a list of origins (input independent): d_weight_acc fp32 [d0][d1][d2] += g at each origin, in
 * order (predict_from_raw_data.py:614). */
int boa_accumulate_weights(const int32_t* h_origins, int n_patches, const int32_t* patch, const float* d_gaussian,
                           float* d_weight_acc, const int32_t* vol_shape, void* stream);

/* `logits /= n` + isinf check + argmax over channels (first maximum wins) + part->global label LUT + non-zero
 * overwrite merge, in one pass.
 * (predict_from_raw_data.py:620-625; _external/nnunetv2/utilities/label_handling/label_handling.py:174-180;
 *  _external/totalsegmentator/nnunet.py:553-556)
 * d_label_inout [V] uint8: when overwrite_nonzero_only != 0 only voxels whose mapped label is non-zero are written.
 * d_nonfinite: int32 counter incremented for every non-finite quotient (caller zeroes it). */
int boa_finalize_argmax(const float* d_logits_acc, const float* d_weight_acc, int C, size_t V, const uint8_t* h_lut,
                        int overwrite_nonzero_only, uint8_t* d_label_inout, int32_t* d_nonfinite, void* stream);

/* The same division and check WITHOUT the argmax, for the entry that returns logits
 * (predict_sliding_window_return_logits, predict_from_raw_data.py:620-625, and the fold mean of
 * predict_logits_from_preprocessed_data :494-500): d_logits_acc[c][v] /= d_weight_acc[v] * folds in place.
 * d_nonfinite: int32 counter incremented when a quotient is not finite (caller zeroes it). */
int boa_normalize_logits(float* d_logits_acc, const float* d_weight_acc, int C, size_t V, float folds,
                         int32_t* d_nonfinite, void* stream);

/* subclassify_tissues numerics: tissue id from (HU, body region) with inclusive HU bounds
 * (_external/body_composition_analysis/tissue/subclassification.py:38-53, tissue/definition.py:6-30). */
int boa_tissue_subclassify(const void* d_ct, int ct_dtype, const uint8_t* d_regions, size_t n, uint8_t* d_tissues,
                           void* stream);

/* Per-slice (outermost axis) per-label voxel counts and HU sums, optionally restricted to voxels where
 * d_mask == mask_value.  counts uint64 [Z][n_labels], hu_sums int64 [Z][n_labels] (either may be NULL; d_ct may be
 * NULL when hu_sums is NULL).  The caller zeroes the outputs.
 * (_external/body_composition_analysis/report/builder.py:403-444,284-305,56-99; body_composition_analysis/commands.py:24-45) */
int boa_slice_label_stats(const uint8_t* d_labels, const uint8_t* d_mask, int mask_value, const void* d_ct,
                          int ct_dtype, int Z, size_t slice_voxels, int n_labels, uint64_t* d_counts,
                          int64_t* d_hu_sums, void* stream);

/* Per-label integer-HU histograms: d_hist uint32 [n_labels][n_bins] += 1 at bin (hu - hu_min) for every voxel with
 * label in [1, n_labels); HU outside [hu_min, hu_min + n_bins) increments d_out_of_range instead.  All statistics of
 * metrics_for_region (count, mean, std, min, median, max, percentiles, fat-window subsets, unions) are exact functions
 * of these histograms.  (compute/measurements.py:74-123,126-148,203-241) */
int boa_label_hu_hist(const void* d_ct, int ct_dtype, const uint8_t* d_labels, size_t n, int n_labels, int hu_min,
                      int n_bins, uint32_t* d_hist, uint32_t* d_out_of_range, void* stream);

/* create_mask (compute/util.py:25-31) optionally combined with an HU window:
 *   mode 0: mask = label in set;  mode 1: ... AND lo <= hu <= hi  (lung fat, compute/measurements.py:134-141);
 *   mode 2: ... AND (hu < lo OR hu > hi)  (region minus fat, compute/measurements.py:29-39).
 * h_label_set: 256 host bytes, non-zero = label selected.  d_mask uint8 [n] (0/1). */
int boa_mask_label_minus_window(const void* d_ct, int ct_dtype, const uint8_t* d_labels, size_t n,
                                const uint8_t* h_label_set, int lo, int hi, int mode, uint8_t* d_mask, void* stream);

/* Binary erosion with a box window of offsets [-before, +after] on every axis, voxels outside the volume treated as
 * foreground: erode_region (compute/measurements.py:61-71) is before = 3, after = 2.  d_tmp: scratch uint8 [V]. */
int boa_erode_box(const uint8_t* d_mask, const int32_t* shape, int before, int after, uint8_t* d_tmp, uint8_t* d_out,
                  void* stream);

/* Slice-thickness resampling of the body-composition path (resample_only_thickness,
 * _external/totalsegmentator/nnunet.py:457-475; scipy.ndimage.zoom(order=3, mode="nearest") with zoom (z, 1, 1),
 * _external/totalsegmentator/resampling.py:24-56): interpolating cubic B-spline along the outermost axis in fp64,
 * truncated toward zero like `astype(np.int32)`.  d_scratch: double [(z_in + 24) * plane].  plane = d1 * d2. */
int boa_resample_z_cubic(const void* d_in, int in_dtype, int z_in, size_t plane, int z_out, double* d_scratch,
                         int16_t* d_out, void* stream);
/* Order-0 zoom of a label map along the outermost axis back to the input grid (nnunet.py:685-687). */
int boa_resample_z_nearest_u8(const uint8_t* d_in, int z_in, size_t plane, int z_out, uint8_t* d_out, void* stream);

/* 3-D resampling to / from the network spacing (change_spacing(img, [1.5]*3, order=3, dtype=int32) before and
 * change_spacing(seg, target_shape=original, order=0) after the networks, _external/totalsegmentator/nnunet.py:466-470,
 * 685-687 -> resampling.py:24-56,129-222; scipy.ndimage.zoom(mode="nearest")).  The cubic zoom is a tensor product:
 * call boa_resample_axis_cubic once per axis on a volume viewed as [outer][n_in][inner].  in_dtype: BOA_DT_I16 /
 * F32 / F64; out_mode 0: fp64 (feed the next pass), 1: int16, 2: int32 (both truncated toward zero like astype()).
 * d_scratch: double [outer * (n_in + 24) * inner]. */
int boa_resample_axis_cubic(const void* d_in, int in_dtype, size_t outer, int n_in, size_t inner, int n_out,
                            double* d_scratch, void* d_out, int out_mode, void* stream);
/* Order-0 zoom of a uint8 label map [z][y][x] from in_shape to out_shape (host int32[3] each). */
int boa_resample_nearest_u8(const uint8_t* d_in, const int32_t* in_shape, const int32_t* out_shape, uint8_t* d_out,
                            void* stream);
/* One axis of skimage.transform.resize(order=3, mode="edge", anti_aliasing=False) = scipy.ndimage.zoom(order=3,
 * mode="nearest", grid_mode=True): what nnU-Net's DefaultPreprocessor resamples the NORMALISED image with
 * (_external/nnunetv2/preprocessing/resampling/default_resampling.py:117-203, default_preprocessor.py:57-90).  Same
 * addressing as boa_resample_axis_cubic; in_dtype fp32 / fp64; out_mode 0 = fp64 (intermediate pass), 3 = fp32. */
int boa_resample_axis_cubic_grid(const void* d_in, int in_dtype, size_t outer, int n_in, size_t inner, int n_out,
                                 double* d_scratch, void* d_out, int out_mode, void* stream);
/* resize(..., clip=True): clip d_data [n_slices][data_slice_voxels] in place to the value range of the matching slice
 * of the resize's input d_ref [n_slices][ref_slice_voxels] (n_slices = 1: whole volume).  d_minmax: 2 * n_slices ints. */
int boa_clip_slices_f32(const float* d_ref, size_t ref_slice_voxels, float* d_data, size_t data_slice_voxels,
                        int n_slices, int32_t* d_minmax, void* stream);
/* Order-0 pick along dim 0 with grid coordinates (order_z = 0 of the separate-z branch, default_resampling.py:176-192). */
int boa_resample_z_nearest_grid_f32(const float* d_in, int z_in, size_t plane, int z_out, float* d_out, void* stream);
/* boa_finalize_argmax when the network ran on a resampled grid: every class channel of logits / n is resampled with
 * order 1 from net_shape to out_shape before the argmax (export_prediction.py:25-38; separate_z: order-0 pick along
 * dim 0, bilinear in plane).  d_logits_acc [C][net_shape], d_weight_acc [net_shape], d_label_inout [out_shape]. */
int boa_finalize_argmax_resampled(const float* d_logits_acc, const float* d_weight_acc, int C, const int32_t* net_shape,
                                  const int32_t* out_shape, int separate_z, const uint8_t* h_lut,
                                  int overwrite_nonzero_only, uint8_t* d_label_inout, int32_t* d_nonfinite, void* stream);

/* In-plane 3x3 median of every z-slice, boundary mode "reflect": scipy.ndimage.median_filter(image, size=[1,3,3]) of
 * subclassify_tissues(median_filtering=True) (_external/body_composition_analysis/tissue/subclassification.py:20-36,
 * CLI --bca-median-filtering).  int16 [z][y][x] -> int16, not in place. */
int boa_median3x3_slices(const int16_t* d_in, const int32_t* shape, int16_t* d_out, void* stream);

/* Connected-component post-processing of the body-composition label maps (between the networks and the tissue rules,
 * _external/body_composition_analysis/infer/infer.py:81-89):
 *   postprocess_region_segmentation (body_regions/postprocess.py:8-40): skimage.measure.label (26-connected) +
 *     regionprops, every component of a label set but the largest -> 255                                   = op 0
 *   postprocess_part_segmentation (body_parts/postprocess.py:7-60): slice-wise cv2 external-contour fill = op 2 on the
 *     complement with mode 1; remove_small_objects(max_size, connectivity=3) on objects / holes           = op 1
 * The set is {v : h_label_set[d_seg[v]] != 0} (256 host bytes), complemented when invert != 0.  mode 0: 26-connected in
 * 3-D, 1: 4-connected inside every z-slice.  op 0: voxels of the set outside its largest component (first in raster
 * order on ties) := fill_value; op 1: voxels in components of size <= threshold := fill_value; op 2: voxels in components
 * that do not touch the border of their slice := fill_value.  Sizes are sum of d_slice_weight[z] (int32 [D], null = 1).
 * d_labels, d_sizes, d_border: int32 [V] scratch (d_border: op 2 only); d_best: 16 bytes scratch.  V < 2^31. */
int boa_cc_filter(uint8_t* d_seg, const int32_t* shape, const uint8_t* h_label_set, int invert, int mode, int op,
                  int threshold, int fill_value, const int32_t* d_slice_weight, int32_t* d_labels, int32_t* d_sizes,
                  int32_t* d_border, void* d_best, void* stream);
/* d_out[v] = label where d_mask[v] != 0 (body_parts/postprocess.py:50). */
int boa_paint_label(const uint8_t* d_mask, size_t n, int label, uint8_t* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  The reference has no multi-GPU inference path: these entry points
 * replace nothing, they extend predict_sliding_window_return_logits (predict_from_raw_data.py:560-631) across GPUs.
 * ------------------------------------------------------------------------------------------------------------ */
/* NCCL path, device side: d_dst[i] += d_src[i] (fp32, round-to-nearest, fixed order chosen by the caller). */
int boa_add_slab(float* d_dst, const float* d_src, size_t n, void* stream);
/* d_dst[c * dst_cstride + i] += d_src[c * src_cstride + i] for c < C, i < n (strides in elements): one launch per
 * received piece [C, len, Y, X]. */
int boa_add_slab_strided(float* d_dst, size_t dst_cstride, const float* d_src, size_t src_cstride, int C, size_t n,
                         void* stream);
/* Peer-memory path.  boa_comm_alloc: device memory that other processes of the box can map - returns the pointer and
 * its 64-byte CUDA IPC handle; boa_comm_open maps a peer's buffer (lazy peer access over NVLink), boa_comm_close
 * unmaps it, boa_comm_free releases an own buffer, boa_comm_zero clears it on a stream. */
#define BOA_IPC_HANDLE_BYTES 64
int boa_comm_alloc(size_t bytes, void** d_ptr, unsigned char* ipc_handle_out);
int boa_comm_open(const unsigned char* ipc_handle, void** d_ptr);
int boa_comm_close(void* d_ptr);
int boa_comm_free(void* d_ptr);
int boa_comm_zero(void* d_ptr, size_t bytes, void* stream);
/* Exchange + reduction + finalize of ONE dim-0 slab in one kernel: rank r's private buffer d_peer_bases[r] holds the
 * partial sums [C][zhi[r] - zlo[r]][Y][X] of its patches (slices zlo[r] .. zhi[r] of the volume); the caller owns the
 * slices slab_lo .. slab_hi, reads every rank's part of them through the mapped pointers, adds them in rank order and
 * then does exactly what boa_finalize_argmax does (divide by d_weight_slab, non-finite flag, argmax with first maximum,
 * LUT, optional non-zero overwrite) into d_label_slab [slab_hi - slab_lo][Y][X].  n_ranks <= 8. */
int boa_reduce_finalize_peers(const void* const* d_peer_bases, const int32_t* zlo, const int32_t* zhi, int n_ranks,
                              int slab_lo, int slab_hi, int Y, int X, const float* d_weight_slab, int C,
                              const uint8_t* h_lut, int overwrite_nonzero_only, uint8_t* d_label_slab,
                              int32_t* d_nonfinite, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BOA_B200_H */
