// tcgen05 implicit GEMM over an explicit per-K-chunk tap list, for sm_100a.
//
// Covers the convolution shapes of PlainConvUNet that the dz-folded kernel (conv_mma.cu) does not.  Per axis the
// kernel size is 1 or 3 and the stride 1 or 2 (nnU-Net plans for anisotropic data - the 5 mm body-composition models -
// carry [1,3,3] kernels and [1,2,2] pools, _external/body_composition_analysis/tasks.py:15-48 + plans_handler.py:36-97):
//   * STRIDED convs (first conv of encoder stages 1..n): a stride-1 gather over the space-to-depth copy of the input
//     ([B][phases][C/8][D/sz][H/sy][W/sx][8], phases = sz*sy*sx, written by the producing conv's epilogue).  On a
//     stride-2 axis output o reads input 2o-1, 2o, 2o+1 = (odd phase, o-1), (even phase, o), (odd phase, o): a K chunk
//     of 16 channels belongs to one phase and only carries the taps that phase can serve - no wasted MACs.
//   * ConvTranspose3d kernel = stride in {1,2}^2 x {2} (decoder up-sampling): a single tap, the output phases sit on N
//     and the epilogue scatters them into the concat buffer.
//   * stride-1 convs with kernels other than 3x3x3 ([1,3,3]: 9 taps per chunk), and plain 3x3x3 (27 taps per chunk)
//     as an independent cross-check of the folded kernel.
// Replaces the cuDNN calls behind nn.Conv3d / nn.ConvTranspose3d of dynamic_network_architectures' PlainConvEncoder /
// UNetDecoder, invoked at _external/nnunetv2/inference/predict_from_raw_data.py:543.
//
// Structure: persistent CTAs of three warpgroups like conv_mma.cu (WG0 = TMA producer + MMA issuer, WG1 = epilogue,
// WG2 = operand transform: the input is the producer's RAW output, normalised in shared memory, conv_xform.cuh);
// A = TMA halo box of the C8 tensor (zero filled out of bounds = conv padding), taps are 16-byte address offsets into
// it; B = weights pre-packed on the host into the smem operand image, one bulk copy per chunk; accumulators: one TMEM
// slot of NC columns per output z-plane of the tile, double buffered.
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "net_kernels.cuh"
#include "conv_epilogue.cuh"
#include "conv_xform.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace boa {

constexpr int TAPS_THREADS = 512;  // four warpgroups (the last two transform), see conv_mma.cu
constexpr int TREGS_WG0 = 96, TREGS_EPI = 224, TREGS_XF = 96, TXF_THREADS = 256;
// transform warps in groups that own ring slots (conv_mma.cu): the stages of these kernels are small (a few hundred
// cycles of MMAs), so four stages are transformed concurrently, two warps each
// The number of groups (p.xf_groups) divides the ring depth, which is trimmed to a multiple of four for that.
constexpr int TT_X = 8, TT_Y = 16;
constexpr int TAPS_MAX_OPS = 27;
constexpr int TAPS_MAX_STAGES = 12;  // smem ring depth: small stages (transposed conv, deep layers) prefetch several tiles ahead
constexpr int TAB_STRIDE = 32;  // ints per class-table entry: [0] n_ops, [1] ops of all earlier chunks, [2..28] a_off,
                                //                             [30] first channel group of the class's phase
constexpr int TAB_GBASE = 30;

struct TapsParams {
  const __half* bpacked;
  const float* bias;
  __half* out;
  double* stats;
  const int32_t* table;  // [n_classes][TAB_STRIDE]: n_ops, ops before this class, a_off[n_ops]
  int n_classes, chunks_per_class;
  int kind;
  int B, kc_count, Ntotal, D, H, W, zt;
  int box_x, box_y, box_z, org_x, org_y, org_z;
  int tsz, tsy;  // transposed conv: output phases along z and y (1 or 2; x is always 2)
  int tiles_x, tiles_y, tiles_z, n_ntiles, total_tiles;
  int in_groups_total, in_group_off;
  int out_groups_total, out_group_off, Cout;
  __half* s2d;   // optional space-to-depth copy of a conv output (nullptr: none)
  int s2d_stride[3];  // strides (z, y, x) of the conv that will read the copy
  InXform xf;    // fused normalisation of the input (xf.scale == nullptr: none)
  int xf_groups; // transform warp groups (divides `stages`)
  int xf_debug;  // profiling aid (BOA_B200_XF_DEBUG): 1 = transform warps only relay the barrier
  int halo;      // box = tile + halo per axis: 2 (3x3x3 stride 1), 1 (stride 2 on the s2d tensor), 0 (transposed)
  int stages;
  int tmap_merged;  // tensor map built with the (channel, x) dimensions merged (tmap.cuh)
  uint32_t a_bytes, a_tx_bytes, b_stage_bytes, b_nt_bytes;  // a_bytes: smem placement (128 B multiple), a_tx: TMA box
};

__device__ __forceinline__ void taps_decode_tile(int t, const TapsParams& p, int& nt, int& b, int& tz, int& ty,
                                                 int& tx) {
  tx = t % p.tiles_x; t /= p.tiles_x;
  ty = t % p.tiles_y; t /= p.tiles_y;
  tz = t % p.tiles_z; t /= p.tiles_z;
  b = t % p.B;
  nt = t / p.B;
}

template <int NC>
__global__ void __launch_bounds__(TAPS_THREADS, 1)
conv_taps_kernel(const __grid_constant__ CUtensorMap tmapA, const TapsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t stage_bytes = p.a_bytes + p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;                            // [TAPS_MAX_STAGES] TMA -> MMA
  uint64_t* empty = bars + TAPS_MAX_STAGES;         // [TAPS_MAX_STAGES] MMA -> TMA
  uint64_t* rawfull = bars + 2 * TAPS_MAX_STAGES;   // [TAPS_MAX_STAGES] TMA -> transform warps
  uint64_t* tfull = bars + 3 * TAPS_MAX_STAGES;     // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;                     // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int32_t* stab = reinterpret_cast<int32_t*>(tempty + 4);  // [n_classes][TAB_STRIDE] tap table (<= 8 classes)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstage = p.stages;
  const bool xform = p.xf.scale != nullptr;
  for (int i = threadIdx.x; i < p.n_classes * TAB_STRIDE; i += blockDim.x) stab[i] = __ldg(p.table + i);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TAPS_MAX_STAGES; ++i) {
      mbar_init(&full[i], xform ? TXF_THREADS / 32 / p.xf_groups : 1);
      mbar_init(&empty[i], 1);
      mbar_init(&rawfull[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmapA);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t buf_cols = (uint32_t)(p.zt * NC);  // <= 256: two accumulator buffers

  if (warp < 4) {
  reg_dealloc<TREGS_WG0>();
  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t it = 0;
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int nt, b, tz, ty, tx;
      taps_decode_tile(tile, p, nt, b, tz, ty, tx);
      for (int kc = 0; kc < p.kc_count; ++kc, ++it) {
        mbar_wait(&empty[st], ph ^ 1);
        const int cls = kc / p.chunks_per_class;
        const int n_ops = __ldg(p.table + cls * TAB_STRIDE);
        // weights are packed chunk-major: all ops of the classes before this one, then this class's earlier chunks
        const int b_off = (__ldg(p.table + cls * TAB_STRIDE + 1) + (kc - cls * p.chunks_per_class) * n_ops) * NC * 2;
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)st * stage_bytes;
          const uint32_t bbytes = (uint32_t)n_ops * NC * 32u;
          uint64_t* ready = xform ? &rawfull[st] : &full[st];
          mbar_arrive_expect_tx(ready, p.a_tx_bytes + bbytes);
          tma_load_c8(sa, &tmapA, ready, p.tmap_merged, tx * TT_X + p.org_x, ty * TT_Y + p.org_y, tz * p.zt + p.org_z,
                      b * p.in_groups_total + p.in_group_off + __ldg(p.table + cls * TAB_STRIDE + TAB_GBASE) +
                          2 * (kc - cls * p.chunks_per_class));
          bulk_load(sa + p.a_bytes,
                    reinterpret_cast<const uint8_t*>(p.bpacked) + (size_t)nt * p.b_nt_bytes + (size_t)b_off * 16u,
                    bbytes, ready);
        }
        __syncwarp();
        if (++st == nstage) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (ONE thread runs the loop)
    if (elect_one()) {
      uint32_t tcount = 0;
      int st = 0;
      uint32_t ph = 0;
      const uint32_t slab16 = (uint32_t)(p.box_x * p.box_y);           // 16-byte units per z-plane of the box
      const uint32_t a_lbo = (uint32_t)p.box_z * slab16 * 16u;         // next channel group (8 channels of K)
      const uint64_t a_desc0 = umma_desc(0, a_lbo, (uint32_t)p.box_x * 16u);
      const uint64_t b_desc0 = umma_desc(0, NC * 16u, 128u);
      const uint32_t idesc = umma_idesc_f16(128, NC);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(&tempty[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dcol0 = tbase + buf * buf_cols;
        for (int kc = 0; kc < p.kc_count; ++kc) {
          const int32_t* tab = stab + (kc / p.chunks_per_class) * TAB_STRIDE;
          const int n_ops = tab[0];
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + (size_t)st * stage_bytes);
          const uint64_t a_base = a_desc0 + (uint64_t)(a0 >> 4);
          const uint64_t b_base = b_desc0 + (uint64_t)((a0 + p.a_bytes) >> 4);
          for (int op = 0; op < n_ops; ++op) {
            const uint64_t ad = a_base + (uint64_t)(uint32_t)tab[2 + op];
            const uint64_t bd = b_base + (uint64_t)(op * NC * 2);
            const uint32_t acc = (kc | op) ? 1u : 0u;
            for (int z = 0; z < p.zt; ++z)
              umma_f16(dcol0 + (uint32_t)(z * NC), ad + (uint64_t)(z * slab16), bd, idesc, acc);
          }
          umma_commit(&empty[st]);
          if (++st == nstage) { st = 0; ph ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
    }
    __syncwarp();
  }
  } else if (warp >= 8) {
    // ===================================================================== operand transform (warps 8..15)
    reg_dealloc<TREGS_XF>();
    if (xform) {
      const int gthreads = TXF_THREADS / p.xf_groups;
      const int grp = (threadIdx.x - 256) / gthreads;
      const int tid = (threadIdx.x - 256) % gthreads;
      uint32_t cnt = 0;  // ring stage counter over (tile, kc)
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, b, tz, ty, tx;
        taps_decode_tile(tile, p, nt, b, tz, ty, tx);
        // box index i <-> tensor coordinate t * T + org + i: the in-volume part of the box
        const int z0 = tz * p.zt + p.org_z, y0 = ty * TT_Y + p.org_y, x0 = tx * TT_X + p.org_x;
        const int zlo = z0 < 0 ? -z0 : 0, zhi = p.D - z0 < p.box_z ? p.D - z0 : p.box_z;
        const int ylo = y0 < 0 ? -y0 : 0, yhi = p.H - y0 < p.box_y ? p.H - y0 : p.box_y;
        const int xlo = x0 < 0 ? -x0 : 0, xhi = p.W - x0 < p.box_x ? p.W - x0 : p.box_x;
        for (int kc = 0; kc < p.kc_count; ++kc, ++cnt) {
          const int st = (int)(cnt % (uint32_t)nstage);
          if (st % p.xf_groups != grp) continue;
          const uint32_t ph = (cnt / (uint32_t)nstage) & 1u;
          // channels of this K chunk: the stride-2 gather walks the 8 phases of the space-to-depth tensor, each
          // holding every channel (chunk index inside the phase = kc % chunks_per_class)
          const int cc = kc % p.chunks_per_class;
          const int g0 = 2 * cc;
          const int skip = (g0 < p.xf.ident_groups ? 1 : 0) | (g0 + 1 < p.xf.ident_groups ? 2 : 0) |
                           (8 * g0 + 8 >= p.xf.channels ? 2 : 0);
          mbar_wait(&rawfull[st], ph);
          if (skip != 3 && p.xf_debug != 1) {
            uint8_t* sa = smem + (size_t)st * stage_bytes;
            const float* sc = p.xf.scale + (size_t)b * p.xf.channels + 16 * cc;
            const float* sh = p.xf.shift + (size_t)b * p.xf.channels + 16 * cc;
            // work items = (z phase, in-plane position) with at most 3 planes each: z phases of 2 for boxes of up to 6
            // planes (NC = 64 / 128), of 4 for the 8..10-plane boxes of NC = 32
            const int zs = p.box_z <= 6 ? 2 : 4;
            if (p.halo == 2) xform_stage<TT_X + 2, TT_Y + 2, 3>(sa, p.box_z, zlo, zhi, ylo, yhi, xlo, xhi, sc, sh, skip, p.xf.slope, tid, gthreads, zs);
            else if (p.halo == 1) xform_stage<TT_X + 1, TT_Y + 1, 3>(sa, p.box_z, zlo, zhi, ylo, yhi, xlo, xhi, sc, sh, skip, p.xf.slope, tid, gthreads, zs);
            else xform_stage<TT_X, TT_Y, 3>(sa, p.box_z, zlo, zhi, ylo, yhi, xlo, xhi, sc, sh, skip, p.xf.slope, tid, gthreads, zs);
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[st]);
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 4..7)
    reg_alloc<TREGS_EPI>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t tcount = 0;
    RunningStats run[NC / 32];
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      int nt, b, tz, ty, tx;
      taps_decode_tile(tile, p, nt, b, tz, ty, tx);
      const uint32_t buf = tcount & 1;
      mbar_wait(&tfull[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const int x = tx * TT_X + (row & 7), y = ty * TT_Y + (row >> 3);
      const bool rowvalid = (x < p.W) && (y < p.H);
      const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + buf * buf_cols;
#pragma unroll
      for (int chunk = 0; chunk < NC / 32; ++chunk) {
        const int nbase = nt * NC + chunk * 32;  // first GEMM column of this 32-wide strip
        if (p.kind != TAPS_TCONV) {
          const size_t zstride = (size_t)p.H * p.W, gstride = (size_t)p.D * zstride;
          uint4* dst = reinterpret_cast<uint4*>(p.out) +
                       ((((size_t)b * p.out_groups_total + p.out_group_off + (nbase >> 3)) * p.D + tz * p.zt) * p.H + y) * p.W + x;
          conv_epilogue_strip(tlane + chunk * 32, NC, p.zt, p.bias + nbase, rowvalid, tz * p.zt, p.D, dst, zstride,
                              gstride, lane, b * p.Cout + nbase, run[chunk], p.stats,
                              s2d_dst(p.s2d, p.s2d_stride, b, p.Cout / 8, nbase >> 3, p.D, p.H, p.W, y, x));
        } else {
          // transposed conv: GEMM column n = (((pz*tsy+py) * Cout/8 + cg) * 2 + px) * 8 + e : the two x-phases of a
          // channel group are neighbours on N, so a thread writes 32 contiguous bytes (xo = 2x, 2x+1) per group
          const int Do = p.tsz * p.D, Ho = p.tsy * p.H, Wo = 2 * p.W;
          const int cgroups = p.Cout / 8;
          float bs[32];
          uint4* d[2];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int q16 = (nbase >> 4) + jj;                 // 16-column block = (pzpy, channel group)
            const int pzpy = q16 / cgroups, cg = q16 - pzpy * cgroups;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8) + 1);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { bs[16 * jj + e] = bb[e]; bs[16 * jj + 8 + e] = bb[e]; }
            const int zo = p.tsz * (tz * p.zt) + pzpy / p.tsy, yo = p.tsy * y + pzpy % p.tsy;
            d[jj] = reinterpret_cast<uint4*>(p.out) +
                    ((((size_t)b * p.out_groups_total + p.out_group_off + cg) * Do + zo) * Ho + yo) * Wo + 2 * x;
          }
          tconv_epilogue_strip(tlane + chunk * 32, NC, p.zt, bs, rowvalid, tz * p.zt, p.D, d[0], d[1],
                               (size_t)p.tsz * Ho * Wo);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
#pragma unroll
    for (int chunk = 0; chunk < NC / 32; ++chunk) stats_flush(run[chunk], p.stats, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ================================================================================================ host side
struct ConvTapsPlan {
  CUtensorMap tmap;
  TapsParams prm;
  __half* d_bpacked = nullptr;
  float* d_bias = nullptr;
  int32_t* d_table = nullptr;
  int nc = 64;
  size_t smem = 0;
  int grid = 0;
};

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

ConvTapsPlan* conv_taps_plan_create(TapsKind kind, const TapsGeom& geo, const float* h_w, const float* h_bias,
                                    int cin_w, int Cout, const ActView& src, int B, const ConvIO& io, double* d_stats) {
  const ActView& dst = io.out;
  const int cin_padded = (cin_w + 15) / 16 * 16;
  const int* ks = geo.ks;
  const int* st = geo.stride;
  for (int a = 0; a < 3; ++a)
    if ((ks[a] != 1 && ks[a] != 3 && kind == TAPS_CONV) || (st[a] != 1 && st[a] != 2)) {
      set_error("conv_taps: kernel sizes must be 1 or 3 and strides 1 or 2");
      return nullptr;
    }
  const int phases = kind == TAPS_CONV ? st[0] * st[1] * st[2] : 1;
  const int Ntotal = kind == TAPS_TCONV ? st[0] * st[1] * st[2] * Cout : Cout;
  const int src_cin_groups = src.groups / phases;  // groups per phase of a space-to-depth source
  const bool sym = kind == TAPS_TCONV ? (st[2] == 2) : (ks[1] == ks[2] && st[1] == st[2]);
  if (Cout % 8 || Ntotal % 32 || cin_padded / 8 > src_cin_groups || (phases > 1 && cin_w % 16) || !sym ||
      src.groups % phases) {
    set_error("conv_taps: unsupported shape kind=%d cin=%d cout=%d src groups=%d kernel %d%d%d stride %d%d%d", (int)kind,
              cin_w, Cout, src.groups, ks[0], ks[1], ks[2], st[0], st[1], st[2]);
    return nullptr;
  }
  ConvTapsPlan* pl = new ConvTapsPlan();
  const int NC = Ntotal % 64 ? 32 : ((Ntotal % 128 == 0 && (kind == TAPS_TCONV || phases > 1)) ? 128 : 64);
  pl->nc = NC;
  TapsParams& p = pl->prm;
  p.kind = (int)kind;
  p.zt = 256 / NC;  // two TMEM accumulator buffers of zt*NC <= 256 columns
  p.B = B; p.Ntotal = Ntotal; p.Cout = Cout; p.D = src.D; p.H = src.H; p.W = src.W;
  // per axis (z, y, x): halo of the box and where it starts relative to the tile.  Kernel 3: one position before the
  // tile; on a stride-2 axis the space-to-depth phases need one extra position, on a stride-1 axis two.
  int halo[3], org[3];
  for (int a = 0; a < 3; ++a) {
    const bool k3 = kind == TAPS_CONV && ks[a] == 3;
    halo[a] = k3 ? (st[a] == 2 ? 1 : 2) : 0;
    org[a] = k3 ? -1 : 0;
  }
  p.org_z = org[0]; p.org_y = org[1]; p.org_x = org[2];
  p.box_x = TT_X + halo[2]; p.box_y = TT_Y + halo[1]; p.box_z = p.zt + halo[0];
  p.halo = halo[2];  // == halo[1]: the transform is instantiated per in-plane box size
  p.tsz = kind == TAPS_TCONV ? st[0] : 1;
  p.tsy = kind == TAPS_TCONV ? st[1] : 1;
  p.tiles_x = (src.W + TT_X - 1) / TT_X;
  p.tiles_y = (src.H + TT_Y - 1) / TT_Y;
  p.tiles_z = (src.D + p.zt - 1) / p.zt;
  p.n_ntiles = Ntotal / NC;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_z * B * p.n_ntiles;
  p.in_groups_total = src.groups_total; p.in_group_off = src.group_off;
  p.out = dst.base;
  p.out_groups_total = dst.groups_total; p.out_group_off = dst.group_off;
  p.s2d = kind == TAPS_TCONV ? nullptr : io.s2d;
  for (int a = 0; a < 3; ++a) p.s2d_stride[a] = io.s2d_stride[a];
  p.xf = io.xf;
  p.xf_debug = getenv("BOA_B200_XF_DEBUG") ? atoi(getenv("BOA_B200_XF_DEBUG")) : 0;
  if (io.xf.scale && (io.xf.channels != cin_w || cin_w % 16 != 0)) {
    set_error("conv_taps: fused input normalisation needs Cin %% 16 == 0 (cin=%d, scale row %d)", cin_w, io.xf.channels);
    conv_taps_plan_destroy(pl);
    return nullptr;
  }
  p.stats = kind == TAPS_TCONV ? nullptr : d_stats;

  // ---- chunk table + packed weights.  A class = one phase of the source that serves at least one tap.
  const int cpp = cin_padded / 16;  // K chunks per phase (or per tensor)
  const int k3 = ks[0] * ks[1] * ks[2];
  struct Op { int a_off, tap; };
  struct Class { int gbase; std::vector<Op> ops; };
  std::vector<Class> classes;
  if (kind == TAPS_TCONV) {
    classes.push_back({0, {{0, 0}}});
  } else {
    // per axis: (box offset, kernel tap) pairs a phase can serve
    auto opts = [&](int a, int bit, int (&off)[3], int (&d)[3]) {
      if (ks[a] == 1) {
        if (st[a] == 2 && bit) return 0;  // a stride-2 axis with kernel 1 reads the even phase only
        off[0] = 0; d[0] = 0;
        return 1;
      }
      if (st[a] == 1) { for (int i = 0; i < 3; ++i) { off[i] = i; d[i] = i; } return 3; }
      if (bit) { off[0] = 0; d[0] = 0; off[1] = 1; d[1] = 2; return 2; }
      off[0] = 1; d[0] = 1;
      return 1;
    };
    for (int phs = 0; phs < phases; ++phs) {
      const int px = phs % st[2], py = (phs / st[2]) % st[1], pz = phs / (st[2] * st[1]);
      int oz[3], dzv[3], oy[3], dyv[3], ox[3], dxv[3];
      const int nz = opts(0, pz, oz, dzv), ny = opts(1, py, oy, dyv), nx = opts(2, px, ox, dxv);
      Class c;
      c.gbase = phs * src_cin_groups;
      for (int a = 0; a < nz; ++a)
        for (int bq = 0; bq < ny; ++bq)
          for (int cx = 0; cx < nx; ++cx)
            c.ops.push_back({(oz[a] * p.box_y + oy[bq]) * p.box_x + ox[cx], (dzv[a] * ks[1] + dyv[bq]) * ks[2] + dxv[cx]});
      if (!c.ops.empty()) classes.push_back(c);
    }
  }
  p.n_classes = (int)classes.size();
  p.chunks_per_class = cpp;
  p.kc_count = p.n_classes * cpp;
  std::vector<int32_t> table((size_t)p.n_classes * TAB_STRIDE, 0);
  int max_ops = 0;
  size_t total_ops = 0;
  for (int c = 0; c < p.n_classes; ++c) {
    const std::vector<Op>& o = classes[c].ops;
    table[(size_t)c * TAB_STRIDE] = (int)o.size();
    table[(size_t)c * TAB_STRIDE + 1] = (int)total_ops;
    for (size_t i = 0; i < o.size(); ++i) table[(size_t)c * TAB_STRIDE + 2 + i] = o[i].a_off;
    table[(size_t)c * TAB_STRIDE + TAB_GBASE] = classes[c].gbase;
    max_ops = std::max(max_ops, (int)o.size());
    total_ops += o.size() * cpp;
  }
  const size_t nt_halves = total_ops * NC * 16;  // [op][kchunk 2][NC rows][8]
  p.b_nt_bytes = (uint32_t)(nt_halves * 2);
  std::vector<__half> hb((size_t)p.n_ntiles * nt_halves);
  const int nph = st[0] * st[1] * st[2];
  for (int nt = 0; nt < p.n_ntiles; ++nt) {
    size_t opi = 0;
    for (int kc = 0; kc < p.kc_count; ++kc)
      for (const Op& op : classes[kc / cpp].ops) {
        __half* blk = hb.data() + (size_t)nt * nt_halves + opi * NC * 16;
        ++opi;
        const int cc = kc % cpp;
        for (int kch = 0; kch < 2; ++kch)
          for (int n = 0; n < NC; ++n)
            for (int e = 0; e < 8; ++e) {
              const int ci = cc * 16 + kch * 8 + e;
              const int ng = nt * NC + n;
              float v = 0.f;
              if (ci < cin_w) {
                if (kind == TAPS_TCONV) {
                  const int px = (ng >> 3) & 1, cg = (ng >> 4) % (Cout / 8), pzpy = ng / (2 * Cout);
                  const int pz = pzpy / st[1], py = pzpy % st[1], co = cg * 8 + (ng & 7);
                  v = h_w[((size_t)ci * Cout + co) * nph + (pz * st[1] + py) * st[2] + px];
                } else {
                  v = h_w[((size_t)ng * cin_w + ci) * k3 + op.tap];
                }
              }
              blk[((size_t)kch * NC + n) * 8 + e] = __float2half_rn(v);
            }
      }
  }
  if (cudaMalloc(&pl->d_bpacked, hb.size() * 2) != cudaSuccess ||
      cudaMalloc(&pl->d_bias, Cout * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&pl->d_table, table.size() * sizeof(int32_t)) != cudaSuccess) {
    set_error("conv_taps: cudaMalloc failed");
    conv_taps_plan_destroy(pl);
    return nullptr;
  }
  cudaMemcpy(pl->d_bpacked, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(pl->d_bias, h_bias, Cout * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(pl->d_table, table.data(), table.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  p.bpacked = pl->d_bpacked; p.bias = pl->d_bias; p.table = pl->d_table;

  if (make_c8_tmap(&pl->tmap, src.base, B * src.groups_total, src.D, src.H, src.W, p.box_x, p.box_y, p.box_z, 2)) {
    set_error("conv_taps: tensor map creation failed");
    conv_taps_plan_destroy(pl);
    return nullptr;
  }
  p.a_tx_bytes = 2u * p.box_z * p.box_y * p.box_x * 16u;
  p.a_bytes = round_up(p.a_tx_bytes, 128u);
  p.b_stage_bytes = round_up((uint32_t)max_ops * NC * 32u, 128u);
  const uint32_t stage = p.a_bytes + p.b_stage_bytes;
  int stages = (int)(200u * 1024u / stage);
  if (stages > TAPS_MAX_STAGES) stages = TAPS_MAX_STAGES;
  if (const char* e = getenv("BOA_B200_TAPS_STAGES")) stages = std::max(2, std::min(stages, atoi(e)));
  if (stages < 2) {
    set_error("conv_taps: stage of %u bytes does not fit twice in shared memory", stage);
    conv_taps_plan_destroy(pl);
    return nullptr;
  }
  if (io.xf.scale && stages >= 4 && !getenv("BOA_B200_TAPS_NOTRIM")) stages = stages / 4 * 4;  // the transform groups share the ring slots evenly
  p.stages = stages;
  p.xf_groups = stages % 4 == 0 ? 4 : (stages % 2 == 0 ? 2 : 1);
  p.tmap_merged = c8_tmap_merged() ? 1 : 0;
  pl->smem = (size_t)stages * stage + (3 * TAPS_MAX_STAGES + 8) * 8 + 8 * TAB_STRIDE * sizeof(int32_t);
  // the attribute is per kernel, not per plan: always opt in to the full 227 KB
  cudaError_t e = NC == 128 ? cudaFuncSetAttribute(conv_taps_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM)
                  : NC == 64 ? cudaFuncSetAttribute(conv_taps_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM)
                             : cudaFuncSetAttribute(conv_taps_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
  if (e != cudaSuccess) {
    set_error("conv_taps: cannot opt in to %zu bytes of shared memory: %s", pl->smem, cudaGetErrorString(e));
    conv_taps_plan_destroy(pl);
    return nullptr;
  }
  pl->grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  return pl;
}

void conv_taps_plan_destroy(ConvTapsPlan* p) {
  if (!p) return;
  if (p->d_bpacked) cudaFree(p->d_bpacked);
  if (p->d_bias) cudaFree(p->d_bias);
  if (p->d_table) cudaFree(p->d_table);
  delete p;
}

int conv_taps_launch(ConvTapsPlan* pl, cudaStream_t s, int nb) {
  if (nb > 0 && nb != pl->prm.B) {  // partial batch: only the tiles of the first nb items (item index is decoded mod B)
    TapsParams& p = pl->prm;
    p.total_tiles = p.total_tiles / p.B * nb;
    p.B = nb;
    pl->grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  }
  if (pl->nc == 128)
    conv_taps_kernel<128><<<pl->grid, TAPS_THREADS, pl->smem, s>>>(pl->tmap, pl->prm);
  else if (pl->nc == 64)
    conv_taps_kernel<64><<<pl->grid, TAPS_THREADS, pl->smem, s>>>(pl->tmap, pl->prm);
  else
    conv_taps_kernel<32><<<pl->grid, TAPS_THREADS, pl->smem, s>>>(pl->tmap, pl->prm);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

}  // namespace boa
