// 3x3x3 stride-1 Conv3d as a tcgen05 implicit GEMM for sm_100a.
//
// Replaces the cuDNN call behind `nn.Conv3d` of dynamic_network_architectures' ConvDropoutNormReLU, invoked at
// _external/nnunetv2/inference/predict_from_raw_data.py:543.
//
// Design (measured basis: profiles/r01_probe_tcgen05.log):
//   * SS-mode tcgen05.mma M=128,K=16 costs max(N/2, 32 + N/4) cycles - the A operand (4 KB/MMA) saturates the
//     128 B/cycle shared-memory port, so N = C_out = 32 would cap at 40 % of the tensor peak.  We therefore FOLD the
//     three dz taps into N: one MMA computes, for one input z-plane ("slab"), the contributions to the three output
//     planes z-1, z, z+1 at once (N = 3*NC: 86 % at NC=32, 100 % at NC=64).  The three N-blocks land in three
//     consecutive TMEM column slots (slot = output plane), so the tap sum happens by accumulation in TMEM.
//   * A operand: a TMA 5-D box (8ch, 10x, 18y, zt+2 planes, 2 channel groups) of the C8 activation tensor, zero
//     filled out of bounds (= conv padding).  The (dy,dx) taps are 16-byte address offsets of the same smem tile
//     (no-swizzle K-major descriptors; M rows = 8 x-voxels x 16 y-rows, SBO = one halo row).
//   * B operand: weights pre-packed on the host into the exact smem image, one bulk copy per K chunk.
//   * Persistent CTAs of three warpgroups (register file re-partitioned with setmaxnreg): WG0 = TMA producer (warp 0) +
//     MMA issuer (warp 1, one elected lane), WG1 = epilogue (TMEM -> regs, +bias, InstanceNorm sum / sum^2, fp16 C8
//     store, optional space-to-depth copy), WG2 = operand transform: the input is the producer's RAW conv output and
//     its InstanceNorm affine + LeakyReLU are applied to the tile in shared memory between the TMA write and the MMAs
//     (conv_xform.cuh) - there is no standalone normalise pass.  Shared-memory ring over K chunks of 16 input
//     channels; TMEM accumulators double-buffered when zt*NC <= 256.
#include <stdlib.h>
#include <vector>
#include "net_kernels.cuh"
#include "conv_epilogue.cuh"
#include "conv_xform.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace boa {

// WG0: producer warp + MMA warp (+2 idle), WG1: 4 epilogue warps, WG2 + WG3: 8 transform warps.  512 threads start with
// 128 registers each and re-partition them with setmaxnreg.  (Eight transform warps at 64 registers spilled the scale /
// shift arrays to local memory inside the element loop - ncu source view, profiles/r02_taps_fused_stalls.txt; four
// warps at 128 registers did not spill but had too little issue bandwidth.)
// Measured (B200, batch 8, tools/perf_probe.py): the transform is bound by instruction issue of its own warps (~36 ALU
// instructions per 16-byte element, 1.8x the tensor's elements because of the halo), so it wants as many warps as the
// register file allows: 16 warps per CTA = 128 * (96 + 224 + 96 + 96) = 65536 registers.
constexpr int MMA_THREADS = 512;
constexpr int REGS_WG0 = 96, REGS_EPI = 224, REGS_XF = 96, XF_THREADS = 256;
// The transform warps work in XF_GROUPS independent groups; group g owns the ring SLOTS s with s % XF_GROUPS == g, so
// that the latency chain of one stage (barrier wait, scale / shift loads, LDS -> math -> STS, proxy fence, arrive)
// overlaps the next stage's instead of serialising with it.  (Ownership by slot, not by stage number: a group then
// waits for the phases of a slot in order - a group that could reach phase k + 1 of a slot before phase k has
// completed would see the parity wait succeed at once.)
// The number of groups divides the ring depth (every group gets the same share of the stages): p.xf_groups.
constexpr int TILE_X = 8, TILE_Y = 16;
constexpr int XB = TILE_X + 2, YB = TILE_Y + 2, SLAB = XB * YB;  // 180 halo positions per z-plane

struct ConvMmaParams {
  const __half* bpacked;
  const float* bias;
  __half* out;
  double* stats;
  int B, kc_count, Cout, D, H, W, zt;
  int tiles_x, tiles_y, tiles_z, n_ntiles, total_tiles;
  int in_groups_total, in_group_off;
  int out_groups_total, out_group_off;  // the output view (a channel-group slice of a concat buffer, or dense)
  __half* s2d;                          // optional space-to-depth copy of the output (nullptr: none)
  int s2d_stride[3];                    // strides (z, y, x) of the conv that will read the copy
  InXform xf;                           // fused normalisation of the input (xf.scale == nullptr: none)
  int xf_groups;                        // transform warp groups (divides `stages`)
  int xf_debug;                         // profiling aid (BOA_B200_XF_DEBUG): 1 = transform warps only relay the barrier
  int stages;      // shared-memory ring depth of the A operand (2..4)
  int b_resident;  // 1: the weights of ALL K chunks stay in shared memory for the whole kernel (n_ntiles == 1), the
                   //    ring carries only the activation tiles
  int tmap_merged;  // tensor map built with the (channel, x) dimensions merged (tmap.cuh)
  int ntaps;  // 9: (dy,dx) taps as address shifts; 1: the in-plane taps already sit on K (first layer), centre only
};

__device__ __forceinline__ void decode_tile(int t, const ConvMmaParams& p, int& nt, int& b, int& tz, int& ty,
                                            int& tx) {
  tx = t % p.tiles_x; t /= p.tiles_x;
  ty = t % p.tiles_y; t /= p.tiles_y;
  tz = t % p.tiles_z; t /= p.tiles_z;
  b = t % p.B;
  nt = t / p.B;
}

// MMAs of one K chunk (16 input channels) of one tile: input z-plane i of the halo box feeds output slots i-2 .. i
// with the weight blocks dz = 2, 1, 0 (one MMA, N = 3*NC, three consecutive TMEM slots); the planes at the box ends
// feed fewer slots.  ZT > 0: the z-tile height is a compile-time constant and the plane loop is fully unrolled, so
// every descriptor / column / instruction-descriptor offset is an immediate - the single issuing thread was spending
// ~40 % of its time on the dependent uniform-datapath arithmetic of the runtime version (ncu source view,
// profiles/r01_fold_issue.txt).  ZT == 0: runtime z-tile (small volumes).
template <int NC, int ZT>
__device__ __forceinline__ void issue_chunk(uint64_t a_base, uint64_t b_base, uint32_t dcol0, bool first_kc, bool taps9,
                                            int zt_rt) {
  constexpr uint32_t BG16 = 96u * NC / 16u;  // one (dy,dx) weight block, in 16-byte units
  const int zt = ZT > 0 ? ZT : zt_rt;
#pragma unroll
  for (int i = 0; i < (ZT > 0 ? ZT + 2 : 18); ++i) {
    if (ZT == 0 && i >= zt + 2) break;
    const int lo = i - 2 > 0 ? i - 2 : 0, hi = i < zt - 1 ? i : zt - 1;
    const int jlo = lo - (i - 2);         // first valid dz block (blocks are ordered dz = 2,1,0)
    const int n = hi - lo + 1;
    const uint64_t ad = a_base + (uint64_t)(i * SLAB);
    const uint64_t bd = b_base + (uint64_t)(jlo * NC);
    const uint32_t dcol = dcol0 + (uint32_t)(lo * NC);
    const uint32_t idesc = umma_idesc_f16(128, (uint32_t)(n * NC));
    if (first_kc && i <= zt - 1) {
      // slot i (= hi) receives its first contribution now: overwrite it, accumulate into the slots below
      if (n > 1) umma_f16(dcol, ad, bd, umma_idesc_f16(128, (uint32_t)((n - 1) * NC)), 1u);
      umma_f16(dcol0 + (uint32_t)(hi * NC), ad, bd + (uint64_t)((n - 1) * NC), umma_idesc_f16(128, NC), 0u);
    } else {
      umma_f16(dcol, ad, bd, idesc, 1u);
    }
    if (taps9) {
#pragma unroll
      for (int g = 1; g < 9; ++g)
        umma_f16(dcol, ad + (uint64_t)((g / 3) * XB + (g % 3)), bd + (uint64_t)(g * BG16), idesc, 1u);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(MMA_THREADS, 1)
conv3_fold_kernel(const __grid_constant__ CUtensorMap tmapA, const ConvMmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int zb = p.zt + 2;
  const uint32_t a_bytes = 2u * zb * SLAB * 16u;
  constexpr uint32_t b_group = 96u * NC;  // [2 kchunks][3*NC rows][16 B]
  const uint32_t b_bytes = (uint32_t)p.ntaps * b_group;
  // smem: [resident weights: kc_count * b_bytes (b_resident only)] [stages x (A tile [+ B chunk])] [barriers]
  const uint32_t stage_bytes = a_bytes + (p.b_resident ? 0u : b_bytes);
  const uint32_t bres_bytes = p.b_resident ? (uint32_t)p.kc_count * b_bytes : 0u;
  uint8_t* const ring = smem + bres_bytes;
  const int nstage = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)nstage * stage_bytes);
  uint64_t* full = bars;         // [4] operand ready (TMA, or the transform warps when the input is normalised here) -> MMA
  uint64_t* empty = bars + 4;    // [4] MMA -> TMA
  uint64_t* tfull = bars + 8;    // [2] MMA -> epilogue
  uint64_t* tempty = bars + 10;  // [2] epilogue -> MMA
  uint64_t* rawfull = bars + 12; // [4] TMA -> transform warps
  uint64_t* bfull = bars + 16;   // [1] resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbuf = (p.zt * NC <= 256) ? 2 : 1;
  const bool xform = p.xf.scale != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full[i], xform ? XF_THREADS / 32 / p.xf_groups : 1);
      mbar_init(&empty[i], 1);
      mbar_init(&rawfull[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(bfull, 1);
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

  if (warp < 4) {
    reg_dealloc<REGS_WG0>();
    if (warp == 0) {
      // ===================================================================== TMA producer
      if (p.b_resident && elect_one()) {  // n_ntiles == 1: one weight set for every tile of this CTA
        mbar_arrive_expect_tx(bfull, bres_bytes);
        for (int kc = 0; kc < p.kc_count; ++kc)
          bulk_load(smem + (size_t)kc * b_bytes, reinterpret_cast<const uint8_t*>(p.bpacked) + (size_t)kc * b_bytes,
                    b_bytes, bfull);
      }
      __syncwarp();
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, b, tz, ty, tx;
        decode_tile(tile, p, nt, b, tz, ty, tx);
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(&empty[st], ph ^ 1);
          if (elect_one()) {
            uint8_t* sa = ring + (size_t)st * stage_bytes;
            uint64_t* ready = xform ? &rawfull[st] : &full[st];
            mbar_arrive_expect_tx(ready, stage_bytes);
            tma_load_c8(sa, &tmapA, ready, p.tmap_merged, tx * TILE_X - 1, ty * TILE_Y - 1, tz * p.zt - 1,
                        b * p.in_groups_total + p.in_group_off + 2 * kc);
            if (!p.b_resident)
              bulk_load(sa + a_bytes,
                        reinterpret_cast<const uint8_t*>(p.bpacked) + (size_t)(nt * p.kc_count + kc) * b_bytes, b_bytes,
                        ready);
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
        if (p.b_resident) {
          mbar_wait(bfull, 0);
          tc_fence_after();
        }
        const uint32_t a_lbo = (uint32_t)zb * SLAB * 16u;  // next channel group (K chunk of 8)
        // descriptors with a zero address field; tap / slab / row-block shifts are added to the low word (16-byte
        // units; shared memory is < 256 KB so the 14-bit address field never carries)
        const uint64_t a_desc0 = umma_desc(0, a_lbo, XB * 16u);
        const uint64_t b_desc0 = umma_desc(0, 3u * NC * 16u, 128u);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
          const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0;
          const uint32_t tph = nbuf == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
          mbar_wait(&tempty[buf], tph ^ 1);
          tc_fence_after();
          const uint32_t dcol0 = tbase + buf * (uint32_t)(p.zt * NC);
          for (int kc = 0; kc < p.kc_count; ++kc) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(ring + (size_t)st * stage_bytes);
            const uint64_t a_base = a_desc0 + (uint64_t)(a0 >> 4);
            const uint64_t b_base =
                b_desc0 + (uint64_t)((p.b_resident ? smem_u32(smem) + (uint32_t)kc * b_bytes : a0 + a_bytes) >> 4);
            const uint64_t a_tap0 = a_base + (uint64_t)(p.ntaps == 1 ? XB + 1 : 0);
            if (p.zt == 8) issue_chunk<NC, 8>(a_tap0, b_base, dcol0, kc == 0, p.ntaps == 9, 8);
            else if (p.zt == 4) issue_chunk<NC, 4>(a_tap0, b_base, dcol0, kc == 0, p.ntaps == 9, 4);
            else issue_chunk<NC, 0>(a_tap0, b_base, dcol0, kc == 0, p.ntaps == 9, p.zt);
            umma_commit(&empty[st]);   // smem stage reusable once these MMAs retire
            if (++st == nstage) { st = 0; ph ^= 1; }
          }
          umma_commit(&tfull[buf]);     // accumulators of this tile complete
        }
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================================================================== epilogue (warps 4..7)
    reg_alloc<REGS_EPI>();
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    uint32_t tcount = 0;
    RunningStats run[NC / 32];
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      int nt, b, tz, ty, tx;
      decode_tile(tile, p, nt, b, tz, ty, tx);
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0;
      const uint32_t tph = nbuf == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
      mbar_wait(&tfull[buf], tph);
      tc_fence_after();
      const int x = tx * TILE_X + (row & 7), y = ty * TILE_Y + (row >> 3);
      const bool rowvalid = (x < p.W) && (y < p.H);
      const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(p.zt * NC);
      const size_t zstride = (size_t)p.H * p.W, gstride = (size_t)p.D * zstride;
#pragma unroll
      for (int chunk = 0; chunk < NC / 32; ++chunk) {
        const int cbase = nt * NC + chunk * 32;
        uint4* dst = reinterpret_cast<uint4*>(p.out) +
                     ((((size_t)b * p.out_groups_total + p.out_group_off + (cbase >> 3)) * p.D + tz * p.zt) * p.H + y) * p.W + x;
        conv_epilogue_strip(tlane + chunk * 32, NC, p.zt, p.bias + cbase, rowvalid, tz * p.zt, p.D, dst, zstride, gstride,
                            lane, b * p.Cout + cbase, run[chunk], p.stats,
                            s2d_dst(p.s2d, p.s2d_stride, b, p.Cout / 8, cbase >> 3, p.D, p.H, p.W, y, x));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
#pragma unroll
    for (int chunk = 0; chunk < NC / 32; ++chunk) stats_flush(run[chunk], p.stats, lane);
  } else {
    // ===================================================================== operand transform (warps 8..15)
    reg_dealloc<REGS_XF>();
    if (xform) {
      const int gthreads = XF_THREADS / p.xf_groups;
      const int grp = (threadIdx.x - 256) / gthreads;
      const int tid = (threadIdx.x - 256) % gthreads;
      uint32_t cnt = 0;  // ring stage counter over (tile, kc) - the same sequence the producer and the MMA warp walk
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, b, tz, ty, tx;
        decode_tile(tile, p, nt, b, tz, ty, tx);
        // box index i <-> volume coordinate t * T - 1 + i: the in-volume part of the box
        const int z0 = tz * p.zt - 1, y0 = ty * TILE_Y - 1, x0 = tx * TILE_X - 1;
        const int zlo = z0 < 0 ? -z0 : 0, zhi = p.D - z0 < zb ? p.D - z0 : zb;
        const int ylo = y0 < 0 ? -y0 : 0, yhi = p.H - y0 < YB ? p.H - y0 : YB;
        const int xlo = x0 < 0 ? -x0 : 0, xhi = p.W - x0 < XB ? p.W - x0 : XB;
        for (int kc = 0; kc < p.kc_count; ++kc, ++cnt) {
          const int st = (int)(cnt % (uint32_t)nstage);
          if (st % p.xf_groups != grp) continue;
          const uint32_t ph = (cnt / (uint32_t)nstage) & 1u;
          const int g0 = 2 * kc;  // first channel group of this K chunk inside the input view
          const int skip = (g0 < p.xf.ident_groups ? 1 : 0) | (g0 + 1 < p.xf.ident_groups ? 2 : 0) |
                           (8 * g0 + 8 >= p.xf.channels ? 2 : 0);
          mbar_wait(&rawfull[st], ph);
          if (skip != 3 && p.xf_debug != 1)
            xform_stage<XB, YB, 3>(ring + (size_t)st * stage_bytes, zb, zlo, zhi, ylo, yhi, xlo, xhi,
                                      p.xf.scale + (size_t)b * p.xf.channels + 16 * kc,
                                      p.xf.shift + (size_t)b * p.xf.channels + 16 * kc, skip, p.xf.slope, tid, gthreads, 4);
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[st]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ================================================================================================ host side
struct ConvMmaPlan {
  CUtensorMap tmap;
  ConvMmaParams prm;
  __half* d_bpacked = nullptr;
  float* d_bias = nullptr;
  int nc = 32;
  size_t smem = 0;
  int grid = 0;
  double macs = 0;
};

ConvMmaPlan* conv_mma_plan_create(const float* h_w, const float* h_bias, int cin_w, int cin_padded, int Cout,
                                  const ActView& src, int B, const ConvIO& io, double* d_stats, bool taps_on_k,
                                  int kz) {
  // taps_on_k (first layer, Cin = 1): the source tensor carries the 9 in-plane neighbours of every voxel as its
  // channels 0..8, h_w is [Cout][1][kz * 9]; only the dz taps remain as (folded) taps.  kz == 1 ([1,3,3] kernels of
  // anisotropic plans): the dz = 0 / 2 blocks are zero (wasted MMAs in a layer that is HBM bound anyway).
  const int ntaps = taps_on_k ? 1 : 9;
  if (taps_on_k) { cin_w = 9; cin_padded = 16; }
  if (cin_padded % 16 || Cout % 32 || cin_padded > src.groups * 8) {
    set_error("conv_mma: unsupported channels cin=%d(padded %d) cout=%d src groups=%d", cin_w, cin_padded, Cout,
              src.groups);
    return nullptr;
  }
  ConvMmaPlan* pl = new ConvMmaPlan();
  // NC = 64 reaches the full MMA rate (N = 192); small volumes (deep stages) cannot fill the machine with tiles, so
  // they take NC = 32 (twice the CTAs, shorter per-tile chains) and a z-tile no taller than the volume.
  const int zt = src.D < 8 ? src.D : 8;
  const int tiles64 = ((src.W + TILE_X - 1) / TILE_X) * ((src.H + TILE_Y - 1) / TILE_Y) * ((src.D + zt - 1) / zt) * B *
                      (Cout / 64 > 0 ? Cout / 64 : 1);
  const int NC = (Cout % 64 == 0 && tiles64 >= sm_count()) ? 64 : 32;
  pl->nc = NC;
  ConvMmaParams& p = pl->prm;
  p.B = B; p.kc_count = cin_padded / 16; p.Cout = Cout; p.D = src.D; p.H = src.H; p.W = src.W; p.zt = zt;
  p.tiles_x = (src.W + TILE_X - 1) / TILE_X;
  p.tiles_y = (src.H + TILE_Y - 1) / TILE_Y;
  p.tiles_z = (src.D + zt - 1) / zt;
  p.n_ntiles = Cout / NC;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_z * B * p.n_ntiles;
  p.in_groups_total = src.groups_total; p.in_group_off = src.group_off;
  p.out = io.out.base; p.out_groups_total = io.out.groups_total; p.out_group_off = io.out.group_off;
  p.s2d = io.s2d;
  for (int a = 0; a < 3; ++a) p.s2d_stride[a] = io.s2d_stride[a];
  p.stats = d_stats;
  p.ntaps = ntaps;
  p.tmap_merged = c8_tmap_merged() ? 1 : 0;
  p.xf = io.xf;
  p.xf_debug = getenv("BOA_B200_XF_DEBUG") ? atoi(getenv("BOA_B200_XF_DEBUG")) : 0;
  if (io.out.groups < Cout / 8 ||
      (io.s2d && (src.D % io.s2d_stride[0] || src.H % io.s2d_stride[1] || src.W % io.s2d_stride[2])) ||
      (io.xf.scale && (taps_on_k || io.xf.channels != cin_w || cin_w % 16 != 0))) {
    set_error("conv_mma: bad output view / space-to-depth copy / input transform for cin=%d cout=%d", cin_w, Cout);
    conv_mma_plan_destroy(pl);
    return nullptr;
  }
  pl->macs = 27.0 * (taps_on_k ? 1 : cin_w) * Cout * (double)src.voxels() * B;

  // pack weights: [nt][kc][g=(dy,dx)][kchunk][n = j*NC + co, j <-> dz = 2-j][8 cin]
  const size_t b_group = 96 * (size_t)NC / 2;  // halves
  std::vector<__half> hb((size_t)p.n_ntiles * p.kc_count * ntaps * b_group);
  for (int nt = 0; nt < p.n_ntiles; ++nt)
    for (int kc = 0; kc < p.kc_count; ++kc)
      for (int g = 0; g < ntaps; ++g) {
        const int dy = g / 3, dx = g % 3;
        __half* blk = hb.data() + (((size_t)nt * p.kc_count + kc) * ntaps + g) * b_group;
        for (int kch = 0; kch < 2; ++kch)
          for (int j = 0; j < 3; ++j)
            for (int co = 0; co < NC; ++co)
              for (int e = 0; e < 8; ++e) {
                const int ci = kc * 16 + kch * 8 + e;
                const int dz = 2 - j;
                float v = 0.f;
                if (ci < cin_w) {
                  const int tap9 = ci == 0 ? 4 : (ci <= 4 ? ci - 1 : ci);
                  if (!taps_on_k) v = h_w[((size_t)(nt * NC + co) * cin_w + ci) * 27 + dz * 9 + dy * 3 + dx];
                  else if (kz == 3) v = h_w[(size_t)(nt * NC + co) * 27 + dz * 9 + tap9];
                  else v = dz == 1 ? h_w[(size_t)(nt * NC + co) * 9 + tap9] : 0.f;
                }
                blk[((size_t)kch * 3 * NC + j * NC + co) * 8 + e] = __float2half_rn(v);
              }
      }
  if (cudaMalloc(&pl->d_bpacked, hb.size() * 2) != cudaSuccess ||
      cudaMalloc(&pl->d_bias, Cout * sizeof(float)) != cudaSuccess) {
    set_error("conv_mma: cudaMalloc failed");
    conv_mma_plan_destroy(pl);
    return nullptr;
  }
  cudaMemcpy(pl->d_bpacked, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(pl->d_bias, h_bias, Cout * sizeof(float), cudaMemcpyHostToDevice);
  p.bpacked = pl->d_bpacked; p.bias = pl->d_bias;

  if (make_c8_tmap(&pl->tmap, src.base, B * src.groups_total, src.D, src.H, src.W, XB, YB, zt + 2, 2)) {
    set_error("conv_mma: tensor map creation failed");
    conv_mma_plan_destroy(pl);
    return nullptr;
  }
  // Weights resident when one weight set serves every tile (Cout == NC) and it fits next to >= 2 activation stages:
  // the ring then carries 58 KB instead of 85 KB per K chunk (the convs are bound by what TMA can bring into an SM)
  // and, for Cin = 32, a third stage fits.  BOA_B200_BRES=0 disables it.
  const size_t a_stage = 2 * (size_t)(zt + 2) * SLAB * 16, b_chunk = (size_t)ntaps * 96 * (size_t)NC;
  const size_t smem_cap = (size_t)MAX_DYN_SMEM - 1024;  // barriers + the loaders' scale / shift slots
  const char* bres_env = getenv("BOA_B200_BRES");
  p.b_resident = (p.n_ntiles == 1 && !(bres_env && atoi(bres_env) == 0) &&
                  p.kc_count * b_chunk + 2 * a_stage <= smem_cap) ? 1 : 0;
  const size_t stage = a_stage + (p.b_resident ? 0 : b_chunk);
  const size_t fixed = p.b_resident ? p.kc_count * b_chunk : 0;
  int stages = (int)((smem_cap - fixed) / stage);
  stages = stages > 4 ? 4 : stages;
  if (stages < 2) {
    set_error("conv_mma: two stages of %zu bytes do not fit in shared memory", stage);
    conv_mma_plan_destroy(pl);
    return nullptr;
  }
  p.stages = stages;
  p.xf_groups = stages % 2 == 0 ? 2 : 1;
  pl->smem = fixed + (size_t)stages * stage + 768;
  cudaError_t e = cudaSuccess;
  for (const void* fn : {(const void*)conv3_fold_kernel<64>, (const void*)conv3_fold_kernel<32>}) {
    cudaError_t e2 = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
    if (e2 != cudaSuccess) e = e2;
  }
  if (e != cudaSuccess) {
    set_error("conv_mma: cannot opt in to %zu bytes of shared memory: %s", pl->smem, cudaGetErrorString(e));
    conv_mma_plan_destroy(pl);
    return nullptr;
  }
  pl->grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  return pl;
}

void conv_mma_plan_destroy(ConvMmaPlan* p) {
  if (!p) return;
  if (p->d_bpacked) cudaFree(p->d_bpacked);
  if (p->d_bias) cudaFree(p->d_bias);
  delete p;
}

double conv_mma_plan_macs(const ConvMmaPlan* p) { return p->macs; }

int conv_mma_launch(ConvMmaPlan* pl, cudaStream_t s, int nb) {
  if (nb > 0 && nb != pl->prm.B) {  // partial batch: only the tiles of the first nb items (item index is decoded mod B)
    ConvMmaParams& p = pl->prm;
    p.total_tiles = p.total_tiles / p.B * nb;
    p.B = nb;
    pl->grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  }
  if (pl->nc == 64)
    conv3_fold_kernel<64><<<pl->grid, MMA_THREADS, pl->smem, s>>>(pl->tmap, pl->prm);
  else
    conv3_fold_kernel<32><<<pl->grid, MMA_THREADS, pl->smem, s>>>(pl->tmap, pl->prm);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

}  // namespace boa
