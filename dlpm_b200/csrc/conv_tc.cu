// K5: implicit-GEMM 3x3 / 1x1 convolution on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA), replacing the cuDNN calls behind unet.py:64,96,143,157,164-168,347,435
// (nn.Conv2d in ResBlock / Upsample / Downsample / AttentionBlock of dlpm/models/unet.py).
//
// GEMM view:  D[M = B*H*W pixels, N = C_out] = sum_k A[M, k] * Wt[N, k],  k = (tap, c_in) [+ 1x1 skip-conv channels]
//   * activations are NHWC bf16; an M tile is 128 pixels = a (Wb x Hb x Nb) box of the 4-D tensor
//     [B, H, W, C].  For filter tap (dy, dx) the A tile is the SAME box shifted by (dy, dx): one
//     cp.async.bulk.tensor.4d per (tap, 64-channel block), zero padding comes from TMA out-of-bounds
//     fill, stride-2 convolutions use the tensor map's element strides.  No im2col buffer exists.
//   * the TMA writes rows of 128 B (64 bf16 channels) in 128B-swizzled 8-row atoms = the canonical
//     K-major UMMA operand layout, so the MMA consumes the tile as it lands.
//   * ResBlock fusion: the 1x1 skip convolution (unet.py:161-168) is appended to the K loop of the
//     block's second 3x3 conv (extra K blocks reading the raw block input through tmS0/tmS1 = the two
//     halves of the skip concatenation, unet.py:489) and identity skips are added in the epilogue,
//     so `skip_connection(x) + h` (unet.py:195) costs no extra pass.
//   * warp-specialised, persistent: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
//     warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> +bias (+residual) -> bf16 -> global).
//     Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
//   * GroupNorm in the epilogue (GNE, template argument): where the accumulator stage holds WHOLE samples (16x16 maps as CTA
//     pairs; 8x8 / 4x4 maps inside one tile) the consumer's GroupNorm (+ scale-shift, SiLU; unet.py:141,153,188-191) is applied
//     straight from TMEM and the normalised rows go to the consumer's input tensor -- no k_gn_apply pass (see the kernel).
//   * thin fp32 output conv (unet.py:435): horizontal taps stacked along N (ConvGeom::n_par == 3), partials of neighbouring
//     pixels summed by lane shuffles in the epilogue.
#include "../../include/dlpm_b200_unet.h"
#include <cstdlib>
#include <string>

#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace dlpm {

using namespace tc;

constexpr int kMaxStages = 8;
constexpr int kConvThreads = 256;
constexpr int kConvThreadsEpi2 = 384;  // + warps 8-11: a second epilogue warp per TMEM lane quadrant (they split the tile's column chunks, or --
                                       // one-chunk tiles, N <= 32 -- the two sub-tiles of the work item)
constexpr int kConvThreadsXF = 512;  // + warps 8-15: the eight transform warps (two per SM sub-partition)
constexpr int kXfWarps = 8;
constexpr int kConvThreadsPost = 512;  // POST kernels: warps 12-15 normalise finished samples ("GroupNorm in the producer's tail")
constexpr int kPostThreads = 128;
constexpr int kPostSmemBytes = 4 * 8 + 16 + 64 * 2 * 4;  // ready[2] / free[2] barriers + per-quad (sum, sum of squares) of one sample
constexpr int kSmemBudget = 204 * 1024;  // operand ring; epilogue staging, barriers, bias and alignment slack come on top (227 KB per CTA)
constexpr int kEpiStageBytes = 2048;     // per epilogue warp: one 32-pixel x 32-channel bf16 chunk (64-byte rows, SWIZZLE_64B) for TMA stores
constexpr int kEpiStageTotal = 8 * kEpiStageBytes;
constexpr int kBiasSmemFloats = 1024;    // the layer's bias vector lives in shared memory: with the whole carve-out given to shared
                                         // memory there is no L1 left and every bias load of every chunk went to L2
constexpr int kSmemExtra = 1024 /*base alignment*/ + 1024 /*staging alignment*/ + kEpiStageTotal + (4 * kMaxStages + 4) * 8 + 16 +
                           kBiasSmemFloats * 4 + kPostSmemBytes;
// GNE kernels ("GroupNorm in the epilogue", below): bias | per-target coefficient tables | exchanged statistics live in the
// bias region + the POST region (unused there) + the staging-alignment slack (their ring is a multiple of 1024 bytes, so the
// slack sits at the END of the allocation, right behind the POST region)
constexpr int kGneRegionBytes = kBiasSmemFloats * 4 + kPostSmemBytes + 1024;

// POST: one GroupNorm (+ scale-shift, + SiLU) that consumes this convolution's output, evaluated by the convolution itself.
// dst is the consumer's NORMALISED input tensor (NHWC bf16, dst_C channels per pixel); this convolution's channel c lands at
// dst channel c_off + c (c_off > 0: second half of a skip concatenation).  cpg = channels per group of the CONSUMER's
// GroupNorm ((C0 + C1) / 32); gamma / beta / ss_off index the consumer's channels.
struct PostTarget {
  __nv_bfloat16* dst;
  const float* gamma;
  const float* beta;
  int64_t ss_off;   // scale at ss[row][ss_off + ch], shift at ss[row][ss_off + dst_C + ch]; < 0: no scale-shift
  int dst_C, c_off, cpg, silu;
};

struct ConvKParams {
  int n_m_tiles, n_n_tiles, stages;
  int Wb, Hb, Nb, H_out, W_out, tiles_per_img;
  int stride, taps, cin_blocks, s0_blocks, s1_blocks;
  int tap_cols, dy0, dx0, out_scale, out_oy, out_ox, H_full, W_full;
  int msub;               // 2: a CTA owns two vertically adjacent 128-pixel sub-tiles fed by ONE (2*Hb+2)-row box and the same
                          //    weight tiles (tall mode only): weight traffic per pixel halves, halo overhead 1.5x -> 1.25x
  int n_par, c_out_pad;   // n_par = 4: the four output-parity 2x2 convs of a folded upsample+conv3x3 share one launch
  int l2_prefetch, xf_dbg;
  int tall, stage_bytes;  // tall: one (Hb+2)-row activation box per (channel block, dx) serves the three dy taps
  int dx_taps;            // 3, or 1 = "dx-stacked" thin convolution (the network's final conv): the three horizontal taps are stacked
                          // along N (N = 3 x 16), ONE unshifted box per channel block feeds them, and the epilogue adds the three
                          // partial results of neighbouring pixels (lane shuffles): a third of the MMAs and of the activation traffic
  int tma_store;          // bf16 NHWC output through shared memory + cp.async.bulk.tensor stores (one 32 x 32 box per warp and chunk)
  int64_t B;
  int C_out, C_out_real, out_mode;
  const float* bias;
  const __nv_bfloat16* residual;
  void* out;
  const float2* ab;       // XF kernels: GroupNorm coefficients of the main input, [B][c_in_total] (a/2, b/2) pairs
  int c0_blocks, c_in_total;  // XF: channel blocks served by the first main source (tmA), the rest come from tmA2
  float* stats;           // GroupNorm partial sums of the output [B][stats_parts][C_out/4][2], or nullptr
  int stats_parts, units_per_img, stats_wpi;  // rows per image, work units per image, epilogue warps per image and unit
  FastDiv fd_ipp, fd_nnt, fd_tpi;  // multiply-high division by items_per_par / n_n_tiles / tiles_per_img: the per-item index math of
                                   // the producer and epilogue warps sits on the critical path of short-K work items
  // POST kernels
  int reverse;            // walk the work items in DESCENDING order (see ConvLaunch::reverse)
  int post_n;             // targets (1 or 2)
  int ipu_log;            // Nb == 1: log2(work items per sample) -- every CTA walks WHOLE samples (unit = CG samples x one N tile)
  int n_units;            // Nb == 1: ceil(B / CG) * n_n_tiles
  int gne_raw;            // GNE kernels: 1 = the raw (un-normalised) output is stored as well (it has other readers)
  int stats_half;         // 4x4 maps: a warp's 32 tile rows are two images, statistics are reduced per half warp
  const float* ss;        // time-embedding scale / shift table [ss_rows][ss_stride] (unet_ops.cu: k_gemv_rows)
  int ss_rows;
  int64_t ss_stride;
  PostTarget post[2];
};

// Coordinates of the it-th work item of a CTA (pair).  Default walk: item = first + it * stride over (parity, M group, N tile).
// POST kernels on maps of >= 128 pixels walk sample-major instead: CTA r of a pair owns sample (unit * CG + r) and runs its
// ipu = tiles_per_img / msub items back to back, so a sample's output (and its GroupNorm statistics) is complete -- and still
// L2-resident -- on ONE CTA when its last accumulator has been drained.
struct ItemCoord { int par, nt, mt, n0, h0; };
template <int CG, bool POST>
__device__ __forceinline__ bool item_coord(const ConvKParams& p, int it, int first, int stride, int cta_rank, int msub, int items_per_par,
                                           int n_items, ItemCoord& c) {
  if (POST && p.Nb == 1) {
    const int su = it >> p.ipu_log, j = it - (su << p.ipu_log);
    int gu = first + su * stride;
    if (gu >= p.n_units) return false;
    if (p.reverse) gu = p.n_units - 1 - gu;
    const int sp = (int)p.fd_nnt.div((uint32_t)gu);
    c.par = 0;
    c.nt = gu - sp * p.n_n_tiles;
    c.n0 = sp * CG + cta_rank;
    c.h0 = j * msub * p.Hb;
    c.mt = c.n0 * p.tiles_per_img + j * msub;
    return true;
  }
  int item = first + it * stride;
  if (item >= n_items) return false;
  if (p.reverse) item = n_items - 1 - item;
  c.par = (int)p.fd_ipp.div((uint32_t)item);
  const int it_in = item - c.par * items_per_par;
  const int mg = (int)p.fd_nnt.div((uint32_t)it_in);
  c.nt = it_in - mg * p.n_n_tiles;
  c.mt = (mg * CG + cta_rank) * msub;
  if (p.Nb == 1) { c.n0 = (int)p.fd_tpi.div((uint32_t)c.mt); c.h0 = (c.mt - c.n0 * p.tiles_per_img) * p.Hb; }
  else { c.n0 = c.mt * p.Nb; c.h0 = 0; }
  return true;
}

// CG = 1: one CTA per 128-pixel tile.  CG = 2: a CTA PAIR (cluster of 2 on one TPC) computes two adjacent 128-pixel
// tiles with tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128 activation rows and only HALF of the weight
// tile (BLOCK_N/2 rows) -- the tensor core reads both halves -- which halves the weight traffic from L2 and the
// shared-memory operand bandwidth per SM (the limiter of the N = 128 layers).  The leader CTA (rank 0) issues the MMAs;
// TMA completions of both CTAs are signalled on the leader's "full" barrier, tcgen05.commit multicasts the "empty" /
// "accumulator ready" arrivals to both CTAs, and both epilogues arrive remotely on the leader's "accumulator free" barrier.
//
// XF = true ("normalise on load", tall mode only): the activation tensor is the RAW input of GroupNorm + SiLU
// (unet.py:141-143,153-157,433-435) and the normalised tensor never exists in memory.  Eight extra warps rewrite every
// activation box in shared memory between its TMA arrival and its MMAs:  x -> silu(a[n,c] * x + b[n,c])  (coefficients
// from the conv-epilogue statistics, unet_ops.cu k_gn_fold), forcing the zero padding back to zero.  The main input may
// be the virtual concatenation of two tensors (tmA | tmA2).  Barrier chain per stage:
//   TMA(A) -> fullA (local) -> transform warps -> xf (leader, 8 x CG arrivals) -+-> MMA -> empty
//   TMA(B) -> full (leader) ----------------------------------------------------+
// Order of the pipeline stages of a "tall" work item: 3 * cin_blocks main stages (activation box + three weight tiles, 12..24
// MMAs each) and the 1x1 skip-conv stages (one K block, 4..8 MMAs each).  Run back to back, the skip stages leave the
// three-stage ring with less MMA work in flight than a TMA round trip takes; spread evenly between the main stages
// (Bresenham) every window of three stages carries enough.  Stage 0 is always a main stage.
__device__ __forceinline__ void tall_slot(int sb, int n_main, int n_skip, bool interleave, bool& main_part, int& idx) {
  if (!interleave || n_skip == 0) {
    main_part = sb < n_main;
    idx = main_part ? sb : sb - n_main;
    return;
  }
  const int n_sb = n_main + n_skip;
  const int before = (sb * n_skip) / n_sb;
  const bool is_skip = ((sb + 1) * n_skip) / n_sb > before;
  main_part = !is_skip;
  idx = is_skip ? before : sb - before;
}

// y = silu(a x + b) (coefficients pre-halved: silu(y) = h + h tanh(h) with h = y / 2, ONE MUFU op) or a x + b, on eight bf16
__device__ __forceinline__ uint4 post_apply8(const uint4& raw, const float (&a)[8], const float (&b)[8], int silu) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float y0 = fmaf(__uint_as_float(w[e] << 16), a[2 * e], b[2 * e]);
    float y1 = fmaf(__uint_as_float(w[e] & 0xffff0000u), a[2 * e + 1], b[2 * e + 1]);
    if (silu) {
      float t0, t1;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(y0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(y1));
      y0 = fmaf(y0, t0, y0);
      y1 = fmaf(y1, t1, y1);
    }
    const __nv_bfloat162 pk = __floats2bfloat162_rn(y0, y1);
    ow[e] = *reinterpret_cast<const uint32_t*>(&pk);
  }
  return o;
}

template <int BLOCK_N, bool XF>
__host__ __device__ constexpr int conv_epi_groups() { return XF ? 1 : 2; }
template <int BLOCK_N, bool XF, bool POST = false>
__host__ __device__ constexpr int conv_threads() {
  return XF ? kConvThreadsXF : (POST ? kConvThreadsPost : (conv_epi_groups<BLOCK_N, XF>() == 2 ? kConvThreadsEpi2 : kConvThreads));
}

// GNE = true ("GroupNorm in the epilogue", CTA pairs on maps of 256 pixels): the pair's accumulator stage holds ONE WHOLE SAMPLE
// (128 pixels x BLOCK_N channels per CTA), so the GroupNorm(s) that consume this convolution's output (unet.py:141,153,188-191)
// are applied straight from TMEM: pass 1 reads the accumulators, adds bias / residual, (optionally stores the raw rows) and
// reduces the per-quad statistics; the two CTAs exchange their totals through distributed shared memory; pass 2 reads the
// accumulators AGAIN and writes silu((v - mean) * rstd * gamma' + beta') into the consumer's input tensor.  No re-read of the
// output through L2 (the POST variant's loss), no separate k_gn_apply pass; the second stage keeps the MMAs of the next sample
// running underneath.
template <int BLOCK_N, int BLOCK_K, int CG, int KS, bool XF, bool POST, int GNE>
__global__ void __launch_bounds__(conv_threads<BLOCK_N, XF, POST>(), 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmS0,
          const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
          const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1, const ConvKParams p) {
  constexpr int A_BYTES = 128 * BLOCK_K * 2;
  constexpr int B_ROWS = BLOCK_N / CG;
  constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  constexpr int SUB_BYTES = A_BYTES + B_BYTES;      // one K block (BLOCK_K channels) of both operands
  // A pipeline stage carries KS K blocks (narrow tiles need fewer barrier round trips per MMA cycle), or -- "tall" mode,
  // 3x3 stride-1 convs whose tile lies inside one image -- ONE activation box of Hb+2 image rows plus the three weight
  // tiles of the taps dy = -1, 0, +1: a dy shift is a whole number of 8-row swizzle atoms (Wb rows), so the three MMAs
  // read the same box at row offsets 0, Wb, 2*Wb and the activation traffic from L2 drops from 9 to 3*(Hb+2)/Hb tiles.
  const int STAGE_BYTES = p.stage_bytes;
  constexpr int SWZ = BLOCK_K * 2;  // bytes per operand row = swizzle span (128 or 64)
  constexpr int MS_MAX = BLOCK_N <= 128 ? 2 : 1;  // sub-tiles per CTA the accumulator space allows (2 stages x MS_MAX x N <= 512)
  constexpr int ACC_COLS = 2 * MS_MAX * BLOCK_N;
  constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc_bf16(128 * CG, BLOCK_N);
  // short-K layers are epilogue-bound with one warp per quadrant (TMEM load -> bias/residual/statistics -> store is a long
  // dependent chain): two warps per quadrant split the tile's column chunks and interleave on the same sub-partition.
  // Tiles with a single chunk (N <= 32: the 32-channel layers of the MNIST network, the final conv) split the work item's
  // two SUB-TILES instead.
  constexpr int EG = conv_epi_groups<BLOCK_N, XF>();
  static_assert(!(XF && POST), "normalise-on-load and the producer-side GroupNorm are alternatives");
  static_assert(!POST || (BLOCK_N >= 128 && EG == 2), "POST kernels: N tiles of 128 / 256 channels");
  static_assert(!GNE || (!XF && !POST && BLOCK_N >= 128 && EG == 2), "GNE kernels: N tiles of 128 / 256 channels");
  static_assert(GNE != 1 || CG == 2, "GNE mode 1 (the pair's accumulator stage = one sample): CTA pairs");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem + p.stages * STAGE_BYTES) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + kEpiStageTotal);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* fullA_bar = empty_bar + kMaxStages;  // XF only
  uint64_t* xf_bar = fullA_bar + kMaxStages;     // XF only
  uint64_t* tfull_bar = xf_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);  // 16-byte aligned (the barrier block is a multiple of 16 bytes)
  uint64_t* ready_bar = reinterpret_cast<uint64_t*>(s_bias + kBiasSmemFloats);  // POST: "sample complete" (epilogue -> post warps), 2 slots
  uint64_t* free_bar = ready_bar + 2;                                            // POST: slot consumed (post warps -> epilogue)
  float* s_pq = reinterpret_cast<float*>(free_bar + 2);                          // POST: [BLOCK_N / 4][2] quad sums of the sample in flight

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform (see tc::elect_one)
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int main_blocks = p.taps * p.cin_blocks;
  const int nkb = main_blocks + p.s0_blocks + p.s1_blocks;
  const int msub = p.msub;
  const int m_units = p.n_m_tiles / msub;             // a unit = msub vertically adjacent 128-pixel tiles of one image
  const int m_groups = (m_units + CG - 1) / CG;        // a work item = CG adjacent units x one N tile
  const int items_per_par = m_groups * p.n_n_tiles;
  const int n_items = items_per_par * p.n_par;
  const int first_item = blockIdx.x / CG, item_stride = gridDim.x / CG;
  // POST: work items per completion unit (a CTA's tile holds whole samples when Nb > 1; else ipu items make one sample)
  const int ipu_log = (POST && p.Nb == 1) ? p.ipu_log : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (XF && p.c0_blocks < p.cin_blocks) prefetch_tmap(&tmA2);
    if (p.s0_blocks) prefetch_tmap(&tmS0);
    if (p.s1_blocks) prefetch_tmap(&tmS1);
    if (p.tma_store) prefetch_tmap(&tmO);
    if (GNE) { prefetch_tmap(&tmP0); if (p.post_n > 1) prefetch_tmap(&tmP1); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
      if (XF) { mbar_init(fullA_bar + s, 1); mbar_init(xf_bar + s, kXfWarps * CG); }
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + a, 1); mbar_init(tempty_bar + a, 4 * EG * CG); }
    if (POST) {
      for (int a = 0; a < 2; ++a) { mbar_init(ready_bar + a, (uint32_t)(4 * EG) << ipu_log); mbar_init(free_bar + a, 1); }
    }
    if (GNE == 1) { mbar_init(xf_bar, BLOCK_N / 2); mbar_init(xf_bar + 1, BLOCK_N / 2); }  // statistics exchange (one barrier per item parity)
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncwarp();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the prologue above overlapped the previous kernel's tail; its outputs may only be touched from here on
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of B) =====================
    {
      const bool elected = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0;; ++it) {
        ItemCoord ic;
        if (!item_coord<CG, POST>(p, it, first_item, item_stride, (int)cta_rank, msub, items_per_par, n_items, ic)) break;
        const int item = first_item + it * item_stride;  // (default walk only: used by the optional L2 prefetch)
        const int par = ic.par, nt = ic.nt, n0 = ic.n0, h0 = ic.h0;
        const int dy_base = p.dy0 + (p.n_par == 4 ? (par >> 1) : 0), dx_base = p.dx0 + (p.n_par == 4 ? (par & 1) : 0);
        const int brow0 = par * p.c_out_pad + nt * BLOCK_N + (int)cta_rank * B_ROWS;
        if (p.tall) {
          const int a_tall_bytes = (msub * p.Hb + 2) * p.Wb * BLOCK_K * 2;
          const int n_sb = p.dx_taps * p.cin_blocks + p.s0_blocks + p.s1_blocks;
          if (!POST && p.l2_prefetch && item + item_stride < n_items) {
            // pull the NEXT work item's activation boxes from HBM into L2 now: its TMA loads then see L2 latency only
            const int it2 = (item + item_stride) % items_per_par;
            const int mt2 = ((it2 / p.n_n_tiles) * CG + (int)cta_rank) * msub;
            const int n2 = mt2 / p.tiles_per_img, h2 = (mt2 - n2 * p.tiles_per_img) * p.Hb;
            if (it2 % p.n_n_tiles == 0 || p.n_n_tiles == 1) {
              for (int cblk = 0; cblk < p.cin_blocks; ++cblk) {
                const bool src0 = !XF || cblk < p.c0_blocks;
                tma_prefetch_4d_e(elected, src0 ? &tmA : &tmA2, (src0 ? cblk : cblk - p.c0_blocks) * BLOCK_K, 0, h2 - 1, n2);
              }
            }
          }
          for (int sb = 0; sb < n_sb; ++sb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* a_dst = smem + stage * STAGE_BYTES;
            uint8_t* b_dst = a_dst + a_tall_bytes;
            bool main_part; int slot_idx;
            tall_slot(sb, p.dx_taps * p.cin_blocks, p.s0_blocks + p.s1_blocks, !XF, main_part, slot_idx);
            const int a_bytes = main_part ? a_tall_bytes : msub * A_BYTES, b_bytes = main_part ? 3 * B_BYTES : B_BYTES;
            // XF: activations complete on this CTA's own fullA barrier (its transform warps wait there), weights on full
            const int bytes = XF ? b_bytes : a_bytes + b_bytes;
            uint32_t lead_full = 0;
            if (CG == 2) {
              lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
              if (leader) mbar_expect_tx_e(elected, full_bar + stage, 2 * bytes);
            } else {
              mbar_expect_tx_e(elected, full_bar + stage, bytes);
            }
            if (XF) mbar_expect_tx_e(elected, fullA_bar + stage, a_bytes);
            const int brow = brow0;
            if (main_part) {
              const int cblk = p.dx_taps == 3 ? slot_idx / 3 : slot_idx, dxi = p.dx_taps == 3 ? slot_idx - cblk * 3 : 1;
              if (XF) {
                const bool src0 = cblk < p.c0_blocks;
                tma_load_4d_e(elected, src0 ? &tmA : &tmA2, fullA_bar + stage, a_dst, (src0 ? cblk : cblk - p.c0_blocks) * BLOCK_K, dxi - 1, h0 - 1, n0);
              } else if (CG == 2) tma_load_4d_pair_e(elected, &tmA, lead_full, a_dst, cblk * BLOCK_K, dxi - 1, h0 - 1, n0);
              else tma_load_4d_e(elected, &tmA, full_bar + stage, a_dst, cblk * BLOCK_K, dxi - 1, h0 - 1, n0);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int kcol = (p.dx_taps == 3 ? (j * 3 + dxi) * p.cin_blocks + cblk : j * p.cin_blocks + cblk) * BLOCK_K;
                if (CG == 2) tma_load_2d_pair_e(elected, &tmB, lead_full, b_dst + j * B_BYTES, kcol, brow);
                else tma_load_2d_e(elected, &tmB, full_bar + stage, b_dst + j * B_BYTES, kcol, brow);
              }
            } else {
              const int e = slot_idx;
              const CUtensorMap* map = e < p.s0_blocks ? &tmS0 : &tmS1;
              const int c_a = (e < p.s0_blocks ? e : e - p.s0_blocks) * BLOCK_K;
              const int kcol = (main_blocks + e) * BLOCK_K;
              if (XF) tma_load_4d_e(elected, map, fullA_bar + stage, a_dst, c_a, 0, h0, n0);
              else if (CG == 2) tma_load_4d_pair_e(elected, map, lead_full, a_dst, c_a, 0, h0, n0);
              else tma_load_4d_e(elected, map, full_bar + stage, a_dst, c_a, 0, h0, n0);
              if (CG == 2) tma_load_2d_pair_e(elected, &tmB, lead_full, b_dst, kcol, brow);
              else tma_load_2d_e(elected, &tmB, full_bar + stage, b_dst, kcol, brow);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        for (int kb0 = 0; kb0 < nkb; kb0 += KS) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          const int cnt = (nkb - kb0) < KS ? (nkb - kb0) : KS;
          uint32_t lead_full = 0;
          if (CG == 2) {
            lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
            if (leader) mbar_expect_tx_e(elected, full_bar + stage, 2 * cnt * SUB_BYTES);
          } else {
            mbar_expect_tx_e(elected, full_bar + stage, cnt * SUB_BYTES);
          }
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            if (ks >= cnt) break;
            const int kb = kb0 + ks;
            uint8_t* a_dst = smem + stage * STAGE_BYTES + ks * SUB_BYTES;
            uint8_t* b_dst = a_dst + A_BYTES;
            int c_a, x_a, y_a;
            const CUtensorMap* map;
            if (kb < main_blocks) {
              const int tap = kb / p.cin_blocks, cblk = kb - tap * p.cin_blocks;
              const int trow = tap / p.tap_cols;
              const int dy = dy_base + trow, dx = dx_base + tap - trow * p.tap_cols;
              map = &tmA; c_a = cblk * BLOCK_K; x_a = dx; y_a = h0 * p.stride + dy;
            } else if (kb < main_blocks + p.s0_blocks) {
              map = &tmS0; c_a = (kb - main_blocks) * BLOCK_K; x_a = 0; y_a = h0;
            } else {
              map = &tmS1; c_a = (kb - main_blocks - p.s0_blocks) * BLOCK_K; x_a = 0; y_a = h0;
            }
            if (CG == 2) {
              tma_load_4d_pair_e(elected, map, lead_full, a_dst, c_a, x_a, y_a, n0);
              tma_load_2d_pair_e(elected, &tmB, lead_full, b_dst, kb * BLOCK_K, brow0);
            } else {
              tma_load_4d_e(elected, map, full_bar + stage, a_dst, c_a, x_a, y_a, n0);
              tma_load_2d_e(elected, &tmB, full_bar + stage, b_dst, kb * BLOCK_K, brow0);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const bool elected = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0;; ++it) {
        ItemCoord ic;
        if (!item_coord<CG, POST>(p, it, first_item, item_stride, (int)cta_rank, msub, items_per_par, n_items, ic)) break;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);  // epilogues (of both CTAs) have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * msub * BLOCK_N);
        if (p.tall) {
          const int a_tall_bytes = (msub * p.Hb + 2) * p.Wb * BLOCK_K * 2;
          const int n_sb = p.dx_taps * p.cin_blocks + p.s0_blocks + p.s1_blocks;
          const int row_units = (p.Wb * BLOCK_K * 2) >> 4, sub_units = p.Hb * row_units;  // descriptor address units (16 bytes)
          for (int sb = 0; sb < n_sb; ++sb) {
            mbar_wait(full_bar + stage, phase);
            if (XF) mbar_wait(xf_bar + stage, phase);  // activation boxes of both CTAs normalised in place
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
            const uint32_t b_addr = a_addr + a_tall_bytes;
            bool main_part; int slot_idx;
            tall_slot(sb, p.dx_taps * p.cin_blocks, p.s0_blocks + p.s1_blocks, !XF, main_part, slot_idx);
            // One descriptor pair per stage; every (dy, sub-tile, K step) operand is that descriptor plus a loop-invariant
            // offset in its 16-byte address field (dy = j - 1 shifts by Wb rows = whole swizzle atoms, sub-tile `sub` starts Hb
            // image rows further down, tap j's weight tile follows tap j-1's, a K step of 16 elements is 32 bytes inside the
            // swizzled row).  Fully unrolled: the issue loop costs a 64-bit add per operand instead of re-deriving addresses
            // and descriptors per MMA (ncu on the N = 32 layers: 24 warp instructions = 109 clk per MMA in the issuing warp
            // against 16 clk of tensor work; the MMA issuer, not TMA or the epilogue, set the pace).
            const uint64_t da_s = make_smem_desc<SWZ>(a_addr), db_s = make_smem_desc<SWZ>(b_addr);
            if (main_part) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int sub = 0; sub < MS_MAX; ++sub) {
                  if (sub < msub) {
                    const uint64_t da0 = da_s + (uint64_t)(uint32_t)(sub * sub_units + j * row_units);
                    const uint64_t db0 = db_s + (uint64_t)(j * (B_BYTES >> 4));
                    const uint32_t d_sub = d_tmem + (uint32_t)(sub * BLOCK_N);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                      const uint32_t accum = (j | k) != 0 ? 1u : (uint32_t)(sb != 0);
                      if (CG == 2) umma_bf16_pair_e(elected, d_sub, da0 + 2 * k, db0 + 2 * k, IDESC, accum);
                      else umma_bf16_e(elected, d_sub, da0 + 2 * k, db0 + 2 * k, IDESC, accum);
                    }
                  }
                }
              }
            } else {  // fused 1x1 skip-conv stage: plain 128-row A tiles, one weight tile
#pragma unroll
              for (int sub = 0; sub < MS_MAX; ++sub) {
                if (sub < msub) {
                  const uint64_t da0 = da_s + (uint64_t)(sub * (A_BYTES >> 4));
                  const uint32_t d_sub = d_tmem + (uint32_t)(sub * BLOCK_N);
#pragma unroll
                  for (int k = 0; k < BLOCK_K / 16; ++k) {
                    if (CG == 2) umma_bf16_pair_e(elected, d_sub, da0 + 2 * k, db_s + 2 * k, IDESC, (sb | k) != 0);
                    else umma_bf16_e(elected, d_sub, da0 + 2 * k, db_s + 2 * k, IDESC, (sb | k) != 0);
                  }
                }
              }
            }
            if (CG == 2) umma_commit_pair_e(elected, empty_bar + stage); else umma_commit_e(elected, empty_bar + stage);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (CG == 2) umma_commit_pair_e(elected, tfull_bar + acc); else umma_commit_e(elected, tfull_bar + acc);
          continue;
        }
        for (int kb0 = 0; kb0 < nkb; kb0 += KS) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const int cnt = (nkb - kb0) < KS ? (nkb - kb0) : KS;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            if (ks >= cnt) break;
            const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES + ks * SUB_BYTES);
            const uint32_t b_addr = a_addr + A_BYTES;
            const uint64_t da0 = make_smem_desc<SWZ>(a_addr), db0 = make_smem_desc<SWZ>(b_addr);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              if (CG == 2) umma_bf16_pair_e(elected, d_tmem, da0 + 2 * k, db0 + 2 * k, IDESC, (kb0 | ks | k) != 0);
              else umma_bf16_e(elected, d_tmem, da0 + 2 * k, db0 + 2 * k, IDESC, (kb0 | ks | k) != 0);
            }
          }
          if (CG == 2) umma_commit_pair_e(elected, empty_bar + stage); else umma_commit_e(elected, empty_bar + stage);  // frees the smem slot(s)
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair_e(elected, tfull_bar + acc); else umma_commit_e(elected, tfull_bar + acc);  // accumulator complete -> epilogue(s)
      }
    }
  } else if (XF && warp >= 8) {
    // ===================== transform warps: GroupNorm + SiLU applied to the activation box in shared memory =====================
    const int xt = (warp - 8) * 32 + lane;   // 0 .. 255
    const int chunk = xt & 7, rl = xt >> 3;  // 16-byte chunk (8 channels) of a 128-byte row; 32 row lanes
    const int box_rows = (msub * p.Hb + 2) * p.Wb;
    const int wshift = 31 - __clz(p.Wb);
    const int n_sb = p.dx_taps * p.cin_blocks + p.s0_blocks + p.s1_blocks;
    const uint32_t xf_remote = CG == 2 ? mapa_u32(smem_u32(xf_bar), 0) : 0u;
    // a thread's rows are rl, rl + 32, ...: Wb divides 32, so its pixel column and its swizzle phase never change
    const int x_base = (rl & (p.Wb - 1)) - 1, y_step = 32 >> wshift;
    const uint32_t row_off = (uint32_t)(rl * 128 + ((chunk ^ (rl & 7)) << 4));
    int stage = 0; uint32_t phase = 0;
    for (int it = 0;; ++it) {
      ItemCoord ic;
      if (!item_coord<CG, false>(p, it, first_item, item_stride, (int)cta_rank, msub, items_per_par, n_items, ic)) break;
      const int n0 = ic.n0, h0 = ic.h0;
      const float4* ab_row = reinterpret_cast<const float4*>(p.ab + (int64_t)(n0 < p.B ? n0 : p.B - 1) * p.c_in_total);
      const int y_first = h0 - 1 + (rl >> wshift);
      for (int sb = 0; sb < n_sb; ++sb) {
        const bool main_part = sb < p.dx_taps * p.cin_blocks;
        const int cblk = p.dx_taps == 3 ? sb / 3 : sb, dxi = p.dx_taps == 3 ? sb - cblk * 3 : 1;  // (dx-stacked: one unshifted box per block)
        float4 co[4];  // (a, b) of this thread's 8 channels; fetched while the box is still in flight
        if (main_part) {
#pragma unroll
          for (int e = 0; e < 4; ++e) co[e] = __ldg(ab_row + (cblk * (BLOCK_K / 2) + chunk * 4 + e));
        }
        mbar_wait(fullA_bar + stage, phase);
        if (main_part && p.xf_dbg != 1) {
          const bool x_ok = (unsigned)(x_base + dxi) < (unsigned)p.W_out;
          uint8_t* ptr = smem + stage * STAGE_BYTES + row_off;
          int y = y_first;
#pragma unroll 5
          for (int r = rl; r < box_rows; r += 32, ptr += 32 * 128, y += y_step) {
            const bool inb = x_ok && (unsigned)y < (unsigned)p.H_out;
            uint4 v = *reinterpret_cast<uint4*>(ptr);
            uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // silu(y) = h + h * tanh(h), h = y / 2 (the fold kernel halves the coefficients); tanh on the packed
              // bf16x2 MUFU path (one op per two elements -- its 2^-9 error is that of the bf16 result anyway)
              const float h0v = fmaf(__uint_as_float(w[e] << 16), co[e].x, co[e].y);
              const float h1v = fmaf(__uint_as_float(w[e] & 0xffff0000u), co[e].z, co[e].w);
              const __nv_bfloat162 hp = __floats2bfloat162_rn(h0v, h1v);
              uint32_t tp;
              asm("tanh.approx.bf16x2 %0, %1;" : "=r"(tp) : "r"(*reinterpret_cast<const uint32_t*>(&hp)));
              const __nv_bfloat162 o = __floats2bfloat162_rn(fmaf(h0v, __uint_as_float(tp << 16), h0v),
                                                             fmaf(h1v, __uint_as_float(tp & 0xffff0000u), h1v));
              w[e] = inb ? *reinterpret_cast<const uint32_t*>(&o) : 0u;
            }
            *reinterpret_cast<uint4*>(ptr) = v;
          }
          if (p.xf_dbg != 2) fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        }
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(xf_remote + (uint32_t)(stage * 8));
          else mbar_arrive(xf_bar + stage);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * EG) {
    // ===================== epilogue =====================
    // the bias vector is only needed here: loaded by the epilogue warps and published with a named barrier among them, so
    // that its L2 round trip is off the path of the TMA producer / MMA issuer (the first accumulator is microseconds away)
    for (int i = (int)threadIdx.x - 128; i < p.c_out_pad; i += 128 * EG) s_bias[i] = __ldg(p.bias + i);
    // GNE: per-target coefficient tables behind the bias vector, A = gamma (1 + scale), B = beta (1 + scale) + shift (halved for the
    // one-MUFU SiLU form), so that pass 2 evaluates y = ((v - mean) rstd) A + B.  Batch-constant in sampling (ss_rows == 1): filled
    // once per kernel; per-sample time steps refill them per item.
    const int CT = p.c_out_pad;  // table rows cover every output channel (mode 1: one N tile = all channels; mode 2: any number of N tiles)
    float* const s_gA = s_bias + CT;
    float* const s_gB = s_gA + p.post_n * CT;
    float* const s_tot = s_gB + p.post_n * CT;       // mode 1: [2 item parities][2 CTA ranks][BLOCK_N / 2] quad (sum, sum of squares) totals
    float* const s_rn = s_tot + 2 * BLOCK_N;         // mode 1: [targets][BLOCK_N / 4] (rstd, -mean rstd) of every quad's group, current item
    auto gne_fill_tables = [&](int64_t n) {
      for (int i = (int)threadIdx.x - 128; i < p.post_n * CT; i += 128 * EG) {
        const int k = i / CT, ch = i - k * CT;
        const PostTarget& tg = p.post[k];
        const int ct = tg.c_off + ch;
        float ga = __ldg(tg.gamma + ct), be = __ldg(tg.beta + ct);
        if (tg.ss_off >= 0) {
          const float* row = p.ss + (p.ss_rows == 1 ? 0 : n * p.ss_stride) + tg.ss_off;
          const float sc = 1.0f + __ldg(row + ct), sh = __ldg(row + tg.dst_C + ct);
          ga *= sc;
          be = be * sc + sh;
        }
        const float osc = tg.silu ? 0.5f : 1.0f;  // silu(y) = h + h tanh(h) with h = y / 2
        s_gA[i] = ga * osc;
        s_gB[i] = be * osc;
      }
    };
    // (mode 2 only runs with one table per kernel: batch-constant rows, or no target that reads them -- gne2_needs_post_warps)
    if (GNE == 2 || (GNE == 1 && p.ss_rows == 1)) gne_fill_tables(0);
    asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int eg = (warp - 4) >> 2;  // which half of the column chunks (EG == 2)
    const int m = q * 32 + lane;
    const int w_in = m % p.Wb, h_in = (m / p.Wb) % p.Hb, n_in = m / (p.Wb * p.Hb);
    // TMA-store path: the warp's packed bf16 chunk (32 pixels x 64 bytes) is staged in shared memory in the SWIZZLE_64B
    // pattern (16-byte granule ^= (row >> 1) & 3: conflict-free st.shared.v4) and leaves as ONE cp.async.bulk.tensor box.
    // Measured (xf_dbg ablation, B = 512): per-lane 16-byte STG cost 0.55 ms of the 4.45 ms of convolutions per forward --
    // 32 half-filled sectors per instruction on the SM -> L2 write path, with the epilogue on the critical path.
    const uint32_t my_stage = smem_u32(epi_stage) + (uint32_t)(warp - 4) * kEpiStageBytes;
    const uint32_t my_row = my_stage + (uint32_t)lane * 64u;
    const int m0 = q * 32;  // first tile row of this warp: coordinates of the store box
    const int w_box = m0 % p.Wb, h_box = (m0 / p.Wb) % p.Hb, n_box = m0 / (p.Wb * p.Hb);
    for (int it = 0;; ++it) {
      ItemCoord ic;
      if (!item_coord<CG, POST>(p, it, first_item, item_stride, (int)cta_rank, msub, items_per_par, n_items, ic)) break;
      const int par = ic.par, nt = ic.nt, mt = ic.mt, n0 = ic.n0, h0 = ic.h0;
      const int out_oy = p.n_par == 4 ? (par >> 1) : p.out_oy, out_ox = p.n_par == 4 ? (par & 1) : p.out_ox;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int64_t nn = (int64_t)n0 + n_in;
      const bool valid = nn < p.B;
      const uint32_t t_row0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * msub * BLOCK_N);
      if constexpr (GNE == 2) {
        // ---------- GroupNorm in the epilogue, 4x4 maps: a warp's 32 tile rows are TWO WHOLE SAMPLES (half warps), so the statistics
        // never leave the warp and the chunk is normalised from the registers it was loaded into -- one TMEM pass, no exchange ----------
        constexpr int NCH = BLOCK_N / 32, CPW = NCH / EG;
        uint4 resv[CPW * 4];
        const bool has_res = p.residual != nullptr && valid;
        const int64_t pix = (nn * p.H_full + h_in) * p.W_full + w_in;
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.C_out + nt * BLOCK_N);
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci)
#pragma unroll
            for (int j = 0; j < 4; ++j) resv[ci * 4 + j] = __ldg(rp + (ci * EG + eg) * 4 + j);
        }
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
        const float inv_hw = 1.0f / (float)(p.H_full * p.W_full);
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c0 = (ci * EG + eg) * 32, col = nt * BLOCK_N + c0;
          uint32_t r[32];
          tmem_ld_x32(t_row0 + c0, r);
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          tmem_ld_wait();
          float2 v[16];
          float st[16];
          {
            const uint32_t bias_s = smem_u32(s_bias) + (uint32_t)col * 4u;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 b0 = lds_f4(bias_s + j * 4), b1 = lds_f4(bias_s + j * 4 + 16);
              v[j / 2 + 0] = __fadd2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), make_float2(b0.x, b0.y));
              v[j / 2 + 1] = __fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), make_float2(b0.z, b0.w));
              v[j / 2 + 2] = __fadd2_rn(make_float2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])), make_float2(b1.x, b1.y));
              v[j / 2 + 3] = __fadd2_rn(make_float2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])), make_float2(b1.z, b1.w));
              if (has_res) {
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&resv[ci * 4 + j / 8]);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[j / 2 + e] = __fadd2_rn(v[j / 2 + e], __bfloat1622float2(rp[e]));
              }
              if (p.gne_raw) {
                uint4 o;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) op[e] = __float22bfloat162_rn(v[j / 2 + e]);
                sts_u4(my_row + (uint32_t)(((j >> 3) ^ ((lane >> 1) & 3)) << 4), o);
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int qi = (j / 4 + h) * 2;
                const float2 s2 = __fadd2_rn(v[j / 2 + 2 * h], v[j / 2 + 2 * h + 1]);
                const float2 q2 = __ffma2_rn(v[j / 2 + 2 * h + 1], v[j / 2 + 2 * h + 1], __fmul2_rn(v[j / 2 + 2 * h], v[j / 2 + 2 * h]));
                st[qi] = s2.x + s2.y;
                st[qi + 1] = q2.x + q2.y;
              }
            }
          }
          if (p.gne_raw) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmO, reinterpret_cast<const void*>(epi_stage + (warp - 4) * kEpiStageBytes), col, w_box, h_box, n0 + n_box);
              bulk_commit();
            }
          }
          float qs[8], qq2[8];  // totals of the chunk's eight quads over this lane's sample
          if (p.stats_half) {
            // 4x4: transposing butterfly inside each half warp (8 + 4 + 2 + 1 shuffles): lane l ends with the total of value l & 15 over
            // its sample; every lane then fetches all sixteen
#pragma unroll
            for (int half = 8, off = 8; half >= 1; half >>= 1, off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int k = 0; k < half; ++k) {
                const float send = up ? st[k] : st[k + half];
                const float keepv = up ? st[k + half] : st[k];
                st[k] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            if (p.stats != nullptr && valid) {  // other (unfused) GroupNorms still read the partial rows
              const int64_t row = nn * p.stats_parts + par;
              p.stats[(row * (p.C_out >> 2) + (col >> 2)) * 2 + (lane & 15)] = st[0];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              qs[k] = __shfl_sync(0xffffffffu, st[0], (lane & 16) | (2 * k));
              qq2[k] = __shfl_sync(0xffffffffu, st[0], (lane & 16) | (2 * k + 1));
            }
          } else {
            // 8x8: a sample is the 64 rows of TWO warps (quadrants 2s, 2s + 1, same column chunks): full-warp transposing butterfly
            // (lane l ends with the warp total of value l >> 1), the two warps swap their sixteen totals through shared memory
#pragma unroll
            for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int k = 0; k < half; ++k) {
                const float send = up ? st[k] : st[k + half];
                const float keepv = up ? st[k + half] : st[k];
                st[k] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            st[0] += __shfl_xor_sync(0xffffffffu, st[0], 1);
            if (p.stats != nullptr && valid && (lane & 1) == 0) {
              const int64_t row = (nn * p.stats_parts + (int64_t)par * p.stats_wpi + (q % p.stats_wpi));
              p.stats[(row * (p.C_out >> 2) + (col >> 2)) * 2 + (lane >> 1)] = st[0];
            }
            float* const s_x = s_gB + p.post_n * CT;  // [8 epilogue warps][16]
            const int bar_id = 2 + (q >> 1) * 2 + eg;  // the two warps of a sample that share this chunk column
            if ((lane & 1) == 0) s_x[(warp - 4) * 16 + (lane >> 1)] = st[0];
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            {
              const float4* mine = reinterpret_cast<const float4*>(s_x + (warp - 4) * 16);
              const float4* other = reinterpret_cast<const float4*>(s_x + (((warp - 4) ^ 1)) * 16);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 a = mine[k], b = other[k];
                // (fixed order: the lower quadrant's total first, so both warps of the sample get the same sums)
                const float4 lo = (q & 1) ? b : a, hi = (q & 1) ? a : b;
                qs[2 * k] = lo.x + hi.x; qq2[2 * k] = lo.y + hi.y; qs[2 * k + 1] = lo.z + hi.z; qq2[2 * k + 1] = lo.w + hi.w;
              }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");  // both have read: the slots may be rewritten (next chunk)
          }
#pragma unroll 1
          for (int k = 0; k < p.post_n; ++k) {
            const PostTarget& tg = p.post[k];
            const int nqg = tg.cpg >> 2;  // quads per group: 1, 2, 4 or 8 (cpg divides 32) -- xor-butterfly on the static arrays
            float gs[8], gq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { gs[j] = qs[j]; gq[j] = qq2[j]; }
#pragma unroll
            for (int step = 1; step < 8; step <<= 1) {
              if (nqg > step) {
                float ts[8], tq[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { ts[j] = gs[j & ~step] + gs[j | step]; tq[j] = gq[j & ~step] + gq[j | step]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) { gs[j] = ts[j]; gq[j] = tq[j]; }
              }
            }
            const float inv_n = inv_hw / (float)tg.cpg;
            float2 rn2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float mean = gs[j] * inv_n;
              const float rstd = rsqrtf(fmaxf(gq[j] * inv_n - mean * mean, 0.f) + 1e-5f);
              rn2[j] = make_float2(rstd, -mean * rstd);
            }
            const uint32_t tabA = smem_u32(s_gA) + (uint32_t)((k * CT + col) * 4), tabB = smem_u32(s_gB) + (uint32_t)((k * CT + col) * 4);
            uint4 o[4];
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 A0 = lds_f4(tabA + j * 4), A1 = lds_f4(tabA + j * 4 + 16), B0 = lds_f4(tabB + j * 4), B1 = lds_f4(tabB + j * 4 + 16);
              const float2 Aa[4] = {make_float2(A0.x, A0.y), make_float2(A0.z, A0.w), make_float2(A1.x, A1.y), make_float2(A1.z, A1.w)};
              const float2 Bb[4] = {make_float2(B0.x, B0.y), make_float2(B0.z, B0.w), make_float2(B1.x, B1.y), make_float2(B1.z, B1.w)};
              __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o[j >> 3]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 rq = rn2[j / 4 + e / 2];
                const float2 t2 = __ffma2_rn(v[j / 2 + e], make_float2(rq.x, rq.x), make_float2(rq.y, rq.y));
                float2 y2 = __ffma2_rn(t2, Aa[e], Bb[e]);
                if (tg.silu) {  // (warp-uniform) tables pre-halved: silu(y) = h + h tanh(h)
                  float2 th;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(y2.x));
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(y2.y));
                  y2 = __ffma2_rn(y2, th, y2);
                }
                op[e] = __float22bfloat162_rn(y2);
              }
            }
            if (lane == 0) bulk_wait_read0();  // the raw box / the previous target's box has been read out
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) sts_u4(my_row + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4), o[j]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(k == 0 ? &tmP0 : &tmP1, reinterpret_cast<const void*>(epi_stage + (warp - 4) * kEpiStageBytes), tg.c_off + col, w_box,
                           h_box, n0 + n_box);
              bulk_commit();
            }
          }
        }
      } else if constexpr (GNE == 1) {
        // ---------- GroupNorm in the epilogue: the pair's accumulator stage is one whole sample (msub == 1, two tiles per image) ----------
        constexpr int NCH = BLOCK_N / 32, CPW = NCH / EG, NQV = BLOCK_N / 2;
        if (p.ss_rows != 1) {  // per-sample time steps: this sample's scale / shift rows (the previous item's pass 2 is over)
          asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");
          gne_fill_tables(nn);
          asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");
        }
        uint4 resv[CPW * 4];
        const bool has_res = p.residual != nullptr;
        const int64_t pix = (nn * p.H_full + (h0 + h_in)) * p.W_full + w_in;
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.C_out);
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci)
#pragma unroll
            for (int j = 0; j < 4; ++j) resv[ci * 4 + j] = __ldg(rp + (ci * EG + eg) * 4 + j);
        }
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
        // v = accumulator + bias (+ identity residual) of eight channels, as four packed pairs (fp32x2 adds: half the issue slots)
        auto gne_v8 = [&](const uint32_t (&r)[32], int ci, int j, float2 (&v)[4]) {
          const uint32_t bias_s = smem_u32(s_bias) + (uint32_t)((ci * EG + eg) * 32 + j) * 4u;
          const float4 b0 = lds_f4(bias_s), b1 = lds_f4(bias_s + 16);
          v[0] = __fadd2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), make_float2(b0.x, b0.y));
          v[1] = __fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), make_float2(b0.z, b0.w));
          v[2] = __fadd2_rn(make_float2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])), make_float2(b1.x, b1.y));
          v[3] = __fadd2_rn(make_float2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])), make_float2(b1.z, b1.w));
          if (has_res) {
            const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&resv[ci * 4 + j / 8]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fadd2_rn(v[e], __bfloat1622float2(rp[e]));
          }
        };
        // ---- pass 1: statistics (and the raw rows when something else reads them)
        float keep[CPW];  // this lane's share of the warp totals, one value per chunk (lanes 2k, 2k+1 hold quad value k)
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c0 = (ci * EG + eg) * 32;
          uint32_t r[32];
          tmem_ld_x32(t_row0 + c0, r);
          if (p.gne_raw) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          }
          tmem_ld_wait();
          float st[16];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float2 v[4];
            gne_v8(r, ci, j, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) { r[j + 2 * e] = __float_as_uint(v[e].x); r[j + 2 * e + 1] = __float_as_uint(v[e].y); }
            if (p.gne_raw) {
              uint4 o;
              __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) op[e] = __float22bfloat162_rn(v[e]);
              sts_u4(my_row + (uint32_t)(((j >> 3) ^ ((lane >> 1) & 3)) << 4), o);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // (sum, sum of squares) of the quad: every quad of the chunk is visited exactly once
              const int qi = (j / 4 + h) * 2;
              const float2 s2 = __fadd2_rn(v[2 * h], v[2 * h + 1]);
              const float2 q2 = __ffma2_rn(v[2 * h + 1], v[2 * h + 1], __fmul2_rn(v[2 * h], v[2 * h]));
              st[qi] = s2.x + s2.y;
              st[qi + 1] = q2.x + q2.y;
            }
          }
          // v = accumulator + bias (+ residual) goes back to TMEM: pass 2 reads it as is, and the residual registers are dead from here
          tmem_st_x32(t_row0 + c0, r);
          if (p.gne_raw) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmO, reinterpret_cast<const void*>(epi_stage + (warp - 4) * kEpiStageBytes), c0, w_box, h0 + h_box, n0 + n_box);
              bulk_commit();
            }
          }
          // transposing butterfly (as in the plain epilogue): lane l ends with the warp total of value l >> 1
#pragma unroll
          for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < half; ++k) {
              const float send = up ? st[k] : st[k + half];
              const float keepv = up ? st[k + half] : st[k];
              st[k] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          st[0] += __shfl_xor_sync(0xffffffffu, st[0], 1);
          keep[ci] = st[0];
          if (p.stats != nullptr && (lane & 1) == 0) {  // other (unfused) GroupNorms still read the partial rows
            const int unit = (mt / msub) % p.units_per_img;
            const int64_t row = (nn * p.stats_parts + (int64_t)(par * p.units_per_img + unit) * p.stats_wpi + (q % p.stats_wpi));
            p.stats[(row * (p.C_out >> 2) + ((nt * BLOCK_N + c0) >> 2)) * 2 + (lane >> 1)] = st[0];
          }
        }
        // warp totals -> the warp's staging buffer (free once the raw box has been read out), CTA totals, exchange with the peer CTA
        if (lane == 0) bulk_wait_read0();  // (raw box of this item, or the previous item's last normalised box)
        __syncwarp();
        if ((lane & 1) == 0) {
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci) sts_f32(my_stage + (uint32_t)((ci * 16 + (lane >> 1)) * 4), keep[ci]);
        }
        tmem_st_wait();
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");
        const int et = (int)threadIdx.x - 128;
        const uint32_t xpar = (uint32_t)(it & 1);
        uint64_t* const xbar = xf_bar + xpar;  // one barrier per item parity: the peer's data for item it + 2 cannot arrive before this phase is over
        if (et < NQV) {
          const int quad = et >> 1, comp = et & 1, chunk = quad >> 3, egc = chunk & (EG - 1), cic = chunk / EG;
          const uint32_t off = (uint32_t)((cic * 16 + (((quad & 7) << 1) | comp)) * 4);
          float tot = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) tot += lds_f32(smem_u32(epi_stage) + (uint32_t)(qq + 4 * egc) * kEpiStageBytes + off);
          const uint32_t slot = (uint32_t)(((xpar * 2 + cta_rank) * NQV + et) * 4);
          sts_f32(smem_u32(s_tot) + slot, tot);
          // the peer's copy travels as an asynchronous DSMEM store that completes transaction bytes on the PEER's barrier: no
          // release fence (which would wait for this thread's outstanding global stores) on the epilogue's critical path
          st_async_f32(mapa_u32(smem_u32(s_tot), cta_rank ^ 1u) + slot, tot, mapa_u32(smem_u32(xbar), cta_rank ^ 1u));
          if (et == 0) mbar_expect_tx(xbar, NQV * 4);  // (counts as this thread's arrival) the peer's NQV values
          else mbar_arrive(xbar);
        }
        mbar_wait(xbar, (uint32_t)((it >> 1) & 1));
        // group statistics once per item: thread = (target, quad) -> (rstd, -mean rstd) of the quad's group, over both CTAs' totals
        // (rank 0 first: the same sum in both CTAs)
        {
          const float inv_hw = 1.0f / (float)(p.H_full * p.W_full);
          for (int i = et; i < p.post_n * (BLOCK_N / 4); i += 128 * EG) {
            const int k = i / (BLOCK_N / 4), quad = i - k * (BLOCK_N / 4);
            const int nqg = p.post[k].cpg >> 2, g0 = quad & ~(nqg - 1);
            const float2* t0 = reinterpret_cast<const float2*>(s_tot + xpar * 2 * NQV) + g0;
            float sA = 0.f, qA = 0.f;
            for (int j = 0; j < nqg; ++j) { const float2 a = t0[j], b = t0[j + NQV / 2]; sA += a.x + b.x; qA += a.y + b.y; }
            const float inv_n = inv_hw / (float)p.post[k].cpg;
            const float mean = sA * inv_n;
            const float rstd = rsqrtf(fmaxf(qA * inv_n - mean * mean, 0.f) + 1e-5f);
            reinterpret_cast<float2*>(s_rn)[i] = make_float2(rstd, -mean * rstd);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");
        // ---- pass 2: v again, normalised rows out
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c0 = (ci * EG + eg) * 32;
          uint32_t r[32];
          tmem_ld_x32(t_row0 + c0, r);
          if (lane == 0) bulk_wait_read0();  // staging buffer: the previous box has been read out (waited for while the TMEM load is in flight)
          __syncwarp();
          tmem_ld_wait();
#pragma unroll 1
          for (int k = 0; k < p.post_n; ++k) {  // (not unrolled: the epilogue's code size is what the instruction cache sees)
            const PostTarget& tg = p.post[k];
            float2 rn2[8];  // (rstd, -mean rstd) of the chunk's eight quads
            {
              const uint32_t rn = smem_u32(s_rn) + (uint32_t)((k * (BLOCK_N / 4) + (c0 >> 2)) * 8);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 a = lds_f4(rn + j * 16);
                rn2[2 * j] = make_float2(a.x, a.y); rn2[2 * j + 1] = make_float2(a.z, a.w);
              }
            }
            if (k > 0) {
              if (lane == 0) bulk_wait_read0();  // the first target's box
              __syncwarp();
            }
            const uint32_t tabA = smem_u32(s_gA) + (uint32_t)((k * CT + c0) * 4), tabB = smem_u32(s_gB) + (uint32_t)((k * CT + c0) * 4);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 A0 = lds_f4(tabA + j * 4), A1 = lds_f4(tabA + j * 4 + 16), B0 = lds_f4(tabB + j * 4), B1 = lds_f4(tabB + j * 4 + 16);
              const float2 Aa[4] = {make_float2(A0.x, A0.y), make_float2(A0.z, A0.w), make_float2(A1.x, A1.y), make_float2(A1.z, A1.w)};
              const float2 Bb[4] = {make_float2(B0.x, B0.y), make_float2(B0.z, B0.w), make_float2(B1.x, B1.y), make_float2(B1.z, B1.w)};
              uint4 o;
              __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {  // channel pair j + 2e, j + 2e + 1 (quad j / 4 + e / 2); fp32x2 FMAs
                const float2 rq = rn2[j / 4 + e / 2];
                const float2 v2 = make_float2(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1]));
                const float2 t2 = __ffma2_rn(v2, make_float2(rq.x, rq.x), make_float2(rq.y, rq.y));
                float2 y2 = __ffma2_rn(t2, Aa[e], Bb[e]);  // = y / 2 (tables pre-halved): silu(y) = h + h tanh(h)
                float2 th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(y2.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(y2.y));
                y2 = __ffma2_rn(y2, th, y2);
                op[e] = __float22bfloat162_rn(y2);
              }
              sts_u4(my_row + (uint32_t)(((j >> 3) ^ ((lane >> 1) & 3)) << 4), o);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(k == 0 ? &tmP0 : &tmP1, reinterpret_cast<const void*>(epi_stage + (warp - 4) * kEpiStageBytes), tg.c_off + c0, w_box,
                           h0 + h_box, n0 + n_box);
              bulk_commit();
            }
          }
        }
      } else if (p.out_mode == CONV_OUT_BF16_NHWC) {
        constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
        // identity-skip rows are fetched BEFORE waiting for the accumulator so their HBM latency hides behind the MMAs
        constexpr bool RES_PREFETCH = !XF && !POST;  // XF / POST kernels run 512 threads (128 registers each): residual rows are read in place
        constexpr int NCH = BLOCK_N / CH;  // column chunks of the tile; this warp owns chunks eg, eg + EG, ...
        constexpr bool SPLIT_SUB = NCH < EG;            // ... or, one-chunk tiles: sub-tiles eg, eg + EG, ... of the work item
        constexpr int CPW = SPLIT_SUB ? NCH : NCH / EG;  // chunks per warp
        uint4 resv[RES_PREFETCH ? MS_MAX : 1][RES_PREFETCH ? CPW * (CH / 8) : 1];
        const bool has_res = p.residual != nullptr && valid;
        if (RES_PREFETCH && has_res) {
#pragma unroll
          for (int sub = 0; sub < MS_MAX; ++sub) {
            if (sub < msub && (!SPLIT_SUB || (sub & (EG - 1)) == eg)) {
              const int64_t pix = (nn * p.H_full + p.out_scale * (h0 + sub * p.Hb + h_in) + out_oy) * p.W_full + p.out_scale * w_in + out_ox;
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.C_out + nt * BLOCK_N);
#pragma unroll
              for (int ci = 0; ci < CPW; ++ci)
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) resv[sub][ci * (CH / 8) + j] = __ldg(rp + (SPLIT_SUB ? ci : ci * EG + eg) * (CH / 8) + j);
            }
          }
        }
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
        int64_t pixs[MS_MAX];
#pragma unroll
        for (int sub = 0; sub < MS_MAX; ++sub)
          pixs[sub] = (nn * p.H_full + p.out_scale * (h0 + sub * p.Hb + h_in) + out_oy) * p.W_full + p.out_scale * w_in + out_ox;
        // GroupNorm statistics of the tensor being written (consumed by k_gn_apply / the fold kernel): per warp, per
        // channel QUAD, (sum, sum of squares) over the warp's 32 pixels (x msub sub-tiles) -- no second pass over the output
        const bool do_stats = CH == 32 && !SPLIT_SUB && p.stats != nullptr && valid;  // (N >= 64 only: conv_stats_parts)
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c0 = (SPLIT_SUB ? ci : ci * EG + eg) * CH;
          float st[CH / 2];
#pragma unroll
          for (int i = 0; i < CH / 2; ++i) st[i] = 0.f;
#pragma unroll
          for (int sub = 0; sub < MS_MAX; ++sub) {
            if (sub < msub && (!SPLIT_SUB || (sub & (EG - 1)) == eg)) {
              const uint32_t t_row = t_row0 + (uint32_t)(sub * BLOCK_N);
              uint32_t r[CH];
              if constexpr (CH == 32) tmem_ld_x32(t_row + c0, r);
              else tmem_ld_x16(t_row + c0, r);
              if (CH == 32 && p.tma_store) {  // the staging buffer is free once the previous box has been read out
                if (lane == 0) bulk_wait_read0();  // (waited for while the TMEM load is in flight)
                __syncwarp();
              }
              tmem_ld_wait();
              if (valid) {
                const int col = nt * BLOCK_N + c0;
                const uint32_t bias_s = smem_u32(s_bias) + (uint32_t)col * 4u;  // explicit shared-space address: LDS.128
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + pixs[sub] * p.C_out + col;
#pragma unroll
                for (int j = 0; j < CH; j += 8) {
                  // packed fp32x2 arithmetic (sm_100 FADD2 / FFMA2): half the issue slots of the bias / residual / statistics math
                  float2 v[4];
                  const float4 b0 = lds_f4(bias_s + j * 4), b1 = lds_f4(bias_s + j * 4 + 16);
                  v[0] = __fadd2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), make_float2(b0.x, b0.y));
                  v[1] = __fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), make_float2(b0.z, b0.w));
                  v[2] = __fadd2_rn(make_float2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])), make_float2(b1.x, b1.y));
                  v[3] = __fadd2_rn(make_float2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])), make_float2(b1.z, b1.w));
                  if (has_res) {
                    uint4 rv;
                    if constexpr (RES_PREFETCH) rv = resv[sub][ci * (CH / 8) + j / 8];
                    else rv = __ldg(reinterpret_cast<const uint4*>(p.residual + pixs[sub] * p.C_out + col + j));
                    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = __fadd2_rn(v[e], __bfloat1622float2(rp[e]));
                  }
                  uint4 o;
                  __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e) op[e] = __float22bfloat162_rn(v[e]);
                  if (CH == 32 && p.tma_store) sts_u4(my_row + (uint32_t)(((j >> 3) ^ ((lane >> 1) & 3)) << 4), o);
                  else *reinterpret_cast<uint4*>(dst + j) = o;
                  if (do_stats) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                      const int qi = (j / 4 + h) * 2;
                      const float2 s2 = __fadd2_rn(v[2 * h], v[2 * h + 1]);
                      const float2 q2 = __ffma2_rn(v[2 * h + 1], v[2 * h + 1], __fmul2_rn(v[2 * h], v[2 * h]));
                      st[qi] += s2.x + s2.y;
                      st[qi + 1] += q2.x + q2.y;
                    }
                  }
                }
              }
              if (CH == 32 && p.tma_store) {
                fence_proxy_async_smem();  // generic-proxy st.shared -> visible to the TMA (async proxy) read
                __syncwarp();
                if (lane == 0) {  // rows of images beyond the batch are clipped by the tensor map bounds
                  const void* src = reinterpret_cast<const void*>(epi_stage + (warp - 4) * kEpiStageBytes);
                  if (p.out_scale == 2)  // folded upsample: the output seen as [B*H_out][py][W_out][px][C], one parity per box
                    tma_store_5d(&tmO, src, nt * BLOCK_N + c0, out_ox, w_box, out_oy, (n0 + n_box) * p.H_out + h0 + sub * p.Hb + h_box);
                  else
                    tma_store_4d(&tmO, src, nt * BLOCK_N + c0, w_box, h0 + sub * p.Hb + h_box, n0 + n_box);
                  bulk_commit();
                }
              }
            }
          }
          if constexpr (CH == 32) {
            if (!SPLIT_SUB && p.stats != nullptr && p.stats_half) {
              // 4x4 maps: the warp's 32 tile rows are TWO images of 16 pixels -- the same transposing butterfly inside each
              // half warp (8+4+2+1 shuffles): lane l ends with the total of value l & 15 over its image
#pragma unroll
              for (int half = 8, off = 8; half >= 1; half >>= 1, off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                  const float send = up ? st[k] : st[k + half];
                  const float keep = up ? st[k + half] : st[k];
                  st[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              if (do_stats) {
                const int64_t row = nn * p.stats_parts + par;
                p.stats[(row * (p.C_out >> 2) + ((nt * BLOCK_N + c0) >> 2)) * 2 + (lane & 15)] = st[0];
              }
            } else if (!SPLIT_SUB && p.stats != nullptr) {  // warp-uniform (so is valid: a warp's 32 pixels belong to one image)
              // transposing butterfly: 16 values over 32 lanes in 8+4+2+1+1 shuffles; lane l ends with the total of value l>>1
#pragma unroll
              for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                  const float send = up ? st[k] : st[k + half];
                  const float keep = up ? st[k + half] : st[k];
                  st[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              st[0] += __shfl_xor_sync(0xffffffffu, st[0], 1);
              if (do_stats && (lane & 1) == 0) {
                const int unit = (mt / msub) % p.units_per_img;
                const int64_t row = (nn * p.stats_parts + (int64_t)(par * p.units_per_img + unit) * p.stats_wpi + (q % p.stats_wpi));
                p.stats[(row * (p.C_out >> 2) + ((nt * BLOCK_N + c0) >> 2)) * 2 + (lane >> 1)] = st[0];
              }
            }
          }
        }
      } else {
        // fp32 NCHW, first C_out_real channels of the (zero-padded) tile: the network's final conv (unet.py:435)
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < MS_MAX; ++sub) {
          if (sub < msub && (sub & (EG - 1)) == eg) {  // N = 16: one chunk, the quadrant's two warps take one sub-tile each
            uint32_t r[16];
            if constexpr (BLOCK_N == 48) {
              // dx-stacked: columns [0,16) hold tap dx = -1 evaluated AT this pixel, [16,32) dx = 0, [32,48) dx = +1; the output pixel w
              // needs the dx = -1 partial of pixel w - 1 and the dx = +1 partial of pixel w + 1 (same image row: a warp's lanes are
              // whole image rows, Wb <= 32), zero beyond the row ends (the horizontal zero padding)
              uint32_t rl[16], rh[16];
              tmem_ld_x16(t_row0 + (uint32_t)(sub * BLOCK_N), rl);
              tmem_ld_x16(t_row0 + (uint32_t)(sub * BLOCK_N + 16), r);
              tmem_ld_x16(t_row0 + (uint32_t)(sub * BLOCK_N + 32), rh);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                if (c < p.C_out_real) {  // (warp-uniform; the padded channels are never stored)
                  const float lo = __shfl_up_sync(0xffffffffu, __uint_as_float(rl[c]), 1);
                  const float hi = __shfl_down_sync(0xffffffffu, __uint_as_float(rh[c]), 1);
                  r[c] = __float_as_uint((w_in == 0 ? 0.f : lo) + __uint_as_float(r[c]) + (w_in == p.Wb - 1 ? 0.f : hi));
                }
              }
            } else {
              tmem_ld_x16(t_row0 + (uint32_t)(sub * BLOCK_N), r);
              tmem_ld_wait();
            }
            if (valid) {
              float* dst = reinterpret_cast<float*>(p.out);
              const int64_t hw = (int64_t)p.H_full * p.W_full;
              const int64_t sp = (int64_t)(p.out_scale * (h0 + sub * p.Hb + h_in) + out_oy) * p.W_full + p.out_scale * w_in + out_ox;
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < p.C_out_real) dst[(nn * p.C_out_real + c) * hw + sp] = __uint_as_float(r[c]) + __ldg(p.bias + c);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // relaxed: the barrier only hands the TMEM stage back (tcgen05.fence above); a release here would wait for every
        // outstanding global store of the warp to be acknowledged by L2, on the critical path of short-K layers
        if (CG == 2) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(tempty_bar + acc), 0));
        else mbar_arrive(tempty_bar + acc);
      }
      if (POST) {
        // this warp's share of the item is in global memory: raw bf16 rows (TMA stores: wait for the bulk group to COMPLETE,
        // not only for its shared-memory reads) and the statistics rows (plain stores of the lanes, ordered by __syncwarp
        // above + the release of the arrive).  One arrival per warp and item on the unit's "sample complete" barrier; the
        // slot is handed back by the post warps (at most two units in flight).
        const int su = it >> ipu_log, slot = su & 1;
        if (lane == 0) {
          if (p.tma_store && p.xf_dbg != 3) { bulk_wait0(); fence_proxy_async_all(); }
          if ((it & ((1 << ipu_log) - 1)) == 0) mbar_wait(free_bar + slot, ((su >> 1) & 1) ^ 1);
          mbar_arrive(ready_bar + slot);
        }
      }
    }
  } else if (POST && warp >= 12) {
    // ===================== post warps: GroupNorm (+ scale-shift, SiLU) of finished samples =====================
    // "GroupNorm in the producer's tail": the GroupNorms that consume this convolution's output (unet.py:141,153,188-191,
    // 212,433) need whole-sample statistics, so they cannot sit in the per-tile epilogue -- but a CTA that walks whole
    // samples can apply them as soon as the sample's last tile has been stored, while the raw rows are still in L2 and the
    // tensor cores work on the next sample.  Replaces the separate k_gn_apply / k_groupnorm_cluster launches: the raw tensor
    // is never re-read from HBM.
    // Work decomposition: a "pair" = (sample of the unit, 8-channel octet of the N tile); a thread owns one pair per round
    // (its affine coefficients live in registers) and every SLOTS-th pixel of it.  Everything a thread needs -- the partial
    // statistics rows, gamma / beta / scale / shift, the raw rows -- is fetched with independent loads issued together
    // (the latency of an L2 round trip under the convolution's own operand traffic is ~2 us: it must be paid once per
    // batch of loads, not once per sample or pixel).
    constexpr int OCT = BLOCK_N / 8;
    const int pt = (warp - 12) * 32 + lane;
    const int HWs = p.H_full * p.W_full;
    const int cq = p.C_out >> 2;
    const int n_pairs = p.Nb * OCT;
    const int ppr = n_pairs < kPostThreads ? n_pairs : kPostThreads;  // pairs per round
    const int SLOTS = kPostThreads / ppr, pslot = pt / ppr;
    for (int su = 0;; ++su) {
      ItemCoord ic;
      if (!item_coord<CG, POST>(p, su << ipu_log, first_item, item_stride, (int)cta_rank, msub, items_per_par, n_items, ic)) break;
      mbar_wait(ready_bar + (su & 1), (su >> 1) & 1);
      fence_proxy_async_all();
      for (int pair = pt % ppr; pair < n_pairs && p.xf_dbg != 1; pair += ppr) {  // (xf_dbg: timing experiments only)
        const int sI = pair / OCT, oct = pair - sI * OCT;
        const int64_t n = (int64_t)ic.n0 + sI;
        if (n >= p.B) continue;
        const int ch0 = ic.nt * BLOCK_N + oct * 8;  // first of this thread's eight conv output channels
        for (int k = 0; k < p.post_n; ++k) {
          const PostTarget& tg = p.post[k];
          // (1) group statistics straight from the partial rows (L2): this octet's two quads belong to one or two groups
          const int nqg = tg.cpg >> 2;  // quads per group
          float mean[2], rstd[2];
          const float inv_n = 1.0f / (float)(tg.cpg * HWs);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ct = tg.c_off + ch0 + 4 * h;             // consumer channel of this quad
            const int g0 = (ct / tg.cpg) * tg.cpg - tg.c_off;  // first conv channel of its group (inside this N tile: host-checked)
            if (h == 1 && nqg > 1) { mean[1] = mean[0]; rstd[1] = rstd[0]; break; }  // both quads in the same group
            const float2* row = reinterpret_cast<const float2*>(p.stats) + (n * p.stats_parts) * cq + (g0 >> 2);
            float sA = 0.f, qA = 0.f;
            for (int r = 0; r < p.stats_parts; ++r)
              for (int j = 0; j < nqg; ++j) { const float2 v = __ldcg(row + (int64_t)r * cq + j); sA += v.x; qA += v.y; }
            mean[h] = sA * inv_n;
            rstd[h] = rsqrtf(fmaxf(qA * inv_n - mean[h] * mean[h], 0.f) + 1e-5f);
          }
          // (2) y = a x + b: a = gamma rstd (1 + scale), b = (beta - mean gamma rstd)(1 + scale) + shift; halved for the SiLU form
          float a[8], b[8];
          const float* ssrow = tg.ss_off >= 0 ? p.ss + (p.ss_rows == 1 ? 0 : n * p.ss_stride) + tg.ss_off : nullptr;
          const float osc = tg.silu ? 0.5f : 1.0f;  // silu(y) = h + h tanh(h), h = y / 2
          {
            const int ct = tg.c_off + ch0;
            const float4 g0v = __ldg(reinterpret_cast<const float4*>(tg.gamma + ct)), g1v = __ldg(reinterpret_cast<const float4*>(tg.gamma + ct) + 1);
            const float4 b0v = __ldg(reinterpret_cast<const float4*>(tg.beta + ct)), b1v = __ldg(reinterpret_cast<const float4*>(tg.beta + ct) + 1);
            const float gg[8] = {g0v.x, g0v.y, g0v.z, g0v.w, g1v.x, g1v.y, g1v.z, g1v.w};
            const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float ga = gg[e] * rstd[e >> 2];
              float be = bb[e] - mean[e >> 2] * ga;
              if (ssrow) {
                const float sc = 1.0f + __ldg(ssrow + ct + e), sh = __ldg(ssrow + tg.dst_C + ct + e);
                ga *= sc;
                be = be * sc + sh;
              }
              a[e] = ga * osc;
              b[e] = be * osc;
            }
          }
          // (3) stream the sample: raw rows from L2 (ld.global.cg: written by this CTA a moment ago), normalised rows out;
          // software-pipelined, two batches of U loads in flight per thread
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.out) + (n * HWs) * p.C_out + ch0);
          uint4* dst = reinterpret_cast<uint4*>(tg.dst + (n * HWs) * tg.dst_C + tg.c_off + ch0);
          const int sstr = p.C_out >> 3, dstr = tg.dst_C >> 3;  // row strides in 16-byte units
          constexpr int U = 8;
          uint4 cur[U], nxt[U];
          int px = pslot;
#pragma unroll
          for (int u = 0; u < U; ++u) if (px + u * SLOTS < HWs) cur[u] = __ldcg(src + (int64_t)(px + u * SLOTS) * sstr);
          for (; px < HWs; px += U * SLOTS) {
            const int pn = px + U * SLOTS;
#pragma unroll
            for (int u = 0; u < U; ++u) if (pn + u * SLOTS < HWs) nxt[u] = __ldcg(src + (int64_t)(pn + u * SLOTS) * sstr);
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (px + u * SLOTS < HWs) {
                const uint4 o = post_apply8(cur[u], a, b, tg.silu);
                if (p.xf_dbg != 2 || o.x == 0x12345678u) dst[(int64_t)(px + u * SLOTS) * dstr] = o;
              }
#pragma unroll
            for (int u = 0; u < U; ++u) cur[u] = nxt[u];
          }
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kPostThreads) : "memory");
      if (pt == 0) mbar_arrive(free_bar + (su & 1));
    }
  }
  // the staging buffers must outlive the TMA engine's reads; the writes themselves are flushed by grid completion
  if (p.tma_store && warp >= 4 && warp < 4 + 4 * EG && lane == 0) bulk_wait_read0();
  tc_fence_before();
  __syncwarp();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int encode_act_map(CUtensorMap* m, const void* base, int64_t B, int H, int W, int C, int block_k, int Wb, int Hb, int Nb,
                          int stride) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DLPM_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)block_k, (cuuint32_t)(Wb * stride), (cuuint32_t)(Hb * stride), (cuuint32_t)Nb};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation [%lld,%d,%d,%d]) failed: %d", (long long)B, H, W, C, (int)r); return DLPM_ERR_CUDA; }
  return DLPM_OK;
}

static int encode_weight_map(CUtensorMap* m, const void* base, int rows, int64_t k_total, int block_n, int block_k) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DLPM_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights [%d,%lld]) failed: %d", rows, (long long)k_total, (int)r); return DLPM_ERR_CUDA; }
  return DLPM_OK;
}

static int encode_out_map(CUtensorMap* m, void* base, int64_t B, int H, int W, int C, int bw, int bh, int bi) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DLPM_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(output [%lld,%d,%d,%d]) failed: %d", (long long)B, H, W, C, (int)r); return DLPM_ERR_CUDA; }
  return DLPM_OK;
}

// Output of a folded upsample conv (pixels (2h + py, 2w + px)) as the 5-D view [B*H_out][py][W_out][px][C] of the dense
// [B][2 H_out][2 W_out][C] tensor: a parity's 32-pixel warp box is (32 ch, 1, bw, 1, rows) -- no element strides needed.
static int encode_out_map_up(CUtensorMap* m, void* base, int64_t B, int H_out, int W_out, int C, int bw, int rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DLPM_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W_out, 2, (cuuint64_t)(B * H_out)};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 4, (cuuint64_t)W_out * C * 4, (cuuint64_t)W_out * C * 8};
  cuuint32_t box[5] = {32, 1, (cuuint32_t)bw, 1, (cuuint32_t)rows};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(upsampled output [%lld,%d,%d,%d]) failed: %d", (long long)B, H_out, W_out, C, (int)r); return DLPM_ERR_CUDA; }
  return DLPM_OK;
}

static int g_tma_store_enabled = 1;
static int g_single_wave_cg1 = 1;  // "conv_single_wave_cg1"
static int g_tall256_enabled = 1;
static int g_xf_dbg = 0;  // timing experiments only: 1 = transform warps skip the math, 2 = skip the proxy fence
static int g_l2_prefetch = 0;  // measured: no gain (the three-stage ring already covers the HBM latency)

int conv_plan(ConvLaunch* L, const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1,
              int C_s1, const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in0, int C_out, ConvGeom geom,
              int stride, const ConvFuse* fuse, bool want_gne) {
  const int C_in2 = fuse ? fuse->C_in2 : 0;
  const int C_in = C_in0 + C_in2;  // K channels per tap: the main input may be the concatenation [in | fuse->in2]
  DLPM_REQUIRE(!fuse || (fuse->ab != nullptr && (fuse->in2 == nullptr) == (C_in2 == 0)), "conv: bad fusion descriptor");
  DLPM_REQUIRE(in && w && bias && out, "conv: NULL tensor");
  DLPM_REQUIRE(C_out <= kBiasSmemFloats, "conv: C_out must be <= 1024 (the bias vector is staged in shared memory)");
  DLPM_REQUIRE(C_in % 32 == 0 && C_in0 % 32 == 0, "conv: channel counts must be multiples of 32");
  DLPM_REQUIRE(geom.tap_rows >= 1 && geom.tap_rows <= 3 && geom.tap_cols >= 1 && geom.tap_cols <= 3, "conv: 1..3 taps per dimension");
  DLPM_REQUIRE(geom.out_scale == 1 || (geom.out_scale == 2 && stride == 1 && !skip0 && !skip1 && !residual &&
                                        out_mode == CONV_OUT_BF16_NHWC), "conv: strided output only for plain stride-1 convs");
  DLPM_REQUIRE(stride == 1 || stride == 2, "conv: stride must be 1 or 2");
  DLPM_REQUIRE(B >= 1 && H >= 1 && W >= 1, "conv: bad shape");
  DLPM_REQUIRE(H % stride == 0 && W % stride == 0, "conv: H, W must be divisible by the stride");
  DLPM_REQUIRE((skip0 == nullptr) == (C_s0 == 0) && (skip1 == nullptr) == (C_s1 == 0), "conv: skip source / channel mismatch");
  DLPM_REQUIRE(!(stride == 2 && (skip0 || skip1 || residual)), "conv: skip fusion needs stride 1");
  const int H_out = H / stride, W_out = W / stride;
  DLPM_REQUIRE(W_out <= 128 && (W_out & (W_out - 1)) == 0 && (H_out & (H_out - 1)) == 0,
               "conv: output H and W must be powers of two with W <= 128");
  const int bk = (C_in0 % 64 == 0 && C_in2 % 64 == 0 && C_s0 % 64 == 0 && C_s1 % 64 == 0) ? 64 : 32;
  if (C_in % bk || C_s0 % bk || C_s1 % bk) { set_error("conv: channel counts must be multiples of 32 (got %d,%d,%d)", C_in, C_s0, C_s1); return DLPM_ERR_UNSUPPORTED; }
  // n_par == 3: "dx-stacked" thin convolution (ConvGeom): weights [3 x 16][3 * C_in], one N tile of 48 columns
  const bool dxs = geom.n_par == 3;
  DLPM_REQUIRE(!dxs || (out_mode == CONV_OUT_F32_NCHW && geom.tap_rows == 3 && geom.tap_cols == 3 && geom.dy0 == -1 && geom.dx0 == -1 &&
                        stride == 1 && geom.out_scale == 1 && !skip0 && !skip1 && !residual && (!fuse || !fuse->in2) && W <= 32 && W % 8 == 0 && H * W >= 128 &&
                        conv_tall_enabled()),
               "conv: the dx-stacked form is for the thin fp32 output conv (3x3, stride 1, rows of <= 32 pixels)");
  const int C_out_pad = dxs ? 48 : (out_mode == CONV_OUT_F32_NCHW ? 16 : C_out);
  int bn;
  if (dxs) { DLPM_REQUIRE(C_out <= 16, "conv: fp32 NCHW output supports <= 16 channels"); bn = 48; }
  else if (out_mode == CONV_OUT_F32_NCHW) { DLPM_REQUIRE(C_out <= 16, "conv: fp32 NCHW output supports <= 16 channels"); bn = 16; }
  else if (C_out % 256 == 0) bn = 256;
  else if (C_out % 128 == 0) bn = 128;
  else if (C_out % 64 == 0) bn = 64;
  else if (C_out % 32 == 0) bn = 32;
  else if (C_out % 16 == 0) bn = 16;
  else { set_error("conv: C_out must be a multiple of 16 (got %d)", C_out); return DLPM_ERR_UNSUPPORTED; }
  L->block_k = bk;
  L->Wb = W_out;
  L->Hb = (128 / W_out) < H_out ? (128 / W_out) : H_out;
  L->Nb = 128 / (L->Wb * L->Hb);
  {  // small problems (4x4 / 8x8 feature maps): prefer more, narrower tiles so that every SM gets work
    const int64_t m_tiles = (L->Nb == 1 ? B * (H_out / L->Hb) : (B + L->Nb - 1) / L->Nb) * (geom.n_par == 4 ? 4 : 1);
    // (not when the GroupNorm is to run in the epilogue: a GNE kernel's N tile is the whole channel range)
    const bool gne_shape = want_gne && L->Nb == 1 && H_out * W_out == 256 && geom.n_par != 4 && C_out_pad <= 256;
    while (!gne_shape && bn > 128 && m_tiles * (C_out_pad / bn) < kNumSMs && C_out_pad % (bn / 2) == 0) bn /= 2;
  }
  L->block_n = bn;
  L->H_out = H_out; L->W_out = W_out;
  L->tiles_per_img = L->Nb == 1 ? H_out / L->Hb : 0;
  L->n_m_tiles = L->Nb == 1 ? (int)(B * L->tiles_per_img) : (int)((B + L->Nb - 1) / L->Nb);
  L->n_n_tiles = C_out_pad / bn;
  L->stride = stride; L->taps = dxs ? 3 : geom.tap_rows * geom.tap_cols;  // (dx-stacked: K = 3 vertical taps x C_in)
  L->dx_taps = dxs ? 1 : 3;
  L->tap_cols = geom.tap_cols; L->dy0 = geom.dy0; L->dx0 = geom.dx0;
  L->out_scale = geom.out_scale; L->out_oy = geom.out_oy; L->out_ox = geom.out_ox;
  L->H_full = H_out * geom.out_scale; L->W_full = W_out * geom.out_scale;
  L->n_par = geom.n_par == 4 ? 4 : 1;
  DLPM_REQUIRE(geom.n_par == 1 || geom.n_par == 3 || geom.n_par == 4, "conv: n_par must be 1, 3 (dx-stacked) or 4 (folded upsample)");
  DLPM_REQUIRE(L->n_par == 1 || (geom.out_scale == 2 && geom.tap_rows == 2 && geom.tap_cols == 2),
               "conv: parity batching is for folded upsample convs (2x2 taps, output scale 2)");
  L->cin_blocks = C_in / bk; L->s0_blocks = C_s0 / bk; L->s1_blocks = C_s1 / bk;
  L->B = B; L->C_out = C_out; L->C_out_real = C_out; L->out_mode = out_mode;
  L->bias = bias; L->residual = reinterpret_cast<const __nv_bfloat16*>(residual); L->out = out;
  // "tall" activation boxes: 3x3 stride-1 convs whose tile lies inside one image.  N = 256 tiles only fit the shared-memory
  // budget as CTA pairs (each CTA stages half of the three weight tiles), so the flavour is decided before the tensor maps.
  const bool tall_geom = L->Nb == 1 && geom.tap_rows == 3 && geom.tap_cols == 3 && geom.dy0 == -1 && geom.dx0 == -1 && stride == 1 &&
                         geom.out_scale == 1 && L->Wb % 8 == 0 && conv_tall_enabled();
  // two vertically adjacent sub-tiles per CTA when the image has an even number of tiles and TMEM has room (N <= 128)
  L->msub = (tall_geom && L->tiles_per_img % 2 == 0 && bn <= 128 && conv_msub_enabled()) ? 2 : 1;
  // CTA pairs (cta_group::2) when there are enough M tiles to keep all 74 pairs busy
  L->cta_group = (conv_cta_group_override() == 1) ? 1
                 : ((bn >= 32 && (int64_t)((L->n_m_tiles / L->msub + 1) / 2) * L->n_n_tiles * L->n_par >= 32) ? 2 : 1);
  if (conv_cta_group_override() == 2 && bn >= 32) L->cta_group = 2;
  // ... but not when single CTAs already fit one wave with one work item each (4x4 maps at batch 512: 128 tiles): a pair then saves
  // nothing it could amortise (half the weight tile of ONE item) and pays the cluster launch, the paired TMEM allocation handshake
  // and the remote barrier hops: measured 20.4 -> 18.3 us per 256 -> 256 layer at 4x4
  if (g_single_wave_cg1 && conv_cta_group_override() == 0 && (int64_t)(L->n_m_tiles / L->msub) * L->n_n_tiles * L->n_par <= kNumSMs)
    L->cta_group = 1;
  if (dxs) L->cta_group = 1;
  // fused GroupNorm targets on a map of 256 pixels: CTA pairs whatever the batch, so that the pair's accumulator stage holds the
  // whole sample and the GroupNorm runs in the epilogue (GNE) instead of the post warps
  if (want_gne && conv_cta_group_override() != 1 && L->Nb == 1 && L->tiles_per_img == 2 && L->msub == 1 && bn >= 128 && geom.n_par != 4) L->cta_group = 2;
  L->tall = (tall_geom && (bn <= 128 || (L->cta_group == 2 && g_tall256_enabled))) ? 1 : 0;
  L->xf = fuse ? 1 : 0;
  L->ab = fuse ? fuse->ab : nullptr;
  L->c0_blocks = C_in0 / bk;
  if (fuse && !(L->tall && bk == 64 && (bn == 16 || bn == 48 || bn == 128 || bn == 256))) {
    set_error("conv: normalise-on-load needs a 3x3 stride-1 conv with one image per tile row block, 64-channel K blocks, N in {16,128,256}");
    return DLPM_ERR_UNSUPPORTED;
  }
  int rc;
  if ((rc = encode_act_map(&L->tmA, in, B, H, W, C_in0, bk, L->Wb, L->tall ? L->msub * L->Hb + 2 : L->Hb, L->Nb, stride))) return rc;
  L->tmA2 = L->tmA;
  if (C_in2 && (rc = encode_act_map(&L->tmA2, fuse->in2, B, H, W, C_in2, bk, L->Wb, L->msub * L->Hb + 2, L->Nb, stride))) return rc;
  L->tmS0 = L->tmA; L->tmS1 = L->tmA;
  if (skip0 && (rc = encode_act_map(&L->tmS0, skip0, B, H_out, W_out, C_s0, bk, L->Wb, L->msub * L->Hb, L->Nb, 1))) return rc;
  if (skip1 && (rc = encode_act_map(&L->tmS1, skip1, B, H_out, W_out, C_s1, bk, L->Wb, L->msub * L->Hb, L->Nb, 1))) return rc;
  const int64_t k_total = (int64_t)L->taps * C_in + C_s0 + C_s1;
  L->c_out_pad = C_out_pad;
  L->stats = nullptr;
  L->post_n = 0;
  L->gne = 0; L->raw_unused = 0;
  L->reverse = 0;
  L->ss = nullptr; L->ss_rows = 1; L->ss_stride = 0;
  if ((rc = encode_weight_map(&L->tmB, w, C_out_pad * L->n_par, k_total, bn / L->cta_group, bk))) return rc;
  // TMA-store epilogue: dense bf16 NHWC outputs with 32-channel chunks; an epilogue warp's 32 tile rows are a (bw, bh, bn) pixel box
  L->tmO = L->tmB;
  L->tma_store = (out_mode == CONV_OUT_BF16_NHWC && bn >= 32 && !fuse && g_tma_store_enabled && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                  B * H_out < (1ll << 31)) ? 1 : 0;
  if (L->tma_store) {
    const int bw = L->Wb < 32 ? L->Wb : 32, bh = L->Hb < 32 / bw ? L->Hb : 32 / bw, bi = 32 / (bw * bh);
    if (geom.out_scale == 2) rc = encode_out_map_up(&L->tmO, out, B, H_out, W_out, C_out, bw, bh * bi);
    else rc = encode_out_map(&L->tmO, out, B, H_out, W_out, C_out, bw, bh, bi);
    if (rc) return rc;
  }
  return DLPM_OK;
}

// Partial-statistics rows per image this conv can emit for the GroupNorm that consumes its output (0 = not supported):
// one row per epilogue warp (4) per work unit of the image per output parity.
int conv_stats_parts(const ConvLaunch& L) {
  if (L.out_mode != CONV_OUT_BF16_NHWC || L.block_n < 64 || L.C_out % 4) return 0;  // one-chunk tiles split sub-tiles over the epilogue warps
  if (L.Nb == 1) return 4 * (L.tiles_per_img / L.msub) * L.n_par;
  // several images per tile: every epilogue warp (32 pixels) lies inside one image, or (4x4 maps) holds exactly two images
  if (L.Wb * L.Hb == 16) return L.n_par;
  return (L.Wb * L.Hb) % 32 == 0 ? ((L.Wb * L.Hb) / 32) * L.n_par : 0;
}

bool conv_post_capable(const ConvLaunch& L) {
  if (L.out_mode != CONV_OUT_BF16_NHWC || L.n_par != 1 || L.out_scale != 1 || L.xf || L.block_n < 128) return false;
  if (conv_stats_parts(L) <= 0 || L.C_out % 8) return false;
  if (L.Nb == 1) {
    const int ipu = L.tiles_per_img / L.msub;
    if (ipu < 1 || (ipu & (ipu - 1)) || ipu * L.msub != L.tiles_per_img) return false;
  }
  return true;
}

int conv_set_post(ConvLaunch* L, int n, const ConvLaunch::Post* targets, const float* ss, int64_t ss_stride) {
  DLPM_REQUIRE(n >= 1 && n <= 2 && targets, "conv_set_post: 1 or 2 targets");
  if (!conv_post_capable(*L)) { set_error("conv: this shape cannot apply the consumer's GroupNorm (N tile %d, out mode %d, parities %d)", L->block_n, L->out_mode, L->n_par); return DLPM_ERR_UNSUPPORTED; }
  for (int k = 0; k < n; ++k) {
    const ConvLaunch::Post& t = targets[k];
    DLPM_REQUIRE(t.dst && t.gamma && t.beta, "conv_set_post: NULL target tensor");
    if (t.cpg < 4 || t.cpg % 4 || L->block_n % t.cpg || t.c_off % t.cpg || t.c_off % 8 || t.dst_C % 8 || t.c_off + L->C_out > t.dst_C ||
        (reinterpret_cast<uintptr_t>(t.dst) & 15u)) {
      set_error("conv_set_post: group size %d / channel offset %d / row width %d do not fit N tiles of %d channels", t.cpg, t.c_off, t.dst_C, L->block_n);
      return DLPM_ERR_UNSUPPORTED;
    }
    DLPM_REQUIRE(t.ss_off < 0 || ss != nullptr, "conv_set_post: scale-shift columns without a table");
    L->post[k] = t;
  }
  L->post_n = n;
  L->ss = ss;
  L->ss_stride = ss_stride;
  // GroupNorm in the epilogue (GNE) instead of the post warps when the pair's accumulator stage holds the whole sample
  L->gne = 0;
  if (const int mode = conv_gne_capable(*L, n)) {
    const int bw = L->Wb < 32 ? L->Wb : 32, bh = L->Hb < 32 / bw ? L->Hb : 32 / bw, bi = 32 / (bw * bh);
    for (int k = 0; k < n; ++k)
      if (int rc = encode_out_map(&L->tmP[k], L->post[k].dst, L->B, L->H_out, L->W_out, L->post[k].dst_C, bw, bh, bi)) return rc;
    L->gne = mode;
  }
  return DLPM_OK;
}

static int g_gne_enabled = 7;  // bit 0: mode 1 (maps of 256 pixels, CTA pairs), bit 1: mode 2 on 4x4 maps, bit 2: mode 2 on 8x8 maps
// 0 = not capable; 1 = the pair's accumulator stage holds one whole sample (two tiles per image, CTA pairs): two TMEM passes + DSMEM
// exchange; 2 = 4x4 maps (eight samples per tile, a sample = half a warp): one pass, statistics by half-warp shuffles
int conv_gne_capable(const ConvLaunch& L, int n_targets) {
  if (!g_gne_enabled || !conv_post_capable(L) || !L.tma_store || (L.block_n != 128 && L.block_n != 256) || L.block_k != 64) return 0;
  for (int k = 0; k < n_targets; ++k)
    if (32 % L.post[k].cpg || L.post[k].c_off % L.post[k].cpg) return 0;
  if ((g_gne_enabled & 2) && L.Nb > 1 && (L.Wb * L.Hb == 16 || (L.Wb * L.Hb == 64 && (g_gne_enabled & 4)))) {
    // bias | A, B per target (all N tiles) | 8x8: the exchange slots of the eight epilogue warps
    if ((1 + 2 * n_targets) * L.c_out_pad * 4 + (L.Wb * L.Hb == 64 ? 512 : 0) > kGneRegionBytes) return 0;
    return 2;
  }
  if (!(g_gne_enabled & 1) || L.cta_group != 2 || L.Nb != 1 || L.msub != 1 || L.tiles_per_img != 2 || L.n_n_tiles != 1) return 0;
  // bias | A, B per target | exchanged totals | group (rstd, -mean rstd)
  if ((3 + 2 * n_targets) * L.block_n * 4 + n_targets * L.block_n * 2 > kGneRegionBytes) return 0;
  for (int k = 0; k < n_targets; ++k)
    if (!L.post[k].silu) return 0;
  return 1;
}

static int g_tall_enabled = 1;
int conv_tall_enabled() { return g_tall_enabled; }
static int g_msub_enabled = 1;
int conv_msub_enabled() { return g_msub_enabled; }
static int g_cta_group_override = -1;
int conv_cta_group_override() {
  if (g_cta_group_override < 0) {
    const char* e = getenv("DLPM_B200_CTA_GROUP");  // 1 / 2 force the MMA flavour (debugging, A/B measurements); unset = auto
    g_cta_group_override = e ? atoi(e) : 0;
  }
  return g_cta_group_override;
}
void conv_set_cta_group_override(int v) { g_cta_group_override = v; }

static int64_t g_n_conv = 0, g_n_post = 0, g_n_gne = 0;

template <int BN, int BK, int CG, bool XF, bool POST = false, int GNE = 0>
static int launch_t(const ConvLaunch& L, cudaStream_t stream) {
  ++g_n_conv;
  if (POST) ++g_n_post;
  if (GNE) ++g_n_gne;
  constexpr int KS = BN <= 128 ? 2 : 1;
  const int B_BYTES = (BN / CG) * BK * 2;
  const int STAGE = L.tall ? (L.msub * L.Hb + 2) * L.Wb * BK * 2 + 3 * B_BYTES : KS * (128 * BK * 2 + B_BYTES);
  int stages = kSmemBudget / STAGE;
  if (stages > kMaxStages) stages = kMaxStages;
  const size_t smem = (size_t)stages * STAGE + kSmemExtra;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc<BN, BK, CG, KS, XF, POST, GNE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + kSmemExtra));
    if (e != cudaSuccess) return cuda_fail(e, "conv smem attribute");
    attr_set = true;
  }
  ConvKParams p;
  p.n_m_tiles = L.n_m_tiles; p.n_n_tiles = L.n_n_tiles; p.stages = stages;
  p.Wb = L.Wb; p.Hb = L.Hb; p.Nb = L.Nb; p.H_out = L.H_out; p.W_out = L.W_out; p.tiles_per_img = L.tiles_per_img;
  p.stride = L.stride; p.taps = L.taps; p.cin_blocks = L.cin_blocks; p.s0_blocks = L.s0_blocks; p.s1_blocks = L.s1_blocks;
  p.tap_cols = L.tap_cols; p.dy0 = L.dy0; p.dx0 = L.dx0; p.out_scale = L.out_scale; p.out_oy = L.out_oy; p.out_ox = L.out_ox;
  p.H_full = L.H_full; p.W_full = L.W_full;
  p.l2_prefetch = g_l2_prefetch; p.xf_dbg = g_xf_dbg;
  p.tma_store = L.tma_store;
  p.dx_taps = L.dx_taps;
  p.tall = L.tall; p.stage_bytes = STAGE; p.n_par = L.n_par; p.c_out_pad = L.c_out_pad; p.msub = L.msub;
  p.B = L.B; p.C_out = L.C_out; p.C_out_real = L.C_out_real; p.out_mode = L.out_mode;
  p.bias = L.bias; p.residual = L.residual; p.out = L.out;
  p.ab = reinterpret_cast<const float2*>(L.ab); p.c0_blocks = L.c0_blocks; p.c_in_total = L.cin_blocks * BK;
  p.stats = L.stats; p.stats_parts = conv_stats_parts(L); p.units_per_img = L.tiles_per_img > 0 ? L.tiles_per_img / L.msub : 1;
  p.stats_wpi = L.Nb == 1 ? 4 : (L.Wb * L.Hb) / 32;
  {
    const int ipp = ((L.n_m_tiles / L.msub + CG - 1) / CG) * L.n_n_tiles;
    p.fd_ipp = FastDiv((uint32_t)(ipp > 0 ? ipp : 1));
    p.fd_nnt = FastDiv((uint32_t)(L.n_n_tiles > 0 ? L.n_n_tiles : 1));
    p.fd_tpi = FastDiv((uint32_t)(L.tiles_per_img > 0 ? L.tiles_per_img : 1));
  }
  int n_items = ((L.n_m_tiles / L.msub + CG - 1) / CG) * L.n_n_tiles * L.n_par;
  p.reverse = L.reverse;
  p.post_n = 0; p.ipu_log = 0; p.n_units = 0; p.ss = L.ss; p.ss_rows = L.ss_rows; p.ss_stride = L.ss_stride;
  p.stats_half = (L.Nb > 1 && L.Wb * L.Hb == 16) ? 1 : 0;
  p.gne_raw = L.raw_unused ? 0 : 1;
  if (GNE && ((stages * STAGE) % 1024 != 0)) { set_error("conv: GNE kernels need an operand ring that is a multiple of 1024 bytes"); return DLPM_ERR_UNSUPPORTED; }
  if (POST || GNE) {
    if ((POST && L.stats == nullptr) || L.post_n < 1) { set_error("conv: POST launch without statistics buffer / targets"); return DLPM_ERR_ARG; }
    p.post_n = L.post_n;
    for (int k = 0; k < L.post_n; ++k) {
      p.post[k].dst = reinterpret_cast<__nv_bfloat16*>(L.post[k].dst); p.post[k].gamma = L.post[k].gamma; p.post[k].beta = L.post[k].beta;
      p.post[k].ss_off = L.post[k].ss_off; p.post[k].dst_C = L.post[k].dst_C; p.post[k].c_off = L.post[k].c_off;
      p.post[k].cpg = L.post[k].cpg; p.post[k].silu = L.post[k].silu;
    }
    if (POST && L.Nb == 1) {  // sample-major walk: a unit = CG samples x one N tile, ipu items each
      int ipu = L.tiles_per_img / L.msub, lg = 0;
      while ((1 << lg) < ipu) ++lg;
      p.ipu_log = lg;
      p.n_units = (int)((L.B + CG - 1) / CG) * L.n_n_tiles;
      n_items = p.n_units;  // (grid sizing below: one CTA group per unit at most)
    }
  }
  const int max_groups = kNumSMs / CG;
  const int grid = (n_items < max_groups ? n_items : max_groups) * CG;
  const int threads = conv_threads<BN, XF, POST>();
  cudaError_t e = launch_ex(k_conv_tc<BN, BK, CG, KS, XF, POST, GNE>, dim3(grid), dim3(threads), smem, stream, CG, L.tmA, L.tmA2, L.tmS0, L.tmS1,
                            L.tmB, L.tmO, L.tmP[0], L.tmP[1], p);
  if (e != cudaSuccess) return cuda_fail(e, CG == 1 ? "conv_tc launch" : "conv_tc pair launch");
  return DLPM_OK;
}

// mode-2 GNE keeps ONE scale / shift table per kernel: a forward with per-sample time steps (ss_rows == B: eight samples with
// different rows in one tile) goes through the post warps, which read the rows per sample
static bool gne2_needs_post_warps(const ConvLaunch& L) {
  if (L.ss_rows == 1) return false;
  for (int k = 0; k < L.post_n; ++k)
    if (L.post[k].ss_off >= 0) return true;
  return false;
}

int conv_launch(const ConvLaunch& L, cudaStream_t stream) {
#define CASE(BN, BK)                                                                \
  if (L.block_n == BN && L.block_k == BK) {                                         \
    if (L.xf) {                                                                     \
      if constexpr (BK == 64 && (BN == 16 || BN == 128 || BN == 256)) {             \
        if (L.cta_group == 2) {                                                     \
          if constexpr (BN >= 32) return launch_t<BN, BK, 2, true>(L, stream);      \
        }                                                                           \
        return launch_t<BN, BK, 1, true>(L, stream);                                \
      }                                                                             \
      set_error("conv: no normalise-on-load kernel for tile N=%d K=%d", BN, BK);   \
      return DLPM_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    if (L.post_n > 0 && L.gne == 1) {                                               \
      if constexpr (BN >= 128 && BK == 64) return launch_t<BN, BK, 2, false, false, 1>(L, stream); \
      set_error("conv: no GroupNorm-in-the-epilogue kernel for tile N=%d K=%d", BN, BK); \
      return DLPM_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    if (L.post_n > 0 && L.gne == 2 && !gne2_needs_post_warps(L)) {                  \
      if constexpr (BN >= 128 && BK == 64) {                                        \
        if (L.cta_group == 2) return launch_t<BN, BK, 2, false, false, 2>(L, stream); \
        return launch_t<BN, BK, 1, false, false, 2>(L, stream);                     \
      }                                                                             \
      set_error("conv: no GroupNorm-in-the-epilogue kernel for tile N=%d K=%d", BN, BK); \
      return DLPM_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    if (L.post_n > 0) {                                                             \
      if constexpr (BN >= 128) {                                                    \
        if (L.cta_group == 2) return launch_t<BN, BK, 2, false, true>(L, stream);   \
        return launch_t<BN, BK, 1, false, true>(L, stream);                         \
      }                                                                             \
      set_error("conv: no producer-side GroupNorm kernel for tile N=%d", BN);       \
      return DLPM_ERR_UNSUPPORTED;                                                  \
    }                                                                               \
    if (L.cta_group == 2) {                                                         \
      if constexpr (BN >= 32) return launch_t<BN, BK, 2, false>(L, stream);         \
    }                                                                               \
    return launch_t<BN, BK, 1, false>(L, stream);                                   \
  }
  CASE(256, 64) CASE(128, 64) CASE(64, 64) CASE(32, 64) CASE(16, 64)
  CASE(256, 32) CASE(128, 32) CASE(64, 32) CASE(32, 32) CASE(16, 32)
  if (L.block_n == 48 && L.post_n == 0 && L.cta_group == 1) {  // dx-stacked thin conv (optionally normalising its input on load)
    if (L.xf && L.block_k == 64) return launch_t<48, 64, 1, true>(L, stream);
    if (!L.xf && L.block_k == 64) return launch_t<48, 64, 1, false>(L, stream);
    if (!L.xf && L.block_k == 32) return launch_t<48, 32, 1, false>(L, stream);
  }
#undef CASE
  set_error("conv: no kernel for tile N=%d K=%d", L.block_n, L.block_k);
  return DLPM_ERR_UNSUPPORTED;
}

}  // namespace dlpm

namespace dlpm { void engine_set_loop_ss_table(bool on); void engine_set_gne_skip_raw(bool on); void engine_set_traverse_alternate(bool on); void attention_set_mma(int on); void attention_set_poly(int v); void gn_apply_set_min_elems(int v); void engine_set_gn_stats(bool on); void engine_set_gn_fuse(int mode); bool process_set_option(const char* name, int value); }
using namespace dlpm;

int dlpm_b200_get_stat(const char* name, int64_t* value) {
  DLPM_REQUIRE(name != nullptr && value != nullptr, "get_stat: NULL argument");
  const std::string n(name);
  if (n == "conv_launches") *value = g_n_conv;
  else if (n == "conv_post_launches") *value = g_n_post;
  else if (n == "conv_gne_launches") *value = g_n_gne;
  else { set_error("get_stat: unknown counter '%s'", name); return DLPM_ERR_ARG; }
  return DLPM_OK;
}

int dlpm_b200_set_option(const char* name, int value) {
  DLPM_REQUIRE(name != nullptr, "set_option: NULL name");
  if (std::string(name) == "pdl") {
    pdl_set_enabled(value != 0);
    return DLPM_OK;
  }
  if (std::string(name) == "gn_fuse") {
    engine_set_gn_fuse(value);
    return DLPM_OK;
  }
  if (std::string(name) == "traverse_alternate") {
    engine_set_traverse_alternate(value != 0);
    return DLPM_OK;
  }
  if (std::string(name) == "gn_stats") {
    engine_set_gn_stats(value != 0);
    return DLPM_OK;
  }
  if (std::string(name) == "loop_ss_table") {  // graph_sample: scale / shift rows of all steps computed once per call (default 1)
    engine_set_loop_ss_table(value != 0);
    return DLPM_OK;
  }
  if (std::string(name) == "gne_skip_raw") {
    engine_set_gne_skip_raw(value != 0);
    return DLPM_OK;
  }
  if (std::string(name) == "conv_gne") {  // GroupNorm in the epilogue for fused targets: bit 0 = 16x16 maps, bit 1 = 4x4 maps, bit 2 = 8x8 maps (0: always the post warps)
    g_gne_enabled = value == 1 ? 7 : value;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_msub") {
    g_msub_enabled = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "xf_dbg") {
    g_xf_dbg = value;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_l2_prefetch") {
    g_l2_prefetch = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "attention_mma") {
    attention_set_mma(value);
    return DLPM_OK;
  }
  if (std::string(name) == "gn_apply_min_elems") {
    gn_apply_set_min_elems(value);
    return DLPM_OK;
  }
  if (std::string(name) == "attention_poly") {
    attention_set_poly(value);
    return DLPM_OK;
  }
  if (std::string(name) == "conv_single_wave_cg1") {
    g_single_wave_cg1 = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_tma_store") {
    g_tma_store_enabled = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_tall256") {
    g_tall256_enabled = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_tall") {
    g_tall_enabled = value != 0;
    return DLPM_OK;
  }
  if (std::string(name) == "conv_cta_group") {
    DLPM_REQUIRE(value >= 0 && value <= 2, "set_option: conv_cta_group must be 0 (auto), 1 or 2");
    conv_set_cta_group_override(value);
    return DLPM_OK;
  }
  if (process_set_option(name, value)) return DLPM_OK;
  set_error("set_option: unknown option '%s'", name);
  return DLPM_ERR_ARG;
}

int dlpm_b200_conv2d(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1, int C_s1,
                     const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in, int C_out, int ksize,
                     int stride, void* stream) {
  DLPM_REQUIRE(ksize == 3 || ksize == 1, "conv: kernel size must be 1 or 3");
  ConvLaunch L;
  if (int rc = conv_plan(&L, in, w, bias, skip0, C_s0, skip1, C_s1, residual, out, out_mode, B, H, W, C_in, C_out,
                         conv_geom_default(ksize), stride, nullptr))
    return rc;
  return conv_launch(L, (cudaStream_t)stream);
}

int dlpm_b200_conv2d_stats(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1, int C_s1,
                           const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in, int C_out, int ksize,
                           int stride, float* stats, int* stats_parts, void* stream) {
  DLPM_REQUIRE(ksize == 3 || ksize == 1, "conv: kernel size must be 1 or 3");
  ConvLaunch L;
  if (int rc = conv_plan(&L, in, w, bias, skip0, C_s0, skip1, C_s1, residual, out, out_mode, B, H, W, C_in, C_out,
                         conv_geom_default(ksize), stride, nullptr))
    return rc;
  const int parts = conv_stats_parts(L);
  if (stats_parts) *stats_parts = parts;
  if (stats == nullptr) return stats_parts ? DLPM_OK : conv_launch(L, (cudaStream_t)stream);
  DLPM_REQUIRE(parts > 0, "conv2d_stats: this shape cannot emit GroupNorm statistics");
  L.stats = stats;
  return conv_launch(L, (cudaStream_t)stream);
}

int dlpm_b200_conv2d_post(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1, int C_s1,
                          const void* residual, void* out, int64_t B, int H, int W, int C_in, int C_out, int ksize, int stride, float* stats,
                          int* stats_parts, void* post_dst, int dst_C, int c_off, int cpg, const float* gamma, const float* beta,
                          const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, void* stream) {
  DLPM_REQUIRE(ksize == 3 || ksize == 1, "conv: kernel size must be 1 or 3");
  ConvLaunch L;
  if (int rc = conv_plan(&L, in, w, bias, skip0, C_s0, skip1, C_s1, residual, out, CONV_OUT_BF16_NHWC, B, H, W, C_in, C_out,
                         conv_geom_default(ksize), stride, nullptr, true))
    return rc;
  const int parts = conv_stats_parts(L);
  if (stats_parts) *stats_parts = parts;
  if (stats == nullptr) {
    if (stats_parts) return conv_post_capable(L) ? DLPM_OK : DLPM_ERR_UNSUPPORTED;
    set_error("conv2d_post: the statistics buffer is required");
    return DLPM_ERR_ARG;
  }
  L.stats = stats;
  ConvLaunch::Post t;
  t.dst = post_dst; t.gamma = gamma; t.beta = beta; t.ss_off = ss ? ss_off : -1; t.dst_C = dst_C; t.c_off = c_off; t.cpg = cpg;
  t.silu = apply_silu;
  if (int rc = conv_set_post(&L, 1, &t, ss, ss_stride)) return rc;
  L.ss_rows = ss_rows;
  return conv_launch(L, (cudaStream_t)stream);
}

int dlpm_b200_conv2d_gn(const void* in, const void* in2, int C_in2, const float* ab, const void* w, const float* bias, const void* skip0,
                        int C_s0, const void* skip1, int C_s1, const void* residual, void* out, int out_mode, int64_t B, int H, int W,
                        int C_in, int C_out, float* stats, int* stats_parts, void* stream) {
  DLPM_REQUIRE(ab != nullptr, "conv2d_gn: NULL coefficient table");
  ConvLaunch L;
  const ConvFuse fuse{in2, C_in2, ab};
  if (int rc = conv_plan(&L, in, w, bias, skip0, C_s0, skip1, C_s1, residual, out, out_mode, B, H, W, C_in, C_out, conv_geom_default(3), 1,
                         &fuse))
    return rc;
  const int parts = conv_stats_parts(L);
  if (stats_parts) *stats_parts = parts;
  if (stats) {
    DLPM_REQUIRE(parts > 0, "conv2d_gn: this shape cannot emit GroupNorm statistics");
    L.stats = stats;
  }
  return conv_launch(L, (cudaStream_t)stream);
}
