// K5: implicit-GEMM convolution on tcgen05 tensor cores (host-side interface).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dlpm {

enum ConvOutMode { CONV_OUT_BF16_NHWC = 0, CONV_OUT_F32_NCHW = 1 };

// Everything the launcher needs for one convolution (tensor maps are built once per plan).
struct ConvLaunch {
  CUtensorMap tmA, tmA2, tmS0, tmS1, tmB;  // main activation (+ second half of a concatenated input), two optional 1x1
                                           // skip-conv sources, packed weights
  CUtensorMap tmO;                   // output tensor, box = one epilogue warp's 32 pixels x 32 channels (tma_store)
  int tma_store;                     // 1 = the epilogue stages bf16 chunks in shared memory and stores them with TMA
  int block_n, block_k;              // tile N (16..256), K block in channels (64 or 32)
  int msub;                          // sub-tiles (128 pixels each) per CTA: 2 in tall mode when the image has an even tile count
  int tall;                          // 1 = one (Hb+2)-row activation box per (channel block, dx) feeds the three dy taps
  int cta_group;                     // 1 = single-CTA MMA, 2 = CTA pairs (tcgen05 cta_group::2, M = 256)
  int n_m_tiles, n_n_tiles;
  int Wb, Hb, Nb;                    // pixel box of one M tile: Wb*Hb*Nb == 128
  int H_out, W_out, tiles_per_img;
  int stride, taps;
  int tap_cols, dy0, dx0;            // tap t reads the input shifted by (dy0 + t / tap_cols, dx0 + t % tap_cols)
  int out_scale, out_oy, out_ox;     // output pixel of tile pixel (h, w): (out_scale*h + out_oy, out_scale*w + out_ox)
  int H_full, W_full;                // spatial size of the output TENSOR (= out_scale * H_out, W_out)
  int n_par, c_out_pad;              // parity batching (folded upsample): 4 weight matrices stacked along rows
  int dx_taps;                       // 3, or 1 = dx-stacked thin conv (ConvGeom::n_par == 3)
  int cin_blocks, s0_blocks, s1_blocks;
  int64_t B;
  int C_out;       // row stride (channels) of the output / residual tensors
  int C_out_real;  // channels actually written (conv_out: 3 of the padded 16)
  int out_mode;
  const float* bias;                // [n_n_tiles * block_n] fp32
  const __nv_bfloat16* residual;    // NHWC bf16 [B, H_out, W_out, C_out] or nullptr
  void* out;
  int xf, c0_blocks;                // normalise-on-load (GroupNorm + SiLU applied to the activation boxes in shared memory)
  const float* ab;                  // its coefficient table [B][C_in][2] = (a/2, b/2)
  float* stats;                     // optional GroupNorm partial sums of the output (see conv_stats_parts), set by the caller
  int reverse;                      // 1: walk the work items (samples) in descending order.  Consecutive kernels of the UNet
                                    // alternate direction so that a consumer starts with the rows its producer wrote LAST -- the
                                    // part of a > L2-sized tensor that is still resident (an LRU cache walked in the same
                                    // direction twice hits nothing)
  // POST ("GroupNorm in the producer's tail", conv_set_post): the GroupNorms that consume this output are applied by the
  // convolution's own post warps as soon as a sample is complete; needs `stats`
  int post_n;
  struct Post {
    void* dst;            // consumer's normalised input, NHWC bf16 [B, H_out, W_out, dst_C]
    const float* gamma;   // consumer GroupNorm parameters, indexed by the consumer's channel (c_off + c)
    const float* beta;
    int64_t ss_off;       // scale / shift columns of the time-embedding table (scale at ss_off + ch, shift at ss_off + dst_C + ch); < 0: none
    int dst_C, c_off, cpg, silu;
  } post[2];
  // GNE ("GroupNorm in the epilogue"): chosen by conv_set_post when a CTA pair's accumulator stage holds one whole sample (maps of
  // 256 pixels): the targets are applied straight from TMEM in a second epilogue pass, no post warps, no re-read through L2
  int gne;
  int raw_unused;                   // 1: nothing reads the raw output (its only consumer is a fused GroupNorm): GNE kernels skip the store
  CUtensorMap tmP[2];               // GNE: the targets' tensors, same pixel box as tmO
  const float* ss;                  // time-embedding table [ss_rows][ss_stride]; ss_rows (1 or B) is patched per forward
  int ss_rows;
  int64_t ss_stride;
};

// Attach up to two fused GroupNorm targets to a planned convolution.  DLPM_ERR_UNSUPPORTED when the shape cannot carry them
// (output not bf16 NHWC, folded-upsample launches, N tiles below 128 channels, groups that are not whole quads of this
// convolution's channels or straddle an N tile).
int conv_set_post(ConvLaunch* L, int n, const ConvLaunch::Post* targets, const float* ss, int64_t ss_stride);
bool conv_post_capable(const ConvLaunch& L);
int conv_gne_capable(const ConvLaunch& L, int n_targets);  // 0 / mode 1 (16x16, CTA pairs) / mode 2 (4x4)

// Fills geometry + tensor maps.  in: NHWC bf16 [B, H, W, C_in]; w: bf16 [C_out_pad][taps*C_in + C_s0 + C_s1] (K contiguous);
// skip sources NHWC bf16 [B, H_out, W_out, C_s*].  Returns DLPM_OK or an error code (message via set_error).
// Tap geometry.  Default (ksize x ksize taps centred on the pixel, dense output) = a plain convolution; the UNet's
// "nearest-upsample x2 then conv3x3" (unet.py:73-75) is run as four 2x2-tap convolutions on the LOW-resolution input,
// one per output parity (out_scale = 2, out_oy/out_ox = parity), with pre-summed weights: 2.25x fewer FLOPs and no
// upsampled tensor in memory.
struct ConvGeom {
  int tap_rows, tap_cols, dy0, dx0, out_scale, out_oy, out_ox;
  int n_par;  // 3: "dx-stacked" thin fp32-output conv (<= 16 channels, rows of <= 32 pixels): weights [3 (dx) x 16][3 (dy) x C_in], i.e. row
              // dx * 16 + co, column dy * C_in + c = w[co][c][dy][dx]; one unshifted activation box per channel block feeds all three
              // horizontal taps (N = 48) and the epilogue sums the partials of neighbouring pixels;
              // 4: run the four parity convs of a folded upsample in one launch (weights stacked [4][C_out][4*C_in]; parity
              // (py, px) uses taps (dy0 + py + a, dx0 + px + b) and writes output pixels (2h + py, 2w + px)); else 1
};
inline ConvGeom conv_geom_default(int ksize) { return ConvGeom{ksize, ksize, -(ksize / 2), -(ksize / 2), 1, 0, 0, 1}; }

// Normalise-on-load fusion: the conv reads the RAW GroupNorm input(s) [in | in2] and applies silu(a*x+b) per (sample, channel)
// while the activation boxes sit in shared memory (coefficients from dlpm_b200_groupnorm_fold with half = 1).
struct ConvFuse {
  const void* in2;  // second half of the concatenated main input (NHWC bf16) or nullptr
  int C_in2;
  const float* ab;  // [B][C_in + C_in2][2]
};

int conv_plan(ConvLaunch* L, const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1,
              int C_s1, const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in, int C_out, ConvGeom geom,
              int stride, const ConvFuse* fuse = nullptr, bool want_gne = false);
int conv_launch(const ConvLaunch& L, cudaStream_t stream);
int conv_stats_parts(const ConvLaunch& L);
int conv_cta_group_override();
int conv_tall_enabled();
int conv_msub_enabled();

}  // namespace dlpm
