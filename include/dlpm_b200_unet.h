/* dlpm_b200_unet.h -- C ABI of the image score network (K5-K7) in libdlpm_b200.so.
 *
 * Replaces UNetModel.forward (dlpm/models/unet.py:463-492) and its sub-modules.  Activations are
 * NHWC bf16 in HBM, accumulation and normalisation statistics are fp32.  Same conventions as
 * dlpm_b200.h (device pointers, void* stream, 0 / negative return, dlpm_b200_last_error()).
 */
#ifndef DLPM_B200_UNET_H_
#define DLPM_B200_UNET_H_
#include <stdint.h>

#include "dlpm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DLPM_CONV_OUT_BF16_NHWC 0
#define DLPM_CONV_OUT_F32_NCHW 1

/* K5. 3x3 (pad 1) / 1x1 convolution, stride 1 or 2, as an implicit GEMM on tcgen05 tensor cores
 * (nn.Conv2d at unet.py:64,96,143,157,164-168,214-215,347,435).
 *   in        NHWC bf16 [B, H, W, C_in]                       (C_in multiple of 32)
 *   w         bf16 [C_out_pad][ksize*ksize*C_in + C_s0 + C_s1], K index = (ky*ksize+kx)*C_in + c, then the
 *             1x1 skip-conv weights for skip0 and skip1 (ResBlock.skip_connection fused into the K loop)
 *   bias      fp32 [C_out_pad]
 *   skip0/1   optional NHWC bf16 [B, H/stride, W/stride, C_s*] inputs of the fused 1x1 skip conv (or NULL, 0)
 *   residual  optional NHWC bf16 [B, H/stride, W/stride, C_out] added in the epilogue (identity skip)
 *   out       DLPM_CONV_OUT_BF16_NHWC: bf16 [B, H/stride, W/stride, C_out] (C_out multiple of 16)
 *             DLPM_CONV_OUT_F32_NCHW : fp32 [B, C_out, H, W], C_out <= 16 and w/bias zero-padded to 16 rows
 */
int dlpm_b200_conv2d(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1,
                     int C_s1, const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in,
                     int C_out, int ksize, int stride, void* stream);

/* K5 with the GroupNorm fusion hooks the UNet engine uses.  Same convolution as dlpm_b200_conv2d, plus:
 *   stats_parts  if non-NULL, receives the number of partial-statistics rows per image this shape emits (0 = the shape
 *                cannot emit statistics: tiles spanning several images, fp32 output, C_out < 32)
 *   stats        NULL, or fp32 [B][*stats_parts][C_out/4][2]: the epilogue writes, per row and per QUAD of output
 *                channels, (sum, sum of squares) of the fp32 results it is storing -- the input of
 *                dlpm_b200_groupnorm_from_stats / dlpm_b200_groupnorm_fold, so that the GroupNorm that follows the
 *                convolution (unet.py:141,153,433) needs no statistics pass over the tensor.
 * Call once with stats == NULL to size the buffer. */
int dlpm_b200_conv2d_stats(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1,
                           int C_s1, const void* residual, void* out, int out_mode, int64_t B, int H, int W, int C_in,
                           int C_out, int ksize, int stride, float* stats, int* stats_parts, void* stream);

/* K5 with the GroupNorm of its CONSUMER applied by the convolution itself ("GroupNorm in the producer's tail"): besides
 * `out` (raw bf16 NHWC, as dlpm_b200_conv2d) and the statistics rows (as dlpm_b200_conv2d_stats) the kernel's post warps
 * write  post_dst[b, y, x, c_off + c] = act( GN(out)[b, y, x, c] * (1 + scale) + shift )  as soon as a sample is complete,
 * reading the raw rows back from L2 -- the separate GroupNorm pass over HBM (unet.py:141,153,188-191,212,433) disappears.
 *   post_dst   NHWC bf16 [B, H/stride, W/stride, dst_C]; this convolution fills channels [c_off, c_off + C_out) (the two
 *              halves of a skip concatenation are filled by their two producers)
 *   cpg        channels per group of the consumer's GroupNorm = its total channels / 32; multiple of 4, divides 128 and c_off
 *   gamma/beta fp32 [dst_C] of the consumer; ss / ss_rows / ss_stride / ss_off as in dlpm_b200_groupnorm_silu (NULL: none)
 * C_out must be a multiple of 128.  Call with stats == NULL and stats_parts != NULL to size the statistics buffer
 * (fp32 [B][*stats_parts][C_out/4][2]); returns DLPM_ERR_UNSUPPORTED for shapes that cannot carry the fusion. */
int dlpm_b200_conv2d_post(const void* in, const void* w, const float* bias, const void* skip0, int C_s0, const void* skip1, int C_s1,
                          const void* residual, void* out, int64_t B, int H, int W, int C_in, int C_out, int ksize, int stride, float* stats,
                          int* stats_parts, void* post_dst, int dst_C, int c_off, int cpg, const float* gamma, const float* beta,
                          const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, void* stream);

/* Tuning / debugging knobs.  "conv_cta_group": 0 = automatic (CTA pairs with tcgen05 cta_group::2 when the problem has
 * enough tiles), 1 = always single-CTA MMAs, 2 = always CTA pairs.  "conv_tall": 1 (default) lets 3x3 stride-1 convs with
 * narrow output tiles load one (rows+2)-tall activation box per horizontal tap and reuse it for the three vertical taps,
 * 0 loads one box per tap.  "attention_mma": 1 (default) = tensor-core attention kernel, 0 = FMA kernel.  "attention_poly":
 * 1 evaluates a quarter of the softmax exponentials on the FMA pipes (measured slower; default -1 / 0 = all on MUFU.EX2).
 * Takes effect for descriptors built afterwards. */
int dlpm_b200_set_option(const char* name, int value);
/* Launch counters since library load (tests / bench use them to prove which kernel flavour ran): "conv_launches",
 * "conv_post_launches" (GroupNorm by the post warps), "conv_gne_launches" (GroupNorm in the epilogue, straight from TMEM). */
int dlpm_b200_get_stat(const char* name, int64_t* value);

/* K6. GroupNorm(min(32,C) groups, eps 1e-5) over the virtual concatenation [in0 | in1] of two NHWC bf16
 * tensors, optional scale-shift conditioning y = GN(x) * (1 + scale) + shift, optional SiLU; bf16 NHWC out
 * (GroupNorm32 + SiLU + use_scale_shift_norm of unet.py:141-142,153-154,188-191,212,433-434; nn.py:17-19).
 *   gamma, beta   fp32 [C0 + C1];   ss: NULL or fp32 [ss_rows][ss_stride] with scale at [ss_off, ss_off+C) and
 *   shift at [ss_off+C, ss_off+2C); ss_rows is 1 (batch-constant timestep) or B. */
int dlpm_b200_groupnorm_silu(void* out, const void* in0, int C0, const void* in1, int C1, int64_t B, int HW,
                             const float* gamma, const float* beta, const float* ss, int ss_rows, int64_t ss_stride,
                             int64_t ss_off, int apply_silu, void* stream);

/* K6 from convolution-epilogue statistics (dlpm_b200_conv2d_stats): same result as dlpm_b200_groupnorm_silu over
 * [in0 | in1], but the statistics come from stats0 / stats1 (partial rows of the two producers, parts0 / parts1 rows per
 * image) and the tensor is streamed exactly once.  Total channels must be a multiple of 128 (group = multiple of 4). */
int dlpm_b200_groupnorm_from_stats(void* out, const void* in0, int C0, const float* stats0, int parts0, const void* in1, int C1,
                                   const float* stats1, int parts1, int64_t B, int HW, const float* gamma, const float* beta,
                                   const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, void* stream);

/* Coefficient table only: ab fp32 [B][C0+C1][2] with GN(x)*(1+scale)+shift == ab[.,c,0] * x + ab[.,c,1]; half != 0
 * stores both coefficients multiplied by 1/2 (the form dlpm_b200_conv2d_gn consumes). */
int dlpm_b200_groupnorm_fold(float* ab, int C0, const float* stats0, int parts0, int C1, const float* stats1, int parts1, int64_t B,
                             int HW, const float* gamma, const float* beta, const float* ss, int ss_rows, int64_t ss_stride,
                             int64_t ss_off, int half, void* stream);

/* K5 + K6 fused ("normalise on load"): out = conv3x3(SiLU(GroupNorm([in | in2]))) (+ fused 1x1 skip conv / residual / statistics
 * as in dlpm_b200_conv2d_stats) without materialising the normalised tensor: the activation boxes are rewritten in shared
 * memory between their TMA arrival and the MMAs (unet.py:141-143,153-157,188-191,433-435).
 *   in, in2   RAW GroupNorm inputs, NHWC bf16 [B, H, W, C_in] and [B, H, W, C_in2] (in2 may be NULL / 0)
 *   ab        fp32 [B][C_in + C_in2][2] from dlpm_b200_groupnorm_fold(..., half = 1)
 *   w         bf16 [C_out_pad][9*(C_in + C_in2) + C_s0 + C_s1]
 * Only 3x3 stride-1 convolutions whose 128-pixel tiles lie inside one image (H*W >= 128, W a multiple of 8), channel
 * counts multiples of 64 and C_out in {<=16 (fp32 NCHW out), multiples of 128}; DLPM_ERR_UNSUPPORTED otherwise. */
int dlpm_b200_conv2d_gn(const void* in, const void* in2, int C_in2, const float* ab, const void* w, const float* bias,
                        const void* skip0, int C_s0, const void* skip1, int C_s1, const void* residual, void* out, int out_mode,
                        int64_t B, int H, int W, int C_in, int C_out, float* stats, int* stats_parts, void* stream);

/* K7. QKVAttention (unet.py:231-250): qkv NHWC bf16 [B, L, 3C] with the reference's channel order (per head:
 * q, k, v blocks of C/heads channels), out NHWC bf16 [B, L, C].  L <= 1024, C/heads <= 64. */
int dlpm_b200_attention(void* out, const void* qkv, int64_t B, int L, int C, int heads, void* stream);

/* Input conv (unet.py:347): x NCHW fp32 [B, C_in<=4, H, W] -> NHWC bf16 [B, H, W, C_out] (C_out multiple of 32);
 * wT fp32 in-major [C_in*9][C_out] (= conv weight [C_out][C_in][3][3] reshaped to [C_out][C_in*9] and transposed). */
int dlpm_b200_conv_in(void* out, const float* x, const float* wT, const float* bias, int64_t B, int C_in, int C_out, int H,
                      int W, void* stream);

/* Tensor-core form of the input conv (unet.py:347), step 1: x NCHW fp32 [B, C, H, W] (3*C <= 32) -> NHWC bf16
 * [B, H, W, 32] with channels [0,C) = bf16(x), [C,2C) = bf16(x - bf16(x)), [2C,3C) = bf16(x), rest 0.  Step 2 is
 * dlpm_b200_conv2d over these 32 channels with weights packed (w_hi, w_hi, w_lo): x*w to 2^-16 relative. */
int dlpm_b200_split_input(void* out, const float* x, int64_t B, int C, int H, int W, void* stream);

/* Input conv that also leaves GroupNorm partial statistics of its output (same layout and protocol as
 * dlpm_b200_conv2d_stats: fp32 [B][*stats_parts][C_out/4][2]; stats == NULL with stats_parts != NULL only sizes). */
int dlpm_b200_conv_in_stats(void* out, const float* x, const float* wT, const float* bias, int64_t B, int C_in, int C_out,
                            int H, int W, float* stats, int* stats_parts, void* stream);

/* Nearest x2 upsample (unet.py:73), NHWC bf16 [B,H,W,C] -> [B,2H,2W,C]. */
int dlpm_b200_upsample2x(void* out, const void* in, int64_t B, int H, int W, int C, void* stream);

/* Timestep embedding + every ResBlock's emb_layers in two launches (nn.py:103-121; unet.py:335-339,145-151,477):
 *   semb = SiLU(time_embed(timestep_embedding(t, mc)))  [rows][4mc];  ss = semb @ W_all^T + b_all  [rows][ss_total].
 * t: device float[rows]; or NULL with t_dev (device int*): t = *t_dev * inv_T when inv_T > 0 (discrete DLPM steps), or
 * t = t[*t_dev] when t is non-NULL as well (table of continuous times, LIM) -- both for CUDA-graph replay.
 * Weights in-major fp32: w0T [mc][4mc], b0, w2T [4mc][4mc], b2, wallT [4mc][ss_total], ball. */
int dlpm_b200_time_embedding(float* ss, float* semb, const float* t, const int* t_dev, float inv_T, int rows, int mc,
                             int64_t ss_total, const float* w0T, const float* b0, const float* w2T, const float* b2,
                             const float* wallT, const float* ball, void* stream);

/* ---- whole network behind one handle ------------------------------------------------------------
 * The Python host (dlpm_b200/score_nets.py::UNetModel) walks the architecture (unet.py:343-437) and hands
 * over an op list + two packed weight blobs; the engine owns activation buffers and TMA descriptors.
 *   header  int64[16] : {n_ops, n_bufs, in_ch, out_ch, H, W, model_channels, ss_total,
 *                        p_w0T, p_b0, p_w2T, p_b2, p_wallT, p_ball, 0, 0}   (p_* = fp32-blob offsets)
 *   ops     int64[n_ops][24], bufs int64[n_bufs] = bf16 elements per sample of each activation buffer
 *   wb      bf16 blob (conv weights, device), wf fp32 blob (everything else, device); both are copied. */
int dlpm_b200_unet_create(void** handle, const int64_t* header, const int64_t* ops, const int64_t* bufs, const void* wb,
                          int64_t n_wb, const float* wf, int64_t n_wf, int64_t max_batch);
/* eps = UNet(x, t): x fp32 NCHW [B,in_ch,H,W]; t device float[t_rows] (t_rows = 1: batch-constant, or B), or
 * t_dev (+ inv_T, or + t as a time table) as above; out fp32 NCHW [B,out_ch,H,W]. */
int dlpm_b200_unet_forward(void* handle, const float* x, const float* t, int t_rows, const int* t_dev, float inv_T,
                           float* out, int64_t B, void* stream);
/* The captured reverse loop: the whole of p_sample_loop_progressive / ddim_sample_loop_progressive
 * (GenerativeLevyProcess.py:291-330, :413-452) or LIM_sampler (LIM/functions/sampler.py:218-258) for the image net in ONE
 * call.  A step = [optional input scaling] -> UNet forward (time from a device-side counter) -> fused update -> counter
 * +-1; it is captured with cudaStreamBeginCapture on `stream` into a CUDA graph, the engine's cached executable graph is
 * updated in place (cudaGraphExecUpdate; instantiated only when the topology changes) and launched once per step.
 * Asynchronous; x (fp32 [B, C, H, W], in place) holds the final sample when the stream has drained.
 *   mode DLPM_LOOP_DLPM : x_{T-1} -> x_0, T - 1 steps; Sigma compact (T, B) or full (flag DLPM_STEP_SIGMA_FULL), sched (T, 4),
 *                         aux = NULL or the device table 1/(1 + barsigma_t) [T] of `input_scaling` (GenerativeLevyProcess.py:177-180)
 *   mode DLPM_LOOP_DLIM : deterministic eta = 0 steps; Sigma unused
 *   mode DLPM_LOOP_LIM_SDE / _ODE : T steps; sched = coefficient rows [T][4] of dlpm_b200_lim_step, aux = continuous times [T];
 *                         isotropic / alpha / clamp_eps as in dlpm_b200_lim_step
 * flags: DLPM_STEP_CLIP_DENOISED, DLPM_STEP_SIGMA_FULL.  seed / offset / sample_base key the in-kernel noise exactly as
 * in the step functions (z of step t at offset + t).  post: NULL or the fused post-processing of the final sample. */
#define DLPM_LOOP_DLPM 0
#define DLPM_LOOP_DLIM 1
#define DLPM_LOOP_LIM_SDE 2
#define DLPM_LOOP_LIM_ODE 3
int dlpm_b200_graph_sample(void* handle, int mode, float* x, const float* Sigma, const float* sched, const float* aux, int T,
                           int64_t B, int flags, int isotropic, float alpha, float clamp_eps, uint64_t seed, uint64_t offset,
                           int64_t sample_base, const dlpm_b200_post_t* post, void* stream);
/* how often dlpm_b200_graph_sample had to instantiate an executable graph / could update the cached one in place */
int dlpm_b200_graph_sample_stats(void* handle, int* instantiations, int* updates);
/* debugging / per-layer parity: copy activation buffer `buf` (bf16, B * elems) of the last forward to dst. */
int dlpm_b200_unet_copy_buffer(void* handle, int buf, void* dst, int64_t B, void* stream);
/* measurement: one forward with a CUDA event after every op; ms_per_op / flops_per_op are HOST arrays of n_ops + 1
 * entries (entry 0 = the two time-embedding launches); flops = 2*M*N*K of the convolutions, 0 for the other ops. */
int dlpm_b200_unet_profile(void* handle, const float* x, const float* t, int t_rows, float* out, int64_t B, float* ms_per_op,
                           double* flops_per_op, void* stream);
int64_t dlpm_b200_unet_workspace_bytes(void* handle);
int dlpm_b200_unet_num_launches(void* handle);
int dlpm_b200_unet_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* DLPM_B200_UNET_H_ */
