/* dlpm_b200.h -- C ABI of libdlpm_b200.so (hand-written sm_100a CUDA behind plain pointers).
 *
 * Drop-in boundary for the DLPM sampling hot path.  Every entry point
 * takes DEVICE pointers + sizes + a cudaStream_t (as void*), launches asynchronously on that
 * stream, allocates nothing visible to the caller unless stated, and returns 0 on success or a
 * negative error code (text via dlpm_b200_last_error()).  No torch types cross this boundary.
 * The Python host (dlpm_b200/_lib.py, ctypes) mirrors the reference's `dlpm/methods` API on top.
 * `file:line` citations are relative to the reference tree (darioShar/DLPM).
 *
 * RNG contract: Philox4x32-R (Salmon et al. 2011) with R = dlpm_b200_philox_rounds() = 7 in the default build (the fewest
 * rounds that pass BigCrush, Table 2 of the paper; -DDLPM_PHILOX_ROUNDS=10 rebuilds with the Random123 / cuRAND default);
 * a variate is a pure function of (seed, stream tag, offset, GLOBAL sample index = sample_base + local index, position in
 * sample), so any sharding of the batch over GPUs reproduces the same numbers.  The reference draws from numpy's / torch's
 * global generators, so no reference stream is pinned by this choice; oracle/philox.py restates the generator with the
 * Random123 known-answer vectors for both round counts.
 */
#ifndef DLPM_B200_H_
#define DLPM_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLPM_B200_ABI_VERSION 1
#define DLPM_OK 0
#define DLPM_ERR_ARG (-1)      /* invalid argument */
#define DLPM_ERR_CUDA (-2)     /* CUDA runtime / driver error */
#define DLPM_ERR_UNSUPPORTED (-3)

int dlpm_b200_abi_version(void);
const char* dlpm_b200_last_error(void);

/* Layout modes for the subordinator A ~ S(alpha/2, 1) scaled as in Distributions.py:45,48. */
#define DLPM_A_COMPACT 0   /* out[n_outer]:           one draw per sample                          */
#define DLPM_A_ISOTROPIC 1 /* out[n_outer * inner]:   one draw per sample, replicated (match_last_dims, :9-28,:46) */
#define DLPM_A_FULL 2      /* out[n_outer * inner]:   independent draw per element (isotropic=False, :48) */

/* K1a. gen_skewed_levy (bem/datasets/Distributions.py:33-51).  alpha in (0,2]; alpha==2 -> A == 2.
 * clamp_a < 0 means "None" (no clamp), else A is clamped to [0, clamp_a] (:49-50). */
int dlpm_b200_stable_A(float* out, int64_t n_outer, int64_t inner, int mode, float alpha, float clamp_a,
                       uint64_t seed, uint64_t offset, int64_t sample_base, void* stream);

/* K1b. gen_sas (Distributions.py:57-73): out = scale * clamp(sqrt(A) * G, +-clamp_eps), G ~ N(0, I).
 * A_in: NULL -> A is drawn in-kernel WITHOUT clamp_a (reference quirk, :64), per sample when
 * isotropic != 0 else per element; otherwise compact A[n_outer] (isotropic) or full A[n_outer*inner].
 * clamp_eps < 0 means None.  scale folds `barsigmas[-1] * gen_eps.generate()` of
 * GenerativeLevyProcess.py:313. */
int dlpm_b200_sas(float* out, const float* A_in, int64_t n_outer, int64_t inner, int isotropic, float alpha,
                  float clamp_eps, float scale, uint64_t seed, uint64_t offset, int64_t sample_base, void* stream);

/* Standard normal fill with the same counter layout (stream tag Z); used for training z_t
 * (GenerativeLevyProcess.py:652) and by tests. */
int dlpm_b200_normal(float* out, int64_t n_outer, int64_t inner, uint64_t seed, uint64_t offset,
                     int64_t sample_base, void* stream);

/* K2. DLPM.sample_A + DLPM.compute_Sigmas (dlpm/methods/dlpm.py:226-239) for the isotropic case on
 * compact (T, n) tables:  Sigma_0 = s_0^2 A_0,  Sigma_t = s_t^2 A_t + g_t^2 Sigma_{t-1}.
 * sched: device float[T*4], row t = (gamma_t, bargamma_t, sigma_t, barsigma_t) (dlpm.py:114-156).
 * A_in NULL -> A_t drawn in-kernel (offset + t keys the step), clamped to [0,clamp_a] if clamp_a>=0.
 * A_out optional (NULL to skip): receives the A table that was used.
 * For isotropic=False pass n = B*D (element-wise chains) and per_element=1 so the Philox position
 * is derived per element. */
int dlpm_b200_sigma_scan(float* Sigma, const float* A_in, float* A_out, const float* sched, int T, int64_t n,
                         int64_t inner, int per_element, float alpha, float clamp_a, uint64_t seed,
                         uint64_t offset, int64_t sample_base, void* stream);

/* Flags for the fused step kernels. */
#define DLPM_STEP_CLIP_DENOISED 1 /* GenerativeLevyProcess.py:186-207: eps <- (x - bg clamp((x - bs eps)/bg,-1,1))/bs */
#define DLPM_STEP_EPS_BF16 2      /* eps tensor is bf16 (network output), else fp32 */
#define DLPM_STEP_SIGMA_FULL 4    /* Sigma is (T, B*D) (isotropic=False) instead of compact (T, B) */

/* K3. One stochastic DLPM reverse step, in place on x (B, D) fp32
 * (p_sample GenerativeLevyProcess.py:225-239 + anterior_mean_variance_dlpm dlpm.py:250-278):
 *   Gamma = 1 - g_t^2 Sigma_{t-1}/Sigma_t ;  x <- (x - bs_t Gamma eps)/g_t + 1[t != 1] sqrt(Gamma Sigma_{t-1}) z
 * z: injected Gaussian (parity tests) or NULL -> drawn in-kernel (offset keys the step).
 * t_dev: optional device int*; when non-NULL the step index is read from *t_dev (CUDA-graph replay)
 * and `t` is ignored.  hist_out optional: also stores the new x there (get_sample_history). */
int dlpm_b200_reverse_step(float* x, const void* eps, const float* Sigma, const float* sched, int t,
                           const int* t_dev, int T, int64_t B, int64_t D, int flags, const float* z,
                           uint64_t seed, uint64_t offset, int64_t sample_base, float* hist_out, void* stream);

/* Post-processing of the FINAL sample fused into the last step's store (bem/GenerationManager.py:50-63: clamp to +-1
 * (images) / +-6 (2-D data), images then mapped by (x+1)/2; DLPM_POST_U8_NHWC additionally quantises like
 * torchvision.utils.save_image -- x*255 + 0.5, clamp to [0,255], truncate -- into uint8 [B, H*W, C]).  The step kernel
 * writes `out` only when it executes the step whose result is final (t == 1; LIM: the last step), so the same argument
 * can be baked into a CUDA graph that is replayed for every step.  x itself always receives the unprocessed x_0. */
#define DLPM_POST_NONE 0
#define DLPM_POST_F32 1        /* out fp32 [B, D] = clamp(x, +-clamp)                 (16-byte aligned) */
#define DLPM_POST_F32_IMAGE 2  /* out fp32 [B, D] = (clamp(x, +-clamp) + 1) / 2 */
#define DLPM_POST_U8_NHWC 3    /* out uint8 [B, D/channels, channels] of the image form */
typedef struct dlpm_b200_post {
  void* out;
  float clamp;
  int mode;
  int channels; /* DLPM_POST_U8_NHWC only: D = channels * H * W */
} dlpm_b200_post_t;
int dlpm_b200_reverse_step_post(float* x, const void* eps, const float* Sigma, const float* sched, int t, const int* t_dev,
                                int T, int64_t B, int64_t D, int flags, const float* z, uint64_t seed, uint64_t offset,
                                int64_t sample_base, float* hist_out, const dlpm_b200_post_t* post, void* stream);
int dlpm_b200_dlim_step_post(float* x, const void* eps, const float* sched, int t, const int* t_dev, int T, int64_t B,
                             int64_t D, int flags, float* hist_out, const dlpm_b200_post_t* post, void* stream);
/* last_step: the step index whose result is the final sample (n_steps - 1). */
int dlpm_b200_lim_step_post(float* x, const void* model_out, const float* coef, int step, const int* step_dev, int64_t B,
                            int64_t D, int flags, int ode, int isotropic, float alpha, float clamp_eps, const float* e_L,
                            uint64_t seed, uint64_t offset, int64_t sample_base, float* hist_out, const dlpm_b200_post_t* post,
                            int last_step, void* stream);

/* K3'. Deterministic DLIM step, eta = 0 (anterior_mean_variance_dlim dlpm.py:281-287):
 *   x <- (x - bs_t eps)/g_t + bs_{t-1} eps. */
int dlpm_b200_dlim_step(float* x, const void* eps, const float* sched, int t, const int* t_dev, int T,
                        int64_t B, int64_t D, int flags, float* hist_out, void* stream);

/* K3''. LIM continuous-time steps (dlpm/methods/LIM/functions/sampler.py:81-181); per-step scalars are
 * computed by the host from the VPSDE (sde.py:35-47) and are batch-constant (sampler.py:229):
 *   SDE: x <- a x + c_score (score_scale * out) + c_noise * e_L,   e_L = clamp(sqrt(A) G) drawn in-kernel
 *        (isotropic: one A per sample) or injected via e_L (parity tests);
 *   ODE: same with c_noise = 0 and no noise.
 * coef: device float[4*n_steps] rows (score_scale, a, c_score, c_noise); step index from `step`
 * or *step_dev. */
int dlpm_b200_lim_step(float* x, const void* model_out, const float* coef, int step, const int* step_dev,
                       int64_t B, int64_t D, int flags, int ode, int isotropic, float alpha, float clamp_eps,
                       const float* e_L, uint64_t seed, uint64_t offset, int64_t sample_base, float* hist_out,
                       void* stream);

/* Device-side step counter helpers for graph replay: *t_dev += delta / *t_dev = value (single thread). */
int dlpm_b200_advance_counter(int* t_dev, int delta, void* stream);
int dlpm_b200_set_counter(int* t_dev, int value, void* stream);

/* Rounds of the Philox4x32 generator this build uses (compile-time DLPM_PHILOX_ROUNDS; 7 by default, see RNG contract). */
int dlpm_b200_philox_rounds(void);

/* Training forward elements, Proposition (9) (dlpm.py:384-401, GenerativeLevyProcess.py:634-661):
 *   x_t = bg_t x0 + sqrt(A bs_t^2) z ;  eps_t = (x_t - bg_t x0)/bs_t,   per-sample t (int64, B).
 * A compact (B) or NULL -> drawn in-kernel (clamp_a applies); z injected or NULL -> in-kernel. */
int dlpm_b200_training_elements(float* x_t, float* eps_t, const float* x0, const int64_t* t, const float* A,
                                const float* z, const float* sched, int T, int64_t B, int64_t D, float alpha,
                                float clamp_a, uint64_t seed, uint64_t offset, int64_t sample_base, void* stream);

/* Model-input scaling of the scale_exploding schedule (GenerativeLevyProcess.py:177-180, :651-654):
 *   out[b, :] = x[b, :] * table[t_b],   table = 1 / (1 + barsigma) (device float[T], host-computed).
 * t_b = t_vec[b] (device int64[B], training) when t_vec != NULL; else the batch-constant *t_dev (graph replay) or t. */
int dlpm_b200_scale_by_step(float* out, const float* x, const float* table, const int64_t* t_vec, int t, const int* t_dev,
                            int T, int64_t B, int64_t D, void* stream);

/* LIM training elements (LIM/functions/loss.py:13-31, GenerativeLevyProcess.py:680-709) for alpha < 2:
 *   x_t = x0 * exp(l_b) + e * (1 - exp(alpha l_b))^(1/alpha),   score = -e / alpha,
 *   l_b = log cos((t_b + s)/(1 + s) pi/2) - log cos(s/(1 + s) pi/2), s = 0.008   (sde.py:35-47, cosine VPSDE)
 * t: device float[B] continuous times; e: injected SaS noise (B, D) or NULL -> clamp(sqrt(A) G) drawn in-kernel with the
 * same streams as dlpm_b200_sas (isotropic: one A per sample). */
int dlpm_b200_lim_training_elements(float* x_t, float* score, const float* x0, const float* t, const float* e, int64_t B,
                                    int64_t D, float alpha, int isotropic, float clamp_eps, uint64_t seed, uint64_t offset,
                                    int64_t sample_base, void* stream);

/* Per-sample loss terms compute_loss_terms (GenerativeLevyProcess.py:19-31): lploss 2 -> sqrt(mean sq),
 * 1 -> mean smooth-L1(beta=1), -1 -> mean sq.  out[B]. pred may be bf16 (flag DLPM_STEP_EPS_BF16). */
int dlpm_b200_loss_terms(float* out, const void* pred, const float* target, int64_t B, int64_t D, float lploss,
                         int flags, void* stream);

/* GenerationManager.generate post-processing (bem/GenerationManager.py:50-63): y = clamp(x,+-c), and for
 * images (x+1)/2.  In place when out == x. */
int dlpm_b200_postprocess(float* out, const float* x, int64_t n, float clamp, int is_image, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4. Score network for the 2-D configs: MLPModel.forward (dlpm/models/Model.py:148-211,
 * DiffusionBlocks.py:125-136), the only runnable configuration (no_a, learnable time embedding,
 * LayerNorm, skip connections).  Weights: one flat fp32 device buffer in this order
 *   time_mlp.0 W[E,1] b[E] | time_mlp.2 W[E,E] b[E] | linear_in W[U,F] b[U] | group_norm_in g[U] b[U] |
 *   per block (nblocks+1, the last being outblocks_mean.0):
 *       mlp_1.1 W[U,U] b[U] | mlp_1.2 g[U] b[U] | t_proj.1 W[U,E] b[U] | mlp_2.1 W[U,U] b[U] | mlp_2.2 g[U] b[U] |
 *   outblocks_mean.1 W[F,U] b[F]
 * (row-major [out,in] as in nn.Linear).  U must be 64, E <= 64, F <= 4 in this build.
 * t: device float[B] (already scaled, GenerativeLevyProcess.py:92-96). */
int dlpm_b200_mlp_forward(float* out, const float* x, const float* t, const float* weights, int64_t B, int F,
                          int U, int E, int nblocks_total, void* stream);

/* Whole reverse chain for the 2-D configs in ONE persistent launch: x_{T-1} -> x_0 with the network,
 * the in-kernel noise and the posterior update fused (p_sample_loop_progressive,
 * GenerativeLevyProcess.py:291-330).  Sigma: compact (T,B) from dlpm_b200_sigma_scan.
 * z: NULL (in-kernel) or injected (T-1, B, F).  hist: NULL or (T, B, F) history (entry 0 = x_init).
 * mode: 0 = DLPM stochastic, 1 = DLIM eta=0. */
int dlpm_b200_mlp_sample_chain(float* x, const float* weights, const float* Sigma, const float* sched, int T,
                               int64_t B, int F, int U, int E, int nblocks_total, int mode, int flags,
                               const float* z, float* hist, uint64_t seed, uint64_t offset, int64_t sample_base,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5-K7. Image score network: UNetModel.forward (dlpm/models/unet.py:463-492) behind one handle.
 * See include/dlpm_b200_unet.h. */

#ifdef __cplusplus
}
#endif
#endif /* DLPM_B200_H_ */
