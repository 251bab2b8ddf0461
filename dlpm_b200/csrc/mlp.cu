// K4: fused score network for the 2-D configs and the persistent whole-chain sampler.
//
// Replaces MLPModel.forward (dlpm/models/Model.py:148-211; DiffusionBlockConditioned.forward,
// dlpm/models/DiffusionBlocks.py:125-136) -- 85 tiny ATen launches per forward in the reference -- and,
// for sampling, the whole p_sample_loop_progressive loop (dlpm/methods/GenerativeLevyProcess.py:291-330)
// with ONE launch: a CTA owns 64 samples for all T-1 steps; the 41 k main-path weights live in
// shared memory (166 KB, loaded once), activations never leave the SM, the Gaussian z_t is drawn
// in registers (same Philox counters as the stand-alone step kernel) and the posterior update of
// dlpm.py:250-278 is applied in place.  fp32 FMA throughout (the reference is fp32; tolerance 1e-3).
//
// Work decomposition: 256 threads = 16 sample-groups (ty) x 16 output-groups (tx); each thread owns
// a 4-sample x 4-output register tile of every 64-wide layer (classic SIMT SGEMM micro-tile:
// 8 LDS.128 per 64 FMA), LayerNorm statistics are reduced over the 16 tx lanes with shuffles.
#include "common.cuh"
#include "rng.cuh"

namespace dlpm {

// samples per CTA: 64 (256 threads) for the stand-alone forward, 80 (320 threads) for the persistent chain -- the largest
// tile whose activations fit next to the 173 KB of weights, so that 10 000 samples (config C1) are ONE wave of 125 CTAs
constexpr int U = 64;        // hidden width (nunits)
constexpr int ROW = 68;      // padded activation row stride (floats): conflict-free column reads
constexpr int MAX_F = 4;
constexpr int MAX_E = 64;

// Packed weight buffer layout (floats).  Built by dlpm_b200/score_nets.py::MLPModel.packed_weights().
struct MlpLayout {
  int F, E, NB;
  __host__ __device__ int tw1() const { return 0; }                 // [E]    time_mlp.0 weight (in=1)
  __host__ __device__ int tb1() const { return E; }                 // [E]
  __host__ __device__ int tw2T() const { return 2 * E; }            // [E][E] in-major
  __host__ __device__ int tb2() const { return 2 * E + E * E; }     // [E]
  __host__ __device__ int tproj() const { return 3 * E + E * E; }   // NB x ( wtT [E][U], bt [U] )
  __host__ __device__ int tproj_stride() const { return E * U + U; }
  __host__ __device__ int main0() const { return tproj() + NB * tproj_stride(); }
  // main block (copied to shared memory verbatim), offsets relative to main0():
  __host__ __device__ int winT() const { return 0; }                // [F][U]
  __host__ __device__ int bin() const { return F * U; }             // [U]
  __host__ __device__ int gin() const { return F * U + U; }         // [U] LayerNorm weight
  __host__ __device__ int bein() const { return F * U + 2 * U; }    // [U] LayerNorm bias
  __host__ __device__ int blk0() const { return F * U + 3 * U; }
  __host__ __device__ int blk_stride() const { return 2 * U * U + 6 * U; }  // w1T b1 g1 be1 w2T b2 g2 be2
  __host__ __device__ int wout() const { return blk0() + NB * blk_stride(); }  // [F][U] row-major (out,in)
  __host__ __device__ int bout() const { return wout() + F * U; }              // [F]
  __host__ __device__ int main_size() const { return bout() + ((F + 3) & ~3); }
  __host__ __device__ int total() const { return main0() + main_size(); }
};

__device__ __forceinline__ float silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// acc[i][j] += sum_k act[4ty+i][k] * WT[k][4tx+j]
template <int K>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ act, const float* __restrict__ WT, float (&acc)[4][4],
                                          int ty, int tx) {
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(act + (4 * ty + i) * ROW + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(WT + (k + kk) * U + 4 * tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
        acc[i][0] = fmaf(av, w.x, acc[i][0]);
        acc[i][1] = fmaf(av, w.y, acc[i][1]);
        acc[i][2] = fmaf(av, w.z, acc[i][2]);
        acc[i][3] = fmaf(av, w.w, acc[i][3]);
      }
    }
  }
}

// bias + LayerNorm([64], eps 1e-5) over the 16 tx lanes, in registers.
__device__ __forceinline__ void bias_layernorm(float (&acc)[4][4], const float* __restrict__ bias,
                                               const float* __restrict__ gamma, const float* __restrict__ beta, int tx) {
  const float4 b = *reinterpret_cast<const float4*>(bias + 4 * tx);
  const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * tx);
  const float4 be = *reinterpret_cast<const float4*>(beta + 4 * tx);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[i][0] += b.x; acc[i][1] += b.y; acc[i][2] += b.z; acc[i][3] += b.w;
    float s = (acc[i][0] + acc[i][1]) + (acc[i][2] + acc[i][3]);
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    const float mean = s * (1.0f / U);
    const float d0 = acc[i][0] - mean, d1 = acc[i][1] - mean, d2 = acc[i][2] - mean, d3 = acc[i][3] - mean;
    float v = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    const float rstd = rsqrtf(v * (1.0f / U) + 1e-5f);
    acc[i][0] = d0 * rstd * g.x + be.x; acc[i][1] = d1 * rstd * g.y + be.y;
    acc[i][2] = d2 * rstd * g.z + be.z; acc[i][3] = d3 * rstd * g.w + be.w;
  }
}

__device__ __forceinline__ void store_tile(float* __restrict__ act, const float (&acc)[4][4], int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(act + (4 * ty + i) * ROW + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

struct MlpSmem {
  float* w;      // main weights (MlpLayout::main_size floats)
  float* actA;   // [ST][ROW]
  float* actB;   // [ST][ROW]
  float* temb;   // [ST][ROW] per-sample time embedding (columns 0..E-1) -- per-sample-t mode
  float* tvec;   // [NB][U] SiLU(t_proj) for a batch-constant t
  float* xs;     // [ST][MAX_F] current x
  float* es;     // [ST][MAX_F] network output
};

template <int ST>
__device__ __forceinline__ MlpSmem carve(float* base, const MlpLayout& L, bool per_sample_t) {
  MlpSmem m;
  m.w = base;
  float* p = base + ((L.main_size() + 3) & ~3);
  m.actA = p; p += ST * ROW;
  m.actB = p; p += ST * ROW;
  m.tvec = p; p += L.NB * U;
  m.xs = p; p += ST * MAX_F;
  m.es = p; p += ST * MAX_F;
  m.temb = per_sample_t ? p : nullptr;
  return m;
}
static size_t mlp_smem_bytes(const MlpLayout& L, bool per_sample_t, int ST) {
  size_t f = ((L.main_size() + 3) & ~3) + 2 * ST * ROW + L.NB * U + 2 * ST * MAX_F + (per_sample_t ? ST * ROW : 0);
  return f * sizeof(float);
}

// time path for a batch-constant t: temb = SiLU(W2 SiLU(w1 t + b1) + b2); tvec[b] = SiLU(Wt_b temb + bt_b)
// (Model.py:190, DiffusionBlocks.py:131).  Uses actB as scratch.  All threads must call.
__device__ void time_path_uniform(const float* __restrict__ gw, const MlpLayout& L, const MlpSmem& m, float t) {
  float* e1 = m.actB;        // [E]
  float* e2 = m.actB + 64;   // [E]
  const int tid = threadIdx.x;
  if (tid < L.E) e1[tid] = silu(fmaf(__ldg(gw + L.tw1() + tid), t, __ldg(gw + L.tb1() + tid)));
  __syncthreads();
  if (tid < L.E) {
    float a = __ldg(gw + L.tb2() + tid);
    for (int k = 0; k < L.E; ++k) a = fmaf(e1[k], __ldg(gw + L.tw2T() + k * L.E + tid), a);
    e2[tid] = silu(a);
  }
  __syncthreads();
  for (int o = tid; o < L.NB * U; o += blockDim.x) {
    const int b = o / U, j = o - b * U;
    const float* wt = gw + L.tproj() + b * L.tproj_stride();
    float a = __ldg(wt + L.E * U + j);
    for (int k = 0; k < L.E; ++k) a = fmaf(e2[k], __ldg(wt + k * U + j), a);
    m.tvec[o] = silu(a);
  }
  __syncthreads();
}

// per-sample time embedding temb[s][0..E) for t[s]
template <int ST>
__device__ void time_embed_per_sample(const float* __restrict__ gw, const MlpLayout& L, const MlpSmem& m,
                                      const float* __restrict__ t, int64_t s0, int64_t B) {
  constexpr int S_TILE = ST;
  float* e1 = m.actB;  // [S][ROW] scratch
  for (int o = threadIdx.x; o < S_TILE * L.E; o += blockDim.x) {
    const int s = o / L.E, j = o - s * L.E;
    const float tv = (s0 + s < B) ? t[s0 + s] : 0.f;
    e1[s * ROW + j] = silu(fmaf(__ldg(gw + L.tw1() + j), tv, __ldg(gw + L.tb1() + j)));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < S_TILE * L.E; o += blockDim.x) {
    const int s = o / L.E, j = o - s * L.E;
    float a = __ldg(gw + L.tb2() + j);
    for (int k = 0; k < L.E; ++k) a = fmaf(e1[s * ROW + k], __ldg(gw + L.tw2T() + k * L.E + j), a);
    m.temb[s * ROW + j] = silu(a);
  }
  __syncthreads();
}

// One full forward for the CTA's 64 samples: reads m.xs, writes m.es.  All threads must call.
template <bool PER_SAMPLE_T, int ST>
__device__ void mlp_forward_tile(const float* __restrict__ gw, const MlpLayout& L, const MlpSmem& m) {
  constexpr int S_TILE = ST;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float* w = m.w;
  float acc[4][4], skip[4][4];
  // inblock: Linear(F,U) -> LayerNorm -> SiLU   (Model.py:98-102,194)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < L.F; ++k) {
    const float4 wv = *reinterpret_cast<const float4*>(w + L.winT() + k * U + 4 * tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xv = m.xs[(4 * ty + i) * MAX_F + k];
      acc[i][0] = fmaf(xv, wv.x, acc[i][0]); acc[i][1] = fmaf(xv, wv.y, acc[i][1]);
      acc[i][2] = fmaf(xv, wv.z, acc[i][2]); acc[i][3] = fmaf(xv, wv.w, acc[i][3]);
    }
  }
  bias_layernorm(acc, w + L.bin(), w + L.gin(), w + L.bein(), tx);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = silu(acc[i][j]);
  float* cur = m.actA;
  float* nxt = m.actB;
  __syncthreads();  // previous users of actA/actB are done
  store_tile(cur, acc, ty, tx);
  __syncthreads();

  for (int b = 0; b < L.NB; ++b) {
    const float* wb = w + L.blk0() + b * L.blk_stride();
    const float* w1T = wb;
    const float* b1 = wb + U * U;
    const float* g1 = b1 + U;
    const float* be1 = g1 + U;
    const float* w2T = be1 + U;
    const float* b2 = w2T + U * U;
    const float* g2 = b2 + U;
    const float* be2 = g2 + U;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { skip[i][j] = acc[i][j]; acc[i][j] = 0.f; }  // x_skip (DiffusionBlocks.py:126-127)
    gemm_tile<U>(cur, w1T, acc, ty, tx);
    bias_layernorm(acc, b1, g1, be1, tx);
    // x = act(mlp_1(x)); x += t_proj(t_emb)   (DiffusionBlocks.py:128-130)
    if (PER_SAMPLE_T) {
      float tp[4][4];
      const float* wt = gw + L.tproj() + b * L.tproj_stride();
      const float4 bt = __ldg(reinterpret_cast<const float4*>(wt + L.E * U + 4 * tx));
#pragma unroll
      for (int i = 0; i < 4; ++i) { tp[i][0] = bt.x; tp[i][1] = bt.y; tp[i][2] = bt.z; tp[i][3] = bt.w; }
      for (int k = 0; k < L.E; ++k) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + k * U + 4 * tx));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ev = m.temb[(4 * ty + i) * ROW + k];
          tp[i][0] = fmaf(ev, wv.x, tp[i][0]); tp[i][1] = fmaf(ev, wv.y, tp[i][1]);
          tp[i][2] = fmaf(ev, wv.z, tp[i][2]); tp[i][3] = fmaf(ev, wv.w, tp[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = silu(acc[i][j]) + silu(tp[i][j]);
    } else {
      const float4 tv = *reinterpret_cast<const float4*>(m.tvec + b * U + 4 * tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = silu(acc[i][0]) + tv.x; acc[i][1] = silu(acc[i][1]) + tv.y;
        acc[i][2] = silu(acc[i][2]) + tv.z; acc[i][3] = silu(acc[i][3]) + tv.w;
      }
    }
    store_tile(nxt, acc, ty, tx);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    gemm_tile<U>(nxt, w2T, acc, ty, tx);
    bias_layernorm(acc, b2, g2, be2, tx);
    // x = act(mlp_2(x) + x_skip)   (DiffusionBlocks.py:133-136)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = silu(acc[i][j] + skip[i][j]);
    store_tile(cur, acc, ty, tx);  // all reads of `cur` for this block finished before the previous barrier
    __syncthreads();
  }
  // outblocks_mean[1]: Linear(U, F)   (Model.py:199)
  if (tid < S_TILE * L.F) {
    const int s = tid / L.F, f = tid - s * L.F;
    const float* wo = w + L.wout() + f * U;
    float a = w[L.bout() + f];
#pragma unroll 4
    for (int k = 0; k < U; k += 4) {
      const float4 av = *reinterpret_cast<const float4*>(cur + s * ROW + k);
      const float4 wv = *reinterpret_cast<const float4*>(wo + k);
      a = fmaf(av.x, wv.x, a); a = fmaf(av.y, wv.y, a); a = fmaf(av.z, wv.z, a); a = fmaf(av.w, wv.w, a);
    }
    m.es[s * MAX_F + f] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ void load_main_weights(const float* __restrict__ gw, const MlpLayout& L, const MlpSmem& m) {
  const float4* src = reinterpret_cast<const float4*>(gw + L.main0());
  float4* dst = reinterpret_cast<float4*>(m.w);
  const int n4 = (L.main_size() + 3) >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
}

constexpr int kFwdTile = 64, kChainTile = 80;

__global__ void __launch_bounds__(kFwdTile * 4, 1) k_mlp_forward(float* __restrict__ out, const float* __restrict__ x,
                                                                 const float* __restrict__ t, const float* __restrict__ gw,
                                                                 int64_t B, MlpLayout L) {
  constexpr int S_TILE = kFwdTile;
  extern __shared__ __align__(16) float smem[];
  const MlpSmem m = carve<S_TILE>(smem, L, true);
  load_main_weights(gw, L, m);
  for (int64_t s0 = (int64_t)blockIdx.x * S_TILE; s0 < B; s0 += (int64_t)gridDim.x * S_TILE) {
    __syncthreads();
    for (int o = threadIdx.x; o < S_TILE * L.F; o += blockDim.x) {
      const int s = o / L.F, f = o - s * L.F;
      m.xs[s * MAX_F + f] = (s0 + s < B) ? x[(s0 + s) * L.F + f] : 0.f;
    }
    __syncthreads();
    time_embed_per_sample<S_TILE>(gw, L, m, t, s0, B);
    mlp_forward_tile<true, S_TILE>(gw, L, m);
    for (int o = threadIdx.x; o < S_TILE * L.F; o += blockDim.x) {
      const int s = o / L.F, f = o - s * L.F;
      if (s0 + s < B) out[(s0 + s) * L.F + f] = m.es[s * MAX_F + f];
    }
  }
}

// DLPM posterior pieces shared with process.cu (same arithmetic, see dlpm.py:250-278)
__device__ __forceinline__ float chain_update(float xv, float ev, float zv, float S1, float St, const float4& row,
                                              float bs_prev, int t, int mode, bool clip) {
  if (clip) {
    const float xs = fminf(fmaxf(__fdiv_rn(__fsub_rn(xv, __fmul_rn(ev, row.w)), row.y), -1.f), 1.f);
    ev = __fdiv_rn(__fsub_rn(xv, __fmul_rn(xs, row.y)), row.w);
  }
  if (mode == 1) return __fadd_rn(__fdiv_rn(__fsub_rn(xv, __fmul_rn(row.w, ev)), row.x), __fmul_rn(bs_prev, ev));
  const float Gamma = __fsub_rn(1.0f, __fdiv_rn(__fmul_rn(__fmul_rn(row.x, row.x), S1), St));
  const float mean = __fdiv_rn(__fsub_rn(xv, __fmul_rn(__fmul_rn(row.w, Gamma), ev)), row.x);
  const float sd = (t == 1) ? 0.f : __fsqrt_rn(__fmul_rn(Gamma, S1));
  return __fadd_rn(mean, __fmul_rn(sd, zv));
}

__global__ void __launch_bounds__(kChainTile * 4, 1) k_mlp_chain(float* __restrict__ x, const float* __restrict__ gw,
                                                      const float* __restrict__ Sigma, const float* __restrict__ sched,
                                                      int T, int64_t B, MlpLayout L, int mode, int flags,
                                                      const float* __restrict__ z, float* __restrict__ hist,
                                                      uint64_t seed, uint64_t offset, int64_t sample_base) {
  constexpr int S_TILE = kChainTile;
  extern __shared__ __align__(16) float smem[];
  const MlpSmem m = carve<S_TILE>(smem, L, false);
  const Philox ph(seed);
  const bool clip = flags & DLPM_STEP_CLIP_DENOISED;
  load_main_weights(gw, L, m);
  const int tid = threadIdx.x;
  const float inv_T = 1.0f / (float)T;  // GenerativeLevyProcess.py:92-96 (rescale_timesteps)
  for (int64_t s0 = (int64_t)blockIdx.x * S_TILE; s0 < B; s0 += (int64_t)gridDim.x * S_TILE) {
    __syncthreads();
    const bool owner = tid < S_TILE * L.F;
    const int s = owner ? tid / L.F : 0, f = owner ? tid - s * L.F : 0;
    const int64_t b = s0 + s;
    const bool live = owner && b < B;
    float xv = live ? x[b * L.F + f] : 0.f;
    if (owner) m.xs[s * MAX_F + f] = xv;
    if (live && hist) hist[b * L.F + f] = xv;  // history entry 0 = x_{T-1} (GenerativeLevyProcess.py:314)
    __syncthreads();
    for (int t = T - 1; t >= 1; --t) {
      time_path_uniform(gw, L, m, (float)t * inv_T);
      mlp_forward_tile<false, S_TILE>(gw, L, m);
      if (owner) {
        const float ev = m.es[s * MAX_F + f];
        if (live) {
          const float4 row = __ldg(reinterpret_cast<const float4*>(sched) + t);
          const float bs_prev = __ldg(sched + 4 * (t - 1) + 3);
          float zv = 0.f, S1 = 1.f, St = 1.f;
          if (mode == 0) {
            S1 = __ldg(Sigma + (int64_t)(t - 1) * B + b);
            St = __ldg(Sigma + (int64_t)t * B + b);
            if (z) zv = z[((int64_t)(T - 1 - t) * B + b) * L.F + f];
            else {
              const float4 q = normal4(philox_at(ph, STREAM_Z, offset + (uint64_t)t, (uint64_t)(b + sample_base), (uint32_t)(f >> 2)));
              const int k = f & 3;
              zv = k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
            }
          }
          xv = chain_update(xv, ev, zv, S1, St, row, bs_prev, t, mode, clip);
          if (hist) hist[((int64_t)(T - t) * B + b) * L.F + f] = xv;
        }
        m.xs[s * MAX_F + f] = xv;
      }
      __syncthreads();
    }
    if (live) x[b * L.F + f] = xv;
  }
}

}  // namespace dlpm

using namespace dlpm;

static int check_mlp_dims(int F, int U_, int E, int NB) {
  if (U_ != U) { set_error("mlp: nunits must be 64 in this build (got %d)", U_); return DLPM_ERR_UNSUPPORTED; }
  if (F < 1 || F > MAX_F || E < 4 || E > MAX_E || (E % 4) || NB < 1 || NB > 16) {
    set_error("mlp: unsupported dims F=%d E=%d blocks=%d", F, E, NB);
    return DLPM_ERR_UNSUPPORTED;
  }
  return DLPM_OK;
}

int dlpm_b200_mlp_forward(float* out, const float* x, const float* t, const float* weights, int64_t B, int F, int U_,
                          int E, int nblocks_total, void* stream) {
  DLPM_REQUIRE(out && x && t && weights, "mlp_forward: NULL tensor");
  DLPM_REQUIRE(B >= 0, "mlp_forward: bad batch");
  if (int rc = check_mlp_dims(F, U_, E, nblocks_total)) return rc;
  if (B == 0) return DLPM_OK;
  MlpLayout L{F, E, nblocks_total};
  constexpr int S_TILE = kFwdTile;
  const size_t smem = mlp_smem_bytes(L, true, S_TILE);
  DLPM_REQUIRE(smem <= 227 * 1024, "mlp_forward: network does not fit in shared memory");
  cudaError_t e = cudaFuncSetAttribute(k_mlp_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "mlp_forward smem attr");
  int64_t tiles = (B + S_TILE - 1) / S_TILE;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  k_mlp_forward<<<grid, S_TILE * 4, smem, (cudaStream_t)stream>>>(out, x, t, weights, B, L);
  DLPM_CHECK_LAUNCH("mlp_forward");
  return DLPM_OK;
}

int dlpm_b200_mlp_sample_chain(float* x, const float* weights, const float* Sigma, const float* sched, int T, int64_t B,
                               int F, int U_, int E, int nblocks_total, int mode, int flags, const float* z, float* hist,
                               uint64_t seed, uint64_t offset, int64_t sample_base, void* stream) {
  DLPM_REQUIRE(x && weights && sched, "mlp_sample_chain: NULL tensor");
  DLPM_REQUIRE(mode == 1 || Sigma, "mlp_sample_chain: Sigma required for the stochastic chain");
  DLPM_REQUIRE(mode == 0 || mode == 1, "mlp_sample_chain: mode must be 0 (DLPM) or 1 (DLIM)");
  DLPM_REQUIRE(T >= 2 && B >= 0, "mlp_sample_chain: bad sizes");
  if (int rc = check_mlp_dims(F, U_, E, nblocks_total)) return rc;
  if (B == 0) return DLPM_OK;
  MlpLayout L{F, E, nblocks_total};
  constexpr int S_TILE = kChainTile;
  const size_t smem = mlp_smem_bytes(L, false, S_TILE);
  DLPM_REQUIRE(smem <= 227 * 1024, "mlp_sample_chain: network does not fit in shared memory");
  cudaError_t e = cudaFuncSetAttribute(k_mlp_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "mlp_sample_chain smem attr");
  int64_t tiles = (B + S_TILE - 1) / S_TILE;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  k_mlp_chain<<<grid, S_TILE * 4, smem, (cudaStream_t)stream>>>(x, weights, Sigma, sched, T, B, L, mode, flags, z, hist, seed, offset,
                                                        sample_base);
  DLPM_CHECK_LAUNCH("mlp_sample_chain");
  return DLPM_OK;
}
