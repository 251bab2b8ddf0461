// K1 (alpha-stable noise), K2 (Sigma scan), K3 (fused reverse steps) and the training-forward
// helpers.  All of these are HBM-bound streaming kernels: 128-bit accesses, grid = multiples of the
// SM count, noise generated in registers so the x_t state makes exactly one read + one write per step.
//
// Reference behaviour being replaced (file:line relative to the reference tree):
//   bem/datasets/Distributions.py:33-73        gen_skewed_levy / gen_sas (host scipy + H2D + 3 eager kernels)
//   dlpm/methods/dlpm.py:226-239               sample_A / compute_Sigmas (T host draws, 4T launches, (T,B,C,H,W) tables)
//   dlpm/methods/dlpm.py:250-297               Gamma_t, posterior mean/variance, DLIM
//   dlpm/methods/GenerativeLevyProcess.py:186-239   clip_denoised path, p_sample
//   dlpm/methods/LIM/functions/sampler.py:81-181    LIM ODE / SDE updates
// Arithmetic uses explicit round-to-nearest intrinsics in the reference's evaluation order (no FMA
// contraction), so with injected noise the results are bit-identical to the reference's fp32 CPU path.
#include <cstdarg>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "rng.cuh"

namespace dlpm {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;  // -1: not initialised -> DLPM_B200_PDL environment variable (default off: measured no gain inside the CUDA graph)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("DLPM_B200_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pdl != 0;
}
void pdl_set_enabled(bool on) { g_pdl = on ? 1 : 0; }

static StableParams make_params(float alpha) {
  StableParams p;
  p.gaussian = (alpha == 2.0f);
  const double ap = (double)alpha / 2.0;
  p.ap = (float)ap;
  p.inv_ap = (float)(1.0 / ap);
  p.r = (float)((1.0 - ap) / ap);
  p.one_m_ap = (float)(1.0 - ap);
  return p;
}

constexpr int kChunkQuads = 2048;  // quads per CTA iteration of the streaming kernels (8 per thread)

__device__ __forceinline__ float clamp_A(float a, float clamp_a) { return clamp_a >= 0.f ? fminf(fmaxf(a, 0.f), clamp_a) : a; }
__device__ __forceinline__ float clamp_sym(float v, float c) { return c >= 0.f ? fminf(fmaxf(v, -c), c) : v; }
__device__ __forceinline__ float sel4(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// per-sample subordinator (one draw per sample; position 0 of the sample's A stream)
template <class P>
__device__ __forceinline__ float sample_A(const P& ph, const StableParams& sp, uint32_t stream, uint64_t offset,
                                          uint64_t sample) {
  if (sp.gaussian) return 2.0f;  // alpha == 2: A == 2, no variates consumed (Distributions.py:40-42)
  const uint4 r = philox_at(ph, stream, offset, sample, 0u);
  return stable_A(sp, r.x, r.y);
}
// four per-element subordinators for quad `pos` of a sample (two Philox blocks: positions 2pos, 2pos+1)
template <class P>
__device__ __forceinline__ float4 element_A4(const P& ph, const StableParams& sp, uint32_t stream, uint64_t offset,
                                             uint64_t sample, uint32_t pos) {
  if (sp.gaussian) return make_float4(2.0f, 2.0f, 2.0f, 2.0f);
  const uint4 r0 = philox_at(ph, stream, offset, sample, 2u * pos + 1u);  // +1: position 0 is the per-sample draw
  const uint4 r1 = philox_at(ph, stream, offset, sample, 2u * pos + 2u);
  return make_float4(stable_A(sp, r0.x, r0.y), stable_A(sp, r0.z, r0.w), stable_A(sp, r1.x, r1.y),
                     stable_A(sp, r1.z, r1.w));
}
template <class P>
__device__ __forceinline__ float4 normal_quad(const P& ph, uint32_t stream, uint64_t offset, uint64_t sample,
                                              uint32_t pos) {
  return normal4(philox_at(ph, stream, offset, sample, pos));
}

// ------------------------------------------------------------------------------------------------
// Skeleton of the vectorised streaming kernels (K1, K3, LIM).  The flat index space of float4 "quads" is cut into
// 32-quad granules (512 B) and every CTA owns ONE contiguous share of them, equal to within a granule, with
// grid = SMs x resident CTAs: a single full wave, no tail.  A thread walks its share in strides of 256 quads and
// carries (sample, position) incrementally -- no per-quad division.  Per-sample scalars (sqrt(A), step coefficients)
// of a sub-chunk of <= kChunkQuads quads are computed once into shared memory.  32-bit indices: nq < 2^31.
// ------------------------------------------------------------------------------------------------
struct QuadSpan {
  uint32_t nq;        // total quads = n_outer * inner / 4
  uint32_t qpr;       // quads per sample row = inner / 4
  uint32_t step_o;    // 256 quads expressed as (samples, quads): 256 = step_o * qpr + step_pos
  uint32_t step_pos;
  uint32_t sub;       // sub-chunk length in quads: touches <= kChunkQuads samples (size of the per-sample smem tables)
  FastDiv fd;         // division by qpr
};
static QuadSpan make_span(int64_t n_outer, int64_t inner) {
  QuadSpan s;
  s.qpr = (uint32_t)(inner / 4);
  s.nq = (uint32_t)(n_outer * (inner / 4));
  s.step_o = 256u / s.qpr;
  s.step_pos = 256u % s.qpr;
  s.fd = FastDiv(s.qpr);
  const uint64_t sub = (uint64_t)(kChunkQuads - 2) * s.qpr;
  s.sub = sub < (uint64_t)kChunkQuads ? (uint32_t)kChunkQuads : (sub > (1u << 30) ? (1u << 30) : (uint32_t)(sub & ~31ull));
  return s;
}
__device__ __forceinline__ void span_advance(const QuadSpan& s, uint32_t& o, uint32_t& pos) {
  pos += s.step_pos;
  o += s.step_o;
  if (pos >= s.qpr) { pos -= s.qpr; ++o; }
}
__device__ __forceinline__ void cta_share(uint32_t nq, uint32_t& q_lo, uint32_t& q_hi) {
  const uint32_t ngran = (nq + 31u) >> 5;
  const uint32_t g0 = (uint32_t)(((uint64_t)ngran * blockIdx.x) / gridDim.x);
  const uint32_t g1 = (uint32_t)(((uint64_t)ngran * (blockIdx.x + 1)) / gridDim.x);
  q_lo = g0 << 5;
  q_hi = (g1 << 5) < nq ? (g1 << 5) : nq;
}

// ------------------------------------------------------------------------------------------------
// K1a  stable_A
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) k_stable_A(float* __restrict__ out, int64_t n_outer, int64_t inner, int mode,
                                                  StableParams sp, float clamp_a, uint64_t seed, uint64_t offset,
                                                  int64_t sample_base, FastDiv fd, int chunk) {
  const Philox ph(seed);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (mode == DLPM_A_COMPACT) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_outer; o += stride)
      out[o] = clamp_A(sample_A(ph, sp, STREAM_A, offset, (uint64_t)(o + sample_base)), clamp_a);
    return;
  }
  if (VEC) {
    // block-contiguous chunks: the per-sample draws of a chunk are computed once into shared memory
    const int64_t qpr = inner >> 2, nq = n_outer * qpr;  // quads per row
    const bool fast = nq < (1ll << 31);
    __shared__ float s_a[kChunkQuads];
    for (int64_t q0 = (int64_t)blockIdx.x * chunk; q0 < nq; q0 += (int64_t)gridDim.x * chunk) {
      const int64_t q1 = q0 + chunk < nq ? q0 + chunk : nq;
      const int64_t o_first = fast ? (int64_t)fd.div((uint32_t)q0) : q0 / qpr;
      if (mode == DLPM_A_ISOTROPIC) {
        const int64_t o_last = fast ? (int64_t)fd.div((uint32_t)(q1 - 1)) : (q1 - 1) / qpr;
        __syncthreads();
        for (int64_t sI = threadIdx.x; sI <= o_last - o_first; sI += blockDim.x)
          s_a[sI] = clamp_A(sample_A(ph, sp, STREAM_A, offset, (uint64_t)(o_first + sI + sample_base)), clamp_a);
        __syncthreads();
      }
      for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        uint32_t o32, pos;
        if (fast) fd.divmod((uint32_t)q, o32, pos);
        const int64_t o = fast ? (int64_t)o32 : q / qpr;
        if (!fast) pos = (uint32_t)(q - o * qpr);
        float4 v;
        if (mode == DLPM_A_ISOTROPIC) {
          const float a_iso = s_a[o - o_first];
          v = make_float4(a_iso, a_iso, a_iso, a_iso);
        } else {
          v = element_A4(ph, sp, STREAM_A, offset, (uint64_t)(o + sample_base), pos);
          v.x = clamp_A(v.x, clamp_a); v.y = clamp_A(v.y, clamp_a); v.z = clamp_A(v.z, clamp_a); v.w = clamp_A(v.w, clamp_a);
        }
        st_stream(reinterpret_cast<float4*>(out) + q, v);
      }
    }
  } else {
    const int64_t n = n_outer * inner;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
      const int64_t o = e / inner;
      const int64_t i = e - o * inner;
      float a;
      if (mode == DLPM_A_ISOTROPIC) a = sample_A(ph, sp, STREAM_A, offset, (uint64_t)(o + sample_base));
      else a = sel4(element_A4(ph, sp, STREAM_A, offset, (uint64_t)(o + sample_base), (uint32_t)(i >> 2)), (int)(i & 3));
      out[e] = clamp_A(a, clamp_a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1b  sas  (and plain normal fill)
// ------------------------------------------------------------------------------------------------
// A_MODE: 0 = none (plain normal), 1 = in-kernel isotropic, 2 = in-kernel per element, 3 = A_in compact, 4 = A_in full
template <bool VEC, int A_MODE>
__global__ void __launch_bounds__(256) k_sas(float* __restrict__ out, const float* __restrict__ A_in, int64_t n_outer,
                                             int64_t inner, StableParams sp, float clamp_eps, float scale,
                                             uint32_t g_stream, uint64_t seed, uint64_t offset, int64_t sample_base,
                                             FastDiv fd, int chunk, int sext) {
  // sext: the row layout of the "sextet" scheme (rng.cuh; rows that are multiples of 384 elements, A_MODE 0 / 1 / 3) -- these generic
  // kernels (unaligned tensors, >= 2^31 quads) then evaluate it quad by quad so that every path gives the same field
  const Philox ph(seed);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int64_t qpr = inner >> 2, nq = n_outer * qpr;
    const bool fast = nq < (1ll << 31);
    __shared__ float s_sa[kChunkQuads];  // sqrt(A) of the samples touched by this chunk
    for (int64_t q0 = (int64_t)blockIdx.x * chunk; q0 < nq; q0 += (int64_t)gridDim.x * chunk) {
      const int64_t q1 = q0 + chunk < nq ? q0 + chunk : nq;
      const int64_t o_first = fast ? (int64_t)fd.div((uint32_t)q0) : q0 / qpr;
      if (A_MODE == 1 || A_MODE == 3) {
        const int64_t o_last = fast ? (int64_t)fd.div((uint32_t)(q1 - 1)) : (q1 - 1) / qpr;
        __syncthreads();
        for (int64_t sI = threadIdx.x; sI <= o_last - o_first; sI += blockDim.x)
          s_sa[sI] = __fsqrt_rn(A_MODE == 1 ? sample_A(ph, sp, STREAM_EPS_A, offset, (uint64_t)(o_first + sI + sample_base))
                                            : __ldg(A_in + o_first + sI));
        __syncthreads();
      }
      for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        uint32_t o32, pos;
        if (fast) fd.divmod((uint32_t)q, o32, pos);
        const int64_t o = fast ? (int64_t)o32 : q / qpr;
        if (!fast) pos = (uint32_t)(q - o * qpr);
        const uint64_t sample = (uint64_t)(o + sample_base);
        float4 g;
        if (sext) {
          g = normal_sextet_quad(ph, g_stream, offset, sample, pos, (A_MODE == 1 || A_MODE == 3) ? s_sa[o - o_first] : 1.0f);
        } else {
          g = normal_quad(ph, g_stream, offset, sample, pos);
          if (A_MODE == 1 || A_MODE == 3) {
            const float sa_iso = s_sa[o - o_first];
            g.x *= sa_iso; g.y *= sa_iso; g.z *= sa_iso; g.w *= sa_iso;
          }
        }
        if (A_MODE == 2 || A_MODE == 4) {
          const float4 a = (A_MODE == 2) ? element_A4(ph, sp, STREAM_EPS_A, offset, sample, pos)
                                         : ld_stream(reinterpret_cast<const float4*>(A_in) + q);
          g.x *= __fsqrt_rn(a.x); g.y *= __fsqrt_rn(a.y); g.z *= __fsqrt_rn(a.z); g.w *= __fsqrt_rn(a.w);
        }
        if (A_MODE != 0) {
          g.x = scale * clamp_sym(g.x, clamp_eps); g.y = scale * clamp_sym(g.y, clamp_eps);
          g.z = scale * clamp_sym(g.z, clamp_eps); g.w = scale * clamp_sym(g.w, clamp_eps);
        }
        st_stream(reinterpret_cast<float4*>(out) + q, g);
      }
    }
  } else {
    const int64_t n = n_outer * inner;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
      const int64_t o = e / inner;
      const int64_t i = e - o * inner;
      const uint64_t sample = (uint64_t)(o + sample_base);
      const uint32_t pos = (uint32_t)(i >> 2);
      float a = 1.0f;
      if (A_MODE == 1) a = sample_A(ph, sp, STREAM_EPS_A, offset, sample);
      else if (A_MODE == 3) a = A_in[o];
      else if (A_MODE == 2) a = sel4(element_A4(ph, sp, STREAM_EPS_A, offset, sample, pos), (int)(i & 3));
      else if (A_MODE == 4) a = A_in[e];
      float g;
      if (sext) g = sel4(normal_sextet_quad(ph, g_stream, offset, sample, pos, A_MODE != 0 ? __fsqrt_rn(a) : 1.0f), (int)(i & 3));
      else g = sel4(normal_quad(ph, g_stream, offset, sample, pos), (int)(i & 3)) * (A_MODE != 0 ? __fsqrt_rn(a) : 1.0f);
      if (A_MODE != 0) g = scale * clamp_sym(g, clamp_eps);
      out[e] = g;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// K1 vectorised production kernels (QuadSpan skeleton above)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stable_A_vec(float* __restrict__ out, const QuadSpan span, int mode, StableParams sp,
                                                      float clamp_a, const __grid_constant__ PhiloxKeys keys, uint64_t offset,
                                                      int64_t sample_base) {
  const PhiloxRef ph(keys);
  __shared__ float s_a[kChunkQuads];
  uint32_t q_lo, q_hi;
  cta_share(span.nq, q_lo, q_hi);
  for (uint32_t q0 = q_lo; q0 < q_hi; q0 += span.sub) {
    const uint32_t q1 = q0 + span.sub < q_hi ? q0 + span.sub : q_hi;
    const uint32_t o_first = span.fd.div(q0);
    if (mode == DLPM_A_ISOTROPIC) {
      const uint32_t o_last = span.fd.div(q1 - 1);
      __syncthreads();
      for (uint32_t sI = threadIdx.x; sI <= o_last - o_first; sI += 256)
        s_a[sI] = clamp_A(sample_A(ph, sp, STREAM_A, offset, (uint64_t)((int64_t)(o_first + sI) + sample_base)), clamp_a);
      __syncthreads();
    }
    uint32_t q = q0 + threadIdx.x, o, pos;
    span.fd.divmod(q, o, pos);
    for (; q < q1; q += 256) {
      float4 v;
      if (mode == DLPM_A_ISOTROPIC) {
        const float a_iso = s_a[o - o_first];
        v = make_float4(a_iso, a_iso, a_iso, a_iso);
      } else {
        v = element_A4(ph, sp, STREAM_A, offset, (uint64_t)((int64_t)o + sample_base), pos);
        v.x = clamp_A(v.x, clamp_a); v.y = clamp_A(v.y, clamp_a); v.z = clamp_A(v.z, clamp_a); v.w = clamp_A(v.w, clamp_a);
      }
      st_stream(reinterpret_cast<float4*>(out) + q, v);
      span_advance(span, o, pos);
    }
  }
}

// A_MODE as in k_sas
template <int A_MODE, bool SCALED>
__global__ void __launch_bounds__(256) k_sas_vec(float* __restrict__ out, const float* __restrict__ A_in, const QuadSpan span,
                                                 StableParams sp, float clamp_eps, float scale, uint32_t g_stream,
                                                 const __grid_constant__ PhiloxKeys keys, uint64_t offset, int64_t sample_base) {
  const PhiloxRef ph(keys);
  __shared__ float s_sa[(A_MODE == 1 || A_MODE == 3) ? kChunkQuads : 1];  // sqrt(A) of the samples touched by a sub-chunk
  uint32_t q_lo, q_hi;
  cta_share(span.nq, q_lo, q_hi);
  for (uint32_t q0 = q_lo; q0 < q_hi; q0 += span.sub) {
    const uint32_t q1 = q0 + span.sub < q_hi ? q0 + span.sub : q_hi;
    const uint32_t o_first = span.fd.div(q0);
    if (A_MODE == 1 || A_MODE == 3) {
      const uint32_t o_last = span.fd.div(q1 - 1);
      __syncthreads();
      for (uint32_t sI = threadIdx.x; sI <= o_last - o_first; sI += 256)
        s_sa[sI] = __fsqrt_rn(A_MODE == 1 ? sample_A(ph, sp, STREAM_EPS_A, offset, (uint64_t)((int64_t)(o_first + sI) + sample_base))
                                          : __ldg(A_in + o_first + sI));
      __syncthreads();
    }
    uint32_t q = q0 + threadIdx.x, o, pos;
    span.fd.divmod(q, o, pos);
    float4* optr = reinterpret_cast<float4*>(out) + q;  // running pointer: no per-iteration IMAD.WIDE on the fmaheavy pipe
    for (; q < q1; q += 256, optr += 256) {
      const uint64_t sample = (uint64_t)((int64_t)o + sample_base);
      float4 g = normal_quad(ph, g_stream, offset, sample, pos);
      bool may_clamp = A_MODE != 0;
      if (A_MODE == 1 || A_MODE == 3) {
        const float sa_iso = s_sa[o - o_first];
        g.x *= sa_iso; g.y *= sa_iso; g.z *= sa_iso; g.w *= sa_iso;
        // |G| <= sqrt(-2 ln 2^-33) = 6.77 on this lattice: the clamp can only bite when 6.77 sqrt(A) exceeds it
        may_clamp = sa_iso * 6.77f > clamp_eps;
      } else if (A_MODE == 2 || A_MODE == 4) {
        const float4 a = (A_MODE == 2) ? element_A4(ph, sp, STREAM_EPS_A, offset, sample, pos)
                                       : ld_stream(reinterpret_cast<const float4*>(A_in) + q);
        g.x *= __fsqrt_rn(a.x); g.y *= __fsqrt_rn(a.y); g.z *= __fsqrt_rn(a.z); g.w *= __fsqrt_rn(a.w);
      }
      if (may_clamp) {
        g.x = clamp_sym(g.x, clamp_eps); g.y = clamp_sym(g.y, clamp_eps);
        g.z = clamp_sym(g.z, clamp_eps); g.w = clamp_sym(g.w, clamp_eps);
      }
      if (SCALED && A_MODE != 0) { g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale; }
      st_stream(optr, g);
      span_advance(span, o, pos);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1 fast fill under the "sextet" scheme (rng.cuh): rows of gps granules of 96 quads; a WARP owns one granule per iteration
// (lane l draws two Philox blocks = twelve normals and writes quads l, 32 + l, 64 + l of the granule: three coalesced 512-byte
// segments); every CTA owns one contiguous share of the granules (single wave), the per-sample sqrt(A) of a sub-chunk lives in
// shared memory as in k_sas_vec.  A_MODE 0 (plain normal), 1 (in-kernel isotropic A), 3 (A_in compact).
// ------------------------------------------------------------------------------------------------
struct GranSpan {
  uint32_t ng, gps;        // granules in total / per sample row
  uint32_t step_o, step_g; // 8 granules (one CTA iteration) expressed as (samples, granules)
  uint32_t sub;            // sub-chunk length in granules: touches <= kChunkQuads samples
  uint32_t qpr;            // quads per row
  FastDiv fd;              // division by gps
};
static GranSpan make_gran_span(int64_t n_outer, int64_t inner) {
  GranSpan s;
  s.qpr = (uint32_t)(inner / 4);
  s.gps = s.qpr / 96u;
  s.ng = (uint32_t)(n_outer * s.gps);
  s.step_o = 8u / s.gps;
  s.step_g = 8u % s.gps;
  s.fd = FastDiv(s.gps);
  const uint64_t sub = (uint64_t)(kChunkQuads - 2) * s.gps;
  s.sub = sub > (1u << 28) ? (1u << 28) : (uint32_t)sub;
  return s;
}

template <int A_MODE, bool SCALED>
__global__ void __launch_bounds__(256) k_fill6(float* __restrict__ out, const float* __restrict__ A_in, const GranSpan span, StableParams sp,
                                               float clamp_eps, float scale, uint32_t g_stream, const __grid_constant__ PhiloxKeys keys,
                                               uint64_t offset, int64_t sample_base) {
  const PhiloxRef ph(keys);
  __shared__ float s_sa[A_MODE != 0 ? kChunkQuads : 1];
  const uint32_t g_lo = (uint32_t)(((uint64_t)span.ng * blockIdx.x) / gridDim.x), g_hi = (uint32_t)(((uint64_t)span.ng * (blockIdx.x + 1)) / gridDim.x);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t q0 = g_lo; q0 < g_hi; q0 += span.sub) {
    const uint32_t q1 = q0 + span.sub < g_hi ? q0 + span.sub : g_hi;
    const uint32_t o_first = span.fd.div(q0);
    if (A_MODE != 0) {
      const uint32_t o_last = span.fd.div(q1 - 1);
      __syncthreads();
      for (uint32_t sI = threadIdx.x; sI <= o_last - o_first; sI += 256)
        s_sa[sI] = __fsqrt_rn(A_MODE == 1 ? sample_A(ph, sp, STREAM_EPS_A, offset, (uint64_t)((int64_t)(o_first + sI) + sample_base))
                                          : __ldg(A_in + o_first + sI));
      __syncthreads();
    }
    uint32_t G = q0 + warp, o, gi;
    span.fd.divmod(G, o, gi);
    float4* optr = reinterpret_cast<float4*>(out) + ((uint64_t)o * span.qpr + gi * 96u + lane);
    const uint64_t step_ptr = (uint64_t)span.step_o * span.qpr + span.step_g * 96u;  // 8 granules further on (before the row wrap)
    for (; G < q1; G += 8) {
      const uint64_t sample = (uint64_t)((int64_t)o + sample_base);
      const uint32_t g = gi * 32u + lane;
      float sa = 1.0f;
      bool may_clamp = false;
      if (A_MODE != 0) {
        sa = s_sa[o - o_first];
        may_clamp = clamp_eps >= 0.f && sa * kSextetMaxAbs > clamp_eps;  // |G| <= 6.23 on this lattice
        if (SCALED && !may_clamp) sa *= scale;
      }
      float z[12];
      {
        float h[6];
        normal6(philox_at(ph, g_stream, offset, sample, 2u * g), sa, h);
#pragma unroll
        for (int i = 0; i < 6; ++i) z[i] = h[i];
        normal6(philox_at(ph, g_stream, offset, sample, 2u * g + 1u), sa, h);
#pragma unroll
        for (int i = 0; i < 6; ++i) z[6 + i] = h[i];
      }
      if (may_clamp) {  // (rare: only samples whose sqrt(A) is within a factor 6.23 of the clamp)
#pragma unroll
        for (int i = 0; i < 12; ++i) { z[i] = clamp_sym(z[i], clamp_eps); if (SCALED) z[i] *= scale; }
      }
      st_stream(optr, make_float4(z[0], z[1], z[2], z[3]));
      st_stream(optr + 32, make_float4(z[4], z[5], z[6], z[7]));
      st_stream(optr + 64, make_float4(z[8], z[9], z[10], z[11]));
      // advance by 8 granules: (o, gi) += (step_o, step_g) with the row wrap
      gi += span.step_g;
      o += span.step_o;
      optr += step_ptr;
      if (gi >= span.gps) { gi -= span.gps; ++o; }  // (the flat quad index is continuous across rows: no pointer correction)
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2  Sigma scan: one thread per chain, T sequential steps, coalesced (T, n) stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sigma_scan(float* __restrict__ Sigma, const float* __restrict__ A_in,
                                                    float* __restrict__ A_out, const float* __restrict__ sched, int T,
                                                    int64_t n, int64_t inner, int per_element, StableParams sp,
                                                    float clamp_a, uint64_t seed, uint64_t offset, int64_t sample_base) {
  const Philox ph(seed);
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n) return;
  uint64_t sample;
  uint32_t pos = 0;
  int lane = 0;
  if (per_element) {
    const int64_t o = c / inner, i = c - o * inner;
    sample = (uint64_t)(o + sample_base);
    pos = (uint32_t)(i >> 2);
    lane = (int)(i & 3);
  } else {
    sample = (uint64_t)(c + sample_base);
  }
  float S = 0.f;
  for (int t = 0; t < T; ++t) {
    const float4 row = __ldg(reinterpret_cast<const float4*>(sched) + t);  // (g, bg, s, bs)
    float a;
    if (A_in) a = A_in[(int64_t)t * n + c];
    else if (per_element) a = clamp_A(sel4(element_A4(ph, sp, STREAM_A, offset + (uint64_t)t, sample, pos), lane), clamp_a);
    else a = clamp_A(sample_A(ph, sp, STREAM_A, offset + (uint64_t)t, sample), clamp_a);
    // dlpm.py:234,238  s[t]**2 * A_t + g[t]**2 * Sigmas[-1]
    const float sa = __fmul_rn(__fmul_rn(row.z, row.z), a);
    S = (t == 0) ? sa : __fadd_rn(sa, __fmul_rn(__fmul_rn(row.x, row.x), S));
    Sigma[(int64_t)t * n + c] = S;
    if (A_out) A_out[(int64_t)t * n + c] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// K3  fused reverse steps
// ------------------------------------------------------------------------------------------------
struct StepCoef {  // per-chain scalars of one DLPM step
  float c1;        // bs_t * Gamma_t
  float g;         // gamma_t
  float sd;        // 1[t != 1] * sqrt(Gamma_t * Sigma_{t-1})
};
__device__ __forceinline__ StepCoef dlpm_coef(float S1, float St, const float4& row, int t) {
  // dlpm.py:250-254,272-278
  const float Gamma = __fsub_rn(1.0f, __fdiv_rn(__fmul_rn(__fmul_rn(row.x, row.x), S1), St));
  StepCoef c;
  c.c1 = __fmul_rn(row.w, Gamma);
  c.g = row.x;
  c.sd = (t == 1) ? 0.f : __fsqrt_rn(__fmul_rn(Gamma, S1));
  return c;
}
__device__ __forceinline__ float clip_eps1(float x, float e, const float4& row) {
  // GenerativeLevyProcess.py:186-207 with dlpm.py:191-202
  const float xs = fminf(fmaxf(__fdiv_rn(__fsub_rn(x, __fmul_rn(e, row.w)), row.y), -1.f), 1.f);
  return __fdiv_rn(__fsub_rn(x, __fmul_rn(xs, row.y)), row.w);
}
__device__ __forceinline__ float dlpm_update1(float x, float e, float z, const StepCoef& c) {
  const float mean = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.c1, e)), c.g);
  return __fadd_rn(mean, __fmul_rn(c.sd, z));
}
// Production variant (in-kernel noise, nothing to be bit-compared against): reciprocal + FMAs instead of the IEEE
// divide; differs from dlpm_update1 by <= 2 ulp.
__device__ __forceinline__ float dlpm_update1_fast(float x, float e, float z, const StepCoef& c, float inv_g) {
  return fmaf(c.sd, z, fmaf(-c.c1, e, x) * inv_g);
}

// GenerationManager.generate post-processing (bem/GenerationManager.py:50-63) fused into the LAST step's store: when the
// step being executed is `at`, the new x_0 is also written to `out` clamped to +-clamp, mapped to (x+1)/2 for images, and
// optionally quantised like torchvision.utils.save_image (x*255 + 0.5, clamp, truncate) into uint8 NHWC -- no extra pass
// over x_0 and nothing to wait for before the device -> host copy.
struct StepPost {
  void* out;   // nullptr = off
  float clamp;
  int mode;    // DLPM_POST_F32 (clamp), DLPM_POST_F32_IMAGE (clamp, (x+1)/2), DLPM_POST_U8_NHWC
  int at;      // step index whose result is final (DLPM / DLIM: 1; LIM: n_steps - 1)
  int C, HW;   // sample layout (C, H*W) for the NHWC mode; D = C * HW
};
__device__ __forceinline__ float post_val(const StepPost& p, float v) {
  v = fminf(fmaxf(v, -p.clamp), p.clamp);
  return p.mode == DLPM_POST_F32 ? v : __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
}
__device__ __forceinline__ void post_store1(const StepPost& p, int64_t b, int64_t i, float v) {
  v = post_val(p, v);
  if (p.mode != DLPM_POST_U8_NHWC) { reinterpret_cast<float*>(p.out)[b * ((int64_t)p.C * p.HW) + i] = v; return; }
  const int c = (int)(i / p.HW), pix = (int)(i - (int64_t)c * p.HW);
  reinterpret_cast<uint8_t*>(p.out)[(b * p.HW + pix) * p.C + c] = (uint8_t)fminf(fmaxf(fmaf(v, 255.0f, 0.5f), 0.0f), 255.0f);
}
__device__ __forceinline__ void post_store4(const StepPost& p, int64_t q, int64_t b, uint32_t pos, const float4& v) {
  if (p.mode != DLPM_POST_U8_NHWC) {
    st_stream(reinterpret_cast<float4*>(p.out) + q, make_float4(post_val(p, v.x), post_val(p, v.y), post_val(p, v.z), post_val(p, v.w)));
    return;
  }
  const int64_t i = (int64_t)pos * 4;
  post_store1(p, b, i, v.x); post_store1(p, b, i + 1, v.y); post_store1(p, b, i + 2, v.z); post_store1(p, b, i + 3, v.w);
}

template <bool VEC, bool EPS_BF16>
__device__ __forceinline__ float4 load_eps4(const void* eps, int64_t q) {
  if (EPS_BF16) return bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(eps) + q));
  return ld_stream(reinterpret_cast<const float4*>(eps) + q);
}
template <bool EPS_BF16>
__device__ __forceinline__ float load_eps1(const void* eps, int64_t e) {
  if (EPS_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(eps)[e]);
  return reinterpret_cast<const float*>(eps)[e];
}

// MODE 0: DLPM stochastic, MODE 1: DLIM eta=0
template <bool VEC, bool EPS_BF16, int MODE>
__global__ void __launch_bounds__(256) k_reverse_step(float* __restrict__ x, const void* __restrict__ eps,
                                                      const float* __restrict__ Sigma, const float* __restrict__ sched,
                                                      int t_imm, const int* __restrict__ t_dev, int T, int64_t B,
                                                      int64_t D, int flags, const float* __restrict__ z,
                                                      uint64_t seed, uint64_t offset, int64_t sample_base,
                                                      float* __restrict__ hist, FastDiv fd, const StepPost post) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = t_dev ? *t_dev : t_imm;
  if (t < 1 || t >= T) return;
  const bool do_post = post.out != nullptr && t == post.at;
  const Philox ph(seed);
  const float4 row = __ldg(reinterpret_cast<const float4*>(sched) + t);
  const float bs_prev = __ldg(sched + 4 * (t - 1) + 3);
  const bool clip = flags & DLPM_STEP_CLIP_DENOISED;
  const bool sig_full = flags & DLPM_STEP_SIGMA_FULL;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint64_t off_t = offset + (uint64_t)t;
  if (VEC) {
    const int64_t qpr = D >> 2, nq = B * qpr;
    const bool fast = nq < (1ll << 31);
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += stride) {
      uint32_t b32, pos;
      if (fast) fd.divmod((uint32_t)q, b32, pos);
      const int64_t b = fast ? (int64_t)b32 : q / qpr;
      if (!fast) pos = (uint32_t)(q - b * qpr);
      float4 xv = ld_rw(reinterpret_cast<const float4*>(x) + q);
      float4 ev = load_eps4<VEC, EPS_BF16>(eps, q);
      if (clip) {
        ev.x = clip_eps1(xv.x, ev.x, row); ev.y = clip_eps1(xv.y, ev.y, row);
        ev.z = clip_eps1(xv.z, ev.z, row); ev.w = clip_eps1(xv.w, ev.w, row);
      }
      float4 o;
      if (MODE == 1) {  // dlpm.py:285-287
        o.x = __fadd_rn(__fdiv_rn(__fsub_rn(xv.x, __fmul_rn(row.w, ev.x)), row.x), __fmul_rn(bs_prev, ev.x));
        o.y = __fadd_rn(__fdiv_rn(__fsub_rn(xv.y, __fmul_rn(row.w, ev.y)), row.x), __fmul_rn(bs_prev, ev.y));
        o.z = __fadd_rn(__fdiv_rn(__fsub_rn(xv.z, __fmul_rn(row.w, ev.z)), row.x), __fmul_rn(bs_prev, ev.z));
        o.w = __fadd_rn(__fdiv_rn(__fsub_rn(xv.w, __fmul_rn(row.w, ev.w)), row.x), __fmul_rn(bs_prev, ev.w));
      } else {
        const float4 zv = z ? ld_stream(reinterpret_cast<const float4*>(z) + q)
                            : normal_quad(ph, STREAM_Z, off_t, (uint64_t)(b + sample_base), pos);
        if (!sig_full) {
          const StepCoef c = dlpm_coef(__ldg(Sigma + (int64_t)(t - 1) * B + b), __ldg(Sigma + (int64_t)t * B + b), row, t);
          if (z) {  // injected noise (parity tests): reference evaluation order, bit-exact
            o.x = dlpm_update1(xv.x, ev.x, zv.x, c); o.y = dlpm_update1(xv.y, ev.y, zv.y, c);
            o.z = dlpm_update1(xv.z, ev.z, zv.z, c); o.w = dlpm_update1(xv.w, ev.w, zv.w, c);
          } else {
            const float inv_g = __frcp_rn(c.g);
            o.x = dlpm_update1_fast(xv.x, ev.x, zv.x, c, inv_g); o.y = dlpm_update1_fast(xv.y, ev.y, zv.y, c, inv_g);
            o.z = dlpm_update1_fast(xv.z, ev.z, zv.z, c, inv_g); o.w = dlpm_update1_fast(xv.w, ev.w, zv.w, c, inv_g);
          }
        } else {
          const float4 s1 = ld_stream(reinterpret_cast<const float4*>(Sigma + (int64_t)(t - 1) * B * D) + q);
          const float4 st = ld_stream(reinterpret_cast<const float4*>(Sigma + (int64_t)t * B * D) + q);
          o.x = dlpm_update1(xv.x, ev.x, zv.x, dlpm_coef(s1.x, st.x, row, t));
          o.y = dlpm_update1(xv.y, ev.y, zv.y, dlpm_coef(s1.y, st.y, row, t));
          o.z = dlpm_update1(xv.z, ev.z, zv.z, dlpm_coef(s1.z, st.z, row, t));
          o.w = dlpm_update1(xv.w, ev.w, zv.w, dlpm_coef(s1.w, st.w, row, t));
        }
      }
      reinterpret_cast<float4*>(x)[q] = o;
      if (hist) st_stream(reinterpret_cast<float4*>(hist) + q, o);
      if (do_post) post_store4(post, q, b, pos, o);
    }
  } else {
    const int64_t n = B * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
      const int64_t b = e / D, i = e - b * D;
      const float xv = x[e];
      float ev = load_eps1<EPS_BF16>(eps, e);
      if (clip) ev = clip_eps1(xv, ev, row);
      float o;
      if (MODE == 1) {
        o = __fadd_rn(__fdiv_rn(__fsub_rn(xv, __fmul_rn(row.w, ev)), row.x), __fmul_rn(bs_prev, ev));
      } else {
        const float zv = z ? z[e] : sel4(normal_quad(ph, STREAM_Z, off_t, (uint64_t)(b + sample_base), (uint32_t)(i >> 2)), (int)(i & 3));
        const float S1 = sig_full ? Sigma[(int64_t)(t - 1) * n + e] : Sigma[(int64_t)(t - 1) * B + b];
        const float St = sig_full ? Sigma[(int64_t)t * n + e] : Sigma[(int64_t)t * B + b];
        o = dlpm_update1(xv, ev, zv, dlpm_coef(S1, St, row, t));
      }
      x[e] = o;
      if (hist) hist[e] = o;
      if (do_post) post_store1(post, b, i, o);
    }
  }
}

// Production variant of the stochastic DLPM step (isotropic, compact Sigma, in-kernel noise, no clipping): the hot loop
// of image sampling, on the QuadSpan skeleton.  The per-sample coefficients (bs*Gamma, sqrt(Gamma*Sigma_{t-1})) of the
// samples touched by a sub-chunk are computed once into shared memory with the reference's exact arithmetic; every
// thread then streams UNROLL quads with all loads issued before the math (memory-level parallelism), z drawn in registers.
template <bool EPS_BF16, int UNROLL, int OCC>
__global__ void __launch_bounds__(256, OCC) k_reverse_step_fast(float* __restrict__ x, const void* __restrict__ eps,
                                                              const float* __restrict__ Sigma, const float* __restrict__ sched,
                                                              int t_imm, const int* __restrict__ t_dev, int T, int64_t B,
                                                              const QuadSpan span, const __grid_constant__ PhiloxKeys keys,
                                                              uint64_t offset, int64_t sample_base, float* __restrict__ hist,
                                                              const StepPost post) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = t_dev ? *t_dev : t_imm;
  if (t < 1 || t >= T) return;
  const bool do_post = post.out != nullptr && t == post.at;
  const PhiloxRef ph(keys);
  const float4 row = __ldg(reinterpret_cast<const float4*>(sched) + t);
  const float inv_g = __frcp_rn(row.x);
  const uint64_t off_t = offset + (uint64_t)t;
  __shared__ float2 s_coef[kChunkQuads];  // (bs*Gamma, 1[t != 1] sqrt(Gamma*Sigma_{t-1})) per sample of the sub-chunk
  uint32_t q_lo, q_hi;
  cta_share(span.nq, q_lo, q_hi);
  for (uint32_t q0 = q_lo; q0 < q_hi; q0 += span.sub) {
    const uint32_t q1 = q0 + span.sub < q_hi ? q0 + span.sub : q_hi;
    const uint32_t b_first = span.fd.div(q0), b_last = span.fd.div(q1 - 1);
    __syncthreads();
    for (uint32_t sI = threadIdx.x; sI <= b_last - b_first; sI += 256) {
      const int64_t b = (int64_t)(b_first + sI);
      const StepCoef c = dlpm_coef(__ldg(Sigma + (int64_t)(t - 1) * B + b), __ldg(Sigma + (int64_t)t * B + b), row, t);
      s_coef[sI] = make_float2(c.c1, c.sd);
    }
    __syncthreads();
    uint32_t qb = q0 + threadIdx.x, o, pos;
    span.fd.divmod(qb, o, pos);
    for (; qb < q1; qb += 256 * UNROLL) {
      float4 xv[UNROLL], ev[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t q = qb + (uint32_t)u * 256u;
        if (q < q1) {
          xv[u] = ld_rw(reinterpret_cast<const float4*>(x) + q);
          ev[u] = load_eps4<true, EPS_BF16>(eps, q);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t q = qb + (uint32_t)u * 256u;
        if (q < q1) {
          const float2 cf = s_coef[o - b_first];
          const float4 zv = normal_quad(ph, STREAM_Z, off_t, (uint64_t)((int64_t)o + sample_base), pos);
          float4 r;
          r.x = fmaf(cf.y, zv.x, fmaf(-cf.x, ev[u].x, xv[u].x) * inv_g);
          r.y = fmaf(cf.y, zv.y, fmaf(-cf.x, ev[u].y, xv[u].y) * inv_g);
          r.z = fmaf(cf.y, zv.z, fmaf(-cf.x, ev[u].z, xv[u].z) * inv_g);
          r.w = fmaf(cf.y, zv.w, fmaf(-cf.x, ev[u].w, xv[u].w) * inv_g);
          reinterpret_cast<float4*>(x)[q] = r;
          if (hist) st_stream(reinterpret_cast<float4*>(hist) + q, r);
          if (do_post) post_store4(post, q, o, pos, r);
        }
        span_advance(span, o, pos);
      }
    }
  }
}

// LIM step (sampler.py:81-181): x <- a x + c_score (sc * out) [+ c_noise e_L]
template <bool VEC, bool EPS_BF16>
__global__ void __launch_bounds__(256) k_lim_step(float* __restrict__ x, const void* __restrict__ mo,
                                                  const float* __restrict__ coef, int step_imm,
                                                  const int* __restrict__ step_dev, int64_t B, int64_t D, int ode,
                                                  int isotropic, StableParams sp, float clamp_eps,
                                                  const float* __restrict__ e_L, uint64_t seed, uint64_t offset,
                                                  int64_t sample_base, float* __restrict__ hist, FastDiv fd, int chunk,
                                                  const StepPost post) {
  const int step = step_dev ? *step_dev : step_imm;
  const bool do_post = post.out != nullptr && step == post.at;
  const Philox ph(seed);
  const float4 cf = __ldg(reinterpret_cast<const float4*>(coef) + step);  // (score_scale, a, c_score, c_noise)
  const uint64_t off_s = offset + (uint64_t)step;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int64_t qpr = D >> 2, nq = B * qpr;
    const bool fast = nq < (1ll << 31);
    const bool iso_draw = !ode && !e_L && isotropic;
    __shared__ float s_sa[kChunkQuads];
    for (int64_t q0 = (int64_t)blockIdx.x * chunk; q0 < nq; q0 += (int64_t)gridDim.x * chunk) {
      const int64_t q1 = q0 + chunk < nq ? q0 + chunk : nq;
      const int64_t b_first = fast ? (int64_t)fd.div((uint32_t)q0) : q0 / qpr;
      if (iso_draw) {
        const int64_t b_last = fast ? (int64_t)fd.div((uint32_t)(q1 - 1)) : (q1 - 1) / qpr;
        __syncthreads();
        for (int64_t sI = threadIdx.x; sI <= b_last - b_first; sI += blockDim.x)
          s_sa[sI] = __fsqrt_rn(sample_A(ph, sp, STREAM_EPS_A, off_s, (uint64_t)(b_first + sI + sample_base)));
        __syncthreads();
      }
      for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        uint32_t b32, pos;
        if (fast) fd.divmod((uint32_t)q, b32, pos);
        const int64_t b = fast ? (int64_t)b32 : q / qpr;
        if (!fast) pos = (uint32_t)(q - b * qpr);
        const uint64_t sample = (uint64_t)(b + sample_base);
        const float4 xv = ld_rw(reinterpret_cast<const float4*>(x) + q);
        const float4 mv = load_eps4<VEC, EPS_BF16>(mo, q);
        float4 o;
        o.x = __fadd_rn(__fmul_rn(cf.y, xv.x), __fmul_rn(cf.z, __fmul_rn(mv.x, cf.x)));
        o.y = __fadd_rn(__fmul_rn(cf.y, xv.y), __fmul_rn(cf.z, __fmul_rn(mv.y, cf.x)));
        o.z = __fadd_rn(__fmul_rn(cf.y, xv.z), __fmul_rn(cf.z, __fmul_rn(mv.z, cf.x)));
        o.w = __fadd_rn(__fmul_rn(cf.y, xv.w), __fmul_rn(cf.z, __fmul_rn(mv.w, cf.x)));
        if (!ode) {
          float4 n4;
          if (e_L) {
            n4 = ld_stream(reinterpret_cast<const float4*>(e_L) + q);
          } else {
            n4 = normal_quad(ph, STREAM_G, off_s, sample, pos);
            if (isotropic) {
              const float sa = s_sa[b - b_first];
              n4.x *= sa; n4.y *= sa; n4.z *= sa; n4.w *= sa;
            } else {
              const float4 a = element_A4(ph, sp, STREAM_EPS_A, off_s, sample, pos);
              n4.x *= __fsqrt_rn(a.x); n4.y *= __fsqrt_rn(a.y); n4.z *= __fsqrt_rn(a.z); n4.w *= __fsqrt_rn(a.w);
            }
            n4.x = clamp_sym(n4.x, clamp_eps); n4.y = clamp_sym(n4.y, clamp_eps);
            n4.z = clamp_sym(n4.z, clamp_eps); n4.w = clamp_sym(n4.w, clamp_eps);
          }
          o.x = __fadd_rn(o.x, __fmul_rn(cf.w, n4.x)); o.y = __fadd_rn(o.y, __fmul_rn(cf.w, n4.y));
          o.z = __fadd_rn(o.z, __fmul_rn(cf.w, n4.z)); o.w = __fadd_rn(o.w, __fmul_rn(cf.w, n4.w));
        }
        reinterpret_cast<float4*>(x)[q] = o;
        if (hist) st_stream(reinterpret_cast<float4*>(hist) + q, o);
        if (do_post) post_store4(post, q, b, pos, o);
      }
    }
  } else {
    const int64_t n = B * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
      const int64_t b = e / D, i = e - b * D;
      const uint64_t sample = (uint64_t)(b + sample_base);
      const float mv = load_eps1<EPS_BF16>(mo, e);
      float o = __fadd_rn(__fmul_rn(cf.y, x[e]), __fmul_rn(cf.z, __fmul_rn(mv, cf.x)));
      if (!ode) {
        float nz;
        if (e_L) {
          nz = e_L[e];
        } else {
          const uint32_t pos = (uint32_t)(i >> 2);
          nz = sel4(normal_quad(ph, STREAM_G, off_s, sample, pos), (int)(i & 3));
          const float a = isotropic ? sample_A(ph, sp, STREAM_EPS_A, off_s, sample)
                                    : sel4(element_A4(ph, sp, STREAM_EPS_A, off_s, sample, pos), (int)(i & 3));
          nz = clamp_sym(nz * __fsqrt_rn(a), clamp_eps);
        }
        o = __fadd_rn(o, __fmul_rn(cf.w, nz));
      }
      x[e] = o;
      if (hist) hist[e] = o;
      if (do_post) post_store1(post, b, i, o);
    }
  }
}

// LIM step on the QuadSpan skeleton (nq < 2^31, 16-byte aligned tensors)
template <bool EPS_BF16>
__global__ void __launch_bounds__(256) k_lim_step_vec(float* __restrict__ x, const void* __restrict__ mo,
                                                      const float* __restrict__ coef, int step_imm,
                                                      const int* __restrict__ step_dev, const QuadSpan span, int ode,
                                                      int isotropic, StableParams sp, float clamp_eps,
                                                      const float* __restrict__ e_L, const __grid_constant__ PhiloxKeys keys,
                                                      uint64_t offset, int64_t sample_base, float* __restrict__ hist,
                                                      const StepPost post) {
  const int step = step_dev ? *step_dev : step_imm;
  const bool do_post = post.out != nullptr && step == post.at;
  const PhiloxRef ph(keys);
  const float4 cf = __ldg(reinterpret_cast<const float4*>(coef) + step);  // (score_scale, a, c_score, c_noise)
  const uint64_t off_s = offset + (uint64_t)step;
  const bool iso_draw = !ode && !e_L && isotropic;
  __shared__ float s_sa[kChunkQuads];
  uint32_t q_lo, q_hi;
  cta_share(span.nq, q_lo, q_hi);
  for (uint32_t q0 = q_lo; q0 < q_hi; q0 += span.sub) {
    const uint32_t q1 = q0 + span.sub < q_hi ? q0 + span.sub : q_hi;
    const uint32_t b_first = span.fd.div(q0);
    if (iso_draw) {
      const uint32_t b_last = span.fd.div(q1 - 1);
      __syncthreads();
      for (uint32_t sI = threadIdx.x; sI <= b_last - b_first; sI += 256)
        s_sa[sI] = __fsqrt_rn(sample_A(ph, sp, STREAM_EPS_A, off_s, (uint64_t)((int64_t)(b_first + sI) + sample_base)));
      __syncthreads();
    }
    uint32_t q = q0 + threadIdx.x, b, pos;
    span.fd.divmod(q, b, pos);
    for (; q < q1; q += 256) {
      const uint64_t sample = (uint64_t)((int64_t)b + sample_base);
      const float4 xv = ld_rw(reinterpret_cast<const float4*>(x) + q);
      const float4 mv = load_eps4<true, EPS_BF16>(mo, q);
      float4 o;
      o.x = __fadd_rn(__fmul_rn(cf.y, xv.x), __fmul_rn(cf.z, __fmul_rn(mv.x, cf.x)));
      o.y = __fadd_rn(__fmul_rn(cf.y, xv.y), __fmul_rn(cf.z, __fmul_rn(mv.y, cf.x)));
      o.z = __fadd_rn(__fmul_rn(cf.y, xv.z), __fmul_rn(cf.z, __fmul_rn(mv.z, cf.x)));
      o.w = __fadd_rn(__fmul_rn(cf.y, xv.w), __fmul_rn(cf.z, __fmul_rn(mv.w, cf.x)));
      if (!ode) {
        float4 n4;
        if (e_L) {
          n4 = ld_stream(reinterpret_cast<const float4*>(e_L) + q);
        } else {
          n4 = normal_quad(ph, STREAM_G, off_s, sample, pos);
          if (isotropic) {
            const float sa = s_sa[b - b_first];
            n4.x *= sa; n4.y *= sa; n4.z *= sa; n4.w *= sa;
          } else {
            const float4 a = element_A4(ph, sp, STREAM_EPS_A, off_s, sample, pos);
            n4.x *= __fsqrt_rn(a.x); n4.y *= __fsqrt_rn(a.y); n4.z *= __fsqrt_rn(a.z); n4.w *= __fsqrt_rn(a.w);
          }
          n4.x = clamp_sym(n4.x, clamp_eps); n4.y = clamp_sym(n4.y, clamp_eps);
          n4.z = clamp_sym(n4.z, clamp_eps); n4.w = clamp_sym(n4.w, clamp_eps);
        }
        o.x = __fadd_rn(o.x, __fmul_rn(cf.w, n4.x)); o.y = __fadd_rn(o.y, __fmul_rn(cf.w, n4.y));
        o.z = __fadd_rn(o.z, __fmul_rn(cf.w, n4.z)); o.w = __fadd_rn(o.w, __fmul_rn(cf.w, n4.w));
      }
      reinterpret_cast<float4*>(x)[q] = o;
      if (hist) st_stream(reinterpret_cast<float4*>(hist) + q, o);
      if (do_post) post_store4(post, q, b, pos, o);
      span_advance(span, b, pos);
    }
  }
}

__global__ void k_advance(int* t, int delta) {
  pdl_launch_dependents();
  pdl_wait();
  *t += delta;
}

// ------------------------------------------------------------------------------------------------
// training forward elements (dlpm.py:384-401)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_training_elements(float* __restrict__ x_t, float* __restrict__ eps_t,
                                                           const float* __restrict__ x0, const int64_t* __restrict__ t,
                                                           const float* __restrict__ A, const float* __restrict__ z,
                                                           const float* __restrict__ sched, int T, int64_t B, int64_t D,
                                                           StableParams sp, float clamp_a, uint64_t seed,
                                                           uint64_t offset, int64_t sample_base) {
  const Philox ph(seed);
  const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
    const int64_t b = e / D, i = e - b * D;
    const uint64_t sample = (uint64_t)(b + sample_base);
    int64_t tb = t[b];
    tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
    const float4 row = __ldg(reinterpret_cast<const float4*>(sched) + tb);
    const float a = A ? A[b] : clamp_A(sample_A(ph, sp, STREAM_A, offset, sample), clamp_a);
    const float zz = z ? z[e] : sel4(normal_quad(ph, STREAM_Z, offset, sample, (uint32_t)(i >> 2)), (int)(i & 3));
    const float Sig = __fmul_rn(a, __fmul_rn(row.w, row.w));        // a_t * bs[t]**2
    const float xt = __fadd_rn(__fmul_rn(row.y, x0[e]), __fmul_rn(__fsqrt_rn(Sig), zz));
    x_t[e] = xt;
    eps_t[e] = __fdiv_rn(__fsub_rn(xt, __fmul_rn(x0[e], row.y)), row.w);
  }
}

// model-input scaling of the exploding schedule: one CTA column per sample row
__global__ void __launch_bounds__(256) k_scale_by_step(float* __restrict__ out, const float* __restrict__ x,
                                                       const float* __restrict__ table, const int64_t* __restrict__ t_vec,
                                                       int t_imm, const int* __restrict__ t_dev, int T, int64_t B, int64_t D) {
  const int64_t n = B * D, stride = (int64_t)gridDim.x * blockDim.x;
  const int t_const = t_dev ? *t_dev : t_imm;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
    int64_t tb = t_vec ? t_vec[e / D] : (int64_t)t_const;
    tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
    out[e] = __fmul_rn(x[e], __ldg(table + tb));
  }
}

// LIM training elements: one CTA per sample; the per-sample VPSDE coefficients use the precise libm functions
__global__ void __launch_bounds__(256) k_lim_training_elements(float* __restrict__ x_t, float* __restrict__ score,
                                                               const float* __restrict__ x0, const float* __restrict__ t,
                                                               const float* __restrict__ e_in, int64_t D, float alpha,
                                                               int isotropic, StableParams sp, float clamp_eps, uint64_t seed,
                                                               uint64_t offset, int64_t sample_base) {
  const Philox ph(seed);
  const int64_t b = blockIdx.x;
  const uint64_t sample = (uint64_t)(b + sample_base);
  const float s = 0.008f;
  const float HALF_PI = 1.57079632679489661923f;
  // sde.py:41-47 in the reference's op order: log(cos((t + s) / (1 + s) * pi / 2)) - log_alpha_0
  const float l0 = (float)log(cos((double)0.008 / (1.0 + 0.008) * 3.14159265358979323846 / 2.0));
  const float la = __fsub_rn(logf(cosf(__fmul_rn(__fdiv_rn(__fadd_rn(t[b], s), 1.008f), HALF_PI))), l0);
  const float x_coeff = expf(la);
  const float sigma = powf(__fsub_rn(1.0f, expf(__fmul_rn(la, alpha))), 1.0f / alpha);
  const float sa_iso = (!e_in && isotropic) ? __fsqrt_rn(sample_A(ph, sp, STREAM_EPS_A, offset, sample)) : 0.f;
  for (int64_t i = threadIdx.x; i < D; i += blockDim.x) {
    float ev;
    if (e_in) {
      ev = e_in[b * D + i];
    } else {
      const uint32_t pos = (uint32_t)(i >> 2);
      const float g = sel4(normal_quad(ph, STREAM_G, offset, sample, pos), (int)(i & 3));
      const float sa = isotropic ? sa_iso : __fsqrt_rn(sel4(element_A4(ph, sp, STREAM_EPS_A, offset, sample, pos), (int)(i & 3)));
      ev = clamp_sym(g * sa, clamp_eps);
    }
    x_t[b * D + i] = __fadd_rn(__fmul_rn(x0[b * D + i], x_coeff), __fmul_rn(ev, sigma));
    score[b * D + i] = __fdiv_rn(-ev, alpha);
  }
}

template <bool PRED_BF16>
__global__ void __launch_bounds__(256) k_loss_terms(float* __restrict__ out, const void* __restrict__ pred,
                                                    const float* __restrict__ target, int64_t D, float lploss) {
  const int64_t b = blockIdx.x;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < D; i += blockDim.x) {
    const float d = load_eps1<PRED_BF16>(pred, b * D + i) - target[b * D + i];
    if (lploss == 1.0f) { const float ad = fabsf(d); acc += ad < 1.f ? 0.5f * d * d : ad - 0.5f; }
    else acc += d * d;
  }
  __shared__ float red[8];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float m = tot / (float)D;
    out[b] = (lploss == 2.0f) ? sqrtf(m) : m;
  }
}

__global__ void __launch_bounds__(256) k_postprocess(float* __restrict__ out, const float* __restrict__ x, int64_t n,
                                                     float clamp, int is_image) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += stride) {
    float v = fminf(fmaxf(x[e], -clamp), clamp);
    if (is_image) v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
    out[e] = v;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// chunk size (quads per CTA iteration, multiple of 256, <= kChunkQuads) and grid for the chunked streaming kernels:
// aim for >= 8 CTAs per SM when the problem is large enough, never more CTAs than chunks.
static inline void chunk_grid(int64_t nq, int* chunk, int* grid) {
  int64_t c = nq / ((int64_t)kNumSMs * 8);
  c = (c / 256) * 256;
  if (c < 256) c = 256;
  if (c > kChunkQuads) c = kChunkQuads;
  int64_t g = (nq + c - 1) / c;
  if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
  if (g < 1) g = 1;
  *chunk = (int)c;
  *grid = (int)g;
}

static int g_sextet_enabled = 1;  // "noise_sextet": 0 = every fill uses the quad scheme (four normals per Philox block)
static int g_k3_variant = 0;   // experiment switches (dlpm_b200_set_option): K3 unroll/occupancy variant,
static int g_stream_ctas = 0;  // CTAs per SM override for the QuadSpan kernels (0 = per-kernel default)
bool process_set_option(const char* name, int value) {
  const std::string n(name);
  if (n == "k3_variant") { g_k3_variant = value; return true; }
  if (n == "stream_ctas") { g_stream_ctas = value; return true; }
  if (n == "noise_sextet") { g_sextet_enabled = value != 0; return true; }
  return false;
}

// one full wave for the QuadSpan kernels: SMs x resident CTAs, never more CTAs than 256-quad tiles
static inline int span_grid(const QuadSpan& sp, int ctas_per_sm) {
  int64_t tiles = ((int64_t)sp.nq + 255) / 256;
  int64_t g = (int64_t)kNumSMs * (g_stream_ctas > 0 ? g_stream_ctas : ctas_per_sm);
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}
static inline bool span_ok(int64_t n_outer, int64_t inner) { return inner % 4 == 0 && n_outer * (inner / 4) < (1ll << 31); }

}  // namespace dlpm

using namespace dlpm;

// C linkage comes from the declarations in include/dlpm_b200.h

int dlpm_b200_abi_version(void) { return DLPM_B200_ABI_VERSION; }
const char* dlpm_b200_last_error(void) { return g_err; }

int dlpm_b200_stable_A(float* out, int64_t n_outer, int64_t inner, int mode, float alpha, float clamp_a, uint64_t seed,
                       uint64_t offset, int64_t sample_base, void* stream) {
  DLPM_REQUIRE(out != nullptr || n_outer == 0, "stable_A: out is NULL");
  DLPM_REQUIRE(alpha > 0.f && alpha <= 2.f, "Wrong value of alpha for skewed levy r.v generation");
  DLPM_REQUIRE(mode >= 0 && mode <= 2 && n_outer >= 0 && inner >= 1, "stable_A: bad mode/size");
  if (n_outer == 0) return DLPM_OK;
  const StableParams sp = make_params(alpha);
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = mode != DLPM_A_COMPACT && (inner % 4 == 0) && aligned16(out);
  const int64_t items = mode == DLPM_A_COMPACT ? n_outer : (vec ? n_outer * inner / 4 : n_outer * inner);
  int grid = grid_for(items, 256), chunk = 256;
  if (vec) chunk_grid(items, &chunk, &grid);
  const FastDiv fd((uint32_t)(vec ? inner / 4 : 1));
  if (vec && span_ok(n_outer, inner)) {
    const QuadSpan span = make_span(n_outer, inner);
    k_stable_A_vec<<<span_grid(span, 8), 256, 0, s>>>(out, span, mode, sp, clamp_a, make_philox_keys(seed), offset, sample_base);
  } else if (vec) k_stable_A<true><<<grid, 256, 0, s>>>(out, n_outer, inner, mode, sp, clamp_a, seed, offset, sample_base, fd, chunk);
  else k_stable_A<false><<<grid, 256, 0, s>>>(out, n_outer, inner, mode, sp, clamp_a, seed, offset, sample_base, fd, chunk);
  DLPM_CHECK_LAUNCH("stable_A");
  return DLPM_OK;
}

template <int A_MODE>
static void launch_sas(bool vec, int grid, int chunk, cudaStream_t s, float* out, const float* A_in, int64_t n_outer, int64_t inner,
                       const StableParams& sp, float clamp_eps, float scale, uint32_t g_stream, uint64_t seed,
                       uint64_t offset, int64_t sample_base) {
  const FastDiv fd((uint32_t)(vec ? inner / 4 : 1));
  // plain normal / isotropic fields whose rows are multiples of 384 elements use the "sextet" scheme (rng.cuh): a property of
  // (mode, row length) only, so every kernel path below produces the same field
  const bool sext = (A_MODE == 0 || A_MODE == 1 || A_MODE == 3) && inner % 384 == 0 && g_sextet_enabled;
  if (vec && span_ok(n_outer, inner) && sext) {
    if constexpr (A_MODE == 0 || A_MODE == 1 || A_MODE == 3) {
      const GranSpan span = make_gran_span(n_outer, inner);
      int64_t gsz = (int64_t)kNumSMs * (g_stream_ctas > 0 ? g_stream_ctas : 8);
      if (gsz > ((int64_t)span.ng + 7) / 8) gsz = ((int64_t)span.ng + 7) / 8;
      if (gsz < 1) gsz = 1;
      if (scale == 1.0f) k_fill6<A_MODE, false><<<(int)gsz, 256, 0, s>>>(out, A_in, span, sp, clamp_eps, scale, g_stream, make_philox_keys(seed), offset, sample_base);
      else k_fill6<A_MODE, true><<<(int)gsz, 256, 0, s>>>(out, A_in, span, sp, clamp_eps, scale, g_stream, make_philox_keys(seed), offset, sample_base);
    }
  } else if (vec && span_ok(n_outer, inner) && !sext) {
    const QuadSpan span = make_span(n_outer, inner);
    if (scale == 1.0f) k_sas_vec<A_MODE, false><<<span_grid(span, 8), 256, 0, s>>>(out, A_in, span, sp, clamp_eps, scale, g_stream, make_philox_keys(seed), offset, sample_base);
    else k_sas_vec<A_MODE, true><<<span_grid(span, 8), 256, 0, s>>>(out, A_in, span, sp, clamp_eps, scale, g_stream, make_philox_keys(seed), offset, sample_base);
  } else if (vec) k_sas<true, A_MODE><<<grid, 256, 0, s>>>(out, A_in, n_outer, inner, sp, clamp_eps, scale, g_stream, seed, offset, sample_base, fd, chunk, sext ? 1 : 0);
  else k_sas<false, A_MODE><<<grid, 256, 0, s>>>(out, A_in, n_outer, inner, sp, clamp_eps, scale, g_stream, seed, offset, sample_base, fd, chunk, sext ? 1 : 0);
}

int dlpm_b200_sas(float* out, const float* A_in, int64_t n_outer, int64_t inner, int isotropic, float alpha,
                  float clamp_eps, float scale, uint64_t seed, uint64_t offset, int64_t sample_base, void* stream) {
  DLPM_REQUIRE(out != nullptr || n_outer == 0, "sas: out is NULL");
  DLPM_REQUIRE(alpha > 0.f && alpha <= 2.f, "Wrong value of alpha for skewed levy r.v generation");
  DLPM_REQUIRE(n_outer >= 0 && inner >= 1, "sas: bad size");
  if (n_outer == 0) return DLPM_OK;
  const StableParams sp = make_params(alpha);
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (inner % 4 == 0) && aligned16(out) && (A_in == nullptr || isotropic || aligned16(A_in));
  int grid = grid_for(n_outer * inner, 256), chunk = 256;
  if (vec) chunk_grid(n_outer * inner / 4, &chunk, &grid);
  const int mode = A_in ? (isotropic ? 3 : 4) : (isotropic ? 1 : 2);
  switch (mode) {
    case 1: launch_sas<1>(vec, grid, chunk, s, out, A_in, n_outer, inner, sp, clamp_eps, scale, STREAM_G, seed, offset, sample_base); break;
    case 2: launch_sas<2>(vec, grid, chunk, s, out, A_in, n_outer, inner, sp, clamp_eps, scale, STREAM_G, seed, offset, sample_base); break;
    case 3: launch_sas<3>(vec, grid, chunk, s, out, A_in, n_outer, inner, sp, clamp_eps, scale, STREAM_G, seed, offset, sample_base); break;
    default: launch_sas<4>(vec, grid, chunk, s, out, A_in, n_outer, inner, sp, clamp_eps, scale, STREAM_G, seed, offset, sample_base); break;
  }
  DLPM_CHECK_LAUNCH("sas");
  return DLPM_OK;
}

int dlpm_b200_normal(float* out, int64_t n_outer, int64_t inner, uint64_t seed, uint64_t offset, int64_t sample_base,
                     void* stream) {
  DLPM_REQUIRE(out != nullptr || n_outer == 0, "normal: out is NULL");
  DLPM_REQUIRE(n_outer >= 0 && inner >= 1, "normal: bad size");
  if (n_outer == 0) return DLPM_OK;
  StableParams sp = make_params(2.0f);
  const bool vec = (inner % 4 == 0) && aligned16(out);
  int grid = grid_for(n_outer * inner, 256), chunk = 256;
  if (vec) chunk_grid(n_outer * inner / 4, &chunk, &grid);
  launch_sas<0>(vec, grid, chunk, (cudaStream_t)stream, out, nullptr, n_outer, inner, sp, -1.f, 1.f, STREAM_Z, seed, offset, sample_base);
  DLPM_CHECK_LAUNCH("normal");
  return DLPM_OK;
}

int dlpm_b200_sigma_scan(float* Sigma, const float* A_in, float* A_out, const float* sched, int T, int64_t n,
                         int64_t inner, int per_element, float alpha, float clamp_a, uint64_t seed, uint64_t offset,
                         int64_t sample_base, void* stream) {
  DLPM_REQUIRE(Sigma && sched && T >= 1 && n >= 0, "sigma_scan: bad arguments");
  DLPM_REQUIRE(alpha > 0.f && alpha <= 2.f, "Wrong value of alpha for skewed levy r.v generation");
  DLPM_REQUIRE(!per_element || inner >= 1, "sigma_scan: inner must be >= 1");
  if (n == 0) return DLPM_OK;
  const StableParams sp = make_params(alpha);
  const int grid = (int)((n + 127) / 128);
  k_sigma_scan<<<grid, 128, 0, (cudaStream_t)stream>>>(Sigma, A_in, A_out, sched, T, n, inner < 1 ? 1 : inner, per_element, sp,
                                                       clamp_a, seed, offset, sample_base);
  DLPM_CHECK_LAUNCH("sigma_scan");
  return DLPM_OK;
}

template <int MODE>
static int launch_step(float* x, const void* eps, const float* Sigma, const float* sched, int t, const int* t_dev, int T,
                       int64_t B, int64_t D, int flags, const float* z, uint64_t seed, uint64_t offset, int64_t sample_base,
                       float* hist, const StepPost& post, void* stream) {
  const bool bf16 = flags & DLPM_STEP_EPS_BF16;
  const bool vec = (D % 4 == 0) && aligned16(x) && (reinterpret_cast<uintptr_t>(eps) % (bf16 ? 8 : 16) == 0) &&
                   (!z || aligned16(z)) && (!hist || aligned16(hist)) &&
                   (!(flags & DLPM_STEP_SIGMA_FULL) || (aligned16(Sigma) && (B * D) % 4 == 0));
  const int grid = grid_for(vec ? B * D / 4 : B * D, 256);
  cudaStream_t s = (cudaStream_t)stream;
  const FastDiv fd((uint32_t)(vec ? D / 4 : 1));
  if (MODE == 0 && vec && !z && !(flags & (DLPM_STEP_CLIP_DENOISED | DLPM_STEP_SIGMA_FULL)) && span_ok(B, D)) {
    const QuadSpan span = make_span(B, D);
    const PhiloxKeys keys = make_philox_keys(seed);
#define LF(U, O)                                                                                                              \
  do {                                                                                                                        \
    const int fgrid = span_grid(span, g_stream_ctas > 0 ? g_stream_ctas : O);                                                 \
    if (bf16) launch_ex(k_reverse_step_fast<true, U, O>, dim3(fgrid), dim3(256), 0, s, 1, x, eps, Sigma, sched, t, t_dev, T, B, span, keys, offset, sample_base, hist, post); \
    else launch_ex(k_reverse_step_fast<false, U, O>, dim3(fgrid), dim3(256), 0, s, 1, x, eps, Sigma, sched, t, t_dev, T, B, span, keys, offset, sample_base, hist, post);     \
  } while (0)
    switch (g_k3_variant) {  // measured within 4 % of each other at B = 4096 (tools/bench_stream.py); (2, 6) is the default
      case 1: LF(4, 4); break;
      case 2: LF(2, 8); break;
      case 3: LF(1, 8); break;
      default: LF(2, 6); break;
    }
#undef LF
    DLPM_CHECK_LAUNCH("reverse_step");
    return DLPM_OK;
  }
#define L(V, H) launch_ex(k_reverse_step<V, H, MODE>, dim3(grid), dim3(256), 0, s, 1, x, eps, Sigma, sched, t, t_dev, T, B, D, flags, z, seed, offset, sample_base, hist, fd, post)
  if (vec) { if (bf16) L(true, true); else L(true, false); }
  else { if (bf16) L(false, true); else L(false, false); }
#undef L
  DLPM_CHECK_LAUNCH("reverse_step");
  return DLPM_OK;
}

// host form of dlpm_b200_post_t -> StepPost (validated)
static int make_post(const dlpm_b200_post_t* post, int at, int64_t D, StepPost* out) {
  StepPost p;
  p.out = nullptr; p.clamp = 0.f; p.mode = 0; p.at = at; p.C = 1; p.HW = (int)D;
  if (post && post->out && post->mode != DLPM_POST_NONE) {
    DLPM_REQUIRE(post->mode == DLPM_POST_F32 || post->mode == DLPM_POST_F32_IMAGE || post->mode == DLPM_POST_U8_NHWC,
                 "post: unknown mode");
    DLPM_REQUIRE(post->clamp > 0.f, "post: clamp must be positive");
    p.out = post->out; p.clamp = post->clamp; p.mode = post->mode;
    if (post->mode == DLPM_POST_U8_NHWC) {
      DLPM_REQUIRE(post->channels >= 1 && D % post->channels == 0 && D < (1ll << 31), "post: channels must divide D");
      p.C = post->channels; p.HW = (int)(D / post->channels);
    } else {
      DLPM_REQUIRE((reinterpret_cast<uintptr_t>(post->out) & 15u) == 0, "post: fp32 output must be 16-byte aligned");
    }
  }
  *out = p;
  return DLPM_OK;
}

int dlpm_b200_reverse_step_post(float* x, const void* eps, const float* Sigma, const float* sched, int t, const int* t_dev,
                                int T, int64_t B, int64_t D, int flags, const float* z, uint64_t seed, uint64_t offset,
                                int64_t sample_base, float* hist_out, const dlpm_b200_post_t* post, void* stream) {
  DLPM_REQUIRE(x && eps && Sigma && sched, "reverse_step: NULL tensor");
  DLPM_REQUIRE(T >= 2 && B >= 0 && D >= 1, "reverse_step: bad sizes");
  DLPM_REQUIRE(t_dev || (t >= 1 && t < T), "reverse_step: t out of range [1, T)");
  if (B == 0) return DLPM_OK;
  StepPost sp;
  if (int rc = make_post(post, 1, D, &sp)) return rc;
  return launch_step<0>(x, eps, Sigma, sched, t, t_dev, T, B, D, flags, z, seed, offset, sample_base, hist_out, sp, stream);
}

int dlpm_b200_reverse_step(float* x, const void* eps, const float* Sigma, const float* sched, int t, const int* t_dev,
                           int T, int64_t B, int64_t D, int flags, const float* z, uint64_t seed, uint64_t offset,
                           int64_t sample_base, float* hist_out, void* stream) {
  return dlpm_b200_reverse_step_post(x, eps, Sigma, sched, t, t_dev, T, B, D, flags, z, seed, offset, sample_base, hist_out,
                                     nullptr, stream);
}

int dlpm_b200_dlim_step_post(float* x, const void* eps, const float* sched, int t, const int* t_dev, int T, int64_t B,
                             int64_t D, int flags, float* hist_out, const dlpm_b200_post_t* post, void* stream) {
  DLPM_REQUIRE(x && eps && sched, "dlim_step: NULL tensor");
  DLPM_REQUIRE(T >= 2 && B >= 0 && D >= 1, "dlim_step: bad sizes");
  DLPM_REQUIRE(t_dev || (t >= 1 && t < T), "dlim_step: t out of range [1, T)");
  if (B == 0) return DLPM_OK;
  StepPost sp;
  if (int rc = make_post(post, 1, D, &sp)) return rc;
  return launch_step<1>(x, eps, sched /*unused Sigma*/, sched, t, t_dev, T, B, D, flags & ~DLPM_STEP_SIGMA_FULL, nullptr, 0, 0, 0,
                        hist_out, sp, stream);
}

int dlpm_b200_dlim_step(float* x, const void* eps, const float* sched, int t, const int* t_dev, int T, int64_t B,
                        int64_t D, int flags, float* hist_out, void* stream) {
  return dlpm_b200_dlim_step_post(x, eps, sched, t, t_dev, T, B, D, flags, hist_out, nullptr, stream);
}

int dlpm_b200_lim_step_post(float* x, const void* model_out, const float* coef, int step, const int* step_dev, int64_t B,
                            int64_t D, int flags, int ode, int isotropic, float alpha, float clamp_eps, const float* e_L,
                            uint64_t seed, uint64_t offset, int64_t sample_base, float* hist_out, const dlpm_b200_post_t* post,
                            int last_step, void* stream) {
  DLPM_REQUIRE(x && model_out && coef, "lim_step: NULL tensor");
  DLPM_REQUIRE(B >= 0 && D >= 1 && (step_dev || step >= 0), "lim_step: bad sizes");
  DLPM_REQUIRE(alpha > 0.f && alpha < 2.f, "lim_step: heavy-tailed branch only (0 < alpha < 2)");
  if (B == 0) return DLPM_OK;
  StepPost pp;
  if (int rc = make_post(post, last_step, D, &pp)) return rc;
  const StableParams sp = make_params(alpha);
  const bool bf16 = flags & DLPM_STEP_EPS_BF16;
  const bool vec = (D % 4 == 0) && aligned16(x) && (reinterpret_cast<uintptr_t>(model_out) % (bf16 ? 8 : 16) == 0) &&
                   (!e_L || aligned16(e_L)) && (!hist_out || aligned16(hist_out));
  int grid = grid_for(B * D, 256), chunk = 256;
  if (vec) chunk_grid(B * D / 4, &chunk, &grid);
  cudaStream_t s = (cudaStream_t)stream;
  const FastDiv fd((uint32_t)(vec ? D / 4 : 1));
#define L(V, H) k_lim_step<V, H><<<grid, 256, 0, s>>>(x, model_out, coef, step, step_dev, B, D, ode, isotropic, sp, clamp_eps, e_L, seed, offset, sample_base, hist_out, fd, chunk, pp)
  if (vec && span_ok(B, D)) {
    const QuadSpan span = make_span(B, D);
    const PhiloxKeys keys = make_philox_keys(seed);
    const int g = span_grid(span, 6);
    if (bf16) k_lim_step_vec<true><<<g, 256, 0, s>>>(x, model_out, coef, step, step_dev, span, ode, isotropic, sp, clamp_eps, e_L, keys, offset, sample_base, hist_out, pp);
    else k_lim_step_vec<false><<<g, 256, 0, s>>>(x, model_out, coef, step, step_dev, span, ode, isotropic, sp, clamp_eps, e_L, keys, offset, sample_base, hist_out, pp);
  } else if (vec) { if (bf16) L(true, true); else L(true, false); }
  else { if (bf16) L(false, true); else L(false, false); }
#undef L
  DLPM_CHECK_LAUNCH("lim_step");
  return DLPM_OK;
}

int dlpm_b200_lim_step(float* x, const void* model_out, const float* coef, int step, const int* step_dev, int64_t B,
                       int64_t D, int flags, int ode, int isotropic, float alpha, float clamp_eps, const float* e_L,
                       uint64_t seed, uint64_t offset, int64_t sample_base, float* hist_out, void* stream) {
  return dlpm_b200_lim_step_post(x, model_out, coef, step, step_dev, B, D, flags, ode, isotropic, alpha, clamp_eps, e_L, seed, offset,
                                 sample_base, hist_out, nullptr, 0, stream);
}

__global__ void k_set_counter(int* t, int value) { *t = value; }
int dlpm_b200_set_counter(int* t_dev, int value, void* stream) {
  DLPM_REQUIRE(t_dev, "set_counter: NULL");
  k_set_counter<<<1, 1, 0, (cudaStream_t)stream>>>(t_dev, value);
  DLPM_CHECK_LAUNCH("set_counter");
  return DLPM_OK;
}

int dlpm_b200_philox_rounds(void) { return DLPM_PHILOX_ROUNDS; }

int dlpm_b200_advance_counter(int* t_dev, int delta, void* stream) {
  DLPM_REQUIRE(t_dev, "advance_counter: NULL");
  launch_ex(k_advance, dim3(1), dim3(1), 0, (cudaStream_t)stream, 1, t_dev, delta);
  DLPM_CHECK_LAUNCH("advance_counter");
  return DLPM_OK;
}

int dlpm_b200_training_elements(float* x_t, float* eps_t, const float* x0, const int64_t* t, const float* A,
                                const float* z, const float* sched, int T, int64_t B, int64_t D, float alpha,
                                float clamp_a, uint64_t seed, uint64_t offset, int64_t sample_base, void* stream) {
  DLPM_REQUIRE(x_t && eps_t && x0 && t && sched, "training_elements: NULL tensor");
  DLPM_REQUIRE(alpha > 0.f && alpha <= 2.f, "Wrong value of alpha for skewed levy r.v generation");
  DLPM_REQUIRE(T >= 1 && B >= 0 && D >= 1, "training_elements: bad sizes");
  if (B == 0) return DLPM_OK;
  const StableParams sp = make_params(alpha);
  k_training_elements<<<grid_for(B * D, 256), 256, 0, (cudaStream_t)stream>>>(x_t, eps_t, x0, t, A, z, sched, T, B, D, sp,
                                                                              clamp_a, seed, offset, sample_base);
  DLPM_CHECK_LAUNCH("training_elements");
  return DLPM_OK;
}

int dlpm_b200_scale_by_step(float* out, const float* x, const float* table, const int64_t* t_vec, int t, const int* t_dev,
                            int T, int64_t B, int64_t D, void* stream) {
  DLPM_REQUIRE(out && x && table, "scale_by_step: NULL tensor");
  DLPM_REQUIRE(T >= 1 && B >= 0 && D >= 1, "scale_by_step: bad sizes");
  DLPM_REQUIRE(t_vec || t_dev || (t >= 0 && t < T), "scale_by_step: t out of range [0, T)");
  if (B == 0) return DLPM_OK;
  k_scale_by_step<<<grid_for(B * D, 256), 256, 0, (cudaStream_t)stream>>>(out, x, table, t_vec, t, t_dev, T, B, D);
  DLPM_CHECK_LAUNCH("scale_by_step");
  return DLPM_OK;
}

int dlpm_b200_lim_training_elements(float* x_t, float* score, const float* x0, const float* t, const float* e, int64_t B,
                                    int64_t D, float alpha, int isotropic, float clamp_eps, uint64_t seed, uint64_t offset,
                                    int64_t sample_base, void* stream) {
  DLPM_REQUIRE(x_t && score && x0 && t, "lim_training_elements: NULL tensor");
  DLPM_REQUIRE(alpha > 0.f && alpha < 2.f, "lim_training_elements: heavy-tailed branch only (0 < alpha < 2)");
  DLPM_REQUIRE(B >= 0 && B < (1ll << 31) && D >= 1, "lim_training_elements: bad sizes");
  if (B == 0) return DLPM_OK;
  k_lim_training_elements<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(x_t, score, x0, t, e, D, alpha, isotropic, make_params(alpha),
                                                                         clamp_eps, seed, offset, sample_base);
  DLPM_CHECK_LAUNCH("lim_training_elements");
  return DLPM_OK;
}

int dlpm_b200_loss_terms(float* out, const void* pred, const float* target, int64_t B, int64_t D, float lploss, int flags,
                         void* stream) {
  DLPM_REQUIRE(out && pred && target, "loss_terms: NULL tensor");
  DLPM_REQUIRE(lploss == 2.0f || lploss == 1.0f || lploss == -1.0f, "loss_terms: lploss must be 2, 1 or -1");
  DLPM_REQUIRE(B >= 0 && D >= 1 && B < (1ll << 31), "loss_terms: bad sizes");
  if (B == 0) return DLPM_OK;
  if (flags & DLPM_STEP_EPS_BF16) k_loss_terms<true><<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(out, pred, target, D, lploss);
  else k_loss_terms<false><<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(out, pred, target, D, lploss);
  DLPM_CHECK_LAUNCH("loss_terms");
  return DLPM_OK;
}

int dlpm_b200_postprocess(float* out, const float* x, int64_t n, float clamp, int is_image, void* stream) {
  DLPM_REQUIRE((out && x) || n == 0, "postprocess: NULL tensor");
  if (n <= 0) return DLPM_OK;
  k_postprocess<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(out, x, n, clamp, is_image);
  DLPM_CHECK_LAUNCH("postprocess");
  return DLPM_OK;
}

