// Counter-based RNG and alpha-stable transforms (device side).
//
// Replaces the reference's host-side scipy draw + H2D copy
// (bem/datasets/Distributions.py:45-51: scipy.stats.levy_stable.rvs -> Chambers-Mallows-Stuck)
// and torch.randn (Distributions.py:65, GenerativeLevyProcess.py:236) with a stateless
// Philox4x32-R generator (R = DLPM_PHILOX_ROUNDS): every variate is a pure function of
//   (seed, stream tag, call offset, GLOBAL sample index, position inside the sample)
// so results do not depend on grid shape or on how the batch is sharded over GPUs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dlpm {

enum : uint32_t { STREAM_A = 0x0Au, STREAM_G = 0x06u, STREAM_Z = 0x5Au, STREAM_EPS_A = 0xEAu };

// The ten round keys (k + r * Weyl constant).  The hot streaming kernels take them precomputed on the host as a
// __grid_constant__ parameter, so every round's key is a constant-bank operand of the LOP3 (no per-iteration
// UIADD3 key schedule); the other kernels build them in registers from the seed.  Both give the same stream.
#ifndef DLPM_PHILOX_ROUNDS
#define DLPM_PHILOX_ROUNDS 7  // Philox4x32-7: the fewest rounds that pass BigCrush (Salmon et al. 2011, Table 2: "Crush-resistant").
                              // Random123 / cuRAND default to 10 for margin; nothing in the reference pins a generator (it draws from
                              // numpy MT19937 + ATen Philox), and the fills are bound by the quarter-rate IMAD.WIDE of the rounds
                              // (profiles/r01_ncu_stream.md: 10 rounds 0.54 of the HBM copy peak, 7 rounds 0.77), so the default is 7.
                              // -DDLPM_PHILOX_ROUNDS=10 (DLPM_B200_NVCC_EXTRA) rebuilds with the cuRAND count; oracle/philox.py
                              // asks the library (dlpm_b200_philox_rounds) and carries the Random123 known answers for both.
#endif
struct PhiloxKeys {
  uint32_t a[10], b[10];
};
__host__ __device__ __forceinline__ PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.a[r] = a;
    k.b[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

// DLPM_PHILOX_ROUNDS rounds, Salmon et al. 2011 constants.  One round = 2 IMAD.WIDE + 2 three-input XORs.
__device__ __forceinline__ uint4 philox_rounds(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < DLPM_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.a[r];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.b[r];
    c3 = (uint32_t)p0;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct Philox {  // keys in registers, built from the seed
  PhiloxKeys k;
  __device__ __forceinline__ Philox(uint64_t seed) : k(make_philox_keys(seed)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    return philox_rounds(k, c0, c1, c2, c3);
  }
};
struct PhiloxRef {  // keys in the kernel's parameter space (constant bank)
  const PhiloxKeys& k;
  __device__ __forceinline__ PhiloxRef(const PhiloxKeys& keys) : k(keys) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    return philox_rounds(k, c0, c1, c2, c3);
  }
};

// Counter layout shared by every kernel (documented in DESIGN.md):
//   c0 = position inside the sample (in units of 4 variates), c1 = global sample index (low 32),
//   c2 = call offset / diffusion step (low 32), c3 = stream tag | sample-index bits 32..39 << 8 | offset bits 32..47 << 16
template <class P>
__device__ __forceinline__ uint4 philox_at(const P& ph, uint32_t stream, uint64_t offset, uint64_t sample, uint32_t pos) {
  const uint32_t c3 = stream | ((uint32_t)((sample >> 32) & 0xFFu) << 8) | ((uint32_t)((offset >> 32) & 0xFFFFu) << 16);
  return ph(pos, (uint32_t)sample, (uint32_t)offset, c3);
}

// uniform in (0, 1] on the 32-bit lattice, centred: never 0 so log() is finite; small values are
// exact, which is what the Gaussian tail (|z| up to 6.6) needs.
// ((float)x + 0.5f) * 2^-32 is at most exactly 1.0f (x rounds up to 2^32), so no clamp is needed.
__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; }

// uniform in [0, 1) with 23 bits, ALU-only (no I2F on the XU pipe).
__device__ __forceinline__ float u01_fast(uint32_t x) { return __uint_as_float((x >> 9) | 0x3f800000u) - 1.0f; }

// MUFU.LG2 without the denormal pre-scaling sequence of __log2f (3 extra instructions): arguments here are >= 2^-33.
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two N(0,1) from two 32-bit words (Box-Muller: lg2 + sqrt + sin + cos = 4 MUFU ops per pair).
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
  const float r = sqrt_approx(-1.3862943611198906f * lg2_ftz(u01(x)));  // sqrt(-2 ln u), ln u = ln2 * lg2 u
  float s, c;
  // angle 2 pi m with m in [1, 2): one turn ahead of 2 pi (m - 1), same sine / cosine, saves the subtraction
  __sincosf(6.28318530717958647692f * __uint_as_float((y >> 9) | 0x3f800000u), &s, &c);
  return make_float2(r * c, r * s);
}

__device__ __forceinline__ float4 normal4(uint4 r) {
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// "Sextet" scheme of the K1 fills (plain normal / isotropic SaS tensors whose rows are multiples of 384 elements): SIX N(0,1)
// from one Philox block instead of four.  The fills are bound by instruction dispatch, a third of it the 14 IMAD.WIDE of
// Philox4x32-7 (profiles/r02_noise.md); Box-Muller does not need 32 + 23 random bits per pair:
//   pair j = 0, 1, 2:  radius uniform u = (k + 1/2) 2^-27 from the TOP 27 bits k of word j  (|z| <= sqrt(2 * 28 ln 2) = 6.23;
//                      the 32-bit lattice of the quad scheme reaches 6.66 -- 4e-10 of the mass lies beyond 6.23),
//                      angle 2 pi a / 2^15 with the 15-bit a = (bits [10 j, 10 j + 10) of word 3) << 5 | (low 5 bits of word j):
//                      32768 directions; the marginal law of r cos(theta) on a lattice of N directions differs from the normal
//                      law only by Bessel terms J_N(t r), i.e. not at all in fp32 for N = 2^15, and radius and angle use
//                      disjoint bits.
// Row layout (pure function of the position, independent of the grid): a row is cut into granules of 96 quads (384 elements);
// generator g = granule * 32 + lane draws the Philox blocks at positions 2 g and 2 g + 1 = twelve normals n[0..12), and quad
// granule * 96 + 32 j + lane (j = 0, 1, 2) holds n[4 j .. 4 j + 4): a warp writes three fully coalesced 512-byte segments.
// oracle/philox.py::normal_sextet restates it; tests/test_gpu_noise.py pins it pointwise.
// ------------------------------------------------------------------------------------------------
constexpr float kSextetMaxAbs = 6.24f;
// pair (z0, z1) of the sextet scheme from word w (radius + 5 angle bits) and the 10 angle bits f10s = (f10 << 13) & 0x7FE000 of word 3,
// both scaled by `scale` (folded into the radius)
__device__ __forceinline__ float2 box_muller27(uint32_t w, uint32_t f10s, float scale) {
  const float u = fmaf((float)(w >> 5), 7.450580596923828125e-9f, 3.7252902984619140625e-9f);  // (k + 1/2) 2^-27
  const float r = sqrt_approx(-1.3862943611198906f * lg2_ftz(u)) * scale;
  float s, c;
  __sincosf(6.28318530717958647692f * __uint_as_float(0x3f800000u | f10s | ((w << 8) & 0x1F00u)), &s, &c);  // angle 2 pi m, m in [1, 2)
  return make_float2(r * c, r * s);
}
__device__ __forceinline__ void normal6(uint4 r, float scale, float (&z)[6]) {
  const float2 a = box_muller27(r.x, (r.w << 13) & 0x7FE000u, scale);
  const float2 b = box_muller27(r.y, (r.w << 3) & 0x7FE000u, scale);
  const float2 c = box_muller27(r.z, (r.w >> 7) & 0x7FE000u, scale);
  z[0] = a.x; z[1] = a.y; z[2] = b.x; z[3] = b.y; z[4] = c.x; z[5] = c.y;
}
// generic (slow) access: the quad at position `pos` of a row under the sextet scheme -- fallback paths only
template <class P>
__device__ __forceinline__ float4 normal_sextet_quad(const P& ph, uint32_t stream, uint64_t offset, uint64_t sample, uint32_t pos, float scale) {
  const uint32_t gran = pos / 96u, rem = pos - gran * 96u, j = rem >> 5, g = gran * 32u + (rem & 31u);
  float n[12];
  {
    float z[6];
    normal6(philox_at(ph, stream, offset, sample, 2u * g), scale, z);
#pragma unroll
    for (int i = 0; i < 6; ++i) n[i] = z[i];
    normal6(philox_at(ph, stream, offset, sample, 2u * g + 1u), scale, z);
#pragma unroll
    for (int i = 0; i < 6; ++i) n[6 + i] = z[i];
  }
  return j == 0 ? make_float4(n[0], n[1], n[2], n[3]) : (j == 1 ? make_float4(n[4], n[5], n[6], n[7]) : make_float4(n[8], n[9], n[10], n[11]));
}

__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exp(1) variate from 32 bits, accurate in the W -> 0 tail (which drives the heavy tail of A):
// W = -log(1 - d), d in (0,1) on the centred 32-bit lattice; series for small d, MUFU lg2 otherwise.
__device__ __forceinline__ float exp1(uint32_t x) {
  const float d = fminf(fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f), 0.99999994f);
  const float series = d * (1.0f + d * (0.5f + d * (0.33333334f + d * 0.25f)));
  const float full = -0.6931471805599453f * lg2_ftz(1.0f - d);
  return d < 0.03125f ? series : full;
}

// SCALE * sin(pi v) for v in [0, 1/2] on the FMA pipes: odd minimax polynomial of degree 9 (relative error 5.3e-9 before
// rounding, so sin keeps RELATIVE accuracy as v -> 0, which MUFU.SIN does not promise).  Replaces MUFU.SIN + range scaling
// + a small-argument branch: fewer instructions and no XU-pipe slot.  (Throughput note, profiles/r01_ncu_stream.md: the
// per-element draw is bound by instruction dispatch with IMAD.WIDE costing 4 slots, not by the XU pipe.)
template <int SCALE>
__device__ __forceinline__ float sinpi_half(float v) {
  const float z = v * v;
  float p = SCALE * 0.07756038554456743f;
  p = fmaf(p, z, SCALE * -0.5982421256741446f);
  p = fmaf(p, z, SCALE * 2.5500697262138807f);
  p = fmaf(p, z, SCALE * -5.167709684792514f);
  p = fmaf(p, z, SCALE * 3.14159263689534f);
  return v * p;
}

// Parameters of the Kanter / CMS transform for alpha' = alpha/2 (precomputed on the host).
struct StableParams {
  float ap;        // alpha' = alpha / 2
  float inv_ap;    // 1 / alpha'
  float r;         // (1 - alpha') / alpha'
  float one_m_ap;  // 1 - alpha'
  int gaussian;    // alpha == 2  ->  A == 2 exactly (Distributions.py:40-42)
};

// A = 2 K,  K = sin(a'U)/sin(U)^(1/a') * (sin((1-a')U)/W)^((1-a')/a'),  U = pi u ~ Unif(0,pi), W ~ Exp(1)
// (identical pointwise to scipy's _rvs_Z1 'otherwise' branch with beta=1
//  times scale 2 cos(pi alpha/4)^(2/alpha), see oracle/stable.py).
// Evaluated in log2 space with the logarithms merged (1/a' = 1 + r):
//   lg2 A = lg2( 2 sin(a'U) / sin U ) + r lg2( sin((1-a')U) / (sin U * W) )
// = one MUFU.RCP + two MUFU.LG2 instead of four LG2; the heavy tail (U -> pi, W -> 0) neither overflows (sin U * W >= 2e-17)
// nor loses precision: sin(U) is evaluated on the reflected argument min(u, 1-u) built from the integer so that U -> pi keeps
// full relative precision.  MUFU ops per draw: LG2 (W), RCP, 2 LG2, EX2 = 5 (was 9: + 3 SIN, + 1 LG2).
// Checked pointwise against oracle/stable.py::kanter_A on the same lattice variates (max relative error 3e-6 incl. both tails).
__device__ __forceinline__ float stable_A(const StableParams& p, uint32_t xu, uint32_t xw) {
  if (p.gaussian) return 2.0f;
  // u in (0,1) on the centred 32-bit lattice; v = min(u, 1-u) is built from the integer so both ends are exact.
  const uint32_t xr = xu ^ (uint32_t)((int32_t)xu >> 31);                      // reflect the upper half (~xu)
  const float v = fmaf((float)xr, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 0.5]
  const float u = (int32_t)xu < 0 ? 1.0f - v : v;
  const float sinU = sinpi_half<1>(v);  // sin(pi u) = sin(pi (1-u))
  const float a1 = p.ap * u;            // (0, a') with a' < 1
  const float s1 = sinpi_half<2>(fminf(a1, 1.0f - a1));
  float a2 = p.one_m_ap * u;            // (0, 1/2] for alpha >= 1
  if (p.one_m_ap > 0.5f) a2 = fminf(a2, 1.0f - a2);
  const float s2 = sinpi_half<1>(a2);
  const float w = exp1(xw);
  const float q = rcp_ftz(sinU * w);
  const float l2 = lg2_ftz((s1 * q) * w) + p.r * lg2_ftz(s2 * q);
  return ex2_ftz(l2);
}

}  // namespace dlpm
