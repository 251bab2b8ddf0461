// K6 / K7 and the small memory-bound UNet operators (NHWC bf16 activations, fp32 statistics).
//   GroupNorm + scale-shift + SiLU   unet.py:141-142,153-154,188-191,212,433-434 (GroupNorm32 nn.py:17-19)
//   QKV attention                    unet.py:231-250
//   input conv 3->C                  unet.py:347          nearest upsample x2   unet.py:73
//   timestep embedding + emb_layers  nn.py:103-121, unet.py:335-339,145-151,477
#include <cooperative_groups.h>

#include "../../include/dlpm_b200_unet.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dlpm {

// x * sigmoid(x) = h + h * tanh(h) with h = x / 2: ONE MUFU op (tanh.approx.f32, abs. error ~5e-4, far below the bf16
// rounding of the result) instead of EX2 + RCP -- the streaming GroupNorm kernels were MUFU-bound (16 ops/clk/SM)
__device__ __forceinline__ float tanh_fast(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return t;
}
__device__ __forceinline__ float silu_half(float h) { return fmaf(h, tanh_fast(h), h); }  // silu(2h)
__device__ __forceinline__ float silu_f(float v) { return silu_half(0.5f * v); }

__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(p[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// K6 GroupNorm (+ scale-shift) (+ SiLU): one CTA per sample, two passes over the sample (2nd pass hits L2).
// Thread layout: lane v = 8-channel vector (C/8 of them), slot = pixel phase; per-channel partial sums in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 512;

__global__ void __launch_bounds__(kGnThreads) k_groupnorm(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ in0,
                                                          int C0, const __nv_bfloat16* __restrict__ in1, int C1, int HW,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ ss, int ss_rows, int64_t ss_stride,
                                                          int64_t ss_off, int apply_silu) {
  extern __shared__ float sm[];
  pdl_launch_dependents();
  pdl_wait();
  const int C = C0 + C1, nvec = C >> 3;
  const int slots = kGnThreads / nvec;
  const int n = blockIdx.x;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec;
  const bool active = slot < slots;
  const int c = v * 8;
  const __nv_bfloat16* src = (c < C0) ? in0 + (int64_t)n * HW * C0 + c : in1 + (int64_t)n * HW * C1 + (c - C0);
  const int src_stride = (c < C0) ? C0 : C1;
  float* part_sum = sm;                       // [slots][C]
  float* part_sq = sm + slots * C;            // [slots][C]
  float* coef_a = sm + 2 * slots * C;         // [C]
  float* coef_b = coef_a + C;                 // [C]
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active) {
    for (int p = slot; p < HW; p += slots) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (int64_t)p * src_stride)), x);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += x[e]; q[e] = fmaf(x[e], x[e], q[e]); }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { part_sum[slot * C + c + e] = s[e]; part_sq[slot * C + c + e] = q[e]; }
  }
  __syncthreads();
  // per-channel totals, then per-group statistics
  for (int ch = threadIdx.x; ch < C; ch += kGnThreads) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < slots; ++k) { a += part_sum[k * C + ch]; b += part_sq[k * C + ch]; }
    coef_a[ch] = a; coef_b[ch] = b;
  }
  __syncthreads();
  const int G = C < 32 ? C : 32, cpg = C / G;
  float mean = 0.f, rstd = 0.f;
  const int my_ch = threadIdx.x;  // thread ch handles channel ch in the coefficient phase
  if (my_ch < C) {
    const int g = my_ch / cpg;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < cpg; ++k) { a += coef_a[g * cpg + k]; b += coef_b[g * cpg + k]; }
    const float inv_n = 1.0f / (float)(cpg * HW);
    mean = a * inv_n;
    const float var = fmaxf(b * inv_n - mean * mean, 0.f);
    rstd = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += kGnThreads) {  // C <= 512 == kGnThreads: one channel per thread
    float ga = __ldg(gamma + ch) * rstd;
    float be = __ldg(beta + ch) - mean * ga;
    if (ss) {
      const float* row = ss + (ss_rows == 1 ? 0 : (int64_t)n * ss_stride) + ss_off;
      const float sc = 1.0f + __ldg(row + ch), sh = __ldg(row + C + ch);
      ga *= sc;
      be = be * sc + sh;
    }
    coef_a[ch] = ga; coef_b[ch] = be;
  }
  __syncthreads();
  if (active) {
    float a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = coef_a[c + e]; b[e] = coef_b[c + e]; }
    __nv_bfloat16* dst = out + (int64_t)n * HW * C + c;
    for (int p = slot; p < HW; p += slots) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (int64_t)p * src_stride)), x);
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] = fmaf(x[e], a[e], b[e]); if (apply_silu) x[e] = silu_f(x[e]); }
      *reinterpret_cast<uint4*>(dst + (int64_t)p * C) = pack8(x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6 (production variant): ONE HBM read + ONE HBM write per GroupNorm.  A thread-block CLUSTER of cs CTAs owns one sample;
// every CTA keeps its HW/cs pixels (<= 96 KB of bf16) in shared memory, per-channel partial sums are exchanged
// through distributed shared memory (fixed rank order -> bitwise deterministic), then the data is normalised from
// shared memory.  No second pass over global memory, no atomics.
// ------------------------------------------------------------------------------------------------
// K6 (fused-statistics variant).  When the tensor was written by k_conv_tc, its epilogue already left per-image partial
// sums per channel QUAD ([B][parts][C/4][2] fp32: sum, sum of squares; conv_tc.cu).  GroupNorm then needs no statistics
// pass and no cluster: gn_fold turns the partials of ONE sample into per-channel affine coefficients
//   y = a[c] * x + b[c],  a = gamma*rstd*(1+scale),  b = (beta - mean*gamma*rstd)*(1+scale) + shift
// (fixed summation order -> bitwise deterministic), and k_gn_apply streams the tensor once (read + write).
// Group boundaries are multiples of 4 channels for every C that is a multiple of 128 (32 groups).
// ------------------------------------------------------------------------------------------------
struct GnFoldArgs {
  const float* st0; int parts0, C0;
  const float* st1; int parts1, C1;
  int HW;
  const float* gamma; const float* beta;
  const float* ss; int ss_rows; int64_t ss_stride, ss_off;
  float out_scale;  // 1, or 0.5 for consumers that evaluate silu(y) as h + h*tanh(h) with h = y/2
};

// block-wide; quad: shared scratch [C/4][2]; coef_a / coef_b: [C] (shared or global); ends with __syncthreads when SYNC
__device__ __forceinline__ void gn_fold(const GnFoldArgs& g, int n, float* quad, float* coef_a, float* coef_b, int a_stride) {
  const int C = g.C0 + g.C1, nq = C >> 2, nq0 = g.C0 >> 2;
  const int pl = threadIdx.x & 7;
  for (int qd = threadIdx.x >> 3; qd < nq; qd += blockDim.x >> 3) {
    const bool first = qd < nq0;
    const float* st = first ? g.st0 : g.st1;
    const int parts = first ? g.parts0 : g.parts1, cq = first ? nq0 : nq - nq0, ql = first ? qd : qd - nq0;
    const float2* row = reinterpret_cast<const float2*>(st) + (int64_t)n * parts * cq + ql;
    float s = 0.f, q = 0.f;
    for (int pp = pl; pp < parts; pp += 8) { const float2 v = __ldg(row + (int64_t)pp * cq); s += v.x; q += v.y; }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); q += __shfl_xor_sync(0xffffffffu, q, off); }
    if (pl == 0) { quad[2 * qd] = s; quad[2 * qd + 1] = q; }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int cpg = C >> 5, gq0 = ((ch / cpg) * cpg) >> 2;
    float sA = 0.f, qA = 0.f;
    for (int k = 0; k < (cpg >> 2); ++k) { sA += quad[2 * (gq0 + k)]; qA += quad[2 * (gq0 + k) + 1]; }
    const float inv_n = 1.0f / (float)(cpg * g.HW);
    const float mean = sA * inv_n;
    const float rstd = rsqrtf(fmaxf(qA * inv_n - mean * mean, 0.f) + 1e-5f);
    float ga = __ldg(g.gamma + ch) * rstd;
    float be = __ldg(g.beta + ch) - mean * ga;
    if (g.ss) {
      const float* row = g.ss + (g.ss_rows == 1 ? 0 : (int64_t)n * g.ss_stride) + g.ss_off;
      const float sc = 1.0f + __ldg(row + ch), sh = __ldg(row + C + ch);
      ga *= sc;
      be = be * sc + sh;
    }
    coef_a[ch * a_stride] = ga * g.out_scale;
    coef_b[ch * a_stride] = be * g.out_scale;
  }
}

constexpr int kGnApplyThreads = 512;

__global__ void __launch_bounds__(kGnApplyThreads) k_gn_apply(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ in0,
                                                              const __nv_bfloat16* __restrict__ in1, GnFoldArgs g, int apply_silu,
                                                              int pix_per_cta, int slices, int reverse) {
  __shared__ float quad[256];
  __shared__ float coef_a[512], coef_b[512];
  pdl_launch_dependents();
  pdl_wait();
  // reverse: last sample first -- the rows the producing convolution wrote last are the ones still in L2 (see ConvLaunch::reverse)
  const int bid = reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int n = bid / slices, p0 = (bid - n * slices) * pix_per_cta;
  const int C = g.C0 + g.C1, nvec = C >> 3, nvec0 = g.C0 >> 3;
  const int slots = kGnApplyThreads / nvec;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec;
  const int c = v * 8;
  const bool first = c < g.C0;
  const int sstride = first ? nvec0 : nvec - nvec0;
  const uint4* src = first ? reinterpret_cast<const uint4*>(in0) + ((int64_t)n * g.HW + p0) * nvec0 + v
                           : reinterpret_cast<const uint4*>(in1) + ((int64_t)n * g.HW + p0) * (nvec - nvec0) + (v - nvec0);
  uint4* dst = reinterpret_cast<uint4*>(out) + ((int64_t)n * g.HW + p0) * nvec + v;
  constexpr int U = 4;
  int p = slot;
  // the first batch of loads does not depend on the coefficients: issue it before the fold so HBM latency overlaps it
  uint4 raw[U];
  const bool pre = slot < slots && p + (U - 1) * slots < pix_per_cta;
  if (pre) {
#pragma unroll
    for (int u = 0; u < U; ++u) raw[u] = ld_stream_u4(src + (int64_t)(p + u * slots) * sstride);
  }
  gn_fold(g, n, quad, coef_a, coef_b, 1);
  __syncthreads();
  if (slot >= slots) return;
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = coef_a[c + e]; b[e] = coef_b[c + e]; }
  for (bool have = pre; p + (U - 1) * slots < pix_per_cta; p += U * slots, have = false) {
    if (!have) {
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ld_stream_u4(src + (int64_t)(p + u * slots) * sstride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float x[8];
      unpack8(raw[u], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] = fmaf(x[e], a[e], b[e]); if (apply_silu) x[e] = silu_half(x[e]); }  // coefficients pre-halved
      dst[(int64_t)(p + u * slots) * nvec] = pack8(x);
    }
  }
  for (; p < pix_per_cta; p += slots) {
    float x[8];
    unpack8(ld_stream_u4(src + (int64_t)p * sstride), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) { x[e] = fmaf(x[e], a[e], b[e]); if (apply_silu) x[e] = silu_half(x[e]); }
    dst[(int64_t)p * nvec] = pack8(x);
  }
}

// coefficient table only ([B][C] float2 = (a, b)), for convolutions that normalise their input on load
__global__ void __launch_bounds__(256) k_gn_fold(float2* __restrict__ ab, GnFoldArgs g) {
  __shared__ float quad[256];
  pdl_launch_dependents();
  pdl_wait();
  const int C = g.C0 + g.C1;
  float* base = reinterpret_cast<float*>(ab + (int64_t)blockIdx.x * C);
  gn_fold(g, blockIdx.x, quad, base, base + 1, 2);
}

// ------------------------------------------------------------------------------------------------
constexpr int kGnClusterSmemData = 96 * 1024;

__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kGnThreads) k_groupnorm_cluster(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ in0,
                                                                  int C0, const __nv_bfloat16* __restrict__ in1, int C1, int HW,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ ss, int ss_rows, int64_t ss_stride,
                                                                  int64_t ss_off, int apply_silu, int pix_per_cta) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int n = blockIdx.x / cs;
  const int C = C0 + C1, nvec = C >> 3, nvec0 = C0 >> 3;
  const int nthreads = blockDim.x;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  // two planes (one per input tensor) so that each is ONE contiguous bulk copy: plane0 [pix][C0], plane1 [pix][C1]
  __nv_bfloat16* plane0 = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* plane1 = plane0 + (size_t)pix_per_cta * C0;
  float* psum = reinterpret_cast<float*>(sm_raw + (size_t)pix_per_cta * C * 2);     // [C] this CTA's per-channel sums
  float* psq = psum + C;                                                            // [C]
  float* coef_a = psq + C;                                                          // [C]
  float* coef_b = coef_a + C;                                                       // [C]
  float* part = coef_b + C;                                                         // [parts][C/2][4] scratch, parts * C/2 <= 512
  uint64_t* bar = reinterpret_cast<uint64_t*>(part + 2048);
  const int p0 = rank * pix_per_cta;
  // phase 1: global -> shared with bulk asynchronous copies (the only read of the tensor)
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes0 = (uint32_t)pix_per_cta * C0 * 2, bytes1 = (uint32_t)pix_per_cta * C1 * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gn_smem_u32(bar)), "r"(bytes0 + bytes1) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gn_smem_u32(plane0)),
                 "l"(in0 + ((int64_t)n * HW + p0) * C0), "r"(bytes0), "r"(gn_smem_u32(bar))
                 : "memory");
    if (C1)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gn_smem_u32(plane1)),
                   "l"(in1 + ((int64_t)n * HW + p0) * C1), "r"(bytes1), "r"(gn_smem_u32(bar))
                   : "memory");
  }
  // affine parameters of (up to) two channels per thread are fetched while the bulk copy is in flight
  float pg[2], pb[2], psc[2], psh[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int ch = threadIdx.x + k * nthreads;
    pg[k] = pb[k] = psh[k] = 0.f;
    psc[k] = 1.f;
    if (ch < C) {
      pg[k] = __ldg(gamma + ch);
      pb[k] = __ldg(beta + ch);
      if (ss) {
        const float* row = ss + (ss_rows == 1 ? 0 : (int64_t)n * ss_stride) + ss_off;
        psc[k] = 1.0f + __ldg(row + ch);
        psh[k] = __ldg(row + C + ch);
      }
    }
  }
  __syncthreads();  // barrier initialised before anyone polls it
  {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(gn_smem_u32(bar)) : "memory");
      if (++spins > (1u << 26)) __trap();
    }
  }
  // phase 2: per-channel sums over this CTA's pixels (2-byte shared loads, consecutive threads = consecutive channels)
  // (channel PAIRS: one 4-byte shared load feeds two channels; consecutive threads = consecutive pairs, conflict-free)
  const int C2 = C >> 1;
  const int parts = nthreads / C2 > 0 ? nthreads / C2 : 1;  // pixel phases per channel pair; parts * C2 <= 512
  for (int item = threadIdx.x; item < parts * C2; item += nthreads) {
    const int cp = item % C2, pt = item / C2, ch = cp * 2;
    const __nv_bfloat162* col = reinterpret_cast<const __nv_bfloat162*>(ch < C0 ? plane0 + ch : plane1 + (ch - C0));
    const int stride2 = (ch < C0 ? C0 : C1) >> 1;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 4
    for (int p = pt; p < pix_per_cta; p += parts) {
      const float2 xv = __bfloat1622float2(col[p * stride2]);
      s0 += xv.x; q0 = fmaf(xv.x, xv.x, q0);
      s1 += xv.y; q1 = fmaf(xv.y, xv.y, q1);
    }
    float4* dstp = reinterpret_cast<float4*>(part) + (pt * C2 + cp);
    *dstp = make_float4(s0, q0, s1, q1);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += nthreads) {
    float sA = 0.f, qA = 0.f;
    for (int pt = 0; pt < parts; ++pt) {
      const float* e = part + ((pt * C2 + (ch >> 1)) * 4) + (ch & 1) * 2;
      sA += e[0];
      qA += e[1];
    }
    psum[ch] = sA;
    psq[ch] = qA;
  }
  if (cs > 1) cluster.sync(); else __syncthreads();
  // phase 3: totals over the cluster (DSMEM reads, fixed order), group statistics, affine coefficients
  for (int ch = threadIdx.x; ch < C; ch += nthreads) {
    float sA = 0.f, qA = 0.f;
    for (int r = 0; r < cs; ++r) {
      sA += cluster.map_shared_rank(psum, r)[ch];
      qA += cluster.map_shared_rank(psq, r)[ch];
    }
    part[ch * 2] = sA;
    part[ch * 2 + 1] = qA;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int ch = threadIdx.x + k * nthreads;
    if (ch < C) {
      const int G = C < 32 ? C : 32, cpg = C / G, g = ch / cpg;
      float sA = 0.f, qA = 0.f;
      for (int kk = 0; kk < cpg; ++kk) { sA += part[(g * cpg + kk) * 2]; qA += part[(g * cpg + kk) * 2 + 1]; }
      const float inv_n = 1.0f / (float)(cpg * HW);
      const float mean = sA * inv_n;
      const float rstd = rsqrtf(fmaxf(qA * inv_n - mean * mean, 0.f) + 1e-5f);
      const float ga = pg[k] * rstd;
      const float be = pb[k] - mean * ga;
      coef_a[ch] = ga * psc[k];
      coef_b[ch] = be * psc[k] + psh[k];
    }
  }
  if (cs > 1) cluster.sync(); else __syncthreads();  // peers done reading this CTA's shared memory; orders coef_*
  // phase 4: normalise from shared memory -> global (the only write)
  const int slots = nthreads / nvec;
  const int v = threadIdx.x % nvec, slot = threadIdx.x / nvec;
  if (slot < slots) {
    const int c = v * 8;
    float a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = coef_a[c + e]; b[e] = coef_b[c + e]; }
    const uint4* srcv = c < C0 ? reinterpret_cast<const uint4*>(plane0) + v : reinterpret_cast<const uint4*>(plane1) + (v - nvec0);
    const int sstride = c < C0 ? nvec0 : nvec - nvec0;
    __nv_bfloat16* dst = out + ((int64_t)n * HW + p0) * C + c;
    for (int p = slot; p < pix_per_cta; p += slots) {
      float x[8];
      unpack8(srcv[p * sstride], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] = fmaf(x[e], a[e], b[e]); if (apply_silu) x[e] = silu_f(x[e]); }
      *reinterpret_cast<uint4*>(dst + (int64_t)p * C) = pack8(x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K7 attention: one CTA per (sample, head); K and V of the head staged in shared memory as fp32;
// one query row per thread with an online softmax.  L <= 1024, d = C/heads <= 64.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_attention(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv, int L,
                                                   int C, int heads) {
  // A query row is shared by TPQ = D/8 adjacent lanes, each owning 8 of the D dimensions of q and of the output: dot
  // products are finished with TPQ-wide shuffles.  (One row per thread left 16 threads per CTA busy at L = 16, the
  // CIFAR middle block: 60 us for 134 MFLOP.)
  constexpr int TPQ = D / 8;
  extern __shared__ float sm[];
  pdl_launch_dependents();
  pdl_wait();
  float* Ks = sm;            // [L][D]
  float* Vs = sm + L * D;    // [L][D]
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const __nv_bfloat16* base = qkv + (int64_t)n * L * 3 * C + h * 3 * D;  // per head: q | k | v blocks of D channels
  for (int i = threadIdx.x; i < L * (D / 8); i += blockDim.x) {  // 16-byte loads: 8 channels of K and of V per iteration
    const int s = i / (D / 8), d8 = (i - s * (D / 8)) * 8;
    const uint4 kr = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)s * 3 * C + D + d8));
    const uint4 vr = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)s * 3 * C + 2 * D + d8));
    const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&kr);
    const __nv_bfloat162* vp = reinterpret_cast<const __nv_bfloat162*>(&vr);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 kf = __bfloat1622float2(kp[e]), vf = __bfloat1622float2(vp[e]);
      Ks[s * D + d8 + 2 * e] = kf.x; Ks[s * D + d8 + 2 * e + 1] = kf.y;
      Vs[s * D + d8 + 2 * e] = vf.x; Vs[s * D + d8 + 2 * e + 1] = vf.y;
    }
  }
  __syncthreads();
  const float scale2 = rsqrtf((float)D);  // (1/sqrt(sqrt(d)))^2 applied to q and k (unet.py:244-247)
  const int part = threadIdx.x % TPQ, d0 = part * 8;
  const int rows_per_pass = blockDim.x / TPQ;
  for (int t0 = 0; t0 < L; t0 += rows_per_pass) {  // uniform trip count: every lane takes part in the shuffles
    const int t = t0 + threadIdx.x / TPQ;
    const bool active = t < L;
    const int tr = active ? t : L - 1;
    float qv[8], o[8];
    {
      const uint4 qr = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)tr * 3 * C + d0));
      const __nv_bfloat162* qp = reinterpret_cast<const __nv_bfloat162*>(&qr);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(qp[e]);
        qv[2 * e] = f.x * scale2; qv[2 * e + 1] = f.y * scale2;
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = 0.f;
    float mx = -INFINITY, den = 0.f;
    for (int s = 0; s < L; ++s) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) dot = fmaf(qv[d], Ks[s * D + d0 + d], dot);
#pragma unroll
      for (int off = TPQ / 2; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
      const float nm = fmaxf(mx, dot);
      const float corr = __expf(mx - nm), pw = __expf(dot - nm);
      den = den * corr + pw;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = fmaf(o[d], corr, pw * Vs[s * D + d0 + d]);
      mx = nm;
    }
    if (active) {
      const float inv = 1.0f / den;
      uint4 ov;
      __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
      for (int e = 0; e < 4; ++e) op[e] = __floats2bfloat162_rn(o[2 * e] * inv, o[2 * e + 1] * inv);
      *reinterpret_cast<uint4*>(out + ((int64_t)n * L + t) * C + h * D + d0) = ov;
    }
  }
}

// Tensor-core form of K7 for head dims 16 / 32 / 64 and L a multiple of 16 (every attention block of the shipped configs:
// MNIST L = 256 / 64 with d = 16 -- 58 % of that network's forward with the FMA kernel above -- and CIFAR L = 16, d = 64).
// Flash-attention dataflow on mma.sync.m16n8k16 (bf16 in, fp32 accumulate): a warp owns 16 query rows, walks the keys in
// blocks of 64 with an online softmax in the exp2 domain, and feeds the S accumulator fragments straight back as the A
// operand of P.V (the m16n8 accumulator layout of two adjacent key tiles IS the m16k16 A layout).  K sits in shared memory
// row-major with (D+8)-element rows, V transposed with (L+8)-element rows: both B-fragment reads are conflict-free 32-bit
// loads.  The contraction dims here (d = 16..64, 16..256 keys) are far below a tcgen05 tile (M = 128 per CTA, operands via
// TMA descriptors); the warp-level MMA is the right granularity for this 0.3 % of the FLOPs.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes: round-to-nearest split x = n + f (magic-number add), cubic minimax of 2^f on
// [-1/2, 1/2] (relative error 7.5e-5, far below the 2^-9 rounding of the bf16 probability it feeds), exponent added to the bits.
// x < -125 (incl. -inf of masked keys) returns 0.
__device__ __forceinline__ float att_ex2_poly(float x) {
  const float xc = fmaxf(x, -125.0f);
  const float tt = xc + 12582912.0f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = xc - (tt - 12582912.0f);    // [-1/2, 1/2]
  float p = fmaf(f, 0.05517163f, 0.24261112f);
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  const float r = __int_as_float(__float_as_int(p) + (__float_as_int(tt) << 23));
  return x < -125.0f ? 0.0f : r;
}

template <int D, bool POLY>
__global__ void __launch_bounds__(256) k_attention_mma(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv, int L,
                                                       int C, int heads, int q_blocks) {
  constexpr int KSTR = D + 8;
  extern __shared__ __align__(16) uint8_t att_smem[];
  pdl_launch_dependents();
  pdl_wait();
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(att_smem);  // [L][KSTR]
  __nv_bfloat16* Vs = Ks + (size_t)L * KSTR;                        // [L][KSTR], row-major like K: the P.V B fragments come from ldmatrix.trans
  const int qb = blockIdx.x % q_blocks, nh = blockIdx.x / q_blocks;
  const int n = nh / heads, h = nh % heads;
  const __nv_bfloat16* base = qkv + (int64_t)n * L * 3 * C + h * 3 * D;  // per head: q | k | v blocks of D channels
  const int64_t rs = (int64_t)3 * C;                                      // row stride of qkv in elements
  for (int i = threadIdx.x; i < L * (D / 8); i += blockDim.x) {
    const int s = i / (D / 8), d8 = (i - s * (D / 8)) * 8;
    const uint4 kr = __ldg(reinterpret_cast<const uint4*>(base + s * rs + D + d8));
    const uint4 vr = __ldg(reinterpret_cast<const uint4*>(base + s * rs + 2 * D + d8));
    *reinterpret_cast<uint4*>(Ks + s * KSTR + d8) = kr;
    *reinterpret_cast<uint4*>(Vs + s * KSTR + d8) = vr;
    // the eight padding columns of a V row hold (1, 0, ..., 0): a ninth channel block whose P.V product is the ROW SUM of the
    // probabilities (of the bf16 values the numerator uses) -- four MMAs per 64 keys instead of 32 FADDs per thread
    if (d8 == 0) *reinterpret_cast<uint4*>(Vs + s * KSTR + D) = make_uint4(0x00003f80u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = qb * (int)(blockDim.x >> 5) * 16 + warp * 16;  // a CTA of w warps owns 16 w query rows of one head
  if (r0 >= L) return;
  uint32_t qa[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    const __nv_bfloat16* q0 = base + (r0 + g) * rs + kk * 16 + 2 * t;
    qa[kk][0] = __ldg(reinterpret_cast<const uint32_t*>(q0));
    qa[kk][1] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 8 * rs));
    qa[kk][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 8));
    qa[kk][3] = __ldg(reinterpret_cast<const uint32_t*>(q0 + 8 * rs + 8));
  }
  // softmax(q.k / sqrt(d)) (unet.py:244-247: q and k are each scaled by d^-1/4), evaluated as exp2(s * c - m * c) with
  // c = log2 e / sqrt(d): the running maximum is kept on the RAW scores (c > 0), so an element costs one FFMA + one MUFU.EX2
  // + one FADD (row sum) -- at d = 16 the softmax, not the MMAs, is the bulk of this kernel (L^2 exponentials per head against
  // 4 d L^2 tensor FLOPs).  POLY evaluates a quarter of the exponentials on the FMA pipes (att_ex2_poly); measured slower
  // (the kernel is issue-bound, the XU pipe is not the limiter), so it is off unless "attention_poly" = 1.
  const float sl2 = rsqrtf((float)D) * 1.4426950408889634f;
  float o[D / 8][4];
#pragma unroll
  for (int nd = 0; nd < D / 8; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY;
  float osum[4] = {0.f, 0.f, 0.f, 0.f};  // P.V against the ones column: [0] / [2] of the lanes with t == 0 are the row sums of rows g / g + 8
  for (int kb = 0; kb < L; kb += 64) {
    const int nkt = (L - kb) >= 64 ? 8 : (L - kb) / 8;  // valid 8-key tiles of this block (L is a multiple of 16)
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {  // (nkt is even: L is a multiple of 16)
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      sc[j + 1][0] = sc[j + 1][1] = sc[j + 1][2] = sc[j + 1][3] = 0.f;
      if (j < nkt) {
        // B fragments of K^T for two 8-key tiles per ldmatrix.x4: matrix l >> 3 = (key tile j + (l >> 4), channel half (l >> 3) & 1)
        const uint32_t ka = (uint32_t)__cvta_generic_to_shared(Ks + (kb + (j + (lane >> 4)) * 8 + (lane & 7)) * KSTR + ((lane >> 3) & 1) * 8);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t b00, b01, b10, b11;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(b00), "=r"(b01), "=r"(b10), "=r"(b11)
                       : "r"(ka + (uint32_t)(kk * 32)));
          mma_bf16_16816(sc[j], qa[kk], b00, b01);
          mma_bf16_16816(sc[j + 1], qa[kk], b10, b11);
        }
      } else {
        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY;
        sc[j + 1][0] = sc[j + 1][1] = sc[j + 1][2] = sc[j + 1][3] = -INFINITY;
      }
    }
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx_lo = fmaxf(mx_lo, fmaxf(sc[j][0], sc[j][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(sc[j][2], sc[j][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float c_lo = att_ex2((m_lo - mx_lo) * sl2), c_hi = att_ex2((m_hi - mx_hi) * sl2);  // first block: ex2(-inf) = 0
    m_lo = mx_lo; m_hi = mx_hi;
    const float ms_lo = -m_lo * sl2, ms_hi = -m_hi * sl2;
    osum[0] *= c_lo; osum[2] *= c_hi;
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) { o[nd][0] *= c_lo; o[nd][1] *= c_lo; o[nd][2] *= c_hi; o[nd][3] *= c_hi; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = att_ex2(fmaf(sc[j][0], sl2, ms_lo)); sc[j][1] = att_ex2(fmaf(sc[j][1], sl2, ms_lo));
      sc[j][2] = att_ex2(fmaf(sc[j][2], sl2, ms_hi));
      sc[j][3] = POLY ? att_ex2_poly(fmaf(sc[j][3], sl2, ms_hi)) : att_ex2(fmaf(sc[j][3], sl2, ms_hi));
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (2 * ks < nkt) {
        uint32_t pa[4];
        pa[0] = pack_bf16(sc[2 * ks][0], sc[2 * ks][1]);
        pa[1] = pack_bf16(sc[2 * ks][2], sc[2 * ks][3]);
        pa[2] = pack_bf16(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
        pa[3] = pack_bf16(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
        // B fragments of V (16 keys x 8 channels, "col" operand) straight from the row-major tile: ldmatrix.x4.trans fetches the two
        // 8-key halves of two adjacent 8-channel blocks (lane l supplies row l & 7 of matrix l >> 3; the 144-byte row pitch keeps the
        // eight rows of a matrix on distinct banks), thread (g, t) receives V[2t .. 2t+1][g] of each
        const uint32_t va = (uint32_t)__cvta_generic_to_shared(Vs + (kb + ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * KSTR + (lane >> 4) * 8);
#pragma unroll
        for (int nd = 0; nd < D / 8; nd += 2) {
          uint32_t b00, b01, b10, b11;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(b00), "=r"(b01), "=r"(b10), "=r"(b11)
                       : "r"(va + (uint32_t)(nd * 16)));
          mma_bf16_16816(o[nd], pa, b00, b01);
          mma_bf16_16816(o[nd + 1], pa, b10, b11);
        }
        {  // the ones column (lanes 0-15 supply the sixteen row addresses of the two 8-key halves)
          uint32_t b0, b1;
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(va + (uint32_t)(D * 2) - (uint32_t)((lane >> 4) * 16)));
          mma_bf16_16816(osum, pa, b0, b1);
        }
      }
    }
  }
  const float l_lo = __shfl_sync(0xffffffffu, osum[0], lane & ~3), l_hi = __shfl_sync(0xffffffffu, osum[2], lane & ~3);
  const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
  __nv_bfloat16* d_lo = out + ((int64_t)n * L + r0 + g) * C + h * D + 2 * t;
  __nv_bfloat16* d_hi = d_lo + (int64_t)8 * C;
#pragma unroll
  for (int nd = 0; nd < D / 8; ++nd) {
    *reinterpret_cast<uint32_t*>(d_lo + nd * 8) = pack_bf16(o[nd][0] * i_lo, o[nd][1] * i_lo);
    *reinterpret_cast<uint32_t*>(d_hi + nd * 8) = pack_bf16(o[nd][2] * i_hi, o[nd][3] * i_hi);
  }
}

// ------------------------------------------------------------------------------------------------
// Network input for the tensor-core input conv: NCHW fp32 [B, C, H, W] -> NHWC bf16 [B, H, W, 32] holding the bf16
// SPLIT of every value, x = hi + lo (+ O(2^-17 |x|)):  channels [0,C) = hi, [C,2C) = lo, [2C,3C) = hi, rest 0.
// With weights packed as (w_hi, w_hi, w_lo) the 3x3 conv over these 32 channels computes
// x_hi*w_hi + x_lo*w_hi + x_hi*w_lo = x*w up to 2^-16 relative -- fp32-grade accuracy for the heavy-tailed x_t on
// the bf16 tensor cores (the fp32 FMA kernel k_conv_in took 2.8 % of the forward).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_input(__nv_bfloat16* __restrict__ out, const float* __restrict__ x, int C, int64_t HW,
                                                     int64_t total_px) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < total_px; px += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = px / HW, sp = px - n * HW;
    __align__(16) __nv_bfloat16 row[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) row[j] = __float2bfloat16(0.f);
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(x + (n * C + c) * HW + sp);
      const __nv_bfloat16 hi = __float2bfloat16(v);
      const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
      row[c] = hi; row[C + c] = lo; row[2 * C + c] = hi;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + px * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = reinterpret_cast<const uint4*>(row)[j];
  }
}

// ------------------------------------------------------------------------------------------------
// input conv: NCHW fp32 -> NHWC bf16, 3x3 pad 1, C_in <= 4 (fp32 FMA: the network input keeps full precision).
// One CTA per (image, band of kRows output rows).  Weights arrive pre-transposed [C_in*9][C_out] (in-major) and are
// copied to shared memory once per CTA; thread = (pixel x, g) owns the 16 channels {32*j + 4*g .. +3, j = 0..3} so the
// four 16-byte weight reads of a warp are contiguous (conflict-free) and broadcast across pixels.
// ------------------------------------------------------------------------------------------------
constexpr int kConvInRows = 4;

__global__ void __launch_bounds__(256) k_conv_in(__nv_bfloat16* __restrict__ out, const float* __restrict__ x,
                                                 const float* __restrict__ wT, const float* __restrict__ bias, int C_in, int C_out,
                                                 int H, int W, float* __restrict__ stats) {
  extern __shared__ float sm[];
  __shared__ float s_red[8][32][2];  // per warp, per channel quad: (sum, sum of squares) -- GroupNorm statistics of the output
  const int K = C_in * 9;
  float* s_w = sm;                    // [K][C_out]
  float* s_b = s_w + K * C_out;       // [C_out]
  float* s_in = s_b + C_out;          // [C_in][kRows + 2][W + 2]
  const int bands = (H + kConvInRows - 1) / kConvInRows;
  const int n = blockIdx.x / bands, y0 = (blockIdx.x % bands) * kConvInRows;
  for (int i = threadIdx.x; i < K * C_out / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wT) + i);
  for (int i = threadIdx.x; i < C_out; i += blockDim.x) s_b[i] = __ldg(bias + i);
  pdl_launch_dependents();
  pdl_wait();  // weights are constants (staged above, overlapping the previous kernel); x is produced upstream
  const int Wp = W + 2, R = kConvInRows + 2;
  for (int i = threadIdx.x; i < C_in * R * Wp; i += blockDim.x) {
    const int ci = i / (R * Wp), r = (i / Wp) % R, xx = i % Wp - 1;
    const int yy = y0 + r - 1;
    s_in[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((int64_t)n * C_in + ci) * H + yy) * W + xx) : 0.f;
  }
  __syncthreads();
  // thread = (pixel column px, 4-channel slice gq, block of 4 channel sets) and computes ALL kRows rows of the band:
  // every 16-byte weight read (4 shared-memory wavefronts) feeds 16 FMAs, so the kernel is FMA- not LDS-bound
  const int sets = C_out >> 5;            // 32-channel sets; a thread owns 4 channels of up to 4 sets
  const int gq = threadIdx.x & 7;         // which 4-channel slice inside each 32-channel set
  const int nsb = (sets + 3) / 4;
  for (int item = threadIdx.x >> 3; item < W * nsb; item += blockDim.x >> 3) {
    const int sb = item / W, px = item - sb * W;
    float acc[kConvInRows][4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (sb * 4 + j) * 32 + gq * 4;
      const float4 bv = (sb * 4 + j) < sets ? *reinterpret_cast<const float4*>(s_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kConvInRows; ++r) { acc[r][j][0] = bv.x; acc[r][j][1] = bv.y; acc[r][j][2] = bv.z; acc[r][j][3] = bv.w; }
    }
    for (int ci = 0; ci < C_in; ++ci) {
      float win[kConvInRows + 2][3];
#pragma unroll
      for (int r = 0; r < kConvInRows + 2; ++r)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) win[r][dx] = s_in[(ci * R + r) * Wp + px + dx];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float* wr = s_w + (ci * 9 + t) * C_out + gq * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((sb * 4 + j) < sets) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + (sb * 4 + j) * 32);
#pragma unroll
            for (int r = 0; r < kConvInRows; ++r) {
              const float v = win[r + t / 3][t % 3];
              acc[r][j][0] = fmaf(v, wv.x, acc[r][j][0]); acc[r][j][1] = fmaf(v, wv.y, acc[r][j][1]);
              acc[r][j][2] = fmaf(v, wv.z, acc[r][j][2]); acc[r][j][3] = fmaf(v, wv.w, acc[r][j][3]);
            }
          }
        }
      }
    }
    if (stats) {  // launch guarantees a single pass of this loop (W * nsb <= 32): quad index = j * 8 + gq
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sA = 0.f, qA = 0.f;
#pragma unroll
        for (int r = 0; r < kConvInRows; ++r) {
          if (y0 + r < H) {
            sA += (acc[r][j][0] + acc[r][j][1]) + (acc[r][j][2] + acc[r][j][3]);
            qA += fmaf(acc[r][j][0], acc[r][j][0], acc[r][j][1] * acc[r][j][1]) + fmaf(acc[r][j][2], acc[r][j][2], acc[r][j][3] * acc[r][j][3]);
          }
        }
        sA += __shfl_xor_sync(0xffffffffu, sA, 8); qA += __shfl_xor_sync(0xffffffffu, qA, 8);
        sA += __shfl_xor_sync(0xffffffffu, sA, 16); qA += __shfl_xor_sync(0xffffffffu, qA, 16);
        if ((threadIdx.x & 31) < 8) { s_red[threadIdx.x >> 5][j * 8 + gq][0] = sA; s_red[threadIdx.x >> 5][j * 8 + gq][1] = qA; }
      }
    }
#pragma unroll
    for (int r = 0; r < kConvInRows; ++r) {
      if (y0 + r >= H) continue;
      __nv_bfloat16* dst = out + (((int64_t)n * H + y0 + r) * W + px) * C_out + gq * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((sb * 4 + j) < sets) {
          uint2 o;
          *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(acc[r][j][0], acc[r][j][1]);
          *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(acc[r][j][2], acc[r][j][3]);
          *reinterpret_cast<uint2*>(dst + (sb * 4 + j) * 32) = o;
        }
      }
    }
  }
  if (stats) {
    __syncthreads();
    const int nq = C_out >> 2, warps_used = (W * 8) >> 5;
    if ((int)threadIdx.x < 2 * nq) {
      const int qd = threadIdx.x >> 1, which = threadIdx.x & 1;
      float a = 0.f;
      for (int w = 0; w < warps_used; ++w) a += s_red[w][qd][which];
      stats[(((int64_t)n * bands + (blockIdx.x % bands)) * nq + qd) * 2 + which] = a;
    }
  }
}

__global__ void __launch_bounds__(256) k_upsample2x(uint4* __restrict__ out, const uint4* __restrict__ in, int64_t B, int H, int W,
                                                    int C8) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = B * (2 * H) * (2 * W) * C8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    const int64_t pix = idx / C8;
    const int xo = (int)(pix % (2 * W)), yo = (int)((pix / (2 * W)) % (2 * H));
    const int64_t n = pix / ((int64_t)4 * W * H);
    out[idx] = __ldg(in + ((n * H + (yo >> 1)) * W + (xo >> 1)) * C8 + c);
  }
}

// ------------------------------------------------------------------------------------------------
// timestep embedding + emb_layers: three launches of one K-split GEMV kernel.
//   out[r][j] = act( bias[j] + sum_k in[r][k] * WT[k][j] ),  CTA = 32 outputs x 8 K-slices (one warp per slice,
//   lanes = consecutive outputs -> coalesced in-major weight reads), reduced through shared memory.
//   input mode 1 builds the sinusoidal embedding of t (nn.py:103-121) on the fly.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gemv_rows(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ t,
                                                   const int* __restrict__ t_dev, float inv_T, int sinus, int K, int64_t N,
                                                   const float* __restrict__ WT, const float* __restrict__ bias, int act_out) {
  extern __shared__ float sm[];  // in row [K], partial [8][32]
  float* s_in = sm;
  float* s_part = sm + K;
  const int r = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  if (sinus) {
    const float tv = t_dev ? (t ? t[*t_dev] : (float)(*t_dev) * inv_T) : t[r];
    const int half = K / 2;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      if (i < 2 * half) {
        const int k = i < half ? i : i - half;
        const float arg = tv * expf(-9.210340371976184f * (float)k / (float)half);  // exp(-ln(10000) k / half)
        s_in[i] = i < half ? cosf(arg) : sinf(arg);
      } else {
        s_in[i] = 0.f;  // odd-dim padding (nn.py:118-119)
      }
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_in[i] = in[(int64_t)r * K + i];
  }
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * 32 + lane;
  const int k0 = (K * warp) / 8, k1 = (K * (warp + 1)) / 8;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < N) {
    int k = k0;
    for (; k + 3 < k1; k += 4) {
      a0 = fmaf(s_in[k], __ldg(WT + (int64_t)k * N + j), a0);
      a1 = fmaf(s_in[k + 1], __ldg(WT + (int64_t)(k + 1) * N + j), a1);
      a2 = fmaf(s_in[k + 2], __ldg(WT + (int64_t)(k + 2) * N + j), a2);
      a3 = fmaf(s_in[k + 3], __ldg(WT + (int64_t)(k + 3) * N + j), a3);
    }
    for (; k < k1; ++k) a0 = fmaf(s_in[k], __ldg(WT + (int64_t)k * N + j), a0);
  }
  s_part[warp * 32 + lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (warp == 0 && j < N) {
    float a = __ldg(bias + j);
#pragma unroll
    for (int wI = 0; wI < 8; ++wI) a += s_part[wI * 32 + lane];
    if (act_out) a = a / (1.0f + expf(-a));
    out[(int64_t)r * N + j] = a;
  }
}

}  // namespace dlpm

using namespace dlpm;

int dlpm_b200_groupnorm_silu(void* out, const void* in0, int C0, const void* in1, int C1, int64_t B, int HW, const float* gamma,
                             const float* beta, const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu,
                             void* stream) {
  DLPM_REQUIRE(out && in0 && gamma && beta, "groupnorm: NULL tensor");
  DLPM_REQUIRE((in1 == nullptr) == (C1 == 0), "groupnorm: in1 / C1 mismatch");
  const int C = C0 + C1;
  DLPM_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && C >= 8 && C <= 512, "groupnorm: channels must be multiples of 8, total <= 512");
  DLPM_REQUIRE(C % (C < 32 ? C : 32) == 0, "groupnorm: channels must be divisible by the group count");
  DLPM_REQUIRE(B >= 0 && HW >= 1 && B < (1ll << 31), "groupnorm: bad sizes");
  DLPM_REQUIRE(!ss || ss_rows == 1 || ss_rows == B, "groupnorm: ss_rows must be 1 or B");
  if (B == 0) return DLPM_OK;
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  auto* i0 = reinterpret_cast<const __nv_bfloat16*>(in0);
  auto* i1 = reinterpret_cast<const __nv_bfloat16*>(in1);
  // cluster variant: smallest power-of-two cluster (<= 8) whose per-CTA slice fits the 96 KB shared-memory budget
  int cs = 1;
  const int64_t bytes = (int64_t)HW * C * 2;
  while (cs < 8 && bytes / cs > kGnClusterSmemData) cs *= 2;
  if (bytes / cs <= 112 * 1024 /* one CTA per SM in the worst case */ && HW % cs == 0 && B * cs < (1ll << 31)) {
    const int pix = HW / cs;
    const size_t smem = (size_t)pix * C * 2 + (size_t)(4 * C + 4 * 512) * sizeof(float) + 16;
    // small slices: 256-thread CTAs so that more of them are co-resident (the kernel is then pure latency)
    int threads = ((int64_t)pix * (C / 8) <= 2048) ? 256 : kGnThreads;
    if (threads < C / 8) threads = kGnThreads;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(k_groupnorm_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "groupnorm smem attribute");
      attr = true;
    }
    cudaError_t e = launch_ex(k_groupnorm_cluster, dim3((unsigned)(B * cs)), dim3((unsigned)threads), smem, (cudaStream_t)stream, cs, o, i0,
                              C0, i1, C1, HW, gamma, beta, ss, ss_rows, ss_stride, ss_off, apply_silu, pix);
    if (e != cudaSuccess) return cuda_fail(e, "groupnorm cluster launch");
    return DLPM_OK;
  }
  // fallback for shapes the cluster variant cannot hold: two passes over global memory
  const int nvec = C / 8, slots = kGnThreads / nvec;
  const size_t smem = (size_t)(2 * slots * C + 2 * C) * sizeof(float);
  static bool attr2 = false;
  if (!attr2) {
    cudaError_t e = cudaFuncSetAttribute(k_groupnorm, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "groupnorm smem attribute");
    attr2 = true;
  }
  cudaError_t e2 = launch_ex(k_groupnorm, dim3((unsigned)B), dim3(kGnThreads), smem, (cudaStream_t)stream, 1, o, i0, C0, i1, C1, HW, gamma,
                             beta, ss, ss_rows, ss_stride, ss_off, apply_silu);
  if (e2 != cudaSuccess) return cuda_fail(e2, "groupnorm launch");
  return DLPM_OK;
}

static int g_gn_apply_min_elems = 131072;
namespace dlpm { void gn_apply_set_min_elems(int v) { g_gn_apply_min_elems = v > 0 ? v : 131072; } }

namespace dlpm {
int groupnorm_from_stats_dir(void* out, const void* in0, int C0, const float* stats0, int parts0, const void* in1, int C1,
                             const float* stats1, int parts1, int64_t B, int HW, const float* gamma, const float* beta,
                             const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, int reverse, void* stream);
}
int dlpm_b200_groupnorm_from_stats(void* out, const void* in0, int C0, const float* stats0, int parts0, const void* in1, int C1,
                                   const float* stats1, int parts1, int64_t B, int HW, const float* gamma, const float* beta,
                                   const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, void* stream) {
  return groupnorm_from_stats_dir(out, in0, C0, stats0, parts0, in1, C1, stats1, parts1, B, HW, gamma, beta, ss, ss_rows, ss_stride, ss_off,
                                  apply_silu, 0, stream);
}

int dlpm::groupnorm_from_stats_dir(void* out, const void* in0, int C0, const float* stats0, int parts0, const void* in1, int C1,
                                   const float* stats1, int parts1, int64_t B, int HW, const float* gamma, const float* beta,
                                   const float* ss, int ss_rows, int64_t ss_stride, int64_t ss_off, int apply_silu, int reverse, void* stream) {
  DLPM_REQUIRE(out && in0 && stats0 && gamma && beta && parts0 >= 1, "groupnorm_from_stats: NULL tensor");
  DLPM_REQUIRE((in1 == nullptr) == (C1 == 0) && (in1 == nullptr) == (stats1 == nullptr), "groupnorm_from_stats: in1 / C1 / stats1 mismatch");
  const int C = C0 + C1;
  DLPM_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && C % 128 == 0 && C <= 512, "groupnorm_from_stats: total channels must be a multiple of 128 (<= 512)");
  DLPM_REQUIRE(B >= 0 && HW >= 1 && B < (1ll << 24), "groupnorm_from_stats: bad sizes");
  DLPM_REQUIRE(!ss || ss_rows == 1 || ss_rows == B, "groupnorm_from_stats: ss_rows must be 1 or B");
  if (B == 0) return DLPM_OK;
  // a CTA streams HW / slices pixels of one image after folding the image's statistics into per-channel coefficients: the
  // fold costs O(C) per CTA, so a slice keeps at least g_gn_apply_min_elems elements ("gn_apply_min_elems")
  // (measured at B = 512, GroupNorm total of one forward: 16 K elements 1.81 ms, 32 K 1.48, 64 K 1.41, 128 K 1.40, 512 K 1.38);
  // small batches are cut further so that the grid still covers the SMs twice
  int slices = 1;
  while (slices * 2 <= 64 && HW % (slices * 2) == 0 && (HW / (slices * 2)) >= 32 &&
         ((int64_t)(HW / (slices * 2)) * C >= g_gn_apply_min_elems ||
          (B * slices < 2 * kNumSMs && (int64_t)(HW / (slices * 2)) * C >= 16384)))
    slices *= 2;
  GnFoldArgs g{stats0, parts0, C0, stats1, parts1, C1, HW, gamma, beta, ss, ss_rows, ss_stride, ss_off, apply_silu ? 0.5f : 1.0f};
  cudaError_t e = launch_ex(k_gn_apply, dim3((unsigned)(B * slices)), dim3(kGnApplyThreads), 0, (cudaStream_t)stream, 1,
                            reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(in0),
                            reinterpret_cast<const __nv_bfloat16*>(in1), g, apply_silu, HW / slices, slices, reverse);
  if (e != cudaSuccess) return cuda_fail(e, "groupnorm_from_stats launch");
  return DLPM_OK;
}

int dlpm_b200_groupnorm_fold(float* ab, int C0, const float* stats0, int parts0, int C1, const float* stats1, int parts1, int64_t B,
                             int HW, const float* gamma, const float* beta, const float* ss, int ss_rows, int64_t ss_stride,
                             int64_t ss_off, int half, void* stream) {
  DLPM_REQUIRE(ab && stats0 && gamma && beta && parts0 >= 1, "groupnorm_fold: NULL tensor");
  DLPM_REQUIRE((C1 == 0) == (stats1 == nullptr), "groupnorm_fold: C1 / stats1 mismatch");
  const int C = C0 + C1;
  DLPM_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && C % 128 == 0 && C <= 512, "groupnorm_fold: total channels must be a multiple of 128 (<= 512)");
  DLPM_REQUIRE(B >= 0 && HW >= 1 && B < (1ll << 31), "groupnorm_fold: bad sizes");
  if (B == 0) return DLPM_OK;
  GnFoldArgs g{stats0, parts0, C0, stats1, parts1, C1, HW, gamma, beta, ss, ss_rows, ss_stride, ss_off, half ? 0.5f : 1.0f};
  cudaError_t e = launch_ex(k_gn_fold, dim3((unsigned)B), dim3(256), 0, (cudaStream_t)stream, 1, reinterpret_cast<float2*>(ab), g);
  if (e != cudaSuccess) return cuda_fail(e, "groupnorm_fold launch");
  return DLPM_OK;
}

static int g_attention_mma = 1;  // dlpm_b200_set_option("attention_mma", 0): FMA kernel (A/B and fallback for other shapes)
static int g_attention_poly = -1;  // -1 = auto = off: measured slower (MNIST forward, 5 + 6 attention blocks: 1.65 vs 1.53 ms) -- the kernel is
                                    // issue / latency-bound, not XU-bound; 1 forces the polynomial variant ("attention_poly")
namespace dlpm { void attention_set_mma(int on) { g_attention_mma = on; } void attention_set_poly(int v) { g_attention_poly = v; } }

int dlpm_b200_attention(void* out, const void* qkv, int64_t B, int L, int C, int heads, void* stream) {
  DLPM_REQUIRE(out && qkv, "attention: NULL tensor");
  DLPM_REQUIRE(heads >= 1 && C % heads == 0 && L >= 1 && L <= 1024, "attention: bad shape");
  const int D = C / heads;
  DLPM_REQUIRE(B * heads < (1ll << 31), "attention: batch too large");
  if (B == 0) return DLPM_OK;
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  auto* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  cudaStream_t s = (cudaStream_t)stream;
  {  // tensor-core path
    const size_t msmem = (size_t)2 * L * (D + 8) * 2;  // K and V tiles of one head, rows of D + 8 bf16
    // 8 warps (128 query rows) per CTA when the head has them: K / V of the head are staged half as often as with 4
    const int rows_per_cta = L >= 128 ? 128 : 64;
    const int q_blocks = (L + rows_per_cta - 1) / rows_per_cta;
    if ((D == 16 || D == 32 || D == 64) && L % 16 == 0 && msmem <= 200 * 1024 && B * heads * q_blocks < (1ll << 31) && g_attention_mma) {
      const int mthreads = L >= 128 ? 256 : (L >= 64 ? 128 : (L / 16) * 32);
      const unsigned mgrid = (unsigned)(B * heads * q_blocks);
      const bool poly = g_attention_poly > 0;
#define ATTM(DD)                                                                                                     \
  case DD: {                                                                                                         \
    auto kern = poly ? k_attention_mma<DD, true> : k_attention_mma<DD, false>;                                       \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);             \
    if (e != cudaSuccess) return cuda_fail(e, "attention smem attribute");                                           \
    cudaError_t e2 = launch_ex(kern, dim3(mgrid), dim3(mthreads), msmem, s, 1, o, q, L, C, heads, q_blocks);         \
    if (e2 != cudaSuccess) return cuda_fail(e2, "attention launch");                                                 \
  } break
      switch (D) { ATTM(16); ATTM(32); ATTM(64); }
#undef ATTM
      DLPM_CHECK_LAUNCH("attention");
      return DLPM_OK;
    }
  }
  const size_t smem = (size_t)2 * L * D * sizeof(float);
  DLPM_REQUIRE(smem <= 200 * 1024, "attention: K/V of one head do not fit in shared memory");
  int threads = L * (D / 8);  // D/8 lanes per query row
  threads = threads < 32 ? 32 : (threads > 256 ? 256 : (threads + 31) / 32 * 32);
  const unsigned grid = (unsigned)(B * heads);
#define ATT(DD)                                                                                            \
  case DD: {                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(k_attention<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "attention smem attribute");                                 \
    cudaError_t e2 = launch_ex(k_attention<DD>, dim3(grid), dim3(threads), smem, s, 1, o, q, L, C, heads);  \
    if (e2 != cudaSuccess) return cuda_fail(e2, "attention launch");                                       \
  } break
  switch (D) {
    ATT(8); ATT(16); ATT(32); ATT(64);
    default: set_error("attention: head dim %d not supported (8/16/32/64)", D); return DLPM_ERR_UNSUPPORTED;
  }
#undef ATT
  DLPM_CHECK_LAUNCH("attention");
  return DLPM_OK;
}

int dlpm_b200_conv_in(void* out, const float* x, const float* wT, const float* bias, int64_t B, int C_in, int C_out, int H, int W,
                      void* stream) {
  return dlpm_b200_conv_in_stats(out, x, wT, bias, B, C_in, C_out, H, W, nullptr, nullptr, stream);
}

int dlpm_b200_conv_in_stats(void* out, const float* x, const float* wT, const float* bias, int64_t B, int C_in, int C_out, int H, int W,
                            float* stats, int* stats_parts, void* stream) {
  if (stats_parts) *stats_parts = (W <= 32 && W % 4 == 0 && C_out <= 128) ? (H + kConvInRows - 1) / kConvInRows : 0;
  if (stats_parts && !stats) return DLPM_OK;
  DLPM_REQUIRE(!stats || (W <= 32 && W % 4 == 0 && C_out <= 128), "conv_in_stats: this shape cannot emit GroupNorm statistics");
  DLPM_REQUIRE(out && x && wT && bias, "conv_in: NULL tensor");
  DLPM_REQUIRE(C_in >= 1 && C_in <= 4 && C_out % 32 == 0 && C_out <= 512, "conv_in: C_in <= 4, C_out multiple of 32 (<= 512)");
  const int bands = (H + kConvInRows - 1) / kConvInRows;
  DLPM_REQUIRE(B * bands < (1ll << 31) && W >= 1 && W <= 256, "conv_in: bad shape");
  if (B == 0) return DLPM_OK;
  const size_t smem = (size_t)(C_out * C_in * 9 + C_out + C_in * (kConvInRows + 2) * (W + 2)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_in, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "conv_in smem attribute");
    attr = true;
  }
  DLPM_REQUIRE(smem <= 96 * 1024, "conv_in: weights do not fit in shared memory");
  cudaError_t e2 = launch_ex(k_conv_in, dim3((unsigned)(B * bands)), dim3(256), smem, (cudaStream_t)stream, 1,
                             reinterpret_cast<__nv_bfloat16*>(out), x, wT, bias, C_in, C_out, H, W, stats);
  if (e2 != cudaSuccess) return cuda_fail(e2, "conv_in launch");
  return DLPM_OK;
}

int dlpm_b200_split_input(void* out, const float* x, int64_t B, int C, int H, int W, void* stream) {
  DLPM_REQUIRE(out && x, "split_input: NULL tensor");
  DLPM_REQUIRE(C >= 1 && 3 * C <= 32 && H >= 1 && W >= 1 && B >= 0, "split_input: needs 3*C <= 32");
  if (B == 0) return DLPM_OK;
  const int64_t total = B * H * W;
  cudaError_t e = launch_ex(k_split_input, dim3((unsigned)grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
                            reinterpret_cast<__nv_bfloat16*>(out), x, C, (int64_t)H * W, total);
  if (e != cudaSuccess) return cuda_fail(e, "split_input launch");
  return DLPM_OK;
}

int dlpm_b200_upsample2x(void* out, const void* in, int64_t B, int H, int W, int C, void* stream) {
  DLPM_REQUIRE(out && in && C % 8 == 0, "upsample2x: NULL tensor or C not a multiple of 8");
  if (B == 0) return DLPM_OK;
  const int64_t total = B * 4 * H * W * (C / 8);
  k_upsample2x<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(out), reinterpret_cast<const uint4*>(in),
                                                                      B, H, W, C / 8);
  DLPM_CHECK_LAUNCH("upsample2x");
  return DLPM_OK;
}

int dlpm_b200_time_embedding(float* ss, float* semb, const float* t, const int* t_dev, float inv_T, int rows, int mc,
                             int64_t ss_total, const float* w0T, const float* b0, const float* w2T, const float* b2,
                             const float* wallT, const float* ball, void* stream) {
  DLPM_REQUIRE(ss && semb && (t || t_dev) && w0T && b0 && w2T && b2 && wallT && ball, "time_embedding: NULL tensor");
  DLPM_REQUIRE(rows >= 1 && mc >= 2 && mc <= 1024 && ss_total >= 1, "time_embedding: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = 4 * mc;
  float* h1 = ss;  // hidden layer [rows][E] parked in the ss buffer, which the third launch rewrites completely
  DLPM_REQUIRE(ss_total >= E, "time_embedding: ss_total must be >= 4*mc");
  dim3 g1((unsigned)((E + 31) / 32), (unsigned)rows);
  cudaError_t e = launch_ex(k_gemv_rows, g1, dim3(256), (size_t)(mc + 256) * sizeof(float), s, 1, h1, (const float*)nullptr, t, t_dev, inv_T,
                            1, mc, (int64_t)E, w0T, b0, 1);
  if (e != cudaSuccess) return cuda_fail(e, "time_embed layer 1");
  // every emb_layers starts with SiLU (unet.py:145-146): semb = SiLU(time_embed(.))
  e = launch_ex(k_gemv_rows, g1, dim3(256), (size_t)(E + 256) * sizeof(float), s, 1, semb, (const float*)h1, (const float*)nullptr,
                (const int*)nullptr, 0.f, 0, E, (int64_t)E, w2T, b2, 1);
  if (e != cudaSuccess) return cuda_fail(e, "time_embed layer 2");
  dim3 g3((unsigned)((ss_total + 31) / 32), (unsigned)rows);
  e = launch_ex(k_gemv_rows, g3, dim3(256), (size_t)(E + 256) * sizeof(float), s, 1, ss, (const float*)semb, (const float*)nullptr,
                (const int*)nullptr, 0.f, 0, E, ss_total, wallT, ball, 0);
  if (e != cudaSuccess) return cuda_fail(e, "emb_layers");
  return DLPM_OK;
}
