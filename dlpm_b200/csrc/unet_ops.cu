// K6 / K7 and the small memory-bound UNet operators (NHWC bf16 activations, fp32 statistics).
//   GroupNorm + scale-shift + SiLU   unet.py:141-142,153-154,188-191,212,433-434 (GroupNorm32 nn.py:17-19)
//   QKV attention                    unet.py:231-250
//   input conv 3->C                  unet.py:347          nearest upsample x2   unet.py:73
//   timestep embedding + emb_layers  nn.py:103-121, unet.py:335-339,145-151,477
#include "../../include/dlpm_b200_unet.h"
#include "common.cuh"

namespace dlpm {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

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
// K7 attention: one CTA per (sample, head); K and V of the head staged in shared memory as fp32;
// one query row per thread with an online softmax.  L <= 1024, d = C/heads <= 64.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_attention(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv, int L,
                                                   int C, int heads) {
  extern __shared__ float sm[];
  float* Ks = sm;            // [L][D]
  float* Vs = sm + L * D;    // [L][D]
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const __nv_bfloat16* base = qkv + (int64_t)n * L * 3 * C + h * 3 * D;  // per head: q | k | v blocks of D channels
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    const int s = i / D, d = i - s * D;
    Ks[i] = __bfloat162float(base[(int64_t)s * 3 * C + D + d]);
    Vs[i] = __bfloat162float(base[(int64_t)s * 3 * C + 2 * D + d]);
  }
  __syncthreads();
  const float scale2 = rsqrtf((float)D);  // (1/sqrt(sqrt(d)))^2 applied to q and k (unet.py:244-247)
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    float qv[D], o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { qv[d] = __bfloat162float(base[(int64_t)t * 3 * C + d]) * scale2; o[d] = 0.f; }
    float mx = -INFINITY, den = 0.f;
    for (int s = 0; s < L; ++s) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) dot = fmaf(qv[d], Ks[s * D + d], dot);
      const float nm = fmaxf(mx, dot);
      const float corr = __expf(mx - nm), pw = __expf(dot - nm);
      den = den * corr + pw;
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = fmaf(o[d], corr, pw * Vs[s * D + d]);
      mx = nm;
    }
    const float inv = 1.0f / den;
    __nv_bfloat16* dst = out + ((int64_t)n * L + t) * C + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) dst[d] = __float2bfloat16(o[d] * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// input conv: NCHW fp32 -> NHWC bf16, 3x3 pad 1, C_in <= 4.  Thread = (pixel, 8 output channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_conv_in(__nv_bfloat16* __restrict__ out, const float* __restrict__ x,
                                                 const float* __restrict__ w, const float* __restrict__ bias, int64_t B, int C_in,
                                                 int C_out, int H, int W) {
  extern __shared__ float sm[];  // w [C_out][C_in*9], bias [C_out]
  const int K = C_in * 9;
  for (int i = threadIdx.x; i < C_out * K; i += blockDim.x) sm[i] = w[i];
  for (int i = threadIdx.x; i < C_out; i += blockDim.x) sm[C_out * K + i] = bias[i];
  __syncthreads();
  const int groups = C_out >> 3;
  const int64_t total = B * H * W * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const int64_t pix = idx / groups;
    const int xw = (int)(pix % W), yh = (int)((pix / W) % H);
    const int64_t n = pix / ((int64_t)W * H);
    float in[36];
    for (int ci = 0; ci < C_in; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
        in[ci * 9 + t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((n * C_in + ci) * H + yy) * W + xx) : 0.f;
      }
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float* wr = sm + (g * 8 + e) * K;
      float a = sm[C_out * K + g * 8 + e];
      for (int k = 0; k < K; ++k) a = fmaf(in[k], wr[k], a);
      acc[e] = a;
    }
    *reinterpret_cast<uint4*>(out + pix * C_out + g * 8) = pack8(acc);
  }
}

__global__ void __launch_bounds__(256) k_upsample2x(uint4* __restrict__ out, const uint4* __restrict__ in, int64_t B, int H, int W,
                                                    int C8) {
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
// timestep embedding (one CTA per row) and the concatenated emb_layers GEMV
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_time_embed(float* __restrict__ semb, const float* __restrict__ t,
                                                    const int* __restrict__ t_dev, float inv_T, int mc,
                                                    const float* __restrict__ w0T, const float* __restrict__ b0,
                                                    const float* __restrict__ w2T, const float* __restrict__ b2) {
  extern __shared__ float sm[];  // e0 [mc], h1 [4mc]
  float* e0 = sm;
  float* h1 = sm + mc;
  const int r = blockIdx.x, E = 4 * mc, half = mc / 2;
  const float tv = t_dev ? (float)(*t_dev) * inv_T : t[r];
  for (int i = threadIdx.x; i < mc; i += blockDim.x) {
    if (i < 2 * half) {
      const int k = i < half ? i : i - half;
      const float freq = expf(-9.210340371976184f * (float)k / (float)half);  // exp(-ln(10000) k / half), nn.py:113-115
      const float arg = tv * freq;
      e0[i] = i < half ? cosf(arg) : sinf(arg);
    } else {
      e0[i] = 0.f;  // odd dim padding (nn.py:118-119)
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    float a = __ldg(b0 + j);
    for (int k = 0; k < mc; ++k) a = fmaf(e0[k], __ldg(w0T + (int64_t)k * E + j), a);
    h1[j] = a / (1.0f + expf(-a));
  }
  __syncthreads();
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    float a = __ldg(b2 + j);
    for (int k = 0; k < E; ++k) a = fmaf(h1[k], __ldg(w2T + (int64_t)k * E + j), a);
    semb[(int64_t)r * E + j] = a / (1.0f + expf(-a));  // every emb_layers starts with SiLU (unet.py:145-146)
  }
}

__global__ void __launch_bounds__(256) k_emb_layers(float* __restrict__ ss, const float* __restrict__ semb, int E, int64_t ss_total,
                                                    const float* __restrict__ wallT, const float* __restrict__ ball) {
  extern __shared__ float sm[];  // semb row [E]
  const int r = blockIdx.y;
  for (int i = threadIdx.x; i < E; i += blockDim.x) sm[i] = semb[(int64_t)r * E + i];
  __syncthreads();
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (o >= ss_total) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 3 < E; k += 4) {
    a0 = fmaf(sm[k], __ldg(wallT + (int64_t)k * ss_total + o), a0);
    a1 = fmaf(sm[k + 1], __ldg(wallT + (int64_t)(k + 1) * ss_total + o), a1);
    a2 = fmaf(sm[k + 2], __ldg(wallT + (int64_t)(k + 2) * ss_total + o), a2);
    a3 = fmaf(sm[k + 3], __ldg(wallT + (int64_t)(k + 3) * ss_total + o), a3);
  }
  for (; k < E; ++k) a0 = fmaf(sm[k], __ldg(wallT + (int64_t)k * ss_total + o), a0);
  ss[(int64_t)r * ss_total + o] = __ldg(ball + o) + ((a0 + a1) + (a2 + a3));
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
  const int nvec = C / 8, slots = kGnThreads / nvec;
  const size_t smem = (size_t)(2 * slots * C + 2 * C) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_groupnorm, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "groupnorm smem attribute");
    attr = true;
  }
  k_groupnorm<<<(unsigned)B, kGnThreads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(in0), C0,
      reinterpret_cast<const __nv_bfloat16*>(in1), C1, HW, gamma, beta, ss, ss_rows, ss_stride, ss_off, apply_silu);
  DLPM_CHECK_LAUNCH("groupnorm");
  return DLPM_OK;
}

int dlpm_b200_attention(void* out, const void* qkv, int64_t B, int L, int C, int heads, void* stream) {
  DLPM_REQUIRE(out && qkv, "attention: NULL tensor");
  DLPM_REQUIRE(heads >= 1 && C % heads == 0 && L >= 1 && L <= 1024, "attention: bad shape");
  const int D = C / heads;
  DLPM_REQUIRE(B * heads < (1ll << 31), "attention: batch too large");
  if (B == 0) return DLPM_OK;
  const size_t smem = (size_t)2 * L * D * sizeof(float);
  DLPM_REQUIRE(smem <= 200 * 1024, "attention: K/V of one head do not fit in shared memory");
  const int threads = L < 32 ? 32 : (L > 256 ? 256 : L);
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  auto* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(B * heads);
#define ATT(DD)                                                                                            \
  case DD: {                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(k_attention<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "attention smem attribute");                                 \
    k_attention<DD><<<grid, threads, smem, s>>>(o, q, L, C, heads);                                        \
  } break
  switch (D) {
    ATT(8); ATT(16); ATT(32); ATT(64);
    default: set_error("attention: head dim %d not supported (8/16/32/64)", D); return DLPM_ERR_UNSUPPORTED;
  }
#undef ATT
  DLPM_CHECK_LAUNCH("attention");
  return DLPM_OK;
}

int dlpm_b200_conv_in(void* out, const float* x, const float* w, const float* bias, int64_t B, int C_in, int C_out, int H, int W,
                      void* stream) {
  DLPM_REQUIRE(out && x && w && bias, "conv_in: NULL tensor");
  DLPM_REQUIRE(C_in >= 1 && C_in <= 4 && C_out % 8 == 0 && C_out <= 512, "conv_in: C_in <= 4, C_out multiple of 8 (<= 512)");
  if (B == 0) return DLPM_OK;
  const size_t smem = (size_t)(C_out * C_in * 9 + C_out) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_in, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "conv_in smem attribute");
    attr = true;
  }
  const int64_t total = B * H * W * (C_out / 8);
  k_conv_in<<<grid_for(total, 256, 4), 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(out), x, w, bias, B, C_in,
                                                                         C_out, H, W);
  DLPM_CHECK_LAUNCH("conv_in");
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
  k_time_embed<<<rows, 512, (size_t)5 * mc * sizeof(float), s>>>(semb, t, t_dev, inv_T, mc, w0T, b0, w2T, b2);
  DLPM_CHECK_LAUNCH("time_embed");
  dim3 grid((unsigned)((ss_total + 255) / 256), (unsigned)rows);
  k_emb_layers<<<grid, 256, (size_t)4 * mc * sizeof(float), s>>>(ss, semb, 4 * mc, ss_total, wallT, ball);
  DLPM_CHECK_LAUNCH("emb_layers");
  return DLPM_OK;
}
