// Decode-path quantized mat-vec for sm_100a: y[M] = W[M,K] · x[K] with the reference's exact arithmetic.
//
// Replaces ggml_compute_forward_mul_mat for N == 1 (reference: models/ggml/ggml.c:11031-11245) together
// with the ops the reference runs immediately before/after it in llm_build_llama / llm_build_falcon
// (models/ggml/llama.cpp:2267-2466, 2598-2775):
//
//   prologue (every CTA, into shared memory; never touches HBM)
//     RMSNorm / LayerNorm with fp64 reductions + separate weight (+bias) multiply  ggml.c:10674-10720, 10605-10654
//     activation quantization to Q8_K (K-quants) or Q8_0 (Q4_0/Q8_0), bit-exact    k_quants.c:1191-1226, ggml.c:1232-1268
//   body (one warp per R rows; lanes stride the row in 16-byte chunks; integer dp4a dots per sub-block)
//     Q4_K k_quants.c:2550-2856 · Q5_K 3081-3430 · Q6_K 3650-4036 · Q4_0 ggml.c:2428-2697 · Q8_0 3321-3420
//   epilogue
//     store | + residual (ggml_add, llama.cpp:2415, 2453) | SiLU-table(gate)·up (ggml.c:3625-3632, llama.cpp:2438-2443)
//     | GELU-table (falcon, ggml.c:3568-3575)
//
// HBM traffic per launch = the weight planes once (algorithmic bytes) + O(K) activations from L2.
#pragma once
#include "device_types.cuh"

namespace ctb {

constexpr int MV_THREADS = 256;
constexpr int MV_WARPS = MV_THREADS / 32;
constexpr int MV_MAX_SEG = 3;

enum : int { NORM_NONE = 0, NORM_RMS = 1, NORM_LAYER = 2 };
enum : int { EPI_STORE = 0, EPI_ADD = 1, EPI_GELU = 2, EPI_ADD2 = 3 };

struct MVSeg {
  DevMat w;
  float* out;          // [M]
  const float* res;    // EPI_ADD: out = acc + res; EPI_ADD2: out = (acc + res) + res2
  const float* res2;
  int epi;
};

struct MVParams {
  const float* x;        // [K] f32 input
  const float* norm_w;   // [K] or null
  const float* norm_b;   // [K] or null (LayerNorm bias)
  float* norm_out;       // optional [K]: CTA 0 writes the normalised vector (result_norm / embeddings)
  float eps;
  int norm_mode;
  int K;
  int act;               // ACT_*
  int nseg;
  int pair_silu;         // 1: seg[0] = gate, seg[1] = up, seg[0].out[i] = silu(gate_i) * up_i
  MVSeg seg[MV_MAX_SEG];
  const uint16_t* silu_tab;   // 65536-entry fp16 tables built on the host exactly like ggml.c:4319-4333
  const uint16_t* gelu_tab;
};

// ---------------------------------------------------------------------------------------------
// Shared-memory view of the quantized activation vector
struct ActView {
  const int8_t* qs;      // [K] (Q8_K / Q8_0) — or K halves (ACT_F16) / K floats (ACT_F32)
  const float* d;        // Q8_K: per 256; Q8_0: per 32 (value already rounded through fp16)
  const int16_t* bs;     // Q8_K: bsums per 16; Q8_0: sum of the 32 quants per block
};

__host__ __device__ inline size_t act_smem_bytes(int act, int K) {
  switch (act) {
    case ACT_Q8_K: return (size_t)K + (size_t)(K / 256) * 4 + (size_t)(K / 16) * 2 + 16;
    case ACT_Q8_0: return (size_t)K + (size_t)(K / 32) * 4 + (size_t)(K / 32) * 2 + 16;
    case ACT_F16: return (size_t)K * 2;
    default: return (size_t)K * 4;
  }
}

__device__ __forceinline__ ActView act_view(int act, int K, uint8_t* smem) {
  ActView a;
  a.qs = (const int8_t*)smem;
  size_t off = (size_t)K;
  off = (off + 15) & ~(size_t)15;
  if (act == ACT_Q8_K) {
    a.d = (const float*)(smem + off);
    off += (size_t)(K / 256) * 4;
    a.bs = (const int16_t*)(smem + off);
  } else if (act == ACT_Q8_0) {
    a.d = (const float*)(smem + off);
    off += (size_t)(K / 32) * 4;
    a.bs = (const int16_t*)(smem + off);
  } else {
    a.d = nullptr;
    a.bs = nullptr;
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// Prologue pieces.  Every float operation is spelled with explicit-rounding intrinsics so nvcc cannot
// contract a*b+c into an FMA the reference does not perform.

// Block-wide fp64 sum over K elements of f(x[i]); result broadcast to all threads.
template <typename F>
__device__ __forceinline__ double block_sum_f64(int K, F f, double* red /* [MV_WARPS] smem */) {
  double s = 0.0;
  for (int i = threadIdx.x; i < K; i += MV_THREADS) s += f(i);
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < MV_WARPS; w++) t += red[w];
  return t;
}

struct NormCtx {
  int mode;
  float mean;    // LayerNorm only
  float scale;
  const float* x;
  const float* w;
  const float* b;
};

__device__ __forceinline__ float norm_apply(const NormCtx& n, int i) {
  float v = n.x[i];
  if (n.mode == NORM_NONE) return v;
  if (n.mode == NORM_LAYER) v = __fsub_rn(v, n.mean);
  v = __fmul_rn(v, n.scale);
  if (n.w) v = __fmul_rn(v, n.w[i]);
  if (n.b) v = __fadd_rn(v, n.b[i]);
  return v;
}

__device__ __forceinline__ NormCtx norm_prepare(int mode, const float* x, const float* w, const float* b, int K, float eps, double* red) {
  NormCtx n{mode, 0.f, 1.f, x, w, b};
  if (mode == NORM_RMS) {
    const double ss = block_sum_f64(K, [&](int i) { float v = x[i]; return (double)__fmul_rn(v, v); }, red);
    const float mean = (float)(ss / (double)K);
    n.scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
  } else if (mode == NORM_LAYER) {
    const double s = block_sum_f64(K, [&](int i) { return (double)x[i]; }, red);
    const float mean = (float)(s / (double)K);
    const double s2 = block_sum_f64(K, [&](int i) { float v = __fsub_rn(x[i], mean); return (double)__fmul_rn(v, v); }, red);
    const float var = (float)(s2 / (double)K);
    n.mean = mean;
    n.scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
  }
  return n;
}

// Q8_K: one warp per 256-element block, lane owns 8 consecutive elements.
// (reference quantize_row_q8_K_reference, k_quants.c:1191-1226: first element with the largest |x| fixes the
//  sign of the scale; iscale = -128/max; q = min(127, rne(iscale*x)); d = 1/iscale; bsums per 16.)
__device__ __forceinline__ void quantize_q8k_block(const float (&v)[8], int lane, int8_t* qs_out /* block base */, float* d_out,
                                                    int16_t* bs_out /* 16 entries */) {
  float amax = 0.f, mx = 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const float ax = fabsf(v[e]);
    if (ax > amax) { amax = ax; mx = v[e]; }
  }
  const float gmax = warp_max(amax);
  if (gmax == 0.f) {
    *(uint2*)(qs_out + lane * 8) = make_uint2(0u, 0u);
    if (lane < 16) bs_out[lane] = 0;
    if (lane == 0) *d_out = 0.f;
    return;
  }
  const unsigned who = __ballot_sync(0xffffffffu, amax == gmax);
  const float maxv = __shfl_sync(0xffffffffu, mx, __ffs(who) - 1);
  const float iscale = __fdiv_rn(-128.f, maxv);
  int q[8];
  int sum = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    q[e] = min(127, __float2int_rn(__fmul_rn(iscale, v[e])));
    sum += q[e];
  }
  uint2 packed;
  packed.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
  packed.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
  *(uint2*)(qs_out + lane * 8) = packed;
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  if ((lane & 1) == 0) bs_out[lane >> 1] = (int16_t)sum;
  if (lane == 0) *d_out = __fdiv_rn(1.f, iscale);
}

// Q8_0, AVX2 semantics (ggml.c:1232-1268): 4 lanes per 32-element block.
__device__ __forceinline__ void quantize_q80_group(const float (&v)[8], int lane, bool valid, int8_t* qs_out /* warp's 256-elem base */,
                                                    float* d_out /* 8 */, int16_t* sum_out /* 8 */) {
  float amax = 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) amax = fmaxf(amax, fabsf(v[e]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  const float d = __fdiv_rn(amax, 127.f);
  const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
  int q[8];
  int sum = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    q[e] = __float2int_rn(__fmul_rn(v[e], id));
    sum += q[e];
  }
  uint2 packed;
  packed.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
  packed.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
  if (valid) *(uint2*)(qs_out + lane * 8) = packed;
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  if (valid && (lane & 3) == 0) {
    d_out[lane >> 2] = h2f(f2h(d));   // the reference stores d as fp16 and multiplies with the converted value
    sum_out[lane >> 2] = (int16_t)sum;
  }
}

// Whole prologue: normalise + quantize x[K] into shared memory.  All MV_THREADS threads must call.
__device__ __forceinline__ void stage_activation(const float* x, const float* nw, const float* nb_, float* norm_out, int norm_mode, float eps,
                                                  int K, int act, uint8_t* smem, double* red, bool write_norm) {
  const NormCtx n = norm_prepare(norm_mode, x, nw, nb_, K, eps, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (act == ACT_Q8_K || act == ACT_Q8_0) {
    int8_t* qs = (int8_t*)smem;
    size_t off = ((size_t)K + 15) & ~(size_t)15;
    float* dd = (float*)(smem + off);
    int16_t* bs = (int16_t*)(smem + off + (size_t)(act == ACT_Q8_K ? K / 256 : K / 32) * 4);
    // K is a multiple of 256 for Q8_K; for Q8_0 a multiple of 32 (last warp-chunk may be partial)
    const int nchunk = (K + 255) / 256;
    for (int c = warp; c < nchunk; c += MV_WARPS) {
      float v[8];
      const int base = c * 256 + lane * 8;
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = (base + e < K) ? norm_apply(n, base + e) : 0.f;
      if (write_norm && norm_out) {
#pragma unroll
        for (int e = 0; e < 8; e++) if (base + e < K) norm_out[base + e] = v[e];
      }
      if (act == ACT_Q8_K) {
        quantize_q8k_block(v, lane, qs + c * 256, dd + c, bs + c * 16);
      } else {
        quantize_q80_group(v, lane, base < K, qs + c * 256, dd + c * 8, bs + c * 8);   // all lanes take part in the shuffles
      }
    }
  } else if (act == ACT_F16) {
    uint16_t* h = (uint16_t*)smem;
    for (int i = threadIdx.x; i < K; i += MV_THREADS) {
      const float v = norm_apply(n, i);
      if (write_norm && norm_out) norm_out[i] = v;
      h[i] = f2h(v);
    }
  } else {
    float* f = (float*)smem;
    for (int i = threadIdx.x; i < K; i += MV_THREADS) {
      const float v = norm_apply(n, i);
      if (write_norm && norm_out) norm_out[i] = v;
      f[i] = v;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Row dot products.  Each returns per-lane partial sums in acc[R]; caller warp-reduces.

// unpack the 12 scale bytes of a Q4_K/Q5_K header (k_quants.c:306-313 get_scale_min_k4, all 8 at once)
__device__ __forceinline__ void unpack_k4(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t& sc03, uint32_t& sc47, uint32_t& m03, uint32_t& m47) {
  sc03 = s0 & 0x3f3f3f3fu;
  m03 = s1 & 0x3f3f3f3fu;
  sc47 = (s2 & 0x0f0f0f0fu) | (((s0 >> 6) & 0x03030303u) << 4);
  m47 = ((s2 >> 4) & 0x0f0f0f0fu) | (((s1 >> 6) & 0x03030303u) << 4);
}
__device__ __forceinline__ int byte_of(uint32_t lo, uint32_t hi, int idx /*0..7*/) {
  const uint32_t w = idx < 4 ? lo : hi;
  return (int)((w >> ((idx & 3) * 8)) & 0xffu);
}

template <int R>
__device__ __forceinline__ void dot_q4k(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const int nb = w.nb;
  const int nchunk = nb * 8;   // 16-byte chunks of the qs plane per row
  const uint8_t* qrow[R];
  const uint8_t* hrow[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = min(row0 + r, w.M - 1);
    qrow[r] = w.qs + (size_t)row * nb * 128;
    hrow[r] = w.sc + (size_t)row * nb * 16;
  }
  for (int c = lane; c < nchunk; c += 32) {
    const int b = c >> 3, cc = c & 7, j = cc >> 1, half = cc & 1;
    int4 q[R];
    int4 h[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      q[r] = ldg_stream16(qrow[r] + (size_t)c * 16);
      h[r] = __ldg((const int4*)(hrow[r] + (size_t)b * 16));
    }
    const int8_t* ab = a.qs + b * 256 + j * 64 + half * 16;
    const int4 alo = *(const int4*)ab;
    const int4 ahi = *(const int4*)(ab + 32);
    const float yd = a.d[b];
    const int bsum = (int)a.bs[b * 16 + 2 * cc] + (int)a.bs[b * 16 + 2 * cc + 1];   // sub-block cc's 32 quants
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t sc03, sc47, m03, m47;
      unpack_k4((uint32_t)h[r].y, (uint32_t)h[r].z, (uint32_t)h[r].w, sc03, sc47, m03, m47);
      const int sc_lo = byte_of(sc03, sc47, 2 * j), sc_hi = byte_of(sc03, sc47, 2 * j + 1);
      const int mn = byte_of(m03, m47, cc);
      int s_lo = 0, s_hi = 0;
      s_lo = __dp4a((int)((uint32_t)q[r].x & 0x0f0f0f0fu), alo.x, s_lo);
      s_lo = __dp4a((int)((uint32_t)q[r].y & 0x0f0f0f0fu), alo.y, s_lo);
      s_lo = __dp4a((int)((uint32_t)q[r].z & 0x0f0f0f0fu), alo.z, s_lo);
      s_lo = __dp4a((int)((uint32_t)q[r].w & 0x0f0f0f0fu), alo.w, s_lo);
      s_hi = __dp4a((int)(((uint32_t)q[r].x >> 4) & 0x0f0f0f0fu), ahi.x, s_hi);
      s_hi = __dp4a((int)(((uint32_t)q[r].y >> 4) & 0x0f0f0f0fu), ahi.y, s_hi);
      s_hi = __dp4a((int)(((uint32_t)q[r].z >> 4) & 0x0f0f0f0fu), ahi.z, s_hi);
      s_hi = __dp4a((int)(((uint32_t)q[r].w >> 4) & 0x0f0f0f0fu), ahi.w, s_hi);
      const float dw = h2f((uint16_t)((uint32_t)h[r].x & 0xffffu));
      const float dmin = h2f((uint16_t)((uint32_t)h[r].x >> 16));
      const int isum = sc_lo * s_lo + sc_hi * s_hi;
      acc[r] += (yd * dw) * (float)isum - (yd * dmin) * (float)(mn * bsum);
    }
  }
}

template <int R>
__device__ __forceinline__ void dot_q5k(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const int nb = w.nb;
  const int nchunk = nb * 8;
  const uint8_t* qrow[R];
  const uint8_t* hrow[R];
  const uint8_t* brow[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = min(row0 + r, w.M - 1);
    qrow[r] = w.qs + (size_t)row * nb * 128;
    hrow[r] = w.sc + (size_t)row * nb * 16;
    brow[r] = w.qh + (size_t)row * nb * 32;
  }
  for (int c = lane; c < nchunk; c += 32) {
    const int b = c >> 3, cc = c & 7, j = cc >> 1, half = cc & 1;
    int4 q[R], h[R], hb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      q[r] = ldg_stream16(qrow[r] + (size_t)c * 16);
      h[r] = __ldg((const int4*)(hrow[r] + (size_t)b * 16));
      hb[r] = __ldg((const int4*)(brow[r] + (size_t)b * 32 + half * 16));
    }
    const int8_t* ab = a.qs + b * 256 + j * 64 + half * 16;
    const int4 alo = *(const int4*)ab;
    const int4 ahi = *(const int4*)(ab + 32);
    const float yd = a.d[b];
    const int bsum = (int)a.bs[b * 16 + 2 * cc] + (int)a.bs[b * 16 + 2 * cc + 1];
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t sc03, sc47, m03, m47;
      unpack_k4((uint32_t)h[r].y, (uint32_t)h[r].z, (uint32_t)h[r].w, sc03, sc47, m03, m47);
      const int sc_lo = byte_of(sc03, sc47, 2 * j), sc_hi = byte_of(sc03, sc47, 2 * j + 1);
      const int mn = byte_of(m03, m47, cc);
      // 5th bit: bit (2j) of qh[l] for the low-nibble sub-block, bit (2j+1) for the high-nibble one (k_quants.c:3385-3394)
      const uint32_t qv[4] = {(uint32_t)q[r].x, (uint32_t)q[r].y, (uint32_t)q[r].z, (uint32_t)q[r].w};
      const uint32_t hv[4] = {(uint32_t)hb[r].x >> (2 * j), (uint32_t)hb[r].y >> (2 * j), (uint32_t)hb[r].z >> (2 * j), (uint32_t)hb[r].w >> (2 * j)};
      const int av_lo[4] = {alo.x, alo.y, alo.z, alo.w};
      const int av_hi[4] = {ahi.x, ahi.y, ahi.z, ahi.w};
      int s_lo = 0, s_hi = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t lo = (qv[i] & 0x0f0f0f0fu) | ((hv[i] << 4) & 0x10101010u);
        const uint32_t hi = ((qv[i] >> 4) & 0x0f0f0f0fu) | ((hv[i] << 3) & 0x10101010u);
        s_lo = __dp4a((int)lo, av_lo[i], s_lo);
        s_hi = __dp4a((int)hi, av_hi[i], s_hi);
      }
      const float dw = h2f((uint16_t)((uint32_t)h[r].x & 0xffffu));
      const float dmin = h2f((uint16_t)((uint32_t)h[r].x >> 16));
      const int isum = sc_lo * s_lo + sc_hi * s_hi;
      acc[r] += (yd * dw) * (float)isum - (yd * dmin) * (float)(mn * bsum);
    }
  }
}

// Q6_K: lane-unit = 64 weights = ql chunks (n, 16h) and (n, 32+16h) + qh chunk (n, 16h); 4 units per block.
// q = (nibble | two high bits << 4) - 32; the "-32" is folded out with the activation's bsums, exactly like the
// reference AVX2 path does (k_quants.c:3758-3830).
template <int R>
__device__ __forceinline__ void dot_q6k(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const int nb = w.nb;
  const int nunit = nb * 4;
  const uint8_t* lrow[R];
  const uint8_t* hrow[R];
  const uint8_t* srow[R];
  const uint16_t* drow[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = min(row0 + r, w.M - 1);
    lrow[r] = w.qs + (size_t)row * nb * 128;
    hrow[r] = w.qh + (size_t)row * nb * 64;
    srow[r] = w.sc + (size_t)row * nb * 16;
    drow[r] = w.d + (size_t)row * nb;
  }
  for (int c = lane; c < nunit; c += 32) {
    const int b = c >> 2, u = c & 3, n = u >> 1, hh = u & 1;
    int4 qa[R], qb[R], qh[R];
    int2 scv[R];
    uint16_t dv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      qa[r] = ldg_stream16(lrow[r] + (size_t)b * 128 + n * 64 + hh * 16);
      qb[r] = ldg_stream16(lrow[r] + (size_t)b * 128 + n * 64 + 32 + hh * 16);
      qh[r] = ldg_stream16(hrow[r] + (size_t)b * 64 + n * 32 + hh * 16);
      scv[r] = __ldg((const int2*)(srow[r] + (size_t)b * 16 + n * 8));
      dv[r] = __ldg(drow[r] + b);
    }
    const int8_t* ab = a.qs + b * 256 + n * 128 + hh * 16;
    const int4 a0 = *(const int4*)(ab);
    const int4 a1 = *(const int4*)(ab + 32);
    const int4 a2 = *(const int4*)(ab + 64);
    const int4 a3 = *(const int4*)(ab + 96);
    const int16_t* bsp = a.bs + b * 16 + n * 8 + hh;
    const int bs0 = bsp[0], bs1 = bsp[2], bs2 = bsp[4], bs3 = bsp[6];
    const float yd = a.d[b];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t A[4] = {(uint32_t)qa[r].x, (uint32_t)qa[r].y, (uint32_t)qa[r].z, (uint32_t)qa[r].w};
      const uint32_t B[4] = {(uint32_t)qb[r].x, (uint32_t)qb[r].y, (uint32_t)qb[r].z, (uint32_t)qb[r].w};
      const uint32_t H[4] = {(uint32_t)qh[r].x, (uint32_t)qh[r].y, (uint32_t)qh[r].z, (uint32_t)qh[r].w};
      const int x0[4] = {a0.x, a0.y, a0.z, a0.w};
      const int x1[4] = {a1.x, a1.y, a1.z, a1.w};
      const int x2[4] = {a2.x, a2.y, a2.z, a2.w};
      const int x3[4] = {a3.x, a3.y, a3.z, a3.w};
      int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t w0 = (A[i] & 0x0f0f0f0fu) | ((H[i] << 4) & 0x30303030u);
        const uint32_t w1 = (B[i] & 0x0f0f0f0fu) | ((H[i] << 2) & 0x30303030u);
        const uint32_t w2 = ((A[i] >> 4) & 0x0f0f0f0fu) | (H[i] & 0x30303030u);
        const uint32_t w3 = ((B[i] >> 4) & 0x0f0f0f0fu) | ((H[i] >> 2) & 0x30303030u);
        s0 = __dp4a((int)w0, x0[i], s0);
        s1 = __dp4a((int)w1, x1[i], s1);
        s2 = __dp4a((int)w2, x2[i], s2);
        s3 = __dp4a((int)w3, x3[i], s3);
      }
      // int8 scales for sub-blocks 8n + hh + {0,2,4,6}
      const uint32_t slo = (uint32_t)scv[r].x, shi = (uint32_t)scv[r].y;
      const int c0 = (int)(int8_t)((slo >> (hh * 8)) & 0xff);
      const int c1 = (int)(int8_t)((slo >> (16 + hh * 8)) & 0xff);
      const int c2 = (int)(int8_t)((shi >> (hh * 8)) & 0xff);
      const int c3 = (int)(int8_t)((shi >> (16 + hh * 8)) & 0xff);
      const int isum = c0 * (s0 - 32 * bs0) + c1 * (s1 - 32 * bs1) + c2 * (s2 - 32 * bs2) + c3 * (s3 - 32 * bs3);
      acc[r] += (h2f(dv[r]) * yd) * (float)isum;
    }
  }
}

// Q4_0: one lane per 32-weight block (16 B of nibbles: byte j = element j | element j+16 << 4, value - 8).
template <int R>
__device__ __forceinline__ void dot_q40(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const int nb = w.nb;
  const uint8_t* qrow[R];
  const uint16_t* drow[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = min(row0 + r, w.M - 1);
    qrow[r] = w.qs + (size_t)row * nb * 16;
    drow[r] = w.d + (size_t)row * nb;
  }
  for (int b = lane; b < nb; b += 32) {
    int4 q[R];
    uint16_t dv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      q[r] = ldg_stream16(qrow[r] + (size_t)b * 16);
      dv[r] = __ldg(drow[r] + b);
    }
    const int4 alo = *(const int4*)(a.qs + b * 32);
    const int4 ahi = *(const int4*)(a.qs + b * 32 + 16);
    const float yd = a.d[b];
    const int ysum = a.bs[b];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int s = 0;
      s = __dp4a((int)((uint32_t)q[r].x & 0x0f0f0f0fu), alo.x, s);
      s = __dp4a((int)((uint32_t)q[r].y & 0x0f0f0f0fu), alo.y, s);
      s = __dp4a((int)((uint32_t)q[r].z & 0x0f0f0f0fu), alo.z, s);
      s = __dp4a((int)((uint32_t)q[r].w & 0x0f0f0f0fu), alo.w, s);
      s = __dp4a((int)(((uint32_t)q[r].x >> 4) & 0x0f0f0f0fu), ahi.x, s);
      s = __dp4a((int)(((uint32_t)q[r].y >> 4) & 0x0f0f0f0fu), ahi.y, s);
      s = __dp4a((int)(((uint32_t)q[r].z >> 4) & 0x0f0f0f0fu), ahi.z, s);
      s = __dp4a((int)(((uint32_t)q[r].w >> 4) & 0x0f0f0f0fu), ahi.w, s);
      acc[r] += (h2f(dv[r]) * yd) * (float)(s - 8 * ysum);
    }
  }
}

template <int R>
__device__ __forceinline__ void dot_q80(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const int nb = w.nb;
  const uint8_t* qrow[R];
  const uint16_t* drow[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = min(row0 + r, w.M - 1);
    qrow[r] = w.qs + (size_t)row * nb * 32;
    drow[r] = w.d + (size_t)row * nb;
  }
  for (int b = lane; b < nb; b += 32) {
    int4 q0[R], q1[R];
    uint16_t dv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      q0[r] = ldg_stream16(qrow[r] + (size_t)b * 32);
      q1[r] = ldg_stream16(qrow[r] + (size_t)b * 32 + 16);
      dv[r] = __ldg(drow[r] + b);
    }
    const int4 a0 = *(const int4*)(a.qs + b * 32);
    const int4 a1 = *(const int4*)(a.qs + b * 32 + 16);
    const float yd = a.d[b];
#pragma unroll
    for (int r = 0; r < R; r++) {
      int s = 0;
      s = __dp4a(q0[r].x, a0.x, s); s = __dp4a(q0[r].y, a0.y, s); s = __dp4a(q0[r].z, a0.z, s); s = __dp4a(q0[r].w, a0.w, s);
      s = __dp4a(q1[r].x, a1.x, s); s = __dp4a(q1[r].y, a1.y, s); s = __dp4a(q1[r].z, a1.z, s); s = __dp4a(q1[r].w, a1.w, s);
      acc[r] += (h2f(dv[r]) * yd) * (float)s;
    }
  }
}

// F16 weights · f16-rounded activations, fp32 accumulate (ggml.c:2392-2426); F32 weights · f32 activations.
template <int R>
__device__ __forceinline__ void dot_f16(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const uint16_t* xa = (const uint16_t*)a.qs;
  for (int i = lane * 8; i < w.K; i += 256) {
    const int4 av = *(const int4*)(xa + i);
    const uint32_t aw[4] = {(uint32_t)av.x, (uint32_t)av.y, (uint32_t)av.z, (uint32_t)av.w};
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int row = min(row0 + r, w.M - 1);
      const int4 wv = ldg_stream16(w.qs + ((size_t)row * w.K + i) * 2);
      const uint32_t ww[4] = {(uint32_t)wv.x, (uint32_t)wv.y, (uint32_t)wv.z, (uint32_t)wv.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        acc[r] += h2f((uint16_t)(ww[e] & 0xffff)) * h2f((uint16_t)(aw[e] & 0xffff));
        acc[r] += h2f((uint16_t)(ww[e] >> 16)) * h2f((uint16_t)(aw[e] >> 16));
      }
    }
  }
}
template <int R>
__device__ __forceinline__ void dot_f32(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  const float* xa = (const float*)a.qs;
  for (int i = lane * 4; i < w.K; i += 128) {
    const float4 av = *(const float4*)(xa + i);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int row = min(row0 + r, w.M - 1);
      const int4 wv = ldg_stream16(w.qs + ((size_t)row * w.K + i) * 4);
      acc[r] += __int_as_float(wv.x) * av.x + __int_as_float(wv.y) * av.y + __int_as_float(wv.z) * av.z + __int_as_float(wv.w) * av.w;
    }
  }
}

template <int R>
__device__ __forceinline__ void dot_rows(const DevMat& w, int row0, const ActView& a, int lane, float (&acc)[R]) {
  switch (w.type) {
    case GT_Q4_K: dot_q4k<R>(w, row0, a, lane, acc); break;
    case GT_Q6_K: dot_q6k<R>(w, row0, a, lane, acc); break;
    case GT_Q5_K: dot_q5k<R>(w, row0, a, lane, acc); break;
    case GT_Q4_0: dot_q40<R>(w, row0, a, lane, acc); break;
    case GT_Q8_0: dot_q80<R>(w, row0, a, lane, acc); break;
    case GT_F16: dot_f16<R>(w, row0, a, lane, acc); break;
    default: dot_f32<R>(w, row0, a, lane, acc); break;
  }
}

__device__ __forceinline__ float table_f16(const uint16_t* tab, float x) { return h2f(__ldg(tab + f2h(x))); }

// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MV_THREADS) k_matvec(const __grid_constant__ MVParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  stage_activation(p.x, p.norm_w, p.norm_b, p.norm_out, p.norm_mode, p.eps, p.K, p.act, smem, red, blockIdx.x == 0);
  const ActView a = act_view(p.act, p.K, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * MV_WARPS + warp, nw = gridDim.x * MV_WARPS;

  if (p.pair_silu) {
    const DevMat& g = p.seg[0].w;
    const DevMat& u = p.seg[1].w;
    const int units = (g.M + R - 1) / R;
    for (int un = gw; un < units; un += nw) {
      const int row0 = un * R;
      float ag[R], au[R];
#pragma unroll
      for (int r = 0; r < R; r++) { ag[r] = 0.f; au[r] = 0.f; }
      dot_rows<R>(g, row0, a, lane, ag);
      dot_rows<R>(u, row0, a, lane, au);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float vg = warp_sum(ag[r]);
        const float vu = warp_sum(au[r]);
        if (lane == 0 && row0 + r < g.M) p.seg[0].out[row0 + r] = __fmul_rn(table_f16(p.silu_tab, vg), vu);
      }
    }
    return;
  }

  int unit_base = 0;
  for (int s = 0; s < p.nseg; s++) {
    const MVSeg& sg = p.seg[s];
    const int units = (sg.w.M + R - 1) / R;
    // continue the global striding across segments so all warps stay busy
    int first = gw - (unit_base % nw);
    if (first < 0) first += nw;
    for (int un = first; un < units; un += nw) {
      const int row0 = un * R;
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; r++) acc[r] = 0.f;
      dot_rows<R>(sg.w, row0, a, lane, acc);
#pragma unroll
      for (int r = 0; r < R; r++) {
        float v = warp_sum(acc[r]);
        const int row = row0 + r;
        if (lane == 0 && row < sg.w.M) {
          if (sg.epi == EPI_ADD) v = __fadd_rn(v, sg.res[row]);
          else if (sg.epi == EPI_ADD2) v = __fadd_rn(__fadd_rn(v, sg.res[row]), sg.res2[row]);
          else if (sg.epi == EPI_GELU) v = table_f16(p.gelu_tab, v);
          sg.out[row] = v;
        }
      }
    }
    unit_base += units;
  }
}

}  // namespace ctb
