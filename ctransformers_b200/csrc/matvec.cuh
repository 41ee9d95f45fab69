// Decode-path quantized mat-vec for sm_100a: y[M] = W[M,K] · x[K], BIT-EXACT with the reference's AVX2 CPU build.
//
// Replaces ggml_compute_forward_mul_mat for N == 1 (reference: models/ggml/ggml.c:11031-11245) together
// with the ops the reference runs immediately before/after it in llm_build_llama / llm_build_falcon
// (models/ggml/llama.cpp:2267-2466, 2598-2775):
//
//   prologue (every CTA, into shared memory; never touches HBM)
//     RMSNorm / LayerNorm with fp64 reductions + separate weight (+bias) multiply  ggml.c:10674-10720, 10605-10654
//     activation quantization to Q8_K (K-quants) or Q8_0 (Q4_0/Q8_0), bit-exact    k_quants.c:1191-1226, ggml.c:1232-1268
//   body — the reference's AVX2 kernels restated lane for lane.  Each AVX2 kernel keeps an 8-lane fp32 accumulator in which
//     lane l holds, per block, (float)(integer dot of elements 4l..4l+3 of every 32-element group, times the sub-block scales)
//     folded in with ONE fmadd per block, blocks in order, and ends with hsum_float_8.  Here 8 GPU lanes play the 8 AVX lanes
//     of one weight row, a warp carries 4 rows, dp4a does the 4-element integer dots, and the per-block fmaf chain and the
//     final shuffle tree reproduce the float order — results equal the reference's bit for bit:
//     Q4_K k_quants.c:2651-2714 · Q5_K 3174-3262 · Q6_K 3794-3872 · Q4_0 ggml.c:2500-2525 · Q8_0 3379-3402 · F16 2392-2426
//   epilogue
//     store | + residual (ggml_add, llama.cpp:2415, 2453) | SiLU-table(gate)·up (ggml.c:3625-3632, llama.cpp:2438-2443)
//     | GELU-table (falcon, ggml.c:3568-3575)
//
// Why exactness matters: the next mat-mul re-quantizes this output to int8; a 1-ulp difference can flip one rounding and the
// logits then differ by ~1e-3 (measured).  HBM traffic per launch = the weight planes once + O(K) activations from L2.
#pragma once
#include "device_types.cuh"

namespace ctb {

constexpr int MV_THREADS = 256;
constexpr int MV_WARPS = MV_THREADS / 32;
constexpr int MV_ROWS = 4;   // quantized weights: rows per warp (one per 8-lane group)
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
// Shared-memory view of the quantized activation vector.
//   Q8_K: qs is stored lane-major per block: byte offset of int8 word (sub-block s, lane l) = ((b*2 + (s>>2))*8 + l)*16 + (s&3)*4,
//         so GPU lane l reads its 8 words of a block with two conflict-free 16-byte loads.  d: per block.  bs: bsums, natural.
//   Q8_0: natural order; d per 32 (already rounded through fp16).
struct ActView {
  const int8_t* qs;
  const float* d;
  const int16_t* bs;
};

__host__ __device__ inline size_t act_smem_bytes(int act, int K) {
  switch (act) {
    case ACT_Q8_K: return (size_t)K + (((size_t)(K / 256) * 4 + 15) & ~(size_t)15) + (size_t)(K / 16) * 2 + 16;
    case ACT_Q8_0: return (size_t)K + (size_t)(K / 32) * 4 + 16;
    case ACT_F16: return (size_t)K * 2;
    default: return (size_t)K * 4;
  }
}

// bytes reserved for the per-block d values of a Q8_K vector (keeps the bsums that follow 16-byte aligned)
__host__ __device__ inline size_t q8k_d_bytes(int K) { return ((size_t)(K / 256) * 4 + 15) & ~(size_t)15; }

__device__ __forceinline__ ActView act_view(int act, int K, uint8_t* smem) {
  ActView a;
  a.qs = (const int8_t*)smem;
  const size_t off = ((size_t)K + 15) & ~(size_t)15;
  a.d = (const float*)(smem + off);
  a.bs = (const int16_t*)(smem + off + q8k_d_bytes(K));
  return a;
}

__device__ __forceinline__ int q8k_word_offset(int b, int s, int l) { return ((b * 2 + (s >> 2)) * 8 + l) * 16 + (s & 3) * 4; }

// ---------------------------------------------------------------------------------------------
// Prologue pieces.  Every float operation is spelled with explicit-rounding intrinsics so nvcc cannot
// contract a*b+c into an FMA the reference does not perform — and fuses exactly where the reference binary does.

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

// Q8_K: one warp per 256-element block, lane owns 8 consecutive elements (two int8 words).
// reference quantize_row_q8_K_reference, k_quants.c:1191-1226: first element with the largest |x| fixes the sign of the
// scale; iscale = -128/max; q = min(127, nearest_int(iscale*x)); d = 1/iscale; bsums per 16.  nearest_int adds 12582912.f and
// reads the mantissa — and the reference BINARY fuses iscale*x + 12582912.f into one vfmadd (the loop is auto-vectorised), so
// the exact product is rounded once.  __fmaf_rn reproduces that.
__device__ __forceinline__ void quantize_q8k_block(const float (&v)[8], int lane, int b, int8_t* qs_base, float* d_out, int16_t* bs_out /* 16 entries */) {
  float amax = 0.f, mx = 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const float ax = fabsf(v[e]);
    if (ax > amax) { amax = ax; mx = v[e]; }
  }
  const float gmax = warp_max(amax);
  const int w0 = 2 * lane, w1 = 2 * lane + 1;
  uint32_t* dst0 = (uint32_t*)(qs_base + q8k_word_offset(b, w0 >> 3, w0 & 7));
  uint32_t* dst1 = (uint32_t*)(qs_base + q8k_word_offset(b, w1 >> 3, w1 & 7));
  if (gmax == 0.f) {
    *dst0 = 0u; *dst1 = 0u;
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
    const float val = __fmaf_rn(iscale, v[e], 12582912.f);
    q[e] = min(127, (__float_as_int(val) & 0x007fffff) - 0x00400000);
    sum += q[e];
  }
  *dst0 = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
  *dst1 = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  if ((lane & 1) == 0) bs_out[lane >> 1] = (int16_t)sum;
  if (lane == 0) *d_out = __fdiv_rn(1.f, iscale);
}

// Q8_0, AVX2 semantics (ggml.c:1232-1268): 4 lanes per 32-element block; d = amax/127 (kept as fp16), id = 127/amax, RNE.
__device__ __forceinline__ void quantize_q80_group(const float (&v)[8], int lane, bool valid, int8_t* qs_out /* warp's 256-elem base */, float* d_out /* 8 */) {
  float amax = 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) amax = fmaxf(amax, fabsf(v[e]));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  const float d = __fdiv_rn(amax, 127.f);
  const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
  int q[8];
#pragma unroll
  for (int e = 0; e < 8; e++) q[e] = __float2int_rn(__fmul_rn(v[e], id));
  uint2 packed;
  packed.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
  packed.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
  if (valid) *(uint2*)(qs_out + lane * 8) = packed;
  if (valid && (lane & 3) == 0) d_out[lane >> 2] = h2f(f2h(d));   // the reference stores d as fp16 and multiplies with the converted value
}

// Whole prologue: normalise + quantize x[K] into shared memory.  All MV_THREADS threads must call.
__device__ __forceinline__ void stage_activation(const float* x, const float* nw, const float* nb_, float* norm_out, int norm_mode, float eps,
                                                  int K, int act, uint8_t* smem, double* red, bool write_norm) {
  const NormCtx n = norm_prepare(norm_mode, x, nw, nb_, K, eps, red);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (act == ACT_Q8_K || act == ACT_Q8_0) {
    int8_t* qs = (int8_t*)smem;
    const size_t off = ((size_t)K + 15) & ~(size_t)15;
    float* dd = (float*)(smem + off);
    int16_t* bs = (int16_t*)(smem + off + q8k_d_bytes(K));
    const int nchunk = (K + 255) / 256;   // Q8_K: K % 256 == 0; Q8_0: K % 32 == 0, the last warp-chunk may be partial
    for (int c = warp; c < nchunk; c += MV_WARPS) {
      float v[8];
      const int base = c * 256 + lane * 8;
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = (base + e < K) ? norm_apply(n, base + e) : 0.f;
      if (write_norm && norm_out) {
#pragma unroll
        for (int e = 0; e < 8; e++) if (base + e < K) norm_out[base + e] = v[e];
      }
      if (act == ACT_Q8_K) quantize_q8k_block(v, lane, c, qs, dd + c, bs + c * 16);
      else quantize_q80_group(v, lane, base < K, qs + c * 256, dd + c * 8);   // all lanes take part in the shuffles
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
// hsum_float_8 (ggml.c:609-615) over the 8 lanes of a group: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); every lane ends with the sum.
__device__ __forceinline__ float group_hsum8(float v) {
  v = v + __shfl_xor_sync(0xffffffffu, v, 4);
  v = v + __shfl_xor_sync(0xffffffffu, v, 2);
  v = v + __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
// the 4-lane tail of the Q4_K mins accumulator: (m0+m2)+(m1+m3), valid in lanes l < 4
__device__ __forceinline__ float group_hsum4(float v) {
  v = v + __shfl_xor_sync(0xffffffffu, v, 2);
  v = v + __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// unpack the 12 scale bytes of a Q4_K/Q5_K header (k_quants.c:306-313 get_scale_min_k4, all 8 at once)
__device__ __forceinline__ void unpack_k4(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t& sc03, uint32_t& sc47, uint32_t& m03, uint32_t& m47) {
  sc03 = s0 & 0x3f3f3f3fu;
  m03 = s1 & 0x3f3f3f3fu;
  sc47 = (s2 & 0x0f0f0f0fu) | (((s0 >> 6) & 0x03030303u) << 4);
  m47 = ((s2 >> 4) & 0x0f0f0f0fu) | (((s1 >> 6) & 0x03030303u) << 4);
}
#define CTB_BYTE(w, i) ((int)(((w) >> ((i) * 8)) & 0xffu))

// Each dot_* returns the finished row value (valid in every lane of the 8-lane group).  `l` = lane & 7.

__device__ __forceinline__ float dot_q4k(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* qrow = w.qs + (size_t)row * nb * 128 + l * 16;
  const uint8_t* hrow = w.sc + (size_t)row * nb * 16;
  float acc = 0.f, acc_m = 0.f;
#pragma unroll 4
  for (int b = 0; b < nb; b++) {
    const int4 q = ldg_stream16(qrow + (size_t)b * 128);
    const int4 h = __ldg((const int4*)(hrow + (size_t)b * 16));
    const int4 a0 = *(const int4*)(a.qs + ((b * 2 + 0) * 8 + l) * 16);
    const int4 a1 = *(const int4*)(a.qs + ((b * 2 + 1) * 8 + l) * 16);
    const float yd = a.d[b];
    uint32_t sc03, sc47, m03, m47;
    unpack_k4((uint32_t)h.y, (uint32_t)h.z, (uint32_t)h.w, sc03, sc47, m03, m47);
    int sumi;
    sumi = CTB_BYTE(sc03, 0) * __dp4a((int)((uint32_t)q.x & 0x0f0f0f0fu), a0.x, 0);
    sumi += CTB_BYTE(sc03, 1) * __dp4a((int)(((uint32_t)q.x >> 4) & 0x0f0f0f0fu), a0.y, 0);
    sumi += CTB_BYTE(sc03, 2) * __dp4a((int)((uint32_t)q.y & 0x0f0f0f0fu), a0.z, 0);
    sumi += CTB_BYTE(sc03, 3) * __dp4a((int)(((uint32_t)q.y >> 4) & 0x0f0f0f0fu), a0.w, 0);
    sumi += CTB_BYTE(sc47, 0) * __dp4a((int)((uint32_t)q.z & 0x0f0f0f0fu), a1.x, 0);
    sumi += CTB_BYTE(sc47, 1) * __dp4a((int)(((uint32_t)q.z >> 4) & 0x0f0f0f0fu), a1.y, 0);
    sumi += CTB_BYTE(sc47, 2) * __dp4a((int)((uint32_t)q.w & 0x0f0f0f0fu), a1.z, 0);
    sumi += CTB_BYTE(sc47, 3) * __dp4a((int)(((uint32_t)q.w >> 4) & 0x0f0f0f0fu), a1.w, 0);
    const float dw = h2f((uint16_t)((uint32_t)h.x & 0xffffu));
    const float dm = h2f((uint16_t)((uint32_t)h.x >> 16));
    acc = __fmaf_rn(__fmul_rn(yd, dw), (float)sumi, acc);
    if (l < 4) {   // mins: lane k of acc_m gets m[2k]*(bsums[4k]+bsums[4k+1]) + m[2k+1]*(bsums[4k+2]+bsums[4k+3])
      const int2 bsv = *(const int2*)(a.bs + b * 16 + 4 * l);
      const int s0 = (int)(short)(bsv.x & 0xffff) + (int)(short)((uint32_t)bsv.x >> 16);
      const int s1 = (int)(short)(bsv.y & 0xffff) + (int)(short)((uint32_t)bsv.y >> 16);
      const uint32_t mw = l < 2 ? m03 : m47;
      const int k2 = (l & 1) * 2;
      const int prod = CTB_BYTE(mw, k2) * s0 + CTB_BYTE(mw, k2 + 1) * s1;
      acc_m = __fmaf_rn(__fmul_rn(-yd, dm), (float)prod, acc_m);
    }
  }
  const float hs = group_hsum8(acc);
  const float ms = group_hsum4(acc_m);   // meaningful in lanes l < 4; lane 0 of the group stores
  return __fadd_rn(hs, ms);
}

__device__ __forceinline__ float dot_q5k(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* qrow = w.qs + (size_t)row * nb * 128 + l * 16;
  const uint8_t* hrow = w.sc + (size_t)row * nb * 16;
  const uint8_t* brow = w.qh + (size_t)row * nb * 32 + l * 4;
  float acc = 0.f, summs = 0.f;
#pragma unroll 4
  for (int b = 0; b < nb; b++) {
    const int4 q = ldg_stream16(qrow + (size_t)b * 128);
    const int4 h = __ldg((const int4*)(hrow + (size_t)b * 16));
    const uint32_t hb = (uint32_t)__ldg((const int*)(brow + (size_t)b * 32));   // bit s of byte e: 5th bit of element 4l+e in sub-block s
    const int4 a0 = *(const int4*)(a.qs + ((b * 2 + 0) * 8 + l) * 16);
    const int4 a1 = *(const int4*)(a.qs + ((b * 2 + 1) * 8 + l) * 16);
    const float yd = a.d[b];
    uint32_t sc03, sc47, m03, m47;
    unpack_k4((uint32_t)h.y, (uint32_t)h.z, (uint32_t)h.w, sc03, sc47, m03, m47);
    const uint32_t qv[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
    const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    int sumi = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t lo = (qv[j] & 0x0f0f0f0fu) | (((hb >> (2 * j)) & 0x01010101u) << 4);
      const uint32_t hi = ((qv[j] >> 4) & 0x0f0f0f0fu) | (((hb >> (2 * j + 1)) & 0x01010101u) << 4);
      const uint32_t scw = j < 2 ? sc03 : sc47;
      sumi += CTB_BYTE(scw, (2 * j) & 3) * __dp4a((int)lo, av[2 * j], 0);
      sumi += CTB_BYTE(scw, (2 * j + 1) & 3) * __dp4a((int)hi, av[2 * j + 1], 0);
    }
    const float dw = h2f((uint16_t)((uint32_t)h.x & 0xffffu));
    const float dm = h2f((uint16_t)((uint32_t)h.x >> 16));
    acc = __fmaf_rn(__fmul_rn(yd, dw), (float)sumi, acc);
    if (l == 0) {   // scalar mins chain of the AVX2 kernel: summs = fma(dmin, hsum, summs) (fused in the reference binary)
      int hsum = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int s = (int)a.bs[b * 16 + 2 * k] + (int)a.bs[b * 16 + 2 * k + 1];
        hsum += CTB_BYTE(k < 4 ? m03 : m47, k & 3) * s;
      }
      summs = __fmaf_rn(__fmul_rn(-yd, dm), (float)hsum, summs);
    }
  }
  return __fadd_rn(group_hsum8(acc), summs);   // lane 0 of the group holds summs and stores
}

__device__ __forceinline__ float dot_q6k(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* lrow = w.qs + (size_t)row * nb * 128 + l * 16;
  const uint8_t* hrow = w.qh + (size_t)row * nb * 64 + l * 8;
  const uint8_t* srow = w.sc + (size_t)row * nb * 16;
  const uint16_t* drow = w.d + (size_t)row * nb;
  const int hi16 = l >> 2;   // elements 0..15 of a 32-group use the even scale, 16..31 the odd one
  float acc = 0.f;
#pragma unroll 4
  for (int b = 0; b < nb; b++) {
    const int4 ql = ldg_stream16(lrow + (size_t)b * 128);     // words: (jj=0,v=0) (0,1) (1,0) (1,1)
    const int2 qh = ldg_stream8(hrow + (size_t)b * 64);       // words: jj=0, jj=1
    const int4 scv = __ldg((const int4*)(srow + (size_t)b * 16));
    const float dw = h2f(__ldg(drow + b));
    const int4 a0 = *(const int4*)(a.qs + ((b * 2 + 0) * 8 + l) * 16);
    const int4 a1 = *(const int4*)(a.qs + ((b * 2 + 1) * 8 + l) * 16);
    const float yd = a.d[b];
    const uint32_t A[2] = {(uint32_t)ql.x, (uint32_t)ql.z}, B[2] = {(uint32_t)ql.y, (uint32_t)ql.w}, H[2] = {(uint32_t)qh.x, (uint32_t)qh.y};
    const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const uint32_t scw[4] = {(uint32_t)scv.x, (uint32_t)scv.y, (uint32_t)scv.z, (uint32_t)scv.w};
    int sumi = 0;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const uint32_t u0 = (A[jj] & 0x0f0f0f0fu) | ((H[jj] << 4) & 0x30303030u);
      const uint32_t u1 = (B[jj] & 0x0f0f0f0fu) | ((H[jj] << 2) & 0x30303030u);
      const uint32_t u2 = ((A[jj] >> 4) & 0x0f0f0f0fu) | (H[jj] & 0x30303030u);
      const uint32_t u3 = ((B[jj] >> 4) & 0x0f0f0f0fu) | ((H[jj] >> 2) & 0x30303030u);
      const uint32_t uu[4] = {u0, u1, u2, u3};
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const int aw = av[jj * 4 + m];
        // (q6 - 32)·q8 over 4 elements = u·q8 - 32·Σq8   (the AVX2 kernel does the same with maddubs(m32s, q8))
        const int s = __dp4a((int)uu[m], aw, 0) - 32 * __dp4a(0x01010101, aw, 0);
        const int sidx = 8 * jj + 2 * m + hi16;   // int8 scale of this 16-element sub-block
        const int scale = (int)(int8_t)CTB_BYTE(scw[sidx >> 2], sidx & 3);
        sumi += scale * s;
      }
    }
    acc = __fmaf_rn(__fmul_rn(yd, dw), (float)sumi, acc);
  }
  return group_hsum8(acc);
}

// Q4_0: natural plane; lane l uses word (l & 3) of the block's 16 nibble bytes, low nibbles for l < 4 (elements 4l..4l+3),
// high nibbles for l >= 4 (elements 16+4(l-4)..).  bytes_from_nibbles_32 - 8, then the s8·s8 dot (ggml.c:2500-2525).
__device__ __forceinline__ float dot_q40(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* qrow = w.qs + (size_t)row * nb * 16 + (l & 3) * 4;
  const uint16_t* drow = w.d + (size_t)row * nb;
  const int shift = (l >> 2) * 4;
  float acc = 0.f;
#pragma unroll 8
  for (int b = 0; b < nb; b++) {
    const uint32_t q = (uint32_t)__ldg((const int*)(qrow + (size_t)b * 16));
    const float dw = h2f(__ldg(drow + b));
    const int aw = *(const int*)(a.qs + b * 32 + l * 4);
    const float yd = a.d[b];
    const uint32_t nib = (q >> shift) & 0x0f0f0f0fu;
    const uint32_t bx = ((nib | 0x80808080u) - 0x08080808u) ^ 0x80808080u;   // per-byte (nib - 8), two's complement, no borrow
    acc = __fmaf_rn(__fmul_rn(dw, yd), (float)__dp4a((int)bx, aw, 0), acc);
  }
  return group_hsum8(acc);
}

__device__ __forceinline__ float dot_q80(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* qrow = w.qs + (size_t)row * nb * 32 + l * 4;
  const uint16_t* drow = w.d + (size_t)row * nb;
  float acc = 0.f;
#pragma unroll 8
  for (int b = 0; b < nb; b++) {
    const int q = __ldg((const int*)(qrow + (size_t)b * 32));
    const float dw = h2f(__ldg(drow + b));
    const int aw = *(const int*)(a.qs + b * 32 + l * 4);
    acc = __fmaf_rn(__fmul_rn(dw, a.d[b]), (float)__dp4a(q, aw, 0), acc);
  }
  return group_hsum8(acc);
}

__device__ __forceinline__ float dot_quant(const DevMat& w, int row, const ActView& a, int l) {
  switch (w.type) {
    case GT_Q4_K: return dot_q4k(w, row, a, l);
    case GT_Q6_K: return dot_q6k(w, row, a, l);
    case GT_Q5_K: return dot_q5k(w, row, a, l);
    case GT_Q4_0: return dot_q40(w, row, a, l);
    default: return dot_q80(w, row, a, l);
  }
}

// GGML_F32x8_REDUCE over a warp that plays 4 accumulators x 8 lanes (lane = 8*j + l): (0+2),(1+3) -> (0+1) -> lo128+hi128 ->
// hadd -> hadd  (ggml.c:1964-1982).  All lanes end with the result.
__device__ __forceinline__ float warp_reduce_f32x8(float v) {
  v = v + __shfl_xor_sync(0xffffffffu, v, 16);
  v = v + __shfl_xor_sync(0xffffffffu, v, 8);
  v = v + __shfl_xor_sync(0xffffffffu, v, 4);
  v = v + __shfl_xor_sync(0xffffffffu, v, 1);
  v = v + __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// F16 weights: ggml_vec_dot_f16 (ggml.c:2392-2426), one warp per row: lane L owns elements 32i+L, one fma per step,
// the reduce above, leftovers (K % 32) added in double.  x has been rounded to f16 in the prologue (ggml.c:11141-11157).
__device__ __forceinline__ float dot_f16_row(const DevMat& w, int row, const uint8_t* smem, int lane) {
  const uint16_t* xa = (const uint16_t*)smem;
  const uint16_t* wr = (const uint16_t*)w.qs + (size_t)row * w.K;
  const int np = w.K & ~31;
  float s = 0.f;
  for (int i = lane; i < np; i += 32) s = __fmaf_rn(h2f(__ldg(wr + i)), h2f(xa[i]), s);
  double sumf = (double)warp_reduce_f32x8(s);
  for (int i = np; i < w.K; i++) sumf += (double)__fmul_rn(h2f(__ldg(wr + i)), h2f(xa[i]));
  return (float)sumf;
}
// F32 weights: ggml_vec_dot_f32 (ggml.c:2330-2365) has the same 4x8-lane shape.
__device__ __forceinline__ float dot_f32_row(const DevMat& w, int row, const uint8_t* smem, int lane) {
  const float* xa = (const float*)smem;
  const float* wr = (const float*)w.qs + (size_t)row * w.K;
  const int np = w.K & ~31;
  float s = 0.f;
  for (int i = lane; i < np; i += 32) s = __fmaf_rn(__ldg(wr + i), xa[i], s);
  float sumf = warp_reduce_f32x8(s);
  for (int i = np; i < w.K; i++) sumf = __fmaf_rn(__ldg(wr + i), xa[i], sumf);
  return sumf;
}

__device__ __forceinline__ float table_f16(const uint16_t* tab, float x) { return h2f(__ldg(tab + f2h(x))); }

__device__ __forceinline__ void store_epilogue(const MVSeg& sg, const MVParams& p, int row, float v) {
  if (sg.epi == EPI_ADD) v = __fadd_rn(v, sg.res[row]);
  else if (sg.epi == EPI_ADD2) v = __fadd_rn(__fadd_rn(v, sg.res[row]), sg.res2[row]);
  else if (sg.epi == EPI_GELU) v = table_f16(p.gelu_tab, v);
  sg.out[row] = v;
}

// rows one work unit (one warp-iteration) covers for a weight type
__host__ __device__ inline int rows_per_unit(int type) { return (type == GT_F16 || type == GT_F32) ? 1 : MV_ROWS; }

// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(MV_THREADS) k_matvec(const __grid_constant__ MVParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  stage_activation(p.x, p.norm_w, p.norm_b, p.norm_out, p.norm_mode, p.eps, p.K, p.act, smem, red, blockIdx.x == 0);
  const ActView a = act_view(p.act, p.K, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l = lane & 7, g = lane >> 3;
  const int gw = blockIdx.x * MV_WARPS + warp, nw = gridDim.x * MV_WARPS;

  if (p.pair_silu) {
    const DevMat& gm = p.seg[0].w;
    const DevMat& um = p.seg[1].w;
    const int units = (gm.M + MV_ROWS - 1) / MV_ROWS;
    for (int un = gw; un < units; un += nw) {
      const int row = un * MV_ROWS + g, rc = min(row, gm.M - 1);
      const float vg = dot_quant(gm, rc, a, l);
      const float vu = dot_quant(um, rc, a, l);
      if (l == 0 && row < gm.M) p.seg[0].out[row] = __fmul_rn(table_f16(p.silu_tab, vg), vu);
    }
    return;
  }

  int unit_base = 0;
  for (int s = 0; s < p.nseg; s++) {
    const MVSeg& sg = p.seg[s];
    const int rpu = rows_per_unit(sg.w.type);
    const int units = (sg.w.M + rpu - 1) / rpu;
    int first = gw - (unit_base % nw);   // continue the global striding across segments so all warps stay busy
    if (first < 0) first += nw;
    for (int un = first; un < units; un += nw) {
      if (rpu == 1) {
        const float v = sg.w.type == GT_F16 ? dot_f16_row(sg.w, un, smem, lane) : dot_f32_row(sg.w, un, smem, lane);
        if (lane == 0) store_epilogue(sg, p, un, v);
      } else {
        const int row = un * MV_ROWS + g;
        const float v = dot_quant(sg.w, min(row, sg.w.M - 1), a, l);
        if (l == 0 && row < sg.w.M) store_epilogue(sg, p, row, v);
      }
    }
    unit_base += units;
  }
}

}  // namespace ctb
