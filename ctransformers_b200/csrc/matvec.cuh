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
//     folded in with ONE fmadd per block, blocks in order, and ends with hsum_float_8.  K-quants: 4 GPU lanes share a weight
//     row (lane u plays AVX lanes u and u+4), a warp carries an 8-row tile, dp4a does the 4-element integer dots, dp2a folds
//     the 6-bit scales; Q4_0 / Q8_0: 8 lanes per row, 4 rows per warp.  The per-block fmaf chain and the final shuffle tree
//     reproduce the float order — results equal the reference's bit for bit:
//     Q4_K k_quants.c:2651-2714 · Q5_K 3174-3262 · Q6_K 3794-3872 · Q4_0 ggml.c:2500-2525 · Q8_0 3379-3402 · F16 2392-2426
//   scheduling (K-quants) — see k_matvec: contiguous equal block ranges per warp, rows cut between warps are folded in order
//     by handing the fp32 state from warp to warp through shared memory
//   epilogue
//     store | + residual (ggml_add, llama.cpp:2415, 2453) | SiLU table of the gate rows (ggml.c:3625-3632; the product with
//     up is formed where ffn_down stages its input, llama.cpp:2438-2443) | GELU table (falcon, ggml.c:3568-3575)
//
// Why exactness matters: the next mat-mul re-quantizes this output to int8; a 1-ulp difference can flip one rounding and the
// logits then differ by ~1e-3 (measured).  HBM traffic per launch = the weight planes once + O(K) activations from L2.
#pragma once
#include "device_types.cuh"
#include "attention.cuh"

namespace ctb {

#ifndef CTB_PF
#define CTB_PF 2     // L2 prefetch policy: 0 none, 1 the warp's whole range before the prologue, 2 rolling, CTB_PFD blocks ahead (measured best)
#define CTB_PFD 4
#endif
#ifndef CTB_LPR
#define CTB_LPR 4     // K-quants: GPU lanes per weight row (each plays 8/CTB_LPR of the reference kernel's 8 AVX lanes)
#endif
#ifndef CTB_THREADS
#define CTB_THREADS 512
#endif
#ifndef CTB_CTAS_PER_SM
#define CTB_CTAS_PER_SM 1
#endif
constexpr int MV_THREADS = CTB_THREADS;    // one persistent CTA per SM (<= 128 registers per thread): the activation prologue is paid once per SM
constexpr int MV_WARPS = MV_THREADS / 32;
constexpr int MV_ROWS = 4;   // Q4_0 / Q8_0: rows per warp (one per 8-lane group, in-lane chain)
constexpr int MV_MAX_SEG = 3;

enum : int { NORM_NONE = 0, NORM_RMS = 1, NORM_LAYER = 2 };
enum : int { EPI_STORE = 0, EPI_ADD = 1, EPI_GELU = 2, EPI_ADD2 = 3, EPI_SILU = 4 };   // GELU / SILU: the reference's fp16-table activation of the row value

struct MVSeg {
  DevMat w;
  float* out;          // [M]
  const float* res;    // EPI_ADD: out = acc + res; EPI_ADD2: out = (acc + res) + res2
  const float* res2;
  int epi;
};

struct MVParams {
  const float* x;        // [K] f32 input
  const float* x2;       // x_mode 1: second operand
  int def_max;           // set by matvec_launch_shape: blocks of parked terms per warp that fit in shared memory
  int x_mode;            // 0: x;  1: x * x2 (ggml_mul of silu(gate) and up, llama.cpp:2438-2443; the SiLU table is applied by the gate rows' epilogue)
  const float* norm_w;   // [K] or null
  const float* norm_b;   // [K] or null (LayerNorm bias)
  float* norm_out;       // optional [K]: CTA 0 writes the normalised vector (result_norm / embeddings)
  float eps;
  int norm_mode;
  int K;
  int act;               // ACT_*
  int nseg;
  MVSeg seg[MV_MAX_SEG];
  const uint16_t* silu_tab;   // 65536-entry fp16 tables built on the host exactly like ggml.c:4319-4333
  const uint16_t* gelu_tab;
  // fused attention tail (QKV launches of the decode step): after its share of the mat-vec a CTA runs attention task(s) for
  // the token — see k_matvec.  attn_counter counts finished row tiles of this launch (zeroed before the step).
  int attn_on;
  AttnParams attn;
  int* attn_counter;
  unsigned long long* trace;   // optional, per CTA 4 + MV_WARPS globaltimer stamps: entry, dependency released, input staged, (unused), each warp's end
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

// Whole prologue: normalise + quantize x[K] into shared memory.  All MV_THREADS threads must call.
// Each thread owns 16 consecutive elements per pass (one bsums group; 16 threads = one Q8_K block, 2 threads = one Q8_0
// block); all global loads of a pass are issued before anything depends on them.
__device__ __forceinline__ void load16(const float* p, int valid, float (&v)[16]) {
  if (valid >= 16) {
    const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)p + 1), c = __ldg((const float4*)p + 2), d = __ldg((const float4*)p + 3);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
  } else {
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = e < valid ? __ldg(p + e) : 0.f;
  }
}

__device__ __forceinline__ double block_sum_f64(double s, double* red /* [MV_WARPS] smem */) {
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < MV_WARPS; w++) t += red[w];
  return t;
}

__device__ __forceinline__ uint32_t pack4(const int* q) {
  return (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
}

// 16 consecutive input elements with the producer's activation applied: the reference's fp16-table SiLU (ggml.c:3625-3632)
// times the up projection, or the fp16-table GELU (ggml.c:3568-3575) — fused here instead of in the producing kernel so that
// gate and up can be two independent row sets there.
__device__ __forceinline__ void load16x(const MVParams& xs, int base, int valid, float (&v)[16]) {
  load16(xs.x + base, valid, v);
  if (xs.x_mode == 1) {          // x = silu_table(gate) (applied once, where the gate row was produced); input = x * up (ggml_mul)
    float u[16];
    load16(xs.x2 + base, valid, u);
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = __fmul_rn(v[e], u[e]);
  }
}

// Norm weight / bias of this thread's first 16 elements: constants of the model, so they are fetched before the kernel waits
// for its predecessor (pdl_wait) and are in registers when the input vector arrives.
struct NormPre { float w0[16], bias0[16]; };
__device__ __forceinline__ void preload_norm(NormPre& np, const float* nw, const float* nb_, int norm_mode, int K) {
  const int t = threadIdx.x;
  if (norm_mode != NORM_NONE && nw) load16(nw + t * 16, K - t * 16, np.w0);
  if (norm_mode != NORM_NONE && nb_) load16(nb_ + t * 16, K - t * 16, np.bias0);
}

__device__ __forceinline__ void stage_activation(const MVParams& xs, const NormPre& np, const float* nw, const float* nb_, float* norm_out, int norm_mode, float eps,
                                                  int K, int act, uint8_t* smem, double* red, bool write_norm, unsigned long long* t_stats = nullptr) {
  const int t = threadIdx.x, lane = t & 31;
  const int passes = (K + MV_THREADS * 16 - 1) / (MV_THREADS * 16);
  // ---- statistics (fp64 sums like ggml.c:10700-10703 / 10630-10645; the order of a double sum does not reach the float result)
  float mean = 0.f, scale = 1.f;
  float v0[16];                          // pass 0's x stays in registers
  load16x(xs, t * 16, K - t * 16, v0);
  if (norm_mode == NORM_RMS) {
    double ss = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * MV_THREADS + t) * 16;
      if ((base & ~511) >= K) continue;     // the whole warp lies past the end of x (warp-uniform): nothing to add
      float v[16];
      if (ps == 0) {
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = v0[e];
      } else load16x(xs, base, K - base, v);
#pragma unroll
      for (int e = 0; e < 16; e++) ss += (double)__fmul_rn(v[e], v[e]);
    }
    ss = block_sum_f64(ss, red);
    const float m = (float)(ss / (double)K);
    scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(m, eps)));
  } else if (norm_mode == NORM_LAYER) {
    double s1 = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * MV_THREADS + t) * 16;
      if ((base & ~511) >= K) continue;
      float v[16];
      load16x(xs, base, K - base, v);
#pragma unroll
      for (int e = 0; e < 16; e++) s1 += (double)v[e];
    }
    s1 = block_sum_f64(s1, red);
    mean = (float)(s1 / (double)K);
    double s2 = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * MV_THREADS + t) * 16;
      if ((base & ~511) >= K) continue;
      float v[16];
      load16x(xs, base, K - base, v);
#pragma unroll
      for (int e = 0; e < 16; e++) { const float d = (base + e < K) ? __fsub_rn(v[e], mean) : 0.f; s2 += (double)__fmul_rn(d, d); }
    }
    s2 = block_sum_f64(s2, red);
    const float var = (float)(s2 / (double)K);
    scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
  }
  if (t_stats && t == 0) *t_stats = globaltimer_ns();
  // ---- normalise + quantize
  int8_t* qs = (int8_t*)smem;
  const size_t off = ((size_t)K + 15) & ~(size_t)15;
  float* dd = (float*)(smem + off);
  int16_t* bs = (int16_t*)(smem + off + q8k_d_bytes(K));
  for (int ps = 0; ps < passes; ps++) {
    const int base = (ps * MV_THREADS + t) * 16;
    const int valid = K - base;
    if ((base & ~511) >= K) continue;       // warp-uniform: this warp has no elements in this pass
    float v[16];
    if (ps == 0) {
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = v0[e];
    } else load16x(xs, base, valid, v);
    if (norm_mode != NORM_NONE) {
      float w[16], bb[16];
      if (ps == 0) {
#pragma unroll
        for (int e = 0; e < 16; e++) { w[e] = np.w0[e]; bb[e] = np.bias0[e]; }
      } else {
        if (nw) load16(nw + base, valid, w);
        if (nb_) load16(nb_ + base, valid, bb);
      }
#pragma unroll
      for (int e = 0; e < 16; e++) {
        float y = v[e];
        if (norm_mode == NORM_LAYER) y = __fsub_rn(y, mean);
        y = __fmul_rn(y, scale);
        if (nw) y = __fmul_rn(y, w[e]);
        if (nb_) y = __fadd_rn(y, bb[e]);
        v[e] = y;
      }
    }
    if (write_norm && norm_out && valid > 0) {
#pragma unroll
      for (int e = 0; e < 16; e++) if (e < valid) norm_out[base + e] = v[e];
    }
    if (act == ACT_Q8_K) {
      // reference quantize_row_q8_K_reference (k_quants.c:1191-1226): the first element with the largest |x| fixes the sign;
      // iscale = -128/max; q = min(127, nearest_int(iscale*x)); the reference BINARY fuses iscale*x + 12582912.f (vfmadd), so
      // the exact product is rounded once — __fmaf_rn.  d = 1/iscale; bsums per 16.  16 lanes (a half-warp) share a block.
      float amax = 0.f, mx = 0.f;
#pragma unroll
      for (int e = 0; e < 16; e++) { const float ax = fabsf(v[e]); if (ax > amax) { amax = ax; mx = v[e]; } }
      float gmax = amax;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
      const unsigned who = __ballot_sync(0xffffffffu, amax == gmax);
      const int hbase = lane & 16, hl = lane & 15;
      const unsigned mine = (who >> hbase) & 0xffffu;
      const float maxv = __shfl_sync(0xffffffffu, mx, hbase + __ffs(mine) - 1);
      if (valid > 0) {
        const int b = base >> 8;
        int q[16];
        int sum = 0;
        if (gmax == 0.f) {
#pragma unroll
          for (int e = 0; e < 16; e++) q[e] = 0;
          if (hl == 0) dd[b] = 0.f;
        } else {
          const float iscale = __fdiv_rn(-128.f, maxv);
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const float val = __fmaf_rn(iscale, v[e], 12582912.f);
            q[e] = min(127, (__float_as_int(val) & 0x007fffff) - 0x00400000);
            sum += q[e];
          }
          if (hl == 0) dd[b] = __fdiv_rn(1.f, iscale);
        }
        bs[b * 16 + hl] = (int16_t)sum;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int wi = hl * 4 + j;
          *(uint32_t*)(qs + q8k_word_offset(b, wi >> 3, wi & 7)) = pack4(q + 4 * j);
        }
      }
    } else if (act == ACT_Q8_0) {
      // quantize_row_q8_0, AVX2 variant (ggml.c:1232-1268): d = amax/127 kept as fp16, id = 127/amax, round-half-even
      float amax = 0.f;
#pragma unroll
      for (int e = 0; e < 16; e++) amax = fmaxf(amax, fabsf(v[e]));
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
      if (valid > 0) {
        const float d = __fdiv_rn(amax, 127.f);
        const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
        int q[16];
#pragma unroll
        for (int e = 0; e < 16; e++) q[e] = __float2int_rn(__fmul_rn(v[e], id));
        *(uint4*)(qs + base) = make_uint4(pack4(q), pack4(q + 4), pack4(q + 8), pack4(q + 12));
        if ((lane & 1) == 0) dd[base >> 5] = h2f(f2h(d));
      }
    } else if (act == ACT_F16) {
      uint16_t* h = (uint16_t*)smem;
#pragma unroll
      for (int e = 0; e < 16; e++) if (e < valid) h[base + e] = f2h(v[e]);
    } else {
      float* f = (float*)smem;
#pragma unroll
      for (int e = 0; e < 16; e++) if (e < valid) f[base + e] = v[e];
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

// ---- K-quants.  A warp task is a tile of 8 consecutive rows.  Lane (g = lane>>2, t = lane&3) owns row g of the tile and plays
// AVX lanes l = t and l = t+4 of the reference kernel for ALL blocks of that row, in order.  Per block it fetches the two
// 16-byte pieces of the lane-major qs plane that hold words l (the 4 lanes of a row read its 128-byte block as two full
// 64-byte segments), computes what int32 lanes l hold in the AVX2 kernel — sumi(b,l) = Σ_sub-blocks scale · dp4a(4 weights,
// 4 activations), the scales applied two at a time with dp2a on int16 pairs — and folds (float)sumi into its private fp32
// accumulator with one fmaf per block.  No shared-memory staging of weights, no synchronisation; the fold order is the
// reference's by construction, and hsum_float_8 starts in-lane (a[t] + a[t+4]) and ends with two shuffles.
__device__ __forceinline__ int pack16(int lo, int hi) { return (int)__byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }

// Σ_s scale_s · dp_s for the 8 sub-block partial dots of one AVX lane; sc03 / sc47 hold the 8 scale bytes
__device__ __forceinline__ int scale_fold(const int (&dp)[8], uint32_t sc03, uint32_t sc47) {
  int s = __dp2a_lo(pack16(dp[0], dp[1]), (int)sc03, 0);
  s = __dp2a_hi(pack16(dp[2], dp[3]), (int)sc03, s);
  s = __dp2a_lo(pack16(dp[4], dp[5]), (int)sc47, s);
  s = __dp2a_hi(pack16(dp[6], dp[7]), (int)sc47, s);
  return s;
}

constexpr int KQ_LPR = CTB_LPR;            // lanes per row
constexpr int KQ_NA = 8 / KQ_LPR;          // AVX lanes per GPU lane: lane u of a row plays l = u + KQ_LPR*e, e = 0..KQ_NA-1
constexpr int KQ_NM = KQ_LPR >= 4 ? 1 : 4 / KQ_LPR;   // Q4_K mins lanes per GPU lane: k = u + KQ_LPR*i
static_assert(KQ_LPR == 4 || KQ_LPR == 2, "lanes per row: 4 or 2");

// activation words of AVX lane l of block b: sub-blocks 0..3 and 4..7
__device__ __forceinline__ void load_act_lane(const ActView& a, int b, int l, int (&av)[8]) {
  const int8_t* base = a.qs + (size_t)b * 256 + l * 16;
  const int4 lo = *(const int4*)base, hi = *(const int4*)(base + 128);
  av[0] = lo.x; av[1] = lo.y; av[2] = lo.z; av[3] = lo.w; av[4] = hi.x; av[5] = hi.y; av[6] = hi.z; av[7] = hi.w;
}

// What one block contributes to one row, for this lane's AVX lanes: p[e] = (float)sumi of lane l_e, dd = y.d·d, and the mins
// terms pm[i]·ddm (Q4_K: mins lanes k = u + KQ_LPR*i; Q5_K: the scalar term in pm[0] of lane u == 0; Q6_K: none).
struct BlockTerms { float p[KQ_NA]; float pm[KQ_NM]; float dd, ddm; };

// Raw block data of one lane, held in registers by the software pipeline below
struct RawQ4K { int4 c[KQ_NA]; int4 ch; };
struct RawQ5K { int4 c[KQ_NA]; int4 ch; uint32_t hb[KQ_NA]; };
struct RawQ6K { int4 ql[KQ_NA]; int2 qh[KQ_NA]; int4 scv; uint16_t d; };
__device__ __forceinline__ void load_raw(RawQ4K& r, const DevMat& w, size_t blk, int u) {
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) r.c[e] = ldg_stream16(w.qs + blk * 128 + (u + KQ_LPR * e) * 16);
  r.ch = ldg_keep16(w.sc + blk * 16);
}
__device__ __forceinline__ void load_raw(RawQ5K& r, const DevMat& w, size_t blk, int u) {
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) {
    r.c[e] = ldg_stream16(w.qs + blk * 128 + (u + KQ_LPR * e) * 16);
    r.hb[e] = (uint32_t)__ldg((const int*)(w.qh + blk * 32 + (u + KQ_LPR * e) * 4));
  }
  r.ch = ldg_keep16(w.sc + blk * 16);
}
__device__ __forceinline__ void load_raw(RawQ6K& r, const DevMat& w, size_t blk, int u) {
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) {
    r.ql[e] = ldg_stream16(w.qs + blk * 128 + (u + KQ_LPR * e) * 16);
    r.qh[e] = ldg_stream8(w.qh + blk * 64 + (u + KQ_LPR * e) * 8);
  }
  r.scv = ldg_keep16(w.sc + blk * 16);
  r.d = __ldg(w.d + blk);
}

// k_quants.c:2651-2714
__device__ __forceinline__ BlockTerms block_terms(const RawQ4K& raw, int b, const ActView& a, int u) {
  const int4 ch = raw.ch;
  uint32_t sc03, sc47, m03, m47;
  unpack_k4((uint32_t)ch.y, (uint32_t)ch.z, (uint32_t)ch.w, sc03, sc47, m03, m47);
  BlockTerms r;
  const float yd = a.d[b];
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) {
    const int4 q = raw.c[e];
    const uint32_t qv[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
    int av[8];
    load_act_lane(a, b, u + KQ_LPR * e, av);
    int dp[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      dp[2 * j] = __dp4a((int)(qv[j] & 0x0f0f0f0fu), av[2 * j], 0);
      dp[2 * j + 1] = __dp4a((int)((qv[j] >> 4) & 0x0f0f0f0fu), av[2 * j + 1], 0);
    }
    r.p[e] = (float)scale_fold(dp, sc03, sc47);
  }
  r.dd = __fmul_rn(yd, h2f((uint16_t)((uint32_t)ch.x & 0xffffu)));
  // mins lane k: m[2k]*(bsums[4k]+bsums[4k+1]) + m[2k+1]*(bsums[4k+2]+bsums[4k+3])
#pragma unroll
  for (int i = 0; i < KQ_NM; i++) {
    const int k = u + KQ_LPR * i;
    const int2 bsv = *(const int2*)(a.bs + b * 16 + 4 * k);
    const int s0 = (int)(short)(bsv.x & 0xffff) + (int)(short)((uint32_t)bsv.x >> 16);
    const int s1 = (int)(short)(bsv.y & 0xffff) + (int)(short)((uint32_t)bsv.y >> 16);
    const uint32_t mw = (k < 2 ? m03 : m47) >> ((k & 1) * 16);
    r.pm[i] = (float)((int)(mw & 0xffu) * s0 + (int)((mw >> 8) & 0xffu) * s1);
  }
  r.ddm = __fmul_rn(-yd, h2f((uint16_t)((uint32_t)ch.x >> 16)));
  return r;
}

// k_quants.c:3174-3262
__device__ __forceinline__ BlockTerms block_terms(const RawQ5K& raw, int b, const ActView& a, int u) {
  const int4 ch = raw.ch;
  uint32_t sc03, sc47, m03, m47;
  unpack_k4((uint32_t)ch.y, (uint32_t)ch.z, (uint32_t)ch.w, sc03, sc47, m03, m47);
  BlockTerms r;
  const float yd = a.d[b];
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) {
    const int4 q = raw.c[e];
    const uint32_t hb = raw.hb[e];
    const uint32_t qv[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
    int av[8];
    load_act_lane(a, b, u + KQ_LPR * e, av);
    int dp[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {   // bit s of a qh byte: 5th bit of the element in sub-block s
      const uint32_t lo = (qv[j] & 0x0f0f0f0fu) | (((hb >> (2 * j)) & 0x01010101u) << 4);
      const uint32_t hi = ((qv[j] >> 4) & 0x0f0f0f0fu) | (((hb >> (2 * j + 1)) & 0x01010101u) << 4);
      dp[2 * j] = __dp4a((int)lo, av[2 * j], 0);
      dp[2 * j + 1] = __dp4a((int)hi, av[2 * j + 1], 0);
    }
    r.p[e] = (float)scale_fold(dp, sc03, sc47);
  }
  r.dd = __fmul_rn(yd, h2f((uint16_t)((uint32_t)ch.x & 0xffffu)));
  int hsum = 0;   // scalar mins term of the AVX2 kernel: Σ_k m[k]·(bsums[2k]+bsums[2k+1]) — used by lane u == 0 only
  if (u == 0) {
#pragma unroll
    for (int k = 0; k < 8; k++) hsum += CTB_BYTE(k < 4 ? m03 : m47, k & 3) * ((int)a.bs[b * 16 + 2 * k] + (int)a.bs[b * 16 + 2 * k + 1]);
  }
#pragma unroll
  for (int i = 0; i < KQ_NM; i++) r.pm[i] = i == 0 ? (float)hsum : 0.f;
  r.ddm = __fmul_rn(-yd, h2f((uint16_t)((uint32_t)ch.x >> 16)));
  return r;
}

// k_quants.c:3794-3872
__device__ __forceinline__ BlockTerms block_terms(const RawQ6K& raw, int b, const ActView& a, int u) {
  const int4 scv = raw.scv;
  const float dw = h2f(raw.d);
  const uint32_t scw[4] = {(uint32_t)scv.x, (uint32_t)scv.y, (uint32_t)scv.z, (uint32_t)scv.w};
  BlockTerms r;
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) {   // AVX lane l: l < 4 = elements 0..15 of each 32-group (even scales), l >= 4 = elements 16..31 (odd scales)
    const int l = u + KQ_LPR * e, par = l >> 2;
    const int4 ql = raw.ql[e];       // words: (jj=0,v=0) (0,1) (1,0) (1,1)
    const int2 qh = raw.qh[e];       // words: jj=0, jj=1
    const uint32_t A[2] = {(uint32_t)ql.x, (uint32_t)ql.z}, B[2] = {(uint32_t)ql.y, (uint32_t)ql.w}, H[2] = {(uint32_t)qh.x, (uint32_t)qh.y};
    int av[8];
    load_act_lane(a, b, l, av);
    int sumi = 0;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const uint32_t uu[4] = {(A[jj] & 0x0f0f0f0fu) | ((H[jj] << 4) & 0x30303030u), (B[jj] & 0x0f0f0f0fu) | ((H[jj] << 2) & 0x30303030u),
                              ((A[jj] >> 4) & 0x0f0f0f0fu) | (H[jj] & 0x30303030u), ((B[jj] >> 4) & 0x0f0f0f0fu) | ((H[jj] >> 2) & 0x30303030u)};
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const int aw = av[jj * 4 + m];
        const int sidx = 8 * jj + 2 * m + par;   // int8 scale of this 16-element sub-block
        const int scale = (int)(int8_t)CTB_BYTE(scw[sidx >> 2], sidx & 3);
        // (q6 - 32)·q8 = u·q8 - 32·Σq8, as the AVX2 kernel does with maddubs(m32s, q8)
        sumi += scale * (__dp4a((int)uu[m], aw, 0) - 32 * __dp4a(0x01010101, aw, 0));
      }
    }
    r.p[e] = (float)sumi;
  }
  r.dd = __fmul_rn(a.d[b], dw); r.ddm = 0.f;
#pragma unroll
  for (int i = 0; i < KQ_NM; i++) r.pm[i] = 0.f;
  return r;
}

// running state of one row's fold in this lane
struct Fold { float a[KQ_NA]; float am[KQ_NM]; };
constexpr int KQ_FOLD_FLOATS = KQ_NA + KQ_NM;
__device__ __forceinline__ void fold_zero(Fold& f) {
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) f.a[e] = 0.f;
#pragma unroll
  for (int i = 0; i < KQ_NM; i++) f.am[i] = 0.f;
}
__device__ __forceinline__ void fold_block(Fold& f, const BlockTerms& x) {
#pragma unroll
  for (int e = 0; e < KQ_NA; e++) f.a[e] = __fmaf_rn(x.dd, x.p[e], f.a[e]);
#pragma unroll
  for (int i = 0; i < KQ_NM; i++) f.am[i] = __fmaf_rn(x.ddm, x.pm[i], f.am[i]);
}
// hsum_float_8 (ggml.c:609-615): res[l] = x[l+4] + x[l]; res[0]+res[2], res[1]+res[3]; then their sum — plus the mins tail.
// The finished row value ends up in every lane of the row's group.
__device__ __forceinline__ float fold_finish(int type, const Fold& f) {
  float r, m;
  if (KQ_LPR == 4) {
    r = __fadd_rn(f.a[1], f.a[0]);                          // lane t holds AVX lanes t and t+4
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  } else {
    // lane u holds AVX lanes u, u+2, u+4, u+6: (x[u+4]+x[u]) + (x[u+6]+x[u+2]) is res[0]+res[2] (u = 0) or res[1]+res[3] (u = 1)
    r = __fadd_rn(__fadd_rn(f.a[KQ_NA / 2], f.a[0]), __fadd_rn(f.a[KQ_NA / 2 + 1 < KQ_NA ? KQ_NA / 2 + 1 : 0], f.a[1 < KQ_NA ? 1 : 0]));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  }
  if (type == GT_Q4_K) {                                    // acc_m: (m0+m2) + (m1+m3)
    if (KQ_LPR == 4) {
      m = f.am[0];
      m = __fadd_rn(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = __fadd_rn(m, __shfl_xor_sync(0xffffffffu, m, 1));
    } else {
      m = __fadd_rn(f.am[0], f.am[KQ_NM - 1]);
      m = __fadd_rn(m, __shfl_xor_sync(0xffffffffu, m, 1));
    }
  } else if (type == GT_Q5_K) {
    m = __shfl_sync(0xffffffffu, f.am[0], 0, KQ_LPR);       // the scalar mins chain lives in lane u == 0
  } else {
    return r;
  }
  return __fadd_rn(r, m);
}

// Hand-off of a row tile's fold state between consecutive warps of a CTA (see k_matvec): warp w receives at most one state
// (for the tile its range starts in the middle of) and posts at most one (for the tile its range ends in the middle of).
constexpr int MV_SMEM_LIMIT = 227 * 1024 / CTB_CTAS_PER_SM - (CTB_CTAS_PER_SM > 1 ? 1024 : 0) - ((MV_WARPS + 1) * (KQ_FOLD_FLOATS * 128 + 4) + MV_WARPS * 8 + 256);   // dynamic shared memory a launch may ask for: 227 KB per CTA minus the static part
#ifndef CTB_DEF_MAX
#define CTB_DEF_MAX 20
#endif
constexpr int MV_DEF_MAX = CTB_DEF_MAX;   // most blocks of a mid-row segment whose terms are parked before the state arrives
#ifndef CTB_RING
#define CTB_RING 2
#endif
constexpr int MV_RING = CTB_RING;     // blocks per lane in flight in the register pipeline (measured: 2 > 3 > 1 > 4 once L2 is prefetched)   // most blocks of a mid-row segment whose terms are parked before the state arrives
constexpr int KQ_PARK_F4 = (KQ_NA + KQ_NM + 1 + 3) / 4;   // float4s per lane and parked block: p[], pm[], and dd or ddm
constexpr int KQ_PARK_BYTES = KQ_PARK_F4 * 16 * 32;       // per warp and parked block
struct Chain {
  float4* buf;                   // warp-private [def_max][KQ_PARK_F4][32] parked block terms
  int def_max;
  volatile float* mail_out;      // [KQ_FOLD_FLOATS][32] floats, the NEXT warp's mailbox
  volatile int* flag_out;
  volatile float* mail_in;       // this warp's mailbox
  volatile int* flag_in;
};

// Ask the memory system for blocks [b0, b1) of this lane's row right away (L2 prefetch, no registers held): the 4 lanes of a
// row take turns over its 128-byte lines.  The register pipeline below then finds its data in L2.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_row(const DevMat& w, size_t rb, int b0, int b1, int u) {
  for (int b = b0 + u; b < b1; b += KQ_LPR) {
    prefetch_l2(w.qs + (rb + b) * 128);
    if (w.type == GT_Q6_K && (b & 1) == 0) prefetch_l2(w.qh + (rb + b) * 64);
    if (w.type == GT_Q5_K && (b & 3) == 0) prefetch_l2(w.qh + (rb + b) * 32);
    if ((b & 7) < KQ_LPR || b - u == b0) prefetch_l2(w.sc + (rb + b) * 16);
  }
}

#define CTB_PIN() asm volatile("" ::: "memory")
template <typename Raw, typename Sink>
__device__ __forceinline__ void stream_blocks(const DevMat& w, size_t rb, int b0, int b1, const ActView& a, int lane, Sink sink) {
  constexpr int D = MV_RING;
  const int t = lane % KQ_LPR;   // which of the row's lanes this is
  if (b0 >= b1) return;
  Raw ring[D];
  const int last = b1 - 1;
#pragma unroll
  for (int i = 0; i < D; i++) load_raw(ring[i], w, rb + min(b0 + i, last), t);   // tail slots re-load the last block (a cache hit)
  for (int b = b0; b < b1; b += D) {
#pragma unroll
    for (int i = 0; i < D; i++) {
      CTB_PIN();
#if CTB_PF == 2
      if (b + i + CTB_PFD < b1 && ((b + i) % KQ_LPR) == t) {   // rolling L2 prefetch: the row's block CTB_PFD ahead, one lane of the row per block
        prefetch_l2(w.qs + (rb + b + i + CTB_PFD) * 128);
        if (((b + i) & 7) < 4) prefetch_l2(w.sc + (rb + b + i + CTB_PFD) * 16);
      }
#endif
#if defined(CTB_EXP_LOADS_ONLY)
      if (b + i < b1) { BlockTerms z{}; const unsigned* rw = (const unsigned*)&ring[i]; unsigned acc = 0;
#pragma unroll
        for (unsigned k = 0; k < sizeof(Raw) / 4; k++) acc ^= rw[k];
        z.dd = __uint_as_float(acc & 0x3fffffu); sink(b + i, z); }
#elif defined(CTB_EXP_COMPUTE_X2)
      if (b + i < b1) { BlockTerms z = block_terms(ring[i], b + i, a, t); CTB_PIN(); BlockTerms z2 = block_terms(ring[i], (b + i) ^ 1 < b1 ? (b + i) ^ 1 : b + i, a, t); z.dd = __fadd_rn(z.dd, __fmul_rn(z2.dd, 1e-30f)); z.p[0] = __fadd_rn(z.p[0], __fmul_rn(z2.p[0], 1e-30f)); sink(b + i, z); }
#else
#if defined(CTB_LOAD_FIRST)
      {   // refill the slot BEFORE computing on its old contents: D blocks stay in flight during the compute
        const Raw cur = ring[i];
        CTB_PIN();
        load_raw(ring[i], w, rb + min(b + i + D, last), t);
        CTB_PIN();
        if (b + i < b1) sink(b + i, block_terms(cur, b + i, a, t));
      }
#else
      if (b + i < b1) sink(b + i, block_terms(ring[i], b + i, a, t));
#endif
#endif
#if !defined(CTB_LOAD_FIRST)
      CTB_PIN();
      load_raw(ring[i], w, rb + min(b + i + D, last), t);
#endif
    }
  }
}


// Blocks [b0, b1) of row `row`.  b0 == 0: folded as they are computed.  b0 > 0: the previous warp owns the row's fold
// state; the terms of up to MV_DEF_MAX blocks are parked in the warp's buffer (all the integer work is done before waiting),
// then the state is received, the parked terms are folded in order and any further blocks are folded directly.
// b1 == nb: the row is finished (returns true, value in `out`); else the state is posted to the next warp.
// The fold order is the reference's for every partition of the row.
template <typename Raw>
__device__ __forceinline__ bool run_segment_typed(const DevMat w, int row, int b0, int b1, const ActView& a, int lane, const Chain& ch, float& out) {
  const int nb = w.nb;
  const size_t rb = (size_t)row * nb;
  Fold f;
  fold_zero(f);
  const int u = lane % KQ_LPR, lead = lane - u;
  int bd = b0;
  if (b0 > 0) {
    bd = min(b1, b0 + ch.def_max);
    // parked per block and lane: p[], pm[] and (u == 0 ? dd : ddm) — dd and ddm are the same for all lanes of a row
    stream_blocks<Raw>(w, rb, b0, bd, a, lane, [&](int b, const BlockTerms& x) {
      float v[KQ_PARK_F4 * 4];
#pragma unroll
      for (int e = 0; e < KQ_NA; e++) v[e] = x.p[e];
#pragma unroll
      for (int i = 0; i < KQ_NM; i++) v[KQ_NA + i] = x.pm[i];
      v[KQ_NA + KQ_NM] = u == 0 ? x.dd : x.ddm;
#pragma unroll
      for (int q = 0; q < KQ_PARK_F4; q++)
        ch.buf[((size_t)(b - b0) * KQ_PARK_F4 + q) * 32 + lane] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    });
    while (*ch.flag_in == 0) { }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < KQ_NA; e++) f.a[e] = ch.mail_in[e * 32 + lane];
#pragma unroll
    for (int i = 0; i < KQ_NM; i++) f.am[i] = ch.mail_in[(KQ_NA + i) * 32 + lane];
    for (int b = b0; b < bd; b++) {
      float v[KQ_PARK_F4 * 4];
#pragma unroll
      for (int q = 0; q < KQ_PARK_F4; q++) {
        const float4 t4 = ch.buf[((size_t)(b - b0) * KQ_PARK_F4 + q) * 32 + lane];
        v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
      }
      BlockTerms x;
#pragma unroll
      for (int e = 0; e < KQ_NA; e++) x.p[e] = v[e];
#pragma unroll
      for (int i = 0; i < KQ_NM; i++) x.pm[i] = v[KQ_NA + i];
      x.dd = __shfl_sync(0xffffffffu, v[KQ_NA + KQ_NM], lead);
      x.ddm = __shfl_sync(0xffffffffu, v[KQ_NA + KQ_NM], lead + 1);
      fold_block(f, x);
    }
  }
  stream_blocks<Raw>(w, rb, bd, b1, a, lane, [&](int, const BlockTerms& x) { fold_block(f, x); });
  if (b1 < nb) {
#pragma unroll
    for (int e = 0; e < KQ_NA; e++) ch.mail_out[e * 32 + lane] = f.a[e];
#pragma unroll
    for (int i = 0; i < KQ_NM; i++) ch.mail_out[(KQ_NA + i) * 32 + lane] = f.am[i];
    __threadfence_block();
    __syncwarp();
    if (lane == 0) *ch.flag_out = 1;
    return false;
  }
  out = fold_finish(w.type, f);
  return true;
}

// KT = the one K-quant type of the launch (smaller kernel, no type switch in the loop), or 0 = decide per matrix
template <int KT>
__device__ __forceinline__ bool run_segment(const DevMat w, int row, int b0, int b1, const ActView& a, int lane, const Chain& ch, float& out) {
  if (KT == GT_Q4_K || (KT == 0 && w.type == GT_Q4_K)) return run_segment_typed<RawQ4K>(w, row, b0, b1, a, lane, ch, out);
  if (KT == GT_Q6_K || (KT == 0 && w.type == GT_Q6_K)) return run_segment_typed<RawQ6K>(w, row, b0, b1, a, lane, ch, out);
  return run_segment_typed<RawQ5K>(w, row, b0, b1, a, lane, ch, out);
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

__device__ __forceinline__ float dot_legacy(const DevMat& w, int row, const ActView& a, int l) {
  return w.type == GT_Q4_0 ? dot_q40(w, row, a, l) : dot_q80(w, row, a, l);
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
  else if (sg.epi == EPI_SILU) v = table_f16(p.silu_tab, v);
  sg.out[row] = v;
}

constexpr int MV_KQ_ROWS = 32 / KQ_LPR;   // K-quants: rows per tile (one warp)

// rows one work unit (one warp task) covers for a weight type
__host__ __device__ inline int rows_per_unit(int type) {
  if (type == GT_F16 || type == GT_F32) return 1;
  return type_is_kquant(type) ? MV_KQ_ROWS : MV_ROWS;
}
// relative cost of one row tile of a K-quant matrix (its bytes per block / 16)
__host__ __device__ inline int tile_cost(int type) { return type == GT_Q6_K ? 13 : (type == GT_Q5_K ? 11 : 9); }

// The K-quant tile space of a launch: the 8-row tiles of all its matrices, concatenated.
struct TileSpace {
  int tiles[MV_MAX_SEG], cost[MV_MAX_SEG], nseg, ntiles;
  long total;   // Σ tiles·cost
  __host__ __device__ __forceinline__ void init(const MVParams& p) {
    nseg = p.nseg; ntiles = 0; total = 0;
#pragma unroll
    for (int s = 0; s < MV_MAX_SEG; s++) {
      tiles[s] = s < p.nseg ? (p.seg[s].w.M + MV_KQ_ROWS - 1) / MV_KQ_ROWS : 0;
      cost[s] = s < p.nseg ? tile_cost(p.seg[s].w.type) : 1;
      ntiles += tiles[s]; total += (long)tiles[s] * cost[s];
    }
  }
  // matrix a tile of the concatenated space belongs to; `tile` becomes the tile index inside that matrix
  __host__ __device__ __forceinline__ int locate(int& tile) const {
    static_assert(MV_MAX_SEG == 3, "locate() is written out for three segments");
    if (tile < tiles[0]) return 0;
    tile -= tiles[0];
    if (tile < tiles[1]) return 1;
    tile -= tiles[1];
    return 2;
  }
  // first tile of CTA c of G: the tile at which the cumulative cost reaches c/G of the total
  __host__ __device__ __forceinline__ int boundary(int c, int G) const {
    if (c >= G) return ntiles;
    long target = total * c / G;
    int base = 0;
#pragma unroll
    for (int s = 0; s < MV_MAX_SEG; s++) {
      const long span = (long)tiles[s] * cost[s];
      if (target < span || s == MV_MAX_SEG - 1) return base + (int)min((long)tiles[s], (target + cost[s] / 2) / cost[s]);
      target -= span; base += tiles[s];
    }
    return ntiles;
  }
};

// ---------------------------------------------------------------------------------------------
// Persistent: one CTA per SM.  K-quant launches: CTA c owns a contiguous range of row tiles (cost-balanced); the blocks of
// those tiles, laid end to end, are cut into MV_WARPS equal contiguous pieces, one per warp, so every warp streams the same
// number of weight bytes whatever the shape.  A row tile cut between warps is folded in order by handing its fp32 state from
// warp to warp (run_segment_typed).  A warp first does the tiles it starts at block 0 (it can post their state early), then
// the tile it joined in the middle.  Other weight types: warp tasks strided over all warps of the grid.
template <int KT, bool ATTN>
static __global__ void __launch_bounds__(MV_THREADS, CTB_CTAS_PER_SM) k_matvec(const __grid_constant__ MVParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  __shared__ float mailbox[MV_WARPS + 1][KQ_FOLD_FLOATS * 32];
  __shared__ int flags[MV_WARPS + 1];
  __shared__ int cta_tiles;        // fused attention: row tiles this CTA finished
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool kq = KT != 0 || type_is_kquant(p.seg[0].w.type);
  if (threadIdx.x <= MV_WARPS) flags[threadIdx.x] = 0;
  if (threadIdx.x == 0) cta_tiles = 0;
  pdl_trigger();
  unsigned long long* const tr = p.trace ? p.trace + (size_t)blockIdx.x * (4 + MV_WARPS) : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = globaltimer_ns();

  // ---- this warp's share of the weights, known before any input is: ask L2 for it while the prologue runs
  TileSpace ts;
  int T0 = 0, nb = 1, s0 = 0, e0 = 0;
  if (kq) {
    ts.init(p);
    T0 = ts.boundary(blockIdx.x, gridDim.x);
    const int T1 = ts.boundary(blockIdx.x + 1, gridDim.x);
    nb = p.K >> 8;
    const int B = (T1 - T0) * nb, Lw = (B + MV_WARPS - 1) / MV_WARPS;
    s0 = min(B, warp * Lw); e0 = min(B, s0 + Lw);
    for (int pos = s0; pos < e0;) {
      int tile = T0 + pos / nb;
      const int b0 = pos % nb, len = min(nb - b0, e0 - pos);
      const int s = ts.locate(tile);
      const DevMat& w = p.seg[s].w;
      const int row = min(tile * MV_KQ_ROWS + lane / KQ_LPR, w.M - 1);
#if CTB_PF == 1
      prefetch_row(w, (size_t)row * nb, b0, b0 + len, lane % KQ_LPR);
#elif CTB_PF == 2
      prefetch_row(w, (size_t)row * nb, b0, min(b0 + len, b0 + CTB_PFD), lane % KQ_LPR);
#endif
      pos += len;
    }
  }

  NormPre np;
  preload_norm(np, p.norm_w, p.norm_b, p.norm_mode, p.K);
  pdl_wait();   // everything above touched only weights and shared memory; the input vector is the predecessor's output
  if (tr && threadIdx.x == 0) tr[1] = globaltimer_ns();
  stage_activation(p, np, p.norm_w, p.norm_b, p.norm_out, p.norm_mode, p.eps, p.K, p.act, smem, red, blockIdx.x == 0, tr ? tr + 3 : nullptr);
  const ActView a = act_view(p.act, p.K, smem);
  if (tr && threadIdx.x == 0) tr[2] = globaltimer_ns();

  if (kq) {
    int tiles_done = 0;
    if (s0 < e0) {
      Chain ch;
      uint8_t* const dyn = smem + ((act_smem_bytes(p.act, p.K) + 15) & ~(size_t)15);
      ch.def_max = p.def_max;
      ch.buf = (float4*)dyn + (size_t)warp * p.def_max * KQ_PARK_F4 * 32;
      ch.mail_in = mailbox[warp]; ch.flag_in = &flags[warp];
      ch.mail_out = mailbox[warp + 1]; ch.flag_out = &flags[warp + 1];
      const int a0 = s0 % nb;
      const int def_len = a0 ? min(nb - a0, e0 - s0) : 0;     // the piece of a tile another warp started
      const int sd = s0 + def_len;                              // tile-aligned from here on
      const int ndirect = (e0 - sd + nb - 1) / nb;
      const int nsegs = ndirect + (def_len ? 1 : 0);
      for (int i = 0; i < nsegs; i++) {
        int tile, b0, b1;
        if (i < ndirect) { const int pos = sd + i * nb; tile = T0 + pos / nb; b0 = 0; b1 = min(nb, e0 - pos); }
        else { tile = T0 + s0 / nb; b0 = a0; b1 = a0 + def_len; }
        const int s = ts.locate(tile);
        const MVSeg sg = s == 0 ? p.seg[0] : (s == 1 ? p.seg[1] : p.seg[2]);   // by value: static param-bank reads, pointers in registers
        const int row = tile * MV_KQ_ROWS + lane / KQ_LPR;
        float v = 0.f;
        const bool done = run_segment<KT>(sg.w, min(row, sg.w.M - 1), b0, b1, a, lane, ch, v);
        if (done && (lane % KQ_LPR) == 0 && row < sg.w.M) store_epilogue(sg, p, row, v);
        tiles_done += done ? 1 : 0;
      }
    }
    if (tr && lane == 0) tr[4 + warp] = globaltimer_ns();
    if (ATTN) {
      // ---- attention for this token, as soon as every row tile of q, k and v is in (they come from all CTAs of this launch)
      if (lane == 0 && tiles_done) atomicAdd(&cta_tiles, tiles_done);
      __threadfence();                 // this lane's output rows are visible device-wide before the tiles are counted
      __syncthreads();                 // ... and every warp is done with the activation / parking buffers
      if (threadIdx.x == 0 && cta_tiles) atomicAdd(p.attn_counter, cta_tiles);   // one device-wide update per CTA
      const int n_cg = p.attn.hd / ATTN_CH, n_tasks = p.attn.n_head * n_cg;
      for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        if (task != (int)blockIdx.x) __syncthreads();   // the previous task's shared-memory reads are over
        attn_body<1>(p.attn, smem, task / n_cg, 0, task % n_cg, p.attn_counter, ts.ntiles);
      }
    }
    return;
  }
  if (KT != 0) return;   // specialised instances carry no code for the other weight types

  const int gw = blockIdx.x * MV_WARPS + warp, nw = gridDim.x * MV_WARPS;
  int first = gw;   // global striding continues across segments so all warps stay busy
  for (int s = 0; s < p.nseg; s++) {
    const MVSeg& sg = p.seg[s];
    const int type = sg.w.type;
    const int rpu = rows_per_unit(type);
    const int units = (sg.w.M + rpu - 1) / rpu;
    int un = first;
    for (; un < units; un += nw) {
      if (rpu == 1) {
        const float v = type == GT_F16 ? dot_f16_row(sg.w, un, smem, lane) : dot_f32_row(sg.w, un, smem, lane);
        if (lane == 0) store_epilogue(sg, p, un, v);
      } else {
        const int row = un * MV_ROWS + (lane >> 3);
        const float v = dot_legacy(sg.w, min(row, sg.w.M - 1), a, lane & 7);
        if ((lane & 7) == 0 && row < sg.w.M) store_epilogue(sg, p, row, v);
      }
    }
    first = un - units;   // where this warp lands in the next segment
  }
}

// host-side launch geometry shared by the engine and the op-level entry points
struct MVLaunch { int grid; size_t smem; int kt; bool attn; };
inline MVLaunch matvec_launch_shape(MVParams& p, int n_sm) {
  MVLaunch L;
  const size_t act = (act_smem_bytes(p.act, p.K) + 15) & ~(size_t)15;
  const bool kq = type_is_kquant(p.seg[0].w.type);
  long units = 0;
  for (int s = 0; s < p.nseg; s++) { const int r = rows_per_unit(p.seg[s].w.type); units += (p.seg[s].w.M + r - 1) / r; }
  L.kt = 0;
  L.attn = false;
  p.def_max = 0;
  if (kq) {
    L.kt = p.seg[0].w.type;
    for (int s = 1; s < p.nseg; s++) if (p.seg[s].w.type != L.kt) L.kt = 0;
    L.attn = p.attn_on != 0;
    if (L.attn && L.kt != GT_Q4_K) L.kt = 0;   // the attention tail is instantiated for the Q4_K and the generic kernel only
    L.grid = (int)std::max<long>(1, std::min<long>(units, (long)n_sm * CTB_CTAS_PER_SM));
    const long room = (long)MV_SMEM_LIMIT - (long)act;
    // a parked (mid-row) segment is never longer than a warp's range nor than a row; shared memory not asked for stays L1.
    // The range comes from the largest CTA of the actual (cost-balanced) partition.
    TileSpace ts;
    ts.init(p);
    long tiles_per_cta = 1;
    for (int c = 0; c < L.grid; c++) tiles_per_cta = std::max<long>(tiles_per_cta, ts.boundary(c + 1, L.grid) - ts.boundary(c, L.grid));
    const long nb = p.K / 256;
    const long range = (tiles_per_cta * nb + MV_WARPS - 1) / MV_WARPS;
    const long need = std::min<long>(range, nb - 1);
    p.def_max = (int)std::max<long>(1, std::min<long>(std::min<long>(MV_DEF_MAX, need), room / (MV_WARPS * KQ_PARK_BYTES)));
    L.smem = act + (size_t)MV_WARPS * p.def_max * KQ_PARK_BYTES;
    if (p.attn_on) L.smem = std::max(L.smem, attn_smem_bytes(p.attn.n_ctx, p.attn.hd));
  } else {
    L.grid = (int)std::max<long>(1, std::min<long>((units + MV_WARPS - 1) / MV_WARPS, (long)n_sm));
    L.smem = act;
  }
  return L;
}


// static: each translation unit launches / configures ITS OWN instantiations of the (static) kernel template
static inline cudaError_t launch_matvec_kernel(const MVLaunch& L, cudaStream_t st, const MVParams& p, bool pdl = false) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(MV_THREADS); cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  if (L.attn) return L.kt == GT_Q4_K ? cudaLaunchKernelEx(&cfg, k_matvec<GT_Q4_K, true>, p) : cudaLaunchKernelEx(&cfg, k_matvec<0, true>, p);
  switch (L.kt) {
    case GT_Q4_K: return cudaLaunchKernelEx(&cfg, k_matvec<GT_Q4_K, false>, p);
    case GT_Q5_K: return cudaLaunchKernelEx(&cfg, k_matvec<GT_Q5_K, false>, p);
    case GT_Q6_K: return cudaLaunchKernelEx(&cfg, k_matvec<GT_Q6_K, false>, p);
    default: return cudaLaunchKernelEx(&cfg, k_matvec<0, false>, p);
  }
}
static inline cudaError_t matvec_set_smem_limit(int bytes) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_matvec<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_matvec<GT_Q4_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_matvec<GT_Q5_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_matvec<GT_Q6_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_matvec<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_matvec<GT_Q4_K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace ctb
