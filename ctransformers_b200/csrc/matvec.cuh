// Decode-path mat-vec pieces shared by both kernels of the step (sm_100a), BIT-EXACT with the reference's AVX2 CPU build.
//
//   stage_activation<NT, BAR>  the prologue of every mat-vec: RMSNorm / LayerNorm with fp64 reductions + separate weight (+bias)
//                              multiply (ggml.c:10674-10720, 10605-10654) and the activation quantization to Q8_K (K-quants,
//                              k_quants.c:1191-1226) or Q8_0 (Q4_0/Q8_0, ggml.c:1232-1268) into shared memory, never HBM
//   k_matvec                   y[M] = W[M,K]·x[K] for the non-K-quant weight types (Q4_0 / Q8_0 / F16 / F32): one CTA per SM,
//                              warp tasks strided over the grid, the reference kernels' lane order restated
//                              (Q4_0 ggml.c:2500-2525 · Q8_0 3379-3402 · F16 2392-2426 · F32 2330-2365)
//   store_epilogue             store | + residual (ggml_add, llama.cpp:2415, 2453) | SiLU / GELU fp16 table (ggml.c:3568-3632)
//
// K-quant weights (Q4_K / Q5_K / Q6_K — everything a Q4_K_M / Q5_K_M file multiplies per token) go through the persistent
// step kernel in stream.cuh, which uses stage_activation and store_epilogue from here.
//
// Why exactness matters: the next mat-mul re-quantizes this output to int8; a 1-ulp difference can flip one rounding and the
// logits then differ by ~1e-3 (measured).  HBM traffic per launch = the weight planes once + O(K) activations from L2.
#pragma once
#include "device_types.cuh"
#include "attention.cuh"

namespace ctb {

#ifndef CTB_THREADS
#define CTB_THREADS 512
#endif
constexpr int MV_THREADS = CTB_THREADS;    // one persistent CTA per SM: the activation prologue is paid once per SM
constexpr int MV_WARPS = MV_THREADS / 32;
constexpr int MV_ROWS = 4;   // Q4_0 / Q8_0: rows per warp (one per 8-lane group, in-lane chain)
constexpr int MV_MAX_SEG = 3;

enum : int { NORM_NONE = 0, NORM_RMS = 1, NORM_LAYER = 2 };
enum : int { EPI_STORE = 0, EPI_ADD = 1, EPI_GELU = 2, EPI_ADD2 = 3, EPI_SILU = 4 };   // GELU / SILU: the reference's fp16-table activation of the row value

// Every spin in the persistent kernels is bounded: a wait that lasts longer than ST_WATCHDOG_NS writes {code, CTA, aux} into
// host-mapped memory and traps — a protocol bug then ends as a launch failure with a message, not as a hung GPU.
#ifndef ST_WATCHDOG_NS
#define ST_WATCHDOG_NS 4000000000ull
#endif
static __device__ int* g_st_dbg = nullptr;   // set by the host (st_set_debug_words): 4 ints of mapped pinned host memory, or null
static __device__ __noinline__ void st_fail(int code, int aux) {
  int* d = g_st_dbg;
  if (d) { d[0] = code; d[1] = (int)blockIdx.x; d[2] = aux; d[3] = (int)threadIdx.x; __threadfence_system(); }
  __trap();
}
#ifndef XC_WATCHDOG_NS
#define XC_WATCHDOG_NS 30000000000ull   // a peer rank may start its launch late (host jitter): 30 s
#endif

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
  int x_mode;            // 0: x;  1: x * x2 (ggml_mul of silu(gate) and up, llama.cpp:2438-2443; the SiLU table is applied by the gate rows' epilogue)
                         // 2: x is a uint2 array of {float bits, exchange number} elements (stream.cuh: XchgParams): the input is
                         //    part[0] + part[1] + ... + part[x_parts-1], parts x_stride elements apart, added in that order
                         //    (tensor-parallel partial sums of a row-parallel mat-vec, one per rank; rank 0's carries the
                         //    residual); an element is valid once its number equals the exchange number the caller passes
  int x_parts, x_stride;
  float* sum_out;        // x_mode 2, optional [K]: CTA 0 writes the summed vector (the residual stream of the next block)
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

// Whole prologue: normalise + quantize x[K] into shared memory.
// Each thread owns 16 consecutive elements per pass (one bsums group; 16 threads = one Q8_K block, 2 threads = one Q8_0
// block); all global loads of a pass are issued before anything depends on them.
// CG: the data was produced earlier in the SAME kernel by other SMs (persistent step kernel): read it from L2 (ld.global.cg)
template <bool CG = false>
__device__ __forceinline__ void load16(const float* p, int valid, float (&v)[16]) {
  if (valid >= 16) {
    float4 a, b, c, d;
    if (CG) { a = __ldcg((const float4*)p); b = __ldcg((const float4*)p + 1); c = __ldcg((const float4*)p + 2); d = __ldcg((const float4*)p + 3); }
    else { a = __ldg((const float4*)p); b = __ldg((const float4*)p + 1); c = __ldg((const float4*)p + 2); d = __ldg((const float4*)p + 3); }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
  } else {
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = e < valid ? (CG ? __ldcg(p + e) : __ldg(p + e)) : 0.f;
  }
}

// named barrier over the first NT threads of the CTA (BAR = 0 with NT = blockDim.x is __syncthreads)
template <int BAR, int NT>
__device__ __forceinline__ void bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory"); }

template <int NT, int BAR>
__device__ __forceinline__ double block_sum_f64(double s, double* red /* [NT / 32] smem */) {
  s = warp_sum(s);
  bar_sync<BAR, NT>();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  bar_sync<BAR, NT>();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; w++) t += red[w];
  return t;
}

__device__ __forceinline__ uint32_t pack4(const int* q) {
  return (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
}

// 16 consecutive input elements with the producer's activation applied: the reference's fp16-table SiLU (ggml.c:3625-3632)
// times the up projection, or the fp16-table GELU (ggml.c:3568-3575) — fused here instead of in the producing kernel so that
// gate and up can be two independent row sets there.
// 16 consecutive {value, number} elements of one rank's partial vector; spins until all carry `epoch`
__device__ __forceinline__ void load16_ll(const uint2* p, unsigned epoch, float (&u)[16]) {
  uint4 q[8];
  unsigned long long t0 = 0;
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[j].x), "=r"(q[j].y), "=r"(q[j].z), "=r"(q[j].w) : "l"(p + 2 * j) : "memory");
    }
#pragma unroll
    for (int j = 0; j < 8; j++) ok &= q[j].y == epoch && q[j].w == epoch;
    if (ok) break;
    if (!t0) t0 = globaltimer_ns();
    else if (globaltimer_ns() - t0 > XC_WATCHDOG_NS) st_fail(10, (int)epoch);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) { u[2 * j] = __uint_as_float(q[j].x); u[2 * j + 1] = __uint_as_float(q[j].z); }
}

template <bool XC = false>
__device__ __forceinline__ void load16x(const MVParams& xs, int base, int valid, float (&v)[16], unsigned epoch = 0) {
  if (XC && xs.x_mode == 2) {   // (K is a multiple of 256 and base of 16: a thread's 16 elements are all valid or all past the end)
    if (valid <= 0) {
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = 0.f;
      return;
    }
    const uint2* ll = (const uint2*)xs.x + (size_t)(epoch & 1u) * xs.x_parts * xs.x_stride;   // the parity half this exchange uses
    load16_ll(ll + base, epoch, v);
    for (int r = 1; r < xs.x_parts; r++) {
      float u[16];
      load16_ll(ll + (size_t)r * xs.x_stride + base, epoch, u);
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = __fadd_rn(v[e], u[e]);
    }
    return;
  }
  load16<true>(xs.x + base, valid, v);
  if (xs.x_mode == 1) {          // x = silu_table(gate) (applied once, where the gate row was produced); input = x * up (ggml_mul)
    float u[16];
    load16<true>(xs.x2 + base, valid, u);
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

// The first NT threads of the CTA must call (named barrier BAR); each owns 16 consecutive elements per pass.
template <int NT, int BAR, bool XC = false>   // XC: the input may be a tensor-parallel exchange (x_mode 2); compiled out otherwise
__device__ __forceinline__ void stage_activation(const MVParams& xs, const NormPre& np, const float* nw, const float* nb_, float* norm_out, int norm_mode, float eps,
                                                  int K, int act, uint8_t* smem, double* red, bool write_norm, unsigned epoch = 0) {
  const int t = threadIdx.x, lane = t & 31;
  const int passes = (K + NT * 16 - 1) / (NT * 16);
  // ---- statistics (fp64 sums like ggml.c:10700-10703 / 10630-10645; the order of a double sum does not reach the float result)
  float mean = 0.f, scale = 1.f;
  float v0[16];                          // pass 0's x stays in registers
  load16x<XC>(xs, t * 16, K - t * 16, v0, epoch);
  if (norm_mode == NORM_RMS) {
    double ss = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * NT + t) * 16;
      if ((base & ~511) >= K) continue;     // the whole warp lies past the end of x (warp-uniform): nothing to add
      float v[16];
      if (ps == 0) {
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = v0[e];
      } else load16x<XC>(xs, base, K - base, v, epoch);
#pragma unroll
      for (int e = 0; e < 16; e++) ss += (double)__fmul_rn(v[e], v[e]);
    }
    ss = block_sum_f64<NT, BAR>(ss, red);
    const float m = (float)(ss / (double)K);
    scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(m, eps)));
  } else if (norm_mode == NORM_LAYER) {
    double s1 = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * NT + t) * 16;
      if ((base & ~511) >= K) continue;
      float v[16];
      load16x<XC>(xs, base, K - base, v, epoch);
#pragma unroll
      for (int e = 0; e < 16; e++) s1 += (double)v[e];
    }
    s1 = block_sum_f64<NT, BAR>(s1, red);
    mean = (float)(s1 / (double)K);
    double s2 = 0.0;
    for (int ps = 0; ps < passes; ps++) {
      const int base = (ps * NT + t) * 16;
      if ((base & ~511) >= K) continue;
      float v[16];
      load16x<XC>(xs, base, K - base, v, epoch);
#pragma unroll
      for (int e = 0; e < 16; e++) { const float d = (base + e < K) ? __fsub_rn(v[e], mean) : 0.f; s2 += (double)__fmul_rn(d, d); }
    }
    s2 = block_sum_f64<NT, BAR>(s2, red);
    const float var = (float)(s2 / (double)K);
    scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
  }
  // ---- normalise + quantize
  int8_t* qs = (int8_t*)smem;
  const size_t off = ((size_t)K + 15) & ~(size_t)15;
  float* dd = (float*)(smem + off);
  int16_t* bs = (int16_t*)(smem + off + q8k_d_bytes(K));
  for (int ps = 0; ps < passes; ps++) {
    const int base = (ps * NT + t) * 16;
    const int valid = K - base;
    if ((base & ~511) >= K) continue;       // warp-uniform: this warp has no elements in this pass
    float v[16];
    if (ps == 0) {
#pragma unroll
      for (int e = 0; e < 16; e++) v[e] = v0[e];
    } else load16x<XC>(xs, base, valid, v, epoch);
    if (XC && write_norm && xs.x_mode == 2 && xs.sum_out && valid > 0) {
#pragma unroll
      for (int e = 0; e < 16; e++) if (e < valid) xs.sum_out[base + e] = v[e];
    }
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
  bar_sync<BAR, NT>();
}

// ---------------------------------------------------------------------------------------------
// hsum_float_8 (ggml.c:609-615) over the 8 lanes of a group: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)); every lane ends with the sum.
__device__ __forceinline__ float group_hsum8(float v) {
  v = v + __shfl_xor_sync(0xffffffffu, v, 4);
  v = v + __shfl_xor_sync(0xffffffffu, v, 2);
  v = v + __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
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

// Q5_0 (ggml.c:2983-3005): the nibble as for Q4_0 plus the fifth bit from qh (bit j = element j): the AVX2 kernel ORs 0xF0 into
// the bytes whose bit is CLEAR, i.e. the int8 value is q5 - 16; then the same s8·s8 dot, fmadd and hsum as Q4_0.
__device__ __forceinline__ float dot_q50(const DevMat& w, int row, const ActView& a, int l) {
  const int nb = w.nb;
  const uint8_t* qrow = w.qs + (size_t)row * nb * 16 + (l & 3) * 4;
  const uint32_t* hrow = (const uint32_t*)w.qh + (size_t)row * nb;
  const uint16_t* drow = w.d + (size_t)row * nb;
  const int shift = (l >> 2) * 4;
  float acc = 0.f;
#pragma unroll 8
  for (int b = 0; b < nb; b++) {
    const uint32_t q = (uint32_t)__ldg((const int*)(qrow + (size_t)b * 16));
    const uint32_t h4 = (__ldg(hrow + b) >> (4 * l)) & 0xfu;               // fifth bits of elements 4l..4l+3
    const float dw = h2f(__ldg(drow + b));
    const int aw = *(const int*)(a.qs + b * 32 + l * 4);
    const float yd = a.d[b];
    const uint32_t nib = (q >> shift) & 0x0f0f0f0fu;
    const uint32_t set = (h4 * 0x00204081u) & 0x01010101u;                   // bit k of h4 -> bit 0 of byte k
    const uint32_t bx = nib | ((set ^ 0x01010101u) * 0xf0u);                 // per byte: nibble - 16 when the bit is clear
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
  return w.type == GT_Q4_0 ? dot_q40(w, row, a, l) : (w.type == GT_Q5_0 ? dot_q50(w, row, a, l) : dot_q80(w, row, a, l));
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

// residuals may have been written earlier in the same (persistent) kernel: L2-coherent loads
__device__ __forceinline__ void store_epilogue(const MVSeg& sg, const MVParams& p, int row, float v) {
  if (sg.epi == EPI_ADD) v = __fadd_rn(v, __ldcg(sg.res + row));
  else if (sg.epi == EPI_ADD2) v = __fadd_rn(__fadd_rn(v, __ldcg(sg.res + row)), __ldcg(sg.res2 + row));
  else if (sg.epi == EPI_GELU) v = table_f16(p.gelu_tab, v);
  else if (sg.epi == EPI_SILU) v = table_f16(p.silu_tab, v);
  sg.out[row] = v;
}

// rows one warp task covers for a (non-K-quant) weight type
__host__ __device__ inline int rows_per_unit(int type) { return (type == GT_F16 || type == GT_F32) ? 1 : MV_ROWS; }

// ---------------------------------------------------------------------------------------------
// Q4_0 / Q8_0 / F16 / F32 weights.  Persistent: one CTA per SM, warp tasks strided over all warps of the grid.
static __global__ void __launch_bounds__(MV_THREADS, 1) k_matvec(const __grid_constant__ MVParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  NormPre np;
  preload_norm(np, p.norm_w, p.norm_b, p.norm_mode, p.K);
  pdl_wait();   // everything above touched only weights and shared memory; the input vector is the predecessor's output
  stage_activation<MV_THREADS, 0>(p, np, p.norm_w, p.norm_b, p.norm_out, p.norm_mode, p.eps, p.K, p.act, smem, red, blockIdx.x == 0);
  const ActView a = act_view(p.act, p.K, smem);
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
struct MVLaunch { int grid; size_t smem; };
inline MVLaunch matvec_launch_shape(const MVParams& p, int n_sm) {
  MVLaunch L;
  long units = 0;
  for (int s = 0; s < p.nseg; s++) { const int r = rows_per_unit(p.seg[s].w.type); units += (p.seg[s].w.M + r - 1) / r; }
  L.grid = (int)std::max<long>(1, std::min<long>((units + MV_WARPS - 1) / MV_WARPS, (long)n_sm));
  L.smem = (act_smem_bytes(p.act, p.K) + 15) & ~(size_t)15;
  return L;
}
constexpr int MV_SMEM_LIMIT = 200 * 1024;

// static: each translation unit launches / configures ITS OWN instantiation of the (static) kernel
static inline cudaError_t launch_matvec_kernel(const MVLaunch& L, cudaStream_t st, const MVParams& p, bool pdl = false) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.grid); cfg.blockDim = dim3(MV_THREADS); cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_matvec, p);
}
static inline cudaError_t matvec_set_smem_limit(int bytes) {
  return cudaFuncSetAttribute(k_matvec, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace ctb
