// Non-matmul stages of the eval graph, each reproducing the reference's rounding points:
//   k_embed     ggml_get_rows on a quantized token_embd            ggml.c:11615-11642 + dequantize_row_* (k_quants.c:784-821, 984-1026, 1123-1166; ggml.c:1483-1608)
//   k_rope_kv   RoPE (mode 0 / neox) in place on Q, on K → fp16 KV  ggml.c:12430-12566, llama.cpp:2303-2335
//   k_attn      K·q (fp16 operands, fp32 acc) → scale → causal mask → fp16-table softmax (fp64 sum) → P→fp16 → V·P
//               llama.cpp:2337-2400, ggml.c:11031 (F16 path), 11390, 11925-11973, 12009-12078
//   k_argmax    greedy pick on device (used by the fused decode loop; ties → lowest id)
#pragma once
#include "device_types.cuh"

namespace ctb {

// ------------------------------------------------------------------------------------------ embed
// token_embd keeps the GGUF array-of-blocks layout (one row is gathered per token; no streaming access).
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t* q, int& sc, int& m) {
  if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
  else { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

__device__ __forceinline__ float dequant_elem(int type, const uint8_t* row, int e) {
  switch (type) {
    case GT_F32: return ((const float*)row)[e];
    case GT_F16: return h2f(((const uint16_t*)row)[e]);
    case GT_Q4_0: {
      const uint8_t* blk = row + (size_t)(e >> 5) * 18;
      const int r = e & 31;
      const float d = h2f((uint16_t)(blk[0] | (blk[1] << 8)));
      const int byte = blk[2 + (r & 15)];
      const int nib = r < 16 ? (byte & 0xF) : (byte >> 4);
      return __fmul_rn((float)(nib - 8), d);
    }
    case GT_Q8_0: {
      const uint8_t* blk = row + (size_t)(e >> 5) * 34;
      const float d = h2f((uint16_t)(blk[0] | (blk[1] << 8)));
      return __fmul_rn((float)(int8_t)blk[2 + (e & 31)], d);
    }
    case GT_Q4_K: case GT_Q5_K: {
      const bool q5 = type == GT_Q5_K;
      const uint8_t* blk = row + (size_t)(e >> 8) * (q5 ? 176 : 144);
      const int r = e & 255, j = r >> 6, within = r & 63, sub = 2 * j + (within >> 5), l = within & 31;
      const float d = h2f((uint16_t)(blk[0] | (blk[1] << 8)));
      const float dmin = h2f((uint16_t)(blk[2] | (blk[3] << 8)));
      int sc, m;
      k4_scale_min(sub, blk + 4, sc, m);
      const uint8_t* qs = blk + (q5 ? 48 : 16);
      const int byte = qs[32 * j + l];
      int q = (sub & 1) ? (byte >> 4) : (byte & 0xF);
      if (q5 && (blk[16 + l] & (1 << sub))) q += 16;
      return __fsub_rn(__fmul_rn(__fmul_rn(d, (float)sc), (float)q), __fmul_rn(dmin, (float)m));
    }
    case GT_Q6_K: {
      const uint8_t* blk = row + (size_t)(e >> 8) * 210;
      const uint8_t* ql = blk; const uint8_t* qh = blk + 128; const int8_t* sc = (const int8_t*)(blk + 192);
      const float d = h2f((uint16_t)(blk[208] | (blk[209] << 8)));
      const int r = e & 255, n = r >> 7, rr = r & 127, k = rr >> 5, l = rr & 31, is = l >> 4;
      const int byte = ql[64 * n + ((k & 1) ? 32 : 0) + l];
      const int nib = (k >= 2) ? (byte >> 4) : (byte & 0xF);
      const int hb = (qh[32 * n + l] >> (2 * k)) & 3;
      const int q = (int)(int8_t)(nib | (hb << 4)) - 32;
      return __fmul_rn(__fmul_rn(d, (float)sc[8 * n + is + 2 * k]), (float)q);
    }
  }
  return 0.f;
}

// grid = N tokens; out[n][K]
static __global__ void k_embed(const uint8_t* table, int type, size_t row_bytes, int K, int n_vocab, const int* tokens, float* out) {
  const int tok = tokens[blockIdx.x];
  const uint8_t* row = table + (size_t)min(max(tok, 0), n_vocab - 1) * row_bytes;
  float* o = out + (size_t)blockIdx.x * K;
  for (int e = threadIdx.x; e < K; e += blockDim.x) o[e] = dequant_elem(type, row, e);
}

// ---------------------------------------------------------------------------------------- rope+kv
struct RopeKVParams {
  float* q;             // [N][n_head*hd]   rotated in place
  const float* k;       // [N][n_kv*hd]
  const float* v;       // [N][n_kv*hd]
  uint16_t* kc;         // this layer's K cache [n_ctx][n_kv*hd] fp16 (RoPE'd K, llama.cpp:2333)
  uint16_t* vc;         // this layer's V cache [n_kv][n_ctx][hd] fp16
  const float2* rope;   // [n_ctx][hd/2] (cos, sin), built on the host with libm exactly like the reference loop
  const int* n_past;    // device scalar
  int n_head, n_kv, hd, n_ctx, neox;
  int q_stride, kv_stride;   // row strides (floats) of q and k/v — falcon reads them out of one fused qkv row
};

// grid = (N, n_head + n_kv), block = hd/2
static __global__ void k_rope_kv(const RopeKVParams p) {
  const int n = blockIdx.x, hh = blockIdx.y, i = threadIdx.x;
  const int pos = *p.n_past + n;
  if (pos >= p.n_ctx) return;
  const float2 cs = p.rope[(size_t)pos * (p.hd / 2) + i];
  const int i0 = p.neox ? i : 2 * i, i1 = p.neox ? i + p.hd / 2 : 2 * i + 1;
  if (hh < p.n_head) {
    float* qh = p.q + (size_t)n * p.q_stride + (size_t)hh * p.hd;
    const float x0 = qh[i0], x1 = qh[i1];
    qh[i0] = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
    qh[i1] = __fadd_rn(__fmul_rn(x0, cs.y), __fmul_rn(x1, cs.x));
  } else {
    const int kh = hh - p.n_head;
    const float* ksrc = p.k + (size_t)n * p.kv_stride + (size_t)kh * p.hd;
    const float* vsrc = p.v + (size_t)n * p.kv_stride + (size_t)kh * p.hd;
    const float x0 = ksrc[i0], x1 = ksrc[i1];
    uint16_t* kd = p.kc + (size_t)pos * (p.n_kv * p.hd) + (size_t)kh * p.hd;
    kd[i0] = f2h(__fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y)));
    kd[i1] = f2h(__fadd_rn(__fmul_rn(x0, cs.y), __fmul_rn(x1, cs.x)));
    uint16_t* vd = p.vc + ((size_t)kh * p.n_ctx + pos) * p.hd;
    vd[2 * i] = f2h(vsrc[2 * i]);
    vd[2 * i + 1] = f2h(vsrc[2 * i + 1]);
  }
}

// ------------------------------------------------------------------------------------------- attn
struct AttnParams {
  const float* q;        // [N][q_stride] (already rotated)
  const uint16_t* kc;    // layer K cache
  const uint16_t* vc;    // layer V cache
  float* out;            // [N][n_head*hd]
  const uint16_t* exp_tab;
  const int* n_past;
  float kq_scale;
  int n_head, n_kv, hd, n_ctx, q_stride;
};

constexpr int ATTN_THREADS = 256;
constexpr int ATTN_WARPS = ATTN_THREADS / 32;

// grid = (n_head, N); dynamic smem = attn_smem_bytes(n_ctx, hd)
static __global__ void __launch_bounds__(ATTN_THREADS) k_attn(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float red_f[ATTN_WARPS];
  __shared__ double red_d[ATTN_WARPS];
  const int h = blockIdx.x, n = blockIdx.y;
  const int hd = p.hd;
  const int T = min(*p.n_past + n + 1, p.n_ctx);
  const int kvh = h / (p.n_head / p.n_kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const size_t ctx_pad = ((size_t)p.n_ctx + 7) & ~(size_t)7;   // keeps every sub-array 16-byte aligned for any n_ctx
  float* sc = (float*)smem;                                  // [ctx_pad]
  uint16_t* p16 = (uint16_t*)(smem + ctx_pad * 4);           // [ctx_pad]
  uint16_t* q16 = p16 + ctx_pad;                             // [hd]
  float* part = (float*)(smem + ctx_pad * 6 + (size_t)hd * 2);   // [ATTN_WARPS][hd]

  const float* qv = p.q + (size_t)n * p.q_stride + (size_t)h * hd;
  for (int i = threadIdx.x; i < hd; i += ATTN_THREADS) q16[i] = f2h(qv[i]);
  __syncthreads();

  // scores: one warp per cached position
  const int per = hd / 32;   // halves per lane (2 or 4; hd is 64 or 128)
  const size_t krow = (size_t)p.n_kv * hd;
  for (int t = warp; t < T; t += ATTN_WARPS) {
    const uint16_t* kr = p.kc + (size_t)t * krow + (size_t)kvh * hd + lane * per;
    float acc = 0.f;
    if (per == 4) {
      const uint2 kk = *(const uint2*)kr;
      const uint2 qq = *(const uint2*)(q16 + lane * 4);
      acc = fmaf(h2f((uint16_t)(kk.x & 0xffff)), h2f((uint16_t)(qq.x & 0xffff)), acc);
      acc = fmaf(h2f((uint16_t)(kk.x >> 16)), h2f((uint16_t)(qq.x >> 16)), acc);
      acc = fmaf(h2f((uint16_t)(kk.y & 0xffff)), h2f((uint16_t)(qq.y & 0xffff)), acc);
      acc = fmaf(h2f((uint16_t)(kk.y >> 16)), h2f((uint16_t)(qq.y >> 16)), acc);
    } else {
      for (int e = 0; e < per; e++) acc = fmaf(h2f(kr[e]), h2f(q16[lane * per + e]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) sc[t] = __fmul_rn(acc, p.kq_scale);
  }
  __syncthreads();

  // softmax (ggml.c:12047-12069)
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < T; t += ATTN_THREADS) mx = fmaxf(mx, sc[t]);
  mx = warp_max(mx);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  mx = red_f[0];
#pragma unroll
  for (int w = 1; w < ATTN_WARPS; w++) mx = fmaxf(mx, red_f[w]);
  double sum = 0.0;
  for (int t = threadIdx.x; t < T; t += ATTN_THREADS) {
    const float val = h2f(__ldg(p.exp_tab + f2h(__fsub_rn(sc[t], mx))));
    sc[t] = val;
    sum += (double)val;
  }
  sum = warp_sum(sum);
  if (lane == 0) red_d[warp] = sum;
  __syncthreads();
  sum = 0.0;
#pragma unroll
  for (int w = 0; w < ATTN_WARPS; w++) sum += red_d[w];
  const float inv = (float)(1.0 / sum);
  for (int t = threadIdx.x; t < T; t += ATTN_THREADS) p16[t] = f2h(__fmul_rn(sc[t], inv));
  __syncthreads();

  // V·P: warp w covers positions w, w+W, ...; lane owns `per` channels
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const uint16_t* vbase = p.vc + (size_t)kvh * p.n_ctx * hd + lane * per;
  for (int t = warp; t < T; t += ATTN_WARPS) {
    const float pt = h2f(p16[t]);
    const uint16_t* vr = vbase + (size_t)t * hd;
    if (per == 4) {
      const uint2 vv = *(const uint2*)vr;
      acc[0] = fmaf(pt, h2f((uint16_t)(vv.x & 0xffff)), acc[0]);
      acc[1] = fmaf(pt, h2f((uint16_t)(vv.x >> 16)), acc[1]);
      acc[2] = fmaf(pt, h2f((uint16_t)(vv.y & 0xffff)), acc[2]);
      acc[3] = fmaf(pt, h2f((uint16_t)(vv.y >> 16)), acc[3]);
    } else {
      for (int e = 0; e < per; e++) acc[e] = fmaf(pt, h2f(vr[e]), acc[e]);
    }
  }
  for (int e = 0; e < per; e++) part[warp * hd + lane * per + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < hd; c += ATTN_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < ATTN_WARPS; w++) s += part[w * hd + c];
    p.out[(size_t)n * p.n_head * hd + (size_t)h * hd + c] = s;
  }
}

__host__ inline size_t attn_smem_bytes(int n_ctx, int hd) {
  return (((size_t)n_ctx + 7) & ~(size_t)7) * 6 + (size_t)hd * 2 + (size_t)ATTN_WARPS * hd * 4;
}

// ----------------------------------------------------------------------------------------- argmax
// single block; writes the id of the largest logit (lowest id on ties) to *out
static __global__ void k_argmax(const float* logits, int n, int* out) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = logits[i];
    if (v > best) { best = v; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
    *out = idx;
  }
}

}  // namespace ctb
