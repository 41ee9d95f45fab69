// Op-level C entry points (include/ctransformers_b200.h, part 2): each mirrors one ggml operator of the hot
// path with plain host pointers, runs the SAME device code the engine runs (stream.cuh / matvec.cuh / attention.cuh), and
// copies the result back.  Used by the parity tests and usable by a maintainer who wants to swap one op.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ctransformers_b200.h"
#include "attention.cuh"
#include "matvec.cuh"
#include "repack.cuh"
#include "stream.cuh"
#include "tables.hpp"

using namespace ctb;

namespace {

#define OPS_CUDA(expr)                                                                                     \
  do {                                                                                                     \
    cudaError_t e__ = (expr);                                                                              \
    if (e__ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " (" #expr ")"); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  explicit DevBuf(size_t n) { OPS_CUDA(cudaMalloc(&p, n ? n : 1)); }
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  template <typename T> T* as() const { return (T*)p; }
};

struct OpsTables {
  uint16_t *silu = nullptr, *gelu = nullptr, *ex = nullptr;
};

// built once per process, on the host with libm, like ggml_init (ggml.c:4319-4333)
OpsTables& tables() {
  static OpsTables t;
  if (!t.silu) {
    std::vector<uint16_t> s(65536), g(65536), e(65536);
    for (int i = 0; i < 65536; i++) {
      const float f = __half2float(__ushort_as_half((uint16_t)i));
      s[i] = __half_as_ushort(__float2half_rn(host_silu(f)));
      g[i] = __half_as_ushort(__float2half_rn(host_gelu(f)));
      e[i] = __half_as_ushort(__float2half_rn(expf(f)));
    }
    OPS_CUDA(cudaMalloc(&t.silu, 65536 * 2)); OPS_CUDA(cudaMalloc(&t.gelu, 65536 * 2)); OPS_CUDA(cudaMalloc(&t.ex, 65536 * 2));
    OPS_CUDA(cudaMemcpy(t.silu, s.data(), 65536 * 2, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(t.gelu, g.data(), 65536 * 2, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(t.ex, e.data(), 65536 * 2, cudaMemcpyHostToDevice));
  }
  return t;
}

size_t raw_row_bytes(int type, int K) {
  switch (type) {
    case GT_F32: return (size_t)K * 4; case GT_F16: return (size_t)K * 2;
    case GT_Q4_0: return (size_t)K / 32 * 18; case GT_Q5_0: return (size_t)K / 32 * 22; case GT_Q8_0: return (size_t)K / 32 * 34;
    case GT_Q4_K: return (size_t)K / 256 * 144; case GT_Q5_K: return (size_t)K / 256 * 176; case GT_Q6_K: return (size_t)K / 256 * 210;
  }
  throw std::runtime_error("unsupported ggml type " + std::to_string(type));
}
int block_elems(int type) { return type_is_kquant(type) ? 256 : (type == GT_Q4_0 || type == GT_Q5_0 || type == GT_Q8_0) ? 32 : 1; }

struct OwnedMat {
  DevMat m;
  std::vector<void*> bufs;
  ~OwnedMat() { for (void* b : bufs) cudaFree(b); }
};

void upload(OwnedMat& o, int type, const void* blocks, int K, int M) {
  if (K % block_elems(type)) throw std::runtime_error("K is not a multiple of the block size");
  const size_t bytes = raw_row_bytes(type, K) * M;
  DevBuf raw(bytes);
  OPS_CUDA(cudaMemcpy(raw.p, blocks, bytes, cudaMemcpyHostToDevice));
  o.m.type = type; o.m.K = K; o.m.M = M; o.m.nb = K / block_elems(type); o.m.bytes = bytes;
  if (type_is_kquant(type)) {   // the stream layout of the step kernel
    const size_t sb = st_matrix_bytes(type, M, o.m.nb);
    uint16_t* st = nullptr;
    OPS_CUDA(cudaMalloc((void**)&st, sb));
    o.bufs.push_back(st);
    k_repack_stream<<<(int)std::min<size_t>((sb / 2 + 255) / 256, 4096), 256>>>(type, raw.as<uint8_t>(), M, o.m.nb, st);
    OPS_CUDA(cudaDeviceSynchronize());
    o.m.st = (const uint8_t*)st;
    return;
  }
  const PlaneSizes ps = plane_sizes(type, M, o.m.nb, bytes);
  uint16_t* pl[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t sz[4] = {ps.qs, ps.qh, ps.sc, ps.d};
  for (int i = 0; i < 4; i++)
    if (sz[i]) { OPS_CUDA(cudaMalloc((void**)&pl[i], sz[i])); o.bufs.push_back(pl[i]); }
  k_repack<<<(int)std::min<size_t>((bytes / 2 + 255) / 256, 4096), 256>>>(type, raw.as<uint16_t>(), bytes / 2, pl[0], pl[1], pl[2], pl[3]);
  OPS_CUDA(cudaDeviceSynchronize());
  o.m.qs = (const uint8_t*)pl[0]; o.m.qh = (const uint8_t*)pl[1]; o.m.sc = (const uint8_t*)pl[2]; o.m.d = pl[3];
}

int sm_count() {
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  return n_sm;
}

// a program of phases through the persistent step kernel, exactly as the engine launches it
void run_phases(std::vector<Phase> phs) {
  static unsigned* d_sync = nullptr;
  if (!d_sync) { OPS_CUDA(cudaMalloc((void**)&d_sync, 64)); OPS_CUDA(cudaMemset(d_sync, 0, 64)); }
  const StepLaunch L = step_launch_shape(phs.data(), (int)phs.size(), sm_count(), step_max_dyn_smem());
  const std::vector<int> hb = step_bounds(phs.data(), (int)phs.size(), L.grid);
  DevBuf dbounds(hb.size() * 4);
  OPS_CUDA(cudaMemcpy(dbounds.p, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  if (L.n_slots < ST_W) throw std::runtime_error("rows too long for the step kernel's shared memory");
  OPS_CUDA(step_set_smem_limit(L.smem));
  DevBuf dprog((phs.size() + 1) * sizeof(Phase));
  OPS_CUDA(cudaMemcpy(dprog.p, phs.data(), phs.size() * sizeof(Phase), cudaMemcpyHostToDevice));
  OPS_CUDA(launch_step(L, 0, dprog.as<Phase>(), dbounds.as<int>(), (int)phs.size(), d_sync));
  OPS_CUDA(cudaGetLastError());
  OPS_CUDA(cudaDeviceSynchronize());
}

void run_matvec(MVParams& p) {
  p.silu_tab = tables().silu;
  p.gelu_tab = tables().gelu;
  if (step_supports(p)) {
    run_phases({matvec_phase(p)});
    return;
  }
  static bool attr = false;
  if (!attr) {
    OPS_CUDA(matvec_set_smem_limit(MV_SMEM_LIMIT));
    attr = true;
  }
  const MVLaunch L = matvec_launch_shape(p, sm_count());
  OPS_CUDA(launch_matvec_kernel(L, 0, p));
  OPS_CUDA(cudaGetLastError());
}

// standalone wrappers around the prologue pieces, so the activation quantizers can be checked bit-for-bit
__global__ void __launch_bounds__(MV_THREADS) k_stage_dump(const float* x, const float* nw, const float* nb, float* norm_out, int mode, float eps, int K, int act,
                                                            uint8_t* dump) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  MVParams q{};
  q.x = x;
  NormPre np;
  preload_norm(np, nw, nb, mode, K);
  stage_activation<MV_THREADS, 0>(q, np, nw, nb, norm_out, mode, eps, K, act, smem, red, true);
  const size_t n = act_smem_bytes(act, K);
  for (size_t i = threadIdx.x; i < n; i += MV_THREADS) dump[i] = smem[i];
}

// the x_mode = 1 input path of the prologue on its own: out[i] = gate_act[i] * up[i] (gate_act = silu_table(gate), from the epilogue)
__global__ void __launch_bounds__(MV_THREADS) k_gate_dump(const float* gate, const float* up, const uint16_t* silu_tab, int M, float* out) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double red[MV_WARPS];
  MVParams q{};
  q.x = gate; q.x2 = up; q.x_mode = 1; q.silu_tab = silu_tab;
  NormPre np;
  preload_norm(np, nullptr, nullptr, NORM_NONE, M);
  stage_activation<MV_THREADS, 0>(q, np, nullptr, nullptr, nullptr, NORM_NONE, 0.f, M, ACT_F32, smem, red, false);
  const float* f = (const float*)smem;
  for (int i = threadIdx.x; i < M; i += MV_THREADS) out[i] = f[i];
}

int guarded(const char* what, const std::function<void()>& fn) {
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device available (no CPU fallback)");
    fn();
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: %s failed: %s\n", what, e.what());
    return -1;
  }
}

void stage_to_host(const float* x, const float* w, const float* b, float* y_norm, int mode, float eps, int K, int act, std::vector<uint8_t>& dump) {
  DevBuf dx((size_t)K * 4), dw((size_t)K * 4), db((size_t)K * 4), dy((size_t)K * 4);
  OPS_CUDA(cudaMemcpy(dx.p, x, (size_t)K * 4, cudaMemcpyHostToDevice));
  if (w) OPS_CUDA(cudaMemcpy(dw.p, w, (size_t)K * 4, cudaMemcpyHostToDevice));
  if (b) OPS_CUDA(cudaMemcpy(db.p, b, (size_t)K * 4, cudaMemcpyHostToDevice));
  const size_t n = act_smem_bytes(act, K);
  DevBuf dd(n);
  if (n > 48 * 1024) OPS_CUDA(cudaFuncSetAttribute(k_stage_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)n));
  k_stage_dump<<<1, MV_THREADS, n>>>(dx.as<float>(), w ? dw.as<float>() : nullptr, b ? db.as<float>() : nullptr, dy.as<float>(), mode, eps, K, act, dd.as<uint8_t>());
  OPS_CUDA(cudaGetLastError());
  dump.resize(n);
  OPS_CUDA(cudaMemcpy(dump.data(), dd.p, n, cudaMemcpyDeviceToHost));
  if (y_norm) OPS_CUDA(cudaMemcpy(y_norm, dy.p, (size_t)K * 4, cudaMemcpyDeviceToHost));
}

}  // namespace

extern "C" {

int ctb_mul_mat(int type, const void* w_blocks, const float* x, float* dst, int K, int M, int N) {
  return guarded("ctb_mul_mat", [&] {
    OwnedMat w;
    upload(w, type, w_blocks, K, M);
    DevBuf dx((size_t)K * N * 4), dy((size_t)M * N * 4);
    OPS_CUDA(cudaMemcpy(dx.p, x, (size_t)K * N * 4, cudaMemcpyHostToDevice));
    for (int n = 0; n < N; n++) {
      MVParams p{};
      p.x = dx.as<float>() + (size_t)n * K; p.norm_mode = NORM_NONE; p.K = K; p.act = act_format_for(type); p.nseg = 1;
      p.seg[0].w = w.m; p.seg[0].out = dy.as<float>() + (size_t)n * M; p.seg[0].epi = EPI_STORE;
      run_matvec(p);
    }
    OPS_CUDA(cudaMemcpy(dst, dy.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  });
}

int ctb_quantize_row_q8_K(const float* x, void* y, int k) {
  return guarded("ctb_quantize_row_q8_K", [&] {
    if (k % 256) throw std::runtime_error("k must be a multiple of 256");
    std::vector<uint8_t> dump;
    stage_to_host(x, nullptr, nullptr, nullptr, NORM_NONE, 0.f, k, ACT_Q8_K, dump);
    const int nb = k / 256;
    const size_t off = ((size_t)k + 15) & ~(size_t)15;
    const float* d = (const float*)(dump.data() + off);
    const int16_t* bs = (const int16_t*)(dump.data() + off + q8k_d_bytes(k));
    uint8_t* out = (uint8_t*)y;   // block_q8_K: float d; int8 qs[256]; int16 bsums[16]  (k_quants.h:121-125)
    for (int b = 0; b < nb; b++) {
      memcpy(out + (size_t)b * 292, d + b, 4);
      for (int e = 0; e < 256; e++) {   // undo the lane-major shared-memory order (matvec.cuh q8k_word_offset)
        const int wd = e >> 2, sb = wd >> 3, ln = wd & 7;
        out[(size_t)b * 292 + 4 + e] = dump[(size_t)((b * 2 + (sb >> 2)) * 8 + ln) * 16 + (sb & 3) * 4 + (e & 3)];
      }
      memcpy(out + (size_t)b * 292 + 260, bs + (size_t)b * 16, 32);
    }
  });
}

int ctb_quantize_row_q8_0(const float* x, void* y, int k) {
  return guarded("ctb_quantize_row_q8_0", [&] {
    if (k % 32) throw std::runtime_error("k must be a multiple of 32");
    std::vector<uint8_t> dump;
    stage_to_host(x, nullptr, nullptr, nullptr, NORM_NONE, 0.f, k, ACT_Q8_0, dump);
    const int nb = k / 32;
    const size_t off = ((size_t)k + 15) & ~(size_t)15;
    const float* d = (const float*)(dump.data() + off);
    uint8_t* out = (uint8_t*)y;   // block_q8_0: fp16 d; int8 qs[32]  (ggml.c:920-925)
    for (int b = 0; b < nb; b++) {
      const uint16_t h = __half_as_ushort(__float2half_rn(d[b]));   // d[b] is already fp16-representable
      memcpy(out + (size_t)b * 34, &h, 2);
      memcpy(out + (size_t)b * 34 + 2, dump.data() + (size_t)b * 32, 32);
    }
  });
}

int ctb_norm(int mode, const float* x, const float* w, const float* b, float* y, int n, float eps) {
  return guarded("ctb_norm", [&] {
    std::vector<uint8_t> dump;
    stage_to_host(x, w, b, y, mode, eps, n, ACT_F32, dump);
  });
}

int ctb_rope(float* x, int n_heads, int head_dim, int pos, int mode, float freq_base, float freq_scale) {
  return guarded("ctb_rope", [&] {
    const int half = head_dim / 2;
    std::vector<float2> tab((size_t)(pos + 1) * half);
    const float theta_scale = powf(freq_base, -2.0f / head_dim);
    for (int p = 0; p <= pos; p++) {
      float theta = freq_scale * (float)p;
      for (int i = 0; i < half; i++) { tab[(size_t)p * half + i] = make_float2(cosf(theta), sinf(theta)); theta *= theta_scale; }
    }
    const size_t nq = (size_t)n_heads * head_dim;
    DevBuf dq(nq * 4), dtab(tab.size() * 8), dst(16);
    OPS_CUDA(cudaMemcpy(dq.p, x, nq * 4, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dtab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    const int st[4] = {0, pos, 0, pos + 1};
    OPS_CUDA(cudaMemcpy(dst.p, st, 16, cudaMemcpyHostToDevice));
    RopeKVParams rp{};
    rp.q = dq.as<float>(); rp.rope = dtab.as<float2>(); rp.state = dst.as<int>(); rp.n_head = n_heads; rp.n_kv = 1; rp.hd = head_dim;
    rp.n_ctx = pos + 1; rp.neox = (mode & 2) ? 1 : 0; rp.q_stride = (int)nq; rp.kv_stride = (int)nq;
    k_rope_kv<<<dim3(1, n_heads), half>>>(rp);   // grid.y == n_head: only the Q-head blocks exist, no KV store happens
    OPS_CUDA(cudaGetLastError());
    OPS_CUDA(cudaMemcpy(x, dq.p, nq * 4, cudaMemcpyDeviceToHost));
  });
}

int ctb_attention(const float* q, const uint16_t* kcache, const uint16_t* vcache, float* out, int n_head, int n_kv, int head_dim, int T,
                  int n_total, float kq_scale) {
  return guarded("ctb_attention", [&] {
    if (head_dim != 64 && head_dim != 128) throw std::runtime_error("head_dim must be 64 or 128");
    if (n_total < T) n_total = T;
    const int cp = kv_ctx_pad(n_total);
    const size_t nq = (size_t)n_head * head_dim;
    // reference layouts in (K [T][n_kv*hd], V transposed [n_kv*hd][T]) -> our permuted device layouts
    std::vector<uint16_t> kp((size_t)n_total * n_kv * head_dim, 0), vp((size_t)n_kv * head_dim * cp, 0);
    for (int t = 0; t < T; t++)
      for (int kh = 0; kh < n_kv; kh++)
        for (int e = 0; e < head_dim; e++)
          kp[k_row(kh, t, n_total, head_dim) + k_perm(e, head_dim)] = kcache[((size_t)t * n_kv + kh) * head_dim + e];
    for (int ch = 0; ch < n_kv * head_dim; ch++)
      for (int t = 0; t < T; t++) vp[(size_t)ch * cp + v_perm(t)] = vcache[(size_t)ch * T + t];
    // the kernel fuses RoPE + KV store for the current position: feed it an identity rotation (cos 1, sin 0 is exact) and the
    // current position's k/v taken back out of the caller's caches (f16 -> f32 -> f16 round-trips exactly)
    const int pos = T - 1;
    std::vector<float> kcur((size_t)n_kv * head_dim), vcur((size_t)n_kv * head_dim);
    for (int kh = 0; kh < n_kv; kh++)
      for (int e = 0; e < head_dim; e++) {
        kcur[(size_t)kh * head_dim + e] = __half2float(__ushort_as_half(kcache[((size_t)pos * n_kv + kh) * head_dim + e]));
        vcur[(size_t)kh * head_dim + e] = __half2float(__ushort_as_half(vcache[((size_t)kh * head_dim + e) * T + pos]));
      }
    std::vector<float2> ident((size_t)n_total * (head_dim / 2), make_float2(1.f, 0.f));
    DevBuf dq(nq * 4), dk(kp.size() * 2), dv(vp.size() * 2), dout(nq * 4), dst(16), dkc(kcur.size() * 4), dvc(vcur.size() * 4), dtab(ident.size() * 8);
    OPS_CUDA(cudaMemcpy(dq.p, q, nq * 4, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dk.p, kp.data(), kp.size() * 2, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dv.p, vp.data(), vp.size() * 2, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dkc.p, kcur.data(), kcur.size() * 4, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dvc.p, vcur.data(), vcur.size() * 4, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dtab.p, ident.data(), ident.size() * 8, cudaMemcpyHostToDevice));
    const int st[4] = {0, pos, 0, n_total};
    OPS_CUDA(cudaMemcpy(dst.p, st, 16, cudaMemcpyHostToDevice));
    AttnParams ap{};
    ap.q = dq.as<float>(); ap.k = dkc.as<float>(); ap.v = dvc.as<float>(); ap.kc = dk.as<uint16_t>(); ap.vc = dv.as<uint16_t>();
    ap.out = dout.as<float>(); ap.exp_tab = tables().ex; ap.rope = dtab.as<float2>(); ap.state = dst.as<int>(); ap.kq_scale = kq_scale;
    ap.n_head = n_head; ap.n_kv = n_kv; ap.hd = head_dim; ap.n_ctx = n_total; ap.q_stride = (int)nq; ap.kv_stride = n_kv * head_dim; ap.neox = 0;
    const size_t smem = attn_smem_bytes(n_total, head_dim);
    OPS_CUDA(cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    k_attn<<<dim3(n_head, 1, head_dim / ATTN_CH), ATTN_THREADS, smem>>>(ap);
    OPS_CUDA(cudaGetLastError());
    OPS_CUDA(cudaMemcpy(out, dout.p, nq * 4, cudaMemcpyDeviceToHost));
  });
}

int ctb_ffn_gate(int type, const void* w1_blocks, const void* w3_blocks, const float* x, float* out, int K, int M) {
  return guarded("ctb_ffn_gate", [&] {
    OwnedMat w1, w3;
    upload(w1, type, w1_blocks, K, M);
    upload(w3, type, w3_blocks, K, M);
    DevBuf dx((size_t)K * 4), dg((size_t)M * 4), du((size_t)M * 4), dy((size_t)M * 4);
    OPS_CUDA(cudaMemcpy(dx.p, x, (size_t)K * 4, cudaMemcpyHostToDevice));
    // as the engine does it: gate and up rows in one launch, SiLU(gate)*up where the down projection stages its input
    MVParams p{};
    p.x = dx.as<float>(); p.norm_mode = NORM_NONE; p.K = K; p.act = act_format_for(type); p.nseg = 2;
    p.seg[0].w = w1.m; p.seg[0].out = dg.as<float>(); p.seg[0].epi = EPI_SILU; p.seg[1].w = w3.m; p.seg[1].out = du.as<float>();
    run_matvec(p);
    const size_t smem = (size_t)M * 4 + 64;
    OPS_CUDA(cudaFuncSetAttribute(k_gate_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    k_gate_dump<<<1, MV_THREADS, smem>>>(dg.as<float>(), du.as<float>(), tables().silu, M, dy.as<float>());
    OPS_CUDA(cudaGetLastError());
    OPS_CUDA(cudaMemcpy(out, dy.p, (size_t)M * 4, cudaMemcpyDeviceToHost));
  });
}

// Host-side view of how a K-quant mat-vec phase is cut up (no GPU needed): the 16-row tile range of every CTA.
int ctb_matvec_partition(const int* types, const int* rows, int nseg, int K, int n_sm, int* first_tile, int* meta) {
  if (nseg < 1 || nseg > MV_MAX_SEG || K <= 0 || K % 256 != 0 || n_sm < 1) return -1;
  MVParams p{};
  p.K = K; p.nseg = nseg; p.act = ACT_Q8_K;
  for (int s = 0; s < nseg; s++) {
    if (!type_is_kquant(types[s]) || rows[s] < 1) return -1;
    p.seg[s].w.type = types[s]; p.seg[s].w.K = K; p.seg[s].w.M = rows[s]; p.seg[s].w.nb = K / 256;
  }
  TileSpace ts;
  ts.init(p);
  const int nb = K / 256;
  long max_items = 0;
  for (int c = 0; c <= n_sm; c++) first_tile[c] = ts.boundary(c, n_sm);
  for (int c = 0; c < n_sm; c++) {
    long items = 0;
    for (int tile = first_tile[c]; tile < first_tile[c + 1]; tile++) {
      int tl = tile;
      const int kb = st_chunk_blocks(types[ts.locate(tl)]);
      items += (nb + kb - 1) / kb;
    }
    max_items = std::max(max_items, items);
  }
  meta[0] = n_sm; meta[1] = ST_SLOT; meta[2] = ST_MAXT; meta[3] = ts.ntiles; meta[4] = ST_W; meta[5] = ST_ROWS; meta[6] = (int)max_items;
  meta[7] = st_chunk_blocks(GT_Q4_K) | (st_chunk_blocks(GT_Q5_K) << 8) | (st_chunk_blocks(GT_Q6_K) << 16);
  return 0;
}

int ctb_get_row(int type, const void* table_blocks, int K, int n_rows, int row, float* out) {
  return guarded("ctb_get_row", [&] {
    const size_t rb = raw_row_bytes(type, K);
    DevBuf dt(rb * n_rows), dtok(4), dout((size_t)K * 4);
    OPS_CUDA(cudaMemcpy(dt.p, table_blocks, rb * n_rows, cudaMemcpyHostToDevice));
    OPS_CUDA(cudaMemcpy(dtok.p, &row, 4, cudaMemcpyHostToDevice));
    k_embed<<<1, 256>>>(dt.as<uint8_t>(), type, rb, K, n_rows, dtok.as<int>(), dout.as<float>());
    OPS_CUDA(cudaGetLastError());
    OPS_CUDA(cudaMemcpy(out, dout.p, (size_t)K * 4, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
