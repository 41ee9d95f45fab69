// See engine.cuh.  Model upload (GGUF blocks → device planes), tables, KV cache, op schedule, CUDA graphs.
#include "engine.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "attention.cuh"
#include "matvec.cuh"
#include "prefill.cuh"
#include "repack.cuh"
#include "sample_gpu.cuh"
#include "stream.cuh"
#include "tables.hpp"
#include "tp_nccl.hpp"

namespace ctb {

#define CTB_CUDA(expr)                                                                                         \
  do {                                                                                                         \
    cudaError_t e__ = (expr);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__) + " (" #expr ")" + watchdog_note());                   \
  } while (0)

// what a trapped persistent kernel left in the host-mapped watchdog words (stream.cuh st_fail)
static int* g_watchdog_words = nullptr;
static std::string watchdog_note() {
  const int* d = g_watchdog_words;
  if (!d) return " [no watchdog words]";
  if (d[0] == 0) return " [watchdog words clear]";
  static const char* const what[] = {"?", "mbarrier", "grid barrier (aux = phase)", "fold hand-off flag (aux = chunk)", "producer: free weight slot (aux = item)",
                                     "consumer: weight item (aux = item)", "producer: free K slot (attention)", "producer: free V slot (attention)",
                                     "consumer: K item (attention)", "consumer: V item (attention)", "tensor-parallel exchange: a peer's element (aux = exchange number)", "?",
                                     "prefill grid barrier (aux = phase)", "?", "prefill producer: free slot", "prefill consumer: weight item"};
  const char* w = (d[0] > 0 && d[0] < (int)(sizeof(what) / sizeof(what[0]))) ? what[d[0]] : "?";
  return " [step-kernel watchdog: wait " + std::to_string(d[0]) + " (" + w + ") timed out in CTA " + std::to_string(d[1]) + ", aux " + std::to_string(d[2]) + ", thread " +
         std::to_string(d[3]) + "]";
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// the engine works on its own device but leaves the caller's current device as it found it
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) {
      cudaError_t e = cudaSetDevice(dev);
      if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (cudaSetDevice)");
    }
  }
  ~DeviceGuard() { int cur = -1; cudaGetDevice(&cur); if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};

// one op of the per-token schedule: a phase of the step kernel, or (mat-vecs over non-K-quant weights) a kernel of its own
struct StepOp {
  Phase ph;
  int mvk = 0;          // PH_MATVEC: which projection (MVK_*)
  bool stream = false;  // PH_MATVEC: runs inside the step kernel
};

// advance the on-device decode state after a greedy pick: state = {token, n_past, step}
// state = {token, position, step, n_total}; pick lives in state[4]
__global__ void k_advance(int* state, int* out_tokens) {
  const int t = state[4];
  out_tokens[state[2]] = t;
  state[0] = t;
  state[1] += 1;
  state[2] += 1;
  state[3] = state[1] + 1;   // a single-token eval: the attention rows have length position + 1
}

// ---- batched prefill (prefill.cuh): the per-token schedule rewritten over PB_T-row buffers
struct PrefillState {
  std::vector<void*> bufs;          // cudaMalloc'ed
  PPhase* d_prog = nullptr;
  int n_phases = 0;
  int* d_state = nullptr;           // [PB_T][4] + n_tok
  int* h_state = nullptr;           // pinned, PF_RING launches deep
  int h_next = 0;
  float* x_final = nullptr;         // batched buffer that holds the last layer's output rows
  int n_slots = 0;
  size_t smem = 0;
  bool ok = false, tried = false;
  ~PrefillState() {
    for (void* b : bufs) cudaFree(b);
    if (h_state) cudaFreeHost(h_state);
  }
};
constexpr int PF_RING = 64;

constexpr size_t UP_CHUNK = (size_t)32 << 20;   // upload pipeline: chunk bytes and buffers in flight
constexpr int UP_BUFS = 3;

static bool supported_matrix_type(uint32_t t) {
  return t == T_F32 || t == T_F16 || t == T_Q4_0 || t == T_Q5_0 || t == T_Q8_0 || t == T_Q4_K || t == T_Q5_K || t == T_Q6_K;
}

size_t engine_arena_bytes(const GGUFFile& g, const HParams& hp) {
  size_t total = 0;
  for (const auto& t : g.tensors) {
    total += align_up(t.nbytes, 256) + 4 * 256;   // up to 4 planes, each 256-aligned
    if (t.ne[1] > 0) total += align_up(t.nbytes / (size_t)t.ne[1] * ST_ROWS, 256);   // K-quants: rows padded to whole 16-row tiles
  }
  total += UP_CHUNK * UP_BUFS + 256;                                            // device staging of the upload pipeline
  const size_t kv = (size_t)hp.n_layer * (hp.n_ctx + 256) * hp.n_embd_gqa() * 2;
  total += 2 * align_up(kv, 256);
  total += 3 * align_up(65536 * 2, 256);
  total += align_up((size_t)hp.n_ctx * (hp.head_dim() / 2) * 8, 256);
  const size_t qkv = (size_t)hp.n_embd + 2 * (size_t)hp.n_embd_gqa();
  total += 4 * (2 * (size_t)hp.n_embd + qkv + 4 * (size_t)hp.n_embd + 2 * (size_t)hp.n_ff + 2 * (size_t)hp.n_vocab) + 64 * 256 + 8192;
  total += ((size_t)hp.n_layer * 10 + 8) * (sizeof(Phase) * 2 + 2 * 1024) + 8192;   // the step programs and their per-CTA tile ranges
  total += 1 << 20;
  return total;
}

void* Engine::alloc(size_t bytes, size_t align) {
  const size_t off = align_up(arena_used_, align);
  if (off + bytes > arena_size_) throw std::runtime_error("device arena exhausted");
  arena_used_ = off + bytes;
  return arena_ + off;
}

// ---- load pipeline (reference: llama_model_loader::load_all_data, llama.cpp:1417-1487 + ggml_cuda_transform_tensor,
// ggml-cuda.cu:6359-6432 — a blocking cudaMemcpy per tensor).  Here a tensor travels in row chunks of <= UP_CHUNK bytes through
// UP_BUFS pinned host buffers and as many device staging buffers: the host thread copies chunk i+1 out of the mmap'ed file
// into pinned memory while chunk i is on the wire (cudaMemcpyAsync from pinned memory is truly asynchronous) and chunk i-1 is
// being repacked on the GPU; buffers are recycled behind events, nothing synchronises per tensor.
struct Uploader {
  uint8_t* host[UP_BUFS] = {nullptr, nullptr, nullptr};
  uint8_t* dev[UP_BUFS] = {nullptr, nullptr, nullptr};
  cudaEvent_t done[UP_BUFS] = {nullptr, nullptr, nullptr};
  cudaStream_t st[UP_BUFS] = {nullptr, nullptr, nullptr};
  int next = 0;
  size_t bytes = 0;
  void init(uint8_t* dev_base) {
    for (int i = 0; i < UP_BUFS; i++) {
      CTB_CUDA(cudaMallocHost(&host[i], UP_CHUNK));
      dev[i] = dev_base + (size_t)i * UP_CHUNK;
      CTB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
      CTB_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    }
  }
  // copies [src, src+n) to the device staging buffer of the next slot; returns the slot (its stream carries the copy)
  int push(const uint8_t* src, size_t n) {
    const int i = next;
    next = (next + 1) % UP_BUFS;
    CTB_CUDA(cudaEventSynchronize(done[i]));        // the slot's previous chunk has been repacked
    memcpy(host[i], src, n);
    CTB_CUDA(cudaMemcpyAsync(dev[i], host[i], n, cudaMemcpyHostToDevice, st[i]));
    bytes += n;
    return i;
  }
  void finish(int i) { CTB_CUDA(cudaEventRecord(done[i], st[i])); }
  void drain() { for (int i = 0; i < UP_BUFS; i++) if (st[i]) CTB_CUDA(cudaStreamSynchronize(st[i])); }
  void release() {
    for (int i = 0; i < UP_BUFS; i++) {
      if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
      if (done[i]) cudaEventDestroy(done[i]);
      if (host[i]) cudaFreeHost(host[i]);
      st[i] = nullptr; done[i] = nullptr; host[i] = nullptr;
    }
  }
};

// rows [row0, row1) and elements [k0, k1) of every row (whole quantization blocks) are kept: a tensor-parallel shard
// (column-parallel = a row range, row-parallel = a K range); the defaults keep the whole tensor.
DevMat Engine::upload_matrix(const GGUFTensor& t, Uploader& up, int want_K, int want_M, int row0, int row1, int k0, int k1) {
  if (!supported_matrix_type(t.type)) throw std::runtime_error("tensor '" + t.name + "': quantization type " + std::to_string(t.type) + " is not supported by the B200 path");
  DevMat m;
  m.type = (int)t.type;
  const int full_K = (int)t.ne[0], full_M = (int)(t.ne[1] * t.ne[2] * t.ne[3]);
  // the reference rejects a tensor whose shape does not follow from the hyper-parameters (llama.cpp:1345-1361 "wrong shape")
  if (full_K != want_K || full_M != want_M)
    throw std::runtime_error("tensor '" + t.name + "' has wrong shape; expected " + std::to_string(want_K) + " x " + std::to_string(want_M) + ", got " +
                             std::to_string(full_K) + " x " + std::to_string(full_M));
  if (row1 < 0) row1 = full_M;
  if (k1 < 0) k1 = full_K;
  const int be = type_block_elems(t.type);
  if (row0 < 0 || row1 > full_M || row0 >= row1 || k0 < 0 || k1 > full_K || k0 >= k1 || k0 % be || k1 % be)
    throw std::runtime_error("tensor '" + t.name + "': shard does not fall on quantization block boundaries");
  m.K = k1 - k0;
  m.M = row1 - row0;
  m.nb = m.K / be;
  const size_t full_row_bytes = t.nbytes / (size_t)full_M;
  const size_t row_bytes = (size_t)m.nb * type_block_bytes(t.type);
  m.bytes = row_bytes * (size_t)m.M;
  const uint8_t* src = t.data + (size_t)row0 * full_row_bytes;
  std::vector<uint8_t> gathered;
  if (m.K != full_K) {   // a K range: the kept blocks of every row, packed
    gathered.resize(m.bytes);
    const size_t off = (size_t)(k0 / be) * type_block_bytes(t.type);
    for (int r = 0; r < m.M; r++) memcpy(gathered.data() + (size_t)r * row_bytes, src + (size_t)r * full_row_bytes + off, row_bytes);
    src = gathered.data();
  }
  int rows_per_chunk = (int)std::max<size_t>(ST_ROWS, UP_CHUNK / row_bytes / ST_ROWS * ST_ROWS);   // whole 16-row tiles
  if (row_bytes * ST_ROWS > UP_CHUNK) throw std::runtime_error("tensor '" + t.name + "': rows too long for the upload staging buffers");
  uint16_t *st = nullptr, *qs = nullptr, *qh = nullptr, *d = nullptr;
  const bool kq = type_is_kquant(m.type);
  if (kq) {
    st = (uint16_t*)alloc(st_matrix_bytes(m.type, m.M, m.nb));
    m.st = (const uint8_t*)st;
  } else {
    const PlaneSizes ps = plane_sizes(m.type, m.M, m.nb, m.bytes);
    qs = (uint16_t*)alloc(ps.qs);
    if (ps.qh) qh = (uint16_t*)alloc(ps.qh);
    if (ps.d) d = (uint16_t*)alloc(ps.d);
    m.qs = (const uint8_t*)qs; m.qh = (const uint8_t*)qh; m.d = d;
  }
  for (int r0 = 0; r0 < m.M; r0 += rows_per_chunk) {
    const int rows = std::min(rows_per_chunk, m.M - r0);
    const size_t n = (size_t)rows * row_bytes;
    const int slot = up.push(src + (size_t)r0 * row_bytes, n);
    if (kq) {
      const size_t sb = st_matrix_bytes(m.type, rows, m.nb);
      const int grid = (int)std::min<size_t>((sb / 2 + 255) / 256, (size_t)sm_count_ * 32);
      k_repack_stream<<<grid, 256, 0, up.st[slot]>>>(m.type, up.dev[slot], rows, m.nb, st + (size_t)(r0 / ST_ROWS) * m.nb * st_block_bytes(m.type) / 2);
    } else {
      const size_t n_u16 = n / 2;
      const size_t blk0 = (size_t)r0 * m.nb;
      const int grid = (int)std::min<size_t>((n_u16 + 255) / 256, (size_t)sm_count_ * 32);
      uint16_t* qdst = qs + ((m.type == GT_Q4_0 || m.type == GT_Q5_0) ? blk0 * 8 : (m.type == GT_Q8_0 ? blk0 * 16 : (size_t)r0 * row_bytes / 2));
      k_repack<<<grid, 256, 0, up.st[slot]>>>(m.type, (const uint16_t*)up.dev[slot], n_u16, qdst, qh ? qh + blk0 * 2 : nullptr, nullptr, d ? d + blk0 : nullptr);
    }
    CTB_CUDA(cudaGetLastError());
    up.finish(slot);
  }
  return m;
}

const float* Engine::upload_vector(const GGUFFile& g, const std::string& name, bool required, int want_n) {
  const GGUFTensor* t = g.tensor(name);
  if (!t) {
    if (required) throw std::runtime_error("tensor '" + name + "' not found");
    return nullptr;
  }
  if (t->type != T_F32) throw std::runtime_error("tensor '" + name + "' must be f32");
  if ((long)t->ne[0] * (long)t->ne[1] != (long)want_n) throw std::runtime_error("tensor '" + name + "' has wrong shape");
  float* d = (float*)alloc(t->nbytes);
  CTB_CUDA(cudaMemcpy(d, t->data, t->nbytes, cudaMemcpyHostToDevice));
  return d;
}

// fp16 <-> fp32 on the host through the F16C-equivalent software path (bit-identical to the device's and the reference's)
static inline float host_h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (!man) bits = sign;
    else { int e = -1; do { man <<= 1; e++; } while (!(man & 0x400u)); man &= 0x3ffu; bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13); }
  } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
  else bits = sign | ((exp + 112) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}
static inline uint16_t host_f2h(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0));
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
  if (ax < 0x33000001u) return (uint16_t)sign;
  const int32_t e = (int32_t)(ax >> 23) - 127;
  const uint32_t m = (ax & 0x7fffffu) | 0x800000u;
  if (e < -14) {
    const uint32_t shift = (uint32_t)(13 + (-14 - e));
    uint32_t r = m >> shift; const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (r & 1))) r++;
    return (uint16_t)(sign | r);
  }
  uint32_t hb = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu);
  const uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (hb & 1))) hb++;
  return (uint16_t)(sign | hb);
}

TPShard tp_shard(int n_embd, int n_head, int n_head_kv, int n_ff, int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world || n_head <= 0 || n_head_kv <= 0 || n_embd % n_head || n_head % n_head_kv) throw std::runtime_error("tensor parallel: bad shape or rank");
  const int hd = n_embd / n_head;
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  const int group = 256 / gcd(256, hd);   // query heads per 256-element block of the attention output (2 for head_dim 128)
  if ((hd * group) % 256 || n_head % group || n_ff % 256) throw std::runtime_error("tensor parallel: heads / n_ff do not tile into 256-element blocks");
  auto split = [&](int units, int r0) {   // first element of part r0 when `units` are dealt as evenly as possible (the first units % world parts get one more)
    const int base = units / world, extra = units % world;
    return r0 * base + std::min(r0, extra);
  };
  TPShard s;
  s.rank = rank; s.world = world;
  s.head0 = split(n_head / group, rank) * group; s.head1 = split(n_head / group, rank + 1) * group;
  s.ff0 = split(n_ff / 256, rank) * 256; s.ff1 = split(n_ff / 256, rank + 1) * 256;
  const int per_kv = n_head / n_head_kv;
  s.kv0 = s.head1 > s.head0 ? s.head0 / per_kv : 0;
  s.kv1 = s.head1 > s.head0 ? (s.head1 - 1) / per_kv + 1 : 0;
  return s;
}

Engine::Engine(const GGUFFile& g, const HParams& hp, int device, const TPShard& tp) : hp_(hp), tp_(tp), device_(device) {
  // everything acquired below is released by release() if the constructor throws (the destructor does not run then)
  try {
    init(g);
  } catch (...) {
    release();
    throw;
  }
}

void Engine::init(const GGUFFile& g) {
  CTB_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  CTB_CUDA(cudaGetDeviceProperties(&prop, device_));
  sm_count_ = prop.multiProcessorCount;
  // refuse what the kernels cannot run BEFORE the model-sized allocations are made
  for (const auto& t : g.tensors)
    if (t.n_dims >= 2 && !supported_matrix_type(t.type))
      throw std::runtime_error("tensor '" + t.name + "': quantization type " + std::to_string(t.type) + " is not supported by the B200 path");
  CTB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  CTB_CUDA(cudaEventCreate(&ev0_));
  CTB_CUDA(cudaEventCreate(&ev1_));

  arena_size_ = engine_arena_bytes(g, hp_);
  CTB_CUDA(cudaMalloc(&arena_, arena_size_));
  const auto t_load0 = std::chrono::steady_clock::now();
  Uploader up;
  struct UpGuard { Uploader& u; ~UpGuard() { u.release(); } } up_guard{up};
  up.init((uint8_t*)alloc(UP_CHUNK * UP_BUFS));

  // ---- weights (shapes follow from the hyper-parameters: llama.cpp:1878-1934 llama, 1948-2012 falcon)
  const std::string pfx = "blk.";
  const int n_embd = hp_.n_embd, gqa = hp_.n_embd_gqa(), n_ff = hp_.n_ff;
  nh_ = hp_.n_head; nkv_ = hp_.n_head_kv; nff_ = hp_.n_ff;
  const int hd0 = hp_.head_dim();
  if (tp_.world > 1) {
    const int per_kv = hp_.n_head / hp_.n_head_kv;
    if (hp_.falcon) throw std::runtime_error("tensor parallel mode covers the llama graph only");
    if (!tp_.comm) throw std::runtime_error("tensor parallel mode needs a communicator");
    if (tp_.head1 <= tp_.head0 || tp_.ff1 <= tp_.ff0) throw std::runtime_error("tensor parallel: more ranks than 256-element blocks to share out");
    if (tp_.head0 % per_kv || tp_.head1 % per_kv) throw std::runtime_error("tensor parallel: a rank's query heads must cover whole KV groups");
    nh_ = tp_.head1 - tp_.head0; nkv_ = tp_.kv1 - tp_.kv0; nff_ = tp_.ff1 - tp_.ff0;
    prefill_on_ = false;   // prompts go through the single-token path (the batched kernel has no exchange step)
  }
  const int q0 = tp_.world > 1 ? tp_.head0 * hd0 : 0, q1 = tp_.world > 1 ? tp_.head1 * hd0 : n_embd;        // rows of wq = K range of wo
  const int g0 = tp_.world > 1 ? tp_.kv0 * hd0 : 0, g1 = tp_.world > 1 ? tp_.kv1 * hd0 : gqa;               // rows of wk / wv
  const int f0 = tp_.world > 1 ? tp_.ff0 : 0, f1 = tp_.world > 1 ? tp_.ff1 : n_ff;                          // rows of w1 / w3 = K range of w2
  {
    const GGUFTensor& te = g.need_tensor("token_embd.weight");
    if (!supported_matrix_type(te.type)) throw std::runtime_error("token_embd.weight: unsupported type");
    if ((int)te.ne[0] != n_embd || (long)(te.ne[1] * te.ne[2] * te.ne[3]) != (long)hp_.n_vocab) throw std::runtime_error("token_embd.weight has wrong shape");
    uint8_t* d = (uint8_t*)alloc(te.nbytes);
    CTB_CUDA(cudaMemcpy(d, te.data, te.nbytes, cudaMemcpyHostToDevice));
    tok_embd_ = d; tok_type_ = (int)te.type;
    tok_row_bytes_ = te.ne[0] / type_block_elems(te.type) * type_block_bytes(te.type);
  }
  out_norm_ = upload_vector(g, "output_norm.weight", true, n_embd);
  out_norm_b_ = upload_vector(g, "output_norm.bias", hp_.falcon, n_embd);
  output_ = upload_matrix(g.need_tensor("output.weight"), up, n_embd, hp_.n_vocab);
  size_t wbytes = output_.bytes;
  layers_.resize(hp_.n_layer);
  for (int il = 0; il < hp_.n_layer; il++) {
    LayerW& L = layers_[il];
    const std::string b = pfx + std::to_string(il) + ".";
    L.attn_norm = upload_vector(g, b + "attn_norm.weight", true, n_embd);
    if (hp_.falcon) {
      L.attn_norm_b = upload_vector(g, b + "attn_norm.bias", true, n_embd);
      L.attn_norm2 = upload_vector(g, b + "attn_norm_2.weight", false, n_embd);
      if (L.attn_norm2) L.attn_norm2_b = upload_vector(g, b + "attn_norm_2.bias", true, n_embd);
      L.wqkv = upload_matrix(g.need_tensor(b + "attn_qkv.weight"), up, n_embd, n_embd + 2 * gqa);
      L.wo = upload_matrix(g.need_tensor(b + "attn_output.weight"), up, n_embd, n_embd);
      L.w3 = upload_matrix(g.need_tensor(b + "ffn_up.weight"), up, n_embd, n_ff);
      L.w2 = upload_matrix(g.need_tensor(b + "ffn_down.weight"), up, n_ff, n_embd);
      wbytes += L.wqkv.bytes + L.wo.bytes + L.w3.bytes + L.w2.bytes;
    } else {
      L.ffn_norm = upload_vector(g, b + "ffn_norm.weight", true, n_embd);
      L.wq = upload_matrix(g.need_tensor(b + "attn_q.weight"), up, n_embd, n_embd, q0, q1);
      L.wk = upload_matrix(g.need_tensor(b + "attn_k.weight"), up, n_embd, gqa, g0, g1);
      L.wv = upload_matrix(g.need_tensor(b + "attn_v.weight"), up, n_embd, gqa, g0, g1);
      L.wo = upload_matrix(g.need_tensor(b + "attn_output.weight"), up, n_embd, n_embd, 0, -1, q0, q1);
      L.w1 = upload_matrix(g.need_tensor(b + "ffn_gate.weight"), up, n_embd, n_ff, f0, f1);
      L.w2 = upload_matrix(g.need_tensor(b + "ffn_down.weight"), up, n_ff, n_embd, 0, -1, f0, f1);
      L.w3 = upload_matrix(g.need_tensor(b + "ffn_up.weight"), up, n_embd, n_ff, f0, f1);
      wbytes += L.wq.bytes + L.wk.bytes + L.wv.bytes + L.wo.bytes + L.w1.bytes + L.w2.bytes + L.w3.bytes;
      if (act_format_for(L.w1.type) != act_format_for(L.w3.type)) throw std::runtime_error("ffn_gate / ffn_up use incompatible quantization families");
    }
  }
  stats.weight_bytes_per_token = wbytes;
  up.drain();
  stats.load_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_load0).count();
  stats.load_bytes = up.bytes;

  // ---- lookup tables, built with the host libm exactly like ggml_init does (ggml.c:4319-4333)
  {
    std::vector<uint16_t> silu(65536), gelu(65536), ex(65536);
    for (int i = 0; i < 65536; i++) {
      const float f = host_h2f((uint16_t)i);
      silu[i] = host_f2h(host_silu(f));
      gelu[i] = host_f2h(host_gelu(f));
      ex[i] = host_f2h(expf(f));
    }
    silu_tab_ = (uint16_t*)alloc(65536 * 2); gelu_tab_ = (uint16_t*)alloc(65536 * 2); exp_tab_ = (uint16_t*)alloc(65536 * 2);
    CTB_CUDA(cudaMemcpy(silu_tab_, silu.data(), 65536 * 2, cudaMemcpyHostToDevice));
    CTB_CUDA(cudaMemcpy(gelu_tab_, gelu.data(), 65536 * 2, cudaMemcpyHostToDevice));
    CTB_CUDA(cudaMemcpy(exp_tab_, ex.data(), 65536 * 2, cudaMemcpyHostToDevice));
  }
  // ---- RoPE table: same recurrence, same libm calls as ggml.c:12482-12529
  {
    const int half = hp_.head_dim() / 2;
    std::vector<float2> tab((size_t)hp_.n_ctx * half);
    const float theta_scale = powf(hp_.rope_base, -2.0f / hp_.n_rot);
    for (int p = 0; p < hp_.n_ctx; p++) {
      float theta = hp_.rope_scale * (float)p;
      for (int i = 0; i < half; i++) {
        tab[(size_t)p * half + i] = make_float2(cosf(theta), sinf(theta));
        theta *= theta_scale;
      }
    }
    rope_ = (float2*)alloc(tab.size() * sizeof(float2));
    CTB_CUDA(cudaMemcpy(rope_, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice));
  }
  // ---- KV cache + workspace
  const size_t gqa_l = (size_t)nkv_ * hp_.head_dim(), qw_l = (size_t)nh_ * hp_.head_dim();   // this rank's K/V and Q widths
  const size_t kv = (size_t)hp_.n_layer * hp_.n_ctx * gqa_l;
  const size_t vv = (size_t)hp_.n_layer * kv_ctx_pad(hp_.n_ctx) * gqa_l;
  kc_ = (uint16_t*)alloc(kv * 2);
  vc_ = (uint16_t*)alloc(vv * 2);
  CTB_CUDA(cudaMemset(kc_, 0, kv * 2));
  CTB_CUDA(cudaMemset(vc_, 0, vv * 2));
  const size_t qkv = qw_l + 2 * gqa_l;
  d_state_ = (int*)alloc(64);
  xa_ = (float*)alloc(hp_.n_embd * 4); xb_ = (float*)alloc(hp_.n_embd * 4);
  qkv_ = (float*)alloc(qkv * 4);
  attn_ = (float*)alloc(hp_.n_embd * 4); attn_o_ = (float*)alloc(hp_.n_embd * 4);
  ffn_ = (float*)alloc((size_t)nff_ * 4);
  ffn2_ = (float*)alloc((size_t)nff_ * 4);
  d_logits_ = (float*)alloc((size_t)hp_.n_vocab * 4);
  d_embd_ = (float*)alloc(hp_.n_embd * 4);
  d_logits_keep_ = (float*)alloc((size_t)hp_.n_vocab * 4);
  d_embd_keep_ = (float*)alloc(hp_.n_embd * 4);
  d_sync_ = (unsigned*)alloc(64);
  CTB_CUDA(cudaMemset(d_state_, 0, 64));
  CTB_CUDA(cudaMemset(d_sync_, 0, 64));
  CTB_CUDA(cudaMallocHost(&h_logits_, (size_t)hp_.n_vocab * 4));
  CTB_CUDA(cudaMallocHost(&h_embd_, (size_t)hp_.n_embd * 4));
  memset(h_logits_, 0, (size_t)hp_.n_vocab * 4);
  memset(h_embd_, 0, (size_t)hp_.n_embd * 4);

  if (const char* e = getenv("CTB_NO_PDL")) pdl_ = !(e[0] == '1');
  if (const char* e = getenv("CTB_NO_SPEC")) spec_on_ = !(e[0] == '1');
  if (const char* e = getenv("CTB_STEP_FUSE")) fused_ = !(e[0] == '0');
  if (const char* e = getenv("CTB_NO_PREFILL")) prefill_on_ = !(e[0] == '1');
  if (const char* e = getenv("CTB_PREFILL_MIN")) prefill_min_ = std::max(1, atoi(e));
  CTB_CUDA(cudaMallocHost(&h_spec_tok_, 16));
  {   // the kernels' watchdog words: host memory the device can write and the host can read after a trapped launch
    CTB_CUDA(cudaHostAlloc(&h_dbg_, 64, cudaHostAllocMapped));
    memset(h_dbg_, 0, 64);
    int* d = nullptr;
    CTB_CUDA(cudaHostGetDevicePointer(&d, h_dbg_, 0));
    CTB_CUDA(st_set_debug_words(d));
    g_watchdog_words = h_dbg_;
  }
  CTB_CUDA(cudaEventCreateWithFlags(&ev_pick_, cudaEventDisableTiming));
  CTB_CUDA(matvec_set_smem_limit(MV_SMEM_LIMIT));
  CTB_CUDA(cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem_bytes(hp_.n_ctx, hp_.head_dim())));
  if (tp_.world > 1) tp_setup_peer();
  build_ops();
  if (tp_peer_)
    for (const StepOp& op : ops_)
      if (op.ph.kind == PH_MATVEC && !op.stream) throw std::runtime_error("tensor parallel (fused exchange): every mat-vec must run in the step kernel (K-quant weights)");
  CTB_CUDA(cudaDeviceSynchronize());
  build_graphs();
}

void Engine::release() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  destroy_graphs();
  delete pf_;
  pf_ = nullptr;
  if (h_logits_) cudaFreeHost(h_logits_);
  if (h_embd_) cudaFreeHost(h_embd_);
  if (h_state_) cudaFreeHost(h_state_);
  if (h_tokens_out_) cudaFreeHost(h_tokens_out_);
  if (d_tokens_out_) cudaFree(d_tokens_out_);
  for (int r = 0; r < 8; r++)
    if (xc_ll_[r] && r != tp_.rank) cudaIpcCloseMemHandle(xc_ll_[r]);
  if (xc_region_) cudaFree(xc_region_);
  xc_region_ = nullptr;
  for (int r = 0; r < 8; r++) xc_ll_[r] = nullptr;
  if (arena_) cudaFree(arena_);
  if (h_spec_tok_) cudaFreeHost(h_spec_tok_);
  if (h_dbg_) cudaFreeHost(h_dbg_);
  h_dbg_ = nullptr;
  if (h_sample_) cudaFreeHost(h_sample_);
  h_sample_ = nullptr;
  if (ev_pick_) cudaEventDestroy(ev_pick_);
  if (ev_sample_) cudaEventDestroy(ev_sample_);
  ev_sample_ = nullptr;
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  if (stream_ && own_stream_) cudaStreamDestroy(stream_);
  h_logits_ = h_embd_ = nullptr; h_state_ = nullptr; h_tokens_out_ = nullptr; d_tokens_out_ = nullptr; arena_ = nullptr; h_spec_tok_ = nullptr;
  ev_pick_ = ev0_ = ev1_ = nullptr; stream_ = nullptr;
}

Engine::~Engine() { release(); }

void Engine::set_stream(cudaStream_t s) {
  if (own_stream_ && stream_) { cudaStreamSynchronize(stream_); cudaStreamDestroy(stream_); }
  stream_ = s;
  own_stream_ = false;
}

static MVSeg seg(const DevMat& w, float* out, int epi = EPI_STORE, const float* res = nullptr, const float* res2 = nullptr) {
  MVSeg s;
  s.w = w; s.out = out; s.res = res; s.res2 = res2; s.epi = epi;
  return s;
}

// The op list of one token through the whole model (the reference rebuilds this graph on every eval, llama.cpp:2872-2876;
// here it is a static schedule whose pointers never change): EMBED, per layer {QKV mat-vec(s), ATTN, WO, UP, DOWN}, then the
// HEAD mat-vec and the greedy PICK.  {token, n_past} are read from d_state_ on the device.
void Engine::push_matvec(MVParams& p, int kind) {
  p.silu_tab = silu_tab_;
  p.gelu_tab = gelu_tab_;
  StepOp op{};
  op.ph = step_supports(p) ? matvec_phase(p) : Phase{};
  op.ph.kind = PH_MATVEC;
  op.ph.mv = p;
  op.mvk = kind;
  op.stream = step_supports(p);
  if (op.stream) op.ph.mv.act = ACT_Q8_K;
  ops_.push_back(op);
}

void Engine::tp_all_reduce(float* buf, int n) {
  const NcclApi& nccl = NcclApi::get();
  nccl.check(nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)tp_.comm, stream_), "all-reduce");
}

// Fused exchange set-up: allocate this rank's region, hand its CUDA IPC handle to the peers (the NCCL communicator carries the 64
// bytes), map theirs.  All-or-nothing across ranks: if any rank cannot map a peer, every rank stays on the NCCL all-reduce path.
void Engine::tp_setup_peer() {
  const NcclApi& nccl = NcclApi::get();
  const int W = tp_.world;
  const size_t bytes = align_up((size_t)2 * W * hp_.n_embd * sizeof(uint2), 256);
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::vector<cudaIpcMemHandle_t> all(W);
  uint8_t* d_h = (uint8_t*)alloc(sizeof(cudaIpcMemHandle_t) * W + 16, 256);
  int* d_ok = (int*)(d_h + sizeof(cudaIpcMemHandle_t) * W);
  if (W > XC_MAX_WORLD || getenv("CTB_TP_NCCL")) ok = 0;
  if (ok && cudaMalloc(&xc_region_, bytes) != cudaSuccess) { xc_region_ = nullptr; ok = 0; }
  if (ok) {
    CTB_CUDA(cudaMemset(xc_region_, 0, bytes));
    if (cudaIpcGetMemHandle(&mine, xc_region_) != cudaSuccess) ok = 0;
  }
  if (!ok) memset(&mine, 0, sizeof(mine));
  cudaGetLastError();
  CTB_CUDA(cudaMemcpy(d_h + sizeof(mine) * tp_.rank, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  nccl.check(nccl.AllGather(d_h + sizeof(mine) * tp_.rank, d_h, sizeof(mine), ncclChar, (ncclComm_t)tp_.comm, stream_), "all-gather of the IPC handles");
  CTB_CUDA(cudaStreamSynchronize(stream_));
  CTB_CUDA(cudaMemcpy(all.data(), d_h, sizeof(mine) * W, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W && ok; r++) {
    void* base = xc_region_;
    if (r != tp_.rank && cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    xc_ll_[r] = (uint2*)base;
  }
  CTB_CUDA(cudaMemcpy(d_ok, &ok, 4, cudaMemcpyHostToDevice));
  nccl.check(nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, (ncclComm_t)tp_.comm, stream_), "all-reduce");
  CTB_CUDA(cudaStreamSynchronize(stream_));
  CTB_CUDA(cudaMemcpy(&ok, d_ok, 4, cudaMemcpyDeviceToHost));
  tp_peer_ = ok != 0;
  if (!tp_peer_ && tp_.rank == 0 && !getenv("CTB_TP_NCCL"))
    fprintf(stderr, "ctransformers-b200: peer memory between the ranks is not available; the tensor-parallel exchange uses NCCL all-reduce\n");
}

void Engine::build_ops() {
  // nh_ / nkv_ / nff_ are this rank's share (the whole model without tensor parallelism)
  const int n_embd = hp_.n_embd, hd = hp_.head_dim(), n_kv = nkv_, gqa = nkv_ * hp_.head_dim(), qw = nh_ * hp_.head_dim();
  const bool tp = tp_.world > 1;
  const bool tp_lead = tp_.rank == 0;   // the rank whose partial sum carries the residual
  int n_xchg = 0;                       // exchanges so far in the program
  bool xchg_due = false;                // fused mode: the next mat-vec phase consumes the exchange the last one produced
  float* x_sum_into = nullptr;
  auto fill_xc = [&](XchgParams& xc, int role) {
    xc.world = tp_.world; xc.rank = tp_.rank; xc.index = n_xchg; xc.n = n_embd; xc.role = role;
    for (int r = 0; r < tp_.world; r++) xc.ll[r] = xc_ll_[r];
  };
  auto push_xchg = [&](float* buf) {    // all-reduce of a row-parallel mat-vec's partial sums (+ the residual, once)
    if (tp_peer_) {                     // fused: that phase's epilogue sends its rows to every rank; the next phase sums them into buf
      fill_xc(ops_.back().ph.xc, 2);
      xchg_due = true;
      x_sum_into = buf;
      return;
    }
    StepOp op{};
    op.ph.kind = PH_XCHG;
    op.ph.em.out = buf; op.ph.em.K = n_embd;
    ops_.push_back(op);
  };
  auto take_input = [&](MVParams& p) {   // fused: the input is the sum of the ranks' vectors of the exchange that is due
    if (!xchg_due) return;
    p.x = (const float*)xc_ll_[tp_.rank]; p.x_mode = 2; p.x_parts = tp_.world; p.x_stride = n_embd; p.sum_out = x_sum_into;
  };
  auto mark_exchange = [&](size_t first_op) {
    if (!xchg_due) return;
    if (ops_.size() != first_op + 1) throw std::runtime_error("tensor parallel (fused exchange): the consumer of an exchange must be one mat-vec phase");
    fill_xc(ops_[first_op].ph.xc, 1);
    n_xchg++;
    xchg_due = false; x_sum_into = nullptr;
  };
  const float kq_scale = 1.0f / sqrtf((float)n_embd / (float)hp_.n_head);
  ops_.clear();
  {
    StepOp op{};
    op.ph.kind = PH_EMBED;
    op.ph.em.table = tok_embd_; op.ph.em.row_bytes = tok_row_bytes_; op.ph.em.tokens = d_state_; op.ph.em.out = xa_;
    op.ph.em.type = tok_type_; op.ph.em.K = n_embd; op.ph.em.n_vocab = hp_.n_vocab;
    ops_.push_back(op);
  }
  float* x = xa_;
  float* y = xb_;
  for (int il = 0; il < hp_.n_layer; il++) {
    const LayerW& L = layers_[il];
    uint16_t* kc = kc_ + (size_t)il * hp_.n_ctx * gqa;
    uint16_t* vc = vc_ + (size_t)il * gqa * kv_ctx_pad(hp_.n_ctx);
    AttnParams ap{};
    ap.kc = kc; ap.vc = vc; ap.out = attn_; ap.exp_tab = exp_tab_; ap.state = d_state_; ap.kq_scale = kq_scale;
    ap.n_head = nh_; ap.n_kv = n_kv; ap.hd = hd; ap.n_ctx = hp_.n_ctx; ap.rope = rope_; ap.neox = hp_.falcon ? 1 : 0;
    auto push_attn = [&]() {
      StepOp op{};
      op.ph.kind = PH_ATTN;
      op.ph.at = ap;
      ops_.push_back(op);
    };

    if (!hp_.falcon) {
      float* q = qkv_; float* k = qkv_ + qw; float* v = qkv_ + qw + gqa;
      {  // attention_norm + wq/wk/wv
        MVParams p{};
        p.x = x; p.norm_w = L.attn_norm; p.norm_mode = NORM_RMS; p.eps = hp_.eps; p.K = n_embd;
        const size_t first_op = ops_.size();
        take_input(p);
        const DevMat* ws[3] = {&L.wq, &L.wk, &L.wv};
        float* outs[3] = {q, k, v};
        bool done[3] = {false, false, false};
        ap.q = q; ap.k = k; ap.v = v; ap.q_stride = qw; ap.kv_stride = gqa;
        for (int i = 0; i < 3; i++) {   // group tensors that share an activation format into one launch
          if (done[i]) continue;
          p.act = act_format_for(ws[i]->type); p.nseg = 0;
          for (int j = i; j < 3; j++)
            if (!done[j] && act_format_for(ws[j]->type) == p.act) { p.seg[p.nseg++] = seg(*ws[j], outs[j]); done[j] = true; }
          push_matvec(p, MVK_QKV);
        }
        mark_exchange(first_op);
      }
      push_attn();
      {  // wo + residual
        MVParams p{};
        p.x = attn_; p.norm_mode = NORM_NONE; p.K = qw; p.act = act_format_for(L.wo.type); p.nseg = 1;
        p.seg[0] = (!tp || tp_lead) ? seg(L.wo, y, EPI_ADD, x) : seg(L.wo, y);
        push_matvec(p, MVK_WO);
        if (tp) push_xchg(y);
      }
      {  // ffn_norm + gate and up projections as two independent row sets: silu(gate) is stored, the product with up is formed
         // where ffn_down stages its input
        MVParams p{};
        p.x = y; p.norm_w = L.ffn_norm; p.norm_mode = NORM_RMS; p.eps = hp_.eps; p.K = n_embd;
        const size_t first_op = ops_.size();
        take_input(p);
        p.act = act_format_for(L.w1.type); p.nseg = 2;
        p.seg[0] = seg(L.w1, ffn_, EPI_SILU); p.seg[1] = seg(L.w3, ffn2_);
        push_matvec(p, MVK_UP);
        mark_exchange(first_op);
      }
      {  // w2 on silu(gate)*up, + residual
        MVParams p{};
        p.x = ffn_; p.x2 = ffn2_; p.x_mode = 1; p.norm_mode = NORM_NONE; p.K = nff_; p.act = act_format_for(L.w2.type); p.nseg = 1;
        p.seg[0] = (!tp || tp_lead) ? seg(L.w2, x, EPI_ADD, y) : seg(L.w2, x);
        push_matvec(p, MVK_DOWN);
        if (tp) push_xchg(x);
      }
      // x now holds the next layer's input
    } else {
      const int qkv_w = (hp_.n_head + 2 * n_kv) * hd;
      float* q = qkv_; float* k = qkv_ + (size_t)hp_.n_head * hd; float* v = k + (size_t)n_kv * hd;
      const bool two_norms = L.attn_norm2 != nullptr;
      const bool fuse = !two_norms && act_format_for(L.wqkv.type) == act_format_for(L.w3.type);
      {  // LayerNorm + wqkv (+ ffn_up → GELU when it shares the normed input)
        MVParams p{};
        p.x = x; p.norm_mode = NORM_LAYER; p.eps = hp_.eps; p.K = n_embd;
        p.norm_w = two_norms ? L.attn_norm2 : L.attn_norm; p.norm_b = two_norms ? L.attn_norm2_b : L.attn_norm_b;
        p.act = act_format_for(L.wqkv.type); p.nseg = 1;
        p.seg[0] = seg(L.wqkv, qkv_);
        if (fuse) { p.seg[1] = seg(L.w3, ffn_, EPI_GELU); p.nseg = 2; }
        ap.q = q; ap.k = k; ap.v = v; ap.q_stride = qkv_w; ap.kv_stride = qkv_w;
        push_matvec(p, MVK_QKV);
      }
      if (!fuse) {
        MVParams p{};
        p.x = x; p.norm_mode = NORM_LAYER; p.eps = hp_.eps; p.K = n_embd; p.norm_w = L.attn_norm; p.norm_b = L.attn_norm_b;
        p.act = act_format_for(L.w3.type); p.nseg = 1;
        p.seg[0] = seg(L.w3, ffn_, EPI_GELU);
        push_matvec(p, MVK_UP);
      }
      push_attn();
      {  // attention output projection
        MVParams p{};
        p.x = attn_; p.norm_mode = NORM_NONE; p.K = n_embd; p.act = act_format_for(L.wo.type); p.nseg = 1;
        p.seg[0] = seg(L.wo, attn_o_);
        push_matvec(p, MVK_WO);
      }
      {  // ffn_down, then + attn_out, then + layer input (llama.cpp:2767-2771 order)
        MVParams p{};
        p.x = ffn_; p.norm_mode = NORM_NONE; p.K = hp_.n_ff; p.act = act_format_for(L.w2.type); p.nseg = 1;
        p.seg[0] = seg(L.w2, y, EPI_ADD2, attn_o_, x);
        push_matvec(p, MVK_DOWN);
      }
      std::swap(x, y);
    }
  }
  n_body_ = (int)ops_.size();
  {
    MVParams p{};
    p.x = x; p.norm_w = out_norm_; p.norm_b = out_norm_b_; p.norm_mode = hp_.falcon ? NORM_LAYER : NORM_RMS; p.eps = hp_.eps; p.K = n_embd;
    const size_t first_op = ops_.size();
    take_input(p);
    p.norm_out = d_embd_; p.act = act_format_for(output_.type); p.nseg = 1;
    p.seg[0] = seg(output_, d_logits_);
    push_matvec(p, MVK_OUT);
    mark_exchange(first_op);

  }
  {
    StepOp op{};
    op.ph.kind = PH_PICK;
    op.ph.pk.logits = d_logits_; op.ph.pk.state = d_state_; op.ph.pk.out_tokens = nullptr; op.ph.pk.n = hp_.n_vocab;   // out_tokens: set in build_graphs
    ops_.push_back(op);
  }
  // ---- the step kernel's shared-memory shape: ring slots fill what the largest activation image leaves
  bool any_stream = false;
  for (const StepOp& op : ops_) any_stream |= op.ph.kind == PH_MATVEC && op.stream;
  std::vector<Phase> phs;
  for (const StepOp& op : ops_) if (op.ph.kind != PH_XCHG && (op.ph.kind != PH_MATVEC || op.stream)) phs.push_back(op.ph);
  const StepLaunch sl = step_launch_shape(phs.data(), (int)phs.size(), sm_count_, step_max_dyn_smem());
  step_grid_ = sl.grid; step_slots_ = sl.n_slots; step_smem_ = sl.smem;
  if (const char* e = getenv("CTB_ST_SLOTS")) {   // A/B knob: fewer ring slots = less prefetch in flight
    const int want = atoi(e);
    if (want >= ST_W && want < step_slots_ && want % ST_W == 0) { step_smem_ -= (size_t)(step_slots_ - want) * ST_SLOT; step_slots_ = want; }
  }
  const bool ring_attn = st_attn_ring_ok(hp_.n_ctx, step_slots_) && !getenv("CTB_NO_RING_ATTN");
  for (StepOp& op : ops_) if (op.ph.kind == PH_ATTN) op.ph.q6 = ring_attn ? 1 : 0;
  if (any_stream && step_slots_ < ST_W) throw std::runtime_error("model rows are too long for the step kernel's shared memory");
  if (any_stream) CTB_CUDA(step_set_smem_limit(step_smem_));
  else fused_ = false;
  d_prog_ = (Phase*)alloc((ops_.size() + 1) * sizeof(Phase), 256);
  d_prog_mv_ = (Phase*)alloc((ops_.size() + 1) * sizeof(Phase), 256);
  d_bounds_ = (int*)alloc((ops_.size() + 1) * (size_t)(sm_count_ + 1) * 4, 256);      // per-CTA tile ranges, one row per phase
  d_bounds_mv_ = (int*)alloc((ops_.size() + 1) * (size_t)(sm_count_ + 1) * 4, 256);
}

void Engine::upload_prog(Phase* dst, int* dst_bounds, const std::vector<StepOp>& ops) {
  std::vector<Phase> phs(ops.size());
  for (size_t i = 0; i < ops.size(); i++) phs[i] = ops[i].ph;
  CTB_CUDA(cudaMemcpy(dst, phs.data(), phs.size() * sizeof(Phase), cudaMemcpyHostToDevice));
  std::vector<Phase> run = phs;
  for (size_t i = 0; i < ops.size(); i++) if (ops[i].ph.kind == PH_XCHG || (ops[i].ph.kind == PH_MATVEC && !ops[i].stream)) run[i].kind = -1;   // not a step-kernel phase
  const std::vector<int> b = step_bounds(run.data(), (int)run.size(), sm_count_);
  CTB_CUDA(cudaMemcpy(dst_bounds, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
}

// Enqueue ops[0, n) on stream_.  Fused mode: maximal runs of ops the step kernel can take become ONE k_step launch (a
// K-quant model: the whole token); anything else (Q4_0 / Q8_0 / F16 / F32 mat-vecs) runs as its own kernel.  Un-fused mode
// (CTB_STEP_FUSE=0): one kernel per op, K-quant mat-vecs as one-phase k_step launches.
void Engine::enqueue_ops(const std::vector<StepOp>& ops, const Phase* d_prog, const int* d_bounds, int n) {
  launches_per_step_ = 0;
  StepLaunch step_shape_;
  step_shape_.grid = step_grid_; step_shape_.n_slots = step_slots_; step_shape_.smem = step_smem_;
  auto capable = [&](const StepOp& op) { return op.ph.kind != PH_XCHG && (op.ph.kind != PH_MATVEC || op.stream); };
  int i = 0;
  while (i < n) {
    const StepOp& op = ops[i];
    if (fused_ && capable(op)) {
      int j = i;
      while (j < n && capable(ops[j])) j++;
      mark(-1);
      CTB_CUDA(launch_step(step_shape_, stream_, d_prog + i, d_bounds + (size_t)i * (sm_count_ + 1), j - i, d_sync_, false, nullptr, tp_peer_));
      launches_per_step_++;
      mark(0);
      i = j;
      continue;
    }
    mark(-1);
    switch (op.ph.kind) {
      case PH_EMBED:
        k_embed<<<1, 256, 0, stream_>>>(op.ph.em.table, op.ph.em.type, op.ph.em.row_bytes, op.ph.em.K, op.ph.em.n_vocab, op.ph.em.tokens, op.ph.em.out);
        launches_per_step_++;
        mark(3);
        break;
      case PH_ATTN: {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(nh_, 1, hp_.head_dim() / ATTN_CH); cfg.blockDim = dim3(ATTN_THREADS);
        cfg.dynamicSmemBytes = attn_smem_bytes(hp_.n_ctx, hp_.head_dim()); cfg.stream = stream_;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_ ? 1 : 0;
        CTB_CUDA(cudaLaunchKernelEx(&cfg, k_attn, op.ph.at));
        launches_per_step_++;
        mark(1);
      } break;
      case PH_XCHG:
        tp_all_reduce(op.ph.em.out, op.ph.em.K);
        mark(3);
        break;
      case PH_PICK:
        k_argmax<<<1, 1024, 0, stream_>>>(d_logits_, hp_.n_vocab, d_state_ + 4);
        k_advance<<<1, 1, 0, stream_>>>(d_state_, d_tokens_out_);
        launches_per_step_ += 2;
        mark(3);
        break;
      default:
        if (op.stream) {
          CTB_CUDA(launch_step(step_shape_, stream_, d_prog + i, d_bounds + (size_t)i * (sm_count_ + 1), 1, d_sync_));
        } else {
          const MVLaunch L = matvec_launch_shape(op.ph.mv, sm_count_);
          CTB_CUDA(launch_matvec_kernel(L, stream_, op.ph.mv, pdl_));
        }
        launches_per_step_++;
        mark(0);
    }
    i++;
  }
  CTB_CUDA(cudaGetLastError());
}

void Engine::mark(int kind) {
  if (!profiling_) return;
  cudaEvent_t e;
  CTB_CUDA(cudaEventCreate(&e));
  CTB_CUDA(cudaEventRecord(e, stream_));
  prof_ev_.push_back(e);
  prof_kind_.push_back(kind);
}

// One eager decode step, one kernel per op (un-fused), a CUDA event around every kernel: the kernel classes' share of a step.
int Engine::profile_step(int token, int n_past, double ms_by_kind[4], int count_by_kind[4]) {
  DeviceGuard dev_guard(device_);
  if (tp_.world > 1) throw std::runtime_error("not available in tensor-parallel mode (every rank must run the same launches)");
  spec_pending_ = false; spec_deferred_ = false; spec_pos_ = -1; spec_streak_ = 0;
  if (h_state_cap_ < 1) { h_state_cap_ = 512; CTB_CUDA(cudaMallocHost(&h_state_, (size_t)h_state_cap_ * 16)); }
  h_state_[0] = token; h_state_[1] = n_past; h_state_[2] = 0; h_state_[3] = n_past + 1;
  CTB_CUDA(cudaMemcpyAsync(d_state_, h_state_, 16, cudaMemcpyHostToDevice, stream_));
  const long keep = launches_per_step_;
  const bool keep_fused = fused_;
  profiling_ = true; fused_ = false;
  try { enqueue_ops(ops_, d_prog_, d_bounds_, n_body_ + 1); } catch (...) { profiling_ = false; fused_ = keep_fused; throw; }
  profiling_ = false; fused_ = keep_fused;
  launches_per_step_ = keep;
  CTB_CUDA(cudaStreamSynchronize(stream_));
  int n = 0;
  for (size_t i = 1; i < prof_ev_.size(); i++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, prof_ev_[i - 1], prof_ev_[i]);
    const int k = prof_kind_[i];
    if (k >= 0 && k < 4) { ms_by_kind[k] += ms; count_by_kind[k]++; n++; }
  }
  for (cudaEvent_t e : prof_ev_) cudaEventDestroy(e);
  prof_ev_.clear(); prof_kind_.clear();
  return n;
}

// The step's mat-vec phases alone (same kernel, same parameters, same order; no attention / embedding / pick), replayed as a
// CUDA graph: their duration under in-step conditions is what bench.py's roofline for the mat-vec uses.  mask: bit k set =
// keep the mat-vecs of kind k (0 = all).  with_attn: keep the attention phases too (times the dependency chain as it is).
double Engine::time_matvec_only(int reps, long* launches, unsigned mask) {
  if (!mask) mask = ~0u;
  if (tp_.world > 1) throw std::runtime_error("not available in tensor-parallel mode (every rank must run the same launches)");
  spec_pending_ = false; spec_deferred_ = false; spec_pos_ = -1; spec_streak_ = 0;
  DeviceGuard dev_guard(device_);
  std::vector<StepOp> sel;
  for (int i = 0; i <= n_body_; i++)
    if (ops_[i].ph.kind == PH_MATVEC && ((mask >> ops_[i].mvk) & 1)) sel.push_back(ops_[i]);
  if (sel.empty()) { if (launches) *launches = 0; return 0.0; }
  upload_prog(d_prog_mv_, d_bounds_mv_, sel);
  cudaStream_t user = stream_, cap;
  CTB_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  const long keep = launches_per_step_;
  cudaGraphExec_t ex = nullptr;
  stream_ = cap;
  try {
    cudaGraph_t g;
    CTB_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    enqueue_ops(sel, d_prog_mv_, d_bounds_mv_, (int)sel.size());
    CTB_CUDA(cudaStreamEndCapture(cap, &g));
    CTB_CUDA(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
  } catch (...) { stream_ = user; launches_per_step_ = keep; cudaStreamDestroy(cap); throw; }
  if (launches) *launches = (long)sel.size();
  stream_ = user; launches_per_step_ = keep;
  cudaStreamDestroy(cap);
  cudaEvent_t e0, e1;
  CTB_CUDA(cudaEventCreate(&e0)); CTB_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) CTB_CUDA(cudaGraphLaunch(ex, stream_));
  CTB_CUDA(cudaEventRecord(e0, stream_));
  for (int i = 0; i < reps; i++) CTB_CUDA(cudaGraphLaunch(ex, stream_));
  CTB_CUDA(cudaEventRecord(e1, stream_));
  CTB_CUDA(cudaStreamSynchronize(stream_));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaGraphExecDestroy(ex);
  return (double)ms / reps;
}

// One fused decode step with the step kernel stamping %globaltimer per phase and CTA (4 stamps: barrier passed, input staged,
// first weight item ready, phase done + 4 stamps of the grid barrier in front of the phase).  out: n_phases x {kind, mvk} then n_phases x n_cta x 8 stamps; returns n_phases or -(words needed).
long Engine::trace_step(int token, int n_past, unsigned long long* out, long cap_words) {
  DeviceGuard dev_guard(device_);
  if (tp_.world > 1) throw std::runtime_error("not available in tensor-parallel mode (every rank must run the same launches)");
  spec_pending_ = false; spec_deferred_ = false; spec_pos_ = -1; spec_streak_ = 0;
  const int n = n_body_ + 1;
  for (int i = 0; i < n; i++)
    if (ops_[i].ph.kind == PH_MATVEC && !ops_[i].stream) return 0;
  const long need = 2L * n + 8L * n * step_grid_;
  if (need > cap_words) return -need;
  unsigned long long* buf = nullptr;
  CTB_CUDA(cudaMalloc(&buf, (size_t)n * step_grid_ * 64));
  CTB_CUDA(cudaMemset(buf, 0, (size_t)n * step_grid_ * 64));
  StepLaunch L;
  L.grid = step_grid_; L.n_slots = step_slots_; L.smem = step_smem_;
  if (h_state_cap_ < 1) { h_state_cap_ = 512; CTB_CUDA(cudaMallocHost(&h_state_, (size_t)h_state_cap_ * 16)); }
  for (int rep = 0; rep < 3; rep++) {   // the last (warm) run is the one read back
    h_state_[0] = token; h_state_[1] = n_past; h_state_[2] = 0; h_state_[3] = n_past + 1;
    CTB_CUDA(cudaMemcpyAsync(d_state_, h_state_, 16, cudaMemcpyHostToDevice, stream_));
    CTB_CUDA(launch_step(L, stream_, d_prog_, d_bounds_, n, d_sync_, false, buf));
    CTB_CUDA(cudaStreamSynchronize(stream_));
  }
  for (int i = 0; i < n; i++) { out[2 * i] = (unsigned long long)ops_[i].ph.kind; out[2 * i + 1] = (unsigned long long)ops_[i].mvk; }
  const cudaError_t e = cudaMemcpy(out + 2 * n, buf, (size_t)n * step_grid_ * 64, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  CTB_CUDA(e);
  return n;
}

void Engine::destroy_graphs() {
  if (graph_full_) cudaGraphExecDestroy(graph_full_);
  if (graph_nolog_) cudaGraphExecDestroy(graph_nolog_);
  if (graph_greedy_) cudaGraphExecDestroy(graph_greedy_);
  graph_full_ = graph_nolog_ = graph_greedy_ = nullptr;
}

void Engine::build_graphs() {
  destroy_graphs();
  if (!d_tokens_out_) {
    tokens_out_cap_ = std::max(hp_.n_ctx, 4096);
    CTB_CUDA(cudaMalloc(&d_tokens_out_, (size_t)tokens_out_cap_ * 4));
    CTB_CUDA(cudaMallocHost(&h_tokens_out_, (size_t)tokens_out_cap_ * 4));
  }
  cudaStream_t user = stream_;
  cudaStream_t cap;
  CTB_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  stream_ = cap;
  ops_.back().ph.pk.out_tokens = d_tokens_out_;
  upload_prog(d_prog_, d_bounds_, ops_);
  auto capture = [&](bool logits, bool greedy) {
    cudaGraph_t g;
    CTB_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    // (fused tensor-parallel exchange: the head phase consumes the last layer's exchange, so every program contains it)
    enqueue_ops(ops_, d_prog_, d_bounds_, n_body_ + (logits || tp_peer_ ? 1 : 0) + (greedy ? 1 : 0));
    CTB_CUDA(cudaStreamEndCapture(cap, &g));
    cudaGraphExec_t ex;
    CTB_CUDA(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
    return ex;
  };
  try {
    graph_nolog_ = capture(false, false);
    graph_greedy_ = capture(true, true);
    graph_full_ = capture(true, false);
  } catch (...) {
    stream_ = user;
    cudaStreamDestroy(cap);
    throw;
  }
  stats.launches = launches_per_step_;
  stream_ = user;
  cudaStreamDestroy(cap);
}

// After an eval: pick the greedy next token on the device and — once the caller has proven to decode greedily (its last
// tokens were exactly those picks) — run the step for that token right away, while the host is still sampling and crossing
// the FFI.  The next eval() that asks for exactly this token at this position only has to fetch the result; any other request
// simply runs after it (stream order) and overwrites the same KV slot, so a wrong guess costs time, never correctness.
void Engine::after_eval(int next_pos) {
  spec_pending_ = false;
  spec_deferred_ = false;
  spec_pos_ = -1;
  // the look-ahead step writes K/V slot next_pos: only when nothing valid can live there (append-only decoding).  A caller
  // that re-evaluates an earlier position and later continues past it keeps its cache contents.
  if (!spec_on_ || next_pos >= hp_.n_ctx || next_pos < kv_high_) return;
  k_argmax<<<1, 1024, 0, stream_>>>(d_logits_, hp_.n_vocab, d_state_ + 4);
  k_advance<<<1, 1, 0, stream_>>>(d_state_, d_tokens_out_);
  CTB_CUDA(cudaMemcpyAsync(h_spec_tok_, d_state_ + 4, 8, cudaMemcpyDeviceToHost, stream_));   // {greedy pick, how many logits equal the maximum}
  CTB_CUDA(cudaEventRecord(ev_pick_, stream_));
  spec_pos_ = next_pos;
  if (spec_streak_ >= 2) {
    // The persistent step kernel fills every SM, so whatever is enqueued behind the look-ahead step waits for all of it.  A
    // caller whose sample() runs the device sampler (sample_gpu.cuh) therefore gets the look-ahead launched from there, right
    // behind the sampler kernel; a caller who takes the greedy pick needs no kernel and gets it launched here, before the wait.
    if (sampler_mode_) spec_deferred_ = true;
    else {
      CTB_CUDA(cudaGraphLaunch(graph_full_, stream_));
      spec_pending_ = true;
    }
  }
}

void Engine::launch_deferred_spec() {
  if (!spec_deferred_) return;
  spec_deferred_ = false;
  CTB_CUDA(cudaGraphLaunch(graph_full_, stream_));
  spec_pending_ = true;
}

// The greedy pick of the last eval (what top_k = 1 without a repetition penalty selects, llama.cpp:3832-3857 + 4215-4240: one
// candidate survives, the draw is certain) if the engine computed it: id, or -1 when it did not (look-ahead off, context full,
// non-appending eval) or when several logits share the maximum (std::partial_sort's choice among equals is the reference's).
int Engine::greedy_pick() {
  if (spec_pos_ < 0) return -1;
  DeviceGuard dev_guard(device_);
  CTB_CUDA(cudaEventSynchronize(ev_pick_));
  sampler_mode_ = false;
  launch_deferred_spec();
  return h_spec_tok_[1] == 1 ? h_spec_tok_[0] : -1;
}

void Engine::eval(const int* tokens, int n, int n_past) {
  if (n <= 0) return;
  std::vector<int> pos(n), nt(n, n_past + n);   // n_total: row length of this eval's attention mat-muls
  for (int i = 0; i < n; i++) pos[i] = n_past + i;
  eval_list(tokens, pos.data(), nt.data(), n);
}

void Engine::decode_one(int token, int pos, int n_total, bool with_logits) {
  if (h_state_cap_ < 1) { h_state_cap_ = 512; CTB_CUDA(cudaMallocHost(&h_state_, (size_t)h_state_cap_ * 16)); }
  int* st = h_state_ + (size_t)(h_state_next_ % h_state_cap_) * 4;
  h_state_next_++;
  st[0] = token; st[1] = pos; st[2] = 0; st[3] = n_total;
  CTB_CUDA(cudaMemcpyAsync(d_state_, st, 16, cudaMemcpyHostToDevice, stream_));
  CTB_CUDA(cudaGraphLaunch(with_logits ? graph_full_ : graph_nolog_, stream_));
}

void Engine::host_views() {
  eager_ = true;
  if (host_fresh_) return;
  DeviceGuard dev_guard(device_);
  CTB_CUDA(cudaMemcpyAsync(h_logits_, d_logits_keep_, (size_t)hp_.n_vocab * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaMemcpyAsync(h_embd_, d_embd_keep_, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaStreamSynchronize(stream_));
  host_fresh_ = true;
}

std::vector<float> Engine::logits_copy() {
  std::vector<float> v((size_t)hp_.n_vocab);
  if (host_fresh_) { memcpy(v.data(), h_logits_, v.size() * 4); return v; }
  DeviceGuard dev_guard(device_);
  CTB_CUDA(cudaMemcpyAsync(v.data(), d_logits_keep_, v.size() * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaStreamSynchronize(stream_));
  return v;
}

int Engine::topk_candidates(const int* last, int n_last, float penalty, int k, int* ids, float* logits) {
  if (n_last > SG_MAX_LAST || k < 1 || k > SG_MAX_OUT / 2) return -1;
  DeviceGuard dev_guard(device_);
  if (!d_sample_) {
    d_sample_ = (SampleGpuOut*)alloc(sizeof(SampleGpuOut));
    d_last_ = (int*)alloc(SG_MAX_LAST * 4 + 16);
    CTB_CUDA(cudaMallocHost(&h_sample_, sizeof(SampleGpuOut) + SG_MAX_LAST * 4));
  }
  if (!ev_sample_) CTB_CUDA(cudaEventCreateWithFlags(&ev_sample_, cudaEventDisableTiming));
  sampler_mode_ = true;
  int* h_last = (int*)(h_sample_ + 1);
  for (int i = 0; i < n_last; i++) h_last[i] = last[i];
  if (n_last > 0) CTB_CUDA(cudaMemcpyAsync(d_last_, h_last, (size_t)n_last * 4, cudaMemcpyHostToDevice, stream_));
  k_sample_topk<<<1, SG_THREADS, 0, stream_>>>(d_logits_keep_, hp_.n_vocab, d_last_, n_last, penalty, std::min(k, hp_.n_vocab), d_sample_);
  CTB_CUDA(cudaGetLastError());
  CTB_CUDA(cudaMemcpyAsync(h_sample_, d_sample_, sizeof(SampleGpuOut), cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaEventRecord(ev_sample_, stream_));
  launch_deferred_spec();                          // the look-ahead step runs while the host finishes the draw
  CTB_CUDA(cudaEventSynchronize(ev_sample_));
  const int n = h_sample_->count;
  if (n < 0 || n > SG_MAX_OUT) return -1;
  for (int i = 0; i < n; i++) { ids[i] = h_sample_->id[i]; logits[i] = h_sample_->logit[i]; }
  return n;
}

void Engine::finish_eval(int next_pos, bool hit) {
  // the look-ahead step (after_eval) overwrites d_logits_ / d_embd_: keep this eval's results where a late request finds them
  CTB_CUDA(cudaMemcpyAsync(d_logits_keep_, d_logits_, (size_t)hp_.n_vocab * 4, cudaMemcpyDeviceToDevice, stream_));
  CTB_CUDA(cudaMemcpyAsync(d_embd_keep_, d_embd_, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToDevice, stream_));
  if (eager_) {
    CTB_CUDA(cudaMemcpyAsync(h_logits_, d_logits_, (size_t)hp_.n_vocab * 4, cudaMemcpyDeviceToHost, stream_));
    CTB_CUDA(cudaMemcpyAsync(h_embd_, d_embd_, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToHost, stream_));
  }
  host_fresh_ = eager_;
  CTB_CUDA(cudaEventRecord(ev1_, stream_));
  after_eval(next_pos);
  CTB_CUDA(cudaEventSynchronize(ev1_));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0_, ev1_);
  stats.last_eval_ms = ms;
  stats.spec_hits += hit ? 1 : 0;
}

void Engine::eval_list(const int* tokens, const int* pos, const int* n_total, int n) {
  if (n <= 0) return;
  DeviceGuard dev_guard(device_);
  if (h_state_cap_ < 1) { h_state_cap_ = 512; CTB_CUDA(cudaMallocHost(&h_state_, (size_t)h_state_cap_ * 16)); }
  bool hit = false;
  if (spec_pos_ >= 0) {
    const bool was_pending = spec_pending_;   // (a deferred look-ahead nobody launched is simply dropped)
    spec_deferred_ = false;
    bool guessed = false;
    if (n == 1 && pos[0] == spec_pos_ && n_total[0] == pos[0] + 1) {
      CTB_CUDA(cudaEventSynchronize(ev_pick_));
      guessed = h_spec_tok_[0] == tokens[0];
    }
    spec_streak_ = guessed ? spec_streak_ + 1 : 0;
    hit = guessed && was_pending;
    spec_pending_ = false;
    spec_pos_ = -1;
  }
  CTB_CUDA(cudaEventRecord(ev0_, stream_));
  for (int i = 0; i < n; i++) kv_high_ = std::max(kv_high_, pos[i] + 1);
  if (!hit) {
    int i = 0;
    while (i < n) {
      // a run of consecutive positions goes through the batched kernel, PB_T tokens per launch
      int j = i + 1;
      while (j < n && pos[j] == pos[j - 1] + 1 && pos[j] < hp_.n_ctx) j++;
      if (j - i >= prefill_min_ && pos[i] < hp_.n_ctx && ensure_prefill()) {
        for (int b = i; b < j; b += PB_T) {
          const int m = std::min(PB_T, j - b);
          prefill_batch(tokens + b, pos + b, n_total + b, m, b + m == n);
        }
      } else {
        for (int k = i; k < j; k++) {
          decode_one(tokens[k], pos[k], n_total[k], k == n - 1);
          if ((k - i) % 128 == 127) CTB_CUDA(cudaStreamSynchronize(stream_));   // keeps the pinned state ring from wrapping under the GPU
        }
      }
      i = j;
    }
  }   // else: the step for this token at this position is already in the stream
  finish_eval(pos[n - 1] + 1, hit);
}

bool Engine::ensure_prefill() {
  if (!prefill_on_ || tp_.world > 1) return false;
  if (pf_ && pf_->tried) return pf_->ok;
  if (!pf_) pf_ = new PrefillState();
  PrefillState& P = *pf_;
  P.tried = true;
  for (int i = 0; i < n_body_; i++)
    if (ops_[i].ph.kind == PH_MATVEC && !ops_[i].stream) return false;   // a non-K-quant layer matrix: single-token path only
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    CTB_CUDA(cudaMalloc(&p, bytes));
    P.bufs.push_back(p);
    CTB_CUDA(cudaMemset(p, 0, bytes));
    return p;
  };
  const int n_embd = hp_.n_embd, gqa = hp_.n_embd_gqa();
  const int qkvw = n_embd + 2 * gqa;
  // decode buffer -> (batched buffer, floats between token rows)
  struct Map { const float* lo; size_t n; float* b; int ld; };
  std::vector<Map> maps;
  auto add = [&](const float* dec, size_t n) { maps.push_back({dec, n, (float*)dalloc((size_t)PB_T * n * 4), (int)n}); };
  add(xa_, n_embd); add(xb_, n_embd); add(qkv_, qkvw); add(attn_, n_embd); add(attn_o_, n_embd); add(ffn_, hp_.n_ff); add(ffn2_, hp_.n_ff);
  auto bat = [&](const float* p, int& ld) -> float* {
    if (!p) { ld = 0; return nullptr; }
    for (const Map& m : maps)
      if (p >= m.lo && p < m.lo + m.n) { ld = m.ld; return m.b + (p - m.lo); }
    throw std::runtime_error("prefill: pointer outside the step workspace");
  };
  P.d_state = (int*)dalloc((PB_T * 4 + 4) * 4);
  CTB_CUDA(cudaMallocHost(&P.h_state, (size_t)PF_RING * (PB_T * 4 + 4) * 4));
  std::vector<PPhase> prog;
  int K_max = 0;
  for (int i = 0; i < n_body_; i++) {
    const StepOp& op = ops_[i];
    PPhase ph{};
    ph.state = P.d_state;
    if (op.ph.kind == PH_EMBED) {
      ph.kind = PP_EMBED;
      ph.em = op.ph.em;
      int ld;
      ph.em.out = bat(op.ph.em.out, ld);
      prog.push_back(ph);
    } else if (op.ph.kind == PH_ATTN) {
      ph.at = op.ph.at;
      int ld, ldo;
      ph.at.q = bat(op.ph.at.q, ld); ph.at.k = bat(op.ph.at.k, ld); ph.at.v = bat(op.ph.at.v, ld);
      ph.at.q_stride = ld; ph.at.kv_stride = ld;
      ph.at.out = bat(op.ph.at.out, ldo);
      ph.at.state = P.d_state;
      ph.kind = PP_KV; prog.push_back(ph);
      ph.kind = PP_ATTN; prog.push_back(ph);
    } else {
      const MVParams& m = op.ph.mv;
      K_max = std::max(K_max, m.K);
      ph.mv = m;
      ph.mv.norm_out = nullptr;
      ph.mv.x = bat(m.x, ph.x_ld);
      ph.mv.x2 = bat(m.x2, ph.x2_ld);
      ph.qbuf = (uint8_t*)dalloc(pb_qbuf_bytes(m.K));   // one buffer per phase: nothing stale can sit in an L1
      ph.kind = PP_QUANT; prog.push_back(ph);
      for (int sgi = 0; sgi < m.nseg; sgi++) {
        ph.mv.seg[sgi].out = bat(m.seg[sgi].out, ph.out_ld[sgi]);
        ph.mv.seg[sgi].res = bat(m.seg[sgi].res, ph.res_ld[sgi]);
        ph.mv.seg[sgi].res2 = bat(m.seg[sgi].res2, ph.res2_ld[sgi]);
      }
      ph.kind = PP_GEMM; prog.push_back(ph);
    }
  }
  int ld;
  P.x_final = bat(ops_[n_body_].ph.mv.x, ld);
  P.n_phases = (int)prog.size();
  P.d_prog = (PPhase*)dalloc((prog.size() + 1) * sizeof(PPhase));
  CTB_CUDA(cudaMemcpy(P.d_prog, prog.data(), prog.size() * sizeof(PPhase), cudaMemcpyHostToDevice));
  const size_t work = pb_work_bytes(K_max, hp_.n_ctx, hp_.head_dim()), room = pstep_max_dyn_smem();
  if (work + 4 * (size_t)ST_SLOT > room) return false;
  P.n_slots = (int)std::min<size_t>(ST_MAX_SLOTS, (room - work) / ST_SLOT) / PB_TEAMS * PB_TEAMS;   // whole per-team sub-rings
  P.smem = (size_t)P.n_slots * ST_SLOT + work;
  CTB_CUDA(pstep_set_smem_limit(P.smem));
  P.ok = true;
  return true;
}

// n <= PB_T tokens at consecutive positions through all layers in one launch; `last`: the list ends here, so the head
// mat-vec (logits + final-norm hidden state of the last token) follows on the single-token kernel.
void Engine::prefill_batch(const int* tokens, const int* pos, const int* n_total, int n, bool last) {
  PrefillState& P = *pf_;
  if (P.h_next % PF_RING == PF_RING - 1) CTB_CUDA(cudaStreamSynchronize(stream_));   // pinned state ring
  int* st = P.h_state + (size_t)(P.h_next++ % PF_RING) * (PB_T * 4 + 4);
  for (int i = 0; i < PB_T; i++) {
    const int k = std::min(i, n - 1);
    st[i * 4] = tokens[k]; st[i * 4 + 1] = pos[k]; st[i * 4 + 2] = 0; st[i * 4 + 3] = n_total[k];
  }
  st[PB_T * 4] = n;
  CTB_CUDA(cudaMemcpyAsync(P.d_state, st, (PB_T * 4 + 4) * 4, cudaMemcpyHostToDevice, stream_));
  CTB_CUDA(launch_pstep(step_grid_, P.n_slots, P.smem, stream_, P.d_prog, P.n_phases, d_sync_));
  if (last) {
    const StepOp& head = ops_[n_body_];
    CTB_CUDA(cudaMemcpyAsync(const_cast<float*>(head.ph.mv.x), P.x_final + (size_t)(n - 1) * hp_.n_embd, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToDevice, stream_));
    const bool keep = fused_;
    fused_ = false;                                    // just this one op, the way the un-fused schedule launches it
    std::vector<StepOp> one(1, head);
    const long keepl = launches_per_step_;
    try { enqueue_ops(one, d_prog_ + n_body_, d_bounds_ + (size_t)n_body_ * (sm_count_ + 1), 1); } catch (...) { fused_ = keep; throw; }
    fused_ = keep;
    launches_per_step_ = keepl;
  }
}

double Engine::decode_greedy(int first_token, int n_past, int n_steps, int* out_tokens) {
  if (n_steps <= 0) return 0.0;
  spec_pending_ = false; spec_deferred_ = false; spec_pos_ = -1; spec_streak_ = 0;
  if (n_steps > tokens_out_cap_) throw std::runtime_error("decode_greedy: too many steps");
  if (n_past + n_steps > hp_.n_ctx) throw std::runtime_error("decode_greedy: would run past the context length");
  kv_high_ = std::max(kv_high_, n_past + n_steps);
  DeviceGuard dev_guard(device_);
  if (h_state_cap_ < 1) { h_state_cap_ = 512; CTB_CUDA(cudaMallocHost(&h_state_, (size_t)h_state_cap_ * 16)); }
  h_state_[0] = first_token; h_state_[1] = n_past; h_state_[2] = 0; h_state_[3] = n_past + 1;
  CTB_CUDA(cudaMemcpyAsync(d_state_, h_state_, 16, cudaMemcpyHostToDevice, stream_));
  CTB_CUDA(cudaEventRecord(ev0_, stream_));
  for (int s = 0; s < n_steps; s++) CTB_CUDA(cudaGraphLaunch(graph_greedy_, stream_));
  CTB_CUDA(cudaEventRecord(ev1_, stream_));
  CTB_CUDA(cudaMemcpyAsync(h_tokens_out_, d_tokens_out_, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaMemcpyAsync(h_logits_, d_logits_, (size_t)hp_.n_vocab * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaMemcpyAsync(h_embd_, d_embd_, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToHost, stream_));
  CTB_CUDA(cudaMemcpyAsync(d_logits_keep_, d_logits_, (size_t)hp_.n_vocab * 4, cudaMemcpyDeviceToDevice, stream_));
  CTB_CUDA(cudaMemcpyAsync(d_embd_keep_, d_embd_, (size_t)hp_.n_embd * 4, cudaMemcpyDeviceToDevice, stream_));
  CTB_CUDA(cudaStreamSynchronize(stream_));
  host_fresh_ = true;
  memcpy(out_tokens, h_tokens_out_, (size_t)n_steps * 4);
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0_, ev1_);
  stats.last_eval_ms = ms;
  return ms;
}

}  // namespace ctb
