// Eval engine: static device arena + fixed op schedule for the two graph shapes of the path
// (llm_build_llama / llm_build_falcon, reference models/ggml/llama.cpp:2162-2798) driven like
// llama_eval_internal (llama.cpp:2835-2981): last token's logits and post-final-norm hidden state end
// up in host memory owned by the LLM object.
//
// B200 design: weights repacked once into the stream layout / coalesced planes (device_types.cuh); KV cache fp16;
// a static op list per token whose K-quant mat-vecs, attention, embedding row and greedy pick run as the phases of
// ONE persistent kernel (stream.cuh), replayed as a CUDA graph with {token, n_past} as device scalars; no per-eval
// graph build, no allocator, no thread pool.
#pragma once
#include <cmath>
#include <memory>
#include <string>
#include <vector>

#include "device_types.cuh"
#include "gguf.hpp"

namespace ctb {

// Tensor-sharded mode (BASELINE configs[4]): one process per GPU; rank r of `world` owns the query heads
// [head0, head1) (with their KV heads), the n_ff range [ff0, ff1) — always whole 256-element blocks — and a replica of the
// embedding table, the norms and the output head.  Column-parallel wq / wk / wv / w1 / w3, row-parallel wo / w2, two
// all-reduces of n_embd floats per layer (NCCL over NVLink, captured in the step's CUDA graph).
struct TPShard {
  int rank = 0, world = 1;
  int head0 = 0, head1 = 0, kv0 = 0, kv1 = 0, ff0 = 0, ff1 = 0;
  void* comm = nullptr;   // ncclComm_t
};
// the shard of rank r (same arithmetic as ctransformers_b200/tp_plan.py; throws when the shape does not tile into 256-blocks)
TPShard tp_shard(int n_embd, int n_head, int n_head_kv, int n_ff, int rank, int world);

struct HParams {
  bool falcon = false;
  int n_vocab = 0, n_ctx_train = 0, n_embd = 0, n_ff = 0, n_head = 0, n_head_kv = 0, n_layer = 0, n_rot = 0;
  float eps = 1e-5f, rope_base = 10000.f, rope_scale = 1.f;
  int n_ctx = 512;
  int head_dim() const { return n_embd / n_head; }
  int n_embd_gqa() const { return head_dim() * n_head_kv; }
};

struct LayerW {
  DevMat wq, wk, wv, wqkv, wo, w1, w2, w3;
  const float* attn_norm = nullptr; const float* attn_norm_b = nullptr;
  const float* attn_norm2 = nullptr; const float* attn_norm2_b = nullptr;
  const float* ffn_norm = nullptr;
};

struct Phase;    // stream.cuh: one phase of the persistent step kernel
struct StepOp;   // engine.cu: one op of the per-token schedule
struct MVParams;
struct Uploader;
struct PrefillState;

struct EvalStats { double last_eval_ms = 0; long launches = 0; size_t weight_bytes_per_token = 0; long spec_hits = 0; double load_ms = 0; size_t load_bytes = 0; };

class Engine {
 public:
  Engine(const GGUFFile& g, const HParams& hp, int device, const TPShard& tp = TPShard());
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  // Evaluate n tokens starting at position n_past; afterwards logits()/embeddings() hold the last token's.
  void eval(const int* tokens, int n, int n_past);
  // Evaluate a token list with explicit per-token position and n_total (= n_past + N of the reference eval call the token
  // belongs to, llm.h:40-54): consecutive positions go through the batched prefill kernel, PB_T tokens per launch.
  void eval_list(const int* tokens, const int* pos, const int* n_total, int n);
  // n_steps greedy decode steps entirely on the device stream (token feedback through k_argmax);
  // out_tokens[n_steps] receives the picked ids.  Returns device-timed milliseconds for the steps.
  double decode_greedy(int first_token, int n_past, int n_steps, int* out_tokens);
  // One eager (un-graphed) decode step at n_past with a CUDA event after every kernel; accumulates the
  // per-class device time.  kinds: 0 mat-vec, 1 attention, 2 rope+kv store, 3 other.  Returns kernel count.
  double time_matvec_only(int reps, long* launches, unsigned mask = 0);
  int profile_step(int token, int n_past, double ms_by_kind[4], int count_by_kind[4]);
  long trace_step(int token, int n_past, unsigned long long* out, long cap_words);

  // Host views of the last token's logits / hidden state (the reference hands out ctx->logits.data(), mutable by the caller,
  // llama.cc:47-51).  Until a caller asks for one, nothing is copied per eval (lazy); from the first request on every eval
  // refreshes them, because the caller may keep the pointer.
  float* logits() { host_views(); return h_logits_; }
  float* embeddings() { host_views(); return h_embd_; }
  bool lazy_logits() const { return !eager_; }
  std::vector<float> logits_copy();   // this eval's logits without switching the engine to eager host views
  // Device half of the sampler (sample_gpu.cuh): candidates >= the k-th largest penalised logit.  Returns their count, or -1
  // when the device path does not apply (window too long, k too large).
  int topk_candidates(const int* last, int n_last, float penalty, int k, int* ids, float* logits);
  // The last eval's greedy pick when the engine has it and it is unambiguous (a unique maximum), else -1.
  int greedy_pick();
  const HParams& hparams() const { return hp_; }
  EvalStats stats;
  void set_stream(cudaStream_t s);   // run on a caller-owned stream (bench: torch's current stream)
  cudaStream_t stream() const { return stream_; }

 private:
  HParams hp_;
  TPShard tp_;
  int nh_ = 0, nkv_ = 0, nff_ = 0;   // this rank's query heads, KV heads and n_ff slice (the whole model when world == 1)
  void tp_all_reduce(float* buf, int n);
  // fused exchange (stream.cuh: XchgParams): this rank's region  uint2 ll[2][world][n_embd]  and every peer's, IPC-mapped
  bool tp_peer_ = false;
  uint8_t* xc_region_ = nullptr;
  uint2* xc_ll_[8] = {nullptr};
  void tp_setup_peer();
  int device_ = 0;
  cudaStream_t stream_ = nullptr;
  bool own_stream_ = true;
  // arena
  uint8_t* arena_ = nullptr;
  size_t arena_size_ = 0, arena_used_ = 0;
  void* alloc(size_t bytes, size_t align = 256);

  // weights
  DevMat output_;
  const uint8_t* tok_embd_ = nullptr; int tok_type_ = 0; size_t tok_row_bytes_ = 0;
  const float* out_norm_ = nullptr; const float* out_norm_b_ = nullptr;
  std::vector<LayerW> layers_;
  uint16_t *silu_tab_ = nullptr, *gelu_tab_ = nullptr, *exp_tab_ = nullptr;
  float2* rope_ = nullptr;
  // KV cache
  uint16_t *kc_ = nullptr, *vc_ = nullptr;
  // workspace
  int* d_state_ = nullptr;     // {token, n_past}
  float *xa_ = nullptr, *xb_ = nullptr, *qkv_ = nullptr, *attn_ = nullptr, *attn_o_ = nullptr, *ffn_ = nullptr, *ffn2_ = nullptr, *d_logits_ = nullptr, *d_embd_ = nullptr, *d_logits_keep_ = nullptr, *d_embd_keep_ = nullptr;
  // host (pinned) results
  float *h_logits_ = nullptr, *h_embd_ = nullptr;
  int* h_state_ = nullptr;     // pinned ring of {token, n_past}
  int h_state_cap_ = 0;
  long h_state_next_ = 0;
  int* h_tokens_out_ = nullptr;
  int* d_tokens_out_ = nullptr;
  int tokens_out_cap_ = 0;

  cudaGraphExec_t graph_full_ = nullptr, graph_nolog_ = nullptr, graph_greedy_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, ev_pick_ = nullptr;
  // speculative next step (see after_eval)
  bool spec_on_ = true, spec_pending_ = false;
  bool spec_deferred_ = false;   // a look-ahead step is due but waits for the device sampler to be enqueued first
  bool sampler_mode_ = false;    // the caller's last sample() ran the device sampler kernel (not the greedy pick)
  cudaEvent_t ev_sample_ = nullptr;
  void launch_deferred_spec();
  int spec_pos_ = -1, spec_streak_ = 0;
  int kv_high_ = 0;              // one past the highest position any eval has written
  int* h_spec_tok_ = nullptr;
  int* h_dbg_ = nullptr;         // host-mapped watchdog words of the persistent kernels
  void after_eval(int next_pos);
  int sm_count_ = 148;
  long launches_per_step_ = 0;

  void init(const GGUFFile& g);
  void release();
  // rows [row0, row1) and K range [k0, k1) of the tensor (defaults: all of it); the shape check is against the FULL tensor
  DevMat upload_matrix(const GGUFTensor& t, struct Uploader& up, int want_K, int want_M, int row0 = 0, int row1 = -1, int k0 = 0, int k1 = -1);
  const float* upload_vector(const GGUFFile& g, const std::string& name, bool required, int want_n);
  // the per-token schedule
  std::vector<StepOp> ops_;      // EMBED, layers..., HEAD, PICK
  int n_body_ = 0;               // ops before the HEAD mat-vec
  Phase* d_prog_ = nullptr;      // device copy of ops_' phases (index = op index)
  Phase* d_prog_mv_ = nullptr;   // scratch program of time_matvec_only
  int *d_bounds_ = nullptr, *d_bounds_mv_ = nullptr;   // per-CTA tile ranges of the two programs
  unsigned* d_sync_ = nullptr;   // grid-barrier words of the step kernel
  int step_grid_ = 0, step_slots_ = 0;   // launch shape of the step kernel: CTAs, ring slots,
  size_t step_smem_ = 0;                 // dynamic shared memory
  bool fused_ = true;            // CTB_STEP_FUSE=0: one kernel per op
  void build_ops();
  void push_matvec(struct MVParams& p, int kind);
  void upload_prog(Phase* dst, int* dst_bounds, const std::vector<StepOp>& ops);
  void enqueue_ops(const std::vector<StepOp>& ops, const Phase* d_prog, const int* d_bounds, int n);
  void build_graphs();
  void destroy_graphs();
  bool eager_ = false;           // host logits / embeddings are refreshed by every eval
  bool host_fresh_ = true;
  void host_views();
  struct SampleGpuOut* d_sample_ = nullptr;
  struct SampleGpuOut* h_sample_ = nullptr;
  int* d_last_ = nullptr;
  // batched prefill (prefill.cuh): built on first use
  struct PrefillState* pf_ = nullptr;
  bool prefill_on_ = true;       // CTB_NO_PREFILL=1: prompts run through the single-token kernel
  int prefill_min_ = 4;          // shortest run of consecutive tokens worth a batched launch
  bool ensure_prefill();
  void prefill_batch(const int* tokens, const int* pos, const int* n_total, int n, bool last);
  void decode_one(int token, int pos, int n_total, bool with_logits);
  void finish_eval(int next_pos, bool hit);
  enum : int { MVK_QKV = 0, MVK_WO = 1, MVK_UP = 2, MVK_DOWN = 3, MVK_OUT = 4 };   // which projection a mat-vec launch is
  bool profiling_ = false;
  bool pdl_ = true;              // programmatic dependent launch between the kernels of a step (CTB_NO_PDL=1 turns it off)
  std::vector<cudaEvent_t> prof_ev_;
  std::vector<int> prof_kind_;
  void mark(int kind);
};

size_t engine_arena_bytes(const GGUFFile& g, const HParams& hp);

}  // namespace ctb
