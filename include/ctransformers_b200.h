/*
 * ctransformers_b200.h — C ABI of libctransformers.so (B200 / sm_100a build).
 *
 * Part 1 is the reference's own FFI for the hot path, unchanged: the 17 `ctransformers_llm_*`
 * functions that ctransformers/llm.py:117-208 binds with ctypes and that models/llm.cc:32-138 defines.
 * An unmodified `ctransformers` Python package drives this library through
 * `AutoModelForCausalLM.from_pretrained(path, lib="<this .so>")` (lib.py:12-15 returns unknown strings verbatim).
 *
 * Part 2 is additive (`ctb_*`): timing/introspection hooks and op-level entry points with plain
 * host pointers that mirror the ggml operators on the path, so a maintainer (or a parity test) can call
 * one operator at a time.  No torch / CUDA types appear in any signature.
 */
#ifndef CTRANSFORMERS_B200_H_
#define CTRANSFORMERS_B200_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ part 1: reference FFI ---- */
typedef struct LLM LLM; /* opaque; reference: class LLM, models/llm.h:13 */

/* reference: struct Config, models/llm.h:6-11 — passed BY VALUE (ctypes ConfigStruct, llm.py:73-79) */
typedef struct ctransformers_config {
  int context_length; /* <= 0: 512, the llama_context default (llama.cpp:5281) */
  int gpu_layers;     /* ignored: every layer always runs on the GPU */
  bool mmap;          /* ignored: the file is always mmap'ed for the upload */
  bool mlock;         /* ignored */
} ctransformers_config;

/* models/llm.cc:36-76.  GGUF llama / falcon only; NULL (+ stderr) on failure or when no CUDA device exists. */
LLM* ctransformers_llm_create(const char* model_path, const char* model_type, const ctransformers_config config);
void ctransformers_llm_delete(LLM* llm);                                                   /* llm.cc:78  */
/* caller provides room for strlen(text)+1 ints (llm.py:335-337); returns the count */
int ctransformers_llm_tokenize(LLM* llm, const char* text, const bool add_bos_token, int* output); /* llm.cc:80-85 */
const char* ctransformers_llm_detokenize(LLM* llm, const int token);     /* llm.cc:87-89; valid until the next call */
bool ctransformers_llm_is_eos_token(LLM* llm, const int token);          /* llm.cc:91-93  */
int ctransformers_llm_eos_token_id(LLM* llm);                            /* llm.cc:95     */
int ctransformers_llm_bos_token_id(LLM* llm);                            /* llm.cc:97     */
int ctransformers_llm_vocab_size(LLM* llm);                              /* llm.cc:99     */
int ctransformers_llm_context_length(LLM* llm);                          /* llm.cc:101    */
const char* ctransformers_llm_architecture(LLM* llm);                    /* llm.cc:103-105: "llama" | "falcon" */
/* llm.cc:107-112 → LLM::BatchEval (llm.h:40-54).  `threads` is accepted and ignored. */
bool ctransformers_llm_batch_eval(LLM* llm, const int* tokens, const int n_tokens, const int n_past, const int batch_size,
                                  const int threads);
float* ctransformers_llm_logits_data(LLM* llm);             /* llm.cc:114: host, writable, n_vocab floats (last token) */
int ctransformers_llm_logits_size(LLM* llm);                /* llm.cc:116 */
const float* ctransformers_llm_embeddings_data(LLM* llm);   /* llm.cc:118-120: last token's post-final-norm state */
int ctransformers_llm_embeddings_size(LLM* llm);            /* llm.cc:122-124 */
/* llm.cc:126-132 → llama_llm::Sample (llama.cc:53-84) */
int ctransformers_llm_sample(LLM* llm, const int* last_tokens, const int n_last, const int top_k, const float top_p,
                             const float temperature, const float repetition_penalty, int seed);
void ctransformers_llm_reset(LLM* llm);                     /* llm.cc:134 */

/* ------------------------------------------------------------------ part 2: additive ---------- */
int ctb_abi_version(void);
double ctb_llm_last_eval_ms(LLM* llm);              /* CUDA-event time of the last batch_eval / decode_greedy */
long ctb_llm_launches_per_token(LLM* llm);          /* kernels in one decode step's CUDA graph (a K-quant model: 1, the persistent step kernel) */
/* batch_eval calls answered by the step the engine had already started for the greedy next token (engine.cu: after_eval);
 * CTB_NO_SPEC=1 in the environment turns that look-ahead off. */
long ctb_llm_speculative_hits(LLM* llm);
unsigned long long ctb_llm_weight_bytes_per_token(LLM* llm); /* algorithmic weight bytes one decode step reads */
/* sample() calls answered by the device-side repetition penalty + top-k (csrc/sample_gpu.cuh): used while no caller holds a
 * host view of the logits (logits_data / embeddings_data never called); otherwise, or when equal logits make the top-k cut
 * ambiguous, the host sampler runs on the full logits exactly as the reference does. */
long ctb_llm_device_samples(LLM* llm);
/* wall-clock milliseconds the weight upload took (mmap -> pinned staging -> H2D -> repack, pipelined; engine.cu Uploader) */
double ctb_llm_load_ms(LLM* llm);
/* Tensor-sharded mode (BASELINE configs[4]; the reference's closest facility is layer offload to ONE GPU, llm.h:20 gpu_layers —
 * it has no multi-GPU path).  One process per GPU: rank 0 calls ctb_tp_unique_id (128 bytes, a ncclUniqueId) and hands the
 * bytes to the other ranks by any host channel; every rank then calls ctb_llm_create_tp.  Each rank keeps its query heads
 * (with their KV heads) and its n_ff slice, cut on 256-element block boundaries (ctb_tp_shard reports the ranges:
 * head0, head1, kv0, kv1, ff0, ff1), and the step sums two n_embd-float vectors per layer across the ranks — inside the step
 * kernel over NVLink peer memory (CUDA IPC), or with NCCL all-reduce between launches (CTB_TP_NCCL=1).  llama graph only; every rank must make the same calls in the same order and ends up with the same logits. */
int ctb_tp_unique_id(void* out, int cap);                     /* bytes written (128), or -needed */
LLM* ctb_llm_create_tp(const char* model_path, const char* model_type, const ctransformers_config config, int rank, int world,
                       const void* unique_id);
int ctb_tp_shard(int n_embd, int n_head, int n_head_kv, int n_ff, int rank, int world, int* out6);
void ctb_llm_set_stream(LLM* llm, void* cuda_stream);        /* run on a caller-owned cudaStream_t */
/* n_steps greedy decode steps with the token fed back on the device (no host round trip per token);
 * returns the device-timed milliseconds, < 0 on error.  Logits of the last step land in logits_data. */
double ctb_llm_decode_greedy(LLM* llm, int first_token, int n_past, int n_steps, int* out_tokens);

/* One eager decode step, one kernel per op (un-fused), with a CUDA event after every kernel; ADDS device milliseconds and launch counts per class into
 * ms_by_kind[4] / count_by_kind[4] (0 mat-vec, 1 attention, 2 rope+kv store, 3 other).  Returns kernels timed, < 0 on error. */
int ctb_llm_profile_step(LLM* llm, int token, int n_past, double* ms_by_kind, int* count_by_kind);
/* The mat-vec phases of one decode step alone (same kernel, parameters and order; attention, embedding and pick left
 * out), replayed reps times as a CUDA graph between two CUDA events: returns milliseconds per step, < 0 on error;
 * *launches = mat-vec phases per step.  KV cache and logits are not meaningful afterwards. */
/* One fused decode step whose step-kernel CTAs stamp %globaltimer (ns) per phase: out = n_phases x {phase kind, projection},
 * then n_phases x n_cta x {barrier passed, input staged, first weight item in shared memory, phase done}.  Returns the number
 * of phases, 0 if the model does not run fused, or -(words needed). */
long ctb_llm_trace_step(LLM* llm, int token, int n_past, unsigned long long* out, long cap_words);
double ctb_llm_time_matvec_only(LLM* llm, int reps, long* launches);
/* Same, restricted to the launches whose kind bit is set in kind_mask (bit 0 QKV, 1 attention output, 2 FFN gate+up,
 * 3 FFN down, 4 output head; 0 = all): per-projection timing under in-graph launch conditions. */
double ctb_llm_time_matvec_kinds(LLM* llm, int reps, long* launches, unsigned kind_mask);

/* Host-only pieces of the boundary, callable without a GPU: the GGUF vocabulary with its SPM / BPE tokenizer
 * (llama.cpp:1648-1760, 3080-3427, 6151-6187) and the sampler chain of llama_llm::Sample (llama.cc:53-84). */
typedef struct ctb_vocab ctb_vocab;
ctb_vocab* ctb_vocab_load(const char* gguf_path);
void ctb_vocab_free(ctb_vocab* v);
int ctb_vocab_size(ctb_vocab* v);
int ctb_vocab_tokenize(ctb_vocab* v, const char* text, bool add_bos, int* out, int cap); /* count, or -needed if cap is too small */
int ctb_vocab_piece(ctb_vocab* v, int token, char* buf, int cap);                        /* bytes written (no NUL), or -needed */
int ctb_sample(const float* logits, int n_vocab, const int* last_tokens, int n_last, int top_k, float top_p, float temperature,
               float repetition_penalty, int seed);

/* Op-level mirrors (host pointers in, host pointers out; return 0 on success).  ggml type ids: 0 F32, 1 F16,
 * 2 Q4_0, 8 Q8_0, 12 Q4_K, 13 Q5_K, 14 Q6_K (models/ggml/ggml.h enum ggml_type). */
/* ggml_mul_mat for quantized src0 (ggml.c:11031-11245): dst[n*M+m] = dot(W row m, quantize(x col n)). */
int ctb_mul_mat(int type, const void* w_blocks, const float* x, float* dst, int K, int M, int N);
/* quantize_row_q8_K (k_quants.c:1191-1241) / quantize_row_q8_0 (ggml.c:1232-1268): reference block bytes out. */
int ctb_quantize_row_q8_K(const float* x, void* y, int k);
int ctb_quantize_row_q8_0(const float* x, void* y, int k);
/* ggml_rms_norm + ggml_mul (mode 1) or ggml_norm + ggml_mul + ggml_add (mode 2) (ggml.c:10674-10720, 10605-10654). */
int ctb_norm(int mode, const float* x, const float* w, const float* b, float* y, int n, float eps);
/* ggml_rope_custom on [n_heads][head_dim] at position pos; mode 0 or 2 (neox) (ggml.c:12430-12566). */
int ctb_rope(float* x, int n_heads, int head_dim, int pos, int mode, float freq_base, float freq_scale);
/* One query token (at position T-1) against T cached positions, all heads, exactly as the reference's attention block:
 * KQ (fp16 operands) -> scale -> softmax (fp16 exp table) -> V·P.  Caches in the reference's own layouts:
 * kcache [T][n_kv*hd] fp16 (rotated K), vcache TRANSPOSED [n_kv*hd][T] fp16 (llama.cpp:2323-2335), q [n_head*hd] already
 * rotated, out [n_head*hd].  n_total = n_past + N of the eval call the token belongs to (>= T): the reference's V·P dot runs
 * over rows of that length and splits them at n_total & ~31 between its SIMD lanes and a scalar tail. */
int ctb_attention(const float* q, const uint16_t* kcache, const uint16_t* vcache, float* out, int n_head, int n_kv, int head_dim,
                  int T, int n_total, float kq_scale);
/* silu(W1 x) * (W3 x) with the fp16 SiLU table (ggml.c:3625-3632) — the fused FFN gate. */
int ctb_ffn_gate(int type, const void* w1_blocks, const void* w3_blocks, const float* x, float* out, int K, int M);
/* ggml_get_rows on a quantized table (ggml.c:11615-11642). */
/* How a K-quant mat-vec phase over nseg matrices (types[], rows[], all K wide) is cut up on a GPU with n_sm SMs — pure host
 * arithmetic, no device needed: first_tile[0..grid] = first 16-row tile of each CTA (byte-balanced), meta = {grid, ring slot
 * bytes, tiles alive per CTA (mailboxes), tiles, consumer warps per CTA, rows per tile, work items of the largest CTA,
 * blocks per work item as Q4_K | Q5_K << 8 | Q6_K << 16}.  0 on success. */
int ctb_matvec_partition(const int* types, const int* rows, int nseg, int K, int n_sm, int* first_tile, int* meta);

int ctb_get_row(int type, const void* table_blocks, int K, int n_rows, int row, float* out);

#ifdef __cplusplus
}
#endif
#endif /* CTRANSFORMERS_B200_H_ */
