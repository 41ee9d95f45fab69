// The drop-in boundary: the reference's 17 `ctransformers_llm_*` C entry points
// (reference: models/llm.cc:32-138, bound by ctransformers/llm.py:117-208), re-implemented on top of the
// B200 engine, plus additive `ctb_*` entry points (declared in include/ctransformers_b200.h).
//
// Same semantics as the reference class LLM / llama_llm (models/llm.h:13-138, models/llms/llama.cc:10-117):
//   * create → nullptr (+ message on stderr) on any failure; nothing throws across the ABI
//   * batch_eval chunks by min(batch_size, n_ctx), clamps n_past to n_ctx - chunk (llm.h:40-54, 124-137)
//   * logits_data is a writable host pointer to the last token's n_vocab logits, valid until the next eval
//   * sample reseeds the RNG on every call (llama.cc:57-60)
// There is NO CPU fallback: without a CUDA device create fails loudly.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/ctransformers_b200.h"
#include "engine.cuh"
#include "gguf.hpp"
#include "sampler.hpp"
#include "tp_nccl.hpp"
#include "vocab.hpp"

using namespace ctb;

struct LLM {
  ~LLM() {
    engine.reset();   // before the communicator its graphs captured
    if (comm) NcclApi::get().CommDestroy((ncclComm_t)comm);
  }
  void* comm = nullptr;   // ncclComm_t of the tensor-sharded mode
  std::unique_ptr<GGUFFile> file;
  Vocab vocab;
  HParams hp;
  std::unique_ptr<Engine> engine;
  std::string arch;
  std::string piece_buf;
  std::mt19937 rng;
  bool has_logits = false;
  long gpu_samples = 0;   // sample() calls answered by the device-side penalty + top-k
};

static bool file_is_gguf(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  uint32_t magic = 0;
  const size_t n = fread(&magic, 1, 4, f);
  fclose(f);
  return n == 4 && magic == 0x46554747u;
}

static HParams read_hparams(const GGUFFile& g, const std::string& arch, int n_ctx_req) {
  HParams hp;
  hp.falcon = arch == "falcon";
  const GGUFValue* toks = g.find("tokenizer.ggml.tokens");
  if (!toks) throw std::runtime_error("key not found in model: tokenizer.ggml.tokens");
  hp.n_vocab = (int)toks->arr_n;
  hp.n_ctx_train = (int)g.need_u32(arch + ".context_length");
  hp.n_embd = (int)g.need_u32(arch + ".embedding_length");
  hp.n_ff = (int)g.need_u32(arch + ".feed_forward_length");
  hp.n_head = (int)g.need_u32(arch + ".attention.head_count");
  hp.n_layer = (int)g.need_u32(arch + ".block_count");
  hp.n_head_kv = (int)g.get_u32(arch + ".attention.head_count_kv", (uint32_t)hp.n_head);
  // reference llama.cpp:1576-1595: model values override the (default) context params
  hp.rope_base = g.get_f32(arch + ".rope.freq_base", 10000.0f);
  const float lin = g.get_f32(arch + ".rope.scale_linear", 1.0f);
  hp.rope_scale = lin != 1.0f ? 1.0f / lin : 1.0f;
  if (hp.n_head <= 0 || hp.n_embd % hp.n_head) throw std::runtime_error("invalid head count");
  hp.n_rot = (int)g.get_u32(arch + ".rope.dimension_count", (uint32_t)(hp.n_embd / hp.n_head));
  if (hp.n_rot != hp.n_embd / hp.n_head) throw std::runtime_error("invalid n_rot");
  hp.eps = hp.falcon ? g.need_f32(arch + ".attention.layer_norm_epsilon") : g.need_f32(arch + ".attention.layer_norm_rms_epsilon");
  hp.n_ctx = n_ctx_req > 0 ? n_ctx_req : 512;   // llama_context_default_params().n_ctx (llama.cpp:5281)
  if (hp.n_head_kv <= 0 || hp.n_head % hp.n_head_kv) throw std::runtime_error("invalid kv head count");
  const int hd = hp.head_dim();
  if (hd != 64 && hd != 128) throw std::runtime_error("unsupported head size " + std::to_string(hd) + " (B200 path handles 64 and 128)");
  return hp;
}

extern "C" {

static LLM* create_llm(const char* model_path, const char* model_type, const ctransformers_config config, int rank, int world, const void* unique_id) {
  try {
    std::string type = model_type ? model_type : "";
    type.erase(std::remove_if(type.begin(), type.end(), [](const char c) { return !std::isalnum((unsigned char)c); }), type.end());
    if (!(type == "gguf" || file_is_gguf(model_path))) {
      fprintf(stderr, "Model type '%s' is not supported by the B200 build (GGUF llama / falcon only).\n", model_type ? model_type : "");
      return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      fprintf(stderr, "ctransformers-b200: no CUDA device available; this library has no CPU fallback.\n");
      return nullptr;
    }
    std::unique_ptr<LLM> llm(new LLM);
    llm->file.reset(new GGUFFile(model_path));
    llm->arch = llm->file->need_str("general.architecture");
    if (llm->arch != "llama" && llm->arch != "falcon") throw std::runtime_error("unknown model architecture: '" + llm->arch + "'");
    llm->hp = read_hparams(*llm->file, llm->arch, config.context_length);
    llm->vocab.load(*llm->file);
    int device = 0;
    if (const char* env = getenv("CT_DEVICE")) device = atoi(env);
    else if (const char* lr = getenv("LOCAL_RANK")) device = atoi(lr) % ndev;
    TPShard tp;
    if (world > 1) {
      if (!unique_id) throw std::runtime_error("tensor parallel: no communicator id");
      tp = tp_shard(llm->hp.n_embd, llm->hp.n_head, llm->hp.n_head_kv, llm->hp.n_ff, rank, world);
      if (cudaSetDevice(device) != cudaSuccess) throw std::runtime_error("cudaSetDevice failed");
      const NcclApi& nccl = NcclApi::get();
      ncclUniqueId id;
      memcpy(&id, unique_id, sizeof(id));
      ncclComm_t comm = nullptr;
      nccl.check(nccl.CommInitRank(&comm, world, id, rank), "communicator init");
      llm->comm = comm;
      tp.comm = comm;
    }
    llm->engine.reset(new Engine(*llm->file, llm->hp, device, tp));
    return llm.release();
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: failed to load model: %s\n", e.what());
    return nullptr;
  } catch (...) {
    fprintf(stderr, "ctransformers-b200: failed to load model\n");
    return nullptr;
  }
}

LLM* ctransformers_llm_create(const char* model_path, const char* model_type, const ctransformers_config config) {
  return create_llm(model_path, model_type, config, 0, 1, nullptr);
}

int ctb_tp_unique_id(void* out, int cap) {
  try {
    if (!out || cap < (int)sizeof(ncclUniqueId)) return -(int)sizeof(ncclUniqueId);
    const NcclApi& nccl = NcclApi::get();
    ncclUniqueId id;
    nccl.check(nccl.GetUniqueId(&id), "unique id");
    memcpy(out, &id, sizeof(id));
    return (int)sizeof(id);
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: %s\n", e.what());
    return 0;
  } catch (...) { return 0; }
}

LLM* ctb_llm_create_tp(const char* model_path, const char* model_type, const ctransformers_config config, int rank, int world, const void* unique_id) {
  if (world < 1 || rank < 0 || rank >= world) {
    fprintf(stderr, "ctransformers-b200: bad tensor-parallel rank %d of %d\n", rank, world);
    return nullptr;
  }
  return create_llm(model_path, model_type, config, rank, world, unique_id);
}

int ctb_tp_shard(int n_embd, int n_head, int n_head_kv, int n_ff, int rank, int world, int* out6) {
  try {
    const TPShard s = tp_shard(n_embd, n_head, n_head_kv, n_ff, rank, world);
    out6[0] = s.head0; out6[1] = s.head1; out6[2] = s.kv0; out6[3] = s.kv1; out6[4] = s.ff0; out6[5] = s.ff1;
    return 0;
  } catch (...) { return -1; }
}

void ctransformers_llm_delete(LLM* llm) { delete llm; }

int ctransformers_llm_tokenize(LLM* llm, const char* text, const bool add_bos_token, int* output) {
  try {
    const std::vector<int> t = llm->vocab.tokenize(text ? text : "", add_bos_token);
    std::copy(t.begin(), t.end(), output);
    return (int)t.size();
  } catch (...) { return 0; }
}

const char* ctransformers_llm_detokenize(LLM* llm, const int token) {
  try {
    llm->piece_buf = llm->vocab.piece(token);
  } catch (...) { llm->piece_buf.clear(); }   // nothing may cross the C ABI
  return llm->piece_buf.c_str();
}

bool ctransformers_llm_is_eos_token(LLM* llm, const int token) { return token == llm->vocab.eos; }
int ctransformers_llm_eos_token_id(LLM* llm) { return llm->vocab.eos; }
int ctransformers_llm_bos_token_id(LLM* llm) { return llm->vocab.bos; }
int ctransformers_llm_vocab_size(LLM* llm) { return llm->hp.n_vocab; }
int ctransformers_llm_context_length(LLM* llm) { return llm->hp.n_ctx; }
const char* ctransformers_llm_architecture(LLM* llm) { return llm->arch.c_str(); }

bool ctransformers_llm_batch_eval(LLM* llm, const int* tokens, const int n_tokens, const int n_past, const int batch_size, const int threads) {
  (void)threads;   // host thread count has no meaning on the GPU path
  try {
    const int n_ctx = llm->hp.n_ctx;
    const int bs = std::max(1, std::min(n_ctx, batch_size));
    // LLM::BatchEval (llm.h:40-54): chunks of batch_size tokens, n_past clamped per chunk (llm.h:126).  The chunk a token belongs
    // to fixes the row length n_total = n_past + N of its attention mat-muls; the engine gets the whole list at once so that
    // prompt chunks can share batched launches.
    std::vector<int> pos(n_tokens), nt(n_tokens);
    int past = n_past;
    for (int start = 0; start < n_tokens; start += bs) {
      const int n = std::min(bs, n_tokens - start);
      const int p = std::max(0, std::min(n_ctx - n, past));
      for (int i = 0; i < n; i++) {
        if (tokens[start + i] < 0 || tokens[start + i] >= llm->hp.n_vocab) throw std::runtime_error("token id out of range");
        pos[start + i] = p + i;
        nt[start + i] = p + n;
      }
      past += n;
    }
    llm->engine->eval_list(tokens, pos.data(), nt.data(), n_tokens);
    if (n_tokens > 0) llm->has_logits = true;
    return true;
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: eval failed: %s\n", e.what());
    return false;
  } catch (...) { return false; }
}

float* ctransformers_llm_logits_data(LLM* llm) { return llm->engine->logits(); }
int ctransformers_llm_logits_size(LLM* llm) { return llm->has_logits ? llm->hp.n_vocab : 0; }
const float* ctransformers_llm_embeddings_data(LLM* llm) { return llm->engine->embeddings(); }
int ctransformers_llm_embeddings_size(LLM* llm) { return llm->has_logits ? llm->hp.n_embd : 0; }

int ctransformers_llm_sample(LLM* llm, const int* last_tokens, const int n_last, const int top_k, const float top_p, const float temperature,
                             const float repetition_penalty, int seed) {
  try {
    if (seed < 0) seed = (int)time(nullptr);
    llm->rng.seed((unsigned)seed);
    if (llm->engine->lazy_logits()) {
      // nobody holds a host view of the logits: penalty + top-k run on the device, only the candidates come back
      if (top_k == 1 && (repetition_penalty == 1.0f || n_last <= 0)) {
        // greedy: one candidate survives top-k, so top-p / temperature / the draw cannot change it (llama.cpp:3832-3857,
        // 4215-4240).  The engine already holds the arg-max of these logits (its look-ahead pick); equal maxima fall through.
        const int pick = llm->engine->greedy_pick();
        if (pick >= 0) {
          llm->gpu_samples++;
          return pick;
        }
      }
      int ids[256];
      float lg[256];
      const int count = llm->engine->topk_candidates(last_tokens, n_last, repetition_penalty, top_k, ids, lg);
      std::vector<Candidate> c;
      if (count > 0 && device_candidates_usable(ids, lg, count, top_k, llm->hp.n_vocab, c)) {
        llm->gpu_samples++;
        return sample_candidates(c, top_k, top_p, temperature, llm->rng);
      }
      // ambiguous cut (equal logits): the reference's own sort over all candidates decides — host path on a private copy
      std::vector<float> all = llm->engine->logits_copy();
      return sample_token(all.data(), llm->hp.n_vocab, last_tokens, n_last, top_k, top_p, temperature, repetition_penalty, llm->rng);
    }
    return sample_token(llm->engine->logits(), llm->hp.n_vocab, last_tokens, n_last, top_k, top_p, temperature, repetition_penalty, llm->rng);
  } catch (...) { return llm->vocab.eos; }
}

void ctransformers_llm_reset(LLM* llm) { (void)llm; /* reference clears only the generic logits_ vector, which GGUF models do not use (llm.h:106, llama.cc:47) */ }

// ----------------------------------------------------------------------------- additive entry points
int ctb_abi_version(void) { return 1; }

double ctb_llm_last_eval_ms(LLM* llm) { return llm->engine->stats.last_eval_ms; }
long ctb_llm_launches_per_token(LLM* llm) { return llm->engine->stats.launches; }
long ctb_llm_speculative_hits(LLM* llm) { return llm->engine->stats.spec_hits; }
unsigned long long ctb_llm_weight_bytes_per_token(LLM* llm) { return (unsigned long long)llm->engine->stats.weight_bytes_per_token; }
long ctb_llm_device_samples(LLM* llm) { return llm->gpu_samples; }
double ctb_llm_load_ms(LLM* llm) { return llm->engine->stats.load_ms; }
void ctb_llm_set_stream(LLM* llm, void* cuda_stream) { llm->engine->set_stream((cudaStream_t)cuda_stream); }

double ctb_llm_decode_greedy(LLM* llm, int first_token, int n_past, int n_steps, int* out_tokens) {
  try {
    const double ms = llm->engine->decode_greedy(first_token, n_past, n_steps, out_tokens);
    llm->has_logits = true;
    return ms;
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: decode_greedy failed: %s\n", e.what());
    return -1.0;
  }
}

int ctb_llm_profile_step(LLM* llm, int token, int n_past, double* ms_by_kind, int* count_by_kind) {
  try {
    return llm->engine->profile_step(token, n_past, ms_by_kind, count_by_kind);
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: profile_step failed: %s\n", e.what());
    return -1;
  }
}

long ctb_llm_trace_step(LLM* llm, int token, int n_past, unsigned long long* out, long cap_words) {
  try {
    return llm->engine->trace_step(token, n_past, out, cap_words);
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: trace_step failed: %s\n", e.what());
    return 0;
  }
}

double ctb_llm_time_matvec_only(LLM* llm, int reps, long* launches) { return ctb_llm_time_matvec_kinds(llm, reps, launches, 0); }

double ctb_llm_time_matvec_kinds(LLM* llm, int reps, long* launches, unsigned kind_mask) {
  try {
    return llm->engine->time_matvec_only(reps < 1 ? 1 : reps, launches, kind_mask);
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: time_matvec_only failed: %s\n", e.what());
    return -1.0;
  }
}

// ---- host-only logic (no GPU needed): tokenizer / detokenizer / sampler on their own
struct ctb_vocab { std::unique_ptr<GGUFFile> file; Vocab vocab; };

ctb_vocab* ctb_vocab_load(const char* gguf_path) {
  try {
    std::unique_ptr<ctb_vocab> v(new ctb_vocab);
    v->file.reset(new GGUFFile(gguf_path));
    v->vocab.load(*v->file);
    return v.release();
  } catch (const std::exception& e) {
    fprintf(stderr, "ctransformers-b200: ctb_vocab_load failed: %s\n", e.what());
    return nullptr;
  }
}
void ctb_vocab_free(ctb_vocab* v) { delete v; }
int ctb_vocab_size(ctb_vocab* v) { return v->vocab.size(); }
int ctb_vocab_tokenize(ctb_vocab* v, const char* text, bool add_bos, int* out, int cap) {
  try {
    const std::vector<int> t = v->vocab.tokenize(text ? text : "", add_bos);
    if ((int)t.size() > cap) return -(int)t.size();
    std::copy(t.begin(), t.end(), out);
    return (int)t.size();
  } catch (...) { return 0; }
}
int ctb_vocab_piece(ctb_vocab* v, int token, char* buf, int cap) {
  const std::string s = v->vocab.piece(token);
  if ((int)s.size() > cap) return -(int)s.size();
  memcpy(buf, s.data(), s.size());
  return (int)s.size();
}
int ctb_sample(const float* logits, int n_vocab, const int* last_tokens, int n_last, int top_k, float top_p, float temperature,
               float repetition_penalty, int seed) {
  try {
    if (seed < 0) seed = (int)time(nullptr);
    std::mt19937 rng((unsigned)seed);
    return sample_token(logits, n_vocab, last_tokens, n_last, top_k, top_p, temperature, repetition_penalty, rng);
  } catch (...) { return -1; }
}

}  // extern "C"
