// Vocabulary, tokenizer and detokenizer behind ctransformers_llm_{tokenize,detokenize,...}.
//
// Host-only, deterministic.  Behaviour follows the reference so token ids are identical:
//   vocab load          models/ggml/llama.cpp:1648-1760
//   tokenize entry      models/ggml/llama.cpp:3389-3423  (BOS, SPM leading space + U+2581 escape)
//   SPM merge order     models/ggml/llama.cpp:3066-3196  (highest score first, ties → leftmost)
//   BPE (gpt2 vocab)    models/ggml/llama.cpp:3228-3387  (GPT-2 pre-split regex, lowest merge rank first)
//   token → text        models/ggml/llama.cpp:6151-6187
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <queue>
#include <regex>
#include <string>
#include <unordered_map>
#include <vector>

#include "gguf.hpp"

namespace ctb {

enum TokenType : int { TT_UNDEFINED = 0, TT_NORMAL = 1, TT_UNKNOWN = 2, TT_CONTROL = 3, TT_USER = 4, TT_UNUSED = 5, TT_BYTE = 6 };

struct Vocab {
  enum Kind { SPM, BPE } kind = SPM;
  struct Entry { std::string text; float score; int type; };
  std::vector<Entry> entries;
  std::unordered_map<std::string, int> lookup;
  std::map<std::pair<std::string, std::string>, int> merge_rank;   // BPE only
  int bos = 1, eos = 2, unk = 0, sep = -1, pad = -1;

  int size() const { return (int)entries.size(); }

  void load(const GGUFFile& g) {
    const GGUFValue* toks = g.find("tokenizer.ggml.tokens");
    if (!toks) throw std::runtime_error("cannot find tokenizer vocab in model file");
    const GGUFValue* scores = g.find("tokenizer.ggml.scores");
    if (!scores) throw std::runtime_error("cannot find tokenizer scores in model file");
    const GGUFValue* types = g.find("tokenizer.ggml.token_type");
    if (!types) throw std::runtime_error("cannot find token type list in GGUF file");
    if (scores->arr_n < toks->arr_n || types->arr_n < toks->arr_n) throw std::runtime_error("tokenizer arrays have inconsistent lengths");
    // gguf_type: 6 = float32, 5 = int32, 8 = string (the reference reads them with exactly these element types, llama.cpp:1657-1674)
    if (toks->arr_type != 8 || scores->arr_type != 6 || types->arr_type != 5 || !scores->arr_data || !types->arr_data)
      throw std::runtime_error("tokenizer arrays have unexpected element types");

    const std::string model = g.need_str("tokenizer.ggml.model");
    if (model == "gpt2") {
      kind = BPE;
      const GGUFValue* merges = g.find("tokenizer.ggml.merges");
      if (!merges) throw std::runtime_error("cannot find tokenizer merges in model file");
      for (size_t i = 0; i < merges->arr_str.size(); i++) {
        const std::string& m = merges->arr_str[i];
        const size_t sp = m.find(' ', 1);
        std::string a, b;
        if (sp != std::string::npos) { a = m.substr(0, sp); b = m.substr(sp + 1); }
        merge_rank.emplace(std::make_pair(a, b), (int)i);
      }
      bos = 11; eos = 11; unk = -1;
    } else {
      if (model != "llama") fprintf(stderr, "ctransformers-b200: unknown tokenizer '%s', using 'llama'\n", model.c_str());
      kind = SPM;
    }

    const float* sc = (const float*)scores->arr_data;
    const int32_t* ty = (const int32_t*)types->arr_data;
    entries.resize(toks->arr_str.size());
    for (size_t i = 0; i < entries.size(); i++) {
      entries[i] = Entry{toks->arr_str[i], sc[i], ty[i]};
      lookup[entries[i].text] = (int)i;
    }
    if (kind == SPM) (void)byte_token('\n');  // the reference resolves the newline byte token at load and fails without it
    bos = (int)g.get_u32("tokenizer.ggml.bos_token_id", (uint32_t)bos);
    eos = (int)g.get_u32("tokenizer.ggml.eos_token_id", (uint32_t)eos);
    unk = (int)g.get_u32("tokenizer.ggml.unknown_token_id", (uint32_t)unk);
    sep = (int)g.get_u32("tokenizer.ggml.seperator_token_id", (uint32_t)sep);
    pad = (int)g.get_u32("tokenizer.ggml.padding_token_id", (uint32_t)pad);
  }

  int byte_token(uint8_t ch) const {
    char buf[8];
    snprintf(buf, sizeof(buf), "<0x%02X>", ch);
    auto it = lookup.find(buf);
    if (it == lookup.end()) throw std::runtime_error(std::string("vocab has no byte token ") + buf);
    return it->second;
  }

  std::vector<int> tokenize(const std::string& raw, bool add_bos) const {
    std::vector<int> out;
    if (add_bos && bos != -1) out.push_back(bos);
    if (raw.empty()) return out;
    if (kind == SPM) {
      std::string text;
      text.reserve(raw.size() * 3 + 3);
      for (char c : std::string(" ") + raw) {
        if (c == ' ') text += "\xe2\x96\x81"; else text += c;
      }
      spm(text, out);
    } else {
      bpe(raw, out);
    }
    return out;
  }

  // Text of one token as the detokenizer emits it.
  std::string piece(int id) const {
    if (id < 0 || id >= size()) return "";
    const Entry& e = entries[id];
    switch (e.type) {
      case TT_NORMAL: {
        if (kind != SPM) return e.text;
        std::string r;
        for (size_t i = 0; i < e.text.size();) {
          if (e.text.compare(i, 3, "\xe2\x96\x81") == 0) { r += ' '; i += 3; } else { r += e.text[i++]; }
        }
        return r;
      }
      case TT_UNKNOWN: return "\xe2\x96\x85";
      case TT_BYTE: return e.text.size() >= 5 ? std::string(1, (char)strtol(e.text.substr(3, 2).c_str(), nullptr, 16)) : std::string();   // "<0xNN>"
      default: return "";
    }
  }

 private:
  struct Sym { int prev, next; const char* p; size_t n; };

  static size_t utf8_len(char c) {
    static const uint8_t len[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
    return len[(uint8_t)c >> 4];
  }

  // ---- SentencePiece-style greedy bigram merging
  struct SpmCand { int l, r; float score; size_t bytes; };
  struct SpmLess {  // max-heap: higher score first, then smaller left index
    bool operator()(const SpmCand& a, const SpmCand& b) const { return a.score < b.score || (a.score == b.score && a.l > b.l); }
  };

  void spm(const std::string& text, std::vector<int>& out) const {
    std::vector<Sym> syms;
    for (size_t off = 0; off < text.size();) {
      size_t n = std::min(utf8_len(text[off]), text.size() - off);
      Sym s{(int)syms.size() - 1, 0, text.data() + off, n};
      off += n;
      s.next = off == text.size() ? -1 : (int)syms.size() + 1;
      syms.push_back(s);
    }
    std::priority_queue<SpmCand, std::vector<SpmCand>, SpmLess> heap;
    std::map<std::string, std::pair<int, int>> split_of;   // merged text → the two symbols it came from
    auto offer = [&](int l, int r) {
      if (l < 0 || r < 0) return;
      std::string joined(syms[l].p, syms[l].n + syms[r].n);
      auto it = lookup.find(joined);
      if (it == lookup.end() || (size_t)it->second >= entries.size()) return;
      heap.push(SpmCand{l, r, entries[it->second].score, joined.size()});
      split_of[joined] = {l, r};
    };
    for (int i = 1; i < (int)syms.size(); i++) offer(i - 1, i);
    while (!heap.empty()) {
      SpmCand c = heap.top();
      heap.pop();
      Sym& L = syms[c.l];
      Sym& R = syms[c.r];
      if (L.n == 0 || R.n == 0 || L.n + R.n != c.bytes) continue;   // stale candidate
      L.n += R.n;
      R.n = 0;
      L.next = R.next;
      if (R.next >= 0) syms[R.next].prev = c.l;
      offer(L.prev, c.l);
      offer(c.l, L.next);
    }
    std::function<void(const Sym&)> emit = [&](const Sym& s) {
      std::string t(s.p, s.n);
      auto it = lookup.find(t);
      if (it != lookup.end()) { out.push_back(it->second); return; }
      auto sp = split_of.find(t);
      if (sp == split_of.end()) {
        for (size_t j = 0; j < s.n; j++) out.push_back(byte_token((uint8_t)s.p[j]));
        return;
      }
      emit(syms[sp->second.first]);
      emit(syms[sp->second.second]);
    };
    for (int i = 0; i != -1 && !syms.empty(); i = syms[i].next) emit(syms[i]);
  }

  // ---- GPT-2 style byte-pair merges (Falcon GGUF vocabularies)
  struct BpeCand { int l, r; std::string text; int rank; };
  struct BpeLess {  // max-heap on: lower rank first, then smaller left index
    bool operator()(const BpeCand& a, const BpeCand& b) const { return a.rank > b.rank || (a.rank == b.rank && a.l > b.l); }
  };

  // merges are stored in GPT-2's byte-level alphabet: ' ' is U+0120, '\n' is U+010A (reference: llama.cpp:962-974)
  static std::string to_merge_alphabet(const std::string& s) {
    std::string r;
    for (char c : s) {
      if (c == ' ') r += "\xc4\xa0"; else if (c == '\n') r += "\xc4\x8a"; else r += c;
    }
    return r;
  }
  int rank_of(const std::string& a0, const std::string& b0) const {
    const std::string a = to_merge_alphabet(a0), b = to_merge_alphabet(b0);
    auto it = merge_rank.find(std::make_pair(a, b));
    return it == merge_rank.end() ? -1 : it->second;
  }

  void bpe(const std::string& raw, std::vector<int>& out) const {
    static const std::regex splitter(R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)");
    std::vector<std::string> words;
    {
      std::string rest = raw;
      std::smatch m;
      while (std::regex_search(rest, m, splitter)) {
        for (auto sub : m) words.push_back(sub);
        rest = m.suffix();
      }
    }
    for (const std::string& word : words) {
      std::vector<Sym> syms;
      for (size_t off = 0; off < word.size();) {
        size_t n = std::min(word.size() - off, utf8_len(word[off]));
        Sym s{(int)syms.size() - 1, 0, word.data() + off, n};
        off += n;
        s.next = off == word.size() ? -1 : (int)syms.size() + 1;
        syms.push_back(s);
      }
      std::priority_queue<BpeCand, std::vector<BpeCand>, BpeLess> heap;
      auto offer = [&](int l, int r) {
        if (l < 0 || r < 0) return;
        std::string a(syms[l].p, syms[l].n), b(syms[r].p, syms[r].n);
        int rk = rank_of(a, b);
        if (rk < 0) return;
        heap.push(BpeCand{l, r, a + b, rk});
      };
      for (int i = 1; i < (int)syms.size(); i++) offer(i - 1, i);
      while (!heap.empty()) {
        BpeCand c = heap.top();
        heap.pop();
        Sym& L = syms[c.l];
        Sym& R = syms[c.r];
        if (L.n == 0 || R.n == 0) continue;
        if (std::string(L.p, L.n) + std::string(R.p, R.n) != c.text) continue;   // stale candidate
        L.n += R.n;
        R.n = 0;
        L.next = R.next;
        if (R.next >= 0) syms[R.next].prev = c.l;
        offer(L.prev, c.l);
        offer(c.l, L.next);
      }
      for (const Sym& s : syms) {
        if (s.n == 0) continue;
        std::string t(s.p, s.n);
        auto it = lookup.find(t);
        if (it != lookup.end()) { out.push_back(it->second); continue; }
        for (char ch : t) {
          auto b = lookup.find(std::string(1, ch));
          if (b == lookup.end()) { fprintf(stderr, "ctransformers-b200: byte not found in vocab: 0x%02x\n", (uint8_t)ch); continue; }
          out.push_back(b->second);
        }
      }
    }
  }
};

}  // namespace ctb
