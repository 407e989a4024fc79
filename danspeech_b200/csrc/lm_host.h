// Host-side n-gram language model: what dsb_beam_create reads from disk before it builds the device tables.
//
// Two on-disk forms give the SAME in-memory model:
//   * ARPA text                    (lm_load.cu: load_arpa)
//   * KenLM binary, probing model  (lm_load.cu: load_klm) -- what the reference's language_models/*.py factories
//     return (e.g. danspeech/language_models/dsl_3gram.py:16-20) and DanSpeechRecognizer.py:89-92 hands to ctcdecode.
// A KenLM probing binary does not store the words of an n-gram, only a 64-bit hash chained over their vocabulary ids;
// the device table is therefore keyed the same way for both forms: key = (chain hash, order) with KenLM's own
// CombineWordHash and KenLM's id assignment (<unk> = 0, then the words in unigram-section order).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace dsb {

// lm/search_hashed.hh (detail::CombineWordHash): the hash of w1..wn is chained from the LAST word backwards,
// h = wn; h = combine(h, w_{n-1}); ... ; h = combine(h, w1).  The unigram "hash" is the word id itself.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint64_t lm_combine_word_hash(uint64_t current, uint32_t next) {
  return (current * 8978948897894561157ULL) ^ ((uint64_t)(1 + next) * 17894857484156487943ULL);
}

struct HostLm {
  int order = 0;
  std::unordered_map<std::string, int> vocab;   // word -> id (<unk> = 0)
  std::vector<std::string> words;               // id -> word
  struct Gram { uint64_t key; int n; float prob, backoff; };   // key: chain hash (n >= 2) or word id (n = 1)
  std::vector<Gram> grams;
  float unk_prob = -100.f;
  int index(const std::string& w) const {
    auto it = vocab.find(w);
    return it == vocab.end() ? 0 : it->second;
  }
};

constexpr int kLmMaxOrder = 5;

uint64_t lm_chain_hash(const int* ids, int n);                     // ids in sentence order
uint64_t murmur_hash64a(const void* key, size_t len, uint64_t seed);
// 1 = KenLM binary, 0 = not, <0 = error (dsb_last_error set)
int is_kenlm_binary(const char* path);
int load_arpa(const char* path, HostLm& lm);
int load_klm(const char* path, HostLm& lm);

}  // namespace dsb
