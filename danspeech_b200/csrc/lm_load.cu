// Language-model files -> HostLm (lm_host.h): ARPA text and KenLM probing binaries.
//
// The reference never parses an LM itself: it passes the path of a KenLM binary (language_models/dsl_3gram.py:16-20,
// DanSpeechRecognizer.py:89-92) to ctcdecode, whose Scorer loads it through KenLM (decoder.py:99-100).  KenLM is an
// absent third-party dependency; the binary layout below restates KenLM's lm/binary_format.cc, lm/vocab.cc,
// lm/search_hashed.{hh,cc}, lm/weights.hh and util/probing_hash_table.hh for "format version 5" files:
//
//   [0, 88)        Sanity: magic "mmap lm http://kheafield.com/code format version 5\n\0" padded to 56 bytes, then
//                  float 0.0, 1.0, -0.5, uint32 1, uint32 0xFFFFFFFF, (4 bytes padding), uint64 1
//   [88, 108)      FixedWidthParameters: uint8 order @88, float probing_multiplier @92, int32 model_type @96
//                  (0 probing, 1 rest-probing, 2 trie, 3 quantised trie, 4 array trie, 5 quantised array trie),
//                  uint8 has_vocabulary @100, uint32 search_version @104
//   [108, ...)     uint64 counts[order]; the header ends at ALIGN8(108 + 8 * order)
//   vocabulary     ProbingVocabularyHeader {uint32 version = 0, uint32 bound}, then an open-addressing table of
//                  {uint64 key = MurmurHash64A(word, seed 0), uint32 id} (12 bytes, packed) with
//                  buckets = max(counts[0] + 1, (uint64)(multiplier * counts[0])), slot = key % buckets, linear probing,
//                  key 0 = empty.  <unk> (id 0) is not in the table.
//   search         unigrams: {float prob, float backoff} x (counts[0] + 1), indexed by id;
//                  orders 2..N-1: tables of {uint64 key, float prob, float backoff} (16 bytes),
//                  order N: table of {uint64 key, float prob} (12 bytes, packed); buckets and probing as above;
//                  key = CombineWordHash chained over the ids from the last word backwards.
//                  The sign bit of prob and a backoff of -0.0 are KenLM state-minimisation flags ("extends left /
//                  right"), not part of the value: prob is read as -|prob|, backoff -0.0 as 0.
//   strings        (has_vocabulary) the words in id order, each NUL-terminated, "<unk>" first, up to the end of the file
//
// Only the default probing model with the vocabulary included is read; trie / quantised / rest-probing files are
// refused by name.  A file whose size, bucket occupancy or vocabulary hashes disagree with this layout is refused with
// the first inconsistency found rather than decoded wrongly.
#include "common.cuh"
#include "lm_host.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace dsb {

uint64_t lm_chain_hash(const int* ids, int n) {
  uint64_t h = (uint64_t)(uint32_t)ids[n - 1];
  for (int i = n - 2; i >= 0; --i) h = lm_combine_word_hash(h, (uint32_t)ids[i]);
  return h;
}

// util/murmur_hash.cc: MurmurHash64A (Austin Appleby), the 64-bit-platform branch of MurmurHashNative
uint64_t murmur_hash64a(const void* key, size_t len, uint64_t seed) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  const int r = 47;
  uint64_t h = seed ^ (len * m);
  const unsigned char* p = reinterpret_cast<const unsigned char*>(key);
  const unsigned char* end = p + (len / 8) * 8;
  for (; p != end; p += 8) {
    uint64_t k;
    memcpy(&k, p, 8);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
  }
  switch (len & 7) {
    case 7: h ^= (uint64_t)p[6] << 48;  // fall through
    case 6: h ^= (uint64_t)p[5] << 40;  // fall through
    case 5: h ^= (uint64_t)p[4] << 32;  // fall through
    case 4: h ^= (uint64_t)p[3] << 24;  // fall through
    case 3: h ^= (uint64_t)p[2] << 16;  // fall through
    case 2: h ^= (uint64_t)p[1] << 8;   // fall through
    case 1: h ^= (uint64_t)p[0]; h *= m;
  }
  h ^= h >> r;
  h *= m;
  h ^= h >> r;
  return h;
}

static const char kKlmMagicPrefix[] = "mmap lm http://kheafield.com/code format version";

int is_kenlm_binary(const char* path) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return set_error(DSB_ERR_IO, "dsb_beam_create: cannot open language model '%s'", path);
  char hdr[sizeof(kKlmMagicPrefix)];
  const size_t n = fread(hdr, 1, sizeof(kKlmMagicPrefix) - 1, fp);
  fclose(fp);
  return n == sizeof(kKlmMagicPrefix) - 1 && memcmp(hdr, kKlmMagicPrefix, n) == 0 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------ ARPA
static int finish_model(const char* path, HostLm& lm) {
  if (lm.order == 0) return set_error(DSB_ERR_IO, "dsb_beam_create: '%s' is not an ARPA language model", path);
  if (lm.order > kLmMaxOrder)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: LM order %d > %d", lm.order, kLmMaxOrder);
  if (lm.words.size() >= (1u << 31)) return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: vocabulary too large");
  return 0;
}

static bool parse_float(const std::string& s, float* out) {
  if (s.empty()) return false;
  char* end = nullptr;
  const float v = strtof(s.c_str(), &end);   // correctly rounded to float, like KenLM's StringToFloat
  if (end == s.c_str() || *end != '\0') return false;
  *out = v;
  return true;
}

int load_arpa(const char* path, HostLm& lm) {
  std::ifstream f(path);
  if (!f) return set_error(DSB_ERR_IO, "dsb_beam_create: cannot open language model '%s'", path);
  lm.vocab["<unk>"] = 0;
  lm.vocab["<UNK>"] = 0;            // KenLM maps both spellings to id 0 (lm/vocab.cc: kUnknownHash, kUnknownCapHash)
  lm.words.push_back("<unk>");
  std::string line;
  int cur = 0, lineno = 0;
  bool in_data = false, ended = false;
  std::vector<long long> declared(kLmMaxOrder + 2, -1), seen(kLmMaxOrder + 2, 0);
  while (std::getline(f, line)) {
    ++lineno;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '\\') {
      if (line == "\\data\\") {
        in_data = true;
      } else if (line == "\\end\\") {
        ended = true;
        break;
      } else if (line.size() > 8 && line.compare(line.size() - 7, 7, "-grams:") == 0) {
        cur = atoi(line.c_str() + 1);
        if (cur < 1) return set_error(DSB_ERR_IO, "dsb_beam_create: %s:%d: bad section header '%s'", path, lineno, line.c_str());
        if (cur > kLmMaxOrder) return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: LM order %d > %d", cur, kLmMaxOrder);
        if (cur > lm.order) lm.order = cur;
        in_data = false;
      } else {
        return set_error(DSB_ERR_IO, "dsb_beam_create: %s:%d: unknown ARPA section '%s'", path, lineno, line.c_str());
      }
      continue;
    }
    if (in_data) {   // "ngram N=count"
      int n = 0;
      long long c = 0;
      if (sscanf(line.c_str(), "ngram %d=%lld", &n, &c) == 2 && n >= 1 && n <= kLmMaxOrder) declared[n] = c;
      continue;
    }
    if (cur == 0) continue;   // free text before \data\ is allowed by the format
    std::vector<std::string> tok;
    std::stringstream ss(line);
    std::string t;
    while (ss >> t) tok.push_back(t);
    float prob = 0.f, backoff = 0.f;
    if ((int)tok.size() != cur + 1 && (int)tok.size() != cur + 2)
      return set_error(DSB_ERR_IO, "dsb_beam_create: %s:%d: a %d-gram line needs %d or %d fields, found %d", path, lineno, cur,
                       cur + 1, cur + 2, (int)tok.size());
    if (!parse_float(tok[0], &prob) || prob > 0.f)
      return set_error(DSB_ERR_IO, "dsb_beam_create: %s:%d: bad log10 probability '%s'", path, lineno, tok[0].c_str());
    if ((int)tok.size() == cur + 2 && !parse_float(tok[cur + 1], &backoff))
      return set_error(DSB_ERR_IO, "dsb_beam_create: %s:%d: bad back-off weight '%s'", path, lineno, tok[cur + 1].c_str());
    int ids[kLmMaxOrder];
    for (int i = 0; i < cur; ++i) {
      const std::string& w = tok[1 + i];
      if (cur == 1) {
        auto it = lm.vocab.find(w);
        if (it == lm.vocab.end()) {
          ids[i] = (int)lm.words.size();
          lm.vocab[w] = ids[i];
          lm.words.push_back(w);
        } else {
          ids[i] = it->second;
        }
      } else {
        ids[i] = lm.index(w);
      }
    }
    if (cur == 1 && ids[0] == 0) lm.unk_prob = prob;
    lm.grams.push_back(HostLm::Gram{lm_chain_hash(ids, cur), cur, prob, backoff});
    ++seen[cur];
  }
  if (lm.order > 0 && !ended)
    return set_error(DSB_ERR_IO, "dsb_beam_create: '%s' ends without \\end\\ (truncated ARPA file?)", path);
  for (int n = 1; n <= lm.order; ++n)
    if (declared[n] >= 0 && declared[n] != seen[n])
      return set_error(DSB_ERR_IO, "dsb_beam_create: '%s' declares %lld %d-grams but holds %lld", path, declared[n], n, seen[n]);
  return finish_model(path, lm);
}

// ------------------------------------------------------------------------------------------ KenLM binary
namespace {
struct Reader {
  const unsigned char* p;
  size_t n;
  template <typename T> T at(size_t off) const {
    T v;
    memcpy(&v, p + off, sizeof(T));
    return v;
  }
};
uint64_t klm_buckets(uint64_t entries, float mult) {
  const uint64_t a = entries + 1, b = (uint64_t)(mult * (float)entries);
  return a > b ? a : b;
}
inline float neg_abs(float v) { return -fabsf(v); }
}  // namespace

int load_klm(const char* path, HostLm& lm) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return set_error(DSB_ERR_IO, "dsb_beam_create: cannot open language model '%s'", path);
  const size_t size = (size_t)f.tellg();
  std::vector<unsigned char> buf(size);
  f.seekg(0);
  if (size && !f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)size))
    return set_error(DSB_ERR_IO, "dsb_beam_create: cannot read '%s'", path);
  Reader r{buf.data(), size};
  static const char magic[] = "mmap lm http://kheafield.com/code format version 5\n\0";   // sizeof = 53 with the final NUL
  if (size < 108) return set_error(DSB_ERR_IO, "dsb_beam_create: '%s' is too short for a KenLM binary header", path);
  if (memcmp(r.p, magic, sizeof(magic)) != 0)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s' is a KenLM binary of format version '%c'; this reader "
                     "understands version 5", path, (char)r.p[sizeof(kKlmMagicPrefix)]);
  if (r.at<float>(56) != 0.0f || r.at<float>(60) != 1.0f || r.at<float>(64) != -0.5f || r.at<uint32_t>(68) != 1u ||
      r.at<uint32_t>(72) != 0xFFFFFFFFu || r.at<uint64_t>(80) != 1ull)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': KenLM sanity header mismatch (written on a platform "
                     "with another byte order or type sizes)", path);
  const int order = r.at<uint8_t>(88);
  const float mult = r.at<float>(92);
  const int type = r.at<int32_t>(96);
  const int has_vocab = r.at<uint8_t>(100);
  const uint32_t search_version = r.at<uint32_t>(104);
  static const char* names[] = {"probing", "rest-probing", "trie", "quantised trie", "array trie", "quantised array trie"};
  const char* tname = (type >= 0 && type < 6) ? names[type] : "unknown type";
  if (order < 1 || order > 16 || size < 108 + 8 * (size_t)order)
    return set_error(DSB_ERR_IO, "dsb_beam_create: '%s': bad KenLM header (order %d)", path, order);
  std::vector<uint64_t> counts(order);
  for (int i = 0; i < order; ++i) counts[i] = r.at<uint64_t>(108 + 8 * (size_t)i);
  if (type != 0)
    return set_error(DSB_ERR_UNSUPPORTED,
                     "dsb_beam_create: '%s' is a KenLM binary (format version 5, %s, order %d, %llu unigrams); only the "
                     "probing model (build_binary's default) is read -- rebuild it with `build_binary probing`, or pass "
                     "the .arpa it was built from", path, tname, order, (unsigned long long)counts[0]);
  if (order < 2 || order > kLmMaxOrder)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': LM order %d outside [2,%d]", path, order, kLmMaxOrder);
  if (!has_vocab)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s' was written without its vocabulary strings (build_binary "
                     "-i / include_vocab off); the dictionary of the decoder needs them", path);
  if (search_version != 0 || !(mult > 1.0f) || !(mult < 100.f))
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': probing search version %u / multiplier %g not understood",
                     path, search_version, (double)mult);

  const size_t header = (108 + 8 * (size_t)order + 7) / 8 * 8;
  const uint64_t vb = klm_buckets(counts[0], mult);
  const size_t vocab_off = header, vocab_bytes = 8 + (size_t)vb * 12;
  size_t off = vocab_off + vocab_bytes;
  const size_t uni_off = off;
  off += ((size_t)counts[0] + 1) * 8;
  std::vector<size_t> tab_off(order + 1, 0);
  std::vector<uint64_t> tab_buckets(order + 1, 0);
  for (int n = 2; n <= order; ++n) {
    tab_off[n] = off;
    tab_buckets[n] = klm_buckets(counts[n - 1], mult);
    off += (size_t)tab_buckets[n] * (n < order ? 16 : 12);
  }
  const size_t str_off = off;
  if (str_off >= size)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': the header describes %zu bytes of tables but the file "
                     "has %zu (layout not understood)", path, str_off, size);
  if (r.at<uint32_t>(vocab_off) != 0)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': probing vocabulary version %u", path, r.at<uint32_t>(vocab_off));
  const uint32_t bound = r.at<uint32_t>(vocab_off + 4);
  if (bound < 1 || bound > counts[0] + 1)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': vocabulary bound %u vs %llu unigrams", path, bound,
                     (unsigned long long)counts[0]);

  // vocabulary strings: exactly `bound` NUL-terminated words, "<unk>" first, ending at the end of the file
  lm.vocab.clear();
  lm.words.clear();
  {
    size_t p = str_off;
    while (p < size) {
      const void* z = memchr(r.p + p, 0, size - p);
      if (!z) return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': unterminated vocabulary string", path);
      const size_t len = (size_t)((const unsigned char*)z - (r.p + p));
      lm.words.emplace_back(reinterpret_cast<const char*>(r.p + p), len);
      p += len + 1;
    }
    if (lm.words.size() != bound || lm.words[0] != "<unk>")
      return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': %zu vocabulary strings for bound %u (first '%s'): "
                       "layout not understood", path, lm.words.size(), bound, lm.words.empty() ? "" : lm.words[0].c_str());
  }
  // every word must be where its MurmurHash says in the probing table, with its own id
  const size_t vt = vocab_off + 8;
  for (uint32_t id = 1; id < bound; ++id) {
    const std::string& w = lm.words[id];
    const uint64_t key = murmur_hash64a(w.data(), w.size(), 0);
    uint64_t slot = key % vb;
    bool found = false;
    for (uint64_t probes = 0; probes < vb; ++probes) {
      const uint64_t k = r.at<uint64_t>(vt + (size_t)slot * 12);
      if (k == key) { found = r.at<uint32_t>(vt + (size_t)slot * 12 + 8) == id; break; }
      if (k == 0) break;
      if (++slot == vb) slot = 0;
    }
    if (!found)
      return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': word %u ('%s') is not where its hash says in the "
                       "probing vocabulary: layout not understood", path, id, w.c_str());
    lm.vocab[w] = (int)id;
  }
  lm.vocab["<unk>"] = 0;
  lm.vocab["<UNK>"] = 0;

  lm.order = order;
  lm.grams.clear();
  for (uint32_t id = 0; id < bound; ++id) {
    const float p = neg_abs(r.at<float>(uni_off + (size_t)id * 8)), b = r.at<float>(uni_off + (size_t)id * 8 + 4);
    lm.grams.push_back(HostLm::Gram{(uint64_t)id, 1, p, b == 0.f ? 0.f : b});
  }
  lm.unk_prob = lm.grams[0].prob;
  for (int n = 2; n <= order; ++n) {
    const size_t esz = n < order ? 16 : 12;
    uint64_t used = 0;
    for (uint64_t s = 0; s < tab_buckets[n]; ++s) {
      const size_t e = tab_off[n] + (size_t)s * esz;
      const uint64_t key = r.at<uint64_t>(e);
      if (key == 0) continue;
      ++used;
      const float p = neg_abs(r.at<float>(e + 8)), b = n < order ? r.at<float>(e + 12) : 0.f;
      lm.grams.push_back(HostLm::Gram{key, n, p, b == 0.f ? 0.f : b});
    }
    if (used != counts[n - 1])
      return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: '%s': the %d-gram table holds %llu entries, the header says "
                       "%llu: layout not understood", path, n, (unsigned long long)used, (unsigned long long)counts[n - 1]);
  }
  return finish_model(path, lm);
}

}  // namespace dsb

// Host-only inspection of a language-model file through the loaders above (no CUDA call): order, number of n-grams
// and words, and an order-independent digest of the loaded model -- (key, order, prob, backoff) of every n-gram and
// (word, id) of every vocabulary entry.  Two files that load into the same model have the same digest.
extern "C" int dsb_lm_inspect(const char* path, int* order, int64_t* n_ngrams, int64_t* n_words, uint64_t* digest) {
  using namespace dsb;
  DSB_REQUIRE(path && path[0], "dsb_lm_inspect: null path");
  HostLm lm;
  const int klm = is_kenlm_binary(path);
  if (klm < 0) return klm;
  if (int e = klm ? load_klm(path, lm) : load_arpa(path, lm)) return e;
  auto mix = [](uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
  };
  uint64_t d = 0;
  for (const HostLm::Gram& g : lm.grams) {
    uint32_t pb, bb;
    const float bo = g.backoff == 0.f ? 0.f : g.backoff;   // -0.0 == 0.0
    memcpy(&pb, &g.prob, 4);
    memcpy(&bb, &bo, 4);
    d += mix(mix(g.key + (uint64_t)g.n) ^ (((uint64_t)pb << 32) | bb));
  }
  for (size_t i = 0; i < lm.words.size(); ++i)
    d += mix(murmur_hash64a(lm.words[i].data(), lm.words[i].size(), 0) ^ mix((uint64_t)i + 1));
  if (order) *order = lm.order;
  if (n_ngrams) *n_ngrams = (int64_t)lm.grams.size();
  if (n_words) *n_words = (int64_t)lm.words.size();
  if (digest) *digest = d;
  return 0;
}
