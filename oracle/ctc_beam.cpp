// Oracle (TEST INFRASTRUCTURE, not product code): CPU restatement of the CTC prefix beam search
// with n-gram LM scoring that the reference calls through the third-party `ctcdecode` package.
//
// The algorithm lives in a dependency that is ABSENT from /root/reference and unpinned
// (parlance/ctcdecode, git master; docs_source/installation.rst:25-30; derived from PaddlePaddle
// DeepSpeech's decoders), with KenLM + OpenFST inside.  Reference call sites:
//   danspeech/deepspeech/decoder.py:96,99-100 (construction), :140 (decode);
//   danspeech/DanSpeechRecognizer.py:89-92 (num_processes=6, cutoff_prob=1.0, cutoff_top_n=40).
// The reference holds no golden vector at this boundary that can be used offline => PARITY UNPINNED.
// This file restates the published algorithm:
//   ctc_beam_search_decoder / get_pruned_log_probs / get_beam_search_result  (ctc_beam_search_decoder.cpp, decoder_utils.cpp)
//   PathTrie::get_path_trie / iterate_to_vec / remove / get_path_vec            (path_trie.cpp)
//   Scorer::get_log_cond_prob / get_sent_log_prob / make_ngram / fill_dictionary (scorer.cpp)
//   KenLM back-off scoring of an ARPA model (lm/model.cc semantics: longest match + skipped back-offs)
// The dictionary FST (words + trailing space, determinised) is language-equivalent to a character trie.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

const float NUM_FLT_INF = std::numeric_limits<float>::max();
const float NUM_FLT_MIN = std::numeric_limits<float>::min();
const double NUM_FLT_LOGE = 0.4342944819;
const double OOV_SCORE = -1000.0;

template <typename T>
T log_sum_exp(const T& x, const T& y) {
  static T num_min = -std::numeric_limits<T>::max();
  if (x <= num_min) return y;
  if (y <= num_min) return x;
  T xmax = std::max(x, y);
  return std::log(std::exp(x - xmax) + std::exp(y - xmax)) + xmax;
}

std::vector<std::string> split_utf8(const std::string& s) {
  std::vector<std::string> out;
  for (size_t i = 0; i < s.size();) {
    unsigned char c = (unsigned char)s[i];
    size_t n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
    out.push_back(s.substr(i, n));
    i += n;
  }
  return out;
}

// ------------------------------------------------------------------ ARPA n-gram model (KenLM semantics)
struct NGramLM {
  int order = 0;
  std::unordered_map<std::string, int> vocab;   // word -> index; <unk> = 0
  std::vector<std::string> words;
  struct Entry { float prob; float backoff; };
  std::vector<std::map<std::vector<int>, Entry>> grams;   // grams[n-1]

  int index(const std::string& w) const {
    auto it = vocab.find(w);
    return it == vocab.end() ? 0 : it->second;
  }
  bool load(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    int cur = 0;
    vocab["<unk>"] = 0;
    words.push_back("<unk>");
    std::vector<long> counts;
    while (std::getline(f, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      if (line.rfind("ngram ", 0) == 0) {
        counts.push_back(atol(line.substr(line.find('=') + 1).c_str()));
        continue;
      }
      if (line[0] == '\\') {
        if (line.find("-grams:") != std::string::npos) {
          cur = atoi(line.c_str() + 1);
          if ((int)grams.size() < cur) grams.resize(cur);
          order = std::max(order, cur);
        } else if (line == "\\end\\") {
          break;
        }
        continue;
      }
      if (cur == 0) continue;
      std::vector<std::string> tok;
      std::stringstream ss(line);
      std::string t;
      while (ss >> t) tok.push_back(t);
      if ((int)tok.size() < cur + 1) continue;
      Entry e;
      e.prob = (float)atof(tok[0].c_str());
      e.backoff = (int)tok.size() > cur + 1 ? (float)atof(tok[cur + 1].c_str()) : 0.0f;
      std::vector<int> key;
      for (int i = 0; i < cur; ++i) {
        const std::string& w = tok[1 + i];
        if (cur == 1) {
          if (w == "<unk>") {
            key.push_back(0);
          } else {
            auto it = vocab.find(w);
            if (it == vocab.end()) {
              vocab[w] = (int)words.size();
              words.push_back(w);
            }
            key.push_back(vocab[w]);
          }
        } else {
          key.push_back(index(w));
        }
      }
      grams[cur - 1][key] = e;
    }
    return order > 0;
  }
  // log10 P(w | ctx) with KenLM back-off: longest matching n-gram plus back-offs of skipped contexts.
  double cond_log10(const std::vector<int>& ctx_in, int w) const {
    std::vector<int> ctx = ctx_in;
    if ((int)ctx.size() > order - 1) ctx.erase(ctx.begin(), ctx.end() - (order - 1));
    double bo = 0.0;
    for (;;) {
      std::vector<int> key = ctx;
      key.push_back(w);
      auto& tab = grams[key.size() - 1];
      auto it = tab.find(key);
      if (it != tab.end()) return bo + it->second.prob;
      if (ctx.empty()) {
        auto u = grams[0].find(std::vector<int>{0});
        return bo + (u != grams[0].end() ? u->second.prob : -100.0);
      }
      auto c = grams[ctx.size() - 1].find(ctx);
      if (c != grams[ctx.size() - 1].end()) bo += c->second.backoff;
      ctx.erase(ctx.begin());
    }
  }
};

// ------------------------------------------------------------------ dictionary (trie of word + ' ')
struct Dictionary {
  std::vector<std::map<int, int>> next;   // state -> (label -> state)
  std::vector<bool> is_final;
  Dictionary() { next.emplace_back(); is_final.push_back(false); }
  void add(const std::vector<int>& labels) {
    int s = 0;
    for (int l : labels) {
      auto it = next[s].find(l);
      if (it == next[s].end()) {
        next[s][l] = (int)next.size();
        s = (int)next.size();
        next.emplace_back();
        is_final.push_back(false);
      } else {
        s = it->second;
      }
    }
    is_final[s] = true;
  }
};

// ------------------------------------------------------------------ scorer
struct PathTrie;
struct Scorer {
  double alpha, beta;
  NGramLM lm;
  bool char_based = true;
  size_t max_order = 0;
  int space_id = -1;
  std::vector<std::string> char_list;
  std::unordered_map<std::string, int> char_map;
  std::unique_ptr<Dictionary> dictionary;
  const std::string START = "<s>", END = "</s>", UNK = "<unk>";

  bool init(double a, double b, const std::string& path, const std::vector<std::string>& labels) {
    alpha = a;
    beta = b;
    char_list = labels;
    for (size_t i = 0; i < labels.size(); ++i) {
      if (labels[i] == " ") space_id = (int)i;
      char_map[labels[i]] = (int)i;
    }
    if (!lm.load(path)) return false;
    max_order = lm.order;
    for (const std::string& w : lm.words)
      if (w != UNK && w != START && w != END && split_utf8(w).size() > 1) char_based = false;
    if (!char_based) {
      dictionary.reset(new Dictionary());
      for (const std::string& w : lm.words) {
        std::vector<int> ids;
        bool ok = true;
        for (const std::string& c : split_utf8(w)) {
          auto it = char_map.find(c);
          if (it == char_map.end()) { ok = false; break; }
          ids.push_back(it->second);
        }
        if (!ok || ids.empty()) continue;
        ids.push_back(space_id);
        dictionary->add(ids);
      }
    }
    return true;
  }
  double get_log_cond_prob(const std::vector<std::string>& words) const {
    double cond = 0.0;
    std::vector<int> ctx;
    for (size_t i = 0; i < words.size(); ++i) {
      int wi = lm.index(words[i]);
      if (wi == 0) return OOV_SCORE;
      cond = lm.cond_log10(ctx, wi);
      ctx.push_back(wi);
    }
    return cond / NUM_FLT_LOGE;
  }
  double get_sent_log_prob(const std::vector<std::string>& words) const {
    std::vector<std::string> sentence;
    if (words.empty()) {
      for (size_t i = 0; i < max_order; ++i) sentence.push_back(START);
    } else {
      for (size_t i = 0; i < max_order - 1; ++i) sentence.push_back(START);
      sentence.insert(sentence.end(), words.begin(), words.end());
    }
    sentence.push_back(END);
    double score = 0.0;
    for (size_t i = 0; i + max_order <= sentence.size(); ++i)
      score += get_log_cond_prob(std::vector<std::string>(sentence.begin() + i, sentence.begin() + i + max_order));
    return score;
  }
  std::string vec2str(const std::vector<int>& v) const {
    std::string s;
    for (int i : v) s += char_list[i];
    return s;
  }
  std::vector<std::string> split_labels(const std::vector<int>& labels) const {
    if (labels.empty()) return {};
    std::string s = vec2str(labels);
    std::vector<std::string> words;
    if (char_based) return split_utf8(s);
    std::string cur;
    for (char c : s) {
      if (c == ' ') { if (!cur.empty()) words.push_back(cur); cur.clear(); }
      else cur += c;
    }
    if (!cur.empty()) words.push_back(cur);
    return words;
  }
  std::vector<std::string> make_ngram(PathTrie* prefix) const;
};

// ------------------------------------------------------------------ prefix trie
struct PathTrie {
  float log_prob_b_prev = -NUM_FLT_INF, log_prob_nb_prev = -NUM_FLT_INF;
  float log_prob_b_cur = -NUM_FLT_INF, log_prob_nb_cur = -NUM_FLT_INF;
  float log_prob_c = -NUM_FLT_INF, score = -NUM_FLT_INF, approx_ctc = -NUM_FLT_INF;
  int character = -1, timestep = 0;
  PathTrie* parent = nullptr;
  bool exists_ = true;
  const Dictionary* dictionary_ = nullptr;
  int dictionary_state_ = 0;
  std::vector<std::pair<int, PathTrie*>> children_;

  ~PathTrie() { for (auto& c : children_) delete c.second; }

  PathTrie* get_path_trie(int new_char, int new_timestep, float cur_log_prob_c) {
    auto child = children_.begin();
    for (; child != children_.end(); ++child) {
      if (child->first == new_char) {
        if (child->second->log_prob_c < cur_log_prob_c) {
          child->second->log_prob_c = cur_log_prob_c;
          child->second->timestep = new_timestep;
        }
        break;
      }
    }
    if (child != children_.end()) {
      if (!child->second->exists_) {
        child->second->exists_ = true;
        child->second->log_prob_b_prev = child->second->log_prob_nb_prev = -NUM_FLT_INF;
        child->second->log_prob_b_cur = child->second->log_prob_nb_cur = -NUM_FLT_INF;
      }
      return child->second;
    }
    int next_state = 0;
    if (dictionary_) {
      auto& arcs = dictionary_->next[dictionary_state_];
      auto it = arcs.find(new_char);
      if (it == arcs.end()) return nullptr;   // would leave the LM vocabulary
      next_state = dictionary_->is_final[it->second] ? 0 : it->second;
    }
    PathTrie* n = new PathTrie;
    n->character = new_char;
    n->timestep = new_timestep;
    n->parent = this;
    n->dictionary_ = dictionary_;
    n->dictionary_state_ = next_state;
    n->log_prob_c = cur_log_prob_c;
    children_.push_back(std::make_pair(new_char, n));
    return n;
  }
  PathTrie* get_path_vec(std::vector<int>& output, std::vector<int>& timesteps, int stop = -1,
                         size_t max_steps = std::numeric_limits<size_t>::max()) {
    if (character == stop || character == -1 || output.size() == max_steps) {
      std::reverse(output.begin(), output.end());
      std::reverse(timesteps.begin(), timesteps.end());
      return this;
    }
    output.push_back(character);
    timesteps.push_back(timestep);
    return parent->get_path_vec(output, timesteps, stop, max_steps);
  }
  void iterate_to_vec(std::vector<PathTrie*>& output) {
    if (exists_) {
      log_prob_b_prev = log_prob_b_cur;
      log_prob_nb_prev = log_prob_nb_cur;
      log_prob_b_cur = -NUM_FLT_INF;
      log_prob_nb_cur = -NUM_FLT_INF;
      score = log_sum_exp(log_prob_b_prev, log_prob_nb_prev);
      output.push_back(this);
    }
    for (auto& c : children_) c.second->iterate_to_vec(output);
  }
  void remove() {
    exists_ = false;
    if (children_.empty()) {
      PathTrie* p = parent;
      for (auto it = p->children_.begin(); it != p->children_.end(); ++it)
        if (it->first == character) { p->children_.erase(it); break; }
      if (p->children_.empty() && !p->exists_) p->remove();
      delete this;
    }
  }
  bool is_empty() const { return character == -1; }
};

std::vector<std::string> Scorer::make_ngram(PathTrie* prefix) const {
  std::vector<std::string> ngram;
  PathTrie* current = prefix;
  PathTrie* node = nullptr;
  for (int order = 0; order < (int)max_order; ++order) {
    std::vector<int> pv, ps;
    if (char_based) {
      node = current->get_path_vec(pv, ps, -1, 1);
      current = node;
    } else {
      node = current->get_path_vec(pv, ps, space_id);
      current = node->parent;   // skip the space
    }
    ngram.push_back(vec2str(pv));
    if (node->character == -1) {
      for (int i = 0; i < (int)max_order - order - 1; ++i) ngram.push_back(START);
      break;
    }
  }
  std::reverse(ngram.begin(), ngram.end());
  return ngram;
}

bool prefix_compare(const PathTrie* x, const PathTrie* y) {
  if (x->score == y->score) return x->character < y->character;
  return x->score > y->score;
}

struct Output { std::vector<int> tokens, timesteps; };

std::vector<std::pair<size_t, float>> get_pruned_log_probs(const std::vector<double>& prob_step, double cutoff_prob,
                                                           size_t cutoff_top_n) {
  std::vector<std::pair<int, double>> prob_idx;
  double log_cutoff_prob = std::log(cutoff_prob);
  for (size_t i = 0; i < prob_step.size(); ++i) prob_idx.push_back(std::make_pair((int)i, prob_step[i]));
  size_t cutoff_len = prob_step.size();
  if (log_cutoff_prob < 0.0 || cutoff_top_n < cutoff_len) {
    std::sort(prob_idx.begin(), prob_idx.end(),
              [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second > b.second; });
    if (log_cutoff_prob < 0.0) {
      double cum_prob = 0.0;
      cutoff_len = 0;
      for (size_t i = 0; i < prob_idx.size(); ++i) {
        cum_prob = log_sum_exp(cum_prob, std::log(prob_idx[i].second));
        cutoff_len += 1;
        if (cum_prob >= cutoff_prob) break;
      }
    }
    cutoff_len = std::min(cutoff_len, cutoff_top_n);
    prob_idx.resize(cutoff_len);
  }
  std::vector<std::pair<size_t, float>> out;
  for (size_t i = 0; i < cutoff_len; ++i)
    out.push_back(std::make_pair((size_t)prob_idx[i].first, (float)std::log(prob_idx[i].second + NUM_FLT_MIN)));
  return out;
}

std::vector<std::pair<double, Output>> beam_search(const std::vector<std::vector<double>>& probs_seq, size_t n_labels,
                                                   int space_id_in, size_t beam_size, double cutoff_prob,
                                                   size_t cutoff_top_n, size_t blank_id, Scorer* ext_scorer) {
  const size_t T = probs_seq.size();
  int space_id = space_id_in >= 0 ? space_id_in : -2;
  PathTrie root;
  root.score = root.log_prob_b_prev = 0.0;
  std::vector<PathTrie*> prefixes;
  prefixes.push_back(&root);
  if (ext_scorer && !ext_scorer->char_based) root.dictionary_ = ext_scorer->dictionary.get();
  (void)n_labels;

  for (size_t t = 0; t < T; ++t) {
    const std::vector<double>& prob = probs_seq[t];
    float min_cutoff = -NUM_FLT_INF;
    bool full_beam = false;
    if (ext_scorer) {
      size_t np = std::min(prefixes.size(), beam_size);
      std::sort(prefixes.begin(), prefixes.begin() + np, prefix_compare);
      float blank_prob = std::log(prob[blank_id]);
      min_cutoff = prefixes[np - 1]->score + blank_prob - std::max(0.0, ext_scorer->beta);
      full_beam = (np == beam_size);
    }
    auto log_prob_idx = get_pruned_log_probs(prob, cutoff_prob, cutoff_top_n);
    for (size_t index = 0; index < log_prob_idx.size(); ++index) {
      size_t c = log_prob_idx[index].first;
      float log_prob_c = log_prob_idx[index].second;
      for (size_t i = 0; i < prefixes.size() && i < beam_size; ++i) {
        PathTrie* prefix = prefixes[i];
        if (full_beam && log_prob_c + prefix->score < min_cutoff) break;
        if (c == blank_id) {
          prefix->log_prob_b_cur = log_sum_exp(prefix->log_prob_b_cur, log_prob_c + prefix->score);
          continue;
        }
        if ((int)c == prefix->character)
          prefix->log_prob_nb_cur = log_sum_exp(prefix->log_prob_nb_cur, log_prob_c + prefix->log_prob_nb_prev);
        PathTrie* prefix_new = prefix->get_path_trie((int)c, (int)t, log_prob_c);
        if (prefix_new != nullptr) {
          float log_p = -NUM_FLT_INF;
          if ((int)c == prefix->character && prefix->log_prob_b_prev > -NUM_FLT_INF)
            log_p = log_prob_c + prefix->log_prob_b_prev;
          else if ((int)c != prefix->character)
            log_p = log_prob_c + prefix->score;
          if (ext_scorer && ((int)c == space_id || ext_scorer->char_based)) {
            PathTrie* to_score = ext_scorer->char_based ? prefix_new : prefix;
            float score = (float)(ext_scorer->get_log_cond_prob(ext_scorer->make_ngram(to_score)) * ext_scorer->alpha);
            log_p += score;
            log_p += (float)ext_scorer->beta;
          }
          prefix_new->log_prob_nb_cur = log_sum_exp(prefix_new->log_prob_nb_cur, log_p);
        }
      }
    }
    prefixes.clear();
    root.iterate_to_vec(prefixes);
    if (prefixes.size() >= beam_size) {
      std::nth_element(prefixes.begin(), prefixes.begin() + beam_size, prefixes.end(), prefix_compare);
      for (size_t i = beam_size; i < prefixes.size(); ++i) prefixes[i]->remove();
      prefixes.resize(beam_size);
    }
  }
  if (ext_scorer && !ext_scorer->char_based) {
    for (size_t i = 0; i < beam_size && i < prefixes.size(); ++i) {
      PathTrie* prefix = prefixes[i];
      if (!prefix->is_empty() && prefix->character != space_id) {
        float score = (float)(ext_scorer->get_log_cond_prob(ext_scorer->make_ngram(prefix)) * ext_scorer->alpha);
        score += (float)ext_scorer->beta;
        prefix->score += score;
      }
    }
  }
  size_t np = std::min(prefixes.size(), beam_size);
  std::sort(prefixes.begin(), prefixes.begin() + np, prefix_compare);
  for (size_t i = 0; i < beam_size && i < prefixes.size(); ++i) {
    double approx_ctc = prefixes[i]->score;
    if (ext_scorer) {
      std::vector<int> output, timesteps;
      prefixes[i]->get_path_vec(output, timesteps);
      auto words = ext_scorer->split_labels(output);
      approx_ctc = approx_ctc - output.size() * ext_scorer->beta;
      approx_ctc -= ext_scorer->get_sent_log_prob(words) * ext_scorer->alpha;
    }
    prefixes[i]->approx_ctc = (float)approx_ctc;
  }
  std::vector<PathTrie*> top(prefixes.begin(), prefixes.begin() + np);
  std::sort(top.begin(), top.end(), prefix_compare);
  std::vector<std::pair<double, Output>> result;
  for (PathTrie* p : top) {
    Output o;
    p->get_path_vec(o.tokens, o.timesteps);
    result.emplace_back(-(double)p->approx_ctc, o);
  }
  return result;
}

struct Decoder {
  std::vector<std::string> labels;
  std::unique_ptr<Scorer> scorer;
  size_t beam = 100, cutoff_top_n = 40, blank = 0;
  double cutoff_prob = 1.0;
  int space_id = -1;
};

}  // namespace

extern "C" {

void* oracle_beam_create(const char* labels_blob, int n_labels, const char* lm_path, double alpha, double beta,
                         int cutoff_top_n, double cutoff_prob, int beam_width, int blank_id) {
  Decoder* d = new Decoder();
  const char* p = labels_blob;
  for (int i = 0; i < n_labels; ++i) {
    d->labels.emplace_back(p);
    p += d->labels.back().size() + 1;
    if (d->labels.back() == " ") d->space_id = i;
  }
  d->beam = beam_width;
  d->cutoff_top_n = cutoff_top_n;
  d->cutoff_prob = cutoff_prob;
  d->blank = blank_id;
  if (lm_path && lm_path[0]) {
    d->scorer.reset(new Scorer());
    if (!d->scorer->init(alpha, beta, lm_path, d->labels)) {
      delete d;
      return nullptr;
    }
  }
  return d;
}

void oracle_beam_destroy(void* h) { delete static_cast<Decoder*>(h); }
int oracle_beam_is_char_based(void* h) {
  Decoder* d = static_cast<Decoder*>(h);
  return d->scorer ? (d->scorer->char_based ? 1 : 0) : -1;
}
int oracle_beam_lm_order(void* h) {
  Decoder* d = static_cast<Decoder*>(h);
  return d->scorer ? (int)d->scorer->max_order : 0;
}
double oracle_lm_cond_log_prob(void* h, const char* words_blob, int n_words) {
  Decoder* d = static_cast<Decoder*>(h);
  std::vector<std::string> w;
  const char* p = words_blob;
  for (int i = 0; i < n_words; ++i) { w.emplace_back(p); p += w.back().size() + 1; }
  return d->scorer->get_log_cond_prob(w);
}

// probs [B,T,C] float; seq_lens [B]; outputs as ctcdecode: tokens/timesteps [B,beam,T], scores [B,beam], lens [B,beam]
void oracle_beam_decode(void* h, const float* probs, const int* seq_lens, int B, int T, int C, int* out_tokens,
                        int* out_timesteps, float* out_scores, int* out_lens, int num_threads) {
  Decoder* d = static_cast<Decoder*>(h);
  const size_t beam = d->beam;
  std::memset(out_tokens, 0, sizeof(int) * (size_t)B * beam * T);
  std::memset(out_timesteps, 0, sizeof(int) * (size_t)B * beam * T);
  std::memset(out_scores, 0, sizeof(float) * (size_t)B * beam);
  std::memset(out_lens, 0, sizeof(int) * (size_t)B * beam);
  auto work = [&](int b) {
    int len = std::min(seq_lens ? seq_lens[b] : T, T);
    std::vector<std::vector<double>> seq(len, std::vector<double>(C));
    for (int t = 0; t < len; ++t)
      for (int c = 0; c < C; ++c) seq[t][c] = probs[((size_t)b * T + t) * C + c];
    auto res = beam_search(seq, C, d->space_id, beam, d->cutoff_prob, d->cutoff_top_n, d->blank, d->scorer.get());
    for (size_t p = 0; p < res.size() && p < beam; ++p) {
      const Output& o = res[p].second;
      for (size_t i = 0; i < o.tokens.size(); ++i) {
        out_tokens[((size_t)b * beam + p) * T + i] = o.tokens[i];
        out_timesteps[((size_t)b * beam + p) * T + i] = o.timesteps[i];
      }
      out_scores[(size_t)b * beam + p] = (float)res[p].first;
      out_lens[(size_t)b * beam + p] = (int)o.tokens.size();
    }
  };
  num_threads = std::max(1, std::min(num_threads, B));
  std::vector<std::thread> pool;
  for (int w = 0; w < num_threads; ++w)
    pool.emplace_back([&, w]() { for (int b = w; b < B; b += num_threads) work(b); });
  for (auto& t : pool) t.join();
}

}  // extern "C"
