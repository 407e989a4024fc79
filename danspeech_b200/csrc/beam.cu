// GPU CTC prefix beam search with a device-resident n-gram LM and dictionary.
//
// Replaces the ctcdecode.CTCBeamDecoder object the reference builds at
// danspeech/deepspeech/decoder.py:99-100 and calls at decoder.py:140 (third-party parlance/ctcdecode:
// prefix beam search + KenLM + dictionary FST on 6 host threads, DanSpeechRecognizer.py:89-92).
// Semantics follow SURVEY.md appendix B: blank / repeat / extend updates per prefix, LM applied at
// word boundaries (word LM, with the vocabulary trie as dictionary) or at every symbol (character LM),
// log(p + FLT_MIN) scores, prefix_compare ordering (score desc, then smaller last symbol), OOV = -1000,
// final partial-word scoring, reported score = -(score - len*beta - alpha*sentence_log_prob).
//
// Mapping: one CTA of 512 threads per utterance (utterances are independent; a batch of 64 occupies 64 SMs).  The
// live beam (<= 128 prefixes), the step's candidates (beam x C) and, with a word LM, the dictionary arcs of every
// live prefix live in shared memory; a step is
//   log-probs -> LM score of closing the current word for NEW prefixes only (cached with the prefix; the back-off
//   chain's table probes issued together) -> candidate scores (pass A settles the symbols the dictionary does not
//   allow and collects the rest per warp, pass B scores them on full warps) -> merge children that are already in
//   the beam -> dense list of valid candidates (two passes, no shared counter) -> top-W in order by rank counting
//   (few candidates) or bitonic sort -> materialise the surviving prefixes in a per-utterance node arena.
// The work is latency / random-access bound, reported as utterances/s and steps/s.
#include "model_types.cuh"
#include "lm_host.h"
#include <float.h>
#include <math.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>

namespace dsb {

constexpr int BM_MAXW = 128;       // max beam width
constexpr int BM_MAXC = 64;        // max classes
constexpr int BM_HIST = 4;         // max LM order - 1
constexpr int BM_THREADS = 512;   // measured: 10.2 ms per batch of 64 vs 12.5 ms with 256 threads
constexpr float BM_NEG = -FLT_MAX;
constexpr float BM_OOV = -1000.0f;
constexpr float BM_LOGE = 0.4342944819f;

struct LmEntry {
  uint64_t k0, k1;
  float prob, backoff;
  uint32_t used, pad;
};

struct BeamTables {
  const int32_t* trans;      // [n_states][C] dictionary arcs (-1 = none; final states map to 0)
  const int32_t* word_at;    // [n_states] vocabulary id of the word spelled by the state (-1 = not a word)
  const int32_t* char_word;  // [C] vocabulary id of each single-symbol token (character LM)
  const LmEntry* lm;
  uint32_t lm_mask;
  int order, id_bos, id_eos;
  int has_lm, char_based;
  float alpha, beta, unk_prob;
};

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
// Table key of an n-gram: (k0, k1) = (KenLM chain hash over the word ids -- the id itself for a unigram --, order).
// A KenLM probing binary stores exactly these hashes and nothing else about an n-gram (lm_host.h), so ARPA text and
// .klm files load into the same table.
__host__ __device__ inline uint32_t key_hash(uint64_t k0, uint64_t k1) { return (uint32_t)mix64(k0 ^ mix64(k1 + 0x9e3779b97f4a7c15ULL)); }

// continue a lookup whose first slot `e` (at index h) has already been loaded
__device__ __forceinline__ bool lm_resolve(const LmEntry* __restrict__ lm, uint32_t mask, LmEntry e, uint32_t h, uint64_t k0,
                                           uint64_t k1, float& prob, float& backoff) {
  for (;;) {
    if (!e.used) return false;
    if (e.k0 == k0 && e.k1 == k1) {
      prob = e.prob;
      backoff = e.backoff;
      return true;
    }
    h = (h + 1) & mask;
    e = lm[h];
  }
}

// Scorer::get_log_cond_prob: natural-log P(words[n-1] | words[0..n-2]) with KenLM back-off; OOV anywhere -> -1000.
// The back-off chain needs up to 2*order-1 table lookups; their keys do not depend on each other's results, so the
// first probe of every one is issued up front (one L2 round trip instead of one per lookup) and the chain is then
// walked over the loaded entries.  Not inlined, and fed with scalars instead of the table struct: as an inlined
// function taking the kernel-parameter struct by reference it forced the parameters into local memory and the whole
// kernel to 255 registers.
__device__ __noinline__ float lm_log_cond_prob_impl(const LmEntry* __restrict__ lm, uint32_t mask, int order, float unk_prob,
                                                    const int* words, int n) {
  for (int i = 0; i < n; ++i)
    if (words[i] == 0) return BM_OOV;
  const int first = n > order ? n - order : 0;   // keep order-1 context words
  constexpr int MAXQ = BM_HIST + 1;
  const int nq = n - first;
  uint64_t fk0[MAXQ], fk1[MAXQ], ck0[MAXQ], ck1[MAXQ];
  uint32_t fh[MAXQ], ch[MAXQ];
  LmEntry fe[MAXQ], ce[MAXQ];
  // query q: the n-gram words[first+q .. n-1] (f) and its context words[first+q .. n-2] (c).  The chain hash runs from
  // the last word backwards, so the hash of a longer history is one combine step on top of the shorter one.
  uint64_t hf = 0, hc = 0;
#pragma unroll
  for (int q = MAXQ - 1; q >= 0; --q) {
    if (q < nq) {
      const int start = first + q;
      hf = (start == n - 1) ? (uint64_t)(uint32_t)words[start] : lm_combine_word_hash(hf, (uint32_t)words[start]);
      fk0[q] = hf;
      fk1[q] = (uint64_t)(n - start);
      fh[q] = key_hash(fk0[q], fk1[q]) & mask;
      fe[q] = lm[fh[q]];
      if (start < n - 1) {
        hc = (start == n - 2) ? (uint64_t)(uint32_t)words[start] : lm_combine_word_hash(hc, (uint32_t)words[start]);
        ck0[q] = hc;
        ck1[q] = (uint64_t)(n - 1 - start);
        ch[q] = key_hash(ck0[q], ck1[q]) & mask;
        ce[q] = lm[ch[q]];
      }
    }
  }
  float bo = 0.f;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    if (q < nq) {
      float pr, b;
      if (lm_resolve(lm, mask, fe[q], fh[q], fk0[q], fk1[q], pr, b)) return (bo + pr) / BM_LOGE;
      if (first + q < n - 1 && lm_resolve(lm, mask, ce[q], ch[q], ck0[q], ck1[q], pr, b)) bo += b;
    }
  }
  return (bo + unk_prob) / BM_LOGE;
}
#define lm_log_cond_prob(T, words, n) lm_log_cond_prob_impl((T).lm, (T).lm_mask, (T).order, (T).unk_prob, words, n)

__device__ __forceinline__ float lse2(float x, float y) {
  if (x <= BM_NEG) return y;
  if (y <= BM_NEG) return x;
  const float m = fmaxf(x, y);
  return logf(expf(x - m) + expf(y - m)) + m;
}
__device__ __forceinline__ uint32_t f2o_desc(float f) {
  uint32_t u = __float_as_uint(f);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;   // ascending order
  return ~u;                                    // descending
}

struct BeamState {   // one live-beam buffer (shared memory)
  int node[BM_MAXW], parent[BM_MAXW], ch[BM_MAXW], dstate[BM_MAXW], ts[BM_MAXW];
  int hist[BM_MAXW][BM_HIST];
  float bprev[BM_MAXW], nbprev[BM_MAXW], score[BM_MAXW], lpc[BM_MAXW];
  // word LM: alpha * log P(word | history) of closing the current word with a space, cached with the prefix (it only
  // depends on hist / dstate, which a surviving prefix keeps); lmok = 0 not computed yet, 1 valid, -1 no dictionary arc
  float lmsp[BM_MAXW];
  int lmwid[BM_MAXW], lmns[BM_MAXW], lmok[BM_MAXW];
  int rowok[BM_MAXW];   // dictionary row of this prefix's state is in the row cache
  int gprobe[BM_MAXW];  // this prefix may have children that exist in the trie outside the beam (see phase 3c)
};

struct BeamSmem {
  BeamState st[2];
  float lp[BM_MAXC];
  int allowed[BM_MAXC];
  float selfb[BM_MAXW], selfnb[BM_MAXW], selfscore[BM_MAXW];
  int pidx[BM_MAXW];
  int surv[BM_MAXW];   // live prefix k is in the next beam: its slot there + 1, else 0
  int pslot[BM_MAXW];  // slot of live prefix k's parent in the current beam (-1: the parent is not live)
  int count, arena_count;
  int n_dead;          // nodes that exist in the trie without being in the beam (kept for their descendants)
  int wsum[BM_THREADS / 32];
  float lpb_raw;
};

struct BeamParams {
  const float* probs;        // [B,T,C]
  const int32_t* seq_lens;   // [B] device
  int B, T, C, W, blank, space, cutoff_top_n;
  int n2;                    // capacity of the sort-key array (power of two >= W*C + W)
  float cutoff_prob;
  // arena [B][max_nodes]
  int32_t* a_parent;
  int32_t* a_info;           // ch | timestep << 8
  int32_t* a_wid;            // word id completed at this node (space nodes), else -1
  float* a_lpc;              // best symbol log-probability seen for this node (PathTrie::log_prob_c; its frame is in a_info)
  int32_t* a_ref;            // 1 while the node is in the beam + number of its children that exist; 0 = removed from the trie
  int max_nodes;
  // node identity: (parent node, symbol) -> arena node, open addressing, [B][h_cap].  A prefix that fell out of the beam
  // and comes back must be the SAME node (ctcdecode's PathTrie keeps removed nodes that still have descendants and
  // revives them, path_trie.cpp get_path_trie): its live descendants keep pointing at it, so a later extension of the
  // revived prefix merges into the live child instead of creating a second prefix with the same labels.
  uint32_t* h_keys;
  int32_t* h_vals;
  uint32_t h_mask;
  int h_shift;
  int32_t* words;            // [B][W][T+2] scratch for the sentence score
  int32_t* out_tokens;
  int32_t* out_ts;
  float* out_scores;
  int32_t* out_lens;
  long long* dbg;            // optional [12] per-phase cycle sums of utterance 0 (DSB_BEAM_DEBUG=1)
  BeamTables tab;
};

__global__ void __launch_bounds__(BM_THREADS)
beam_kernel(const BeamParams p) {
  extern __shared__ __align__(16) unsigned char bsm_raw[];
  BeamSmem& sm = *reinterpret_cast<BeamSmem*>(bsm_raw);
  uint64_t* keys = reinterpret_cast<uint64_t*>(bsm_raw + ((sizeof(BeamSmem) + 15) & ~(size_t)15));
  float* cand = reinterpret_cast<float*>(keys + p.n2);             // [W*C] child candidate scores
  int32_t* cand_aux = reinterpret_cast<int32_t*>(cand + p.W * p.C);   // [W*C][2] next dictionary state, completed word id
  // word LM: the dictionary arcs trans[dstate][:] of every live prefix, carried along with the prefix (two buffers like
  // the beam state), so that only NEW prefixes touch the table in global memory
  int32_t* rows0 = cand_aux + 2 * p.W * p.C;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int C = p.C, W = p.W;
  const BeamTables& T = p.tab;
  const int HN = T.order > 1 ? T.order - 1 : 0;
  const int len = min(p.seq_lens ? p.seq_lens[b] : p.T, p.T);
  int32_t* a_parent = p.a_parent + (size_t)b * p.max_nodes;
  int32_t* a_info = p.a_info + (size_t)b * p.max_nodes;
  int32_t* a_wid = p.a_wid + (size_t)b * p.max_nodes;
  float* a_lpc = p.a_lpc + (size_t)b * p.max_nodes;
  int32_t* a_ref = p.a_ref + (size_t)b * p.max_nodes;
  uint32_t* h_keys = p.h_keys + (size_t)b * (p.h_mask + 1);
  int32_t* h_vals = p.h_vals + (size_t)b * (p.h_mask + 1);

  int cur = 0, n_active = 1;
  if (tid == 0) {
    BeamState& s = sm.st[0];
    s.node[0] = 0; s.parent[0] = -1; s.ch[0] = -1; s.dstate[0] = 0; s.ts[0] = 0;
    for (int h = 0; h < BM_HIST; ++h) s.hist[0][h] = T.id_bos;
    s.bprev[0] = 0.f; s.nbprev[0] = BM_NEG; s.score[0] = 0.f; s.lpc[0] = BM_NEG;
    s.lmok[0] = 0; s.rowok[0] = 0; s.gprobe[0] = 0;
    a_parent[0] = -1; a_info[0] = 0xFF; a_wid[0] = -1;
    a_lpc[0] = BM_NEG; a_ref[0] = 1 << 28;   // the root is never removed
    sm.arena_count = 1;
    sm.n_dead = 0;
  }
  __syncthreads();

  const int i_first = tid / C, c_first = tid % C, i_step = BM_THREADS / C, c_step = BM_THREADS % C;
  // x / C for x < 2^16 by a multiply-high (exact: C <= 64); a runtime division is ~25 instructions, three of them on
  // the quarter-rate conversion pipe, and sat in most of the per-item loops
  const unsigned magicC = 0xFFFFFFFFu / (unsigned)C + 1u;
  auto divC = [&](int x) -> int { return C > 1 ? (int)__umulhi((unsigned)x, magicC) : x; };
  long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool dbg = p.dbg != nullptr && b == 0 && tid == 0;
  for (int t = 0; t < len; ++t) {
    long long c0 = dbg ? clock64() : 0;
    auto lap = [&](int i) { if (dbg) { const long long c1 = clock64(); ph[i] += c1 - c0; c0 = c1; } };
    BeamState& S = sm.st[cur];
    BeamState& Nx = sm.st[cur ^ 1];
    const float* pr = p.probs + ((size_t)b * p.T + t) * C;
    // ---- phase 0: log-probabilities, vocabulary pruning (get_pruned_log_probs) ----
    if (tid < C) {
      const float pv = pr[tid];
      sm.lp[tid] = (float)log((double)pv + (double)FLT_MIN);
      if (tid == p.blank) sm.lpb_raw = (float)log((double)pv);
      int allowed = 1;
      if (p.cutoff_prob < 1.0f || p.cutoff_top_n < C) {
        int rank = 0;
        for (int c2 = 0; c2 < C; ++c2) {
          const float q = pr[c2];
          rank += (q > pv) || (q == pv && c2 < tid);
        }
        if (p.cutoff_prob < 1.0f) {
          // upstream accumulates log_sum_exp(cum, log p) from cum = 0.0 and stops at cum >= cutoff_prob
          // (i.e. log(1 + sum p) >= cutoff_prob); reproduce on the sorted order
          double sum_before = 0.0;   // sum of probabilities ranked strictly before this symbol, plus own
          for (int c2 = 0; c2 < C; ++c2) {
            const float q = pr[c2];
            if ((q > pv) || (q == pv && c2 <= tid)) sum_before += (double)q;
          }
          // symbol with rank r is kept iff the stop condition was not met at any earlier rank:
          // cum after rank r-1 = log(1 + sum_{rank<r}) < cutoff
          const double prev = sum_before - (double)pv;
          allowed = (rank == 0) || (log(1.0 + prev) < (double)p.cutoff_prob);
        }
        if (rank >= p.cutoff_top_n) allowed = 0;
      }
      sm.allowed[tid] = allowed;
    }
#pragma unroll 1
    for (int k = tid; k < n_active; k += BM_THREADS) {
      sm.pidx[k] = -1;
      sm.pslot[k] = -1;
      sm.selfnb[k] = BM_NEG;
      sm.selfb[k] = BM_NEG;
    }
    if (tid == 0) sm.count = 0;
    __syncthreads();
    lap(0);
    const bool full_beam = T.has_lm && n_active == W;
    const float min_cutoff = T.has_lm ? S.score[n_active - 1] + sm.lpb_raw - fmaxf(0.f, T.beta) : BM_NEG;

    // ---- phase 2: which live prefixes are children of other live prefixes ----
    for (int k = tid >> 3; k < n_active; k += BM_THREADS >> 3) {   // 8 threads per prefix k scan the beam for its parent
      const int pk = S.parent[k];
      for (int i = tid & 7; i < n_active; i += 8)
        if (S.node[i] == pk && i != k) { sm.pidx[k] = i; sm.pslot[k] = i; }
    }
    lap(1);
    // ---- phase 2b (word LM): the LM score of closing the current word, one thread per prefix that has not got it
    //      cached yet (new prefixes); the dependent hash probes of all prefixes run side by side instead of inside
    //      the (prefix, symbol) loop ----
    const bool word_lm = T.has_lm && !T.char_based;
    int32_t* rowsS = rows0 + cur * W * C;
    int32_t* rowsN = rows0 + (cur ^ 1) * W * C;
    if (word_lm) {
#pragma unroll 1
      for (int idx = tid; idx < n_active * C; idx += BM_THREADS) {
        const int k = divC(idx);
        if (S.rowok[k] == 0) rowsS[idx] = T.trans[(size_t)S.dstate[k] * C + (idx - k * C)];
      }
      __syncthreads();
    }
    lap(2);
    if (word_lm && p.space >= 0) {
#pragma unroll 1
      for (int k = tid; k < n_active; k += BM_THREADS) {
        if (S.lmok[k] == 0) {
          const int ns = rowsS[k * C + p.space];
          int ok = -1;
          if (ns >= 0) {
            int words[BM_HIST + 1];
            for (int h = 0; h < HN; ++h) words[h] = S.hist[k][h];
            const int wid = T.word_at[S.dstate[k]];
            words[HN] = wid < 0 ? 0 : wid;
            S.lmsp[k] = lm_log_cond_prob(T, words, HN + 1) * T.alpha;
            S.lmwid[k] = wid;
            ok = 1;
          }
          S.lmns[k] = ns;
          S.lmok[k] = ok;
        }
      }
      __syncthreads();
    }
    lap(3);
    // ---- phase 3: candidate scores, one thread per (prefix, symbol) ----
    // Pass A over all (prefix, symbol) items (indices advance without a division): a symbol the dictionary does not
    // allow after this prefix (the common case with a word LM) is settled here; the others are collected in a work
    // list so that pass B runs its long, branchy body on full warps instead of on one or two lanes of every warp.
    // Every warp keeps its own list (no shared counter: same-address shared-memory atomics from 16 warps serialise at
    // ~90 cycles each and were the largest item of the step) and runs pass B over it without a block barrier.
    const int iters = (n_active * C + BM_THREADS - 1) / BM_THREADS;
    int* wlist = reinterpret_cast<int*>(keys) + (tid >> 5) * (32 * iters);   // the key array is free until the compaction
    int n_work = 0;
    {
      const unsigned lane = tid & 31, lt = (1u << lane) - 1u;
      int i = i_first, c = c_first;
      const int total = n_active * C;
#pragma unroll 1
      for (int base = 0; base < total; base += BM_THREADS) {
        const int idx = base + tid;
        bool todo = false;
        if (idx < total) {
          todo = c == p.blank || c == S.ch[i] || !word_lm || (c == p.space ? S.lmns[i] : rowsS[idx]) >= 0;
          if (!todo) cand[idx] = BM_NEG;
        }
        const unsigned m = __ballot_sync(0xffffffffu, todo);
        if (todo) wlist[n_work + __popc(m & lt)] = idx;
        n_work += __popc(m);
        i += i_step;
        c += c_step;
        if (c >= C) { c -= C; ++i; }
      }
    }
    __syncwarp();
    lap(10);
#pragma unroll 1
    for (int li = tid & 31; li < n_work; li += 32) {
      const int idx = wlist[li];
      const int i = divC(idx), c = idx - i * C;
      const int ci = S.ch[i];
      const float lpc = sm.lp[c], sc = S.score[i];
      const bool pass = sm.allowed[c] && !(full_beam && lpc + sc < min_cutoff);
      if (c == p.blank) {
        if (pass) sm.selfb[i] = lpc + sc;
        cand[idx] = BM_NEG;
        continue;
      }
      const bool word_space = word_lm && c == p.space;
      int next_state = 0, wid = -1;
      if (word_lm) next_state = word_space ? S.lmns[i] : rowsS[idx];
      float v = BM_NEG;
      if (pass) {
        if (c == ci) sm.selfnb[i] = lpc + S.nbprev[i];   // repeated symbol without a blank
        if (next_state >= 0) {
          float log_p = BM_NEG;
          if (c == ci) {
            if (S.bprev[i] > BM_NEG) log_p = lpc + S.bprev[i];
          } else {
            log_p = lpc + sc;
          }
          if (word_space) {
            wid = S.lmwid[i];
            log_p += S.lmsp[i];
            log_p += T.beta;
          } else if (T.has_lm && T.char_based) {
            int words[BM_HIST + 1];
            for (int h = 0; h < HN; ++h) words[h] = S.hist[i][h];
            wid = T.char_word[c];
            words[HN] = wid < 0 ? 0 : wid;
            const float lm = lm_log_cond_prob(T, words, HN + 1) * T.alpha;
            log_p += lm;
            log_p += T.beta;
          }
          v = log_p;
          if (!(v > BM_NEG)) v = BM_NEG;
        }
      }
      cand[idx] = v;
      if (v > BM_NEG) {
        cand_aux[idx * 2 + 0] = next_state;
        cand_aux[idx * 2 + 1] = wid;
      }
    }
    lap(11);
    __syncthreads();
    lap(4);
    // ---- phase 3b: a child that is already in the beam absorbs its parent's extension ----
#pragma unroll 1
    for (int k = tid; k < n_active; k += BM_THREADS) {
      const int i = sm.pidx[k];
      float nb = sm.selfnb[k];
      if (i >= 0) {
        const int c = S.ch[k];
        const float lpc = sm.lp[c];
        const bool pass = sm.allowed[c] && !(full_beam && lpc + S.score[i] < min_cutoff);
        if (pass) {
          nb = lse2(nb, cand[i * C + c]);
          cand[i * C + c] = BM_NEG;
          if (S.lpc[k] < lpc) {   // PathTrie::get_path_trie keeps the time step of the best symbol probability
            S.lpc[k] = lpc;
            S.ts[k] = t;
            a_info[S.node[k]] = (c & 0xFF) | (t << 8);
            a_lpc[S.node[k]] = lpc;
          }
        }
      }
      sm.selfnb[k] = nb;
      sm.surv[k] = 0;
      sm.selfscore[k] = lse2(sm.selfb[k], nb);
    }
    __syncthreads();
    // ---- phase 3c: PathTrie::get_path_trie also refreshes (log_prob_c, timestep) of a child that EXISTS in the trie but
    //      is not in the beam (removed, kept for its descendants), whether or not the extension survives this step ----
    //      Such a child of a live prefix comes about in two ways only -- it left the beam with descendants while its
    //      parent was live, or its parent came back into the beam (was revived) -- and both mark the parent (gprobe),
    //      so only the candidates of marked prefixes pay for a hash probe in global memory.
#pragma unroll 1
    for (int idx = tid; sm.n_dead > 0 && idx < n_active * C; idx += BM_THREADS) {
      if (!(cand[idx] > BM_NEG)) continue;
      const int i = divC(idx), c = idx - i * C;
      if (!S.gprobe[i]) continue;
      const int par = S.node[i];
      // (a_ref and h_keys change through atomics, which act on L2: read them past the L1)
      const uint32_t hk = ((uint32_t)par << 8) | (uint32_t)c;
      uint32_t slot = (hk * 0x9E3779B1u) >> p.h_shift;
      for (;;) {
        const uint32_t k = __ldcg(&h_keys[slot]);
        if (k == 0xFFFFFFFFu) break;
        if (k == hk) {
          const int id = __ldcg(&h_vals[slot]);
          if (id < p.max_nodes && __ldcg(&a_ref[id]) > 0 && a_lpc[id] < sm.lp[c]) {
            a_lpc[id] = sm.lp[c];
            a_info[id] = (c & 0xFF) | (t << 8);
          }
          break;
        }
        slot = (slot + 1) & p.h_mask;
      }
    }
    lap(5);
    // ---- phase 4: compaction of the valid candidates + bitonic sort by prefix_compare ----
    {
      // Dense list of the valid candidates without a shared counter: every warp counts its own (pass 1), the 16 warp
      // totals give each warp its first slot, pass 2 writes the keys.  q in [0, BM_MAXW): live prefixes, then the children.
      const unsigned lane = tid & 31, lt = (1u << lane) - 1u;
      const int total = n_active * C;
      auto item_key = [&](int q, uint64_t& key) -> bool {
        if (q < BM_MAXW) {
          if (q < n_active && sm.selfscore[q] > BM_NEG) {
            key = ((uint64_t)f2o_desc(sm.selfscore[q]) << 32) | ((uint64_t)((S.ch[q] + 1) & 0xFF) << 16) | (uint64_t)q;
            return true;
          }
        } else if (q - BM_MAXW < total) {
          const int idx = q - BM_MAXW;
          const float v = cand[idx];
          if (v > BM_NEG) {
            key = ((uint64_t)f2o_desc(v) << 32) | ((uint64_t)((idx - divC(idx) * C + 1) & 0xFF) << 16) | (uint64_t)q;
            return true;
          }
        }
        return false;
      };
      int mine = 0;
#pragma unroll 1
      for (int base = 0; base < BM_MAXW + total; base += BM_THREADS) {
        uint64_t key;
        mine += __popc(__ballot_sync(0xffffffffu, item_key(base + tid, key)));
      }
      if (lane == 0) sm.wsum[tid >> 5] = mine;
      __syncthreads();
      int slot = 0, all = 0;
      for (int w2 = 0; w2 < BM_THREADS / 32; ++w2) {
        const int v = sm.wsum[w2];
        if (w2 < (tid >> 5)) slot += v;
        all += v;
      }
#pragma unroll 1
      for (int base = 0; base < BM_MAXW + total; base += BM_THREADS) {
        uint64_t key = 0;
        const bool valid = item_key(base + tid, key);
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) keys[slot + __popc(m & lt)] = key;
        slot += __popc(m);
      }
      if (tid == 0) sm.count = all;
    }
    __syncthreads();
    lap(6);
    const int count = sm.count;
    if (count <= BM_THREADS / 2) {
      // few candidates (word LM: the dictionary leaves ~200): two threads rank each key by counting the smaller ones
      // in one half of the list each (broadcast 128-bit reads, no barrier inside); the best W land in sorted order
      const int ki = tid >> 1, part = tid & 1;
      const uint64_t key = ki < count ? keys[ki] : ~0ULL;
      const int pairs = (count + 1) >> 1, half = (pairs + 1) >> 1;   // the list as ulonglong2 pairs, split in two
      int rank = 0;
      if (ki < count) {
        const ulonglong2* kp = reinterpret_cast<const ulonglong2*>(keys);
        const int q0 = part * half, q1 = min(pairs, q0 + half);
#pragma unroll 2
        for (int q = q0; q < q1; ++q) {
          const ulonglong2 v = kp[q];
          rank += (v.x < key) + ((2 * q + 1 < count) && v.y < key);
        }
      }
      rank += __shfl_xor_sync(0xffffffffu, rank, 1);
      __syncthreads();
      if (part == 0 && ki < count && rank < W) keys[rank] = key;
      __syncthreads();
    } else if (count <= BM_THREADS) {
      const uint64_t key = tid < count ? keys[tid] : ~0ULL;
      int rank = 0;
      if (tid < count) {
#pragma unroll 4
        for (int q = 0; q < count; ++q) rank += keys[q] < key;
      }
      __syncthreads();
      if (tid < count && rank < W) keys[rank] = key;
      __syncthreads();
    } else {
      int n2 = 64;
      while (n2 < count) n2 <<= 1;
      for (int i = count + tid; i < n2; i += BM_THREADS) keys[i] = ~0ULL;
      __syncthreads();
      for (int k2 = 2; k2 <= n2; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < n2; i += BM_THREADS) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const uint64_t a = keys[i], bb = keys[ixj];
              const bool up = (i & k2) == 0;
              if ((a > bb) == up) {
                keys[i] = bb;
                keys[ixj] = a;
              }
            }
          }
          __syncthreads();
        }
      }
    }
    lap(7);
    // ---- phase 5: the new live beam ----
    const int m = min(W, count);
#pragma unroll 1
    for (int r = tid; r < m; r += BM_THREADS) {
      const uint64_t key = keys[r];
      const int code = (int)(key & 0xFFFF);
      if (code < BM_MAXW) {   // surviving prefix
        const int k = code;
        Nx.node[r] = S.node[k]; Nx.parent[r] = S.parent[k]; Nx.ch[r] = S.ch[k]; Nx.dstate[r] = S.dstate[k];
        Nx.ts[r] = S.ts[k]; Nx.lpc[r] = S.lpc[k];
        for (int h = 0; h < BM_HIST; ++h) Nx.hist[r][h] = S.hist[k][h];
        Nx.bprev[r] = sm.selfb[k]; Nx.nbprev[r] = sm.selfnb[k]; Nx.score[r] = sm.selfscore[k];
        Nx.lmsp[r] = S.lmsp[k]; Nx.lmwid[r] = S.lmwid[k]; Nx.lmns[r] = S.lmns[k]; Nx.lmok[r] = S.lmok[k];
        Nx.rowok[r] = 1; Nx.gprobe[r] = S.gprobe[k];
        sm.pidx[r] = k;   // (pidx is free again here) source slot of the dictionary row, copied below by all threads
        sm.surv[k] = r + 1;
      } else {                // new prefix: parent i extended by symbol c
        const int idx = code - BM_MAXW;
        const int i = divC(idx), c = idx - i * C;
        // the node of (parent, symbol): the one created earlier if this prefix has been in the beam before, else new
        int id = -1;
        bool fresh = false;
        {
          const uint32_t hk = ((uint32_t)S.node[i] << 8) | (uint32_t)c;
          uint32_t slot = (hk * 0x9E3779B1u) >> p.h_shift;
          for (;;) {   // one atomic round trip per probe: claim the slot if it is empty, else see who has it
            const uint32_t old = atomicCAS(&h_keys[slot], 0xFFFFFFFFu, hk);   // (parent, symbol) pairs of a step are distinct
            if (old == 0xFFFFFFFFu) {
              id = atomicAdd(&sm.arena_count, 1);
              h_vals[slot] = id;
              fresh = true;
              break;
            }
            if (old == hk) { id = __ldcg(&h_vals[slot]); break; }   // has been in the trie before
            slot = (slot + 1) & p.h_mask;
          }
        }
        const float v = cand[idx];
        const int wid = cand_aux[idx * 2 + 1];
        // PathTrie lifetime: a removed node that still has descendants is REVIVED with the frame of its best symbol
        // probability (refreshed in phase 3c); a node that was deleted (no descendants left) starts afresh
        const bool revived = !fresh && id < p.max_nodes && __ldcg(&a_ref[id]) > 0;
        if (id < p.max_nodes) {
          if (revived) {
            atomicAdd(&a_ref[id], 1);
            atomicSub(&sm.n_dead, 1);
          } else {
            a_ref[id] = 1;
            a_lpc[id] = sm.lp[c];
            a_info[id] = (c & 0xFF) | (t << 8);
            atomicAdd(&a_ref[S.node[i]], 1);
          }
        }
        Nx.node[r] = id; Nx.parent[r] = S.node[i]; Nx.ch[r] = c;
        Nx.ts[r] = revived ? (a_info[id] >> 8) : t;
        Nx.lpc[r] = revived ? a_lpc[id] : sm.lp[c];
        Nx.dstate[r] = cand_aux[idx * 2 + 0];
        const bool shift = T.has_lm && (T.char_based || c == p.space);
        for (int h = 0; h < BM_HIST; ++h) {
          int hv = S.hist[i][h];
          if (shift && HN > 0) hv = (h + 1 < HN) ? S.hist[i][h + 1] : (h == HN - 1 ? (wid < 0 ? 0 : wid) : hv);
          Nx.hist[r][h] = hv;
        }
        Nx.bprev[r] = BM_NEG; Nx.nbprev[r] = v; Nx.score[r] = v;
        Nx.lmok[r] = 0; Nx.rowok[r] = 0; Nx.gprobe[r] = revived ? 1 : 0;
        sm.pidx[r] = -1;
        if (fresh && id < p.max_nodes) {
          a_parent[id] = S.node[i];
          a_wid[id] = (T.has_lm && !T.char_based && c == p.space) ? (wid < 0 ? 0 : wid) : -1;
        }
      }
    }
    __syncthreads();
    // prefixes that left the beam: PathTrie::remove() -- the node goes away unless it has children, and so do its
    // ancestors that are neither in the beam nor have other children (reference counts: 1 for being live + children)
#pragma unroll 1
    for (int k = tid; k < n_active; k += BM_THREADS) {
      if (sm.surv[k]) continue;
      int n = S.node[k];
      if (n <= 0 || n >= p.max_nodes) continue;
      if (atomicSub(&a_ref[n], 1) != 1) {      // still has children: stays in the trie, outside the beam
        atomicAdd(&sm.n_dead, 1);
        const int ps = sm.pslot[k];             // its parent, if it stays in the beam, has such a child from now on
        if (ps >= 0 && sm.surv[ps]) Nx.gprobe[sm.surv[ps] - 1] = 1;
        continue;
      }
      n = S.parent[k];                          // deleted: release the parent (and its removed ancestors in turn)
      while (n > 0 && n < p.max_nodes) {
        if (atomicSub(&a_ref[n], 1) != 1) break;
        atomicSub(&sm.n_dead, 1);               // a live node keeps its own reference, so this one was a removed node
        n = a_parent[n];
      }
    }
    if (word_lm) {
#pragma unroll 1
      for (int idx = tid; idx < m * C; idx += BM_THREADS) {
        const int r = divC(idx), k = sm.pidx[r];
        if (k >= 0) rowsN[idx] = rowsS[k * C + (idx - r * C)];
      }
      __syncthreads();
    }
    lap(8);
    n_active = m;
    cur ^= 1;
    if (n_active == 0) break;
  }
  if (dbg)
    for (int i = 0; i < 12; ++i) p.dbg[i] = ph[i];

  // ---- final: score the unfinished last word (word LM), order, approximate CTC score, back-trace ----
  BeamState& S = sm.st[cur];
  if (T.has_lm && !T.char_based) {
    for (int k = tid; k < n_active; k += BM_THREADS) {
      if (S.ch[k] >= 0 && S.ch[k] != p.space) {
        int words[BM_HIST + 1];
        for (int h = 0; h < HN; ++h) words[h] = S.hist[k][h];
        const int wid = T.word_at[S.dstate[k]];
        words[HN] = wid < 0 ? 0 : wid;
        float sc = lm_log_cond_prob(T, words, HN + 1) * T.alpha;
        sc += T.beta;
        S.score[k] += sc;
      }
    }
    __syncthreads();
  }
  for (int k = tid; k < n_active; k += BM_THREADS)
    keys[k] = ((uint64_t)f2o_desc(S.score[k]) << 32) | ((uint64_t)((S.ch[k] + 1) & 0xFF) << 16) | (uint64_t)k;
  int n2 = 64;
  while (n2 < n_active) n2 <<= 1;
  for (int i = n_active + tid; i < n2; i += BM_THREADS) keys[i] = ~0ULL;
  __syncthreads();
  for (int k2 = 2; k2 <= n2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += BM_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = keys[i], bb = keys[ixj];
          const bool up = (i & k2) == 0;
          if ((a > bb) == up) { keys[i] = bb; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < W; r += BM_THREADS) {
    int32_t* otok = p.out_tokens + ((size_t)b * W + r) * p.T;
    int32_t* ots = p.out_ts + ((size_t)b * W + r) * p.T;
    if (r >= n_active) {
      p.out_scores[(size_t)b * W + r] = 0.f;
      p.out_lens[(size_t)b * W + r] = 0;
      continue;
    }
    const int k = (int)(keys[r] & 0xFFFF);
    // path length
    int n = 0;
    for (int node = S.node[k]; node > 0; node = a_parent[node]) ++n;
    n = min(n, p.T);
    int pos = n - 1;
    int32_t* wbuf = p.words + ((size_t)b * W + r) * (p.T + 2);
    int nwords = 0;   // collected backwards
    for (int node = S.node[k]; node > 0 && pos >= 0; node = a_parent[node], --pos) {
      const int info = a_info[node];
      const int c = info & 0xFF;
      otok[pos] = c;
      ots[pos] = info >> 8;
      if (T.has_lm) {
        if (T.char_based) wbuf[nwords++] = T.char_word[c];
        else if (c == p.space && a_wid[node] >= 0) wbuf[nwords++] = a_wid[node];
      }
    }
    double approx = (double)S.score[k];
    if (T.has_lm) {
      // words are stored newest-first in wbuf[0..nwords); a trailing partial word comes first
      int total = nwords;
      int trailing = -1;
      if (!T.char_based && S.ch[k] >= 0 && S.ch[k] != p.space) {
        const int wid = T.word_at[S.dstate[k]];
        trailing = wid < 0 ? 0 : wid;
        total += 1;
      }
      auto word_at_pos = [&](int i) -> int {   // i-th word of the sentence, oldest first
        if (trailing >= 0 && i == total - 1) return trailing;
        const int from_end = (trailing >= 0) ? (total - 2 - i) : (total - 1 - i);
        return wbuf[from_end];
      };
      const int order = T.order;
      const int pad = total == 0 ? order : order - 1;
      const int slen = pad + total + 1;
      double sent = 0.0;
      for (int i = 0; i + order <= slen; ++i) {
        int win[BM_HIST + 1];
        for (int j = 0; j < order; ++j) {
          const int q = i + j;
          win[j] = q < pad ? T.id_bos : (q < pad + total ? word_at_pos(q - pad) : T.id_eos);
        }
        sent += (double)lm_log_cond_prob(T, win, order);
      }
      approx = approx - (double)n * (double)T.beta - sent * (double)T.alpha;
    }
    p.out_scores[(size_t)b * W + r] = (float)(-(double)(float)approx);
    p.out_lens[(size_t)b * W + r] = n;
  }
}

// ------------------------------------------------------------------------------------------ host
static std::vector<std::string> utf8_split(const std::string& s) {
  std::vector<std::string> out;
  for (size_t i = 0; i < s.size();) {
    unsigned char c = (unsigned char)s[i];
    size_t n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
    out.push_back(s.substr(i, n));
    i += n;
  }
  return out;
}

}  // namespace dsb

struct dsb_beam {
  int C = 0, W = 0, blank = 0, space = -1, cutoff_top_n = 40;
  float cutoff_prob = 1.f;
  dsb::BeamTables tab{};
  int64_t n_ngrams = 0;
  std::vector<void*> owned;
};

using namespace dsb;

template <typename T>
static int beam_upload(dsb_beam* d, const std::vector<T>& h, const T** out) {
  void* q = nullptr;
  DSB_CUDA(cudaMalloc(&q, sizeof(T) * (h.empty() ? 1 : h.size())));
  d->owned.push_back(q);
  if (!h.empty()) DSB_CUDA(cudaMemcpy(q, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(q);
  return 0;
}

extern "C" int dsb_beam_create(const char* labels_utf8, int n_labels, const char* lm_path, float alpha, float beta,
                               int cutoff_top_n, float cutoff_prob, int beam_width, int blank_id,
                               int log_probs_input, dsb_beam** out) {
  DSB_REQUIRE(labels_utf8 && out && n_labels >= 2 && n_labels <= BM_MAXC, "dsb_beam_create: bad labels (n=%d)", n_labels);
  DSB_REQUIRE(beam_width >= 1 && beam_width <= BM_MAXW, "dsb_beam_create: beam_width %d not in [1,%d]", beam_width,
              BM_MAXW);
  DSB_REQUIRE(blank_id >= 0 && blank_id < n_labels, "dsb_beam_create: blank_id out of range");
  if (log_probs_input) return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: log_probs_input is not supported");
  std::vector<std::string> labels;
  const char* p = labels_utf8;
  for (int i = 0; i < n_labels; ++i) {
    labels.emplace_back(p);
    p += labels.back().size() + 1;
  }
  dsb_beam* d = new dsb_beam();
  d->C = n_labels;
  d->W = beam_width;
  d->blank = blank_id;
  d->cutoff_top_n = cutoff_top_n;
  d->cutoff_prob = cutoff_prob;
  for (int i = 0; i < n_labels; ++i)
    if (labels[i] == " ") d->space = i;
  BeamTables& T = d->tab;
  T.alpha = alpha;
  T.beta = beta;
  T.order = 1;
  T.unk_prob = -100.f;
  auto fail = [&](int e) {
    for (void* q : d->owned) cudaFree(q);
    delete d;
    return e;
  };
  if (lm_path && lm_path[0]) {
    HostLm lm;
    const int klm = is_kenlm_binary(lm_path);
    if (klm < 0) return fail(klm);
    if (int e = klm ? load_klm(lm_path, lm) : load_arpa(lm_path, lm)) return fail(e);
    if (lm.order > BM_HIST + 1) return fail(set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_create: LM order %d > %d", lm.order, BM_HIST + 1));
    T.has_lm = 1;
    T.order = lm.order;
    T.unk_prob = lm.unk_prob;
    T.id_bos = lm.index("<s>");
    T.id_eos = lm.index("</s>");
    T.char_based = 1;
    for (const std::string& w : lm.words)
      if (w != "<unk>" && w != "<s>" && w != "</s>" && utf8_split(w).size() > 1) T.char_based = 0;
    // n-gram hash table
    size_t cap = 64;
    while (cap < lm.grams.size() * 2) cap <<= 1;
    std::vector<LmEntry> table(cap);
    memset(table.data(), 0, sizeof(LmEntry) * cap);
    for (const HostLm::Gram& g : lm.grams) {
      const uint64_t k0 = g.key, k1 = (uint64_t)g.n;
      uint32_t h = key_hash(k0, k1) & (uint32_t)(cap - 1);
      while (table[h].used && !(table[h].k0 == k0 && table[h].k1 == k1)) h = (h + 1) & (uint32_t)(cap - 1);
      table[h].k0 = k0; table[h].k1 = k1; table[h].prob = g.prob; table[h].backoff = g.backoff; table[h].used = 1;
    }
    d->n_ngrams = (int64_t)lm.grams.size();
    T.lm_mask = (uint32_t)(cap - 1);
    if (int e = beam_upload(d, table, &T.lm)) return fail(e);
    std::unordered_map<std::string, int> char_map;
    for (int i = 0; i < n_labels; ++i) char_map[labels[i]] = i;
    std::vector<int32_t> char_word(n_labels, 0);
    for (int i = 0; i < n_labels; ++i) char_word[i] = lm.index(labels[i]);
    if (int e = beam_upload(d, char_word, &T.char_word)) return fail(e);
    // dictionary trie over (word + ' '), final states folded back to the root
    std::vector<int32_t> trans(n_labels, -1), word_at(1, -1);
    if (!T.char_based) {
      if (d->space < 0) return fail(set_error(DSB_ERR_INVALID, "dsb_beam_create: word LM needs a space label"));
      for (size_t wi = 0; wi < lm.words.size(); ++wi) {
        const std::string& w = lm.words[wi];
        std::vector<int> ids;
        bool ok = true;
        for (const std::string& c : utf8_split(w)) {
          auto it = char_map.find(c);
          if (it == char_map.end() || it->second == d->space) { ok = false; break; }
          ids.push_back(it->second);
        }
        if (!ok || ids.empty()) continue;
        int s = 0;
        for (int c : ids) {
          int nx = trans[(size_t)s * n_labels + c];
          if (nx < 0) {
            nx = (int)word_at.size();
            word_at.push_back(-1);
            trans.resize(trans.size() + n_labels, -1);
            trans[(size_t)s * n_labels + c] = nx;
          }
          s = nx;
        }
        word_at[s] = (int)wi;
        trans[(size_t)s * n_labels + d->space] = 0;   // word + ' ' is final -> back to the start state
      }
    }
    if (int e = beam_upload(d, trans, &T.trans)) return fail(e);
    if (int e = beam_upload(d, word_at, &T.word_at)) return fail(e);
  }
  *out = d;
  return 0;
}

extern "C" void dsb_beam_destroy(dsb_beam* d) {
  if (!d) return;
  for (void* q : d->owned) cudaFree(q);
  delete d;
}

extern "C" int dsb_beam_lm_order(const dsb_beam* d) { return d && d->tab.has_lm ? d->tab.order : 0; }
extern "C" int dsb_beam_lm_is_char_based(const dsb_beam* d) { return d && d->tab.has_lm ? d->tab.char_based : -1; }
extern "C" int64_t dsb_beam_lm_num_ngrams(const dsb_beam* d) { return d ? d->n_ngrams : 0; }

namespace {
struct BeamWs {
  size_t o_len, o_parent, o_info, o_wid, o_lpc, o_ref, o_words, o_hkeys, o_hvals, total;
  int max_nodes, h_cap, h_log2;
};
BeamWs beam_ws(const dsb_beam* d, int B, int T) {
  BeamWs w{};
  w.max_nodes = 1 + d->W * T;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += dsb::align_up(bytes, 256);
    return o;
  };
  w.o_len = take(sizeof(int32_t) * B);
  w.o_parent = take(sizeof(int32_t) * (size_t)B * w.max_nodes);
  w.o_info = take(sizeof(int32_t) * (size_t)B * w.max_nodes);
  w.o_wid = take(sizeof(int32_t) * (size_t)B * w.max_nodes);
  w.o_lpc = take(sizeof(float) * (size_t)B * w.max_nodes);
  w.o_ref = take(sizeof(int32_t) * (size_t)B * w.max_nodes);
  w.o_words = take(sizeof(int32_t) * (size_t)B * d->W * (T + 2));
  w.h_log2 = 10;
  while ((1 << w.h_log2) < 2 * w.max_nodes) ++w.h_log2;
  w.h_cap = 1 << w.h_log2;
  w.o_hkeys = take(sizeof(uint32_t) * (size_t)B * w.h_cap);
  w.o_hvals = take(sizeof(int32_t) * (size_t)B * w.h_cap);
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t dsb_beam_workspace_bytes(const dsb_beam* d, int B, int T) {
  if (!d || B <= 0 || T <= 0) return 0;
  return beam_ws(d, B, T).total;
}

extern "C" int dsb_beam_decode(dsb_beam* d, const float* probs, const int32_t* seq_lens, int B, int T, int C,
                               int32_t* out_tokens, int32_t* out_timesteps, float* out_scores, int32_t* out_lens,
                               void* workspace, size_t workspace_bytes, void* stream) {
  DSB_REQUIRE(d && probs && seq_lens && out_tokens && out_timesteps && out_scores && out_lens && workspace,
              "dsb_beam_decode: null argument");
  DSB_REQUIRE(B > 0 && T > 0 && C == d->C, "dsb_beam_decode: bad shape B=%d T=%d C=%d (decoder has %d labels)", B, T, C,
              d->C);
  const BeamWs w = beam_ws(d, B, T);
  if (workspace_bytes < w.total)
    return set_error(DSB_ERR_WORKSPACE, "dsb_beam_decode: workspace %zu < required %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope scope(ST_BEAM, st);
  char* base = reinterpret_cast<char*>(workspace);
  int32_t* d_len = reinterpret_cast<int32_t*>(base + w.o_len);
  DSB_CUDA(cudaMemcpyAsync(d_len, seq_lens, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  BeamParams p{};
  p.probs = probs;
  p.seq_lens = d_len;
  p.B = B; p.T = T; p.C = C; p.W = d->W; p.blank = d->blank; p.space = d->space;
  p.cutoff_top_n = d->cutoff_top_n;
  p.cutoff_prob = d->cutoff_prob;
  p.a_parent = reinterpret_cast<int32_t*>(base + w.o_parent);
  p.a_info = reinterpret_cast<int32_t*>(base + w.o_info);
  p.a_wid = reinterpret_cast<int32_t*>(base + w.o_wid);
  p.a_lpc = reinterpret_cast<float*>(base + w.o_lpc);
  p.a_ref = reinterpret_cast<int32_t*>(base + w.o_ref);
  p.max_nodes = w.max_nodes;
  p.h_keys = reinterpret_cast<uint32_t*>(base + w.o_hkeys);
  p.h_vals = reinterpret_cast<int32_t*>(base + w.o_hvals);
  p.h_mask = (uint32_t)w.h_cap - 1u;
  p.h_shift = 32 - w.h_log2;
  DSB_CUDA(cudaMemsetAsync(p.h_keys, 0xFF, sizeof(uint32_t) * (size_t)B * w.h_cap, st));
  p.words = reinterpret_cast<int32_t*>(base + w.o_words);
  p.out_tokens = out_tokens;
  p.out_ts = out_timesteps;
  p.out_scores = out_scores;
  p.out_lens = out_lens;
  p.tab = d->tab;
  static const bool debug = getenv("DSB_BEAM_DEBUG") != nullptr;
  long long* dbg = nullptr;
  if (debug) {
    DSB_CUDA(cudaMalloc(&dbg, 12 * sizeof(long long)));
    DSB_CUDA(cudaMemsetAsync(dbg, 0, 12 * sizeof(long long), st));
  }
  p.dbg = dbg;
  int n2 = BM_THREADS;   // capacity of the key array: also holds the per-warp work lists of the candidate pass
  while (n2 < BM_MAXW + d->W * C + BM_THREADS) n2 <<= 1;
  p.n2 = n2;
  const size_t smem = ((sizeof(BeamSmem) + 15) & ~(size_t)15) + sizeof(uint64_t) * n2 +
                      (size_t)d->W * C * (sizeof(float) + 2 * sizeof(int32_t)) +
                      ((d->tab.has_lm && !d->tab.char_based) ? 2 * (size_t)d->W * C * sizeof(int32_t) : 0);
  if (smem > 227 * 1024)
    return set_error(DSB_ERR_UNSUPPORTED, "dsb_beam_decode: beam width %d x %d classes needs %zu bytes of shared memory", d->W, C, smem);
  DSB_CUDA(cudaFuncSetAttribute(beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DSB_CUDA(cudaMemsetAsync(out_tokens, 0, sizeof(int32_t) * (size_t)B * d->W * T, st));
  DSB_CUDA(cudaMemsetAsync(out_timesteps, 0, sizeof(int32_t) * (size_t)B * d->W * T, st));
  beam_kernel<<<B, BM_THREADS, smem, st>>>(p);
  DSB_CHECK_LAUNCH();
  if (debug) {
    long long h[12];
    DSB_CUDA(cudaStreamSynchronize(st));
    DSB_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    const double n = seq_lens[0] > 0 ? seq_lens[0] : 1;
    fprintf(stderr, "[beam debug] utterance 0, %d steps, cycles/step: prep %.0f | pairs %.0f | rows %.0f | lm %.0f | candidates %.0f | "
            "merge %.0f | compact %.0f | select %.0f | new beam %.0f | work items of warp 0 %.0f | candidates = pass A %.0f + pass B %.0f + barrier\n",
            (int)n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n);
  }
  return 0;
}
