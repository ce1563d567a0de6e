// GPU token-passing Viterbi beam search: one thread block per utterance walks the
// frames of its utterance with the log-likelihood rows already resident in HBM, so neither the
// [frames x pdfs] matrix nor one host core per utterance is needed to get from PCM to words.
//
// Reference: Decoder::Decode / InitDecoding / ProcessEmitting / ProcessNonemitting / BestPath
// (src/decoder.cc:39-339), Fst::IterateArcs (src/fst.cc:94-129), pk_decodable_loglikelihood
// (src/decodable.cc:24-31). What is kept exactly:
//   * arc cost arithmetic: total = double(token cost) + double(arc weight) + double(acoustic cost),
//     stored back as float (src/decoder.cc:276-291, Token::cost_ is a float);
//   * the beam: weight_cutoff = float(best + beam) on the previous frame's tokens (:141-200 with
//     fewer than kBeamSize = 30000 tokens), next cutoff = min over kept arcs of total + beam,
//     ProcessNonemitting(float(next cutoff)) (:203-237);
//   * BestPath: min over tokens of cost + final(state), weight = that + final(state) once more
//     (the reference adds the final weight twice, :300-339).
// What differs: the reference tightens the next cutoff while it walks the token list, so arcs seen
// early are tested against a looser bound and may create tokens that the final bound would not;
// those tokens lie outside the next frame's beam and are dropped there. Here every arc is tested
// against the final (tightest) bound. The max-active estimate by sampling (:141-200, more than
// 30000 tokens) is replaced by a hard per-utterance capacity that fails the utterance. Ties between
// equal float costs are broken by arc index instead of by list order.

#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "decoder.cuh"

namespace pkb {

namespace {

constexpr int kVitThreads = 256;
// kVitGroup lanes share one token's arcs: 8 for graphs with a few arcs per state, a whole warp for
// dense ones (the launcher picks by the average out-degree)
constexpr unsigned long long kEmptyVal = ~0ull;
constexpr uint32_t kNoArc = 0xffffffffu;

// total order on floats / doubles as unsigned integers
__device__ __forceinline__ uint32_t ord32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ unsigned long long ord64(double d) {
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(d));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long o) {
  return __longlong_as_double(static_cast<long long>((o >> 63) ? (o & 0x7fffffffffffffffull) : ~o));
}
__device__ __forceinline__ unsigned long long pack(float cost, uint32_t arc) {
  return (static_cast<unsigned long long>(ord32(cost)) << 32) | arc;
}
__device__ __forceinline__ float cost_of(unsigned long long v) { return unord32(static_cast<uint32_t>(v >> 32)); }

struct FstDev {
  int num_states, start, has_eps;
  const float *final_w;
  const int32_t *arc_begin, *arc_src, *arc_dst, *arc_il, *arc_ol;
  const float *arc_w;
};

// One frame's tokens: open-addressing table state -> (cost, winning arc) plus the list of used slots.
struct Tab {
  int *keys;                 // state + 1, 0 = empty
  unsigned long long *vals;  // pack(cost, arc); kEmptyVal when unused
  int *bp;                   // index of the token's newest word record, -1 none, -2 not resolved yet
  int *list;                 // slots in use
};

struct Work {
  Tab tab[2];
  int *frontier[2];
  int *inq;        // [H] slot is queued for the epsilon closure
  int *log_prev;   // word records: previous record, output label
  int *log_ol;
};

__device__ __forceinline__ uint32_t hash_state(int s) { return static_cast<uint32_t>(s) * 2654435761u; }

// slot of `state`, or -1
__device__ __forceinline__ int tab_find(const Tab &t, int state, uint32_t mask) {
  uint32_t s = hash_state(state) & mask;
  for (uint32_t n = 0; n <= mask; ++n) {
    const int k = t.keys[s];
    if (k == state + 1) return static_cast<int>(s);
    if (k == 0) return -1;
    s = (s + 1) & mask;
  }
  return -1;
}

// slot of `state`, claiming an empty one (and appending it to the list) when it is new; -1 on overflow
__device__ __forceinline__ int tab_insert(const Tab &t, int state, uint32_t mask, int *n_tok, int max_tok) {
  uint32_t s = hash_state(state) & mask;
  for (uint32_t n = 0; n <= mask; ++n) {  // bounded: a full table (capacity overflow) must not spin
    int k = *reinterpret_cast<volatile int *>(&t.keys[s]);
    if (k == 0) {
      k = atomicCAS(&t.keys[s], 0, state + 1);
      if (k == 0) {
        const int i = atomicAdd(n_tok, 1);
        if (i >= max_tok) return -1;
        t.list[i] = static_cast<int>(s);
        t.bp[s] = -2;
        return static_cast<int>(s);
      }
    }
    if (k == state + 1) return static_cast<int>(s);
    s = (s + 1) & mask;
  }
  return -1;
}

template <int kVitGroup>
__global__ void __launch_bounds__(kVitThreads)
viterbi_kernel(FstDev fst, float beam, int max_tok, uint32_t mask, int max_log, int max_words,
               const float *__restrict__ loglik, int num_pdfs, const int64_t *__restrict__ row_off,
               const int32_t *__restrict__ num_frames, int n_utts, const int32_t *__restrict__ tid2pdf,
               char *work_base, size_t work_stride, int tables_in_smem, int32_t *__restrict__ words_out,
               int32_t *__restrict__ n_words_out, float *__restrict__ weight_out) {
  __shared__ int s_ntok[2], s_nfront[2], s_log, s_err, s_unres;
  __shared__ unsigned long long s_min;

  // carve this block's workspace. Small graphs (every state fits the token capacity chosen by the
  // launcher) keep the two token tables, the lists and the queue in shared memory: the per-frame
  // chain of dependent look-ups and atomics then never leaves the SM. The word records stay global.
  extern __shared__ __align__(16) char s_tab[];
  const uint32_t H = mask + 1;
  char *gp = work_base + static_cast<size_t>(blockIdx.x) * work_stride;
  char *wp = tables_in_smem ? s_tab : gp;
  Work w;
  for (int i = 0; i < 2; ++i) {
    w.tab[i].vals = reinterpret_cast<unsigned long long *>(wp); wp += sizeof(unsigned long long) * H;
  }
  for (int i = 0; i < 2; ++i) {
    w.tab[i].keys = reinterpret_cast<int *>(wp); wp += sizeof(int) * H;
    w.tab[i].bp = reinterpret_cast<int *>(wp); wp += sizeof(int) * H;
    w.tab[i].list = reinterpret_cast<int *>(wp); wp += sizeof(int) * max_tok;
    w.frontier[i] = reinterpret_cast<int *>(wp); wp += sizeof(int) * max_tok;
  }
  w.inq = reinterpret_cast<int *>(wp); wp += sizeof(int) * H;
  if (tables_in_smem) wp = gp;  // the global workspace then holds the word records only
  w.log_prev = reinterpret_cast<int *>(wp); wp += sizeof(int) * max_log;
  w.log_ol = reinterpret_cast<int *>(wp);

  const int tid = threadIdx.x;
  if (tables_in_smem) {
    // the invariant the global workspace gets from the launcher's memset
    for (uint32_t i = tid; i < H; i += kVitThreads) {
      for (int k = 0; k < 2; ++k) {
        w.tab[k].keys[i] = 0;
        w.tab[k].vals[i] = kEmptyVal;
      }
      w.inq[i] = 0;
    }
    __syncthreads();
  }
  constexpr int kVitGroups = kVitThreads / kVitGroup;
  const int grp = tid / kVitGroup, gl = tid % kVitGroup;

  // Epsilon closure of table `c` (ProcessNonemitting, src/decoder.cc:203-237) under `cutoff`, then
  // the word back-pointers of every token of the frame. `p` is the previous frame's table.
  auto close_and_resolve = [&](int c, int p, double cutoff) {
    const Tab &tc = w.tab[c];
    if (fst.has_eps) {
      // frontier = every token of the frame
      const int n0 = s_ntok[c];
      for (int i = tid; i < n0; i += kVitThreads) w.frontier[0][i] = tc.list[i];
      if (tid == 0) { s_nfront[0] = n0 < max_tok ? n0 : max_tok; s_nfront[1] = 0; }
      __syncthreads();
      int fb = 0;
      for (;;) {
        const int nf = s_nfront[fb];
        if (nf == 0) break;
        // the queue flags of this round's states are cleared before any of them is expanded, so an
        // improvement that lands while a state is being expanded always re-queues it
        for (int i = tid; i < nf; i += kVitThreads) w.inq[w.frontier[fb][i]] = 0;
        __syncthreads();
        for (int i = tid; i < nf; i += kVitThreads) {
          const int slot = w.frontier[fb][i];
          const int state = tc.keys[slot] - 1;
          const float cost = cost_of(*reinterpret_cast<volatile unsigned long long *>(&tc.vals[slot]));
          for (int a = __ldg(&fst.arc_begin[state]); a < __ldg(&fst.arc_begin[state + 1]); ++a) {
            if (__ldg(&fst.arc_il[a]) != 0) continue;
            const double total = static_cast<double>(cost) + static_cast<double>(__ldg(&fst.arc_w[a]));
            if (total > cutoff) continue;
            const int s2 = tab_insert(tc, __ldg(&fst.arc_dst[a]), mask, &s_ntok[c], max_tok);
            if (s2 < 0) { s_err = 1; continue; }
            // InsertTok replaces only on a strictly lower cost (src/decoder.cc:127-134): CAS loop
            const unsigned long long nv = pack(static_cast<float>(total), static_cast<uint32_t>(a));
            unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&tc.vals[s2]);
            bool won = false;
            while ((nv >> 32) < (old >> 32)) {
              const unsigned long long seen = atomicCAS(&tc.vals[s2], old, nv);
              if (seen == old) { won = true; break; }
              old = seen;
            }
            if (won && atomicExch(&w.inq[s2], 1) == 0) {
              const int j = atomicAdd(&s_nfront[fb ^ 1], 1);
              if (j < max_tok) w.frontier[fb ^ 1][j] = s2; else s_err = 1;
            }
          }
        }
        __syncthreads();
        if (tid == 0) { s_nfront[fb] = 0; if (s_nfront[fb ^ 1] > max_tok) s_nfront[fb ^ 1] = max_tok; }
        fb ^= 1;
        __syncthreads();
        if (s_err) break;
      }
    }
    // word back-pointers (Decoder::InsertTok, src/decoder.cc:107-117): a token inherits the record of
    // the token its winning arc left from, plus one new record when that arc carries an output label
    for (int pass = 0; pass < 256; ++pass) {
      if (tid == 0) s_unres = 0;
      __syncthreads();
      const int n = min(s_ntok[c], max_tok);
      for (int i = tid; i < n; i += kVitThreads) {
        const int slot = tc.list[i];
        if (tc.bp[slot] != -2) continue;
        const uint32_t a = static_cast<uint32_t>(tc.vals[slot]);
        int pb;
        if (a == kNoArc) {
          pb = -1;  // the start token
        } else {
          const int src = __ldg(&fst.arc_src[a]);
          if (__ldg(&fst.arc_il[a]) != 0) {
            const int ps = tab_find(w.tab[p], src, mask);
            pb = ps >= 0 ? w.tab[p].bp[ps] : -1;
          } else {
            const int cs = tab_find(tc, src, mask);
            pb = cs >= 0 ? *reinterpret_cast<volatile int *>(&tc.bp[cs]) : -1;
            if (pb == -2) { atomicAdd(&s_unres, 1); continue; }
          }
          const int ol = __ldg(&fst.arc_ol[a]);
          if (ol != 0) {
            const int r = atomicAdd(&s_log, 1);
            if (r >= max_log) { s_err = 2; pb = -1; }
            else { w.log_prev[r] = pb; w.log_ol[r] = ol; pb = r; }
          }
        }
        tc.bp[slot] = pb;
      }
      __syncthreads();
      const int unres = s_unres;  // read by everyone before thread 0 may reset it in the next pass
      __syncthreads();
      if (unres == 0) break;
      if (pass == 255 && tid == 0) s_err = 3;
    }
    __syncthreads();
  };

  auto clear_tab = [&](int c) {
    const Tab &t = w.tab[c];
    const int n = min(s_ntok[c], max_tok);
    for (int i = tid; i < n; i += kVitThreads) {
      const int slot = t.list[i];
      t.keys[slot] = 0;
      t.vals[slot] = kEmptyVal;
      w.inq[slot] = 0;
    }
    __syncthreads();
    if (tid == 0) s_ntok[c] = 0;
    __syncthreads();
  };

  for (int u = blockIdx.x; u < n_utts; u += gridDim.x) {
    const int T = num_frames[u];
    const float *ll0 = loglik + row_off[u] * num_pdfs;
    if (tid == 0) { s_ntok[0] = s_ntok[1] = 0; s_log = 0; s_err = 0; }
    __syncthreads();
    // ---- InitDecoding (src/decoder.cc:82-101): the start state at cost 0, then its epsilon closure
    int cur = 0;
    if (tid == 0) {
      const int s = tab_insert(w.tab[0], fst.start, mask, &s_ntok[0], max_tok);
      w.tab[0].vals[s] = pack(0.0f, kNoArc);
    }
    __syncthreads();
    close_and_resolve(0, 1, INFINITY);

    bool alive = true;
    for (int f = 0; f < T && alive && !s_err; ++f) {
      const int prev = cur;
      cur ^= 1;
      const Tab &tp = w.tab[prev], &tc = w.tab[cur];
      const int n_prev = s_ntok[prev];
      const float *ll = ll0 + static_cast<int64_t>(f) * num_pdfs;
      // ---- GetCutoff (src/decoder.cc:141-200) below kBeamSize tokens: best cost + beam
      if (tid == 0) s_min = ~0ull;
      __syncthreads();
      {
        uint32_t m = 0xffffffffu;
        for (int i = tid; i < n_prev; i += kVitThreads) m = min(m, static_cast<uint32_t>(tp.vals[tp.list[i]] >> 32));
        atomicMin(&s_min, static_cast<unsigned long long>(m));
      }
      __syncthreads();
      const float best = unord32(static_cast<uint32_t>(s_min));
      const float weight_cutoff = static_cast<float>(static_cast<double>(best) + static_cast<double>(beam));
      __syncthreads();
      // ---- ProcessEmitting, pass 1: the bound on the next frame (min over kept arcs of total + beam)
      if (tid == 0) s_min = ~0ull;
      __syncthreads();
      {
        // eight lanes per token, lanes over its arcs: 32 tokens of the block are in flight at once
        unsigned long long m = ~0ull;
        for (int i = grp; i < n_prev; i += kVitGroups) {
          const int slot = tp.list[i];
          const float cost = cost_of(tp.vals[slot]);
          if (cost > weight_cutoff) continue;
          const int state = tp.keys[slot] - 1;
          const int a1 = __ldg(&fst.arc_begin[state + 1]);
          for (int a = __ldg(&fst.arc_begin[state]) + gl; a < a1; a += kVitGroup) {
            const int il = __ldg(&fst.arc_il[a]);
            if (il == 0) continue;
            const int pdf = __ldg(&tid2pdf[il]);
            const float ac = -__ldg(&ll[pdf]);
            const double total = static_cast<double>(cost) + static_cast<double>(__ldg(&fst.arc_w[a])) +
                                 static_cast<double>(ac);
            m = min(m, ord64(total));
          }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((tid & 31) == 0) atomicMin(&s_min, m);
      }
      __syncthreads();
      if (s_min == ~0ull) { alive = false; break; }  // no arc left the beam: the reference has no tokens either
      const double next_cutoff = unord64(s_min) + static_cast<double>(beam);
      // ---- pass 2: create / improve the tokens of this frame
      for (int i = grp; i < n_prev; i += kVitGroups) {
        const int slot = tp.list[i];
        const float cost = cost_of(tp.vals[slot]);
        if (cost > weight_cutoff) continue;
        const int state = tp.keys[slot] - 1;
        const int a1 = __ldg(&fst.arc_begin[state + 1]);
        for (int a = __ldg(&fst.arc_begin[state]) + gl; a < a1; a += kVitGroup) {
          const int il = __ldg(&fst.arc_il[a]);
          if (il == 0) continue;
          const int pdf = __ldg(&tid2pdf[il]);
          const float ac = -__ldg(&ll[pdf]);
          const double total = static_cast<double>(cost) + static_cast<double>(__ldg(&fst.arc_w[a])) +
                               static_cast<double>(ac);
          if (total > next_cutoff) continue;
          const int s2 = tab_insert(tc, __ldg(&fst.arc_dst[a]), mask, &s_ntok[cur], max_tok);
          if (s2 < 0) { s_err = 1; continue; }
          atomicMin(&tc.vals[s2], pack(static_cast<float>(total), static_cast<uint32_t>(a)));
        }
      }
      __syncthreads();
      // ProcessEmitting returns the bound as a float (src/decoder.cc:240)
      if (!s_err) close_and_resolve(cur, prev, static_cast<double>(static_cast<float>(next_cutoff)));
      clear_tab(prev);
    }

    // ---- BestPath (src/decoder.cc:300-339)
    if (tid == 0) s_min = ~0ull;
    __syncthreads();
    const Tab &tb = w.tab[cur];
    const int n = (alive && !s_err) ? min(s_ntok[cur], max_tok) : 0;
    for (int i = tid; i < n; i += kVitThreads) {
      const int slot = tb.list[i];
      const float c = cost_of(tb.vals[slot]) + fst.final_w[tb.keys[slot] - 1];
      if (c != INFINITY) atomicMin(&s_min, pack(c, static_cast<uint32_t>(slot)));
    }
    __syncthreads();
    if (tid == 0) {
      int32_t *wo = words_out + static_cast<size_t>(u) * max_words;
      if (s_err) {
        n_words_out[u] = -s_err;
        weight_out[u] = 0.0f;
      } else if (s_min == ~0ull) {
        n_words_out[u] = 0;  // no token in a final state: Hypothesis({}, 0)
        weight_out[u] = 0.0f;
      } else {
        const int slot = static_cast<int>(static_cast<uint32_t>(s_min));
        float weight = cost_of(s_min);
        weight += fst.final_w[tb.keys[slot] - 1];
        weight_out[u] = weight;
        int nw = 0;
        for (int r = tb.bp[slot]; r >= 0; r = w.log_prev[r]) ++nw;
        n_words_out[u] = nw;
        int k = nw;
        for (int r = tb.bp[slot]; r >= 0; r = w.log_prev[r]) {
          --k;
          if (k < max_words) wo[k] = w.log_ol[r];
        }
      }
    }
    __syncthreads();
    if (s_err) {
      // a capacity overflow leaves claimed slots that are in no list: wipe both tables
      for (uint32_t i = tid; i < H; i += kVitThreads) {
        for (int k = 0; k < 2; ++k) {
          w.tab[k].keys[i] = 0;
          w.tab[k].vals[i] = kEmptyVal;
        }
        w.inq[i] = 0;
      }
      __syncthreads();
      if (tid == 0) s_ntok[0] = s_ntok[1] = 0;
      __syncthreads();
    } else {
      clear_tab(0);
      clear_tab(1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small graphs (pocketkaldi's own territory: command / small-vocabulary grammars): every state has a
// slot, so the token table is three dense shared-memory arrays per frame and the search runs
// arc-parallel over the whole FST -- all lanes busy, 32-bit native shared-memory atomics, each
// thread keeps the totals of its arcs in registers across the passes of a frame.
//   cost[s]  ord32(token cost), kInactive when the state holds no token
//   warc[s]  winning arc; epsilon arcs carry bit 31 so that an emitting arc of equal cost wins, as
//            it does in the reference where emitting arcs are inserted first (src/decoder.cc:127-134)
//   bp[s]    newest word record of the token
constexpr uint32_t kInactive = 0xffffffffu;

template <int kArcsPerThread>
__global__ void __launch_bounds__(kVitThreads, kArcsPerThread <= 8 ? 3 : 1)
viterbi_dense_kernel(FstDev fst, int num_arcs, float beam, int max_log, int max_words,
                     const float *__restrict__ loglik, int num_pdfs, const int64_t *__restrict__ row_off,
                     const int32_t *__restrict__ num_frames, int n_utts, const int32_t *__restrict__ tid2pdf,
                     char *work_base, size_t work_stride, int32_t *__restrict__ words_out,
                     int32_t *__restrict__ n_words_out, float *__restrict__ weight_out) {
  extern __shared__ __align__(16) char s_dense[];
  __shared__ int s_log, s_err, s_unres, s_changed;
  __shared__ unsigned long long s_min;
  __shared__ unsigned int s_best;
  const int S = fst.num_states;
  uint32_t *cost[2], *warc[2];
  int *bp[2];
  {
    uint32_t *p = reinterpret_cast<uint32_t *>(s_dense);
    cost[0] = p; cost[1] = p + S; warc[0] = p + 2 * S; warc[1] = p + 3 * S;
    bp[0] = reinterpret_cast<int *>(p + 4 * S); bp[1] = reinterpret_cast<int *>(p + 5 * S);
  }
  int *log_prev = reinterpret_cast<int *>(work_base + static_cast<size_t>(blockIdx.x) * work_stride);
  int *log_ol = log_prev + max_log;
  const int tid = threadIdx.x;
  // acoustic costs of the frame, one slot per DISTINCT pdf the graph's arcs read (a word-loop
  // graph reads each pdf from ~15 arcs): bit set of used pdfs, its per-word prefix counts, the pdf
  // of every slot and two frames of costs
  const int W = (num_pdfs + 31) >> 5;
  const int Dmax = min(num_arcs, num_pdfs);
  uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_dense) + 6 * S;
  int *s_wpre = reinterpret_cast<int *>(s_bits + W);
  int *s_pdf = s_wpre + W;
  float *s_ll = reinterpret_cast<float *>(s_pdf + Dmax);  // [2][Dmax]
  __shared__ int s_D;

  // this thread's arcs (the same in every pass of every frame)
  // source | destination << 16 (at most 2048 states; source 0xffff: no arc), weight, and the pdf
  // of an emitting arc (-1: epsilon arc)
  uint32_t a_sd[kArcsPerThread];
  float a_w[kArcsPerThread];
  int a_pdf[kArcsPerThread];
#pragma unroll
  for (int k = 0; k < kArcsPerThread; ++k) {
    const int a = tid + k * kVitThreads;
    const bool ok = a < num_arcs;
    a_sd[k] = ok ? (static_cast<uint32_t>(__ldg(&fst.arc_src[a])) | (static_cast<uint32_t>(__ldg(&fst.arc_dst[a])) << 16))
                 : 0xffffu;
    a_w[k] = ok ? __ldg(&fst.arc_w[a]) : 0.0f;
    const int il = ok ? __ldg(&fst.arc_il[a]) : 0;
    a_pdf[k] = (ok && il != 0) ? __ldg(&tid2pdf[il]) : -1;
  }
  auto src_of = [&](int k) { return static_cast<int>(a_sd[k] & 0xffffu); };
  auto dst_of = [&](int k) { return static_cast<int>(a_sd[k] >> 16); };
  auto has_arc = [&](int k) { return (a_sd[k] & 0xffffu) != 0xffffu; };
  // slots of the distinct pdfs; a_pdf[k] becomes the slot of arc k's pdf
  for (int w = tid; w < W; w += kVitThreads) s_bits[w] = 0u;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kArcsPerThread; ++k)
    if (a_pdf[k] >= 0) atomicOr(&s_bits[a_pdf[k] >> 5], 1u << (a_pdf[k] & 31));
  __syncthreads();
  if (tid < 32) {
    int running = 0;
    for (int base = 0; base < W; base += 32) {
      const int w = base + tid;
      const int cnt = w < W ? __popc(s_bits[w]) : 0;
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (w < W) s_wpre[w] = running + inc - cnt;
      running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) s_D = running;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kArcsPerThread; ++k) {
    if (a_pdf[k] < 0) continue;
    const int pdf = a_pdf[k];
    const int slot = s_wpre[pdf >> 5] + __popc(s_bits[pdf >> 5] & ((1u << (pdf & 31)) - 1u));
    s_pdf[slot] = pdf;  // every arc of the pdf writes the same value
    a_pdf[k] = slot;
  }
  __syncthreads();
  const int D = s_D;
  // The acoustic costs of the NEXT frame are loaded (one load per distinct pdf) while the current
  // frame is searched: the row's DRAM latency is off the critical path of the frame step.

  // epsilon closure of frame table c under `cutoff`, winners, word back-pointers
  auto finish_frame = [&](int c, int p, double cutoff, const double (&td)[kArcsPerThread], const bool (&live)[kArcsPerThread]) {
    if (fst.has_eps) {
      for (int round = 0; round < 4096; ++round) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kArcsPerThread; ++k) {
          if (!has_arc(k) || a_pdf[k] >= 0) continue;
          const uint32_t cs = *reinterpret_cast<volatile uint32_t *>(&cost[c][src_of(k)]);
          if (cs == kInactive) continue;
          const double total = static_cast<double>(unord32(cs)) + static_cast<double>(a_w[k]);
          if (total > cutoff) continue;
          const uint32_t o = ord32(static_cast<float>(total));
          if (atomicMin(&cost[c][dst_of(k)], o) > o) s_changed = 1;
        }
        __syncthreads();
        const int changed = s_changed;  // read by everyone before thread 0 resets it
        __syncthreads();
        if (!changed) break;
      }
    }
    // winners against the final costs: emitting arcs from the totals kept in registers, epsilon
    // arcs from their source's final cost
#pragma unroll
    for (int k = 0; k < kArcsPerThread; ++k) {
      if (!has_arc(k)) continue;
      const uint32_t a = static_cast<uint32_t>(tid + k * kVitThreads);
      if (a_pdf[k] >= 0) {
        if (live[k] && ord32(static_cast<float>(td[k])) == cost[c][dst_of(k)]) atomicMin(&warc[c][dst_of(k)], a);
      } else if (fst.has_eps) {
        const uint32_t cs = cost[c][src_of(k)];
        if (cs == kInactive) continue;
        const double total = static_cast<double>(unord32(cs)) + static_cast<double>(a_w[k]);
        if (total <= cutoff && ord32(static_cast<float>(total)) == cost[c][dst_of(k)])
          atomicMin(&warc[c][dst_of(k)], a | 0x80000000u);
      }
    }
    __syncthreads();
    // word back-pointers (Decoder::InsertTok, src/decoder.cc:107-117)
    for (int s = tid; s < S; s += kVitThreads) bp[c][s] = cost[c][s] != kInactive ? -2 : -1;
    for (int pass = 0; pass < 256; ++pass) {
      if (tid == 0) s_unres = 0;
      __syncthreads();
      for (int s = tid; s < S; s += kVitThreads) {
        if (bp[c][s] != -2) continue;
        const uint32_t wa = warc[c][s];
        int pb;
        if (wa == kInactive) {
          pb = -1;  // the start token
        } else {
          const int a = static_cast<int>(wa & 0x7fffffffu);
          const int src = __ldg(&fst.arc_src[a]);
          if (wa >> 31) {
            pb = *reinterpret_cast<volatile int *>(&bp[c][src]);
            if (pb == -2) { atomicAdd(&s_unres, 1); continue; }
          } else {
            pb = bp[p][src];
          }
          const int ol = __ldg(&fst.arc_ol[a]);
          if (ol != 0) {
            const int r = atomicAdd(&s_log, 1);
            if (r >= max_log) { s_err = 2; pb = -1; }
            else { log_prev[r] = pb; log_ol[r] = ol; pb = r; }
          }
        }
        bp[c][s] = pb;
      }
      __syncthreads();
      const int unres = s_unres;
      __syncthreads();
      if (unres == 0) break;
      if (pass == 255 && tid == 0) s_err = 3;
    }
    __syncthreads();
  };

  // Closes a frame: clears the previous table, reduces the best cost of the new one (the next
  // frame's GetCutoff) and re-arms the two shared minima. Everyone is past a barrier on entry.
  auto end_frame = [&](int c, int p) {
    if (tid == 0) {
      s_best = kInactive;
      s_min = ~0ull;
    }
    __syncthreads();
    uint32_t m = kInactive;
    for (int s = tid; s < S; s += kVitThreads) {
      cost[p][s] = kInactive;
      warc[p][s] = kInactive;
      m = min(m, cost[c][s]);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0 && m != kInactive) atomicMin(&s_best, m);
    __syncthreads();
  };

  for (int u = blockIdx.x; u < n_utts; u += gridDim.x) {
    const int T = num_frames[u];
    const float *ll0 = loglik + row_off[u] * num_pdfs;
    for (int s = tid; s < S; s += kVitThreads) {
      cost[0][s] = cost[1][s] = kInactive;
      warc[0][s] = warc[1][s] = kInactive;
      bp[0][s] = bp[1][s] = -1;
    }
    if (tid == 0) { s_log = 0; s_err = 0; }
    __syncthreads();
    // ---- InitDecoding (src/decoder.cc:82-101)
    int cur = 0;
    if (tid == 0) cost[0][fst.start] = ord32(0.0f);
    if (T > 0)
      for (int d = tid; d < D; d += kVitThreads) s_ll[d] = __ldg(&ll0[s_pdf[d]]);
    __syncthreads();
    {
      double td[kArcsPerThread];
      bool live[kArcsPerThread];
#pragma unroll
      for (int k = 0; k < kArcsPerThread; ++k) { td[k] = 0.0; live[k] = false; }
      finish_frame(0, 1, INFINITY, td, live);
      end_frame(0, 1);
    }
    bool alive = true;
    for (int f = 0; f < T && alive && !s_err; ++f) {
      const int prev = cur;
      cur ^= 1;
      const float *ll_cur = s_ll + (f & 1) * Dmax;
      float ll_next[kArcsPerThread];  // D <= num_arcs <= kArcsPerThread * kVitThreads
      {
        const float *lln = ll0 + static_cast<int64_t>(f + 1) * num_pdfs;
#pragma unroll
        for (int k = 0; k < kArcsPerThread; ++k) {
          const int d = tid + k * kVitThreads;
          ll_next[k] = (f + 1 < T && d < D) ? __ldg(&lln[s_pdf[d]]) : 0.0f;
        }
      }
      // ---- GetCutoff below kBeamSize tokens: the best cost was reduced when the table was finished
      if (s_best == kInactive) { alive = false; break; }
      const float weight_cutoff = static_cast<float>(static_cast<double>(unord32(s_best)) + static_cast<double>(beam));
      // ---- ProcessEmitting, pass 1: totals of this thread's arcs, bound on the next frame
      double td[kArcsPerThread];
      bool live[kArcsPerThread];
      {
        unsigned long long m = ~0ull;
#pragma unroll
        for (int k = 0; k < kArcsPerThread; ++k) {
          live[k] = false;
          td[k] = 0.0;
          if (a_pdf[k] < 0) continue;
          const uint32_t cs = cost[prev][src_of(k)];
          if (cs == kInactive) continue;
          const float c = unord32(cs);
          if (c > weight_cutoff) continue;
          const float ac = -ll_cur[a_pdf[k]];
          td[k] = static_cast<double>(c) + static_cast<double>(a_w[k]) + static_cast<double>(ac);
          live[k] = true;
          m = min(m, ord64(td[k]));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((tid & 31) == 0) atomicMin(&s_min, m);
      }
      __syncthreads();
      if (s_min == ~0ull) { alive = false; break; }
      const double next_cutoff = unord64(s_min) + static_cast<double>(beam);
      // ---- pass 2: token costs of this frame
#pragma unroll
      for (int k = 0; k < kArcsPerThread; ++k) {
        live[k] = live[k] && td[k] <= next_cutoff;
        if (live[k]) atomicMin(&cost[cur][dst_of(k)], ord32(static_cast<float>(td[k])));
      }
      __syncthreads();
      finish_frame(cur, prev, static_cast<double>(static_cast<float>(next_cutoff)), td, live);
      {
        float *ll_nx = s_ll + ((f + 1) & 1) * Dmax;  // last read two frames ago
#pragma unroll
        for (int k = 0; k < kArcsPerThread; ++k) {
          const int d = tid + k * kVitThreads;
          if (d < D) ll_nx[d] = ll_next[k];
        }
      }
      end_frame(cur, prev);
    }

    // ---- BestPath (src/decoder.cc:300-339)
    if (tid == 0) s_min = ~0ull;
    __syncthreads();
    if (alive && !s_err) {
      for (int s = tid; s < S; s += kVitThreads) {
        if (cost[cur][s] == kInactive) continue;
        const float c = unord32(cost[cur][s]) + fst.final_w[s];
        if (c != INFINITY) atomicMin(&s_min, pack(c, static_cast<uint32_t>(s)));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int32_t *wo = words_out + static_cast<size_t>(u) * max_words;
      if (s_err) {
        n_words_out[u] = -s_err;
        weight_out[u] = 0.0f;
      } else if (s_min == ~0ull) {
        n_words_out[u] = 0;
        weight_out[u] = 0.0f;
      } else {
        const int s = static_cast<int>(static_cast<uint32_t>(s_min));
        float weight = cost_of(s_min);
        weight += fst.final_w[s];
        weight_out[u] = weight;
        int nw = 0;
        for (int r = bp[cur][s]; r >= 0; r = log_prev[r]) ++nw;
        n_words_out[u] = nw;
        int k = nw;
        for (int r = bp[cur][s]; r >= 0; r = log_prev[r]) {
          --k;
          if (k < max_words) wo[k] = log_ol[r];
        }
      }
    }
    __syncthreads();
  }
}

// the two value arrays lead every block's workspace
__global__ void viterbi_init_kernel(char *work_base, size_t work_stride, uint32_t n) {
  unsigned long long *vals = reinterpret_cast<unsigned long long *>(work_base + work_stride * blockIdx.y);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    vals[i] = kEmptyVal;
}

}  // namespace

int fst_build(Ctx *c, int num_states, int start, const float *final_w, const int32_t *first_arc,
              int num_arcs, const int32_t *arcs_raw, pkb_fst **out) {
  PKB_REQUIRE(c && out, "pkb_fst: NULL argument");
  PKB_REQUIRE(num_states > 0 && num_arcs >= 0, "pkb_fst: %d states, %d arcs", num_states, num_arcs);
  PKB_REQUIRE(start >= 0 && start < num_states, "pkb_fst: start state %d out of range", start);
  PKB_REQUIRE(final_w && first_arc && (arcs_raw || num_arcs == 0), "pkb_fst: NULL table");
  // arcs of state s: [first_arc[s], first arc of the next state that has arcs) -- Fst::CountArcs,
  // src/fst.cc:94-110
  std::vector<int32_t> begin(num_states + 1, num_arcs), src(num_arcs, 0);
  int32_t next = num_arcs;
  for (int s = num_states - 1; s >= 0; --s) {
    if (first_arc[s] >= 0) {
      PKB_REQUIRE(first_arc[s] <= next, "pkb_fst: arcs are not sorted by source state (state %d)", s);
      begin[s] = first_arc[s];
      next = first_arc[s];
    } else {
      begin[s] = next;
    }
  }
  bool has_eps = false;
  int max_il = 0;
  std::vector<int32_t> dst(num_arcs), il(num_arcs), ol(num_arcs);
  std::vector<float> wt(num_arcs);
  for (int s = 0; s < num_states; ++s)
    for (int a = begin[s]; a < begin[s + 1]; ++a) src[a] = s;
  for (int a = 0; a < num_arcs; ++a) {
    dst[a] = arcs_raw[4 * a];
    il[a] = arcs_raw[4 * a + 1];
    ol[a] = arcs_raw[4 * a + 2];
    memcpy(&wt[a], &arcs_raw[4 * a + 3], sizeof(float));
    PKB_REQUIRE(dst[a] >= 0 && dst[a] < num_states, "pkb_fst: arc %d leads to state %d", a, dst[a]);
    PKB_REQUIRE(il[a] >= 0, "pkb_fst: arc %d has a negative input label", a);
    has_eps = has_eps || il[a] == 0;
    max_il = std::max(max_il, il[a]);
  }
  pkb_fst *f = new pkb_fst();
  f->c = c;
  f->num_states = num_states;
  f->num_arcs = num_arcs;
  f->start = start;
  f->has_eps = has_eps;
  f->max_ilabel = max_il;
  const size_t n1 = static_cast<size_t>(num_states) + 1, na = std::max(num_arcs, 1);
  const size_t bytes = 4 * (static_cast<size_t>(num_states) + n1 + 5 * na);
  std::vector<char> host(bytes, 0);
  char *h = host.data();
  size_t o_fin = 0, o_beg = o_fin + 4 * num_states, o_src = o_beg + 4 * n1, o_dst = o_src + 4 * na,
         o_il = o_dst + 4 * na, o_ol = o_il + 4 * na, o_w = o_ol + 4 * na;
  memcpy(h + o_fin, final_w, 4 * static_cast<size_t>(num_states));
  memcpy(h + o_beg, begin.data(), 4 * n1);
  if (num_arcs) {
    memcpy(h + o_src, src.data(), 4 * static_cast<size_t>(num_arcs));
    memcpy(h + o_dst, dst.data(), 4 * static_cast<size_t>(num_arcs));
    memcpy(h + o_il, il.data(), 4 * static_cast<size_t>(num_arcs));
    memcpy(h + o_ol, ol.data(), 4 * static_cast<size_t>(num_arcs));
    memcpy(h + o_w, wt.data(), 4 * static_cast<size_t>(num_arcs));
  }
  int rc = f->buf.ensure(bytes);
  if (rc == PKB_OK) rc = upload(c, f->buf.p, h, bytes);
  if (rc != PKB_OK) {
    f->buf.release();
    delete f;
    return rc;
  }
  const char *d = f->buf.as<char>();
  f->d_final = reinterpret_cast<const float *>(d + o_fin);
  f->d_arc_begin = reinterpret_cast<const int32_t *>(d + o_beg);
  f->d_arc_src = reinterpret_cast<const int32_t *>(d + o_src);
  f->d_arc_dst = reinterpret_cast<const int32_t *>(d + o_dst);
  f->d_arc_il = reinterpret_cast<const int32_t *>(d + o_il);
  f->d_arc_ol = reinterpret_cast<const int32_t *>(d + o_ol);
  f->d_arc_w = reinterpret_cast<const float *>(d + o_w);
  *out = f;
  return PKB_OK;
}

int launch_viterbi(Ctx *c, const pkb_fst *fst, const ViterbiConfig &cfg, const float *d_loglik,
                   int num_pdfs, const int64_t *d_row_off, const int32_t *d_num_frames, int n_utts,
                   const int32_t *d_tid2pdf, int n_tids, DevBuf *work, int32_t *d_words,
                   int32_t *d_n_words, float *d_weight) {
  (void)n_tids;
  if (n_utts == 0) return PKB_OK;
  PKB_REQUIRE(cfg.max_tokens >= 16 && cfg.max_log >= 16 && cfg.max_words >= 1 && cfg.beam > 0.0f,
              "pkb_batch_decode: bad configuration");
  FstDev fd;
  fd.num_states = fst->num_states;
  fd.start = fst->start;
  fd.has_eps = fst->has_eps ? 1 : 0;
  fd.final_w = fst->d_final;
  fd.arc_begin = fst->d_arc_begin;
  fd.arc_src = fst->d_arc_src;
  fd.arc_dst = fst->d_arc_dst;
  fd.arc_il = fst->d_arc_il;
  fd.arc_ol = fst->d_arc_ol;
  fd.arc_w = fst->d_arc_w;
  static const char *env_dense = getenv("PKB_VIT_DENSE");  // tuning knob: 0 forces the table kernel
  if (fst->num_states <= 2048 && fst->num_arcs <= 16 * kVitThreads && !(env_dense && atoi(env_dense) == 0)) {
    // dense kernel: no token capacity to run out of; the global workspace holds the word records only
    const size_t stride = (2 * sizeof(int) * static_cast<size_t>(cfg.max_log) + 255) & ~static_cast<size_t>(255);
    // one resident wave at most (the workspace is per block); the kernels are built for three
    // blocks per SM up to eight arcs per thread
    const int grid = std::min(n_utts, c->sm_count * 4);
    PKB_TRY(work->ensure(stride * grid));
    const size_t smem = sizeof(uint32_t) * (6 * static_cast<size_t>(fst->num_states) +
                                            2 * static_cast<size_t>((num_pdfs + 31) / 32) +
                                            3 * static_cast<size_t>(std::min(fst->num_arcs, num_pdfs)));
    LaunchScope scope(c, PKB_KERNEL_MISC);
#define PKB_VIT_DENSE(APT)                                                                               \
  do {                                                                                                   \
    if (smem > 48 * 1024)                                                                                \
      PKB_CUDA(cudaFuncSetAttribute(viterbi_dense_kernel<APT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    static_cast<int>(smem)));                                            \
    viterbi_dense_kernel<APT><<<grid, kVitThreads, smem, c->stream>>>(                                   \
        fd, fst->num_arcs, cfg.beam, cfg.max_log, cfg.max_words, d_loglik, num_pdfs, d_row_off,          \
        d_num_frames, n_utts, d_tid2pdf, work->as<char>(), stride, d_words, d_n_words, d_weight);        \
  } while (0)
    if (fst->num_arcs <= 4 * kVitThreads) PKB_VIT_DENSE(4);
    else if (fst->num_arcs <= 8 * kVitThreads) PKB_VIT_DENSE(8);
    else PKB_VIT_DENSE(16);
#undef PKB_VIT_DENSE
    PKB_CUDA(cudaGetLastError());
    return PKB_OK;
  }
  // a graph with at most 512 states can never hold more tokens than states: its tables go to
  // shared memory (capacity = the state count), whatever max_tokens says
  const bool small = fst->num_states <= 512;
  const int max_tok = small ? std::max(16, fst->num_states) : cfg.max_tokens;
  uint32_t H = 64;
  while (H < 2u * static_cast<uint32_t>(max_tok)) H <<= 1;
  const size_t table_bytes = 2 * sizeof(unsigned long long) * H +
                             2 * (2 * sizeof(int) * H + 2 * sizeof(int) * max_tok) + sizeof(int) * H;
  const size_t log_bytes = 2 * sizeof(int) * static_cast<size_t>(cfg.max_log);
  const size_t stride_raw = (small ? 0 : table_bytes) + log_bytes;
  const size_t stride = (stride_raw + 255) & ~static_cast<size_t>(255);
  // persistent blocks: a few per SM, each walking its share of the utterances with one workspace
  const int grid = std::min(n_utts, c->sm_count * 8);
  const size_t bytes = stride * grid;
  PKB_TRY(work->ensure(bytes));
  if (!small) {
    // invariant the kernel relies on (and restores per utterance): keys 0, values empty, queue flags 0
    PKB_CUDA(cudaMemsetAsync(work->p, 0, bytes, c->stream));
    LaunchScope scope(c, PKB_KERNEL_MISC);
    viterbi_init_kernel<<<dim3(4, grid), 256, 0, c->stream>>>(work->as<char>(), stride, 2 * H);
    PKB_CUDA(cudaGetLastError());
  }
  // (staging each frame's log-likelihood row in shared memory with cp.async was measured slower,
  // 44 vs 26 ms for 128 utterances: the larger carve-out takes the L1 space that keeps the FST hot)
  static const char *env_group = getenv("PKB_VIT_GROUP");  // tuning knob: 8 or 32
  const size_t dyn_smem = small ? table_bytes : 0;
  const bool wide = env_group ? atoi(env_group) == 32
                              : static_cast<double>(fst->num_arcs) >= 12.0 * fst->num_states;
  LaunchScope scope(c, PKB_KERNEL_MISC);
#define PKB_VIT_LAUNCH(G)                                                                                  \
  do {                                                                                                     \
    if (dyn_smem > 48 * 1024)                                                                              \
      PKB_CUDA(cudaFuncSetAttribute(viterbi_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    static_cast<int>(dyn_smem)));                                          \
    viterbi_kernel<G><<<grid, kVitThreads, dyn_smem, c->stream>>>(                                         \
        fd, cfg.beam, max_tok, H - 1, cfg.max_log, cfg.max_words, d_loglik, num_pdfs, d_row_off,           \
        d_num_frames, n_utts, d_tid2pdf, work->as<char>(), stride, small ? 1 : 0, d_words, d_n_words,      \
        d_weight);                                                                                         \
  } while (0)
  if (wide) PKB_VIT_LAUNCH(32);
  else PKB_VIT_LAUNCH(8);
#undef PKB_VIT_LAUNCH
  PKB_CUDA(cudaGetLastError());
  return PKB_OK;
}

}  // namespace pkb
