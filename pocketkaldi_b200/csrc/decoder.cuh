// GPU token-passing Viterbi beam search over a pocketkaldi FST.
//
// Replaces, for whole batches whose log-likelihoods are already resident on the device,
// Decoder::Decode + BestPath (src/decoder.cc:39-339) over Fst (src/fst.cc:29-129) with the
// decodable look-up of src/decodable.cc:24-31 -- the consumer that otherwise pulls the
// [frames x pdfs] matrix over PCIe and walks it on one host core per utterance.
#pragma once

#include <vector>

#include "common.cuh"

struct pkb_fst {
  pkb::Ctx *c = nullptr;
  int num_states = 0, num_arcs = 0, start = 0;
  bool has_eps = false;  // some arc has input label 0
  int max_ilabel = 0;    // largest input label (must index the model's tid2pdf map)
  pkb::DevBuf buf;       // one allocation, pointers below
  const float *d_final = nullptr;      // [num_states]
  const int32_t *d_arc_begin = nullptr;  // [num_states + 1]: arcs of state s are [begin[s], begin[s + 1])
  const int32_t *d_arc_src = nullptr;    // [num_arcs]
  const int32_t *d_arc_dst = nullptr, *d_arc_il = nullptr, *d_arc_ol = nullptr;
  const float *d_arc_w = nullptr;
};

namespace pkb {

// Builds the device FST from the file layout of src/fst.cc:29-92: final[state], first_arc[state]
// (-1: no arcs) and arcs {next_state, input_label, output_label, weight} sorted by source state.
int fst_build(Ctx *c, int num_states, int start, const float *final_w, const int32_t *first_arc,
              int num_arcs, const int32_t *arcs_raw /* 4 x int32 per arc, weight as float bits */,
              pkb_fst **out);

struct ViterbiConfig {
  float beam = 16.0f;        // Decoder::beam_ (src/decoder.cc:29)
  int max_tokens = 4096;     // tokens per frame and utterance (the reference caps at 30000 by sampling)
  int max_log = 1 << 18;     // word back-pointer records per utterance
  int max_words = 256;       // words returned per utterance
};

// Decodes n_utts utterances. loglik: [rows][num_pdfs] scaled log-likelihoods where utterance u owns
// rows [row_off[u], row_off[u] + num_frames[u]). tid2pdf: device map of input labels.
// words_out [n_utts][max_words] in spoken order, n_words_out[u] (-1: the search failed: token or
// log capacity exceeded), weight_out[u] = Hypothesis::weight() of Decoder::BestPath.
int launch_viterbi(Ctx *c, const pkb_fst *fst, const ViterbiConfig &cfg, const float *d_loglik,
                   int num_pdfs, const int64_t *d_row_off, const int32_t *d_num_frames, int n_utts,
                   const int32_t *d_tid2pdf, int n_tids, DevBuf *work, int32_t *d_words,
                   int32_t *d_n_words, float *d_weight);

}  // namespace pkb
