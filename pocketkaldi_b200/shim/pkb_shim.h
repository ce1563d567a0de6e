// Reference-shaped C++ shims: pocketkaldi's Fbank / CMVN / Nnet / AcousticModel classes and
// the pk_decodable_* functions, implemented on the C ABI of libpkb200 (include/pkb200.h) with
// batches of one utterance. Compiling the reference's callers (src/pocketkaldi.cc,
// src/decoder.cc, src/main.cc) with `-include pkb_shim.h` makes them use these declarations:
// the include guards of the five reference headers this file replaces are defined below, so
// the reference's own fbank.h / cmvn.h / nnet.h / am.h / decodable.h expand to nothing.
//
//   replaced interface                              reference declaration
//   pocketkaldi::Fbank::Compute                     src/fbank.h:46-53
//   pocketkaldi::CMVN::CMVN / GetFrame              src/cmvn.h:17-26
//   pocketkaldi::Nnet::Read / Propagate             src/nnet.h:88-96
//   pocketkaldi::AcousticModel::Read / Compute ...  src/am.h:21-38
//   pk_decodable_init / destroy / loglikelihood /   src/decodable.h:15-41
//     islastframe
//
// Everything else the callers need (pk_vector_t, pk_matrix_t, Status, Configuration,
// util::ReadableFile, Matrix<float>::Read ...) comes from the unmodified reference headers
// and objects. There is no CPU fallback: a CUDA failure aborts with the library's message,
// where the reference would assert.
#ifndef PKB_SHIM_H_
#define PKB_SHIM_H_

#define POCKETKALDI_FBANK_H_
#define POCKETKALDI_ONLINE_CMVN_H_
#define POCKETKALDI_NNET_H_
#define POCKETKALDI_AM_H_
#define POCKETKALDI_DECODABLE_H_

// compile-time constants the replaced headers export (src/fbank.h:7-13, src/cmvn.h:10-11)
#define PK_SAMPLERATE 16000
#define PK_FRAMESHIFT_MS 10.0
#define PK_FRAMELENGTH_MS 25.0
#define PK_FBANK_DIM 40
#define PK_FBANK_LOWFREQ 20
#define PK_FBANK_HIGHFREQ (PK_SAMPLERATE / 2)
#define PK_PREEMPH_COEFF 0.97
#define PK_ONLINECMVN_WINDOW 600
#define PK_ONLINECMVN_GLOBALFRAMES 200

#include <stdint.h>

#include <vector>

#include "configuration.h"
#include "matrix.h"
#include "pocketkaldi.h"
#include "status.h"
#include "util.h"
#include "vector.h"

struct pkb_am;

namespace pocketkaldi {

class Fbank {
 public:
  Fbank();
  ~Fbank();
  // wave: float samples in int16 range; fbank_feature: resized to {nrow = 40, ncol = T}
  void Compute(const pk_vector_t *wave, pk_matrix_t *fbank_feature);
};

class CMVN {
 public:
  CMVN(const pk_vector_t *global_stats, const pk_matrix_t *raw_feats);  // both borrowed
  ~CMVN();
  // frames must be requested as 0, 1, 2, ... (as in the reference, src/cmvn.cc:38)
  void GetFrame(int frame, pk_vector_t *feats);

 private:
  const pk_vector_t *global_stats_;
  const pk_matrix_t *raw_feats_;
  std::vector<float> normalised_;  // whole utterance, computed on the GPU at frame 0
  int cached_frame_;
};

class Nnet {
 public:
  Nnet();
  ~Nnet();
  Status Read(util::ReadableFile *fd);
  void Propagate(const pk_matrix_t *in, pk_matrix_t *out) const;

 private:
  Nnet(const Nnet &);
  Nnet &operator=(const Nnet &);
  pkb_am *model_;
};

class AcousticModel {
 public:
  AcousticModel();
  ~AcousticModel();
  Status Read(const Configuration &conf);
  int TransitionIdToPdfId(int transition_id) const;
  void Compute(const pk_matrix_t *frames, pk_matrix_t *loglikelihood);
  int num_pdfs() const;
  // Compute with the decodable's prob_scale folded into the GPU epilogue
  void ComputeScaled(const pk_matrix_t *frames, float prob_scale, pk_matrix_t *loglikelihood);
  // the C-ABI model behind this object (batch drivers, lazy decodable)
  pkb_am *handle() const { return model_; }

 private:
  AcousticModel(const AcousticModel &);
  AcousticModel &operator=(const AcousticModel &);
  pkb_am *model_;
};

}  // namespace pocketkaldi

using pocketkaldi::AcousticModel;
using pocketkaldi::Nnet;

// Lazy mode. With PKB_DECODABLE_CHUNK=<frames> in the environment,
// pk_decodable_init returns as soon as the GPU work is queued; the matrix lands in page-locked
// memory chunk by chunk and pk_decodable_loglikelihood waits only for the chunk that holds the
// frame it is asked for, so Decoder::Decode (src/decoder.cc:49) overlaps with the copy-out.
struct pkb_lazy_decodable;
struct pkb_event;

typedef struct pk_decodable_t {
  // {nrow = num_pdfs, ncol = frames}. Eager mode: malloc-family host memory like the reference.
  // Lazy / attached mode: page-locked memory that `lazy` (or the caller) owns.
  pk_matrix_t log_prob;
  AcousticModel *am;
  pkb_lazy_decodable *lazy;  // NULL in eager mode
  int frames_ready;          // frames [0, frames_ready) are valid in log_prob
} pk_decodable_t;

POCKETKALDI_EXPORT
void pk_decodable_init(pk_decodable_t *self, AcousticModel *am, float prob_scale,
                       const pk_matrix_t *feats);
POCKETKALDI_EXPORT
void pk_decodable_destroy(pk_decodable_t *self);
POCKETKALDI_EXPORT
float pk_decodable_loglikelihood(pk_decodable_t *self, int frame, int trans_id);
POCKETKALDI_EXPORT
bool pk_decodable_islastframe(pk_decodable_t *self, int frame);

// The process-wide context the shim objects live on (device from PKB_DEVICE).
struct pkb_ctx;
POCKETKALDI_EXPORT
pkb_ctx *pkb_shim_context();

// Extension for batch drivers: a decodable over `frames` rows of [num_pdfs] floats that a batch
// pipeline is copying (or has copied) into caller-owned host memory; `ready` (may be NULL) is
// waited for on the first look-up. pk_decodable_destroy leaves the memory and the event alone.
POCKETKALDI_EXPORT
void pk_decodable_attach(pk_decodable_t *self, AcousticModel *am, float *log_prob, int frames,
                         pkb_event *ready);

// Same over the compact output of a batch (pkb_batch_set_compact, include/pkb200.h): rows of IEEE
// half bits plus one FP32 offset per frame; the look-up returns
// prob_scale * (half(h16[frame][pdf]) + off[frame]), i.e. the tail of the GPU epilogue moves into
// pk_decodable_loglikelihood and the rows cross PCIe at half the size.
POCKETKALDI_EXPORT
void pk_decodable_attach_compact(pk_decodable_t *self, AcousticModel *am, const uint16_t *h16,
                                 const float *off, int frames, float prob_scale, pkb_event *ready);

#endif  // PKB_SHIM_H_
