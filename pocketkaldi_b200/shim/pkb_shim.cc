// Implementation of the reference-shaped shims on the libpkb200 C ABI. See pkb_shim.h.

#include "pkb_shim.h"

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "pkb200.h"

namespace {

// One context per process: device from PKB_DEVICE (default 0).
pkb_ctx_t *shim_ctx() {
  static pkb_ctx_t *ctx = nullptr;
  if (ctx == nullptr) {
    const char *dev = getenv("PKB_DEVICE");
    if (pkb_create(dev ? atoi(dev) : 0, &ctx) != PKB_OK) {
      fprintf(stderr, "pocketkaldi-b200: %s\n", pkb_last_error());
      abort();
    }
  }
  return ctx;
}

// GEMM arithmetic from PKB_PRECISION: "bf16", "fp16", "bf16x3", "fp16x3", "fp16c8" or "fp16r" (default:
// bf16x3, a mode that meets the parity bar).
int shim_precision() {
  const char *p = getenv("PKB_PRECISION");
  if (p != nullptr && strcmp(p, "bf16") == 0) return PKB_PREC_BF16;
  if (p != nullptr && strcmp(p, "fp16") == 0) return PKB_PREC_FP16;
  if (p != nullptr && strcmp(p, "fp16x3") == 0) return PKB_PREC_FP16X3;
  if (p != nullptr && strcmp(p, "fp16c8") == 0) return PKB_PREC_FP16C8;
  if (p != nullptr && strcmp(p, "fp16r") == 0) return PKB_PREC_FP16R;
  return PKB_PREC_BF16X3;
}

void must(int rc, const char *what) {
  if (rc != PKB_OK) {
    fprintf(stderr, "pocketkaldi-b200: %s: %s\n", what, pkb_last_error());
    abort();
  }
}

pocketkaldi::Status to_status(int rc) {
  using pocketkaldi::Status;
  if (rc == PKB_OK) return Status::OK();
  std::string msg = pkb_last_error();
  // the library message already carries the reference's "IOError: " / "Corruption: " prefix
  const char *prefixes[] = {"IOError: ", "Corruption: "};
  for (const char *p : prefixes)
    if (msg.compare(0, strlen(p), p) == 0) msg = msg.substr(strlen(p));
  if (rc == PKB_ERR_IO) return Status::IOError(msg);
  if (rc == PKB_ERR_CORRUPT) return Status::Corruption(msg);
  if (rc == PKB_ERR_UNSUPPORTED) return Status::NotImplemented(msg);
  return Status::RuntimeError(msg);
}

}  // namespace

namespace pocketkaldi {

// ---------------------------------------------------------------- Fbank
Fbank::Fbank() { shim_ctx(); }
Fbank::~Fbank() {}

void Fbank::Compute(const pk_vector_t *wave, pk_matrix_t *fbank_feature) {
  const int32_t n = wave->dim;
  const int frames = pkb_fbank_num_frames(n);
  if (frames == 0) {
    pk_matrix_resize(fbank_feature, 0, 0);
    return;
  }
  pk_matrix_resize(fbank_feature, PK_FBANK_DIM, frames);
  must(pkb_fbank_f32(shim_ctx(), wave->data, &n, 1, fbank_feature->data, nullptr),
       "Fbank::Compute");
}

// ---------------------------------------------------------------- CMVN
CMVN::CMVN(const pk_vector_t *global_stats, const pk_matrix_t *raw_feats)
    : global_stats_(global_stats), raw_feats_(raw_feats), cached_frame_(-1) {}

CMVN::~CMVN() {
  raw_feats_ = nullptr;
  cached_frame_ = 0;
}

void CMVN::GetFrame(int frame, pk_vector_t *feats) {
  assert(cached_frame_ == frame - 1 && "CMVN frames must be requested in order");
  assert(frame >= 0 && frame < raw_feats_->ncol);
  if (normalised_.empty()) {
    assert(raw_feats_->nrow == PK_FBANK_DIM && global_stats_->dim == PK_FBANK_DIM + 1);
    const int32_t frames = raw_feats_->ncol;
    normalised_.resize(static_cast<size_t>(frames) * PK_FBANK_DIM);
    must(pkb_cmvn(shim_ctx(), raw_feats_->data, &frames, 1, global_stats_->data,
                  normalised_.data()),
         "CMVN::GetFrame");
  }
  if (feats->dim != PK_FBANK_DIM) pk_vector_resize(feats, PK_FBANK_DIM);
  memcpy(feats->data, &normalised_[static_cast<size_t>(frame) * PK_FBANK_DIM],
         sizeof(float) * PK_FBANK_DIM);
  cached_frame_ = frame;
}

// ---------------------------------------------------------------- Nnet
Nnet::Nnet() : model_(nullptr) {}
Nnet::~Nnet() { pkb_am_destroy(model_); }

Status Nnet::Read(util::ReadableFile *fd) {
  // NNT0 / LAY0 container (format of src/nnet.cc:80-147); the MAT0 / VEC0 payloads are read
  // with the reference's own Matrix / Vector readers.
  int32_t section_size = 0, num_layers = 0;
  PK_CHECK_STATUS(fd->ReadAndVerifyString("NNT0"));
  PK_CHECK_STATUS(fd->ReadValue<int32_t>(&section_size));
  PK_CHECK_STATUS(fd->ReadValue<int32_t>(&num_layers));
  std::vector<int32_t> types, out_dims, in_dims;
  std::vector<std::vector<float> > weights, biases;
  for (int i = 0; i < num_layers; ++i) {
    int32_t layer_size = 0, type = 0;
    PK_CHECK_STATUS(fd->ReadAndVerifyString("LAY0"));
    PK_CHECK_STATUS(fd->ReadValue<int32_t>(&layer_size));
    PK_CHECK_STATUS(fd->ReadValue<int32_t>(&type));
    if (layer_size != 4) {
      return Status::Corruption(util::Format(
          "read_layer: section_size == 4 expected, but {} found ({})", layer_size, fd->filename()));
    }
    if (type == 0) {
      Matrix<float> W;
      Vector<float> b;
      PK_CHECK_STATUS(W.Read(fd));
      PK_CHECK_STATUS(b.Read(fd));
      if (b.Dim() != W.NumRows()) {
        return Status::Corruption(util::Format("linear layer: W has {} rows but b has {} ({})",
                                               W.NumRows(), b.Dim(), fd->filename()));
      }
      std::vector<float> w(static_cast<size_t>(W.NumRows()) * W.NumCols());
      for (int r = 0; r < W.NumRows(); ++r)
        memcpy(&w[static_cast<size_t>(r) * W.NumCols()], W.Data() + static_cast<size_t>(r) * W.Stride(),
               sizeof(float) * W.NumCols());
      weights.push_back(w);
      biases.push_back(std::vector<float>(b.Data(), b.Data() + b.Dim()));
      out_dims.push_back(W.NumRows());
      in_dims.push_back(W.NumCols());
    } else if (type < 0 || type > 3) {
      return Status::Corruption(
          util::Format("read_layer: unexpected layer type: {} ({})", type, fd->filename()));
    }
    types.push_back(type);
  }
  std::vector<const float *> wp, bp;
  for (size_t i = 0; i < weights.size(); ++i) {
    wp.push_back(weights[i].data());
    bp.push_back(biases[i].data());
  }
  pkb_am_destroy(model_);
  model_ = nullptr;
  return to_status(pkb_am_create(shim_ctx(), static_cast<int>(types.size()), types.data(), wp.data(),
                                 bp.data(), out_dims.data(), in_dims.data(), nullptr, 0, 0, 0,
                                 nullptr, 0, shim_precision(), &model_));
}

void Nnet::Propagate(const pk_matrix_t *in, pk_matrix_t *out) const {
  assert(model_ != nullptr && "Nnet::Propagate before Read");
  // column-major {nrow = dim, ncol = T} is row-major [T][dim] (src/nnet.cc:150)
  const int rows = in->ncol, dim = in->nrow;
  pk_matrix_resize(out, pkb_am_num_pdfs(model_), rows);
  must(pkb_nnet_propagate(shim_ctx(), model_, in->data, rows, dim, out->data), "Nnet::Propagate");
}

// ---------------------------------------------------------------- AcousticModel
AcousticModel::AcousticModel() : model_(nullptr) {}
AcousticModel::~AcousticModel() { pkb_am_destroy(model_); }

Status AcousticModel::Read(const Configuration &conf) {
  pkb_am_destroy(model_);
  model_ = nullptr;
  return to_status(pkb_am_load(shim_ctx(), conf.filename().c_str(), shim_precision(), &model_));
}

int AcousticModel::TransitionIdToPdfId(int transition_id) const {
  return pkb_am_tid2pdf(model_, transition_id);
}

int AcousticModel::num_pdfs() const { return pkb_am_num_pdfs(model_); }

void AcousticModel::ComputeScaled(const pk_matrix_t *frames, float prob_scale,
                                  pk_matrix_t *loglikelihood) {
  assert(model_ != nullptr && "AcousticModel::Compute before Read");
  const int32_t rows = frames->ncol;
  pk_matrix_resize(loglikelihood, num_pdfs(), rows);
  must(pkb_am_compute(shim_ctx(), model_, frames->data, &rows, 1, frames->nrow, prob_scale,
                      loglikelihood->data),
       "AcousticModel::Compute");
}

void AcousticModel::Compute(const pk_matrix_t *frames, pk_matrix_t *loglikelihood) {
  ComputeScaled(frames, 1.0f, loglikelihood);
}

}  // namespace pocketkaldi

pkb_ctx *pkb_shim_context() { return shim_ctx(); }

// ---------------------------------------------------------------- decodable
// Storage of a lazy decodable: page-locked matrix + one event per chunk. cudaHostAlloc costs
// milliseconds, and the reference builds one decodable after the other (src/pocketkaldi.cc:210-247),
// so the last released storage is kept for the next utterance.
struct pkb_lazy_decodable {
  float *pinned = nullptr;
  size_t capacity = 0;  // floats
  std::vector<pkb_event_t *> events;
  int chunk = 0;
  pkb_event_t *attached = nullptr;  // attach mode: the caller's event (not owned)
  bool attached_mode = false;
  // compact attach mode: [frames][pdfs] IEEE half bits + one offset per frame (not owned)
  const uint16_t *h16 = nullptr;
  const float *off = nullptr;
  float scale = 1.0f;
};

namespace {
// IEEE half -> float without relying on F16C: 1-5-10 bits, subnormals and inf/nan included
inline float half_bits_to_float(uint16_t h) {
  const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {
      int e = -1;
      do { man <<= 1; ++e; } while ((man & 0x400u) == 0);
      bits = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | man << 13;
  } else {
    bits = sign | (exp + 127 - 15) << 23 | man << 13;
  }
  float f;
  memcpy(&f, &bits, sizeof(f));
  return f;
}
}  // namespace

namespace {

pkb_lazy_decodable *g_spare = nullptr;

int decodable_chunk_frames() {
  const char *e = getenv("PKB_DECODABLE_CHUNK");
  return e != nullptr ? atoi(e) : 0;
}

pkb_lazy_decodable *lazy_acquire(size_t floats, int n_events) {
  pkb_lazy_decodable *l = g_spare != nullptr ? g_spare : new pkb_lazy_decodable;
  g_spare = nullptr;
  if (l->capacity < floats) {
    pkb_host_free(l->pinned);
    l->pinned = nullptr;
    l->capacity = 0;
    void *p = nullptr;
    must(pkb_host_alloc(&p, floats * sizeof(float)), "pk_decodable_init (page-locked buffer)");
    l->pinned = static_cast<float *>(p);
    l->capacity = floats;
  }
  while (static_cast<int>(l->events.size()) < n_events) {
    pkb_event_t *e = nullptr;
    must(pkb_event_create(shim_ctx(), &e), "pk_decodable_init (event)");
    l->events.push_back(e);
  }
  return l;
}

void lazy_release(pkb_lazy_decodable *l) {
  if (l->attached_mode) {
    delete l;
    return;
  }
  if (g_spare == nullptr) {
    g_spare = l;
    return;
  }
  for (pkb_event_t *e : l->events) pkb_event_destroy(e);
  pkb_host_free(l->pinned);
  delete l;
}

}  // namespace

void pk_decodable_init(pk_decodable_t *self, AcousticModel *am, float prob_scale,
                       const pk_matrix_t *feats) {
  self->am = am;
  self->lazy = nullptr;
  const int chunk = decodable_chunk_frames();
  const int frames = feats->ncol, pdfs = am->num_pdfs();
  if (chunk <= 0 || frames == 0) {
    pk_matrix_init(&self->log_prob, pdfs, frames);
    am->ComputeScaled(feats, prob_scale, &self->log_prob);
    self->frames_ready = frames;
    return;
  }
  const int n_chunks = (frames + chunk - 1) / chunk;
  pkb_lazy_decodable *l = lazy_acquire(static_cast<size_t>(frames) * pdfs, n_chunks);
  l->chunk = chunk;
  self->lazy = l;
  self->log_prob.nrow = pdfs;
  self->log_prob.ncol = frames;
  self->log_prob.data = l->pinned;
  self->frames_ready = 0;
  must(pkb_am_compute_chunked(shim_ctx(), am->handle(), feats->data, frames, feats->nrow, prob_scale,
                              l->pinned, chunk, l->events.data(), n_chunks),
       "pk_decodable_init");
}

void pk_decodable_attach(pk_decodable_t *self, AcousticModel *am, float *log_prob, int frames,
                         pkb_event *ready) {
  pkb_lazy_decodable *l = new pkb_lazy_decodable;
  l->attached_mode = true;
  l->attached = ready;
  self->am = am;
  self->lazy = l;
  self->log_prob.nrow = am->num_pdfs();
  self->log_prob.ncol = frames;
  self->log_prob.data = log_prob;
  self->frames_ready = ready != nullptr ? 0 : frames;
}

void pk_decodable_attach_compact(pk_decodable_t *self, AcousticModel *am, const uint16_t *h16,
                                 const float *off, int frames, float prob_scale, pkb_event *ready) {
  pk_decodable_attach(self, am, nullptr, frames, ready);
  self->lazy->h16 = h16;
  self->lazy->off = off;
  self->lazy->scale = prob_scale;
}

void pk_decodable_destroy(pk_decodable_t *self) {
  if (self->lazy == nullptr) {
    pk_matrix_destroy(&self->log_prob);
  } else {
    pkb_lazy_decodable *l = self->lazy;
    // the copies still in flight write into the storage: let them finish before it is reused
    if (!l->attached_mode && self->frames_ready < self->log_prob.ncol)
      must(pkb_event_wait(l->events[(self->log_prob.ncol - 1) / l->chunk]), "pk_decodable_destroy");
    lazy_release(l);
    self->lazy = nullptr;
    self->log_prob.data = nullptr;
    self->log_prob.nrow = self->log_prob.ncol = 0;
  }
  self->am = NULL;
}

float pk_decodable_loglikelihood(pk_decodable_t *self, int frame, int trans_id) {
  if (frame >= self->frames_ready) {
    pkb_lazy_decodable *l = self->lazy;
    assert(l != nullptr && frame < self->log_prob.ncol);
    if (l->attached_mode) {
      must(pkb_event_wait(l->attached), "pk_decodable_loglikelihood");
      self->frames_ready = self->log_prob.ncol;
    } else {
      // chunks complete in stream order: waiting for this frame's chunk covers the earlier ones
      const int c = frame / l->chunk;
      must(pkb_event_wait(l->events[c]), "pk_decodable_loglikelihood");
      const int upto = (c + 1) * l->chunk;
      self->frames_ready = upto < self->log_prob.ncol ? upto : self->log_prob.ncol;
    }
  }
  const int pdf_id = self->am->TransitionIdToPdfId(trans_id);
  if (self->lazy != nullptr && self->lazy->h16 != nullptr) {
    // compact rows: the last step of the epilogue (offset, prob_scale) happens per look-up
    const pkb_lazy_decodable *l = self->lazy;
    return l->scale * (half_bits_to_float(l->h16[static_cast<size_t>(frame) * self->log_prob.nrow + pdf_id]) +
                       l->off[frame]);
  }
  return self->log_prob.data[static_cast<size_t>(frame) * self->log_prob.nrow + pdf_id];
}

bool pk_decodable_islastframe(pk_decodable_t *self, int frame) {
  assert(frame < self->log_prob.ncol);
  return frame == self->log_prob.ncol - 1;
}
