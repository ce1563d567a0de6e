// TEST INFRASTRUCTURE -- not product code.
//
// C-API wrapper around the UNMODIFIED pocketkaldi reference, compiled by
// oracle/build_ref.sh from the sources where they lie under /root/reference.
// It only exists so that tests/ and bench.py's reference arm can drive the
// reference's own classes from Python (ctypes). Nothing in pocketkaldi_b200/
// may link or load this.
//
// Every function is a thin call into a reference entry point:
//   ref_fbank          -> pocketkaldi::Fbank::Compute          (src/fbank.cc:267)
//   ref_cmvn           -> pocketkaldi::CMVN::GetFrame x T      (src/cmvn.cc:103)
//   ref_srfft          -> pk_srfft_compute                     (src/srfft.cc:371)
//   ref_gemm           -> pocketkaldi::MatMat                  (src/matrix.cc:418)
//   ref_nnet_propagate -> pocketkaldi::Nnet::Read / Propagate  (src/nnet.cc:132,149)
//   ref_am_*           -> pocketkaldi::AcousticModel           (src/am.cc:23,90)
//   ref_decodable_fill -> pk_decodable_init/loglikelihood      (src/decodable.cc:8,24)
//   ref_decode_wav     -> pk_load / pk_read_audio / pk_process (src/pocketkaldi.cc:72,176)
//   ref_time_path      -> the per-utterance hot path of pk_process, timed on a
//                         std::thread pool

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "am.h"
#include "cmvn.h"
#include "configuration.h"
#include "decodable.h"
#include "fbank.h"
#include "matrix.h"
#include "nnet.h"
#include "pocketkaldi.h"
#include "srfft.h"
#include "vector.h"

using pocketkaldi::AcousticModel;
using pocketkaldi::CMVN;
using pocketkaldi::Configuration;
using pocketkaldi::Fbank;
using pocketkaldi::Nnet;
using pocketkaldi::Status;

namespace {

void set_err(char *err, int errlen, const std::string &msg) {
  if (err == nullptr || errlen <= 0) return;
  strncpy(err, msg.c_str(), errlen - 1);
  err[errlen - 1] = '\0';
}

// Runs fbank -> CMVN for one utterance, leaving normalised feats in *feats.
void front_end(Fbank *fbank, const pk_vector_t *global_stats,
               const float *wave, int n, pk_matrix_t *raw, pk_matrix_t *feats) {
  pk_vector_t w;
  w.dim = n;
  w.data = const_cast<float *>(wave);
  fbank->Compute(&w, raw);
  pk_matrix_resize(feats, raw->nrow, raw->ncol);
  if (raw->ncol == 0) return;
  CMVN cmvn(global_stats, raw);
  for (int t = 0; t < raw->ncol; ++t) {
    pk_vector_t col = pk_matrix_getcol(feats, t);
    cmvn.GetFrame(t, &col);
  }
}

}  // namespace

extern "C" {

int ref_fbank_num_frames(int n) {
  return n < 400 ? 0 : 1 + (n - 400) / 160;
}

// out: [T][40] row-major (== the reference's column-major 40 x T). Returns T.
int ref_fbank(const float *wave, int n, float *out, int out_cap_frames) {
  static thread_local Fbank *fbank = nullptr;
  if (fbank == nullptr) fbank = new Fbank();
  pk_vector_t w;
  w.dim = n;
  w.data = const_cast<float *>(wave);
  pk_matrix_t m;
  pk_matrix_init(&m, 0, 0);
  fbank->Compute(&w, &m);
  int T = m.ncol;
  if (T > out_cap_frames) { pk_matrix_destroy(&m); return -1; }
  if (T > 0) memcpy(out, m.data, sizeof(float) * T * m.nrow);
  pk_matrix_destroy(&m);
  return T;
}

// raw/out: [T][40]; global_stats: 41 floats (40 sums + count).
int ref_cmvn(const float *raw, int T, const float *global_stats, float *out) {
  pk_matrix_t m;
  m.nrow = PK_FBANK_DIM;
  m.ncol = T;
  m.data = const_cast<float *>(raw);
  pk_vector_t g;
  g.dim = PK_FBANK_DIM + 1;
  g.data = const_cast<float *>(global_stats);
  CMVN cmvn(&g, &m);
  for (int t = 0; t < T; ++t) {
    pk_vector_t col;
    col.dim = PK_FBANK_DIM;
    col.data = out + t * PK_FBANK_DIM;
    cmvn.GetFrame(t, &col);
  }
  return T;
}

// In-place forward real FFT of n floats (n a power of two >= 4).
int ref_srfft(float *data, int n) {
  pk_srfft_t fft;
  pk_srfft_init(&fft, n);
  std::vector<float> buf(n);
  pk_srfft_compute(&fft, data, n, true, buf.data(), n);
  pk_srfft_destroy(&fft);
  return 0;
}

// C[m x n] = A[m x k] * B[k x n], all row-major, through the packed SGEMM.
int ref_gemm(const float *A, const float *B, float *C, int m, int k, int n) {
  pocketkaldi::SubMatrix<float> a(const_cast<float *>(A), m, k, k);
  pocketkaldi::SubMatrix<float> b(const_cast<float *>(B), k, n, n);
  pocketkaldi::SubMatrix<float> c(C, m, n, n);
  pocketkaldi::GEMM<float> sgemm;
  pocketkaldi::MatMat(a, b, &c, &sgemm);
  return 0;
}

// in: [T][D] row-major; out: [T][out_dim]. Returns out_dim, or <0 on error.
int ref_nnet_propagate(const char *nnet_path, const float *in, int T, int D,
                       float *out, int out_cap_floats, char *err, int errlen) {
  Nnet nnet;
  pocketkaldi::util::ReadableFile fd;
  Status s = fd.Open(nnet_path);
  if (!s.ok()) { set_err(err, errlen, s.what()); return -1; }
  s = nnet.Read(&fd);
  if (!s.ok()) { set_err(err, errlen, s.what()); return -1; }
  pk_matrix_t mi;
  mi.nrow = D;
  mi.ncol = T;
  mi.data = const_cast<float *>(in);
  pk_matrix_t mo;
  pk_matrix_init(&mo, 0, 0);
  nnet.Propagate(&mi, &mo);
  int od = mo.nrow;
  if (od * mo.ncol > out_cap_floats) { pk_matrix_destroy(&mo); return -2; }
  memcpy(out, mo.data, sizeof(float) * od * mo.ncol);
  pk_matrix_destroy(&mo);
  return od;
}

void *ref_am_load(const char *conf_path, char *err, int errlen) {
  Configuration conf;
  Status s = conf.Read(conf_path);
  if (!s.ok()) { set_err(err, errlen, s.what()); return nullptr; }
  AcousticModel *am = new AcousticModel();
  s = am->Read(conf);
  if (!s.ok()) { set_err(err, errlen, s.what()); delete am; return nullptr; }
  return am;
}

void ref_am_free(void *am) { delete static_cast<AcousticModel *>(am); }

int ref_am_num_pdfs(void *am) {
  return static_cast<AcousticModel *>(am)->num_pdfs();
}

int ref_am_tid2pdf(void *am, int tid) {
  return static_cast<AcousticModel *>(am)->TransitionIdToPdfId(tid);
}

// feats: [T][dim]; out: [T][num_pdfs] = log(max(softmax,1e-20)) - log_prior.
int ref_am_compute(void *am_, const float *feats, int T, int dim, float *out) {
  AcousticModel *am = static_cast<AcousticModel *>(am_);
  pk_matrix_t f;
  f.nrow = dim;
  f.ncol = T;
  f.data = const_cast<float *>(feats);
  pk_matrix_t o;
  pk_matrix_init(&o, am->num_pdfs(), T);
  am->Compute(&f, &o);
  memcpy(out, o.data, sizeof(float) * o.nrow * o.ncol);
  pk_matrix_destroy(&o);
  return am->num_pdfs();
}

// out[t][k] = pk_decodable_loglikelihood(frame t, tids[k]); also returns the
// islastframe answer for every frame in last[t].
int ref_decodable_fill(void *am_, float prob_scale, const float *feats, int T,
                       int dim, const int *tids, int n_tids, float *out,
                       uint8_t *last) {
  AcousticModel *am = static_cast<AcousticModel *>(am_);
  pk_matrix_t f;
  f.nrow = dim;
  f.ncol = T;
  f.data = const_cast<float *>(feats);
  pk_decodable_t d;
  pk_decodable_init(&d, am, prob_scale, &f);
  for (int t = 0; t < T; ++t) {
    for (int k = 0; k < n_tids; ++k) {
      out[t * n_tids + k] = pk_decodable_loglikelihood(&d, t, tids[k]);
    }
    last[t] = pk_decodable_islastframe(&d, t) ? 1 : 0;
  }
  pk_decodable_destroy(&d);
  return 0;
}

// Full reference pipeline on one wav: hypothesis string + loglik per frame.
int ref_decode_wav(const char *conf_path, const char *wav_path, char *hyp,
                   int hyplen, float *llpf, char *err, int errlen) {
  pk_t rec;
  pk_status_t st;
  pk_status_init(&st);
  pk_init(&rec);
  pk_load(&rec, conf_path, &st);
  if (!st.ok) { set_err(err, errlen, st.message); return -1; }
  pk_utterance_t utt;
  pk_utterance_init(&utt);
  pk_read_audio(&utt, wav_path, &st);
  if (!st.ok) {
    set_err(err, errlen, st.message);
    pk_utterance_destroy(&utt);
    pk_destroy(&rec);
    return -2;
  }
  pk_process(&rec, &utt);
  set_err(hyp, hyplen, utt.hyp ? utt.hyp : "");
  *llpf = utt.loglikelihood_per_frame;
  pk_utterance_destroy(&utt);
  pk_destroy(&rec);
  return 0;
}

// Reads a wav through the reference reader. Returns number of samples.
int ref_read_wav(const char *wav_path, float *out, int cap) {
  pk_status_t st;
  pk_status_init(&st);
  pk_vector_t v;
  pk_vector_init(&v, 0, NAN);
  pk_16kpcm_read(wav_path, &v, &st);
  if (!st.ok) return -1;
  int n = v.dim;
  if (n > cap) { pk_vector_destroy(&v); return -2; }
  memcpy(out, v.data, sizeof(float) * n);
  pk_vector_destroy(&v);
  return n;
}

// Times the reference hot path (fbank -> CMVN -> AcousticModel::Compute, or
// fbank -> CMVN only when am == NULL) over n_utts utterances of int16 PCM
// (pcm[u * samples_per_utt ...]) with n_threads workers pulling whole
// utterances. Returns wall seconds of the best of `repeats` passes; *frames_out
// is the number of frames one pass produced. checksum_out (optional) receives
// the sum of all outputs so the work cannot be optimised away.
double ref_time_path(void *am_, const float *global_stats, const int16_t *pcm,
                     int n_utts, int samples_per_utt, int n_threads,
                     int repeats, long long *frames_out, double *checksum_out) {
  AcousticModel *am = static_cast<AcousticModel *>(am_);
  pk_vector_t g;
  g.dim = PK_FBANK_DIM + 1;
  g.data = const_cast<float *>(global_stats);
  double best = 1e30;
  long long frames_total = 0;
  double checksum = 0.0;
  for (int rep = 0; rep < repeats; ++rep) {
    std::atomic<int> next(0);
    std::atomic<long long> frames(0);
    std::vector<double> sums(n_threads, 0.0);
    auto worker = [&](int tid) {
      Fbank fbank;
      std::vector<float> wave(samples_per_utt);
      pk_matrix_t raw, feats, ll;
      pk_matrix_init(&raw, 0, 0);
      pk_matrix_init(&feats, 0, 0);
      pk_matrix_init(&ll, 0, 0);
      for (;;) {
        int u = next.fetch_add(1);
        if (u >= n_utts) break;
        const int16_t *src = pcm + (size_t)u * samples_per_utt;
        for (int i = 0; i < samples_per_utt; ++i) wave[i] = src[i];
        front_end(&fbank, &g, wave.data(), samples_per_utt, &raw, &feats);
        frames += feats.ncol;
        double s = 0.0;
        if (am != nullptr && feats.ncol > 0) {
          pk_matrix_resize(&ll, am->num_pdfs(), feats.ncol);
          am->Compute(&feats, &ll);
          for (int i = 0; i < ll.nrow * ll.ncol; i += 97) s += ll.data[i];
        } else {
          for (int i = 0; i < feats.nrow * feats.ncol; i += 7) s += feats.data[i];
        }
        sums[tid] += s;
      }
      pk_matrix_destroy(&raw);
      pk_matrix_destroy(&feats);
      pk_matrix_destroy(&ll);
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; ++i) pool.emplace_back(worker, i);
    for (auto &th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    double sec = std::chrono::duration<double>(t1 - t0).count();
    if (sec < best) best = sec;
    frames_total = frames.load();
    checksum = 0.0;
    for (double s : sums) checksum += s;
  }
  if (frames_out) *frames_out = frames_total;
  if (checksum_out) *checksum_out = checksum;
  return best;
}

}  // extern "C"
