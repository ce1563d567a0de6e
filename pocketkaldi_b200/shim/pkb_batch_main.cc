// pocketkaldi_b200_batch <model-file> <list.scp | x.wav> [--threads N] [--batch-utts M] [--compact 0|1]
//                        [--gpu-decode 0|1]
//
// Batch counterpart of the reference CLI (src/main.cc): same arguments, same
// "<file>\t<hyp>\t<loglikelihood per frame>" lines in list order, but
//   * the wave files of the list are validated up front and read by several threads as int16
//     straight into page-locked staging (the reference reads one file at a time
//     into a float vector, src/main.cc:34-46 + src/pcm_reader.cc:45-220),
//   * fbank -> CMVN -> nnet run once per sub-batch on the GPU (pkb_batch_*), and
//   * the reference's own Viterbi decoder (src/decoder.cc, unchanged) runs on a pool of host
//     threads, each utterance starting as soon as ITS rows have reached host memory while the
//     copy-out of later utterances is still in flight (pk_decodable_attach), or
//   * with --gpu-decode 1 the search itself runs on the GPU (pkb_batch_decode): the
//     log-likelihoods never leave HBM and only word ids come back.
// Model loading is the reference's pk_load compiled against the shim header.

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "configuration.h"
#include "decoder.h"
#include "pkb200.h"
#include "pocketkaldi.h"
#include "symbol_table.h"

using pocketkaldi::Decoder;

namespace {

// PKB_CLI_TIMING=1: wall-clock milestones on stderr
double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
void milestone(const char *what) {
  static const bool on = getenv("PKB_CLI_TIMING") != nullptr;
  static double t0 = now_s(), last = t0;
  if (!on) return;
  const double t = now_s();
  fprintf(stderr, "[pkb batch] %-28s +%.3f s (%.3f s)\n", what, t - last, t - t0);
  last = t;
}

void die(const char *what) {
  printf("pocketkaldi: %s: %s\n", what, pkb_last_error());
  exit(1);
}
#define CHECK(expr) do { if ((expr) != PKB_OK) die(#expr); } while (0)

struct Result {
  std::string hyp;
  float llpf = 0.0f;
};

// The decode half of pk_process (src/pocketkaldi.cc:209-243) over an attached decodable.
Result decode_one(pk_t *rec, float *rows, const uint16_t *rows16, const float *off, int frames,
                  pkb_event_t *ready) {
  Result r;
  Decoder decoder(rec->fst);
  pk_decodable_t dec;
  if (rows16 != nullptr) pk_decodable_attach_compact(&dec, rec->am, rows16, off, frames, 0.1f, ready);
  else pk_decodable_attach(&dec, rec->am, rows, frames, ready);
  decoder.Decode(&dec);
  Decoder::Hypothesis hyp = decoder.BestPath();
  std::vector<int> words = hyp.words();
  std::reverse(words.begin(), words.end());
  for (size_t i = 0; i < words.size(); ++i) {
    if (i != 0) r.hyp += ' ';
    r.hyp += pk_symboltable_get(rec->symbol_table, words[i]);
  }
  if (!words.empty()) r.llpf = hyp.weight() / frames;
  pk_decodable_destroy(&dec);
  return r;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    puts("Usage: pocketkaldi_b200_batch <model-file> <input-file> [--threads N] [--batch-utts M]");
    puts("  Input-file:");
    puts("    *.wav: decode this file.");
    puts("    *.scp: decode audios listed in it.");
    return 1;
  }
  int n_threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  int batch_utts = 256;
  // compact rows (half the PCIe bytes, finished per look-up) are the default; --compact 0 moves
  // the FP32 matrix like pk_decodable_init does
  bool compact = true;
  bool gpu_decode = false;
  for (int i = 3; i + 1 < argc; i += 2) {
    if (strcmp(argv[i], "--threads") == 0) n_threads = std::max(1, atoi(argv[i + 1]));
    else if (strcmp(argv[i], "--batch-utts") == 0) batch_utts = std::max(1, atoi(argv[i + 1]));
    else if (strcmp(argv[i], "--compact") == 0) compact = atoi(argv[i + 1]) != 0;
    else if (strcmp(argv[i], "--gpu-decode") == 0) gpu_decode = atoi(argv[i + 1]) != 0;
  }
  const char *model_file = argv[1], *input_file = argv[2];

  milestone("start");
  pk_t rec;
  pk_status_t status;
  pk_status_init(&status);
  pk_init(&rec);
  pk_load(&rec, model_file, &status);
  if (!status.ok) {
    printf("pocketkaldi: %s\n", status.message);
    return 1;
  }
  milestone("pk_load");
  pkb_ctx_t *ctx = pkb_shim_context();
  pkb_am_t *am = rec.am->handle();
  const int pdfs = rec.am->num_pdfs();

  pkb_fst_t *gpu_fst = nullptr;
  if (gpu_decode) {
    pocketkaldi::Configuration conf;
    pocketkaldi::Status st = conf.Read(model_file);
    const std::string fst_path = st.ok() ? conf.GetPathOrElse("fst", "") : std::string();
    if (fst_path.empty()) {
      printf("pocketkaldi: Unable to find key 'fst' in %s\n", model_file);
      return 1;
    }
    CHECK(pkb_fst_load(ctx, fst_path.c_str(), &gpu_fst));
  }

  pkb_wavlist_t *list = nullptr;
  const size_t len = strlen(input_file);
  if (len >= 4 && strcmp(input_file + len - 4, ".wav") == 0) {
    CHECK(pkb_wavlist_create(&input_file, 1, &list));
  } else {
    CHECK(pkb_scp_open(input_file, &list));
  }
  milestone("fst + list headers");
  const int n_files = pkb_wavlist_size(list);
  const int32_t *num_samples = pkb_wavlist_num_samples(list);

  std::vector<Result> results(n_files);
  for (int first = 0; first < n_files; first += batch_utts) {
    const int n = std::min(batch_utts, n_files - first);
    std::vector<int64_t> frame_off(n + 1, 0);
    int64_t samples = 0;
    for (int u = 0; u < n; ++u) {
      frame_off[u + 1] = frame_off[u] + pkb_fbank_num_frames(num_samples[first + u]);
      samples += num_samples[first + u];
    }
    const int64_t frames = frame_off[n];
    pkb_batch_t *batch = nullptr;
    CHECK(pkb_batch_create(ctx, am, n, num_samples + first, rec.cmvn_global_stats->data, 0.1f, &batch));
    if (gpu_decode) {
      // the whole search on the device: only the word ids and one weight per utterance come back
      void *pcm_g = nullptr;
      CHECK(pkb_host_alloc(&pcm_g, std::max<int64_t>(samples, 1) * sizeof(int16_t)));
      CHECK(pkb_wavlist_read_i16(list, first, n, static_cast<int16_t *>(pcm_g), n_threads));
      CHECK(pkb_batch_set_pcm_i16(batch, static_cast<const int16_t *>(pcm_g)));
      CHECK(pkb_batch_run(batch, PKB_STAGE_ALL | PKB_STAGE_NO_FEATS));
      milestone("read wavs + queue batch");
      const int max_words = 1024;
      std::vector<int32_t> words(static_cast<size_t>(n) * max_words), n_words(n);
      std::vector<float> weight(n);
      CHECK(pkb_batch_decode(batch, gpu_fst, 0.0f, 0, max_words, words.data(), n_words.data(), weight.data()));
      milestone("acoustic + GPU Viterbi");
      for (int u = 0; u < n; ++u) {
        if (num_samples[first + u] == 0) continue;
        if (n_words[u] < 0) {
          printf("pocketkaldi: %s: the GPU search ran out of its token capacity (code %d)\n",
                 pkb_wavlist_path(list, first + u), n_words[u]);
          return 1;
        }
        Result &r = results[first + u];
        for (int i = 0; i < std::min(n_words[u], max_words); ++i) {
          if (i != 0) r.hyp += ' ';
          r.hyp += pk_symboltable_get(rec.symbol_table, words[static_cast<size_t>(u) * max_words + i]);
        }
        const int64_t T = frame_off[u + 1] - frame_off[u];
        if (n_words[u] > 0 && T > 0) r.llpf = weight[u] / T;
      }
      pkb_host_free(pcm_g);
      pkb_batch_destroy(batch);
      continue;
    }
    void *pcm = nullptr, *ll = nullptr, *off = nullptr;
    if (compact) CHECK(pkb_batch_set_compact(batch, 1));
    CHECK(pkb_host_alloc(&pcm, std::max<int64_t>(samples, 1) * sizeof(int16_t)));
    CHECK(pkb_host_alloc(&ll, std::max<int64_t>(frames, 1) * pdfs * (compact ? sizeof(uint16_t) : sizeof(float))));
    CHECK(pkb_host_alloc(&off, std::max<int64_t>(frames, 1) * sizeof(float)));
    CHECK(pkb_wavlist_read_i16(list, first, n, static_cast<int16_t *>(pcm), n_threads));
    CHECK(pkb_batch_set_pcm_i16(batch, static_cast<const int16_t *>(pcm)));
    CHECK(pkb_batch_run(batch, PKB_STAGE_ALL));
    std::vector<pkb_event_t *> ready(n, nullptr);
    for (int u = 0; u < n; ++u) {
      const int64_t T = frame_off[u + 1] - frame_off[u];
      if (T > 0 && compact) {
        CHECK(pkb_batch_get_rows(batch, PKB_BUF_LOGLIK16, frame_off[u], T,
                                 static_cast<uint16_t *>(ll) + frame_off[u] * pdfs));
        CHECK(pkb_batch_get_rows(batch, PKB_BUF_LOGLIK_OFF, frame_off[u], T,
                                 static_cast<float *>(off) + frame_off[u]));
      } else if (T > 0) {
        CHECK(pkb_batch_get_rows(batch, PKB_BUF_LOGLIK, frame_off[u], T,
                                 static_cast<float *>(ll) + frame_off[u] * pdfs));
      }
      CHECK(pkb_event_create(ctx, &ready[u]));
      CHECK(pkb_event_record(ctx, ready[u]));
    }
    std::atomic<int> next(0);
    auto work = [&]() {
      for (;;) {
        const int u = next.fetch_add(1);
        if (u >= n) return;
        if (num_samples[first + u] == 0) continue;  // pk_process: empty utterance -> empty hyp
        results[first + u] = decode_one(
            &rec, compact ? nullptr : static_cast<float *>(ll) + frame_off[u] * pdfs,
            compact ? static_cast<uint16_t *>(ll) + frame_off[u] * pdfs : nullptr,
            static_cast<float *>(off) + frame_off[u], static_cast<int>(frame_off[u + 1] - frame_off[u]),
            ready[u]);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::min(n_threads, n); ++t) pool.emplace_back(work);
    milestone("read wavs + queue batch");
    for (auto &t : pool) t.join();
    milestone("host decoder threads");
    CHECK(pkb_sync(ctx));
    for (pkb_event_t *e : ready) pkb_event_destroy(e);
    pkb_host_free(pcm);
    pkb_host_free(ll);
    pkb_host_free(off);
    pkb_batch_destroy(batch);
  }
  milestone("batches done");
  for (int i = 0; i < n_files; ++i)
    printf("%s\t%s\t%f\n", pkb_wavlist_path(list, i), results[i].hyp.c_str(), results[i].llpf);
  pkb_wavlist_destroy(list);
  pkb_fst_destroy(gpu_fst);
  pk_destroy(&rec);
  return 0;
}
