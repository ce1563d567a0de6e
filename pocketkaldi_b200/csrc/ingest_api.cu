// Batched ingestion: strict WAV header validation + int16 staging, and .scp lists read by a
// pool of host threads. Host code only (no kernel, no CUDA call).
//
// Mirrors pk_16kpcm_read (src/pcm_reader.cc:45-220): the same checks in the same order with the
// same message texts, so a file the reference rejects is rejected here for the same reason.
// List handling follows process_scp / pk_readable_readline (src/main.cc:34-46,
// src/util.cc:130-160).

#include <stdio.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

struct pkb_wavlist {
  std::vector<std::string> paths;
  std::vector<int32_t> num_samples;
  std::vector<int32_t> bits;
};

namespace pkb {
namespace {

constexpr int kWavHeaderBytes = 44;

struct WavInfo {
  int32_t num_samples = 0;
  int32_t bits = 0;
};

int32_t le32(const unsigned char *p) {
  return static_cast<int32_t>(static_cast<uint32_t>(p[0]) | static_cast<uint32_t>(p[1]) << 8 |
                              static_cast<uint32_t>(p[2]) << 16 | static_cast<uint32_t>(p[3]) << 24);
}
int16_t le16(const unsigned char *p) {
  return static_cast<int16_t>(static_cast<uint16_t>(p[0]) | static_cast<uint16_t>(p[1]) << 8);
}

struct File {
  FILE *fp = nullptr;
  ~File() {
    if (fp) fclose(fp);
  }
};

#define WAV_CORRUPT(...)          \
  do {                            \
    set_error(__VA_ARGS__);       \
    return PKB_ERR_CORRUPT;       \
  } while (0)

// Opens `path`, validates the canonical header and leaves the file positioned on the samples.
int open_wav(const char *path, File &f, WavInfo &info) {
  f.fp = fopen(path, "rb");
  if (f.fp == nullptr) {
    set_error("unable to open: %s", path);  // src/util.cc:75
    return PKB_ERR_IO;
  }
  fseek(f.fp, 0, SEEK_END);
  const long long file_size = ftell(f.fp);
  fseek(f.fp, 0, SEEK_SET);
  unsigned char h[kWavHeaderBytes];
  if (fread(h, 1, kWavHeaderBytes, f.fp) != kWavHeaderBytes) {
    set_error("failed to read: %s", path);
    return PKB_ERR_IO;
  }
  if (memcmp(h, "RIFF", 4) != 0) WAV_CORRUPT("chunk_name == 'RIFF' expected: %s", path);
  const long long chunk_size = le32(h + 4);
  if (chunk_size != file_size - 8)
    WAV_CORRUPT("chunk_size == %lld expected, but %lld found: %s", file_size - 8, chunk_size, path);
  if (memcmp(h + 8, "WAVE", 4) != 0) WAV_CORRUPT("Format == 'WAVE' expected: %s", path);
  if (memcmp(h + 12, "fmt ", 4) != 0) WAV_CORRUPT("subchunk1 == 'fmt ' expected: %s", path);
  const int subchunk1_size = le32(h + 16);
  if (subchunk1_size != 16)
    WAV_CORRUPT("subchunk1_size == 16 expected, but %d found: %s", subchunk1_size, path);
  const int audio_format = le16(h + 20);
  if (audio_format != 1)
    WAV_CORRUPT("audio_format == 1 (PCM) expected, but %d found: %s", audio_format, path);
  const int num_channels = le16(h + 22);
  if (num_channels != 1)
    WAV_CORRUPT("num_channels == 1 (mono) expected, but %d found: %s", num_channels, path);
  const int sample_rate = le32(h + 24);
  if (sample_rate != kSampleRate)
    WAV_CORRUPT("sample_rate == 16000 expected, but %d found: %s", sample_rate, path);
  const int bytes_rate = le32(h + 28);
  const int block_align = le16(h + 32);
  const int bits = le16(h + 34);
  if (bytes_rate != sample_rate * bits / 8)
    WAV_CORRUPT("bytes_rate == %d expected, but %d found: %s", sample_rate * bits / 8, bytes_rate, path);
  if (block_align != bits / 8)
    WAV_CORRUPT("block_align == %d expected, but %d found: %s", bits / 8, block_align, path);
  if (memcmp(h + 36, "data", 4) != 0) WAV_CORRUPT("subchunk2 == 'data' expected: %s", path);
  const long long subchunk2_size = le32(h + 40);
  if (subchunk2_size != file_size - kWavHeaderBytes)
    WAV_CORRUPT("subchunk2_size == %lld expected, but %lld found: %s", file_size - kWavHeaderBytes,
                subchunk2_size, path);
  if (bits != 8 && bits != 16 && bits != 32)
    WAV_CORRUPT("bits_per_sample == 8, 16 or 32 expected, but %d found: %s", bits, path);
  info.bits = bits;
  info.num_samples = static_cast<int32_t>(subchunk2_size / (bits / 8));
  return PKB_OK;
}

int read_exact(File &f, void *dst, size_t bytes, const char *path) {
  if (bytes != 0 && fread(dst, 1, bytes, f.fp) != bytes) {
    set_error("failed to read: %s", path);
    return PKB_ERR_IO;
  }
  return PKB_OK;
}

// T = int16_t or float. Samples are copied unscaled; 8-bit samples are read as signed bytes
// like the reference does (src/pcm_reader.cc:194-196).
template <typename T>
int read_samples(const char *path, T *dst, int32_t capacity, int32_t *num_samples) {
  File f;
  WavInfo info;
  PKB_TRY(open_wav(path, f, info));
  if (num_samples) *num_samples = info.num_samples;
  if (sizeof(T) == sizeof(int16_t) && info.bits == 32) {
    set_error("32-bit samples do not fit the int16 path (use pkb_wav_read_f32): %s", path);
    return PKB_ERR_UNSUPPORTED;
  }
  PKB_REQUIRE(dst != nullptr, "read_wav: null destination");
  PKB_REQUIRE(capacity >= info.num_samples, "read_wav: %d samples do not fit a buffer of %d: %s",
              info.num_samples, capacity, path);
  const size_t n = static_cast<size_t>(info.num_samples);
  if (info.bits == 16 && sizeof(T) == sizeof(int16_t)) {
    return read_exact(f, dst, n * 2, path);  // little-endian host: the file bytes are the samples
  }
  std::vector<unsigned char> raw(n * (info.bits / 8));
  PKB_TRY(read_exact(f, raw.data(), raw.size(), path));
  for (size_t i = 0; i < n; ++i) {
    if (info.bits == 8) dst[i] = static_cast<T>(static_cast<int8_t>(raw[i]));
    else if (info.bits == 16) dst[i] = static_cast<T>(le16(&raw[2 * i]));
    else dst[i] = static_cast<T>(le32(&raw[4 * i]));
  }
  return PKB_OK;
}

int build_list(std::vector<std::string> &&paths, pkb_wavlist_t **out) {
  std::unique_ptr<pkb_wavlist> l(new pkb_wavlist);
  l->paths = std::move(paths);
  l->num_samples.resize(l->paths.size());
  l->bits.resize(l->paths.size());
  for (size_t i = 0; i < l->paths.size(); ++i) {
    File f;
    WavInfo info;
    PKB_TRY(open_wav(l->paths[i].c_str(), f, info));
    l->num_samples[i] = info.num_samples;
    l->bits[i] = info.bits;
  }
  *out = l.release();
  return PKB_OK;
}

}  // namespace
}  // namespace pkb

using namespace pkb;

extern "C" {

int pkb_wav_probe(const char *path, int32_t *num_samples, int32_t *bits_per_sample) {
  PKB_REQUIRE(path != nullptr, "pkb_wav_probe: null path");
  File f;
  WavInfo info;
  PKB_TRY(open_wav(path, f, info));
  if (num_samples) *num_samples = info.num_samples;
  if (bits_per_sample) *bits_per_sample = info.bits;
  return PKB_OK;
}

int pkb_wav_read_i16(const char *path, int16_t *dst, int32_t capacity, int32_t *num_samples) {
  PKB_REQUIRE(path != nullptr, "pkb_wav_read_i16: null path");
  return read_samples<int16_t>(path, dst, capacity, num_samples);
}

int pkb_wav_read_f32(const char *path, float *dst, int32_t capacity, int32_t *num_samples) {
  PKB_REQUIRE(path != nullptr, "pkb_wav_read_f32: null path");
  return read_samples<float>(path, dst, capacity, num_samples);
}

int pkb_scp_open(const char *scp_path, pkb_wavlist_t **list) {
  PKB_REQUIRE(scp_path != nullptr && list != nullptr, "pkb_scp_open: null argument");
  *list = nullptr;
  File f;
  f.fp = fopen(scp_path, "rb");
  if (f.fp == nullptr) {
    set_error("unable to open: %s", scp_path);
    return PKB_ERR_IO;
  }
  std::vector<std::string> paths;
  char line[2048];  // src/main.cc:41
  while (fgets(line, sizeof(line), f.fp) != nullptr) {
    size_t n = strlen(line);
    while (n > 0 && (line[n - 1] == '\r' || line[n - 1] == '\n')) line[--n] = '\0';
    paths.emplace_back(line);
  }
  if (ferror(f.fp)) {
    set_error("%s", scp_path);
    return PKB_ERR_IO;
  }
  return build_list(std::move(paths), list);
}

int pkb_wavlist_create(const char *const *paths, int n_paths, pkb_wavlist_t **list) {
  PKB_REQUIRE(list != nullptr && n_paths >= 0 && (paths != nullptr || n_paths == 0),
              "pkb_wavlist_create: bad arguments");
  *list = nullptr;
  std::vector<std::string> v;
  for (int i = 0; i < n_paths; ++i) {
    PKB_REQUIRE(paths[i] != nullptr, "pkb_wavlist_create: null path %d", i);
    v.emplace_back(paths[i]);
  }
  return build_list(std::move(v), list);
}

void pkb_wavlist_destroy(pkb_wavlist_t *list) { delete list; }

int pkb_wavlist_size(const pkb_wavlist_t *list) { return list ? static_cast<int>(list->paths.size()) : 0; }

const char *pkb_wavlist_path(const pkb_wavlist_t *list, int i) {
  if (list == nullptr || i < 0 || i >= static_cast<int>(list->paths.size())) return nullptr;
  return list->paths[i].c_str();
}

const int32_t *pkb_wavlist_num_samples(const pkb_wavlist_t *list) {
  return list ? list->num_samples.data() : nullptr;
}

int pkb_wavlist_read_i16(const pkb_wavlist_t *list, int first, int count, int16_t *dst, int n_threads) {
  PKB_REQUIRE(list != nullptr, "pkb_wavlist_read_i16: null list");
  const int n = static_cast<int>(list->paths.size());
  PKB_REQUIRE(first >= 0 && count >= 0 && first + count <= n,
              "pkb_wavlist_read_i16: range [%d, %d) outside the %d files of the list", first,
              first + count, n);
  if (count == 0) return PKB_OK;
  PKB_REQUIRE(dst != nullptr, "pkb_wavlist_read_i16: null destination");
  std::vector<int64_t> off(count);
  int64_t total = 0;
  for (int i = 0; i < count; ++i) {
    off[i] = total;
    total += list->num_samples[first + i];
  }
  if (n_threads <= 0) n_threads = static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
  n_threads = std::min(n_threads, count);
  // set_error() keeps one message per thread; the first failing worker's message and code are
  // handed back to the calling thread.
  std::atomic<int> next(0);
  std::mutex mu;
  int rc = PKB_OK;
  std::string msg;
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= count) return;
      int32_t got = 0;
      const int r = read_samples<int16_t>(list->paths[first + i].c_str(), dst + off[i],
                                          list->num_samples[first + i], &got);
      if (r != PKB_OK || got != list->num_samples[first + i]) {
        std::lock_guard<std::mutex> g(mu);
        if (rc == PKB_OK) {
          rc = r != PKB_OK ? r : PKB_ERR_IO;
          msg = r != PKB_OK ? pkb_last_error()
                            : std::string("file changed size since the list was opened: ") + list->paths[first + i];
        }
        return;
      }
    }
  };
  if (n_threads == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  if (rc != PKB_OK) set_error("%s", msg.c_str());
  return rc;
}

}  // extern "C"
