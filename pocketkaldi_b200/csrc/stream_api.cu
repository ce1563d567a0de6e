// Streaming entry points: n_streams audio streams advance in lock step, chunk by chunk, with
// carried state, so that the concatenated outputs equal the whole-utterance outputs of
// Fbank::Compute -> CMVN::GetFrame -> AcousticModel::Compute (src/fbank.cc:267-292,
// src/cmvn.cc:103-115, src/am.cc:90-115). The reference has no streaming API (SURVEY.md
// section 5); the carried state is exactly what its whole-utterance loops keep implicitly:
//   * PCM tail: the samples after the start of the next frame (N - T*160, < 400 of them),
//   * CMVN: the float running sums of ComputeStats (src/cmvn.cc:35-71), the frame counter and a
//     600-frame ring of raw features for the sliding-window subtraction,
//   * splice: the last left+right normalised frames (src/am.cc:65-88); a frame is emitted once
//     its right context exists, pkb_stream_flush replicates the last frame as the reference does
//     at the utterance end.

#include <algorithm>

#include "nnet.cuh"

namespace pkb {
namespace {

// next[s][0..tail) = prev[s][prev_len - tail .. prev_len)
__global__ void stream_tail_kernel(const int16_t *__restrict__ prev, int prev_stride, int prev_len,
                                   int16_t *__restrict__ next, int next_stride, int tail, int n_streams) {
  const int s = blockIdx.y;
  if (s >= n_streams) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x)
    next[static_cast<size_t>(s) * next_stride + i] =
        prev[static_cast<size_t>(s) * prev_stride + prev_len - tail + i];
}

// carry rows: next[s][0..carry) = prev[s][prev_rows - carry .. prev_rows)
__global__ void stream_shift_kernel(const __nv_bfloat16 *__restrict__ prev, int prev_rows,
                                    __nv_bfloat16 *__restrict__ next, int next_rows, int carry,
                                    int dim_pad, int n_streams) {
  const int s = blockIdx.y;
  if (s >= n_streams) return;
  const int n = carry * dim_pad;
  const __nv_bfloat16 *src = prev + (static_cast<size_t>(s) * prev_rows + prev_rows - carry) * dim_pad;
  __nv_bfloat16 *dst = next + static_cast<size_t>(s) * next_rows * dim_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// rows [from, from + count) of every stream := row from - 1 (utterance-end replication)
__global__ void stream_replicate_kernel(__nv_bfloat16 *__restrict__ win, int rows, int from, int count,
                                        int dim_pad, int n_streams) {
  const int s = blockIdx.y;
  if (s >= n_streams) return;
  __nv_bfloat16 *base = win + static_cast<size_t>(s) * rows * dim_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count * dim_pad; i += gridDim.x * blockDim.x)
    base[static_cast<size_t>(from) * dim_pad + i] = base[static_cast<size_t>(from - 1) * dim_pad + (i % dim_pad)];
}

// One thread per (stream, dim): continues the recurrence of cmvn_kernel (cmvn.cu) from the
// carried state for `n_new` frames starting at global frame index t0.
__global__ void cmvn_stream_kernel(const float *__restrict__ raw /* [S][n_new][40] */, int n_new,
                                   int64_t t0, const float *__restrict__ tab, float *__restrict__ stat,
                                   float *__restrict__ ring /* [S][600][40] */,
                                   __nv_bfloat16 *__restrict__ p_hi, __nv_bfloat16 *__restrict__ p_lo,
                                   int win_rows, int carry, int left, int dim_pad, int n_streams,
                                   int fp16) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = g / kMel, d = g % kMel;
  if (s >= n_streams) return;
  const float gd = tab[2 * kCmvnWindow + d];
  float st = stat[s * kMel + d];
  float *rg = ring + static_cast<size_t>(s) * kCmvnWindow * kMel + d;
  __nv_bfloat16 *ph = p_hi + static_cast<size_t>(s) * win_rows * dim_pad + d;
  __nv_bfloat16 *pl = p_lo ? p_lo + static_cast<size_t>(s) * win_rows * dim_pad + d : nullptr;
  for (int k = 0; k < n_new; ++k) {
    const int64_t t = t0 + k;
    const float x = raw[(static_cast<size_t>(s) * n_new + k) * kMel + d];
    double acc = (t > 0 ? static_cast<double>(st) : 0.0) + static_cast<double>(x);
    const int slot = static_cast<int>(t % kCmvnWindow);
    if (t >= kCmvnWindow) acc += -1.0 * static_cast<double>(rg[static_cast<size_t>(slot) * kMel]);
    rg[static_cast<size_t>(slot) * kMel] = x;
    st = static_cast<float>(acc);
    const int ti = t < kCmvnWindow ? static_cast<int>(t) : kCmvnWindow - 1;
    float sm = st;
    if (t < kCmvnWindow - 1) sm = __fadd_rn(sm, __fmul_rn(tab[ti], gd));
    const float v = __fadd_rn(x, __fmul_rn(-tab[kCmvnWindow + ti], sm));
    const __nv_bfloat16 h = operand_bits(v, fp16);
    const __nv_bfloat16 l = operand_bits(v - operand_value(h, fp16), fp16);
    const size_t row = static_cast<size_t>(carry + k) * dim_pad;
    ph[row] = h;
    if (pl) pl[row] = l;
    if (t == 0)  // utterance start: the left context replicates frame 0 (src/am.cc:75)
      for (int r = 0; r < left; ++r) {
        ph[static_cast<size_t>(r) * dim_pad] = h;
        if (pl) pl[static_cast<size_t>(r) * dim_pad] = l;
      }
  }
  stat[s * kMel + d] = st;
}

}  // namespace
}  // namespace pkb

struct pkb_stream {
  pkb::Ctx *c = nullptr;
  pkb_am *am = nullptr;
  int S = 0, C = 0;
  float scale = 1.0f;
  float global[PKB_CMVN_STATS_DIM];
  int L = 0, R = 0, Dp = 0, P = 0;
  int max_new = 0, max_rows = 0, pcm_stride = 0;
  int tail = 0;                   // carried samples per stream
  int prev_pcm_len = 0, prev_pcm_stride = 0;
  int64_t n_feat = 0, n_emit = 0; // frames normalised / emitted so far
  int prev_rows = 0;              // rows per stream of the current feature window
  int pcur = 0, fcur = 0;         // ping-pong indices of the PCM and feature windows
  int meta_len = -1;              // window length the cached fbank metadata was built for
  pkb::DevBuf pcm[2], raw, stat, ring, win_hi[2], win_lo[2], out;
  pkb::Workspace ws;
  pkb::BatchMeta meta;
};

namespace {

// Runs the nnet over the current feature window (rows per stream = `rows`) and copies the first
// `emit` rows of every stream to the host.
int stream_emit(pkb_stream *st, int rows, int emit, float *loglik_out) {
  pkb::Ctx *c = st->c;
  pkb_am *am = st->am;
  if (emit <= 0) return PKB_OK;
  const int64_t gemm_rows = static_cast<int64_t>(st->S) * rows - (st->L + st->R);
  PKB_TRY(pkb::workspace_ensure(am, &st->ws, gemm_rows));
  PKB_TRY(st->out.ensure(static_cast<size_t>(gemm_rows) * st->P * sizeof(float)));
  pkb::InputView in;
  in.hi = st->win_hi[st->fcur].as<__nv_bfloat16>();
  in.lo = am->planes == 2 ? st->win_lo[st->fcur].as<__nv_bfloat16>() : nullptr;
  in.rows = gemm_rows;
  in.cols = (st->L + st->R + 1) * st->Dp;
  in.pitch_elems = st->Dp;
  PKB_TRY(pkb::nnet_forward(am, &st->ws, in, &am->splice_stage, pkb::kFinalLoglik, st->scale,
                            st->out.as<float>()));
  const size_t row_bytes = static_cast<size_t>(st->P) * sizeof(float);
  const int max_frames = pkb_stream_max_frames(st);
  PKB_CUDA(cudaMemcpy2DAsync(loglik_out, max_frames * row_bytes, st->out.p, rows * row_bytes,
                             emit * row_bytes, st->S, cudaMemcpyDeviceToHost, c->stream));
  return PKB_OK;
}

}  // namespace

extern "C" {

int pkb_stream_create(pkb_ctx_t *c, pkb_am_t *am, int n_streams, int chunk_samples,
                      const float *global_stats, float prob_scale, pkb_stream_t **out) {
  PKB_REQUIRE(c && am && out && global_stats, "pkb_stream_create: NULL argument");
  PKB_REQUIRE(am->c == c, "pkb_stream_create: model belongs to another context");
  PKB_REQUIRE(n_streams > 0, "pkb_stream_create: n_streams must be positive");
  PKB_REQUIRE(chunk_samples > 0 && chunk_samples % pkb::kShift == 0,
              "pkb_stream_create: chunk_samples must be a positive multiple of %d", pkb::kShift);
  PKB_REQUIRE(am->has_splice_stage && am->feat_dim == pkb::kMel,
              "pkb_stream_create: the model's feature dim must be %d", pkb::kMel);
  PKB_CUDA(cudaSetDevice(c->device));
  pkb_stream *st = new pkb_stream();
  st->c = c;
  st->am = am;
  st->S = n_streams;
  st->C = chunk_samples;
  st->scale = prob_scale;
  memcpy(st->global, global_stats, sizeof(st->global));
  st->L = am->left;
  st->R = am->right;
  st->Dp = am->feat_dim_pad;
  st->P = am->num_pdfs;
  st->max_new = chunk_samples / pkb::kShift + 1;
  st->max_rows = st->L + st->R + std::max(st->max_new, st->R);  // flush appends R replicated rows
  st->pcm_stride = pkb::kFrame + chunk_samples;
  int rc = PKB_OK;
  do {
    const size_t S = n_streams;
    for (int i = 0; i < 2 && rc == PKB_OK; ++i) {
      rc = st->pcm[i].ensure(S * st->pcm_stride * sizeof(int16_t));
      if (rc == PKB_OK) rc = st->win_hi[i].ensure(S * st->max_rows * st->Dp * 2);
      if (rc == PKB_OK && am->planes == 2) rc = st->win_lo[i].ensure(S * st->max_rows * st->Dp * 2);
    }
    if (rc != PKB_OK) break;
    if ((rc = st->raw.ensure(S * st->max_new * pkb::kMel * sizeof(float))) != PKB_OK) break;
    if ((rc = st->stat.ensure(S * pkb::kMel * sizeof(float))) != PKB_OK) break;
    if ((rc = st->ring.ensure(S * pkb::kCmvnWindow * pkb::kMel * sizeof(float))) != PKB_OK) break;
    cudaMemsetAsync(st->stat.p, 0, S * pkb::kMel * sizeof(float), c->stream);
    for (int i = 0; i < 2; ++i) {
      cudaMemsetAsync(st->win_hi[i].p, 0, S * st->max_rows * st->Dp * 2, c->stream);
      if (am->planes == 2) cudaMemsetAsync(st->win_lo[i].p, 0, S * st->max_rows * st->Dp * 2, c->stream);
    }
    if ((rc = pkb::prepare_cmvn_tables(c, global_stats)) != PKB_OK) break;
  } while (0);
  if (rc != PKB_OK) {
    pkb_stream_destroy(st);
    return rc;
  }
  *out = st;
  return PKB_OK;
}

void pkb_stream_destroy(pkb_stream_t *st) {
  if (!st) return;
  if (st->c) {
    cudaSetDevice(st->c->device);
    cudaStreamSynchronize(st->c->stream);
  }
  for (int i = 0; i < 2; ++i) {
    st->pcm[i].release();
    st->win_hi[i].release();
    st->win_lo[i].release();
  }
  st->raw.release();
  st->stat.release();
  st->ring.release();
  st->out.release();
  st->ws.release();
  st->meta.dev.release();
  delete st;
}

int pkb_stream_max_frames(const pkb_stream_t *st) { return st ? st->max_new + st->R : 0; }

int pkb_stream_push_i16(pkb_stream_t *st, const int16_t *pcm, float *loglik_out, int32_t *frames_out) {
  PKB_REQUIRE(st && pcm, "pkb_stream_push_i16: NULL argument");
  pkb::Ctx *c = st->c;
  pkb_am *am = st->am;
  PKB_CUDA(cudaSetDevice(c->device));
  const int S = st->S, C = st->C;
  const int pn = st->pcur ^ 1;
  // ---- assemble [tail | chunk] per stream
  const int win_len = st->tail + C;
  int16_t *pw = st->pcm[pn].as<int16_t>();
  if (st->tail > 0) {
    pkb::LaunchScope scope(c, PKB_KERNEL_MISC);
    pkb::stream_tail_kernel<<<dim3(2, S), 256, 0, c->stream>>>(
        st->pcm[st->pcur].as<int16_t>(), st->prev_pcm_stride, st->prev_pcm_len, pw, win_len, st->tail, S);
    PKB_CUDA(cudaGetLastError());
  }
  PKB_CUDA(cudaMemcpy2DAsync(pw + st->tail, static_cast<size_t>(win_len) * 2, pcm,
                             static_cast<size_t>(C) * 2, static_cast<size_t>(C) * 2, S,
                             cudaMemcpyHostToDevice, c->stream));
  st->pcur = pn;
  const int n_new = pkb_fbank_num_frames(win_len);
  st->prev_pcm_stride = win_len;
  st->prev_pcm_len = win_len;
  st->tail = win_len - n_new * pkb::kShift;
  int emit = 0;
  if (n_new > 0) {
    // ---- fbank of the new frames (same kernel as the batch path)
    if (st->meta_len != win_len) {
      std::vector<int32_t> ns(S, win_len);
      PKB_TRY(st->meta.build_from_samples(ns.data(), S));
      PKB_TRY(st->meta.upload(c->stream));
      st->meta_len = win_len;
    }
    PKB_TRY(pkb::launch_fbank_i16(c, pw, st->meta, st->raw.as<float>()));
    // ---- feature window: carried context rows + the new frames
    PKB_TRY(pkb::prepare_cmvn_tables(c, st->global));
    const int carry = static_cast<int>(st->n_feat - st->n_emit) + st->L;
    const int rows = carry + n_new;
    PKB_REQUIRE(rows <= st->max_rows, "pkb_stream_push_i16: window overflow (%d rows)", rows);
    const int fn = st->fcur ^ 1;
    __nv_bfloat16 *hi = st->win_hi[fn].as<__nv_bfloat16>();
    __nv_bfloat16 *lo = am->planes == 2 ? st->win_lo[fn].as<__nv_bfloat16>() : nullptr;
    if (st->n_feat > 0) {
      pkb::LaunchScope scope(c, PKB_KERNEL_MISC);
      pkb::stream_shift_kernel<<<dim3(2, S), 256, 0, c->stream>>>(
          st->win_hi[st->fcur].as<__nv_bfloat16>(), st->prev_rows, hi, rows, carry, st->Dp, S);
      if (lo)
        pkb::stream_shift_kernel<<<dim3(2, S), 256, 0, c->stream>>>(
            st->win_lo[st->fcur].as<__nv_bfloat16>(), st->prev_rows, lo, rows, carry, st->Dp, S);
      PKB_CUDA(cudaGetLastError());
    }
    {
      pkb::LaunchScope scope(c, PKB_KERNEL_CMVN);
      const int threads = S * pkb::kMel;
      pkb::cmvn_stream_kernel<<<(threads + 127) / 128, 128, 0, c->stream>>>(
          st->raw.as<float>(), n_new, st->n_feat, c->cmvn_tab.as<float>(), st->stat.as<float>(),
          st->ring.as<float>(), hi, lo, rows, carry, st->L, st->Dp, S, am->fp16);
      PKB_CUDA(cudaGetLastError());
    }
    st->n_feat += n_new;
    st->prev_rows = rows;
    st->fcur = fn;
    emit = std::max(0, rows - (st->L + st->R));
    if (emit > 0) {
      PKB_REQUIRE(loglik_out, "pkb_stream_push_i16: loglik_out is NULL");
      PKB_TRY(stream_emit(st, rows, emit, loglik_out));
      st->n_emit += emit;
    }
  }
  if (frames_out) *frames_out = emit;
  PKB_CUDA(cudaStreamSynchronize(c->stream));
  return PKB_OK;
}

int pkb_stream_flush(pkb_stream_t *st, float *loglik_out, int32_t *frames_out) {
  PKB_REQUIRE(st, "pkb_stream_flush: stream is NULL");
  pkb::Ctx *c = st->c;
  PKB_CUDA(cudaSetDevice(c->device));
  int emit = 0;
  const int pending = static_cast<int>(st->n_feat - st->n_emit);
  if (pending > 0) {
    PKB_REQUIRE(loglik_out, "pkb_stream_flush: loglik_out is NULL");
    // rows = carry (pending + L); append R replicas of the last frame (src/am.cc:76)
    const int carry = pending + st->L;
    const int rows = carry + st->R;
    PKB_REQUIRE(rows <= st->max_rows, "pkb_stream_flush: window overflow");
    // the window is stored with row pitch prev_rows per stream: re-pitch into the other buffer
    const int nxt = st->fcur ^ 1;
    pkb_am *am = st->am;
    __nv_bfloat16 *hi = st->win_hi[nxt].as<__nv_bfloat16>();
    __nv_bfloat16 *lo = am->planes == 2 ? st->win_lo[nxt].as<__nv_bfloat16>() : nullptr;
    {
      pkb::LaunchScope scope(c, PKB_KERNEL_MISC);
      pkb::stream_shift_kernel<<<dim3(2, st->S), 256, 0, c->stream>>>(
          st->win_hi[st->fcur].as<__nv_bfloat16>(), st->prev_rows, hi, rows, carry, st->Dp, st->S);
      pkb::stream_replicate_kernel<<<dim3(1, st->S), 256, 0, c->stream>>>(hi, rows, carry, st->R, st->Dp, st->S);
      if (lo) {
        pkb::stream_shift_kernel<<<dim3(2, st->S), 256, 0, c->stream>>>(
            st->win_lo[st->fcur].as<__nv_bfloat16>(), st->prev_rows, lo, rows, carry, st->Dp, st->S);
        pkb::stream_replicate_kernel<<<dim3(1, st->S), 256, 0, c->stream>>>(lo, rows, carry, st->R, st->Dp, st->S);
      }
      PKB_CUDA(cudaGetLastError());
    }
    st->fcur = nxt;
    st->prev_rows = rows;
    emit = rows - (st->L + st->R);
    PKB_TRY(stream_emit(st, rows, emit, loglik_out));
    st->n_emit += emit;
  }
  if (frames_out) *frames_out = emit;
  PKB_CUDA(cudaStreamSynchronize(c->stream));
  // reset for a new utterance
  st->tail = 0;
  st->n_feat = st->n_emit = 0;
  st->prev_rows = 0;
  PKB_CUDA(cudaMemsetAsync(st->stat.p, 0, static_cast<size_t>(st->S) * pkb::kMel * sizeof(float), c->stream));
  return PKB_OK;
}

}  // extern "C"
