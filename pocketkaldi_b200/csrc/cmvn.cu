// Online sliding-window cepstral mean normalisation for sm_100a, plus the
// synthetic-PCM generator and the checksum reduction used by the batch pipeline.
//
// Replaces CMVN::GetFrame called for t = 0..T-1 (src/cmvn.cc:103-115):
//   ComputeStats (:35-71)  stats_t = float( double(stats_{t-1}) + x_t - x_{t-600} )
//   SmoothStats  (:73-92)  stats += float(min(600-n,200)/Gc) * G     (float mul, float add)
//   Apply        (:94-101) y = x + (-float(1/count)) * stats         (float mul, float add)
// The running sum is a float that is widened to double only inside a step, so
// the recurrence is inherently sequential per (utterance, dimension). The
// kernel keeps that exact rounding sequence -- one thread per chain -- which
// makes the output bit-identical to the reference for identical raw input;
// parallelism comes from the n_utts * 40 independent chains. The frame-count
// element of the stats vector depends only on t, so alpha_t and scale_t are
// tabulated on the host with the reference's double/float arithmetic.

#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace pkb {

namespace {

constexpr int kCmvnUnroll = 8;

__global__ void __launch_bounds__(160)
cmvn_kernel(const float *__restrict__ raw, const int64_t *__restrict__ frame_off,
            const int32_t *__restrict__ num_frames, int n_utts,
            const float *__restrict__ tab /* alpha[600], scale[600], global[41] */,
            float *__restrict__ out, __nv_bfloat16 *__restrict__ p_hi,
            __nv_bfloat16 *__restrict__ p_lo, const int64_t *__restrict__ pad_off, int left,
            int right, int dim_pad, int fp16) {
  __shared__ float s_alpha[kCmvnWindow];
  __shared__ float s_scale[kCmvnWindow];
  for (int i = threadIdx.x; i < kCmvnWindow; i += blockDim.x) {
    s_alpha[i] = tab[i];
    s_scale[i] = tab[kCmvnWindow + i];
  }
  __syncthreads();
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int u = static_cast<int>(g / kMel);
  const int d = static_cast<int>(g % kMel);
  if (u >= n_utts) return;
  const int T = num_frames[u];
  if (T == 0) return;
  const float gd = tab[2 * kCmvnWindow + d];
  const float *x = raw + frame_off[u] * kMel + d;
  float *y = out ? out + frame_off[u] * kMel + d : nullptr;
  __nv_bfloat16 *ph = p_hi ? p_hi + pad_off[u] * dim_pad + d : nullptr;
  __nv_bfloat16 *pl = p_lo ? p_lo + pad_off[u] * dim_pad + d : nullptr;

  // The chain itself is ~4 dependent operations per frame; what has to be hidden is the load
  // latency with only a few warps per SM at small batch sizes. The loads of the next group of
  // kCmvnUnroll frames are therefore issued before the current group is consumed (two register
  // sets, software pipelined).
  float stat = 0.0f;
  float xa[kCmvnUnroll], pa[kCmvnUnroll], xb[kCmvnUnroll], pb[kCmvnUnroll];
  auto load = [&](int t0, float (&xv)[kCmvnUnroll], float (&xp)[kCmvnUnroll]) {
#pragma unroll
    for (int i = 0; i < kCmvnUnroll; ++i) {
      const int t = t0 + i;
      xv[i] = t < T ? x[static_cast<int64_t>(t) * kMel] : 0.0f;
      xp[i] = (t < T && t >= kCmvnWindow) ? x[static_cast<int64_t>(t - kCmvnWindow) * kMel] : 0.0f;
    }
  };
  auto consume = [&](int t0, const float (&xv)[kCmvnUnroll], const float (&xp)[kCmvnUnroll]) {
#pragma unroll
    for (int i = 0; i < kCmvnUnroll; ++i) {
      const int t = t0 + i;
      if (t < T) {
        double acc = static_cast<double>(stat) + static_cast<double>(xv[i]);
        if (t >= kCmvnWindow) acc += -1.0 * static_cast<double>(xp[i]);
        stat = static_cast<float>(acc);
        const int ti = t < kCmvnWindow ? t : kCmvnWindow - 1;
        float s = stat;
        if (t < kCmvnWindow - 1) s = __fadd_rn(s, __fmul_rn(s_alpha[ti], gd));
        const float v = __fadd_rn(xv[i], __fmul_rn(-s_scale[ti], s));
        if (y) y[static_cast<int64_t>(t) * kMel] = v;
        if (ph) {
          const __nv_bfloat16 h = operand_bits(v, fp16);
          const __nv_bfloat16 l = operand_bits(v - operand_value(h, fp16), fp16);
          const int64_t row = left + t;
          ph[row * dim_pad] = h;
          if (pl) pl[row * dim_pad] = l;
          if (t == 0)
            for (int r = 0; r < left; ++r) {
              ph[static_cast<int64_t>(r) * dim_pad] = h;
              if (pl) pl[static_cast<int64_t>(r) * dim_pad] = l;
            }
          if (t == T - 1)
            for (int r = 0; r < right; ++r) {
              ph[(row + 1 + r) * dim_pad] = h;
              if (pl) pl[(row + 1 + r) * dim_pad] = l;
            }
        }
      }
    }
  };
  load(0, xa, pa);
  for (int t0 = 0; t0 < T; t0 += 2 * kCmvnUnroll) {
    load(t0 + kCmvnUnroll, xb, pb);
    consume(t0, xa, pa);
    load(t0 + 2 * kCmvnUnroll, xa, pa);
    consume(t0 + kCmvnUnroll, xb, pb);
  }
}

// splitmix64 finaliser; must match pocketkaldi_b200/synth.py
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void synth_pcm_kernel(int16_t *__restrict__ pcm, const int64_t *__restrict__ sample_off,
                                 const int32_t *__restrict__ num_samples, int n_utts, uint64_t seed,
                                 uint64_t first_utt) {
  const int u = blockIdx.y;
  if (u >= n_utts) return;
  const uint64_t key = mix64(seed * 0x9E3779B97F4A7C15ull + (first_utt + u));
  const int n = num_samples[u];
  int16_t *dst = pcm + sample_off[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t h = mix64(key + static_cast<uint64_t>(i) * 0x9E3779B97F4A7C15ull);
    const int64_t s = static_cast<int64_t>((h & 0xffff) + ((h >> 16) & 0xffff) +
                                           ((h >> 32) & 0xffff) + (h >> 48)) - 2 * 65535;
    dst[i] = static_cast<int16_t>((s * 5196) >> 16);
  }
}

__global__ void checksum_kernel(const float *__restrict__ x, int64_t n, double *sum) {
  double acc = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    acc += static_cast<double>(x[i]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
    atomicAdd(sum, t);
  }
}

}  // namespace

// Host restatement of the count-dependent part of SmoothStats / Apply
// (src/cmvn.cc:73-101): for n = min(t+1, 600) frames in the window,
// alpha = float(min(600-n, 200) / Gc) (0 once n == 600), count' = n + alpha*Gc
// in float, scale = float(1 / double(count')).
int prepare_cmvn_tables(Ctx *c, const float *global_stats) {
  if (c->cmvn_valid && memcmp(c->cmvn_global, global_stats, sizeof(c->cmvn_global)) == 0)
    return PKB_OK;
  PKB_REQUIRE(global_stats[kMel] > 0.0f, "cmvn: global frame count must be positive");
  std::vector<float> tab(2 * kCmvnWindow + PKB_CMVN_STATS_DIM);
  const float gc = global_stats[kMel];
  for (int t = 0; t < kCmvnWindow; ++t) {
    float count_f = static_cast<float>(t + 1);
    double count = count_f;
    float alpha = 0.0f;
    if (count < kCmvnWindow) {
      double from_global = kCmvnWindow - count;
      double global_count = gc;
      if (from_global > kCmvnGlobal) from_global = kCmvnGlobal;
      alpha = static_cast<float>(from_global / global_count);
      volatile float prod = alpha * gc;
      count_f = count_f + prod;
    }
    double cnt = count_f;
    tab[t] = alpha;
    tab[kCmvnWindow + t] = static_cast<float>(1 / cnt);
  }
  memcpy(&tab[2 * kCmvnWindow], global_stats, sizeof(float) * PKB_CMVN_STATS_DIM);
  PKB_TRY(c->cmvn_tab.ensure(tab.size() * sizeof(float)));
  PKB_CUDA(cudaMemcpyAsync(c->cmvn_tab.p, tab.data(), tab.size() * sizeof(float),
                           cudaMemcpyHostToDevice, c->stream));
  PKB_CUDA(cudaStreamSynchronize(c->stream));  // tab is a stack-scoped host vector
  memcpy(c->cmvn_global, global_stats, sizeof(c->cmvn_global));
  c->cmvn_valid = true;
  return PKB_OK;
}

int launch_cmvn(Ctx *c, const float *d_raw, const BatchMeta &m, float *d_out,
                const PaddedPlanes *planes) {
  if (m.n_utts == 0 || m.total_frames == 0) return PKB_OK;
  PKB_REQUIRE(c->cmvn_valid, "cmvn: tables not prepared");
  const int64_t threads = static_cast<int64_t>(m.n_utts) * kMel;
  const int block = 160;
  const int grid = static_cast<int>((threads + block - 1) / block);
  LaunchScope scope(c, PKB_KERNEL_CMVN);
  cmvn_kernel<<<grid, block, 0, c->stream>>>(
      d_raw, m.d_frame_off, m.d_num_frames, m.n_utts, c->cmvn_tab.as<float>(), d_out,
      planes ? planes->hi : nullptr, planes ? planes->lo : nullptr,
      planes ? planes->d_pad_off : nullptr, planes ? planes->left : 0,
      planes ? planes->right : 0, planes ? planes->dim_pad : 0, planes ? planes->fp16 : 0);
  PKB_CUDA(cudaGetLastError());
  return PKB_OK;
}

int launch_synth_pcm(Ctx *c, int16_t *d_pcm, const BatchMeta &m, uint64_t seed,
                     uint64_t first_utt) {
  if (m.n_utts == 0) return PKB_OK;
  int max_n = 0;
  for (int32_t n : m.num_samples) max_n = std::max(max_n, n);
  if (max_n == 0) return PKB_OK;
  // grid.y is limited to 65535: loop over slabs of utterances
  const int bx = std::min((max_n + 255) / 256, 64);
  for (int u0 = 0; u0 < m.n_utts; u0 += 32768) {
    const int nu = std::min(32768, m.n_utts - u0);
    LaunchScope scope(c, PKB_KERNEL_MISC);
    synth_pcm_kernel<<<dim3(bx, nu), 256, 0, c->stream>>>(d_pcm, m.d_sample_off + u0,
                                                          m.d_num_samples + u0, nu, seed,
                                                          first_utt + u0);
    PKB_CUDA(cudaGetLastError());
  }
  return PKB_OK;
}

int launch_checksum(Ctx *c, const float *d, int64_t n, double *d_sum) {
  PKB_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), c->stream));
  if (n == 0) return PKB_OK;
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, c->sm_count * 8));
  LaunchScope scope(c, PKB_KERNEL_MISC);
  checksum_kernel<<<grid, 256, 0, c->stream>>>(d, n, d_sum);
  PKB_CUDA(cudaGetLastError());
  return PKB_OK;
}

}  // namespace pkb
