// Batched log-mel filterbank for sm_100a.
//
// Replaces the per-frame loop of pocketkaldi::Fbank::Compute (src/fbank.cc:267-292):
// ExtractWindow/ProcessWindow (:44-100), pk_srfft_compute (src/srfft.cc:371-461),
// ComputePowerSpectrum (:193-211), Melbanks::Compute (:165-184), floor + log (:244-245).
//
// Mapping. A block of 4 warps works on a tile of 8 consecutive frames of one
// utterance, two frames per warp; the warps never synchronise with each other. Half a
// warp (16 lanes) owns one frame and stages its 400 samples into the frame's own
// shared-memory region (which later serves as the FFT transpose tile):
//   * the 512-point real FFT is done, as in the reference, as a 256-point complex
//     FFT of the packed signal plus a real post-pass; the 256 points are factored
//     16 x 16: lane j runs a register-resident radix-16 DFT over points j + 16q,
//     the half-warp transposes through a padded (bank-conflict-free) shared tile,
//     and lane k1 runs the second radix-16 DFT, ending with bins k1 + 16*k2;
//   * the post-pass needs bin 256-k, which lives in lane 16-k1 at a static
//     register index: one width-16 shuffle per value, exact twiddles from a table
//     (the reference uses a float recurrence, srfft.cc:387-394; the table is
//     closer to the true DFT and agrees with it to float rounding);
//   * the power spectrum goes to shared memory and the 492 non-zero mel weights
//     are spread evenly over the 32 lanes (16 each), partial sums are combined in
//     a fixed order, so results are deterministic.
// All tables are computed on the host with the reference's float formulas.

#include <float.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace pkb {

namespace {

// per-frame shared region: 16 x 17 complex transpose tile (>= 400 staged samples, >= 256 power
// values); 8 extra float2 make consecutive regions start 16 banks apart, so the 32-bit accesses
// of the two half-warps of a warp never collide
constexpr int kXStride = 17;                                          // padded transpose row
constexpr int kXRegion = 16 * kXStride + 8;                           // float2 per frame
constexpr int kPartSlots = 96;

// Optional dither (not in the reference, default off): counter-based N(0,1) pair for the sample
// pair `pair` of frame t of utterance u -- splitmix64 finaliser + Box-Muller.
__device__ __forceinline__ float2 dither_pair(uint64_t seed, int u, int t, int pair) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(u) * 0x100000001B3ull +
                                                static_cast<uint64_t>(t) * 1024ull + static_cast<uint64_t>(pair) + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u1 = (static_cast<float>(static_cast<uint32_t>(z >> 40)) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
  const float u2 = static_cast<float>(static_cast<uint32_t>(z) >> 8) * (1.0f / 16777216.0f);            // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * (-i)
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

// Forward 16-point DFT in registers. Output register r holds bin (r >> 2) + 4 * (r & 3).
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
  // a[4*k1 + n2] *= W16^(n2*k1)
  a[5] = cmul(a[5], make_float2(c1, -s1));    // k1=1,n2=1: W^1
  a[6] = cmul(a[6], make_float2(h, -h));      // k1=1,n2=2: W^2
  a[7] = cmul(a[7], make_float2(s1, -c1));    // k1=1,n2=3: W^3
  a[9] = cmul(a[9], make_float2(h, -h));      // k1=2,n2=1: W^2
  a[10] = mul_neg_i(a[10]);                   // k1=2,n2=2: W^4
  a[11] = cmul(a[11], make_float2(-h, -h));   // k1=2,n2=3: W^6
  a[13] = cmul(a[13], make_float2(s1, -c1));  // k1=3,n2=1: W^3
  a[14] = cmul(a[14], make_float2(-h, -h));   // k1=3,n2=2: W^6
  a[15] = cmul(a[15], make_float2(-c1, s1));  // k1=3,n2=3: W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
}

__host__ __device__ constexpr int bin_of_reg(int r) { return (r >> 2) + 4 * (r & 3); }
__host__ __device__ constexpr int reg_of_bin(int k) { return 4 * (k & 3) + (k >> 2); }

template <typename SampleT>
// 7 blocks (28 warps) per SM: 72 registers and 25 KB of shared memory per block. The FFT twiddle
// tables are read through L1 (__ldg) instead of being staged (measured 1.072 M -> 1.034 M
// cycles per 511 k frames when that bought the seventh block).
__global__ void __launch_bounds__(128, 7)
fbank_kernel(const SampleT *__restrict__ pcm, const int64_t *__restrict__ sample_off,
             const int32_t *__restrict__ num_samples, const int32_t *__restrict__ num_frames,
             const int64_t *__restrict__ frame_off, const int32_t *__restrict__ tile_prefix,
             const int32_t *__restrict__ tile_utt, int n_tiles, FbankTables tab, float *__restrict__ out,
             float dither, uint64_t dither_seed) {
  __shared__ __align__(16) float s_ham[kFrame];
  __shared__ float4 s_melw[128];
  __shared__ uint32_t s_melb[128];
  __shared__ uint32_t s_melc[32];
  __shared__ uint32_t s_mels[kMel];
  __shared__ __align__(16) float2 s_x[4][2][kXRegion];  // per warp, per frame: samples, transpose tile, power
  __shared__ float s_part[4][2][kPartSlots];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int half = lane >> 4, j = lane & 15;

  for (int i = tid; i < kFrame; i += 128) s_ham[i] = tab.hamming[i];
  s_melw[tid] = tab.mel_w[tid];
  s_melb[tid] = tab.mel_bins[tid];
  if (tid < 32) s_melc[tid] = tab.mel_ctl[tid];
  if (tid < kMel) s_mels[tid] = tab.mel_sum[tid];
  __syncthreads();  // the only block-wide barrier: the tables are staged

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int u = __ldg(tile_utt + tile);
    const int t0 = (tile - tile_prefix[u]) * kFramesPerTile;
    const int T = num_frames[u];
    const int f = warp * 2 + half;  // frame inside the tile
    const int t = t0 + f;
    const bool valid = t < T;
    float2 *xb = s_x[warp][half];
    float *x = reinterpret_cast<float *>(xb);

    // ---- stage this frame's 400 samples (src/fbank.cc:74-100: they always lie inside the
    //      utterance, T = 1 + (n - 400) / 160). 16-byte vector loads when the frame start is aligned.
    if (valid) {
      const SampleT *fs = pcm + sample_off[u] + static_cast<int64_t>(t) * kShift;
      constexpr int kVec = 16 / sizeof(SampleT);
      if ((reinterpret_cast<uintptr_t>(fs) & 15) == 0) {
        const uint4 *v = reinterpret_cast<const uint4 *>(fs);
        for (int i = j; i < kFrame / kVec; i += 16) {
          const uint4 w = __ldg(v + i);
          if (sizeof(SampleT) == 2) {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            float d[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              d[2 * q] = static_cast<float>(static_cast<int16_t>(ww[q] & 0xffffu));
              d[2 * q + 1] = static_cast<float>(static_cast<int16_t>(ww[q] >> 16));
            }
            *reinterpret_cast<float4 *>(x + i * 8) = make_float4(d[0], d[1], d[2], d[3]);
            *reinterpret_cast<float4 *>(x + i * 8 + 4) = make_float4(d[4], d[5], d[6], d[7]);
          } else {
            *reinterpret_cast<uint4 *>(x + i * 4) = w;
          }
        }
      } else {
        for (int i = j; i < kFrame; i += 16) x[i] = static_cast<float>(fs[i]);
      }
    }
    __syncwarp();

    // ---- window: DC removal, pre-emphasis, Hamming (src/fbank.cc:44-69) ----
    float2 a[16];
    float xm1[13];
    float sum = 0.0f;
#pragma unroll
    for (int q = 0; q < 13; ++q) {
      const int m = j + 16 * q;  // complex point = samples 2m, 2m+1
      float2 v = make_float2(0.0f, 0.0f);
      float p = 0.0f;
      if (q < 12 || j < 8) {
        v = *reinterpret_cast<const float2 *>(x + 2 * m);
        p = x[m == 0 ? 0 : 2 * m - 1];
        if (dither != 0.0f) {
          const float2 n = dither_pair(dither_seed, u, t, m);
          v.x = fmaf(dither, n.x, v.x);
          v.y = fmaf(dither, n.y, v.y);
          p = m == 0 ? v.x : fmaf(dither, dither_pair(dither_seed, u, t, m - 1).y, p);
        }
      }
      a[q] = v;
      xm1[q] = p;
      sum += v.x + v.y;
    }
    __syncwarp();  // every lane has read its samples: the region becomes the transpose tile
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
    const float mean = sum / static_cast<float>(kFrame);
#pragma unroll
    for (int q = 0; q < 13; ++q) {
      const int m = j + 16 * q;
      if (q < 12 || j < 8) {
        const float a0 = a[q].x - mean, a1 = a[q].y - mean, am = xm1[q] - mean;
        const float y0 = fmaf(-0.97f, am, a0);  // sample 0: x0 - 0.97*x0 (am == a0)
        const float y1 = fmaf(-0.97f, a0, a1);
        const float2 w = *reinterpret_cast<const float2 *>(s_ham + 2 * m);
        a[q] = make_float2(y0 * w.x, y1 * w.y);
      }
    }
    a[13] = a[14] = a[15] = make_float2(0.0f, 0.0f);  // zero padding 400 -> 512

    // ---- 256-point complex FFT = 16 x 16 ----
    dft16(a);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int k1 = bin_of_reg(r);
      float2 v = a[r];
      if (k1 != 0) v = cmul(v, __ldg(tab.tw_pass + k1 * 16 + j));
      xb[k1 * kXStride + j] = v;
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) a[jj] = xb[j * kXStride + jj];  // lane j now plays k1 = j
    __syncwarp();
    dft16(a);  // a[r] = Z[k1 + 16*k2], k2 = bin_of_reg(r), k1 = j

    // ---- real-FFT post-pass + power spectrum (srfft.cc:396-440, fbank.cc:193-211) ----
    // Bins k and 256-k come from the same pair (Z_k, Z_{256-k}):
    //   A_k = (E - iT)/2, A_{256-k} = conj((E + iT)/2),  E = Z_k + conj(Z_{256-k}),
    //   T = W512^k (Z_k - conj(Z_{256-k})),
    // so the two power values share E and T. Lane j handles its own
    // bins j + 16*k2 for k2 < 8 together with their partners (which live in lane 16-j at the
    // static register 15-r), i.e. 8 pairs instead of 16 single bins.
    float *pw = reinterpret_cast<float *>(xb);  // 256 floats, reuses the transpose tile
    const int src_lane = (16 - j) & 15;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int k2 = bin_of_reg(r);
      if (k2 >= 8) continue;  // compile-time: the other half is produced as partners
      float2 p;
      p.x = __shfl_sync(0xffffffffu, a[15 - r].x, src_lane, 16);
      p.y = __shfl_sync(0xffffffffu, a[15 - r].y, src_lane, 16);
      // lane 0: the partner of bin 16*k2 is this lane's own bin 16*(16-k2) (k2 = 0: itself)
      const float2 own = a[reg_of_bin((16 - k2) & 15)];
      if (j == 0) p = own;
      const float2 z = a[r];
      const float2 e = make_float2(z.x + p.x, z.y - p.y);
      const float2 o = make_float2(z.x - p.x, z.y + p.y);
      const float2 tw = __ldg(tab.tw_real + k2 * 16 + j);
      const float2 tt = cmul(tw, o);
      // formed as complex sums first: the difference form |E|^2 + |T|^2 -/+ 2 Re(..) would cancel
      // when one bin of the pair is much weaker than the other
      const float ax = e.x + tt.y, ay = e.y - tt.x;  // E - iT
      const float bx = e.x - tt.y, by = e.y + tt.x;  // E + iT
      const int k = j + 16 * k2;
      pw[k] = 0.25f * fmaf(ax, ax, ay * ay);
      if (k != 0) pw[256 - k] = 0.25f * fmaf(bx, bx, by * by);
    }
    if (j == 0) {
      const float2 z = a[reg_of_bin(8)];  // bin 128 pairs with itself: |A_128|^2 = |Z_128|^2
      pw[128] = fmaf(z.x, z.x, z.y * z.y);
    }
    __syncwarp();

    // ---- mel filterbank, floor, log (fbank.cc:165-184, :244-245) ----
    // this lane's 16 (weight, bin) entries are the same for both frames of the warp: one load
    float wv[16];
    uint32_t bb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = s_melw[q * 32 + lane];
      wv[4 * q] = w.x; wv[4 * q + 1] = w.y; wv[4 * q + 2] = w.z; wv[4 * q + 3] = w.w;
      bb[q] = s_melb[q * 32 + lane];
    }
    const uint32_t ctl = s_melc[lane];
#pragma unroll
    for (int fr = 0; fr < 2; ++fr) {
      const float *pf = reinterpret_cast<const float *>(s_x[warp][fr]);
      float *part = s_part[warp][fr];
      uint32_t slot = ctl >> 16;
      float acc = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc = fmaf(wv[4 * q + i], pf[(bb[q] >> (8 * i)) & 0xffu], acc);
          if ((ctl >> (4 * q + i)) & 1u) {
            part[slot++] = acc;
            acc = 0.0f;
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int fr = 0; fr < 2; ++fr) {
      const int tt = t0 + warp * 2 + fr;
      if (tt < T) {
        const float *part = s_part[warp][fr];
        float *dst = out + (frame_off[u] + tt) * kMel;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int m = lane + 32 * rep;
          if (m < kMel) {
            const uint32_t s = s_mels[m];
            const uint32_t first = s & 0xffffu, cnt = s >> 16;  // cnt in 1..3
            float e = part[first];
            if (cnt > 1) e += part[first + 1];
            if (cnt > 2) e += part[first + 2];
            e = fmaxf(e, FLT_EPSILON);
            dst[m] = logf(e);
          }
        }
      }
    }
    __syncwarp();  // the next tile's staging overwrites the power values
  }
}

template <typename SampleT>
int launch_fbank(Ctx *c, const SampleT *d_pcm, const BatchMeta &m, float *d_raw) {
  if (m.n_tiles == 0) return PKB_OK;
  int occ = 0;
  PKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fbank_kernel<SampleT>, 128, 0));
  if (occ < 1) occ = 1;
  int grid = std::min(m.n_tiles, c->sm_count * occ);
  LaunchScope scope(c, PKB_KERNEL_FBANK);
  fbank_kernel<SampleT><<<grid, 128, 0, c->stream>>>(d_pcm, m.d_sample_off, m.d_num_samples,
                                                     m.d_num_frames, m.d_frame_off,
                                                     m.d_tile_prefix, m.d_tile_utt, m.n_tiles, c->fb,
                                                     d_raw, c->dither, c->dither_seed);
  PKB_CUDA(cudaGetLastError());
  return PKB_OK;
}

}  // namespace

int launch_fbank_i16(Ctx *c, const int16_t *d_pcm, const BatchMeta &m, float *d_raw) {
  return launch_fbank<int16_t>(c, d_pcm, m, d_raw);
}
int launch_fbank_f32(Ctx *c, const float *d_wave, const BatchMeta &m, float *d_raw) {
  return launch_fbank<float>(c, d_wave, m, d_raw);
}

// ---------------------------------------------------------------- host tables
namespace {

float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }  // src/fbank.h:29-31

}  // namespace

int build_fbank_tables(Ctx *c) {
  // Hamming window, src/fbank.cc:249-256 (M_2PI is 6.28318530718 there, fbank.cc:18-20).
  // PKB_WINDOW_POVEY (not in the reference, off by default): Kaldi's pow(0.5 - 0.5 cos, 0.85).
  std::vector<float> ham(kFrame);
  {
    float a = 6.28318530718 / (kFrame - 1);
    for (int i = 0; i < kFrame; ++i) {
      float i_fl = static_cast<float>(i);
      if (c->window_type == PKB_WINDOW_POVEY) ham[i] = pow(0.5 - 0.5 * cos(a * i_fl), 0.85);
      else ham[i] = 0.54 - 0.46 * cos(a * i_fl);
    }
  }
  // exact twiddles
  std::vector<float2> twp(256), twr(256);
  const double two_pi = 6.283185307179586476925286766559005;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int j = 0; j < 16; ++j) {
      double ang = -two_pi * (j * k1) / 256.0;
      twp[k1 * 16 + j] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
    }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int k1 = 0; k1 < 16; ++k1) {
      double ang = -two_pi * (k1 + 16 * k2) / 512.0;
      twr[k2 * 16 + k1] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
    }
  // mel weights, src/fbank.cc:103-163, as a flat (filter, bin, weight) list
  struct Entry { int filter, bin; float w; };
  std::vector<Entry> ent;
  {
    const int nbins = kNfft / 2;
    float sample_freq = kSampleRate;
    float fft_bin_width = sample_freq / kNfft;
    float mel_low = mel_scale(20), mel_high = mel_scale(kSampleRate / 2);
    float delta = (mel_high - mel_low) / (kMel + 1);
    for (int m = 0; m < kMel; ++m) {
      float left = mel_low + m * delta;
      float center = mel_low + (m + 1) * delta;
      float right = mel_low + (m + 2) * delta;
      for (int i = 0; i < nbins; ++i) {
        float freq = fft_bin_width * i;
        float mel = mel_scale(freq);
        if (mel > left && mel < right) {
          float w = (mel <= center) ? (mel - left) / (center - left)
                                    : (right - mel) / (right - center);
          ent.push_back({m, i, w});
        }
      }
    }
  }
  if (ent.size() > 512) {
    set_error("mel table has %zu entries (> 512)", ent.size());
    return PKB_ERR_INVALID;
  }
  const size_t n_real = ent.size();
  while (ent.size() < 512) ent.push_back({kMel - 1, 0, 0.0f});
  std::vector<float4> melw(128);
  std::vector<uint32_t> melb(128), melc(32), mels(kMel, 0);
  std::vector<int> first(kMel, -1), count(kMel, 0);
  int slot = 0;
  for (int l = 0; l < 32; ++l) {
    uint32_t mask = 0;
    int slot0 = slot;
    for (int i = 0; i < 16; ++i) {
      const Entry &e = ent[16 * l + i];
      float *w = reinterpret_cast<float *>(&melw[(i >> 2) * 32 + l]);
      w[i & 3] = e.w;
      if ((i & 3) == 0) melb[(i >> 2) * 32 + l] = 0;
      melb[(i >> 2) * 32 + l] |= static_cast<uint32_t>(e.bin) << (8 * (i & 3));
      // padding entries (weight 0 on bin 0, past the last real entry) never flush: they add an
      // exact 0 to an accumulator nobody reads, instead of a fourth partial slot to filter 39
      const size_t idx = static_cast<size_t>(16 * l + i);
      const bool pad = idx >= n_real;
      bool flush = !pad && ((i == 15) || idx + 1 == n_real || ent[idx + 1].filter != e.filter);
      if (flush) {
        mask |= 1u << i;
        if (first[e.filter] < 0) first[e.filter] = slot;
        count[e.filter]++;
        slot++;
      }
    }
    melc[l] = mask | (static_cast<uint32_t>(slot0) << 16);
  }
  if (slot > kPartSlots) {
    set_error("mel partial slots %d > %d", slot, kPartSlots);
    return PKB_ERR_INVALID;
  }
  for (int m = 0; m < kMel; ++m) {
    // the kernel's combine step sums at most three partial slots per filter
    if (count[m] < 1 || count[m] > 3) {
      set_error("mel filter %d spans %d partial slots (kernel combines 1..3)", m, count[m]);
      return PKB_ERR_INVALID;
    }
    // slots of one filter are consecutive because entries are sorted by filter
    mels[m] = static_cast<uint32_t>(first[m]) | (static_cast<uint32_t>(count[m]) << 16);
  }

  size_t off_ham = 0;
  size_t off_twp = off_ham + sizeof(float) * 400;
  size_t off_twr = off_twp + sizeof(float2) * 256;
  size_t off_melw = off_twr + sizeof(float2) * 256;
  size_t off_melb = off_melw + sizeof(float4) * 128;
  size_t off_melc = off_melb + sizeof(uint32_t) * 128;
  size_t off_mels = off_melc + sizeof(uint32_t) * 32;
  size_t total = off_mels + sizeof(uint32_t) * kMel;
  std::vector<char> host(total);
  memcpy(&host[off_ham], ham.data(), sizeof(float) * 400);
  memcpy(&host[off_twp], twp.data(), sizeof(float2) * 256);
  memcpy(&host[off_twr], twr.data(), sizeof(float2) * 256);
  memcpy(&host[off_melw], melw.data(), sizeof(float4) * 128);
  memcpy(&host[off_melb], melb.data(), sizeof(uint32_t) * 128);
  memcpy(&host[off_melc], melc.data(), sizeof(uint32_t) * 32);
  memcpy(&host[off_mels], mels.data(), sizeof(uint32_t) * kMel);
  PKB_TRY(c->tables.ensure(total));
  PKB_TRY(upload(c, c->tables.p, host.data(), total));
  char *base = c->tables.as<char>();
  c->fb.hamming = reinterpret_cast<const float *>(base + off_ham);
  c->fb.tw_pass = reinterpret_cast<const float2 *>(base + off_twp);
  c->fb.tw_real = reinterpret_cast<const float2 *>(base + off_twr);
  c->fb.mel_w = reinterpret_cast<const float4 *>(base + off_melw);
  c->fb.mel_bins = reinterpret_cast<const uint32_t *>(base + off_melb);
  c->fb.mel_ctl = reinterpret_cast<const uint32_t *>(base + off_melc);
  c->fb.mel_sum = reinterpret_cast<const uint32_t *>(base + off_mels);
  return PKB_OK;
}

}  // namespace pkb
