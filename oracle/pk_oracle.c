/* TEST INFRASTRUCTURE -- CPU restatement ("port") of pocketkaldi's acoustic
 * front half. Not product code: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py may load this library; the CUDA product in
 * pocketkaldi_b200/ never does.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function here
 * against (1) the reference's own known-answer vectors and golden files
 * (test/srfft_test.cc:11-271, test/data/fbankmat_en-us-hello.wav.txt,
 * test/data/fbankcmvnmat_en-us-hello.wav.txt, test/nnet_test.cc:25-72,
 * test/gemm_test.cc:39-60), committed under tests/golden/, and (2) the
 * unmodified reference compiled into oracle/_ref/libpkref.so.
 *
 * This is a restatement, not a copy: it follows the reference's arithmetic
 * (which values are float, which are double, which sums are sequential), but
 * the FFT is an ordinary iterative radix-2 instead of the reference's
 * recursive split-radix, so FFT outputs agree to float rounding, not bit for
 * bit. Everything after raw fbank (CMVN, splice, layers, AM epilogue,
 * decodable scale) reproduces the reference's rounding sequence exactly.
 *
 * Build with -ffp-contract=off (oracle/Makefile): the reference is compiled
 * for baseline x86-64, so its float multiply-adds are NOT fused except inside
 * the AVX2 SGEMM micro-kernel, which pko_linear restates with explicit fmaf.
 *
 * Each function names the reference lines it restates (paths relative to
 * /root/reference).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PKO_SR 16000
#define PKO_SHIFT 160       /* src/fbank.cc:15  (int)(16000*0.001*10.0) */
#define PKO_FRAME 400       /* src/fbank.cc:16  (int)(16000*0.001*25.0) */
#define PKO_NFFT 512        /* src/fbank.cc:24-33 round up to power of two */
#define PKO_NMEL 40         /* src/fbank.h:10 */
#define PKO_CMVN_WIN 600    /* src/cmvn.h:10 */
#define PKO_CMVN_GLOBAL 200 /* src/cmvn.h:11 */

/* ------------------------------------------------------------------ framing */

/* src/fbank.cc:35-42 (snip-edges frame count). */
int pko_num_frames(int num_samples) {
  if (num_samples < PKO_FRAME) return 0;
  return 1 + (num_samples - PKO_FRAME) / PKO_SHIFT;
}

/* src/fbank.cc:249-256: float a = 6.28318530718/399; w[i] = 0.54 - 0.46*cos(a*(float)i),
 * the product a*i is a float multiply, cos and the affine step are double. */
void pko_hamming(float *w) {
  float a = (float)(6.28318530718 / (PKO_FRAME - 1));
  for (int i = 0; i < PKO_FRAME; ++i) {
    float fi = (float)i;
    w[i] = (float)(0.54 - 0.46 * cos((double)(a * fi)));
  }
}

/* src/fbank.h:29-31. */
static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

/* src/fbank.cc:103-163. weights is [40][256] dense (zero outside each
 * triangle); offset/width give the non-zero run the reference stores. */
void pko_mel_table(float *weights, int *offset, int *width) {
  const int nbins = PKO_NFFT / 2;
  float sample_freq = PKO_SR;
  float bin_width = sample_freq / PKO_NFFT;
  float mel_lo = mel_scale(20);
  float mel_hi = mel_scale(PKO_SR / 2);
  float delta = (mel_hi - mel_lo) / (PKO_NMEL + 1);
  memset(weights, 0, sizeof(float) * PKO_NMEL * nbins);
  for (int m = 0; m < PKO_NMEL; ++m) {
    float left = mel_lo + m * delta;
    float center = mel_lo + (m + 1) * delta;
    float right = mel_lo + (m + 2) * delta;
    int first = -1, last = -1;
    for (int i = 0; i < nbins; ++i) {
      float mel = mel_scale(bin_width * i);
      if (mel > left && mel < right) {
        float w = (mel <= center) ? (mel - left) / (center - left)
                                  : (right - mel) / (right - center);
        weights[m * nbins + i] = w;
        if (first < 0) first = i;
        last = i;
      }
    }
    offset[m] = first;
    width[m] = last + 1 - first;
  }
}

/* src/fbank.cc:74-100 + :44-69: copy 400 samples, zero-pad to 512, subtract
 * the float mean, pre-emphasis with the double literal 0.97, Hamming. */
void pko_window(const float *wave, int frame, const float *hamming, float *out) {
  const float *src = wave + (size_t)frame * PKO_SHIFT;
  float sum = 0;
  for (int i = 0; i < PKO_FRAME; ++i) { out[i] = src[i]; sum += src[i]; }
  for (int i = PKO_FRAME; i < PKO_NFFT; ++i) out[i] = 0.0f;
  float mean = sum / PKO_FRAME;
  for (int i = 0; i < PKO_FRAME; ++i) out[i] -= mean;
  for (int i = PKO_FRAME - 1; i > 0; --i)
    out[i] = (float)((double)out[i] - 0.97 * (double)out[i - 1]);
  out[0] = (float)((double)out[0] - 0.97 * (double)out[0]);
  for (int i = 0; i < PKO_FRAME; ++i) out[i] *= hamming[i];
}

/* ---------------------------------------------------------------------- FFT */

/* In-place complex FFT of n points (interleaved re,im), forward sign.
 * Restates the *result* of complexfft_compute (src/srfft.cc:319-341 over
 * :95-317); the butterfly order differs (radix-2 DIT), see file header. */
static void cfft(float *x, int n) {
  int logn = 0;
  while ((1 << logn) < n) ++logn;
  for (int i = 0; i < n; ++i) {
    int j = 0;
    for (int b = 0; b < logn; ++b) j |= ((i >> b) & 1) << (logn - 1 - b);
    if (j > i) {
      float tr = x[2 * i], ti = x[2 * i + 1];
      x[2 * i] = x[2 * j]; x[2 * i + 1] = x[2 * j + 1];
      x[2 * j] = tr; x[2 * j + 1] = ti;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    for (int k = 0; k < half; ++k) {
      double ang = -2.0 * M_PI * k / len;
      float wr = (float)cos(ang), wi = (float)sin(ang);
      for (int s = 0; s < n; s += len) {
        float *a = x + 2 * (s + k), *b = x + 2 * (s + k + half);
        float tr = b[0] * wr - b[1] * wi;
        float ti = b[0] * wi + b[1] * wr;
        b[0] = a[0] - tr; b[1] = a[1] - ti;
        a[0] += tr; a[1] += ti;
      }
    }
  }
}

/* src/srfft.cc:371-461, forward branch: n/2-point complex FFT of the packed
 * input, then the real-FFT post-pass with the reference's float twiddle
 * recurrence kN *= rootN (:387-394, complex_mul :343-347). Output packing
 * [Re0, Re(n/2), Re1, Im1, ...]. n is the real length (power of two). */
void pko_srfft(float *data, int n) {
  int n2 = n / 2;
  cfft(data, n2);
  float root_re = (float)cos((double)(float)(6.283185307179586476925286766559005 / n * -1));
  float root_im = (float)sin((double)(float)(6.283185307179586476925286766559005 / n * -1));
  float k_re = 1.0f, k_im = 0.0f;
  for (int k = 1; 2 * k <= n2; ++k) {
    float t = k_re * root_re - k_im * root_im;
    k_im = k_re * root_im + k_im * root_re;
    k_re = t;
    float ck_re = (float)(0.5 * (data[2 * k] + data[n - 2 * k]));
    float ck_im = (float)(0.5 * (data[2 * k + 1] - data[n - 2 * k + 1]));
    float dk_re = (float)(0.5 * (data[2 * k + 1] + data[n - 2 * k + 1]));
    float dk_im = (float)(-0.5 * (data[2 * k] - data[n - 2 * k]));
    data[2 * k] = ck_re + (dk_re * k_re - dk_im * k_im);
    data[2 * k + 1] = ck_im + (dk_re * k_im + dk_im * k_re);
    int kd = n2 - k;
    if (kd != k) {
      data[2 * kd] = ck_re + (dk_re * -k_re - (-dk_im) * k_im);
      data[2 * kd + 1] = -ck_im + (dk_re * k_im + (-dk_im) * -k_re);
    }
  }
  float zeroth = data[0] + data[1], n2th = data[0] - data[1];
  data[0] = zeroth;
  data[1] = n2th;
}

/* ------------------------------------------------------------------- fbank */

/* src/fbank.cc:193-211 power spectrum, :165-184 mel (sequential float dot over
 * the non-zero run, src/vector.cc:251-262), :244-245 floor FLT_EPSILON and
 * log in double. */
static void frame_to_logmel(float *win, const float *melw, const int *off,
                            const int *wid, float *out40) {
  const int half = PKO_NFFT / 2;
  pko_srfft(win, PKO_NFFT);
  float first = win[0] * win[0], last = win[1] * win[1];
  for (int i = 1; i < half; ++i) {
    float re = win[2 * i], im = win[2 * i + 1];
    win[i] = re * re + im * im;
  }
  win[0] = first;
  win[half] = last;
  for (int m = 0; m < PKO_NMEL; ++m) {
    const float *w = melw + m * half + off[m];
    const float *p = win + off[m];
    float e = 0.0f;
    for (int i = 0; i < wid[m]; ++i) e += w[i] * p[i];
    if (e < FLT_EPSILON) e = FLT_EPSILON;
    out40[m] = (float)log((double)e);
  }
}

/* src/fbank.cc:267-292. wave: n float samples (int16 range, unscaled).
 * out: [T][40]. Returns T. */
int pko_fbank(const float *wave, int n, float *out) {
  int T = pko_num_frames(n);
  if (T == 0) return 0;
  float hamming[PKO_FRAME];
  float *melw = (float *)malloc(sizeof(float) * PKO_NMEL * (PKO_NFFT / 2));
  int off[PKO_NMEL], wid[PKO_NMEL];
  pko_hamming(hamming);
  pko_mel_table(melw, off, wid);
  float win[PKO_NFFT];
  for (int t = 0; t < T; ++t) {
    pko_window(wave, t, hamming, win);
    frame_to_logmel(win, melw, off, wid, out + (size_t)t * PKO_NMEL);
  }
  free(melw);
  return T;
}

/* -------------------------------------------------------------------- CMVN */

/* src/cmvn.cc:103-115 over :35-71 (ComputeStats), :73-92 (SmoothStats),
 * :94-101 (Apply). The running stats are carried in float and widened to
 * double only inside a step; smoothing and the final subtraction are float
 * multiply-then-add (src/vector.cc:358-380 AddVec). raw/out: [T][40];
 * global: 40 sums + count. */
void pko_cmvn(const float *raw, int T, const float *global, float *out) {
  float cached[PKO_NMEL + 1];
  for (int t = 0; t < T; ++t) {
    double acc[PKO_NMEL + 1];
    float stats[PKO_NMEL + 1];
    const float *x = raw + (size_t)t * PKO_NMEL;
    for (int d = 0; d <= PKO_NMEL; ++d) acc[d] = (t > 0) ? (double)cached[d] : 0.0;
    for (int d = 0; d < PKO_NMEL; ++d) acc[d] += x[d];
    acc[PKO_NMEL] += 1.0;
    if (t - PKO_CMVN_WIN >= 0) {
      const float *p = raw + (size_t)(t - PKO_CMVN_WIN) * PKO_NMEL;
      for (int d = 0; d < PKO_NMEL; ++d) acc[d] += -1.0 * p[d];
      acc[PKO_NMEL] -= 1.0;
    }
    for (int d = 0; d <= PKO_NMEL; ++d) { stats[d] = (float)acc[d]; cached[d] = stats[d]; }
    double count = stats[PKO_NMEL];
    if (count < PKO_CMVN_WIN) {
      double from_global = PKO_CMVN_WIN - count;
      double gcount = global[PKO_NMEL];
      if (from_global > PKO_CMVN_GLOBAL) from_global = PKO_CMVN_GLOBAL;
      float alpha = (float)(from_global / gcount);
      for (int d = 0; d <= PKO_NMEL; ++d) {
        stats[d] += alpha * global[d];
      }
    }
    double cnt = stats[PKO_NMEL];
    float scale = (float)(1 / cnt);
    float *y = out + (size_t)t * PKO_NMEL;
    for (int d = 0; d < PKO_NMEL; ++d) {
      y[d] = x[d] + -scale * stats[d];
    }
  }
}

/* ------------------------------------------------------------------ splice */

/* src/am.cc:65-88: frames t-left..t+right, indices clamped to [0, T-1].
 * feats: [T][dim]; out: [T][(left+right+1)*dim]. */
void pko_splice(const float *feats, int T, int dim, int left, int right, float *out) {
  int w = left + right + 1;
  for (int t = 0; t < T; ++t) {
    for (int c = 0; c < w; ++c) {
      int s = t + c - left;
      if (s < 0) s = 0;
      if (s >= T) s = T - 1;
      memcpy(out + ((size_t)t * w + c) * dim, feats + (size_t)s * dim, sizeof(float) * dim);
    }
  }
}

/* ------------------------------------------------------------------ layers */

/* src/nnet.cc:22-36 over src/matrix.cc:418-436 and src/gemm.cc:69-125:
 * y = x W^T + b. W is [out][in] as stored on disk. Each output accumulates
 * its products sequentially in k with fused multiply-add inside a K-chunk of
 * 512 (KC, src/gemm.h:50; the micro-kernel is vfmadd231ps,
 * src/gemm_haswell.cc:122-283) and chunks are summed through C
 * (src/gemm.cc:98-100); the bias is added afterwards (src/nnet.cc:32-35). */
void pko_linear(const float *x, int T, int in, const float *W, const float *b,
                int out, float *y) {
  float *wt = (float *)malloc(sizeof(float) * (size_t)in * out); /* [in][out] */
  for (int o = 0; o < out; ++o)
    for (int k = 0; k < in; ++k) wt[(size_t)k * out + o] = W[(size_t)o * in + k];
  float *acc = (float *)malloc(sizeof(float) * out);
  for (int t = 0; t < T; ++t) {
    const float *xr = x + (size_t)t * in;
    float *yr = y + (size_t)t * out;
    for (int o = 0; o < out; ++o) yr[o] = 0.0f;
    for (int k0 = 0; k0 < in; k0 += 512) {
      int k1 = k0 + 512 < in ? k0 + 512 : in;
      for (int o = 0; o < out; ++o) acc[o] = 0.0f;
      for (int k = k0; k < k1; ++k) {
        float a = xr[k];
        const float *wr = wt + (size_t)k * out;
        for (int o = 0; o < out; ++o) acc[o] = fmaf(a, wr[o], acc[o]);
      }
      if (k0 == 0) for (int o = 0; o < out; ++o) yr[o] = acc[o];
      else for (int o = 0; o < out; ++o) yr[o] += acc[o];
    }
    for (int o = 0; o < out; ++o) yr[o] += b[o];
  }
  free(acc);
  free(wt);
}

/* src/nnet.cc:49-60. */
void pko_relu(float *x, size_t n) {
  for (size_t i = 0; i < n; ++i) if (x[i] < 0.0f) x[i] = 0.0f;
}

/* src/nnet.cc:62-75: float sequential sum of squares (src/vector.cc:251-262),
 * sqrt(D / sum) in double, float scale; no floor. */
void pko_normalize(float *x, int T, int D) {
  for (int t = 0; t < T; ++t) {
    float *r = x + (size_t)t * D;
    float s = 0.0f;
    for (int i = 0; i < D; ++i) {
      s += r[i] * r[i];
    }
    double sq = s;
    float scale = (float)sqrt((float)D / sq);
    for (int i = 0; i < D; ++i) r[i] *= scale;
  }
}

/* src/nnet.cc:38-47 over src/vector.cc:264-277: expf, sequential float sum,
 * divide; no max subtraction. */
void pko_softmax(float *x, int T, int D) {
  for (int t = 0; t < T; ++t) {
    float *r = x + (size_t)t * D;
    float s = 0;
    for (int i = 0; i < D; ++i) { float e = expf(r[i]); r[i] = e; s += e; }
    for (int i = 0; i < D; ++i) r[i] /= s;
  }
}

/* src/am.cc:106-112: floor at (float)1e-20, log in double, minus log-prior
 * (AddVec alpha=-1: float multiply then add). log_prior already holds logs
 * (src/am.cc:42-43). */
void pko_am_epilogue(float *p, int T, int P, const float *log_prior) {
  const float floor_val = (float)1.0e-20;
  for (int t = 0; t < T; ++t) {
    float *r = p + (size_t)t * P;
    for (int i = 0; i < P; ++i) {
      float v = r[i] < floor_val ? floor_val : r[i];
      v = (float)log((double)v);
      r[i] = v + -1.0f * log_prior[i];
    }
  }
}

/* src/am.cc:42-43: prior probabilities -> log (double log, float store). */
void pko_log_prior(const float *prior, int P, float *log_prior) {
  for (int i = 0; i < P; ++i) log_prior[i] = (float)log((double)prior[i]);
}

/* src/decodable.cc:15 over src/matrix.cc:99-103. */
void pko_scale(float *x, size_t n, float s) {
  for (size_t i = 0; i < n; ++i) x[i] *= s;
}

/* src/decodable.cc:24-31: log_prob[frame][tid2pdf[tid]]. */
float pko_decodable_loglikelihood(const float *log_prob, int P, const int32_t *tid2pdf,
                                  int frame, int tid) {
  return log_prob[(size_t)frame * P + tid2pdf[tid]];
}

/* C[m x n] = A[m x k] B[k x n], naive float (src/matrix.cc:393-410
 * SimpleMatMat, the reference's own GEMM oracle). */
void pko_simple_matmat(const float *A, const float *B, float *C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      float s = 0.0f;
      for (int q = 0; q < k; ++q) s += A[(size_t)i * k + q] * B[(size_t)q * n + j];
      C[(size_t)i * n + j] = s;
    }
}
