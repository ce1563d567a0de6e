// FP64 add and FP32<->FP64 conversion rates / dependent-chain latency on sm_100a: the CMVN kernel
// reproduces the reference's running sum, which is one such chain per (utterance, dim) and frame
// (src/cmvn.cc:38-101).
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void probe(double *out, long long *cyc, int iters) {
  double d[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { d[i] = 1.0 + i + threadIdx.x * 1e-3; f[i] = 1.0f + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {        // independent DADDs
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = d[i] + 1.0000001;
    } else if (OP == 1) { // independent float -> double -> float round trips
#pragma unroll
      for (int i = 0; i < 8; ++i) { double t = (double)f[i]; asm volatile("" : "+d"(t)); f[i] = (float)t; }
    } else if (OP == 2) { // the CMVN chain: stat = (float)((double)stat + (double)x - (double)y)
      float stat = f[0];
#pragma unroll
      for (int i = 0; i < 8; ++i) stat = (float)((double)stat + (double)f[i & 3] - (double)f[4 + (i & 3)]);
      f[0] = stat;
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  const char *names[] = {"DADD (independent)", "F32->F64->F32 (independent)", "CMVN chain step (dependent)"};
  for (int w = 1; w <= 4; w *= 2) {
    for (int op = 0; op < 3; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) probe<0><<<148, 128 * w>>>(out, cyc, iters);
        if (op == 1) probe<1><<<148, 128 * w>>>(out, cyc, iters);
        if (op == 2) probe<2><<<148, 128 * w>>>(out, cyc, iters);
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-30s warps/smsp=%d  cycles per warp-op(step) per smsp = %.1f   (per warp: %.1f)\n", names[op], w,
             (double)h[0] / iters / 8 / w, (double)h[0] / iters / 8);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
