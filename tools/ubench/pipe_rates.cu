// Issue-rate probe for the epilogue's instruction mix (FFMA, FMNMX, F2FP pack) on sm_100a.
// Prints cycles per warp-instruction per SM sub-partition with W warps resident per sub-partition.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void probe(float *out, long long *cyc, int iters) {
  float a[16], b = threadIdx.x * 1e-3f;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i + b;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0001f, b);
    } else if (OP == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
    } else if (OP == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) a[2 * i] = __uint_as_float(u[i]);   // keep a dependence so nothing is hoisted
    } else if (OP == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) a[2 * i] = __uint_as_float(u[i]);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  const char *names[] = {"FFMA", "FMNMX", "F2FP.BF16", "F2FP.F16"};
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    const int threads = 128 * warps_per_smsp;
    for (int op = 0; op < 4; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) probe<0><<<148, threads>>>(out, cyc, iters);
        if (op == 1) probe<1><<<148, threads>>>(out, cyc, iters);
        if (op == 2) probe<2><<<148, threads>>>(out, cyc, iters);
        if (op == 3) probe<3><<<148, threads>>>(out, cyc, iters);
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const int n_inst = (op < 2 ? 16 : 8);
      printf("%-10s warps/smsp=%d  cycles per warp-instr per smsp = %.2f\n", names[op], warps_per_smsp,
             (double)h[0] / iters / n_inst / warps_per_smsp);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
