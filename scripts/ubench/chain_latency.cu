// Micro-benchmark: cycles per iteration of dependent chains built from the instructions of the
// lattice chain kernel (one warp alone on an SM sub-partition).  nvcc -arch=sm_100a -o chain_latency chain_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float pow2i(int d) { return __int_as_float((max(d, -127) + 127) << 23); }

template <int MODE>
__global__ void k(float *out, long long *cyc, int iters, float a, float b, int ia) {
  float m = 1.0f + threadIdx.x * 1e-3f;
  int o = threadIdx.x;
  __shared__ float sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {            // FFMA only
      m = fmaf(m, a, b);
    } else if (MODE == 1) {     // SHFL + FFMA
      float nb = __shfl_up_sync(0xffffffffu, m, 1);
      m = fmaf(nb, a, m * b);
    } else if (MODE == 2) {     // SHFL + SEL + FFMA
      float nb = __shfl_up_sync(0xffffffffu, m, 1);
      nb = (threadIdx.x == 0) ? a : nb;
      m = fmaf(nb, a, m * b);
    } else if (MODE == 3) {     // integer frame chain: SHFL + IADD + MAX3
      int nbo = __shfl_up_sync(0xffffffffu, o, 1);
      o = max(max(nbo + ia, o + ia), -(1 << 28));
    } else if (MODE == 4) {     // both chains, coupled like the kernel
      float nb = __shfl_up_sync(0xffffffffu, m, 1);
      int nbo = __shfl_up_sync(0xffffffffu, o, 1);
      int EA = nbo + ia, EB = o + ia;
      int on = max(max(EA, EB), -(1 << 28));
      float gx = a * pow2i(EA - on), gy = b * pow2i(EB - on);
      m = fmaf(nb, gx, m * gy);
      o = on;
    } else if (MODE == 5) {     // LDS dependent chain (pointer chase through smem)
      o = __float_as_int(sm[o & 63]) & 63;
    } else if (MODE == 6) {     // ALU dependent chain: IADD
      o = o + ia; o = o ^ ia;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = m + o;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter_ops) {
  float *out; long long *cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<MODE><<<1, 32>>>(out, cyc, iters, 0.999f, 0.001f, 1);
  k<MODE><<<1, 32>>>(out, cyc, iters, 0.999f, 0.001f, 1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %7.1f cycles/iter\n", name, (double)c / iters);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("SHFL + FMUL/FFMA", 1);
  run<2>("SHFL + SEL + FFMA", 1);
  run<3>("SHFL + IADD + IMAX3 (frame chain)", 1);
  run<4>("both chains coupled (as the kernel)", 1);
  run<5>("LDS pointer chase", 1);
  run<6>("IADD + XOR", 1);
  return 0;
}
