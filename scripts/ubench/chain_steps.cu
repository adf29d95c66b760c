// Micro-benchmark: the per-step body of the lattice chain kernel built up feature by feature
// (one warp alone on an SM sub-partition) to see what each piece costs beyond the ~41-cycle
// dependency chain.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ float pow2i(int d) { return __int_as_float((max(d, -127) + 127) << 23); }
constexpr int kNegI = -(1 << 28);

template <int MODE>
__global__ void k(float *out, long long *cyc, float2 *gout, const float4 *gin, int iters, int P) {
  extern __shared__ float4 ring[];          // [64][P] arcs
  __shared__ float2 edge[64];
  for (int i = threadIdx.x; i < 64 * P; i += blockDim.x) ring[i] = gin[i];
  if (threadIdx.x < 64) edge[threadIdx.x] = make_float2(0.f, __int_as_float(kNegI));
  __syncthreads();
  const int lane = threadIdx.x & 31, r0 = threadIdx.x;
  const bool lane_in = lane == 0, publish = (lane == 31) && (MODE >= 4);
  float m = (r0 == 0) ? 1.f : 0.f;
  int o = (r0 == 0) ? 0 : kNegI;
  float4 a4 = ring[r0];
  float2 ev = edge[0];
  float2 *pa = gout + r0;
  int nb_o_sh = __shfl_up_sync(0xffffffffu, o, 1);
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
    const int el = i & 63, eln = (i + 1) & 63;
    float4 b4 = a4;
    float2 ev2 = ev;
    if (MODE >= 1) b4 = ring[eln * P + r0];                 // arc prefetch
    if (MODE >= 3) ev2 = edge[el];                           // edge read
    float nb_m = __shfl_up_sync(0xffffffffu, m, 1);
    nb_m = lane_in ? ev.x : nb_m;
    const int nb_o = lane_in ? __float_as_int(ev.y) : nb_o_sh;
    const int EA = nb_o + __float_as_int(a4.y), EB = o + __float_as_int(a4.w);
    const int on = max(max(EA, EB), kNegI);
    nb_o_sh = __shfl_up_sync(0xffffffffu, on, 1);
    const float gx = a4.x * pow2i(EA - on);
    const float raw = fmaf(nb_m, gx, m * (a4.z * pow2i(EB - on)));
    if (MODE >= 2) { pa[0] = make_float2(raw, __int_as_float(on)); pa += P; if ((i & 255) == 255) pa -= 256 * P; }
    m = raw; o = on;
    if (MODE >= 4 && publish) edge[el] = make_float2(m, __int_as_float(o));
    if (MODE >= 5 && (i & 15) == 15) {                       // periodic normalisation
      const int bits = __float_as_int(m);
      const bool alive = m > 0.f;
      o = alive ? o + ((bits >> 23) - 127) : o;
      m = alive ? __int_as_float((bits & 0x007fffff) | 0x3f800000) : m;
      nb_o_sh = __shfl_up_sync(0xffffffffu, o, 1);
    }
    a4 = b4; ev = ev2;
  }
  long long t1 = clock64();
  out[threadIdx.x] = m + o;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads) {
  const int P = 128;
  float *out; long long *cyc; float2 *gout; float4 *gin;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&gout, sizeof(float2) * P * 300); cudaMalloc(&gin, sizeof(float4) * 64 * P);
  float4 *h = new float4[64 * P];
  for (int i = 0; i < 64 * P; ++i) {
    int e1 = -3, e2 = -2;
    float f1, f2;
    memcpy(&f1, &e1, 4); memcpy(&f2, &e2, 4);
    h[i] = make_float4(1.3f, f1, 1.1f, f2);
  }
  cudaMemcpy(gin, h, sizeof(float4) * 64 * P, cudaMemcpyHostToDevice);
  const int iters = 20000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * P * 16);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<1, threads, 64 * P * 16>>>(out, cyc, gout, gin, iters, P);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %3d thr %7.1f cycles/step\n", name, threads, (double)c / iters);
}

int main() {
  for (int threads : {32, 128}) {
    run<0>("chains only (arcs in registers)", threads);
    run<1>("+ arc prefetch from smem (LDS.128)", threads);
    run<2>("+ global store of (m, o) (STG.64)", threads);
    run<3>("+ edge read (LDS.64)", threads);
    run<4>("+ predicated publish (STS.64)", threads);
    run<5>("+ normalisation every 16 steps", threads);
  }
  return 0;
}
