// Micro-benchmark: flat step loop of the lattice chain kernel with the chunk-boundary work as a
// rarely taken (warp-uniform) slow path every CH steps, linear ring walk with wrap.  Forward direction,
// one row per lane.  Measures cycles per step for 1 and 4 warps per block, 1 and 64 blocks.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float pow2i(int d) { return __int_as_float((max(d, -127) + 127) << 23); }
constexpr int kNegI = -(1 << 28);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int MODE>
__global__ void k(float *out, long long *cyc, float2 *gout, const float4 *gin, int iters, int P, int CH) {
  extern __shared__ float4 ring[];          // [64][P] arcs
  __shared__ float2 edge[64];
  __shared__ uint64_t bar[2];
  for (int i = threadIdx.x; i < 64 * P; i += blockDim.x) ring[i] = gin[i];
  if (threadIdx.x < 64) edge[threadIdx.x] = make_float2(0.f, __int_as_float(kNegI));
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive(&bar[0]); mbar_arrive(&bar[1]); }   // phase 0 complete: wait(0) always succeeds
  __syncthreads();
  const int lane = threadIdx.x & 31, r0 = threadIdx.x;
  const bool lane_in = lane == 0, publish = (lane == 31);
  float m = (r0 == 0) ? 1.f : 0.f;
  int o = (r0 == 0) ? 0 : kNegI;
  const float4 *xrow = ring + r0;            // this lane's column, current row
  const float4 *xend = ring + r0 + 64 * P;
  const float2 *pe = edge; float2 *po = edge;
  const float2 *eend = edge + 64;
  float4 a4 = xrow[0];
  float2 ev = pe[0];
  float2 *pa = gout + (size_t)blockIdx.x * P * 300 + r0;
  int nb_o_sh = __shfl_up_sync(0xffffffffu, o, 1);
  int el = 0;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
    const bool last = (el == CH - 1);                     // warp-uniform
    if (MODE >= 1 && last) {                              // slow path A: the next stage must have landed
      mbar_wait(&bar[0], 0);
      mbar_wait(&bar[1], 0);
    }
    const float4 *xn = xrow + P; xn = (xn == xend) ? ring + r0 : xn;
    const float4 b4 = *xn;
    const float2 ev2 = *pe;
    float nb_m = __shfl_up_sync(0xffffffffu, m, 1);
    nb_m = lane_in ? ev.x : nb_m;
    const int nb_o = lane_in ? __float_as_int(ev.y) : nb_o_sh;
    const int EA = nb_o + __float_as_int(a4.y), EB = o + __float_as_int(a4.w);
    const int on = max(max(EA, EB), kNegI);
    nb_o_sh = __shfl_up_sync(0xffffffffu, on, 1);
    const float gx = a4.x * pow2i(EA - on);
    const float raw = fmaf(nb_m, gx, m * (a4.z * pow2i(EB - on)));
    pa[0] = make_float2(raw, __int_as_float(on)); pa += P; if ((i & 255) == 255) pa -= 256 * P;
    m = raw; o = on;
    if (publish) *po = make_float2(m, __int_as_float(o));
    if (MODE >= 1 && last) {                              // slow path B: chunk finished
      const int bits = __float_as_int(m);
      const bool alive = m > 0.f;
      o = alive ? o + ((bits >> 23) - 127) : o;
      m = alive ? __int_as_float((bits & 0x007fffff) | 0x3f800000) : m;
      nb_o_sh = __shfl_up_sync(0xffffffffu, o, 1);
      if (publish) mbar_arrive(&bar[1]);
      if (MODE >= 2 && lane == 0) mbar_arrive(&bar[0]);   // stands in for the stage recycling of the tail warp
      // keep the barriers in "phase complete" state for the next wait: one more arrive completes phase 1 -> parity 0 again
      if (publish) mbar_arrive(&bar[1]);
      if (MODE >= 2 && lane == 0) mbar_arrive(&bar[0]);
    }
    a4 = b4; ev = ev2;
    xrow = xn;
    ++pe; pe = (pe == eend) ? edge : pe;
    ++po; po = (po == eend) ? edge : po;
    el = last ? 0 : el + 1;
  }
  long long t1 = clock64();
  out[threadIdx.x] = m + o;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads, int blocks, int CH) {
  const int P = 128;
  float *out; long long *cyc; float2 *gout; float4 *gin;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&gout, sizeof(float2) * P * 300 * 64); cudaMalloc(&gin, sizeof(float4) * 64 * P);
  float4 *h = new float4[64 * P];
  for (int i = 0; i < 64 * P; ++i) {
    int e1 = -3, e2 = -2; float f1, f2;
    memcpy(&f1, &e1, 4); memcpy(&f2, &e2, 4);
    h[i] = make_float4(1.3f, f1, 1.1f, f2);
  }
  cudaMemcpy(gin, h, sizeof(float4) * 64 * P, cudaMemcpyHostToDevice);
  const int iters = 20000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * P * 16);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<blocks, threads, 64 * P * 16>>>(out, cyc, gout, gin, iters, P, CH);
  cudaError_t e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s %3d thr x %2d blocks CH=%2d %7.1f cycles/step %s\n", name, threads, blocks, CH, (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  run<0>("flat loop, no chunk work", 128, 64, 16);
  run<1>("flat loop + boundary slow path", 32, 1, 16);
  run<1>("flat loop + boundary slow path", 128, 64, 16);
  run<1>("flat loop + boundary slow path", 128, 64, 8);
  return 0;
}
