// pipes.cu -- issue-rate micro-benchmarks that size the fused Gauss-Newton kernel's instruction budget on B200:
// FFMA (3-register), FFMA2 (packed f32x2), FMUL2, FFMA + ALU mix, MUFU.RCP, TEX.  Prints warp-instructions per
// clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_IT 4096
#define UNR 16

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, long long* cyc, float a, float b, cudaTextureObject_t tex)
{
  float r[UNR];
  u64 q[UNR];
  int ii[UNR];
  for (int j = 0; j < UNR; ++j) { r[j] = threadIdx.x * 0.001f + j; q[j] = pk(r[j], r[j] + 1.f); ii[j] = threadIdx.x + j; }
  u64 pa = pk(a, b), pb = pk(b, a);
  float x[UNR], y[UNR]; u64 q2[UNR];
  for (int j = 0; j < UNR; ++j) { x[j] = a + j * 0.01f + threadIdx.x * 1e-6f; y[j] = b - j * 0.02f; q2[j] = pk(x[j], y[j]); }
  int flag = (a > 0.f) ? (threadIdx.x | 1) : 0;
  long long t0 = clock64();
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (MODE == 0) r[j] = fmaf(r[j], a, b);                                     // FFMA reg,reg,reg
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[j]) : "l"(pa), "l"(pb));
      if (MODE == 2) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(q[j]) : "l"(pa));
      if (MODE == 3) { r[j] = fmaf(r[j], a, b); ii[j] = (ii[j] ^ it) + j; }        // FFMA + ALU (LOP3/IADD3)
      if (MODE == 4) r[j] = __frcp_rn(r[j]) ;                                      // placeholder, replaced below
      if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[j]));      // MUFU.RCP
      if (MODE == 6) r[j] = tex2D<float>(tex, r[j], r[j] + 1.5f);                  // TEX
      if (MODE == 7) { r[j] = fmaf(r[j], a, b); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[j]) : "l"(pa), "l"(pb)); }
      if (MODE == 8) r[j] = r[j] + a;                                              // FADD
      if (MODE == 9) { r[j] = fmaf(r[j], a, b); r[j] = fmaxf(r[j], a); }           // FFMA + FMNMX(alu)
      if (MODE == 10) r[j] = fmaf(x[j], y[(j + 5) % UNR], r[j]);                   // FFMA, three distinct registers
      if (MODE == 11) asm volatile("{.reg .pred p; setp.ne.s32 p, %3, 0; @p fma.rn.f32 %0, %1, %2, %0;}" : "+f"(r[j]) : "f"(x[j]), "f"(y[(j + 5) % UNR]), "r"(flag));
      if (MODE == 12) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(q[j]) : "l"(q2[j]), "l"(q2[(j + 5) % UNR]));
      if (MODE == 13) { r[j] = fmaf(x[j], y[(j + 5) % UNR], r[j]); r[j] = fmaxf(r[j], x[(j + 3) % UNR]); }  // FFMA + FMNMX distinct regs
      if (MODE == 14) { r[j] = fmaf(x[j], y[(j + 5) % UNR], r[j]); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j])); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < UNR; ++j) s += r[j] + lo(q[j]) + ii[j] + x[j] + y[j] + lo(q2[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_j, cudaTextureObject_t tex, int warps_per_sm)
{
  int threads = 256, blocks_per_sm = warps_per_sm / 8;
  int nb = 148 * blocks_per_sm;
  float* out; long long* cyc;
  cudaMalloc(&out, nb * threads * 4); cudaMalloc(&cyc, nb * 8);
  kern<MODE><<<nb, threads>>>(out, cyc, 1.0001f, 0.5f, tex);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<nb, threads>>>(out, cyc, 1.0001f, 0.5f, tex);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 8]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < nb; ++i) c += h[i]; c /= nb;
  double winstr = (double)N_IT * UNR * instr_per_j * warps_per_sm;  // warp instructions per SM
  double ev = winstr / (ms * 1e-3 * 1.965e9);  // per SM per clock from the event time at 1965 MHz
  printf("%-28s warps/SM %2d : %.3f warp-instr/clk/SMSP (events), %.3f (clock64), %.1f us, %s\n", name, warps_per_sm, ev / 4,
         winstr / c / 4, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main()
{
  cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
  cudaMallocArray(&arr, &cd, 640, 480);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.filterMode = cudaFilterModeLinear; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  for (int w : {16}) {
    run<0>("FFMA r,r,r", 1, tex, w);
    run<1>("FFMA2", 1, tex, w);
    run<2>("FMUL2", 1, tex, w);
    run<3>("FFMA + LOP3 + IADD", 3, tex, w);
    run<5>("MUFU.RCP", 1, tex, w);
    run<7>("FFMA + FFMA2", 2, tex, w);
    run<8>("FADD", 1, tex, w);
    run<9>("FFMA + FMNMX", 2, tex, w);
    run<10>("FFMA 3 distinct regs", 1, tex, w);
    run<11>("@p FFMA 3 distinct regs", 1, tex, w);
    run<12>("FFMA2 3 distinct pairs", 1, tex, w);
    run<13>("FFMA + FMNMX distinct", 2, tex, w);
    run<14>("FFMA + MUFU.RCP", 2, tex, w);
  }
  run<6>("TEX 2D linear (same texel)", 1, tex, 16);
  return 0;
}
