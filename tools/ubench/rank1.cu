// rank1.cu -- which form of the 27-term rank-1 update  acc += s * (r, e)(r, e)^T  issues fastest on a B200 SM.
// The fused Gauss-Newton kernel spends 2 x 33 FMA-pipe instructions per pixel on it; ncu shows the kernel bound by
// dispatch stalls (register-bank pressure), so the accumulate is measured in isolation, 16 warps / SM like the kernel,
// rows produced by a few cheap instructions per pixel so that the accumulate dominates.
//   S0  scalar, predicated FFMA (round-1 kernel)          S1  scalar, plain FFMA, zeroed weight
//   S2  packed pairs of rows with swapped copies           S3  packed row pairs x broadcast scalar column
//   S4  packed over two pixels (54 accumulators)           S5  as S1 with the 27 sums in column-major order
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rank1 rank1.cu ; prints cycles per pixel-constraint / SMSP.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_IT 2048

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ void pfma(float& d, float a, float b, int flag)
{
  asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\n@p fma.rn.f32 %0, %1, %2, %0;\n}" : "+f"(d) : "f"(a), "f"(b), "r"(flag));
}

template <int S>
__global__ void __launch_bounds__(256, 2) kern(float* out, long long* cyc, float a0, float b0)
{
  float acc[27];
  u64 a2[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) { acc[k] = 0.f; a2[k] = 0ull; }
  float t = threadIdx.x * 1e-3f + a0;
  const long long c0 = clock64();
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int p = 0; p < 4; p += (S == 4 ? 2 : 1)) {
      // rows of one (S4: two) pixel(s): 7 dependent-free cheap instructions per pixel
      float r[7], q[7];
      t += b0;
#pragma unroll
      for (int i = 0; i < 7; ++i) { r[i] = fmaf(t, (float)(i + 1) * 0.125f, (float)p); q[i] = fmaf(t, (float)(i + 2) * 0.25f, (float)p); }
      const float s = t * 0.5f, s1 = t * 0.25f;
      if (S == 0) {
        const int flag = (s > -1e30f);
        int sh = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float si = s * r[i];
#pragma unroll
          for (int j = i; j < 7; ++j) { pfma(acc[sh], si, r[j], flag); ++sh; }
        }
      } else if (S == 1) {
        int sh = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float si = s * r[i];
#pragma unroll
          for (int j = i; j < 7; ++j) { acc[sh] = fmaf(si, r[j], acc[sh]); ++sh; }
        }
      } else if (S == 5) {
        // column-major: the column value r[j] is the operand shared by consecutive FMAs
        float sr[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) sr[i] = s * r[i];
        int sh = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
#pragma unroll
          for (int i = 0; i <= j && i < 6; ++i) { acc[sh] = fmaf(sr[i], r[j], acc[sh]); ++sh; }
        }
      } else if (S == 2) {
        const u64 A = pk(r[0], r[1]), B = pk(r[2], r[3]), C = pk(r[4], r[5]);
        const u64 As = pk(r[1], r[0]), Bs = pk(r[3], r[2]), Cs = pk(r[5], r[4]);
        const u64 SS = pk(s, s), E = pk(r[6], r[6]);
        const u64 sA = mul2(SS, A), sB = mul2(SS, B), sC = mul2(SS, C);
        fma2(a2[0], sA, A); fma2(a2[1], sA, As); fma2(a2[2], sA, B); fma2(a2[3], sA, Bs); fma2(a2[4], sA, C);
        fma2(a2[5], sA, Cs); fma2(a2[6], sB, B); fma2(a2[7], sB, Bs); fma2(a2[8], sB, C); fma2(a2[9], sB, Cs);
        fma2(a2[10], sC, C); fma2(a2[11], sC, Cs); fma2(a2[12], sA, E); fma2(a2[13], sB, E); fma2(a2[14], sC, E);
      } else if (S == 3) {
        // row pairs (r0,r1) (r2,r3) (r4,r5) scaled once, times a broadcast column value
        const u64 SS = pk(s, s);
        const u64 sP0 = mul2(SS, pk(r[0], r[1])), sP1 = mul2(SS, pk(r[2], r[3])), sP2 = mul2(SS, pk(r[4], r[5]));
        int sh = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) { fma2(a2[sh], sP0, pk(r[j], r[j])); ++sh; }
#pragma unroll
        for (int j = 2; j < 7; ++j) { fma2(a2[sh], sP1, pk(r[j], r[j])); ++sh; }
#pragma unroll
        for (int j = 4; j < 7; ++j) { fma2(a2[sh], sP2, pk(r[j], r[j])); ++sh; }
      } else if (S == 4) {
        // two pixels per packed instruction: lane 0 = pixel p (rows r, weight s), lane 1 = pixel p + 1 (rows q, s1)
        u64 R[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) R[i] = pk(r[i], q[i]);
        const u64 SS = pk(s, s1);
        int sh = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const u64 si = mul2(SS, R[i]);
#pragma unroll
          for (int j = i; j < 7; ++j) { fma2(a2[sh], si, R[j]); ++sh; }
        }
      }
    }
  }
  const long long c1 = clock64();
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 27; ++k) { float lo, hi; upk(a2[k], lo, hi); sum += acc[k] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int S>
void run(const char* name)
{
  const int nb = 148 * 2, threads = 256;
  float* out; long long* cyc;
  cudaMalloc(&out, nb * threads * 4); cudaMalloc(&cyc, nb * 8);
  kern<S><<<nb, threads>>>(out, cyc, 1.0001f, 1e-6f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<S><<<nb, threads>>>(out, cyc, 1.0001f, 1e-6f);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 2]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < nb; ++i) c += h[i]; c /= nb;
  // per SMSP: 4 warps, each N_IT * 4 pixel-constraints
  const double per = c / (4.0 * N_IT * 4);
  printf("%-44s %7.1f us  %6.2f clk per pixel-constraint per SMSP (clock64)  %6.2f (events @1965 MHz)  %s\n", name, ms * 1e3, per,
         ms * 1e-3 * 1.965e9 / (4.0 * N_IT * 4), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main()
{
  run<0>("S0 scalar predicated FFMA");
  run<1>("S1 scalar plain FFMA");
  run<5>("S5 scalar plain FFMA, column-major");
  run<2>("S2 packed, swapped pairs");
  run<3>("S3 packed row pairs x broadcast column");
  run<4>("S4 packed over two pixels");
  return 0;
}
