// texpat.cu -- what a texture gather of the fused Gauss-Newton kernel costs as a function of how a warp's 32 lanes are
// laid over the image.  The fused kernel gives every lane 4 CONSECUTIVE keyframe pixels (one LDS.128 per map), so a
// warp-wide TEX instruction samples at x = x0 + 4 * lane + k: 32 lanes spread over 128 pixels and two rows (bilinear),
// i.e. ~32 sectors per instruction.  The alternative is x = x0 + lane + 32 * k (32 lanes over 32 pixels, ~10 sectors).
// Prints warp-TEX per clock per SM for both layouts, linear and point filtering, and for a plain LDG gather.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o texpat texpat.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int COLS = 640, ROWS = 480, NIMG = 32;  // 32 streams' worth of current-frame maps (39 MB: L2 resident, not L1)
constexpr int CHUNKS_PER_IMG = COLS * ROWS / 128;

// MODE 0: stride-4 lanes, 1: contiguous lanes; FILTER via the texture object; MODE 2/3: the same two layouts with LDG
template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, cudaTextureObject_t* tex, const float* raw, int iters, float dx, float dy)
{
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int c = (warp + it * nwarps) % (CHUNKS_PER_IMG * NIMG);
    const int img = c / CHUNKS_PER_IMG, ci = c % CHUNKS_PER_IMG;
    const int y = ci / 5, xc = (ci % 5) * 128;
    const cudaTextureObject_t t = tex[img];
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = (MODE & 1) ? xc + lane + 32 * k : xc + 4 * lane + k;
      const float xf = (float)x + dx, yf = (float)y + dy;
      if (MODE < 2) v[k] = tex2D<float>(t, xf + 0.5f, yf + 0.5f);
      else {
        const int xi = min(max(__float2int_rd(xf), 0), COLS - 1), yi = min(max(__float2int_rd(yf), 0), ROWS - 1);
        v[k] = __ldg(raw + ((size_t)img * ROWS + yi) * COLS + xi);
      }
    }
    acc += (v[0] + v[1]) + (v[2] + v[3]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, cudaTextureObject_t* tex, const float* raw, float* out, float dx, float dy)
{
  const int nb = 148 * 2, iters = 512;
  kern<MODE><<<nb, 256>>>(out, tex, raw, iters, dx, dy);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<nb, 256>>>(out, tex, raw, iters, dx, dy);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double wtex_per_sm = (double)iters * 4 * 16;  // 16 warps per SM
  printf("%-44s %8.1f us  %.4f warp-gathers/clk/SM  (%.1f clk per warp-gather per SM)  %s\n", name, ms * 1e3,
         wtex_per_sm / (ms * 1e-3 * 1.965e9), (ms * 1e-3 * 1.965e9) / wtex_per_sm, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  float* raw; cudaMalloc(&raw, (size_t)NIMG * ROWS * COLS * 4);
  cudaMemset(raw, 0, (size_t)NIMG * ROWS * COLS * 4);
  cudaTextureObject_t hl[NIMG], hp[NIMG], *dl, *dp;
  for (int i = 0; i < NIMG; ++i)
    for (int f = 0; f < 2; ++f) {
      cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D;
      rd.res.pitch2D.devPtr = raw + (size_t)i * ROWS * COLS; rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
      rd.res.pitch2D.width = COLS; rd.res.pitch2D.height = ROWS; rd.res.pitch2D.pitchInBytes = COLS * 4;
      cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = f ? cudaAddressModeBorder : cudaAddressModeClamp;
      td.filterMode = f ? cudaFilterModePoint : cudaFilterModeLinear; td.readMode = cudaReadModeElementType;
      cudaCreateTextureObject(f ? &hp[i] : &hl[i], &rd, &td, nullptr);
    }
  cudaMalloc(&dl, sizeof(hl)); cudaMalloc(&dp, sizeof(hp));
  cudaMemcpy(dl, hl, sizeof(hl), cudaMemcpyHostToDevice); cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
  for (int rep = 0; rep < 2; ++rep) {
    const float dx = rep ? 2.3f : 0.3f, dy = rep ? 1.4f : 0.4f;
    printf("offset (%.1f, %.1f)\n", dx, dy);
    run<0>("TEX linear f32, lanes x0 + 4*lane + k", dl, raw, out, dx, dy);
    run<1>("TEX linear f32, lanes x0 + lane + 32*k", dl, raw, out, dx, dy);
    run<0>("TEX point  f32, lanes x0 + 4*lane + k", dp, raw, out, dx, dy);
    run<1>("TEX point  f32, lanes x0 + lane + 32*k", dp, raw, out, dx, dy);
    run<2>("LDG gather,     lanes x0 + 4*lane + k", dl, raw, out, dx, dy);
    run<3>("LDG gather,     lanes x0 + lane + 32*k", dl, raw, out, dx, dy);
  }
  return 0;
}
