// tex_probe.cu -- characterises the texture unit's bilinear filter (cudaFilterModeLinear, float texels,
// unnormalised coordinates, clamp) so that the software sampler in common.cuh can reproduce what the
// reference's warpIntensityWithTrafo3DInvDepth (src/cuda/warping_registration.cu:943, :493) gets from hardware.
// Test tooling only.  Build: nvcc -arch=sm_100a -o tex_probe tex_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__global__ void sample1d(cudaTextureObject_t tex, float base, int n, float step, float y, float* out)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = tex2D<float>(tex, base + k * step + 0.5f, y);
}

__global__ void sample2d(cudaTextureObject_t tex, const float* xs, const float* ys, int n, float* out)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = tex2D<float>(tex, xs[k], ys[k]);
}

static cudaTextureObject_t make_tex(float* d, size_t pitch, int w, int h)
{
  cudaResourceDesc r = {};
  r.resType = cudaResourceTypePitch2D;
  r.res.pitch2D.devPtr = d; r.res.pitch2D.pitchInBytes = pitch; r.res.pitch2D.width = w; r.res.pitch2D.height = h;
  r.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc t = {};
  t.readMode = cudaReadModeElementType; t.addressMode[0] = t.addressMode[1] = cudaAddressModeClamp;
  t.filterMode = cudaFilterModeLinear; t.normalizedCoords = 0;
  cudaTextureObject_t tex = 0;
  cudaCreateTextureObject(&tex, &r, &t, nullptr);
  return tex;
}

int main()
{
  const int W = 640, H = 8;
  float* d; size_t pitch;
  cudaMallocPitch(&d, &pitch, W * sizeof(float), H);
  std::vector<float> ramp(W * H);
  for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) ramp[j * W + i] = (float)i;
  cudaMemcpy2D(d, pitch, ramp.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice);
  cudaTextureObject_t tex = make_tex(d, pitch, W, H);
  const int N = 1 << 16;  // 65536 steps of 1/65536 px
  float* dout; cudaMalloc(&dout, N * 4);
  std::vector<float> out(N);
  for (float base : {3.f, 300.f, 630.f}) {
    sample1d<<<N / 256, 256>>>(tex, base, N, 1.f / N, 1.5f, dout);
    cudaMemcpy(out.data(), dout, N * 4, cudaMemcpyDeviceToHost);
    // alpha_hw(k) = out - base; find step positions
    int nsteps = 0; double first_step = -1, worst_round = 0, worst_trunc = 0; int not_mult = 0;
    float prev = out[0];
    for (int k = 0; k < N; ++k) {
      double frac = (double)k / N;
      double a = (double)out[k] - base;
      if (fabs(a * 256 - llround(a * 256)) > 1e-3) ++not_mult;
      worst_round = fmax(worst_round, fabs(a - floor(frac * 256 + 0.5) / 256));
      worst_trunc = fmax(worst_trunc, fabs(a - floor(frac * 256) / 256));
      if (out[k] != prev) { if (first_step < 0) first_step = frac * 256; ++nsteps; prev = out[k]; }
    }
    printf("base %.0f: distinct steps %d, first step at frac*256 = %.4f, not-multiple-of-1/256: %d, max|a-round| = %.5f, max|a-trunc| = %.5f\n",
           base, nsteps, first_step, not_mult, worst_round, worst_trunc);
    // print the fractional positions (in 1/256 units) of the first 4 steps
    prev = out[0]; int shown = 0;
    for (int k = 0; k < N && shown < 4; ++k) if (out[k] != prev) { printf("   step to %.6f at frac*256 = %.4f\n", out[k] - base, (double)k / N * 256); prev = out[k]; ++shown; }
  }
  // 2-D: random texels, random coordinates
  const int W2 = 64, H2 = 64, M = 1 << 16;
  std::vector<float> img(W2 * H2), xs(M), ys(M), res(M);
  srand(7);
  for (auto& v : img) v = 255.f * rand() / RAND_MAX;
  for (int k = 0; k < M; ++k) { xs[k] = 1.f + 61.f * rand() / RAND_MAX; ys[k] = 1.f + 61.f * rand() / RAND_MAX; }
  float *d2, *dx, *dy, *dr; size_t p2;
  cudaMallocPitch(&d2, &p2, W2 * 4, H2);
  cudaMemcpy2D(d2, p2, img.data(), W2 * 4, W2 * 4, H2, cudaMemcpyHostToDevice);
  cudaMalloc(&dx, M * 4); cudaMalloc(&dy, M * 4); cudaMalloc(&dr, M * 4);
  cudaMemcpy(dx, xs.data(), M * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, ys.data(), M * 4, cudaMemcpyHostToDevice);
  cudaTextureObject_t tex2 = make_tex(d2, p2, W2, H2);
  sample2d<<<M / 256, 256>>>(tex2, dx, dy, M, dr);
  cudaMemcpy(res.data(), dr, M * 4, cudaMemcpyDeviceToHost);
  const char* names[] = {"round(frac*256)", "trunc(frac*256)", "round((x-0.5)*256) fixed", "trunc((x-0.5)*256) fixed", "exact"};
  for (int mode = 0; mode < 5; ++mode) {
    double worst = 0, mean = 0;
    for (int k = 0; k < M; ++k) {
      double a, b; int i0, j0;
      if (mode == 0 || mode == 1 || mode == 4) {
        float xB = xs[k] - 0.5f, yB = ys[k] - 0.5f;
        i0 = (int)floorf(xB); j0 = (int)floorf(yB);
        double fa = (double)xB - i0, fb = (double)yB - j0;
        if (mode == 0) { a = floor(fa * 256 + 0.5) / 256; b = floor(fb * 256 + 0.5) / 256; }
        else if (mode == 1) { a = floor(fa * 256) / 256; b = floor(fb * 256) / 256; }
        else { a = fa; b = fb; }
      } else {
        double fx = ((double)xs[k] - 0.5) * 256, fy = ((double)ys[k] - 0.5) * 256;
        long long qx = mode == 2 ? llround(fx) : (long long)floor(fx), qy = mode == 2 ? llround(fy) : (long long)floor(fy);
        i0 = (int)(qx >> 8); j0 = (int)(qy >> 8); a = (qx & 255) / 256.0; b = (qy & 255) / 256.0;
      }
      int i1 = i0 + 1, j1 = j0 + 1;
      double t00 = img[j0 * W2 + i0], t10 = img[j0 * W2 + i1], t01 = img[j1 * W2 + i0], t11 = img[j1 * W2 + i1];
      double v = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
      double e = fabs(v - res[k]);
      worst = fmax(worst, e); mean += e;
    }
    printf("2-D formula with %-28s: mean |err| = %.3e, max |err| = %.3e\n", names[mode], mean / M, worst);
  }
  return 0;
}
