// scale_est.cu -- drop-in residual sampling, scale / nu estimation and chi-square on explicit error vectors.
#include "scale_core.cuh"

namespace rgbid {

namespace {

// K11: errorGridStrideKernel (src/cuda/sigmaFuncs.cu:116-134)
__global__ void compute_error_kernel(ImgB im1, ImgB im0, float* __restrict__ error, int kept_rows, int kept_cols,
                                     int stride)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= kept_cols || y >= kept_rows) return;
  error[(size_t)y * kept_cols + x] = im1.row(0, stride * y)[stride * x] - im0.row(0, stride * y)[stride * x];
}

// K12 / K13 / K14 on explicit error vectors: one 8-CTA cluster, samples read from global memory.
__global__ void __cluster_dims__(kScaleCluster, 1, 1) __launch_bounds__(kScaleThreads, kScaleMinBlocks)
    scale_from_errors_kernel(const float* __restrict__ err0, const float* __restrict__ err1, int n, int op, int mest,
                             float bias0, float sigma0, float bias1, float sigma1, ScaleState* __restrict__ out)
{
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ ScaleShared sh;
  const int rank = (int)cluster.block_rank();
  const int chunk = (n + kScaleCluster - 1) / kScaleCluster;
  const int begin = min(rank * chunk, n);
  const int n_local = min(chunk, n - begin);
  ScaleSlot s0, s1;
  slot_init(s0, op, mest, bias0, sigma0, err0 != nullptr);
  slot_init(s1, op, mest, bias1, sigma1, err1 != nullptr);
  scale_rounds(cluster, sh, s0, s1, err0 ? err0 + begin : nullptr, err1 ? err1 + begin : nullptr, n_local);
  if (rank == 0 && threadIdx.x == 0) {
    ScaleState st;
    st.bias_int = s0.out_bias; st.sigma_int = s0.out_sigma; st.nu_int = s0.nu; st.irls_iters_int = s0.irls_iters;
    st.bias_depthinv = s1.out_bias; st.sigma_depthinv = s1.out_sigma; st.nu_depthinv = s1.nu;
    st.irls_iters_depthinv = s1.irls_iters;
    *out = st;
  }
}

// K15: normalizeAndAppendErrorsKernel + computeChiSquaredKernelPartial + finalReductionChiSquaredKernel
// (sigmaFuncs.cu:137-150, 541-647) in one pass; out2 = [sum rho, N] accumulated with double atomics
// (two addresses, one atomic pair per CTA).
__device__ __forceinline__ float chi_rho(float e, int mest)
{
  float rho = (e * e) / 2.f;
  if (mest == RGBID_HUBER) { if (fabsf(e) > 1.345f) rho = 1.345f * (fabsf(e) - 1.345f / 2.f); }
  else if (mest == RGBID_TUKEY) {
    if (fabsf(e) < 4.685f) {
      float a1 = (e / 4.685f) * (e / 4.685f);
      float a2 = (1.f - a1) * (1.f - a1) * (1.f - a1);
      rho = ((4.685f * 4.685f) / 6.f) * (1.f - a2);
    } else rho = (4.685f * 4.685f) / 6.f;
  } else if (mest == RGBID_STUDENT) rho = ((5.f + 1.f) / 2.f) * logf(1.f + (e * e) / 5.f);
  return rho;
}

__global__ void __launch_bounds__(256) chi_square_kernel(const float* __restrict__ err_int,
                                                         const float* __restrict__ err_depth, int n, float sigma_int,
                                                         float sigma_depth, int mest, double* __restrict__ out2)
{
  float s_rho = 0.f, s_n = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += gridDim.x * blockDim.x) {
    float e = (i < n) ? err_int[i] / sigma_int : err_depth[i - n] / sigma_depth;
    if (!(isinf(e) || isnan(e))) { s_rho += chi_rho(e, mest); s_n += 1.f; }
  }
  double d_rho = warp_sum((double)s_rho), d_n = warp_sum((double)s_n);
  __shared__ double sm[2][8];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[0][wid] = d_rho; sm[1][wid] = d_n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += sm[0][w]; b += sm[1][w]; }
    atomicAdd(&out2[0], a);
    atomicAdd(&out2[1], b);
  }
}

}  // namespace

void launch_compute_error(const LaunchCtx& L, ImgB im1, ImgB im0, float* error, int kept_rows, int kept_cols,
                          int stride)
{
  dim3 block(32, 8), grid((kept_cols + 31) / 32, (kept_rows + 7) / 8);
  compute_error_kernel<<<grid, block, 0, L.stream>>>(im1, im0, error, kept_rows, kept_cols, stride);
  ++*L.launches;
}

void launch_scale_from_errors(const LaunchCtx& L, const float* err0, const float* err1, int n, int op, int mest,
                              float bias0, float sigma0, float bias1, float sigma1, ScaleState* out)
{
  static PerDevice table;
  table.once([] {
    if (kScaleCluster > 8 &&
        cudaFuncSetAttribute(scale_from_errors_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return false;
    return upload_nu_table();
  });
  scale_from_errors_kernel<<<kScaleCluster, kScaleThreads, 0, L.stream>>>(err0, err1, n, op, mest, bias0, sigma0,
                                                                          bias1, sigma1, out);
  ++*L.launches;
}

void launch_chi_square(const LaunchCtx& L, const float* err_int, const float* err_depth, int n, float sigma_int,
                       float sigma_depth, int mest, double* out2)
{
  cudaMemsetAsync(out2, 0, 2 * sizeof(double), L.stream);
  int grid = min((2 * n + 255) / 256, L.num_sms * 4);
  if (grid < 1) grid = 1;
  chi_square_kernel<<<grid, 256, 0, L.stream>>>(err_int, err_depth, n, sigma_int, sigma_depth, mest, out2);
  ++*L.launches;
}

}  // namespace rgbid
