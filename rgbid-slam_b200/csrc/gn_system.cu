// gn_system.cu -- the Gauss-Newton iteration kernels (the hot path).
//
// One reference iteration = warp inverse depth (K4) + warp intensity (K5) + two residual samplers (K11) +
// up to 34 tiny scale-estimation launches with host round trips (K12/K13) + partial sums (K1) + final
// reduction (K3) + device->host copy + host LLT / exp-map / pose update.  Here it is two launches and no
// host involvement:
//
//   gn_scale_kernel : 8-CTA cluster per frame pair.  Warps + samples the sub-sampled residuals straight into
//                     shared memory and runs the IRLS sigma / nu bisection rounds over DSMEM (scale_core.cuh).
//   gn_build_kernel : fused warp + bilinear sample + residual + 2x6 Jacobian rows + Student/M-estimator weights
//                     + the 27 upper-triangular J^T J | J^T r sums.  float4-vectorised coalesced reads of the 6
//                     keyframe maps, gathers of the 2 current-frame maps through the read-only path, per-thread
//                     FP32 accumulators, FP64 from the warp level up, deterministic per-CTA partials, and a
//                     last-block-done tail that sums the partials in fixed order, solves the 6x6 system
//                     (Cholesky), applies the SE(3) update and refreshes the pixel-space transforms of every
//                     pyramid level -- or inverts A for the covariance.
//
// Algorithmic HBM bytes: 32 B per keyframe pixel per iteration (6 keyframe maps + 2 current-frame maps,
// fp32); this kernel is bandwidth bound (about 5 flop/B), tensor cores do not apply.
#if defined(RGBID_TAIL_PROBE) && RGBID_TAIL_PROBE
#include <cstdio>  // diagnostic build only
#endif
#include <cstdlib>
#include "scale_core.cuh"
#include "bulk_copy.cuh"

namespace rgbid {

namespace {

#ifndef RGBID_BUILD_THREADS
#define RGBID_BUILD_THREADS 256
#endif
#ifndef RGBID_BUILD_MINBLOCKS
#define RGBID_BUILD_MINBLOCKS 2
#endif
constexpr int kBuildThreads = RGBID_BUILD_THREADS;
constexpr int kBuildMinBlocks = RGBID_BUILD_MINBLOCKS;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kAcc = 27;
constexpr int kAccChi = 31;  // + [rho_int, n_int, rho_depthinv, n_depthinv]

struct PixelParams {
  float fx, fy, cx, cy;
  float inv_sigma_int, inv_sigma_depthinv;
  float inv_sigma2_int, inv_sigma2_depthinv;
  float bias_over_sigma_int, bias_over_sigma_depthinv;
  float nu_int, nu_depthinv;
  int mestimator, weighting, student_nu;
  int chi_mestimator;
};

// fast_log: the end-of-frame chi^2 is a reported statistic (compared at 1e-3): log via lg2.approx (2 ulp) instead of the
// ~20-instruction logf, twice per pixel of the covariance pass.  The CHI_SQUARED termination test compares consecutive
// RMSEs and keeps the precise logarithm.
__device__ __forceinline__ float chi_rho_dev(float e, int mest, bool fast_log = false)
{
  float rho = (e * e) / 2.f;
  if (mest == RGBID_HUBER) { if (fabsf(e) > 1.345f) rho = 1.345f * (fabsf(e) - 1.345f / 2.f); }
  else if (mest == RGBID_TUKEY) {
    if (fabsf(e) < 4.685f) {
      float a1 = (e / 4.685f) * (e / 4.685f);
      float a2 = (1.f - a1) * (1.f - a1) * (1.f - a1);
      rho = ((4.685f * 4.685f) / 6.f) * (1.f - a2);
    } else rho = (4.685f * 4.685f) / 6.f;
  } else if (mest == RGBID_STUDENT) rho = ((5.f + 1.f) / 2.f) * (fast_log ? __logf(1.f + (e * e) / 5.f) : logf(1.f + (e * e) / 5.f));
  return rho;
}

// One keyframe pixel of computeSystemGridStride / computeStudentNuSystemGridStride
// (src/cuda/estimate_VO.cu:176-262 constraints, :295-329 / :384-418 weights + accumulation).
//
// The reference scales each Jacobian row and residual by 1/sigma and accumulates w * row_i * row_j.  Here the
// rows stay un-normalised and 1/sigma^2 is folded into the weight (s = w / sigma^2), which is the same sum with
// 12 fewer multiplies per pixel; each of the 27 terms is two FMAs straight into the accumulator.
template <bool CHI>
__device__ __forceinline__ void accumulate_pixel(float* __restrict__ acc, int x, int y, float w0, float i0, float gwx,
                                                 float gwy, float gix, float giy, float w1, float i1,
                                                 const PixelParams& pp)
{
  const float px = (__int2float_rn(x) - pp.cx) / pp.fx;
  const float py = (__int2float_rn(y) - pp.cy) / pp.fy;

  float rd[6], ri[6];
  float err_d = 0.f, err_i = 0.f, s_d = 0.f, s_i = 0.f, wgt_d = 0.f, wgt_i = 0.f;
  const bool valid_d = !(isnan(w0) || isnan(w1) || isnan(gwx) || isnan(gwy));
  const bool valid_i = !(isnan(w0) || isnan(i0) || isnan(i1) || isnan(gix) || isnan(giy));

  // invDepthConstraint (estimate_VO.cu:214-262)
  if (valid_d) {
    float g0 = gwx * pp.fx, g1 = gwy * pp.fy;
    float g2 = -(g0 * px + g1 * py);
    float inv_w0 = 1.f / w0;
    float n0 = g0 * inv_w0, n1 = g1 * inv_w0, n2 = g2 * inv_w0 + 1.f;
    // |n . p| / (|n| |p|) with p = (px, py, 1)
    float n_factor = fabsf(n0 * px + n1 * py + n2) * rsqrtf((n0 * n0 + n1 * n1 + n2 * n2) * (px * px + py * py + 1.f));
    float h2 = g2 + w1;
    rd[0] = g0 * w0; rd[1] = g1 * w0; rd[2] = g2 * w0 + w0 * w1;
    rd[3] = h2 * py - g1; rd[4] = g0 - h2 * px; rd[5] = g1 * px - g0 * py;  // -(g' x p), g' = (g0, g1, h2)
    err_d = w0 - w1;                                                        // -(w1 - w0)
    float eu = err_d * pp.inv_sigma_depthinv - pp.bias_over_sigma_depthinv;
    wgt_d = pp.student_nu ? (pp.nu_depthinv + 1.f) / (pp.nu_depthinv + eu * eu) : mest_weight(eu, pp.mestimator);
    if (pp.weighting == RGBID_PHOT_ONLY) wgt_d = 0.f;
    s_d = n_factor * wgt_d * pp.inv_sigma2_depthinv;
  }
  // intensityConstraint (estimate_VO.cu:176-212)
  if (valid_i) {
    float g0 = gix * pp.fx, g1 = giy * pp.fy;
    float g2 = -(g0 * px + g1 * py);
    ri[0] = g0 * w0; ri[1] = g1 * w0; ri[2] = g2 * w0;
    ri[3] = g2 * py - g1; ri[4] = g0 - g2 * px; ri[5] = g1 * px - g0 * py;  // -(g x p)
    err_i = i0 - i1;
    float eu = err_i * pp.inv_sigma_int - pp.bias_over_sigma_int;
    wgt_i = pp.student_nu ? (pp.nu_int + 1.f) / (pp.nu_int + eu * eu) : mest_weight(eu, pp.mestimator);
    if (pp.weighting == RGBID_GEOM_ONLY) wgt_i = 0.f;
    if (pp.weighting == RGBID_MIN_WEIGHT) wgt_i = fminf(wgt_d, wgt_i);  // wgt_d is 0 if the depth row is invalid
    s_i = wgt_i * pp.inv_sigma2_int;
  }
  if (CHI) {
    // end-of-frame chi^2 on all finite full-resolution residuals (src/visodo.cpp:1411-1414,
    // sigmaFuncs.cu:137-150, 541-611) with the reference scales 5 / 0.0025
    if (pp.chi_mestimator >= 0) {
      float ei = (i1 - i0) / 5.f, ed = (w1 - w0) / 0.0025f;
      if (!(isnan(ei) || isinf(ei))) { acc[27] += chi_rho_dev(ei, pp.chi_mestimator); acc[28] += 1.f; }
      if (!(isnan(ed) || isinf(ed))) { acc[29] += chi_rho_dev(ed, pp.chi_mestimator); acc[30] += 1.f; }
    }
  }
  // an invalid (or zero-weight) constraint contributes nothing: the reference multiplies stale rows by weight 0
  if (s_i != 0.f) {
    int shift = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float si = s_i * ri[i];
#pragma unroll
      for (int j = i; j < 6; ++j) { acc[shift] = fmaf(si, ri[j], acc[shift]); ++shift; }
      acc[shift] = fmaf(si, err_i, acc[shift]); ++shift;
    }
  }
  if (s_d != 0.f) {
    int shift = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float sd = s_d * rd[i];
#pragma unroll
      for (int j = i; j < 6; ++j) { acc[shift] = fmaf(sd, rd[j], acc[shift]); ++shift; }
      acc[shift] = fmaf(sd, err_d, acc[shift]); ++shift;
    }
  }
}

struct BuildShared {
  double warp_part[kBuildWarps][kAccChi];
  double total[kAccChi];
  Proj proj;
};

// Block reduce + publish the per-CTA partial + elect the last CTA of this pair + fixed-order final sum.
// Returns true in the threads of WARP 0 of the last CTA (false everywhere else), with sh.total[0..NACC) holding the
// pair's sums.
//
// wscratch: NACC x 32 floats of shared memory owned by the calling WARP.  The lane sums go through it instead of
// through shuffles: a double butterfly costs 10 SHFL + 5 DADD per value (27 values: 400 instructions per warp, at one
// SHFL per clock per SM -- RGBID_TAIL_PROBE showed ~14 000 clocks between the end of the pixel loop and the election);
// here lane k adds the 32 lane values of sum k in double (4 chains of 8, rotated start: conflict-free banks), fixed order.
//
// After the one CTA-wide barrier everything is done by warp 0 alone -- cross-warp sum, partial store, ONE cumulative
// fence + ticket by lane 0 behind a warp barrier (instead of a fence in each of the 27 storing threads, a second fence
// in all 256 and three more CTA barriers: RGBID_TAIL_PROBE showed 12 000 clocks from the end of the loop to the start of
// the solve), then, in the last CTA, the fixed-order final sum with every lane's loads in flight together.
template <int NACC>
__device__ __forceinline__ bool reduce_and_elect(BuildShared& sh, const float* acc, float* wscratch,
                                                 double* __restrict__ partials, int partial_stride,
                                                 unsigned int* __restrict__ counter, int nblk, int blk)
{
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) wscratch[k * 32 + lane] = acc[k];
  __syncwarp();
  if (lane < NACC) {
    const float* row = wscratch + lane * 32;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      s0 += (double)row[(j + 0 + lane) & 31];
      s1 += (double)row[(j + 1 + lane) & 31];
      s2 += (double)row[(j + 2 + lane) & 31];
      s3 += (double)row[(j + 3 + lane) & 31];
    }
    sh.warp_part[wid][lane] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (wid != 0) return false;
  if (lane < NACC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kBuildWarps; ++w) v += sh.warp_part[w][lane];
    partials[(size_t)blk * partial_stride + lane] = v;
  }
  __syncwarp();  // orders the 27 stores before lane 0's fence (the fence is cumulative)
  unsigned ticket = 0;
  if (lane == 0) {
    __threadfence();
    ticket = atomicAdd(counter, 1u);
    __threadfence();  // the other CTAs' partials are visible to the loads below (ordered behind the warp barrier)
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != (unsigned)(nblk - 1)) return false;
  // fixed-order final sum: lane k owns value k; four independent chains over the CTAs, combined in order
  if (lane < NACC) {
    const double* col = partials + lane;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int c = 0;
#pragma unroll 2
    for (; c + 4 <= nblk; c += 4) {
      const double a0 = __ldcg(col + (size_t)(c + 0) * partial_stride), a1 = __ldcg(col + (size_t)(c + 1) * partial_stride);
      const double a2 = __ldcg(col + (size_t)(c + 2) * partial_stride), a3 = __ldcg(col + (size_t)(c + 3) * partial_stride);
      v0 += a0; v1 += a1; v2 += a2; v3 += a3;
    }
    for (; c < nblk; ++c) v0 += __ldcg(col + (size_t)c * partial_stride);
    sh.total[lane] = (v0 + v1) + (v2 + v3);
  }
  if (lane == 0) *counter = 0u;  // ready for the next launch
  __syncwarp();
  return true;
}

// only: the one level the next launch reads (-1: all levels)
__device__ __forceinline__ void refresh_proj(GnState& st, const double* R, const double* t, int levels, float fx0,
                                             float fy0, float cx0, float cy0, int only = -1)
{
  // R is a product of rotations re-orthogonalised at every step: its inverse is its transpose to rounding (the
  // reference calls Eigen's general inverse(), src/visodo.cpp:1066-1067 -- a division chain this tail cannot afford)
  double Ri[9], ti[3];
  mat3_transpose(R, Ri);
  mat3_vec(Ri, t, ti);
  ti[0] = -ti[0]; ti[1] = -ti[1]; ti[2] = -ti[2];
  for (int l = 0; l < levels; ++l) {
    if (only >= 0 && l != only) continue;
    float div = (float)(1 << l);  // Intr::operator()(level), src/internal.h:128-132
    projective_pose(Ri, ti, fx0 / div, fy0 / div, cx0 / div, cy0 / div, st.proj[l].r, st.proj[l].t);
  }
}

// Tail executed by one thread of the last CTA: the host part of one reference iteration
// (src/visodo.cpp:1242-1274 / src/keyframe_align.cpp:312-350), of the covariance pass (:1382-1415) or of the
// CHI_SQUARED termination test (:1134-1164).  Three separate functions, so that the per-iteration path is small (it is
// cold code executed once per launch by a single thread: RGBID_TAIL_PROBE measured ~11 000 clocks for ~800 instructions)
// and gets its own register allocation instead of inheriting the spills of the 6x6 Gauss-Jordan inverse.
#if RGBID_TAIL_PROBE
__device__ long long g_tail_stamp[8];  // diagnostic build: clock64 at the stages of gn_tail_update, pair 0 only
__device__ const void* g_tail_probe_state;
#define RGBID_STAMP(k) if ((const void*)&st == g_tail_probe_state) g_tail_stamp[k] = clock64()
#else
#define RGBID_STAMP(k)
#endif

// The per-iteration tail in three separate (noinline) stages.  They are ~800 instructions of code that was last executed
// a launch ago: RGBID_TAIL_PROBE shows every stage 2.5-4x slower on its first execution than when repeated (solve
// 4 200-5 500 clocks against 1 700-2 100, update 1 600-2 800 / 620, commit 4 200-6 400 / 1 100).  Letting idle warps of
// every CTA pre-execute the stages on dummy data behind the reduction's barrier shortened the tail (17 500 -> 14 000
// clocks) but made the launch SLOWER (90.4 vs 84.7 us: 4 x 288 lanes of cold double-precision code at once) -- removed,
// profiles/r02_tail_probe.txt.
#ifndef RGBID_TAIL_INLINE
#define RGBID_TAIL_INLINE 1  // the three stages fall through behind the reduction (sequential instruction prefetch): 0.4-0.6 us per launch
#endif
#if RGBID_TAIL_INLINE
#define RGBID_STAGE_ATTR __forceinline__
#else
#define RGBID_STAGE_ATTR __noinline__
#endif
__device__ RGBID_STAGE_ATTR void gn_stage_solve(const double* tot, double* x) { llt_solve_packed(tot, x); }

__device__ RGBID_STAGE_ATTR bool gn_stage_update(const double* x, double* R, double* t) { return gn_update_lean(x, R, t); }

__device__ RGBID_STAGE_ATTR void gn_stage_commit(GnState& st, const double* Rin, const double* tin, const double* x, bool bad,
                                             const GnParams& P)
{
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = Rin[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = tin[i];
  if (P.dry_tail) {
    // everything is executed, but the pose is not committed and the projection goes to the slot of ANOTHER level than
    // the one being timed (rewritten by gn_init before any real run)
    refresh_proj(st, R, t, P.levels, P.fx0, P.fy0, P.cx0, P.cy0, (P.level + 1) % P.levels);
    return;
  }
  if (bad) {
    // lost: keep the previous pose, covariance 100 I (src/visodo.cpp:1265-1274)
    st.status = RGBID_ERR_NAN;
    for (int i = 0; i < 9; ++i) R[i] = st.R0[i];
    for (int i = 0; i < 3; ++i) t[i] = st.t0[i];
    for (int i = 0; i < 36; ++i) st.cov[i] = (i % 7 == 0) ? 100.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) st.R[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) st.t[i] = t[i];
  st.iters_done[P.sched_level] = P.sched_iter + 1;
  int next = P.next_level;
  if (P.conv_eps > 0.f) {
    // CONVERGENCE termination (BASELINE config 2): this update was small enough -> the level's remaining launches
    // are no-ops for this pair, and whichever level iterates next needs its projection
    const double n2 = (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) + (x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
    if (n2 < (double)P.conv_eps * (double)P.conv_eps) { st.skip_level = P.sched_level; next = -1; }
  }
  refresh_proj(st, R, t, P.levels, P.fx0, P.fy0, P.cx0, P.cy0, next);
}

__device__ __forceinline__ void gn_tail_update(GnState& st, const double* tot, const GnParams& P, double* x)
{
  RGBID_STAMP(0);
  // pose in registers: every access through `st` is a global load / store the compiler may not reorder
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = st.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = st.t[i];
#if RGBID_TAIL_PROBE
  if (R[0] + t[0] == 12345.678) st.status = 1;  // consume the loads before the stamp
#endif
  RGBID_STAMP(1);
  if (P.termination == RGBID_TERM_CHI_SQUARED) {  // the increment may have to be undone by the next test
#pragma unroll
    for (int i = 0; i < 9; ++i) st.Rprev[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) st.tprev[i] = t[i];
  }
  gn_stage_solve(tot, x);
  RGBID_STAMP(2);
  const bool bad = gn_stage_update(x, R, t);
  RGBID_STAMP(3);
  gn_stage_commit(st, R, t, x, bad, P);
  RGBID_STAMP(4);
}

// GnParams BY VALUE in the two out-of-line functions: a reference would take the address of the kernel parameter and make
// every thread of every CTA copy it to its local-memory frame in the prologue (160 B x 73 728 threads = the ~12 MB of
// DRAM writes per launch the round-1 captures show); by value the copy is made by the one thread that calls
__device__ __noinline__ void gn_tail_cov(GnState& st, const double* tot, const GnParams P, bool chi)
{
  if (P.compute_cov) {
    double A[36], bv[6];
    unpack_system(tot, A, bv);
    for (int i = 0; i < 36; ++i) st.lastA[i] = A[i];
    // the normal matrix is symmetric positive definite: Cholesky-based inverse in registers; the general Gauss-Jordan
    // (pivot search, row swaps through local memory, six divisions) only for a matrix that is not
    double Ai[36];
    if (inverse6_spd_packed(tot, Ai)) {
      for (int i = 0; i < 36; ++i) st.cov[i] = Ai[i];
    } else {
      inverse6(A, st.cov);
    }
  }
  if (chi && P.chi_mestimator >= 0) {
    // computeChiSquare host part, sigmaFuncs.cu:1286-1287
    float n = (float)(tot[28] + tot[30]);
    float chi2 = (float)(tot[27] + tot[29]) / n;
    float z = (chi2 - n) / sqrtf(2.f * n);
    st.chi_square = chi2; st.ndof = n; st.chi_test = 0.5f * (1.f + erff(z / sqrtf(2.f)));
    if (P.chi_test) {
      // CHI_SQUARED termination (src/visodo.cpp:1139-1163): RMSE of all level-0 residuals at the current pose
      const float rmse = sqrtf(chi2) / sqrtf(n);
      if (P.chi_test == 2 && rmse > st.rmse_prev) {
        // undo the previous increment and end the iterations of this level
        double R[9], t[3];
        for (int i = 0; i < 9; ++i) { R[i] = st.Rprev[i]; st.R[i] = R[i]; }
        for (int i = 0; i < 3; ++i) { t[i] = st.tprev[i]; st.t[i] = t[i]; }
        st.skip_level = P.sched_level;
        st.iters_done[P.sched_level] = P.sched_iter - 1;  // the undone iteration does not count
        refresh_proj(st, R, t, P.levels, P.fx0, P.fy0, P.cx0, P.cy0, -1);
      } else {
        st.rmse_prev = rmse;
      }
    }
  }
}

__device__ __noinline__ void gn_tail_trace(const GnState& st, const double* tot, const GnParams P, const ScaleState* sc,
                                           rgbid_iter_trace* __restrict__ trace, int b, const double* x)
{
  rgbid_iter_trace& T = trace[(size_t)b * P.trace_stride + P.iter_index];
  T.level = P.level; T.iter = P.iter_index;
  for (int i = 0; i < 27; ++i) T.sums27[i] = tot[i];
  if (sc != nullptr && P.use_scale) {
    T.sigma_int = sc->sigma_int; T.sigma_depthinv = sc->sigma_depthinv; T.bias_int = sc->bias_int;
    T.bias_depthinv = sc->bias_depthinv; T.nu_int = sc->nu_int; T.nu_depthinv = sc->nu_depthinv;
    T.irls_iters_int = sc->irls_iters_int; T.irls_iters_depthinv = sc->irls_iters_depthinv;
  } else {
    T.sigma_int = 5.f; T.sigma_depthinv = 0.0025f; T.bias_int = 0.f; T.bias_depthinv = 0.f;
    T.nu_int = 5.f; T.nu_depthinv = 5.f; T.irls_iters_int = 0; T.irls_iters_depthinv = 0;
  }
  for (int i = 0; i < 6; ++i) T.x[i] = x[i];
  for (int i = 0; i < 9; ++i) T.R[i] = st.R[i];
  for (int i = 0; i < 3; ++i) T.t[i] = st.t[i];
}

__device__ __forceinline__ void gn_tail(GnState& st, const double* tot, const GnParams& P, const ScaleState* sc,
                                        rgbid_iter_trace* __restrict__ trace, int b, bool chi)
{
  double x[6] = {0, 0, 0, 0, 0, 0};
  if (P.compute_cov || chi) gn_tail_cov(st, tot, P, chi);
  if (P.update_pose) gn_tail_update(st, tot, P, x);
  // the trace record (~60 stores of cold code at the very end of the launch) is only written when the caller of this
  // run asked for traces (GnState::trace_on, rgbid_aligner_set_trace)
  if (trace != nullptr && P.trace_stride > 0 && P.iter_index >= 0 && P.iter_index < P.trace_stride && st.trace_on)
    gn_tail_trace(st, tot, P, sc, trace, b, x);
}

// a launch is a no-op for a pair that is lost or whose level has been ended by cfg.termination; uniform over all CTAs
// of the pair because both fields are only written by the tail of an earlier launch
__device__ __forceinline__ bool pair_skipped(const GnState& st, const GnParams& P)
{
  return st.status != RGBID_OK || st.skip_level == P.sched_level;
}

__device__ __forceinline__ PixelParams make_pixel_params(const GnParams& P, const ScaleState* sc)
{
  PixelParams pp;
  pp.fx = P.fx; pp.fy = P.fy; pp.cx = P.cx; pp.cy = P.cy;
  float sigma_int = 5.f, sigma_d = 0.0025f, bias_int = 0.f, bias_d = 0.f, nu_int = 5.f, nu_d = 5.f;
  if (P.use_scale && sc != nullptr) {
    sigma_int = sc->sigma_int; sigma_d = sc->sigma_depthinv; bias_int = sc->bias_int; bias_d = sc->bias_depthinv;
    nu_int = sc->nu_int; nu_d = sc->nu_depthinv;
  }
  pp.inv_sigma_int = 1.f / sigma_int; pp.inv_sigma_depthinv = 1.f / sigma_d;
  pp.inv_sigma2_int = pp.inv_sigma_int * pp.inv_sigma_int;
  pp.inv_sigma2_depthinv = pp.inv_sigma_depthinv * pp.inv_sigma_depthinv;
  pp.bias_over_sigma_int = bias_int / sigma_int; pp.bias_over_sigma_depthinv = bias_d / sigma_d;
  pp.nu_int = nu_int; pp.nu_depthinv = nu_d;
  pp.mestimator = P.mestimator; pp.weighting = P.weighting; pp.student_nu = P.student_nu;
  pp.chi_mestimator = P.chi_mestimator;
  return pp;
}

// ------------------------------------------------------------------------------------------------------------
// gn_build_kernel.  grid = (ctas_per_pair, batch); VEC = pixels per thread step (4: float4 path, 1: scalar).
// ------------------------------------------------------------------------------------------------------------
template <int VEC, bool CHI, bool TEX>
__global__ void __launch_bounds__(kBuildThreads, kBuildMinBlocks)
    gn_build_kernel(const GnLevelMaps M, const GnParams P, GnState* __restrict__ states,
                    const ScaleState* __restrict__ scales, double* __restrict__ partials, int partial_stride,
                    unsigned int* __restrict__ counters, rgbid_iter_trace* __restrict__ trace)
{
  constexpr int NACC = CHI ? kAccChi : kAcc;
  const int b = blockIdx.y + P.first;
  GnState& st = states[b];
  grid_dep_wait();
  if (pair_skipped(st, P)) return;
  __shared__ BuildShared sh;
  if (threadIdx.x < 12) {
    const float* src = (const float*)&st.proj[P.level];
    ((float*)&sh.proj)[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  const Proj proj = sh.proj;
  const ScaleState* sc = scales ? &scales[b] : nullptr;
  const PixelParams pp = make_pixel_params(P, sc);
  const bool geom_is_warped = (P.mode == RGBID_MODE_TRACKER);

  float acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.f;

  CurFrame cur;
  cur.Wc = M.Wc.row(b, 0); cur.Ic = M.Ic.row(b, 0);
  cur.wpitch = M.Wc.pitch; cur.ipitch = M.Ic.pitch;
  cur.texW = TEX ? M.texW[b] : 0; cur.texI = TEX ? M.texI[b] : 0;
  const int cols = P.cols, rows = P.rows;
  const int upr = cols / VEC;  // units per row
  const int total = upr * rows;
  for (int u = blockIdx.x * kBuildThreads + threadIdx.x; u < total; u += gridDim.x * kBuildThreads) {
    const int y = u / upr, x0 = (u - y * upr) * VEC;
    float w0[VEC], i0[VEC], gwx[VEC], gwy[VEC], gix[VEC], giy[VEC];
    if (VEC == 4) {
      *(float4*)w0 = __ldg((const float4*)(M.W0.row(b, y) + x0));
      *(float4*)i0 = __ldg((const float4*)(M.I0.row(b, y) + x0));
      *(float4*)gwx = __ldg((const float4*)(M.gWx.row(b, y) + x0));
      *(float4*)gwy = __ldg((const float4*)(M.gWy.row(b, y) + x0));
      *(float4*)gix = __ldg((const float4*)(M.gIx.row(b, y) + x0));
      *(float4*)giy = __ldg((const float4*)(M.gIy.row(b, y) + x0));
    } else {
      w0[0] = __ldg(M.W0.row(b, y) + x0); i0[0] = __ldg(M.I0.row(b, y) + x0);
      gwx[0] = __ldg(M.gWx.row(b, y) + x0); gwy[0] = __ldg(M.gWy.row(b, y) + x0);
      gix[0] = __ldg(M.gIx.row(b, y) + x0); giy[0] = __ldg(M.gIy.row(b, y) + x0);
    }
    float w1[VEC], i1[VEC];
    if (P.prewarped) {  // uniform
      // WARP_ORDER = warpFirst above level 0: Wc / Ic hold pyrDown^level(warp_0(current frame)), src/visodo.cpp:1078-1105
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        w1[k] = __ldg(M.Wc.row(b, y) + x0 + k);
        i1[k] = __ldg(M.Ic.row(b, y) + x0 + k);
      }
    } else if (TEX) {
      // all inverse-depth fetches of this thread in flight, then all intensity fetches
      WarpCoord wc[VEC];
      float fetched[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) wc[k] = warp_stage1(proj, x0 + k, y, w0[k], cols, rows);
#pragma unroll
      for (int k = 0; k < VEC; ++k) fetched[k] = tex2D<float>(cur.texW, wc[k].xt, wc[k].yt);
#pragma unroll
      for (int k = 0; k < VEC; ++k) w1[k] = warp_stage2(proj, x0 + k, y, w0[k], fetched[k], wc[k], cols, rows, geom_is_warped);
#pragma unroll
      for (int k = 0; k < VEC; ++k) fetched[k] = tex2D<float>(cur.texI, wc[k].xt, wc[k].yt);
#pragma unroll
      for (int k = 0; k < VEC; ++k) i1[k] = warp_stage3(fetched[k], wc[k]);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        warp_pixel<false>(proj, x0 + k, y, w0[k], cur, cols, rows, geom_is_warped, w1[k], i1[k]);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      accumulate_pixel<CHI>(acc, x0 + k, y, w0[k], i0[k], gwx[k], gwy[k], gix[k], giy[k], w1[k], i1[k], pp);
  }
  grid_dep_launch();

  __shared__ float scratch[kBuildWarps][kAccChi * 32];
  if (!reduce_and_elect<NACC>(sh, acc, scratch[threadIdx.x >> 5], partials + (size_t)b * gridDim.x * partial_stride,
                              partial_stride, &counters[b], gridDim.x, blockIdx.x))
    return;
  if (threadIdx.x == 0) gn_tail(st, sh.total, P, sc, trace, b, CHI);
}

// ------------------------------------------------------------------------------------------------------------
// gn_build_fast_kernel: the shipped configuration (Student-t weights, INDEPENDENT weighting, texture gathers, flat
// keyframe maps) of gn_build_kernel, rebuilt around the measured limits of the B200 SM (tools/ubench/pipes.cu,
// texpat.cu): the path needs ~140 FP32 FMA-pipe instructions out of 200 per pixel, half of them lose a dispatch cycle to
// register-bank conflicts, and its two gathers per pixel keep the texture unit ~2/3 busy -- the pixel loop runs at 0.74
// of the HBM peak, limited by the dispatch port and the texture unit, not by DRAM (profiles/README.md, round 2).
// What it does differently from the generic kernel:
//   * the six keyframe maps arrive through the TMA engine: each warp owns two rings of shared-memory slots (W0:
//     kStagesW x 512 B, the five other maps: kStagesL x 2560 B), one elected lane issues 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) several chunks ahead and the warp waits on an mbarrier -- no CTA-wide barrier in
//     the loop, no staging registers, DRAM latency hidden;
//   * the texture gathers of chunk i + 1 / i + 2 are issued while the constraints of chunk i are accumulated
//     (software pipeline, unrolled by two so that no register set has to be copied);
//   * the projection is factored as  K R K^-1 (x, y, 1)^T / w + K t : the pixel-dependent part is computed once and
//     reused by the second projection of tracker mode (2 FMAs per coordinate instead of 5);
//   * the inverse-depth texture uses border addressing (0 outside), so "outside the image" falls out of the
//     reference's own `res > 0` test and only the intensity fetch needs an explicit in-image test (four compares on
//     the ALU pipe); validity is carried by NaN propagation into the weight and ONE compare per constraint;
//   * the 2 x 27 accumulations are FMAs predicated on that compare (no selects, no sanitising of the rows);
//   * the bulk copies carry an L2 evict-first policy (streamed once per launch), and in the warpFirst schedule above
//     level 0 (PREW) the pre-warped current-frame maps travel through the same ring instead of being gathered.
// Arithmetic differs from the generic kernel only in rounding (same formulas re-associated); the parity tests
// hold both against the oracle and the reference's own kernels.
// ------------------------------------------------------------------------------------------------------------
#ifndef RGBID_TAIL_PROBE
#define RGBID_TAIL_PROBE 0  // gn_build_fast_kernel: clock64 break-down of the last CTA (diagnostic build only)
#endif
#ifndef RGBID_PX
#define RGBID_PX 4
#endif
#ifndef RGBID_EVICT_FIRST
#define RGBID_EVICT_FIRST 1  // keyframe-map bulk copies carry an L2 evict-first policy (see bulk_copy.cuh)
#endif
constexpr int kPx = RGBID_PX;                   // pixels per lane and chunk (4: 128-bit shared-memory loads; 2: half the live gather state)
constexpr int kChunkPx = 32 * kPx;              // one warp-chunk
struct __align__(4 * kPx) LaneVec { float v[kPx]; };
constexpr int kChunkBytes = kChunkPx * 4;       // per map
// Two rings per warp: the keyframe inverse depth is needed by three pipeline stages (gather, second projection,
// constraints) and therefore lives two iterations longer than the other five maps.
#ifndef RGBID_STAGES_W
#define RGBID_STAGES_W 5
#endif
#ifndef RGBID_STAGES_L
#define RGBID_STAGES_L 3
#endif
constexpr int kStagesW = RGBID_STAGES_W;        // W0 ring: 5 x 512 B
constexpr int kStagesL = RGBID_STAGES_L;        // I0, gWx, gWy, gIx, gIy ring: 3 x 2560 B
// PREW (WARP_ORDER = warpFirst above level 0): the late ring also carries the pre-warped current-frame maps W1, I1
__host__ __device__ constexpr int late_maps(bool prew) { return prew ? 7 : 5; }
__host__ __device__ constexpr int late_bytes(bool prew) { return late_maps(prew) * kChunkBytes; }
__host__ __device__ constexpr int warp_ring_bytes(bool prew) { return kStagesW * kChunkBytes + kStagesL * late_bytes(prew); }  // 10 / 13 KiB
__host__ __device__ constexpr int fast_smem_bytes(bool prew) { return kBuildWarps * warp_ring_bytes(prew); }               // 80 / 104 KiB per CTA

struct FastGeom {
  int npx;        // rows * cols
  int nchunks;    // ceil(npx / 128)
  float colsf, rowsf, inv_cols;
};

// WITH_B = false (covariance pass): only the 21 J^T W J terms -- the pass inverts A and never looks at J^T W r
// (src/visodo.cpp:1382-1415), so the six residual terms stay zero
template <bool WITH_B>
__device__ __forceinline__ void accumulate_scalar(float* acc, float s, const float* r, float e, int flag)
{
  int shift = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float si = s * r[i];
#pragma unroll
    for (int j = i; j < 6; ++j) { pfma(acc[shift], si, r[j], flag); ++shift; }
    if (WITH_B) pfma(acc[shift], si, e, flag);
    ++shift;
  }
}

// CHIM: 0 no chi^2 sums, 1 chi^2 with the M-estimator chosen at run time, 2 chi^2 specialised for Student (the default)
template <bool TRACKER, int CHIM, bool PREW = false>
__global__ void __launch_bounds__(kBuildThreads, kBuildMinBlocks)
    gn_build_fast_kernel(const GnLevelMaps M, const GnParams P, const FastGeom G, GnState* __restrict__ states,
                         const ScaleState* __restrict__ scales, double* __restrict__ partials, int partial_stride,
                         unsigned int* __restrict__ counters, rgbid_iter_trace* __restrict__ trace)
{
  constexpr bool CHI = (CHIM != 0);
  constexpr int kLateBytes = late_bytes(PREW), kWarpRingBytes = warp_ring_bytes(PREW);
  static_assert(!PREW || (TRACKER && CHIM == 0), "pre-warped maps only exist in the tracker's warpFirst iterations");
  constexpr int NACC = CHI ? kAccChi : kAcc;
  const int chi_mest = (CHIM == 2) ? (int)RGBID_STUDENT : P.chi_mestimator;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ BuildShared sh;
  __shared__ __align__(8) unsigned long long bars[kBuildWarps * (kStagesW + kStagesL)];
  const int b = blockIdx.y + P.first;
  GnState& st = states[b];
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: bulk-copy operands stay in uniform registers
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kBuildWarps * (kStagesW + kStagesL); ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  // chunks [c_begin, c_end) of this stream belong to this CTA; warp w takes c_begin + w, + 8, ...
  const int c_begin = (int)(((long long)blockIdx.x * G.nchunks) / gridDim.x);
  const int c_end = (int)(((long long)(blockIdx.x + 1) * G.nchunks) / gridDim.x);
  const int my_first = c_begin + wid;
  const int my_n = (c_end - my_first + kBuildWarps - 1) / kBuildWarps;  // may be <= 0

  const uint32_t ringW = smem_u32(ring) + (uint32_t)wid * kWarpRingBytes;
  const uint32_t ringL = ringW + kStagesW * kChunkBytes;
  const uint32_t barW = smem_u32(&bars[wid * (kStagesW + kStagesL)]);
  const uint32_t barL = barW + kStagesW * 8u;
  const char* gW0 = (const char*)M.W0.row(b, 0);
  const char* gI0 = (const char*)M.I0.row(b, 0);
  const char* gWx = (const char*)M.gWx.row(b, 0);
  const char* gWy = (const char*)M.gWy.row(b, 0);
  const char* gIx = (const char*)M.gIx.row(b, 0);
  const char* gIy = (const char*)M.gIy.row(b, 0);
  const char* gW1 = (const char*)M.Wc.row(b, 0);  // PREW: pyrDown^level(warp_0(current frame)), src/visodo.cpp:1078-1105
  const char* gI1 = (const char*)M.Ic.row(b, 0);
  (void)gW1; (void)gI1;
  // (one elected lane) bulk copies of this warp's i-th chunk; the last chunk of a stream runs into the map's NaN
  // padding (see launch_gn_build), so every copy is a full 512 bytes
#if RGBID_EVICT_FIRST
  const uint64_t l2_stream = l2_policy_evict_first();
#define RGBID_G2S(dst, src, bytes, bar) bulk_g2s_hint(dst, src, bytes, bar, l2_stream)
#else
#define RGBID_G2S(dst, src, bytes, bar) bulk_g2s(dst, src, bytes, bar)
#endif
  auto issue_w = [&](int i) {
    const int s = i % kStagesW;
    const size_t off = (size_t)(my_first + i * kBuildWarps) * kChunkBytes;
    mbar_arrive_expect_tx(barW + (uint32_t)s * 8u, kChunkBytes);
    RGBID_G2S(ringW + (uint32_t)s * kChunkBytes, gW0 + off, kChunkBytes, barW + (uint32_t)s * 8u);
  };
  auto issue_l = [&](int i) {
    const int s = i % kStagesL;
    const size_t off = (size_t)(my_first + i * kBuildWarps) * kChunkBytes;
    const uint32_t dst = ringL + (uint32_t)s * kLateBytes, bar = barL + (uint32_t)s * 8u;
    mbar_arrive_expect_tx(bar, kLateBytes);
    RGBID_G2S(dst + 0 * kChunkBytes, gI0 + off, kChunkBytes, bar);
    RGBID_G2S(dst + 1 * kChunkBytes, gWx + off, kChunkBytes, bar);
    RGBID_G2S(dst + 2 * kChunkBytes, gWy + off, kChunkBytes, bar);
    RGBID_G2S(dst + 3 * kChunkBytes, gIx + off, kChunkBytes, bar);
    RGBID_G2S(dst + 4 * kChunkBytes, gIy + off, kChunkBytes, bar);
    if (PREW) {
      // written by the warp / pyramid kernels of this very iteration and read once: same streaming policy
      RGBID_G2S(dst + 5 * kChunkBytes, gW1 + off, kChunkBytes, bar);
      RGBID_G2S(dst + 6 * kChunkBytes, gI1 + off, kChunkBytes, bar);
    }
  };
#undef RGBID_G2S
  // the keyframe maps were written before this Gauss-Newton schedule started: the first bulk copies may be in flight
  // while the previous kernel (scale estimation, or the previous iteration's solve) is still finishing
  if (elect_one()) {
#pragma unroll
    for (int i = 0; i < kStagesW; ++i)
      if (i < my_n) issue_w(i);
    if (!PREW) {
#pragma unroll
      for (int i = 0; i < kStagesL; ++i)
        if (i < my_n) issue_l(i);
    }
  }
  grid_dep_wait();
  if (PREW) {
    // the pre-warped maps were written by the kernels of this very iteration: their copies wait for the dependency
    if (elect_one()) {
#pragma unroll
      for (int i = 0; i < kStagesL; ++i)
        if (i < my_n) issue_l(i);
    }
  }
#if RGBID_TAIL_PROBE
  const long long probe_t0 = clock64();  // the pixel loop is counted from the end of the dependency wait
#endif
  // a skipped pair leaves with its bulk copies in flight: they land in this CTA's own shared memory, which stays
  // allocated until the copies have completed
  const bool skipped = pair_skipped(st, P);
  if (tid < 12 && !skipped) ((float*)&sh.proj)[tid] = ((const float*)&st.proj[P.level])[tid];
  __syncthreads();
  if (skipped) {
    // drain: wait for every copy this warp issued before the CTA exits
    for (int i = 0; i < kStagesW && i < my_n; ++i) mbar_wait(barW + (uint32_t)i * 8u, 0u);
    for (int i = 0; i < kStagesL && i < my_n; ++i) mbar_wait(barL + (uint32_t)i * 8u, 0u);
    return;
  }

  // per-stream constants
  const float r0 = sh.proj.r[0], r3 = sh.proj.r[3], r6 = sh.proj.r[6];
  const float t0 = sh.proj.t[0], t1 = sh.proj.t[1], tz = sh.proj.t[2];
  const ScaleState* sc = scales ? &scales[b] : nullptr;
  float sigma_i = 5.f, sigma_d = 0.0025f, bias_i = 0.f, bias_d = 0.f, nu_i = 5.f, nu_d = 5.f;
  if (P.use_scale && sc != nullptr) {
    sigma_i = sc->sigma_int; sigma_d = sc->sigma_depthinv; bias_i = sc->bias_int; bias_d = sc->bias_depthinv;
    nu_i = sc->nu_int; nu_d = sc->nu_depthinv;
  }
  if (!P.student_nu) { nu_i = 5.f; nu_d = 5.f; }  // computeWeight(STUDENT), estimate_VO.cu:160-163
  const float is_i = 1.f / sigma_i, is_d = 1.f / sigma_d;
  const float bos_i = bias_i / sigma_i, bos_d = bias_d / sigma_d;
  const float c_i = (nu_i + 1.f) * (is_i * is_i), c_d = (nu_d + 1.f) * (is_d * is_d);  // (nu + 1) / sigma^2
  const float ifx = 1.f / P.fx, ify = 1.f / P.fy;
  const cudaTextureObject_t texW = PREW ? 0 : M.texW[b], texI = PREW ? 0 : M.texI[b];

  float accs[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) accs[k] = 0.f;
  float chi[4] = {0.f, 0.f, 0.f, 0.f};

  // Software pipeline over this warp's chunks.  With 128 registers only four warps share a scheduler, so the two
  // dependent texture latencies of a pixel (inverse depth -> warped inverse depth -> intensity, tracker mode) are
  // covered by independent arithmetic of the SAME warp.  Iteration i
  //   S1  chunk i + 1 : warped inverse depth from the gathered one, second projection, intensity gather issued
  //   S2  chunk i + 2 : first projection, inverse-depth gather issued
  //   S3  chunk i     : both constraints + 2 x 27 accumulations
  // and 12 + 4 registers travel between iterations.  KeyframeAlign mode samples the intensity where it sampled the
  // inverse depth, so both gathers are issued in S2 and S1 disappears.
  struct ChunkGeom { float xf0, yf, rcx, rcy, rcz; };
  const float r1 = sh.proj.r[1], r2 = sh.proj.r[2], r4 = sh.proj.r[kPx], r5 = sh.proj.r[5], r7 = sh.proj.r[7], r8 = sh.proj.r[8];
  const float idxf0 = __int2float_rn(my_first * kChunkPx + lane * kPx) + 0.5f;  // pixel index + 0.5, exact below 2^23
  auto chunk_geom = [&](int i) {
    ChunkGeom g;
    const float idxh = fmaf(__int2float_rn(i), (float)(kBuildWarps * kChunkPx), idxf0);
    g.yf = floorf(idxh * G.inv_cols);  // exact for rows * cols <= 2.5 M (checked by the launcher)
    g.xf0 = fmaf(-g.yf, G.colsf, idxh) - 0.5f;
    g.rcx = fmaf(r1, g.yf, r2);
    g.rcy = fmaf(r4, g.yf, r5);
    g.rcz = fmaf(r7, g.yf, r8);
    return g;
  };
  // plain shared-memory loads (not volatile asm): the mbarrier waits carry a memory clobber, so the loads cannot move
  // above them, and the compiler is free to schedule them and to pick destination registers that are not the target of
  // a texture fetch still in flight (a volatile LDS stalled 9 % of all samples on exactly that hazard)
  const unsigned char* ringW_p = ring + (size_t)wid * kWarpRingBytes + (size_t)lane * (4 * kPx);
  const unsigned char* ringL_p = ringW_p + kStagesW * kChunkBytes;
  auto lds_w0 = [&](int i, float* w0) { *(LaneVec*)w0 = *(const LaneVec*)(ringW_p + (size_t)(i % kStagesW) * kChunkBytes); };
  // floor(xt) in [0, cols) && floor(yt) in [0, rows) (warping_registration.cu:490-491) as four float compares -- exact,
  // false for the NaN coordinates of an invalid geometry, and on the ALU pipe (this kernel is bound by the FMA pipe);
  // returned as 0 / NaN so that it can be added to the sample later
  auto in_image_nan = [&](float xt, float yt) {
    return (xt >= 0.f && xt < G.colsf && yt >= 0.f && yt < G.rowsf) ? 0.f : qnanf();
  };
  // S2: first projection (geometry = keyframe inverse depth) of chunk i and its gather(s)
  auto gather = [&](int i, float* w2, float* wcs, float* i1, float* pinf) {
    mbar_wait(barW + (uint32_t)(i % kStagesW) * 8u, (uint32_t)(i / kStagesW) & 1u);
    float w0[kPx], xt[kPx], yt[kPx];
    lds_w0(i, w0);
    const ChunkGeom g = chunk_geom(i);
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      const float xf = g.xf0 + (float)k;
      const float z = 1.f / w0[k];
      const float Xc = fmaf(fmaf(r0, xf, g.rcx), z, t0), Yc = fmaf(fmaf(r3, xf, g.rcy), z, t1);
      const float Zc = fmaf(fmaf(r6, xf, g.rcz), z, tz);
      const float wc = 1.f / Zc;
      wcs[k] = wc;
      xt[k] = fmaf(Xc, wc, 0.5f); yt[k] = fmaf(Yc, wc, 0.5f);
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) w2[k] = tex2D<float>(texW, xt[k], yt[k]);  // border addressing: 0 outside
    if (!TRACKER) {
#pragma unroll
      for (int k = 0; k < kPx; ++k) i1[k] = tex2D<float>(texI, xt[k], yt[k]);
#pragma unroll
      for (int k = 0; k < kPx; ++k) pinf[k] = in_image_nan(xt[k], yt[k]);
    }
  };
  // trafo3DKernelInvDepthGridStride, warping_registration.cu:533-538, operation for operation (v1z is
  // algebraically (K R K^-1 (x, y, 1))_z, but the residual w0 - w1 is small against w and feels the rounding);
  // a / b is written a * (1 / b): what div.approx does for |b| < 2^126, without its range fix-up
  auto warped_invdepth = [&](float w0, float wc, float w2) {
    const float v1z = (1.f / wc - tz) * w0;
    const float res = (v1z * (1.f / (1.f - w2 * tz))) * w2;
    return (res > 0.f) ? res : qnanf();  // NaN, <= 0 and the border value 0 are all invalid
  };
  // S1 (tracker mode): intensity is sampled where the keyframe pixel lands with the WARPED inverse depth as geometry
  // (src/visodo.cpp:1121-1126)
  auto second_projection = [&](int i, const float* w2, const float* wcs, float* w1, float* i1, float* pinf) {
    const ChunkGeom g = chunk_geom(i);
    float w0[kPx], xt[kPx], yt[kPx];
    lds_w0(i, w0);
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      const float xf = g.xf0 + (float)k;
      w1[k] = warped_invdepth(w0[k], wcs[k], w2[k]);
      const float zz = 1.f / w1[k];
      const float Xc = fmaf(fmaf(r0, xf, g.rcx), zz, t0), Yc = fmaf(fmaf(r3, xf, g.rcy), zz, t1);
      const float wc1 = 1.f / fmaf(fmaf(r6, xf, g.rcz), zz, tz);
      xt[k] = fmaf(Xc, wc1, 0.5f); yt[k] = fmaf(Yc, wc1, 0.5f);
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) i1[k] = tex2D<float>(texI, xt[k], yt[k]);
#pragma unroll
    for (int k = 0; k < kPx; ++k) pinf[k] = in_image_nan(xt[k], yt[k]);
  };

  float w2n[kPx], wcn[kPx];                  // S2 -> S1 (tracker) / S3 (align): gathered inverse depth, 1 / Zc
  // S1 -> S3: warped inverse depth, raw intensity sample, 0 / NaN in-image flag.  Two sets, used alternately by a
  // loop unrolled by two, so that nothing has to be copied between iterations.
  float w1a[kPx], i1a[kPx], pina[kPx], w1b[kPx], i1b[kPx], pinb[kPx];
  if (my_n > 0 && !PREW) {
    if (TRACKER) {
      gather(0, w2n, wcn, nullptr, nullptr);
      second_projection(0, w2n, wcn, w1a, i1a, pina);
      if (my_n > 1) gather(1, w2n, wcn, nullptr, nullptr);
    } else {
      gather(0, w2n, wcn, i1a, pina);
    }
  }
  // one iteration: S1 + S2 fill the `n` set for chunk i + 1, S3 consumes the `c` set of chunk i
  auto iteration = [&](int i, float* w1c, float* i1c, float* pinc, float* w1n, float* i1n, float* pinn) {
    float* w1 = w1c; float* i1 = i1c; float* pin = pinc;
    if (PREW) {
      // nothing to gather: W1 / I1 arrive with the keyframe maps (read below, once the late ring has landed)
    } else if (TRACKER) {
      if (i + 1 < my_n) second_projection(i + 1, w2n, wcn, w1n, i1n, pinn);  // warp-uniform
      if (i + 2 < my_n) gather(i + 2, w2n, wcn, nullptr, nullptr);
    } else {
      float w0a[kPx];
      lds_w0(i, w0a);
#pragma unroll
      for (int k = 0; k < kPx; ++k) w1c[k] = warped_invdepth(w0a[k], wcn[k], w2n[k]);
      if (i + 1 < my_n) gather(i + 1, w2n, wcn, i1n, pinn);
    }

    // --- S3: inverse-depth constraint + accumulation -----------------------------------------------------------
    const ChunkGeom g = chunk_geom(i);
    const float py = (g.yf - P.cy) * ify;
    const float py2p1 = fmaf(py, py, 1.f);
    const unsigned char* buf = ringL_p + (size_t)(i % kStagesL) * kLateBytes;
    mbar_wait(barL + (uint32_t)(i % kStagesL) * 8u, (uint32_t)(i / kStagesL) & 1u);
    float w0[kPx], gwx[kPx], gwy[kPx];
    if (PREW) mbar_wait(barW + (uint32_t)(i % kStagesW) * 8u, (uint32_t)(i / kStagesW) & 1u);  // no gather stage waited for it
    lds_w0(i, w0);
    *(LaneVec*)gwx = *(const LaneVec*)(buf + 1 * kChunkBytes);
    *(LaneVec*)gwy = *(const LaneVec*)(buf + 2 * kChunkBytes);
    if (PREW) {
      *(LaneVec*)w1 = *(const LaneVec*)(buf + 5 * kChunkBytes);
      *(LaneVec*)i1 = *(const LaneVec*)(buf + 6 * kChunkBytes);
    }
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      const float xf = g.xf0 + (float)k;
      const float px = (xf - P.cx) * ifx;
      // invDepthConstraint (estimate_VO.cu:214-262).  The reference's n = (g0, g1, g2) / w0 + (0, 0, 1) satisfies
      // n . p = 1 identically (g2 = -(g0 px + g1 py)), so n_factor = |n . p| / (|n| |p|) = |w0| / (|m| |p|) with
      // m = (g0, g1, g2 + w0).
      const float gd0 = gwx[k] * P.fx, gd1 = gwy[k] * P.fy;
      const float gd2 = -fmaf(gd0, px, gd1 * py);
      const float m2 = gd2 + w0[k];
      const float mm = fmaf(gd0, gd0, fmaf(gd1, gd1, m2 * m2));
      const float nf = fabsf(w0[k]) * rsqrtf(mm * fmaf(px, px, py2p1));
      const float h2 = gd2 + w1[k];
      float rd[6];
      rd[0] = gd0 * w0[k]; rd[1] = gd1 * w0[k]; rd[2] = h2 * w0[k];
      rd[3] = fmaf(h2, py, -gd1); rd[kPx] = fmaf(-h2, px, gd0); rd[5] = fmaf(gd1, px, -(gd0 * py));
      const float ed = w0[k] - w1[k];
      const float eud = fmaf(ed, is_d, -bos_d);
      const float sd = nf * (c_d * (1.f / fmaf(eud, eud, nu_d)));  // NaN if any of w0, w1, gwx, gwy is NaN
      const int fd = (sd > 0.f);
      if (CHI) {
        // end-of-frame chi^2 on all finite full-resolution residuals (src/visodo.cpp:1411-1414,
        // sigmaFuncs.cu:137-150, 541-611) with the reference scales 5 / 0.0025
        if (P.chi_mestimator >= 0) {
          const float cd = (w1[k] - w0[k]) / 0.0025f;
          if (!(isnan(cd) || isinf(cd))) { chi[2] += chi_rho_dev(cd, chi_mest, P.chi_test == 0); chi[3] += 1.f; }
        }
      }
      accumulate_scalar<!CHI>(accs, sd, rd, ed, fd);
    }

    // --- S3: intensity constraint + accumulation ---------------------------------------------------------------
    float i0[kPx], gix[kPx], giy[kPx];
    *(LaneVec*)i0 = *(const LaneVec*)(buf + 0 * kChunkBytes);
    *(LaneVec*)gix = *(const LaneVec*)(buf + 3 * kChunkBytes);
    *(LaneVec*)giy = *(const LaneVec*)(buf + 4 * kChunkBytes);
#pragma unroll
    for (int k = 0; k < kPx; ++k) {
      const float xf = g.xf0 + (float)k;
      const float px = (xf - P.cx) * ifx;
      // max(0, min(r, 255)): a NaN sample becomes 255 like in the reference (warping_registration.cu:493-494);
      // + 0 / NaN: outside the image or invalid geometry
      // (PREW: the warp kernel has clamped already and wrote NaN where the sample is invalid)
      const float i1v = PREW ? i1[k] : fmaxf(0.f, fminf(i1[k], 255.f)) + pin[k];
      // intensityConstraint (estimate_VO.cu:176-212)
      const float gi0 = gix[k] * P.fx, gi1 = giy[k] * P.fy;
      const float gi2 = -fmaf(gi0, px, gi1 * py);
      float ri[6];
      ri[0] = gi0 * w0[k]; ri[1] = gi1 * w0[k]; ri[2] = gi2 * w0[k];
      ri[3] = fmaf(gi2, py, -gi1); ri[kPx] = fmaf(-gi2, px, gi0); ri[5] = fmaf(gi1, px, -(gi0 * py));
      const float ei = i0[k] - i1v;
      const float eui = fmaf(ei, is_i, -bos_i);
      float si = c_i * (1.f / fmaf(eui, eui, nu_i));
      si = fmaf(0.f, gi2, si);  // NaN gradients invalidate the row (a NaN w0 gives a NaN i1, i0 and i1 enter ei)
      const int fi = (si > 0.f);
      if (CHI) {
        if (P.chi_mestimator >= 0) {
          const float ci = (i1v - i0[k]) / 5.f;
          if (!(isnan(ci) || isinf(ci))) { chi[0] += chi_rho_dev(ci, chi_mest, P.chi_test == 0); chi[1] += 1.f; }
        }
      }
      accumulate_scalar<!CHI>(accs, si, ri, ei, fi);
    }
    __syncwarp();
    if (elect_one()) {
      if (i + kStagesW < my_n) issue_w(i + kStagesW);
      if (i + kStagesL < my_n) issue_l(i + kStagesL);
    }
  };
  for (int i = 0; i < my_n; i += 2) {
    iteration(i, w1a, i1a, pina, w1b, i1b, pinb);
    if (i + 1 < my_n) iteration(i + 1, w1b, i1b, pinb, w1a, i1a, pina);
  }
  grid_dep_launch();  // the next kernel's CTAs may take the slots this grid frees while its last CTAs reduce and solve

  float acc[NACC];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = accs[k];
  if (CHI) { acc[27] = chi[0]; acc[28] = chi[1]; acc[29] = chi[2]; acc[30] = chi[3]; }
  // every bulk copy this warp issued has been waited for and consumed: its ring (10 KiB) is free for the lane sums
  float* wscratch = (float*)(ring + (size_t)wid * kWarpRingBytes);
  static_assert(kWarpRingBytes >= kAccChi * 32 * (int)sizeof(float), "ring too small for the lane-sum scratch");

#if RGBID_TAIL_PROBE
  const long long probe_t1 = clock64();
#endif
  const bool last = reduce_and_elect<NACC>(sh, acc, wscratch, partials + (size_t)b * gridDim.x * partial_stride,
                                           partial_stride, &counters[b], gridDim.x, blockIdx.x);
  if (!last) return;
#if RGBID_TAIL_PROBE
  // -DRGBID_TAIL_PROBE=1 (diagnostic build, tools/scale_round_probe.py): clocks of the last CTA of pair 0 -- pixel loop
  // (from the end of the dependency wait), CTA reduction + election + final sum, serial tail and its stages
  if (b == 0 && threadIdx.x == 0) g_tail_probe_state = &st;
  const long long probe_t2 = clock64();
  if (threadIdx.x == 0) {
    gn_tail(st, sh.total, P, sc, trace, b, CHI);
    const long long probe_t3 = clock64();
    if (b == 0) {
      printf("tail probe level %d iter %d cta %d | loop %lld reduce+elect %lld tail %lld\n", P.level, P.iter_index,
             (int)blockIdx.x, probe_t1 - probe_t0, probe_t2 - probe_t1, probe_t3 - probe_t2);
      printf("   tail stages | call %lld loads %lld solve %lld update %lld commit %lld return %lld\n", g_tail_stamp[0] - probe_t2,
             g_tail_stamp[1] - g_tail_stamp[0], g_tail_stamp[2] - g_tail_stamp[1], g_tail_stamp[3] - g_tail_stamp[2],
             g_tail_stamp[4] - g_tail_stamp[3], probe_t3 - g_tail_stamp[4]);
    }
  }
#else
  if (threadIdx.x == 0) gn_tail(st, sh.total, P, sc, trace, b, CHI);
#endif
}

// ------------------------------------------------------------------------------------------------------------
// Un-fused drop-in for buildSystem(StudentNu)GridStride on pre-warped maps (one pair)
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kBuildThreads, 2)
    build_system_kernel(ImgB W0, ImgB I0, ImgB gWx, ImgB gWy, ImgB gIx, ImgB gIy, ImgB W1, ImgB I1, PixelParams pp,
                        double* __restrict__ partials, unsigned int* __restrict__ counter, double* __restrict__ out27)
{
  __shared__ BuildShared sh;
  float acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.f;
  const int upr = W0.cols / VEC, total = upr * W0.rows;
  for (int u = blockIdx.x * kBuildThreads + threadIdx.x; u < total; u += gridDim.x * kBuildThreads) {
    const int y = u / upr, x0 = (u - y * upr) * VEC;
    float w0[VEC], i0[VEC], gwx[VEC], gwy[VEC], gix[VEC], giy[VEC], w1[VEC], i1[VEC];
    if (VEC == 4) {
      *(float4*)w0 = __ldg((const float4*)(W0.row(0, y) + x0));
      *(float4*)i0 = __ldg((const float4*)(I0.row(0, y) + x0));
      *(float4*)gwx = __ldg((const float4*)(gWx.row(0, y) + x0));
      *(float4*)gwy = __ldg((const float4*)(gWy.row(0, y) + x0));
      *(float4*)gix = __ldg((const float4*)(gIx.row(0, y) + x0));
      *(float4*)giy = __ldg((const float4*)(gIy.row(0, y) + x0));
      *(float4*)w1 = __ldg((const float4*)(W1.row(0, y) + x0));
      *(float4*)i1 = __ldg((const float4*)(I1.row(0, y) + x0));
    } else {
      w0[0] = W0.row(0, y)[x0]; i0[0] = I0.row(0, y)[x0]; gwx[0] = gWx.row(0, y)[x0]; gwy[0] = gWy.row(0, y)[x0];
      gix[0] = gIx.row(0, y)[x0]; giy[0] = gIy.row(0, y)[x0]; w1[0] = W1.row(0, y)[x0]; i1[0] = I1.row(0, y)[x0];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      accumulate_pixel<false>(acc, x0 + k, y, w0[k], i0[k], gwx[k], gwy[k], gix[k], giy[k], w1[k], i1[k], pp);
  }
  __shared__ float scratch[kBuildWarps][kAcc * 32];
  if (!reduce_and_elect<kAcc>(sh, acc, scratch[threadIdx.x >> 5], partials, kAccChi, counter, gridDim.x, blockIdx.x)) return;
  if (threadIdx.x < kAcc) out27[threadIdx.x] = sh.total[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------------------
// gn_scale_kernel: fused warp + residual sampling + scale estimation; one 8-CTA cluster per pair.
// Sampling geometry of computeErrorGridStride (sigmaFuncs.cu:711-747): sample (s y, s x) -> index y*kept_cols+x.
// ------------------------------------------------------------------------------------------------------------
template <bool TEX>
__global__ void __cluster_dims__(kScaleCluster, 1, 1) __launch_bounds__(kScaleThreads, kScaleMinBlocks)
    gn_scale_kernel(const GnLevelMaps M, const GnParams P, const GnState* __restrict__ states,
                    ScaleState* __restrict__ scales)
{
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float smem_samples[];
  __shared__ ScaleShared sh;
  __shared__ Proj s_proj;
  const int b = blockIdx.x / kScaleCluster + P.first;
  const int rank = (int)cluster.block_rank();
  const GnState& st = states[b];
  grid_dep_launch();  // the system kernel that follows may start staging its keyframe tiles

  const int n = P.kept_rows * P.kept_cols;
  const int chunk = (n + kScaleCluster - 1) / kScaleCluster;
  const int begin = min(rank * chunk, n);
  const int n_local = min(chunk, n - begin);
  float* samp_int = smem_samples;
  float* samp_dep = smem_samples + chunk;
  const int s = P.sample_stride;
  // Pose-independent part first: the keyframe values of this CTA's samples go to shared memory while the previous
  // kernel (the last CTAs of the system kernel: final sum, 6x6 solve, pose update) is still running.
  for (int il = threadIdx.x; il < n_local; il += kScaleThreads) {
    const int i = begin + il;
    const int ys = i / P.kept_cols, xs = i - ys * P.kept_cols;
    samp_dep[il] = __ldg(M.W0.row(b, s * ys) + s * xs);
    samp_int[il] = __ldg(M.I0.row(b, s * ys) + s * xs);
  }
  grid_dep_wait();
  if (pair_skipped(st, P)) return;  // uniform over the whole cluster
  if (threadIdx.x < 12) ((float*)&s_proj)[threadIdx.x] = ((const float*)&st.proj[P.level])[threadIdx.x];
  __syncthreads();
  const Proj proj = s_proj;

  const bool geom_is_warped = (P.mode == RGBID_MODE_TRACKER);
  CurFrame cur;
  cur.Wc = M.Wc.row(b, 0); cur.Ic = M.Ic.row(b, 0);
  cur.wpitch = M.Wc.pitch; cur.ipitch = M.Ic.pitch;
  cur.texW = TEX ? M.texW[b] : 0; cur.texI = TEX ? M.texI[b] : 0;
  for (int il = threadIdx.x; il < n_local; il += kScaleThreads) {
    const int i = begin + il;
    const int ys = i / P.kept_cols, xs = i - ys * P.kept_cols;
    const int x = s * xs, y = s * ys;
    const float w0 = samp_dep[il], i0 = samp_int[il];  // staged above by this very thread
    float w1, i1;
    if (P.prewarped) { w1 = __ldg(M.Wc.row(b, y) + x); i1 = __ldg(M.Ic.row(b, y) + x); }  // uniform; see gn_build_kernel
    else warp_pixel<TEX>(proj, x, y, w0, cur, P.cols, P.rows, geom_is_warped, w1, i1);
    samp_int[il] = i1 - i0;
    samp_dep[il] = w1 - w0;
  }
  __syncthreads();

  ScaleSlot s_int, s_dep;
  slot_init(s_int, P.sigma_op, P.mestimator, 0.f, 5.f, true);      // seeds: src/visodo.cpp:1168-1173
  slot_init(s_dep, P.sigma_op, P.mestimator, 0.f, 0.0025f, true);  //        src/keyframe_align.cpp:284-289
  scale_rounds(cluster, sh, s_int, s_dep, samp_int, samp_dep, n_local);

  if (rank == 0 && threadIdx.x == 0) {
    ScaleState out;
    out.bias_int = s_int.out_bias; out.sigma_int = s_int.out_sigma;
    out.bias_depthinv = s_dep.out_bias; out.sigma_depthinv = s_dep.out_sigma;
    out.nu_depthinv = s_dep.nu;
    // tracker: nu_int = max(nu_int, nu_depthinv) (src/visodo.cpp:1186);
    // KeyframeAlign passes nu_depthinv for both residuals (src/keyframe_align.cpp:308)
    out.nu_int = (P.mode == RGBID_MODE_TRACKER) ? fmaxf(s_int.nu, s_dep.nu) : s_dep.nu;
    out.irls_iters_int = s_int.irls_iters; out.irls_iters_depthinv = s_dep.irls_iters;
    scales[b] = out;
  }
}

__global__ void gn_init_kernel(GnState* __restrict__ states, const double* __restrict__ R_init,
                               const double* __restrict__ t_init, int batch, int levels, float fx0, float fy0,
                               float cx0, float cy0, const int* __restrict__ trace_flag)
{
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  GnState& st = states[b];
  for (int i = 0; i < 9; ++i) { st.R[i] = R_init[9 * b + i]; st.R0[i] = st.R[i]; }
  for (int i = 0; i < 3; ++i) { st.t[i] = t_init[3 * b + i]; st.t0[i] = st.t[i]; }
  for (int i = 0; i < 36; ++i) { st.cov[i] = 0.0; st.lastA[i] = 0.0; }
  st.status = RGBID_OK; st.iter_count = 0;
  st.skip_level = -1; st.rmse_prev = 9999.f;
  st.trace_on = (trace_flag == nullptr) ? 1 : (*trace_flag != 0);
  for (int l = 0; l < RGBID_MAX_LEVELS; ++l) st.iters_done[l] = 0;
  st.chi_square = 0.f; st.chi_test = 0.f; st.ndof = 0.f;
  refresh_proj(st, st.R, st.t, levels, fx0, fy0, cx0, cy0);
}

__global__ void export_systems_kernel(const GnState* __restrict__ states, double* __restrict__ out, int batch)
{
  const int b = blockIdx.x, k = threadIdx.x;  // 48 threads per pair
  if (b >= batch || k >= 48) return;
  const GnState& st = states[b];
  double v = (k < 36) ? st.cov[k] : (k < 45) ? st.R[k - 36] : st.t[k - 45];
  if (st.status != RGBID_OK) v = __longlong_as_double(0x7ff8000000000000ll);
  out[(size_t)b * 48 + k] = v;
}

// every stream of the map owns whole 128-pixel chunks (the aligner NaN-fills the padding and nothing writes it),
// so the fast kernel may read the last chunk in full
inline bool padded(const ImgB& m, const GnParams& P)
{
  const size_t need = (((size_t)P.rows * P.cols + 127) / 128) * 512;
  return m.sstride >= need;
}

inline bool aligned16(const ImgB& m) { return ((uintptr_t)m.p % 16 == 0) && (m.pitch % 16 == 0) && (m.sstride % 16 == 0); }

}  // namespace

int gn_build_grid_x(int rows, int cols, int batch, int num_sms)
{
  // One balanced wave: the kernel is resident at kBuildMinBlocks CTAs / SM, so the whole launch (all pairs)
  // should use at most that many CTAs per SM, every thread should get the same number k of units, and a pair never gets more than
  // num_sms CTAs so that the last-block final sum stays short.
  const int units = (cols % 4 == 0) ? (cols / 4) * rows : cols * rows;
  int cap = (kBuildMinBlocks * num_sms) / batch;
  if (cap < 1) cap = 1;
  if (cap > num_sms) cap = num_sms;
  const int k = (units + cap * kBuildThreads - 1) / (cap * kBuildThreads);  // units per thread
  int g = (units + k * kBuildThreads - 1) / (k * kBuildThreads);
  return g < 1 ? 1 : g;
}

int gn_prepare_device()
{
  static PerDevice prepared;
  cudaError_t err = cudaSuccess;
  prepared.once([&err] {
    if (!upload_nu_table()) { err = cudaGetLastError(); if (err == cudaSuccess) err = cudaErrorUnknown; return false; }
    cudaError_t e = cudaSuccess;
#define RGBID_FAST_ATTR(T, C) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_build_fast_kernel<T, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_smem_bytes(false))
    RGBID_FAST_ATTR(true, 0); RGBID_FAST_ATTR(true, 1); RGBID_FAST_ATTR(true, 2);
    RGBID_FAST_ATTR(false, 0); RGBID_FAST_ATTR(false, 1); RGBID_FAST_ATTR(false, 2);
#undef RGBID_FAST_ATTR
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gn_build_fast_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast_smem_bytes(true));
    if (kScaleCluster > 8) {  // more than the portable cluster size
      if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_scale_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_scale_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    err = e;
    return e == cudaSuccess;
  });
  return (int)err;
}

void launch_export_systems(const LaunchCtx& L, const GnState* states, double* out, int batch)
{
  export_systems_kernel<<<batch, 64, 0, L.stream>>>(states, out, batch);
  ++*L.launches;
}

void launch_gn_init(const LaunchCtx& L, GnState* states, const double* R_init, const double* t_init, int batch,
                    int levels, float fx0, float fy0, float cx0, float cy0, const int* trace_flag)
{
  gn_init_kernel<<<(batch + 63) / 64, 64, 0, L.stream>>>(states, R_init, t_init, batch, levels, fx0, fy0, cx0, cy0, trace_flag);
  ++*L.launches;
}

void launch_gn_scale(const LaunchCtx& L, const GnLevelMaps& M, const GnParams& P, GnState* states, ScaleState* scales)
{
  const int n = P.kept_rows * P.kept_cols;
  const int chunk = (n + kScaleCluster - 1) / kScaleCluster;
  size_t smem = (size_t)2 * chunk * sizeof(float);
  static PerDevice smem_limit;
  gn_prepare_device();  // no-op after rgbid_aligner_create
  if (smem > 48 * 1024)
    smem_limit.at_least(smem, [](size_t bytes) {
      cudaFuncSetAttribute(gn_scale_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      cudaFuncSetAttribute(gn_scale_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    });
  if (M.texW != nullptr && M.texI != nullptr)
    launch_kernel_pdl(gn_scale_kernel<true>, dim3(kScaleCluster * P.batch), dim3(kScaleThreads), smem, L.stream, L.pdl, M, P, states, scales);
  else
    launch_kernel_pdl(gn_scale_kernel<false>, dim3(kScaleCluster * P.batch), dim3(kScaleThreads), smem, L.stream, L.pdl, M, P, states, scales);
  ++*L.launches;
}

template <int VEC, bool CHI>
static void launch_gn_build_t(const LaunchCtx& L, dim3 grid, bool tex, const GnLevelMaps& M, const GnParams& P,
                              GnState* states, const ScaleState* scales, double* partials, int partial_stride,
                              unsigned int* counters, rgbid_iter_trace* trace)
{
  if (tex)
    launch_kernel_pdl(gn_build_kernel<VEC, CHI, true>, grid, dim3(kBuildThreads), 0, L.stream, L.pdl, M, P, states, scales, partials, partial_stride, counters, trace);
  else
    launch_kernel_pdl(gn_build_kernel<VEC, CHI, false>, grid, dim3(kBuildThreads), 0, L.stream, L.pdl, M, P, states, scales, partials, partial_stride, counters, trace);
}

void launch_gn_build(const LaunchCtx& L, const GnLevelMaps& M, const GnParams& P, GnState* states,
                     const ScaleState* scales, double* partials, int partial_stride, unsigned int* counters,
                     rgbid_iter_trace* trace)
{
  bool vec = (P.cols % 4 == 0) && aligned16(M.W0) && aligned16(M.I0) && aligned16(M.gWx) && aligned16(M.gWy) &&
             aligned16(M.gIx) && aligned16(M.gIy);
  const bool chi = (P.chi_mestimator >= 0);
  const bool tex = (M.texW != nullptr && M.texI != nullptr);
  // Fast path: flat keyframe maps (pitch == cols * 4, so a chunk of 128 pixels is one contiguous segment per map),
  // texture gathers (inverse-depth texture with border addressing), Student-t weights (estimated nu, or the fixed
  // nu = 5 of computeWeight(STUDENT): the same expression), INDEPENDENT weighting.
  static const bool no_fast = [] { const char* e = getenv("RGBID_NO_FAST"); return e && e[0] == '1'; }();
  const size_t flat = (size_t)P.cols * sizeof(float);
  const bool prew = (P.prewarped != 0);
  const bool fast = !no_fast && vec && (prew ? (P.mode == RGBID_MODE_TRACKER && !chi) : (tex && M.tex_border)) &&
                    P.weighting == RGBID_INDEPENDENT &&
                    (P.student_nu || P.mestimator == RGBID_STUDENT) && M.W0.pitch == flat && M.I0.pitch == flat &&
                    M.gWx.pitch == flat && M.gWy.pitch == flat && M.gIx.pitch == flat && M.gIy.pitch == flat &&
                    (long long)P.rows * P.cols <= 2500000ll && padded(M.W0, P) && padded(M.I0, P) && padded(M.gWx, P) &&
                    padded(M.gWy, P) && padded(M.gIx, P) && padded(M.gIy, P) &&
                    (!prew || (M.Wc.pitch == flat && M.Ic.pitch == flat && padded(M.Wc, P) && padded(M.Ic, P) &&
                               aligned16(M.Wc) && aligned16(M.Ic)));
  if (fast) {
    FastGeom G;
    G.npx = P.rows * P.cols;
    G.nchunks = (G.npx + kChunkPx - 1) / kChunkPx;
    G.colsf = (float)P.cols; G.rowsf = (float)P.rows; G.inv_cols = 1.f / (float)P.cols;
    // one balanced wave of kBuildMinBlocks CTAs / SM over all pairs; a pair never gets more CTAs than it has 8-chunk groups
    int cap = (kBuildMinBlocks * L.num_sms) / (P.batch_total > 0 ? P.batch_total : P.batch);
    if (cap < 1) cap = 1;
    if (cap > L.num_sms) cap = L.num_sms;
    int gx = (G.nchunks + kBuildWarps - 1) / kBuildWarps;
    if (gx > cap) gx = cap;
    dim3 grid(gx, P.batch);
    gn_prepare_device();  // no-op after rgbid_aligner_create
    const bool tracker = (P.mode == RGBID_MODE_TRACKER);
    const int chim = !chi ? 0 : (P.chi_mestimator == RGBID_STUDENT ? 2 : 1);
#define RGBID_FAST_LAUNCH(T, C) \
    launch_kernel_pdl(gn_build_fast_kernel<T, C>, grid, dim3(kBuildThreads), fast_smem_bytes(false), L.stream, L.pdl, M, P, G, states, scales, partials, partial_stride, counters, trace)
    if (prew)
      launch_kernel_pdl(gn_build_fast_kernel<true, 0, true>, grid, dim3(kBuildThreads), fast_smem_bytes(true), L.stream, L.pdl, M, P, G,
                        states, scales, partials, partial_stride, counters, trace);
    else if (tracker) { if (chim == 0) RGBID_FAST_LAUNCH(true, 0); else if (chim == 1) RGBID_FAST_LAUNCH(true, 1); else RGBID_FAST_LAUNCH(true, 2); }
    else { if (chim == 0) RGBID_FAST_LAUNCH(false, 0); else if (chim == 1) RGBID_FAST_LAUNCH(false, 1); else RGBID_FAST_LAUNCH(false, 2); }
#undef RGBID_FAST_LAUNCH
    ++*L.launches;
    return;
  }
  dim3 grid(gn_build_grid_x(P.rows, P.cols, P.batch_total > 0 ? P.batch_total : P.batch, L.num_sms), P.batch);
  if (vec) {
    if (chi) launch_gn_build_t<4, true>(L, grid, tex, M, P, states, scales, partials, partial_stride, counters, trace);
    else launch_gn_build_t<4, false>(L, grid, tex, M, P, states, scales, partials, partial_stride, counters, trace);
  } else {
    if (chi) launch_gn_build_t<1, true>(L, grid, tex, M, P, states, scales, partials, partial_stride, counters, trace);
    else launch_gn_build_t<1, false>(L, grid, tex, M, P, states, scales, partials, partial_stride, counters, trace);
  }
  ++*L.launches;
}

void launch_build_system(const LaunchCtx& L, ImgB W0, ImgB I0, ImgB gWx, ImgB gWy, ImgB gIx, ImgB gIy, ImgB W1,
                         ImgB I1, const rgbid_system_params& sp, double* partials, unsigned int* counter,
                         double* out27)
{
  PixelParams pp;
  pp.fx = sp.fx; pp.fy = sp.fy; pp.cx = sp.cx; pp.cy = sp.cy;
  pp.inv_sigma_int = 1.f / sp.sigma_int; pp.inv_sigma_depthinv = 1.f / sp.sigma_depthinv;
  pp.inv_sigma2_int = pp.inv_sigma_int * pp.inv_sigma_int;
  pp.inv_sigma2_depthinv = pp.inv_sigma_depthinv * pp.inv_sigma_depthinv;
  pp.bias_over_sigma_int = sp.bias_int / sp.sigma_int;
  pp.bias_over_sigma_depthinv = sp.bias_depthinv / sp.sigma_depthinv;
  pp.nu_int = sp.nu_int; pp.nu_depthinv = sp.nu_depthinv;
  pp.mestimator = sp.mestimator; pp.weighting = sp.weighting; pp.student_nu = sp.student_nu;
  pp.chi_mestimator = -1;
  bool vec = (W0.cols % 4 == 0) && aligned16(W0) && aligned16(I0) && aligned16(gWx) && aligned16(gWy) &&
             aligned16(gIx) && aligned16(gIy) && aligned16(W1) && aligned16(I1);
  int grid = gn_build_grid_x(W0.rows, W0.cols, 1, L.num_sms);
  if (vec) build_system_kernel<4><<<grid, kBuildThreads, 0, L.stream>>>(W0, I0, gWx, gWy, gIx, gIy, W1, I1, pp, partials, counter, out27);
  else build_system_kernel<1><<<grid, kBuildThreads, 0, L.stream>>>(W0, I0, gWx, gWy, gIx, gIy, W1, I1, pp, partials, counter, out27);
  ++*L.launches;
}

}  // namespace rgbid
