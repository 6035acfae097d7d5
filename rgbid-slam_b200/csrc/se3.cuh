// se3.cuh -- double-precision SE(3) / 6x6 algebra shared by device kernels and the C++ host classes.
//
// Restates the host-side Gauss-Newton step of the reference (src/visodo.cpp:1242-1263,
// src/keyframe_align.cpp:312-335) and its Lie maps (src/util_funcs.cpp:31-155); the reference does
// this on the host with Eigen after a device->host copy per iteration, here it runs in the tail of
// the reduction kernel so the loop never leaves the GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RGBID_HD __host__ __device__ __forceinline__
// every fixed-count loop is unrolled on the device: the Gauss-Newton tail runs in ONE thread per frame pair while the
// GPU waits, so it must be straight-line register code with instruction-level parallelism, not loops over local
// memory
#define RGBID_UNROLL _Pragma("unroll")
#else
#define RGBID_HD inline
#define RGBID_UNROLL
#endif

namespace rgbid {

// Reciprocal and reciprocal square root for the serial Gauss-Newton tail.  On the device a double division / rsqrt is
// a ~40-instruction dependent chain with range fix-ups; the tail runs in one thread per frame pair while the GPU
// waits (RGBID_TAIL_PROBE: ~11 000 clocks per launch before this), so both start from the float SFU seed and take two
// Newton steps in double (relative error <= 2 ulp; operands outside the float range take the IEEE path).
RGBID_HD double rcp_d(double x)
{
#if defined(__CUDA_ARCH__)
  const double ax = fabs(x);
  if (ax > 1e-30 && ax < 1e30) {
    double y = (double)__frcp_rn((float)x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
  }
#endif
  return 1.0 / x;
}

RGBID_HD double rsqrt_d(double x)
{
#if defined(__CUDA_ARCH__)
  if (x > 1e-30 && x < 1e30) {
    double y = (double)rsqrtf((float)x);
    double h = 0.5 * x * y;
    double e = fma(-h, y, 0.5);
    y = fma(y, e, y);
    h = 0.5 * x * y;
    e = fma(-h, y, 0.5);
    return fma(y, e, y);
  }
  return rsqrt(x);  // zero / negative / NaN pivots keep their IEEE results (inf / NaN -> NaN pose -> "lost")
#else
  return 1.0 / sqrt(x);
#endif
}

// sin(theta) / theta and (1 - cos(theta)) / theta^2 from t = theta^2.  Device: even Taylor series to t^7 for
// theta < 0.5 rad (truncation < 5e-17; a Gauss-Newton increment is ~1e-3 rad), libm otherwise and on the host.
RGBID_HD void sinc_cosc(double t, double* a, double* b)
{
#if defined(__CUDA_ARCH__)
  if (t < 0.25) {
    *a = 1.0 + t * (-1.0 / 6.0 + t * (1.0 / 120.0 + t * (-1.0 / 5040.0 + t * (1.0 / 362880.0 + t * (-1.0 / 39916800.0 +
         t * (1.0 / 6227020800.0 - t * (1.0 / 1307674368000.0)))))));
    *b = 0.5 + t * (-1.0 / 24.0 + t * (1.0 / 720.0 + t * (-1.0 / 40320.0 + t * (1.0 / 3628800.0 + t * (-1.0 / 479001600.0 +
         t * (1.0 / 87178291200.0 - t * (1.0 / 20922789888000.0)))))));
    return;
  }
#endif
  const double theta = sqrt(t);
  *a = sin(theta) / theta;
  *b = (1.0 - cos(theta)) / (theta * theta);
}

RGBID_HD void mat3_mul(const double* A, const double* B, double* C)
{
  double T[9];
  RGBID_UNROLL
  for (int i = 0; i < 3; ++i)
    RGBID_UNROLL
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}

RGBID_HD void mat3_vec(const double* A, const double* v, double* r)
{
  double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}

RGBID_HD void mat3_transpose(const double* A, double* At)
{
  double T[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) At[i] = T[i];
}

// General 3x3 inverse by cofactors (the reference calls Eigen's .inverse() on rotations,
// src/visodo.cpp:1066, rather than transposing).
RGBID_HD void mat3_inverse(const double* M, double* Mi)
{
  double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  double id = rcp_d(M[0] * c00 + M[1] * c01 + M[2] * c02);
  double T[9];
  T[0] = c00 * id; T[1] = (M[2] * M[7] - M[1] * M[8]) * id; T[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  T[3] = c01 * id; T[4] = (M[0] * M[8] - M[2] * M[6]) * id; T[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  T[6] = c02 * id; T[7] = (M[1] * M[6] - M[0] * M[7]) * id; T[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) Mi[i] = T[i];
}

// Orthogonal polar factor (= U V^T of the SVD; forceOrthogonalisation, src/util_funcs.cpp:150-155)
// by the quadratically convergent Newton iteration X <- (X + X^-T) / 2.
RGBID_HD void force_orthogonal(const double* M, double* R)
{
  double X[9];
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) X[i] = M[i];
  for (int it = 0; it < 12; ++it) {
    double Xi[9], d = 0.0;
    mat3_inverse(X, Xi);
    RGBID_UNROLL
    for (int i = 0; i < 3; ++i)
      RGBID_UNROLL
      for (int j = 0; j < 3; ++j) {
        double y = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
        d += fabs(y - X[3 * i + j]);
        R[3 * i + j] = y;
      }
    RGBID_UNROLL
    for (int i = 0; i < 9; ++i) X[i] = R[i];
    // quadratic convergence: a step below 1e-14 leaves an error far below one ulp (a tighter test never fires on
    // rounding noise and burns all 12 iterations in the single-thread tail of every Gauss-Newton launch)
    if (d < 1e-14) break;
  }
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) R[i] = X[i];
}

RGBID_HD void skew3(const double* w, double* S)
{
  S[0] = 0; S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2]; S[4] = 0; S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}

// expMapRot, src/util_funcs.cpp:125-148 (small-angle switch at 1e-5)
RGBID_HD void exp_map_rot(const double* omega, double* R)
{
  const double theta2 = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double O[9], O2[9], M[9], a, b;
  skew3(omega, O);
  mat3_mul(O, O, O2);
  if (theta2 < 0.00001 * 0.00001) { a = 1.0; b = 0.5; }
  else sinc_cosc(theta2, &a, &b);
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) M[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  force_orthogonal(M, R);
}

// expMap, src/util_funcs.cpp:85-123: T = [R | Q v]
RGBID_HD void exp_map(const double* omega, const double* v, double* R, double* t)
{
  double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double O[9], O2[9], M[9], Q[9], a, b, c;
  skew3(omega, O);
  mat3_mul(O, O, O2);
  if (theta < 0.00001) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
  else {
    a = sin(theta) / theta;
    b = (1.0 - cos(theta)) / (theta * theta);
    c = (1.0 - a) / (theta * theta);
  }
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) {
    double I = (i % 4 == 0) ? 1.0 : 0.0;
    M[i] = I + a * O[i] + b * O2[i];
    Q[i] = I + b * O[i] + c * O2[i];
  }
  force_orthogonal(M, R);
  mat3_vec(Q, v, t);
}

// logMap, src/util_funcs.cpp:31-82: twist = [v; omega]
RGBID_HD void log_map(const double* Rin, const double* trans, double* twist)
{
  double R[9];
  force_orthogonal(Rin, R);
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c), theta2 = theta * theta;
  double th_by_sinth = (s < 1e-5) ? 1.0 + (1.0 / 6.0) * theta2 + (7.0 / 360.0) * theta2 * theta2 : theta / s;
  double vth = th_by_sinth / 2.0;
  double om[3] = {rx * vth, ry * vth, rz * vth};
  double O[9], O2[9], Q[9], Qi[9], b, cc;
  skew3(om, O);
  mat3_mul(O, O, O2);
  double th = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  if (th < 0.00001) { b = 0.5; cc = 1.0 / 6.0; }
  else { b = (1.0 - cos(theta)) / (theta * theta); cc = (1.0 - (sin(theta) / theta)) / (theta * theta); }
  RGBID_UNROLL
  for (int i = 0; i < 9; ++i) Q[i] = ((i % 4 == 0) ? 1.0 : 0.0) + b * O[i] + cc * O2[i];
  mat3_inverse(Q, Qi);
  mat3_vec(Qi, trans, twist);
  twist[3] = om[0]; twist[4] = om[1]; twist[5] = om[2];
}

// Unpack the 27 upper-triangular sums [A00..A05,b0, A11..A15,b1, ...] (src/cuda/estimate_VO.cu:771-786)
RGBID_HD void unpack_system(const double* s27, double* A36, double* b6)
{
  int shift = 0;
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i)
    RGBID_UNROLL
    for (int j = i; j < 7; ++j) {
      double v = s27[shift++];
      if (j == 6) b6[i] = v;
      else { A36[j * 6 + i] = v; A36[i * 6 + j] = v; }
    }
}

// Cholesky solve of the 6x6 normal equations (A.llt().solve(b), src/visodo.cpp:1249).
// A non-SPD matrix yields NaN in x, which reaches the NaN-pose guard exactly as in the reference.
RGBID_HD void llt_solve6(const double* A, const double* b, double* x)
{
  // One reciprocal square root per column instead of a square root and (5 - j) + 2 divisions: this runs in a
  // single thread at the end of every Gauss-Newton launch while the rest of the GPU waits, and a double-precision
  // division is a ~40-instruction dependent chain there.
  double L[36], inv[6];
  RGBID_UNROLL
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  RGBID_UNROLL
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    RGBID_UNROLL
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
    const double r = rsqrt_d(d);
    inv[j] = r;
    L[j * 6 + j] = d * r;
    RGBID_UNROLL
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      RGBID_UNROLL
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s * r;
    }
  }
  double y[6];
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    RGBID_UNROLL
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
    y[i] = s * inv[i];
  }
  RGBID_UNROLL
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    RGBID_UNROLL
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s * inv[i];
  }
}

// The same solve straight from the 27 packed sums [A00..A05,b0, A11..A15,b1, ...]: right-looking Cholesky on the
// augmented upper triangle [A | b] in place (27 + 6 doubles live instead of 36 + 36 + 12: the Gauss-Newton tail is one
// thread at the kernel's register limit, and the unpacked form spilled to local memory), trailing updates independent
// of each other (instruction-level parallelism for the single thread).  x = A^-1 b as A.llt().solve(b).
RGBID_HD void llt_solve_packed(const double* s27, double* x)
{
  double u[6][7], inv[6];
  {
    int shift = 0;
    RGBID_UNROLL
    for (int i = 0; i < 6; ++i)
      RGBID_UNROLL
      for (int j = i; j < 7; ++j) u[i][j] = s27[shift++];
  }
  RGBID_UNROLL
  for (int k = 0; k < 6; ++k) {
    const double r = rsqrt_d(u[k][k]);
    inv[k] = r;
    RGBID_UNROLL
    for (int j = k; j < 7; ++j) u[k][j] *= r;  // row k of U = L^T, and y[k] in column 6
    RGBID_UNROLL
    for (int i = k + 1; i < 6; ++i)
      RGBID_UNROLL
      for (int j = i; j < 7; ++j) u[i][j] -= u[k][i] * u[k][j];
  }
  RGBID_UNROLL
  for (int i = 5; i >= 0; --i) {
    double s = u[i][6];
    RGBID_UNROLL
    for (int k = i + 1; k < 6; ++k) s -= u[i][k] * x[k];
    x[i] = s * inv[i];
  }
}

// 6x6 inverse by Gauss-Jordan with partial pivoting (covariance = A^-1, src/visodo.cpp:1409).
RGBID_HD bool inverse6(const double* A, double* Ai)
{
  double M[6][12];
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i)
    RGBID_UNROLL
    for (int j = 0; j < 6; ++j) { M[i][j] = A[i * 6 + j]; M[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  RGBID_UNROLL
  for (int c = 0; c < 6; ++c) {
    int p = c;
    RGBID_UNROLL
    for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0) return false;
    if (p != c) for (int j = 0; j < 12; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    double ip = rcp_d(M[c][c]);
    RGBID_UNROLL
    for (int j = 0; j < 12; ++j) M[c][j] *= ip;
    RGBID_UNROLL
    for (int r = 0; r < 6; ++r) if (r != c) {
      double f = M[r][c];
      if (f != 0.0) for (int j = 0; j < 12; ++j) M[r][j] -= f * M[c][j];
    }
  }
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ai[i * 6 + j] = M[i][6 + j];
  return true;
}

// Inverse of a symmetric positive definite 6x6 matrix given as the 27 packed sums (upper triangle row by row, the
// residual column ignored): A = U^T U (rsqrt pivots, as llt_solve_packed), V = U^-1, A^-1 = V V^T.  Straight-line,
// register resident, no divisions and no pivot search -- the covariance pass's A is the Gauss-Newton normal matrix.
// Returns false (Ai untouched) when a pivot is not positive; the caller then takes the general inverse6.
RGBID_HD bool inverse6_spd_packed(const double* s27, double* Ai)
{
  double u[6][6], inv[6];
  {
    int shift = 0;
    RGBID_UNROLL
    for (int i = 0; i < 6; ++i) {
      RGBID_UNROLL
      for (int j = i; j < 6; ++j) u[i][j] = s27[shift++];
      ++shift;  // J^T r
    }
  }
  bool ok = true;
  RGBID_UNROLL
  for (int k = 0; k < 6; ++k) {
    ok = ok && (u[k][k] > 0.0);
    const double r = rsqrt_d(u[k][k]);
    inv[k] = r;
    RGBID_UNROLL
    for (int j = k; j < 6; ++j) u[k][j] *= r;
    RGBID_UNROLL
    for (int i = k + 1; i < 6; ++i)
      RGBID_UNROLL
      for (int j = i; j < 6; ++j) u[i][j] -= u[k][i] * u[k][j];
  }
  if (!ok) return false;
  double v[6][6];  // upper triangular inverse of U
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i) {
    v[i][i] = inv[i];
    RGBID_UNROLL
    for (int j = i + 1; j < 6; ++j) {
      double s = 0.0;
      RGBID_UNROLL
      for (int k = i; k < j; ++k) s += v[i][k] * u[k][j];
      v[i][j] = -s * inv[j];
    }
  }
  RGBID_UNROLL
  for (int i = 0; i < 6; ++i)
    RGBID_UNROLL
    for (int j = i; j < 6; ++j) {
      double s = 0.0;
      RGBID_UNROLL
      for (int k = j; k < 6; ++k) s += v[i][k] * v[j][k];
      Ai[i * 6 + j] = s;
      Ai[j * 6 + i] = s;
    }
  return true;
}

// One Gauss-Newton update T <- T_inc T with x = [trans; rot], R_inc^-1 = expMapRot(rot),
// t_inc = -R_inc trans (src/visodo.cpp:1252-1263).  Returns true if the pose became NaN.
RGBID_HD bool gn_update(const double* x, double* R, double* t)
{
  double Rinc_inv[9], Rinc[9], tinc[3], tn[3];
  exp_map_rot(x + 3, Rinc_inv);
  mat3_inverse(Rinc_inv, Rinc);
  mat3_vec(Rinc, x, tinc);
  mat3_vec(Rinc, t, tn);
  RGBID_UNROLL
  for (int k = 0; k < 3; ++k) t[k] = tn[k] - tinc[k];
  mat3_mul(Rinc, R, R);
  double nr = 0, nt = 0;
  RGBID_UNROLL
  for (int k = 0; k < 9; ++k) nr += R[k] * R[k];
  RGBID_UNROLL
  for (int k = 0; k < 3; ++k) nt += t[k] * t[k];
  return (nr != nr) || (nt != nt);
}

// The same update for the device tail, as few dependent instructions as the algebra allows (the tail is one thread per
// frame pair at the end of every launch): R_inc = exp(-[w]x) = I - a [w]x + b [w]x^2 directly instead of
// inverse(orthogonalise(exp([w]x))), one Newton-Schulz step X (3 I - X^T X) / 2 for the polar factor (Rodrigues'
// formula is orthogonal to rounding, so one quadratically convergent step is exact to double precision; the
// reference runs an SVD), no division.  Differs from gn_update by rounding only (~1e-16).
RGBID_HD bool gn_update_lean(const double* x, double* R, double* t)
{
  const double wx = x[3], wy = x[4], wz = x[5];
  const double xx = wx * wx, yy = wy * wy, zz = wz * wz;
  const double theta2 = xx + yy + zz;
  double a, b;
  if (theta2 < 0.00001 * 0.00001) { a = 1.0; b = 0.5; }
  else sinc_cosc(theta2, &a, &b);
  // O = [w]x, O^2 = w w^T - theta^2 I ; X = I - a O + b O^2
  const double bxy = b * wx * wy, bxz = b * wx * wz, byz = b * wy * wz;
  double X[9] = {1.0 - b * (yy + zz), bxy + a * wz, bxz - a * wy,
                 bxy - a * wz, 1.0 - b * (xx + zz), byz + a * wx,
                 bxz + a * wy, byz - a * wx, 1.0 - b * (xx + yy)};
  // E = X^T X (symmetric), Rinc = X (3 I - E) / 2
  const double e00 = X[0] * X[0] + X[3] * X[3] + X[6] * X[6], e01 = X[0] * X[1] + X[3] * X[4] + X[6] * X[7];
  const double e02 = X[0] * X[2] + X[3] * X[5] + X[6] * X[8], e11 = X[1] * X[1] + X[4] * X[4] + X[7] * X[7];
  const double e12 = X[1] * X[2] + X[4] * X[5] + X[7] * X[8], e22 = X[2] * X[2] + X[5] * X[5] + X[8] * X[8];
  const double F[9] = {1.5 - 0.5 * e00, -0.5 * e01, -0.5 * e02, -0.5 * e01, 1.5 - 0.5 * e11, -0.5 * e12,
                       -0.5 * e02, -0.5 * e12, 1.5 - 0.5 * e22};
  double Rinc[9];
  mat3_mul(X, F, Rinc);
  const double d[3] = {t[0] - x[0], t[1] - x[1], t[2] - x[2]};  // t <- Rinc t - Rinc x_trans
  mat3_vec(Rinc, d, t);
  mat3_mul(Rinc, R, R);
  const double chk = ((R[0] + R[1]) + (R[2] + R[3])) + ((R[4] + R[5]) + (R[6] + R[7])) + ((R[8] + t[0]) + (t[1] + t[2]));
  return (chk != chk) || (fabs(chk) > 1e300);
}

// K R K^-1 and K t in float (Eigen float products of src/visodo.cpp:1110-1114, :1496-1500).
RGBID_HD void projective_pose(const double* R, const double* t, float fx, float fy, float cx, float cy,
                              float* Rp, float* tp)
{
  float Rf[9], T[9];
  RGBID_UNROLL
  for (int k = 0; k < 9; ++k) Rf[k] = (float)R[k];
  float tf0 = (float)t[0], tf1 = (float)t[1], tf2 = (float)t[2];
  const float K[9] = {fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f};
  const float Ki[9] = {1.f / fx, 0.f, -cx / fx, 0.f, 1.f / fy, -cy / fy, 0.f, 0.f, 1.f};
  RGBID_UNROLL
  for (int i = 0; i < 3; ++i)
    RGBID_UNROLL
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = K[3 * i] * Rf[j] + K[3 * i + 1] * Rf[3 + j] + K[3 * i + 2] * Rf[6 + j];
  RGBID_UNROLL
  for (int i = 0; i < 3; ++i)
    RGBID_UNROLL
    for (int j = 0; j < 3; ++j)
      Rp[3 * i + j] = T[3 * i] * Ki[j] + T[3 * i + 1] * Ki[3 + j] + T[3 * i + 2] * Ki[6 + j];
  RGBID_UNROLL
  for (int i = 0; i < 3; ++i) tp[i] = K[3 * i] * tf0 + K[3 * i + 1] * tf1 + K[3 * i + 2] * tf2;
}

// K R^-1 K^-1 and -K R^-1 t (src/visodo.cpp:1066-1067, 1108-1114)
RGBID_HD void projective_inverse_pose(const double* R, const double* t, float fx, float fy, float cx, float cy,
                                      float* Rp, float* tp)
{
  double Ri[9], ti[3];
  mat3_inverse(R, Ri);
  mat3_vec(Ri, t, ti);
  ti[0] = -ti[0]; ti[1] = -ti[1]; ti[2] = -ti[2];
  projective_pose(Ri, ti, fx, fy, cx, cy, Rp, tp);
}

}  // namespace rgbid
