// Host build of rgbid-slam_b200/csrc/se3.cuh (the solver-tail algebra is __host__ __device__): the Cholesky-based inverse of
// the packed normal matrix against the general Gauss-Jordan inverse, and the packed Cholesky solve against it.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../rgbid-slam_b200/csrc/se3.cuh"

int main()
{
  using namespace rgbid;
  double worst_inv = 0.0, worst_solve = 0.0;
  srand(12345);
  for (int trial = 0; trial < 500; ++trial) {
    double J[40][6], e[40];
    const double scale = (double)((trial % 7) + 1);
    for (int k = 0; k < 40; ++k) {
      for (int i = 0; i < 6; ++i) J[k][i] = (rand() / (double)RAND_MAX - 0.5) * scale * ((i < 3) ? 100.0 : 1.0);
      e[k] = rand() / (double)RAND_MAX - 0.5;
    }
    double A[36] = {0}, b[6] = {0};
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 40; ++k) s += J[k][i] * J[k][j]; A[i * 6 + j] = s; }
      double s = 0; for (int k = 0; k < 40; ++k) s += J[k][i] * e[k]; b[i] = s;
    }
    double s27[27]; int sh = 0;
    for (int i = 0; i < 6; ++i) { for (int j = i; j < 6; ++j) s27[sh++] = A[i * 6 + j]; s27[sh++] = b[i]; }
    double X[36], Y[36], x[6];
    if (!inverse6_spd_packed(s27, X)) { printf("FAIL: positive definite matrix rejected\n"); return 1; }
    if (!inverse6(A, Y)) { printf("FAIL: inverse6\n"); return 1; }
    double m = 0, d = 0;
    for (int i = 0; i < 36; ++i) { m = fmax(m, fabs(Y[i])); d = fmax(d, fabs(X[i] - Y[i])); }
    worst_inv = fmax(worst_inv, d / m);
    llt_solve_packed(s27, x);
    double mx = 0, dx = 0;
    for (int i = 0; i < 6; ++i) {
      double xi = 0; for (int j = 0; j < 6; ++j) xi += Y[i * 6 + j] * b[j];
      mx = fmax(mx, fabs(xi)); dx = fmax(dx, fabs(xi - x[i]));
    }
    worst_solve = fmax(worst_solve, dx / mx);
  }
  double zero[27] = {0}, X[36];
  const bool singular_rejected = !inverse6_spd_packed(zero, X);
  printf("worst_inverse_rel %.3e worst_solve_rel %.3e singular_rejected %d\n", worst_inv, worst_solve, (int)singular_rejected);
  return (worst_inv < 1e-9 && worst_solve < 1e-9 && singular_rejected) ? 0 : 1;
}
