/*
 * oracle.h -- CPU restatement of RGBiD-SLAM's dense alignment path (TEST INFRASTRUCTURE ONLY).
 * See oracle.c for the reference citations.  Nothing in the product may include this header.
 */
#ifndef RGBID_ORACLE_H_
#define RGBID_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/internal.h:66-72 */
enum { ORC_LSQ = 0, ORC_HUBER = 1, ORC_TUKEY = 2, ORC_STUDENT = 3 };
enum { ORC_SIGMA_MAD = 0, ORC_SIGMA_PDF = 1, ORC_SIGMA_CONS = 2 };
enum { ORC_INDEPENDENT = 0, ORC_MIN_WEIGHT = 1, ORC_GEOM_ONLY = 2, ORC_PHOT_ONLY = 3 };

enum { ORC_MODE_TRACKER = 0, ORC_MODE_ALIGN = 1 };
enum { ORC_TEX_FRAC_ROUND = 0, ORC_TEX_FRAC_TRUNC = 1, ORC_TEX_FRAC_EXACT = 2 };

#define ORC_MAX_LEVELS 8

typedef struct {
  float fx, fy, cx, cy;
  int mestimator, weighting, student_nu;
  float sigma_depthinv, sigma_int, bias_depthinv, bias_int, nu_depthinv, nu_int;
} orc_system_params;

typedef struct {
  int rows, cols, levels, finest_level;
  int iterations[ORC_MAX_LEVELS];
  int mode;            /* ORC_MODE_TRACKER | ORC_MODE_ALIGN */
  int mestimator;      /* tracker: Mestimator_ (steers the sigma estimator only) */
  int weighting;
  int sigma_estimator; /* tracker only */
  int nsamples;        /* 10000 tracker, 19200 align */
  float fx, fy, cx, cy; /* level-0 intrinsics */
  int warp_first;      /* tracker only: WARP_ORDER = warpFirst (src/visodo.cpp:1078-1105); 0 = pyrFirst */
  int termination;     /* ORC_TERM_*: tracker's TERMINATION_CRITERIA (src/internal.h:112, src/visodo.cpp:1134-1164) */
  float conv_eps;      /* ORC_TERM_CONVERGENCE (BASELINE config 2, not in the reference): |x| < conv_eps ends a level */
} orc_align_config;

enum { ORC_TERM_ALL_ITERS = 0, ORC_TERM_CHI_SQUARED = 1, ORC_TERM_CONVERGENCE = 2 };

typedef struct {
  const float* W_kf[ORC_MAX_LEVELS];
  const float* I_kf[ORC_MAX_LEVELS];
  const float* gWx_kf[ORC_MAX_LEVELS];
  const float* gWy_kf[ORC_MAX_LEVELS];
  const float* gIx_kf[ORC_MAX_LEVELS];
  const float* gIy_kf[ORC_MAX_LEVELS];
  const float* gWx_cov[ORC_MAX_LEVELS]; /* bilateral-filtered ("covOnly") gradients, tracker */
  const float* gWy_cov[ORC_MAX_LEVELS];
  const float* gIx_cov[ORC_MAX_LEVELS];
  const float* gIy_cov[ORC_MAX_LEVELS];
  const float* W_cur[ORC_MAX_LEVELS];
  const float* I_cur[ORC_MAX_LEVELS];
} orc_pyramids;

typedef struct {
  int level, iter;
  double sums27[27];
  float sigma_int, sigma_depthinv, bias_int, bias_depthinv, nu_int, nu_depthinv;
  int irls_iters_int, irls_iters_depthinv;
  double x[6];
  double R[9], t[3];
} orc_iter_trace;

typedef struct {
  double cov_sums27[27];
  float chi_square, chi_test, ndof;
} orc_frame_stats;

void orc_set_tex_frac_mode(int mode);

void orc_depth_to_invdepth(const uint16_t* src, float* dst, int rows, int cols, float factor_depth);
void orc_intensity(const uint8_t* rgb, float* dst, int rows, int cols);
void orc_decompose_rgb(const uint8_t* rgb, float* r, float* g, float* b, int rows, int cols);
void orc_pyr_down(const float* src, int srows, int scols, float* dst);
void orc_gradient(const float* src, int rows, int cols, float* gx, float* gy);
void orc_bilateral(const float* src, int rows, int cols, float* dst, float sigma_floatmap);

void orc_warp_invdepth(const float* src, const float* depth_prev, float* dst, int rows, int cols,
                       const float* Rp, const float* tp);
void orc_warp_intensity(const float* src, const float* depth_prev, float* dst, int rows, int cols,
                        const float* Rp, const float* tp);
void orc_warp_invdepth_weighted(const float* src, const float* depth_prev, float* dst,
                                float* weight_warped, int rows, int cols, const float* Rp,
                                const float* tp);
void orc_integrate_warped_frame(const float* wsrc, const float* wweight, float* dst, float* dweight,
                                int rows, int cols);
float orc_visibility_ratio(const float* depth_src, const float* depth_dst, int rows, int cols,
                           const float* Rp, const float* tp, uint8_t* overlap_mask,
                           double* n_visible, double* n_valid);

void orc_error_geometry(int rows, int cols, int min_nsamples, int* kept_rows, int* kept_cols,
                        int* stride);
int orc_compute_error(const float* im1, const float* im0, int rows, int cols, int min_nsamples,
                      float* error);
double orc_digamma(double x);
int orc_sigma_nu_student(const float* err, int n, float* bias, float* sigma, float* nu, int mest);
void orc_nu_student(const float* err, int n, float bias, float sigma, float* nu);
int orc_sigma_pdf(const float* err, int n, float* bias, float* sigma, int mest);
void orc_chi_square(const float* err_int, const float* err_depth, int n, float sigma_int,
                    float sigma_depth, int mest, float* chi_squared, float* chi_test, float* ndof);

void orc_unpack_system(const double* sums27, double* A36, double* b6);
void orc_build_system(const float* W0, const float* I0, const float* gWx, const float* gWy,
                      const float* gIx, const float* gIy, const float* W1, const float* I1, int rows,
                      int cols, const orc_system_params* P, double* sums27, double* A36, double* b6);

void orc_vmap(const float* depth_inv, int rows, int cols, float fx, float fy, float cx, float cy,
              float* vmap);
void orc_nmap_gradients(const float* depth_inv, const float* gx, const float* gy, int rows, int cols,
                        float fx, float fy, float cx, float cy, float* nmap);

void orc_mat3_inverse(const double* M, double* Mi);
void orc_force_orthogonal(const double* M, double* R);
void orc_exp_map_rot(const double* omega, double* R);
void orc_exp_map(const double* omega, const double* v, double* R, double* t);
void orc_log_map(const double* R, const double* trans, double* twist);
int orc_llt_solve6(const double* A, const double* b, double* x);
int orc_inverse6(const double* A, double* Ai);
void orc_projective_pose(const double* R, const double* t, float fx, float fy, float cx, float cy,
                         float* Rp, float* tp);
void orc_projective_inverse_pose(const double* R, const double* t, float fx, float fy, float cx,
                                 float cy, float* Rp, float* tp);
int orc_gn_update(const double* A36, const double* b6, double* R, double* t, double* x_out);

void orc_set_num_threads(int n); /* OpenMP threads of the row-parallel loops (no-op without OpenMP) */
int orc_max_threads(void);

int orc_align(const orc_align_config* C, const orc_pyramids* P, double* R, double* t, double* cov36,
              orc_iter_trace* trace, int trace_cap, int* n_trace, orc_frame_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
