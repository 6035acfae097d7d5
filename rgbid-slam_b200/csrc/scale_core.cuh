// scale_core.cuh -- robust scale (bias, sigma) and Student-t nu estimation inside one thread-block cluster.
//
// Reference behaviour (src/cuda/sigmaFuncs.cu:858-1222): up to 10 IRLS iterations for (bias, sigma) and up
// to 6 evaluations of the nu likelihood equation, each a pair of tiny kernels + cudaStreamSynchronize +
// device->host copy, with host-side control flow (bisection, float digamma).  Here the whole procedure
// runs inside ONE kernel: an 8-CTA cluster owns the residual samples of one frame pair (both residual
// types), keeps them in shared memory, and performs one cluster-wide reduction per round through
// distributed shared memory; the control flow (convergence test, bisection) is evaluated redundantly and
// identically by two warps of every CTA (one per residual type), so no host round trip is needed.
#pragma once
#include <cooperative_groups.h>
#if RGBID_SCALE_PROBE
#include <cstdio>  // -DRGBID_SCALE_PROBE=1: per-segment clock counts of the rounds, tools/scale_round_probe.py
#endif
#include "kernels.cuh"

namespace rgbid {

namespace cg = cooperative_groups;

#ifndef RGBID_SCALE_CLUSTER
#define RGBID_SCALE_CLUSTER 8
#endif
#ifndef RGBID_SCALE_THREADS
#define RGBID_SCALE_THREADS 512
#endif
#ifndef RGBID_SCALE_MINB
#define RGBID_SCALE_MINB 2
#endif
constexpr int kScaleCluster = RGBID_SCALE_CLUSTER;  // CTAs per frame pair (8 = portable cluster size limit; 16 needs the non-portable attribute)
constexpr int kScaleMinBlocks = RGBID_SCALE_MINB;
constexpr int kScaleThreads = RGBID_SCALE_THREADS;  // threads per CTA (256 measured the same: the per-SM instruction total of a round does not change, see profiles/README.md)
constexpr int kScaleVals = 12;      // reduced values per round (6 per residual slot)

// C(nu) = -psi(nu/2) + ln(nu/2) + f + 1 + psi((nu+1)/2) - ln((nu+1)/2)   (sigmaFuncs.cu:966), float arithmetic
// evaluated left to right.  psi is boost::math::digamma on a float argument (src/cuda/device.hpp:76-80), which
// evaluates in double and rounds to float.  The bisection of sigmaFuncs.cu:946-1047 only ever evaluates C at
// nu in {2, 2.5, 3, ..., 10}, so the four nu-dependent terms are tabulated on the host once (17 x 4 floats in
// constant memory) instead of running a double-precision digamma in every thread every round.
struct NuTerms { float neg_psi_half, log_half, psi_half1, log_half1; };
static __constant__ NuTerms c_nu_table[17];  // one copy per translation unit (internal linkage)

static inline double digamma_host(double x)
{
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
             f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

// Must be called once per translation unit that launches a kernel using c_nu_float (the table has internal
// linkage, so this function must too).
static bool upload_nu_table()
{
  NuTerms h[17];
  for (int k = 0; k < 17; ++k) {
    float nu = 2.f + 0.5f * (float)k;
    float a = nu / 2.f, b = (nu + 1.f) / 2.f;
    h[k].neg_psi_half = -(float)digamma_host((double)a);
    h[k].log_half = logf(a);
    h[k].psi_half1 = (float)digamma_host((double)b);
    h[k].log_half1 = logf(b);
  }
  return cudaMemcpyToSymbol(c_nu_table, h, sizeof(h)) == cudaSuccess;  // synchronous, legacy stream: never inside a capture
}

__device__ __forceinline__ float c_nu_float(float nu, float fw)
{
  const NuTerms t = c_nu_table[__float2int_rn((nu - 2.f) * 2.f)];
  float a = __fadd_rn(t.neg_psi_half, t.log_half);
  a = __fadd_rn(a, fw);
  a = __fadd_rn(a, 1.f);
  a = __fadd_rn(a, t.psi_half1);
  return __fsub_rn(a, t.log_half1);
}

enum ScalePhase { PH_IRLS = 0, PH_NU_INIT = 1, PH_NU_BISECT = 2, PH_DONE = 3 };

struct ScaleSlot {
  int phase, op, mest;
  int it, j, lsq, irls_iters;
  float bias, sigma;          // sh.bias / sh.sigma of the reference's handler
  float out_bias, out_sigma;  // values returned to the caller
  float nu, nu_up, nu_down, nu_new, C_up, C_down;
};

__device__ __forceinline__ void slot_init(ScaleSlot& s, int op, int mest, float bias, float sigma, bool enabled)
{
  s.op = op; s.mest = mest;
  s.it = 0; s.j = 0; s.lsq = 1; s.irls_iters = 0;
  s.bias = bias; s.sigma = sigma; s.out_bias = bias; s.out_sigma = sigma;
  s.nu = 5.f; s.nu_up = 10.f; s.nu_down = 2.f; s.nu_new = 0.f; s.C_up = 0.f; s.C_down = 0.f;
  s.phase = !enabled ? PH_DONE : (op == SCALE_NU_ONLY ? PH_NU_INIT : PH_IRLS);
}

// Per-sample contribution of one round.  acc[0..5]:
//   IRLS     : sum w r^2, sum w r, sum w, N
//   NU_INIT  : sum ln w(nu=2), sum w(2), sum ln w(10), sum w(10), N
//   NU_BISECT: sum ln w(nu_new), sum w(nu_new), -, -, N
__device__ __forceinline__ void slot_accumulate(const ScaleSlot& s, float e, float* acc)
{
  if (isinf(e) || isnan(e)) return;  // sigmaFuncs.cu:206,308,436
  if (s.phase == PH_IRLS) {
    float weight = 1.f, valid = 1.f;
    if (s.op == SCALE_SIGMA_NU) {
      // partialBiasAndSigmaStudent, sigmaFuncs.cu:281-360 (nu fixed at 5 during IRLS, :907)
      if (!s.lsq) {
        float en = (e - s.bias) / s.sigma;
        weight = (5.f + 1.f) / (5.f + en * en);
      }
    } else {
      // partialBiasAndSigma, sigmaFuncs.cu:179-278 (first iteration runs with Mestimator = LSQ, :813)
      float en = (e - s.bias) / s.sigma;
      int m = s.lsq ? RGBID_LSQ : s.mest;
      if (m == RGBID_HUBER) { if (fabsf(en) > 1.345f) weight = 1.345f / fabsf(en); }
      else if (m == RGBID_TUKEY) {
        if (fabsf(en) < 4.685f) { float a = (en / 4.685f) * (en / 4.685f); weight = (1.f - a) * (1.f - a); }
        else { weight = 0.f; valid = 0.f; }
      } else if (m == RGBID_STUDENT) weight = (5.f + 1.f) / (5.f + en * en);
    }
    float wr = e * weight;
    acc[0] += wr * e; acc[1] += wr; acc[2] += weight; acc[5] += valid;
  } else {
    // partialFuncWeightsNu, sigmaFuncs.cu:412-475
    float en = (e - s.bias) / s.sigma;
    float e2 = en * en;
    if (s.phase == PH_NU_INIT) {
      float w2 = (2.f + 1.f) / (2.f + e2), w10 = (10.f + 1.f) / (10.f + e2);
      acc[0] += logf(w2); acc[1] += w2; acc[2] += logf(w10); acc[3] += w10;
    } else {
      float w = (s.nu_new + 1.f) / (s.nu_new + e2);
      acc[0] += logf(w); acc[1] += w;
    }
    acc[5] += 1.f;
  }
}

// RGBID_FAST_NU_LOG=1 (default): ln of the Student weights through lg2.approx (3 instructions, 2 ulp) instead of logf
// (~20 instructions) in the nu rounds.  The sums feed only the SIGN of C(nu) in the bisection (nu itself is one of
// {2, 2.5, ..., 10}); the parity suite and tools/stress_parity.py see identical nu with both.  The kernel is half issue
// bound: 21.6 -> 19.7 us per launch at level 0, 35.0 -> 28.3 us at levels 1 and 2 (tools/gpu_batch48.sh).
#ifndef RGBID_FAST_NU_LOG
#define RGBID_FAST_NU_LOG 1
#endif
#if RGBID_FAST_NU_LOG
#define RGBID_NU_LOG(x) __logf(x)
#else
#define RGBID_NU_LOG(x) logf(x)
#endif

// One round over this CTA's samples.  The phase is uniform over the cluster, so it is tested once per round, not
// once per sample; the Student-t phases (every round of the shipped configuration) are written out with the
// divisions as multiplications by one reciprocal -- what div.approx computes, minus its per-call range fix-up.
// This loop is where the scale kernel spends its time (38 400 samples x ~14 rounds per frame pair and launch).
__device__ __forceinline__ void slot_accumulate_all(const ScaleSlot& s, const float* __restrict__ samples, int n, float* acc)
{
  const int tid = threadIdx.x;
  if (s.phase == PH_IRLS && s.op == SCALE_SIGMA_NU) {
    // partialBiasAndSigmaStudent, sigmaFuncs.cu:281-360 (nu fixed at 5 during IRLS, :907)
    if (s.lsq) {
#pragma unroll 5
      for (int i = tid; i < n; i += kScaleThreads) {
        const float e = samples[i];
        if (fabsf(e) < __int_as_float(0x7f800000)) {  // finite (sigmaFuncs.cu:308)
          acc[0] += e * e; acc[1] += e; acc[2] += 1.f; acc[5] += 1.f;
        }
      }
    } else {
      const float bias = s.bias, rsigma = 1.f / s.sigma;
#pragma unroll 5
      for (int i = tid; i < n; i += kScaleThreads) {
        const float e = samples[i];
        if (fabsf(e) < __int_as_float(0x7f800000)) {
          const float en = (e - bias) * rsigma;
          const float weight = (5.f + 1.f) * (1.f / (5.f + en * en));
          const float wr = e * weight;
          acc[0] += wr * e; acc[1] += wr; acc[2] += weight; acc[5] += 1.f;
        }
      }
    }
  } else if (s.phase == PH_NU_INIT) {
    // partialFuncWeightsNu, sigmaFuncs.cu:412-475 at nu = 2 and nu = 10
    const float bias = s.bias, rsigma = 1.f / s.sigma;
#pragma unroll 5
    for (int i = tid; i < n; i += kScaleThreads) {
      const float e = samples[i];
      if (fabsf(e) < __int_as_float(0x7f800000)) {
        const float en = (e - bias) * rsigma;
        const float e2 = en * en;
        const float w2 = (2.f + 1.f) * (1.f / (2.f + e2)), w10 = (10.f + 1.f) * (1.f / (10.f + e2));
        acc[0] += RGBID_NU_LOG(w2); acc[1] += w2; acc[2] += RGBID_NU_LOG(w10); acc[3] += w10; acc[5] += 1.f;
      }
    }
  } else if (s.phase == PH_NU_BISECT) {
    const float bias = s.bias, rsigma = 1.f / s.sigma, nu = s.nu_new, nu1 = s.nu_new + 1.f;
#pragma unroll 5
    for (int i = tid; i < n; i += kScaleThreads) {
      const float e = samples[i];
      if (fabsf(e) < __int_as_float(0x7f800000)) {
        const float en = (e - bias) * rsigma;
        const float w = nu1 * (1.f / (nu + en * en));
        acc[0] += RGBID_NU_LOG(w); acc[1] += w; acc[5] += 1.f;
      }
    }
  } else {
    for (int i = tid; i < n; i += kScaleThreads) slot_accumulate(s, samples[i], acc);
  }
}

// Fixed-order sum of N doubles read through `at(i)` as four interleaved partial sums: a dependent chain of N / 4 + 2
// additions instead of N (a double add is ~35 clk of latency here and one thread per value is doing this).
template <int N, class F>
__device__ __forceinline__ double sum4(F at)
{
  double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
  for (int i = 0; i < N; i += 4) { q0 += at(i); q1 += at(i + 1); q2 += at(i + 2); q3 += at(i + 3); }
  return (q0 + q1) + (q2 + q3);
}

// Host-side control flow of computeSigmaAndNuStudent / computeSigmaPdf / computeNuStudent, advanced by
// one round given the cluster-wide totals tot[0..5] of this slot.
__device__ __forceinline__ void slot_advance(ScaleSlot& s, const double* tot)
{
  if (s.phase == PH_IRLS) {
    // finalReductionBiasAndSigma, sigmaFuncs.cu:362-409 (float arithmetic on the reduced sums)
    float fwr2 = (float)tot[0], fwr = (float)tot[1], fw = (float)tot[2], fn = (float)tot[5];
    float m0 = fwr / fw;
    float m1 = sqrtf((fwr2 - 2.f * m0 * fwr + m0 * m0 * fw) / fn);
    s.out_bias = m0; s.out_sigma = m1;
    float sigma_prev = s.sigma;
    s.bias = m0; s.sigma = m1;
    s.lsq = (s.op == SCALE_SIGMA_NU) ? (s.mest == RGBID_LSQ) : 0;
    ++s.irls_iters;
    bool conv = (s.it > 0) && ((fabsf(m1 - sigma_prev) / sigma_prev) < 0.1f);
    ++s.it;
    if (conv || s.it >= 10) s.phase = (s.op == SCALE_SIGMA_NU) ? PH_NU_INIT : PH_DONE;
  } else if (s.phase == PH_NU_INIT) {
    float fn = (float)tot[5];
    float f_down = ((float)tot[0] - (float)tot[1]) / fn;  // finalReductionFuncWeightsNu, :478-516
    float f_up = ((float)tot[2] - (float)tot[3]) / fn;
    s.C_down = c_nu_float(2.f, f_down);
    s.C_up = c_nu_float(10.f, f_up);
    if (s.C_up * s.C_down > 0.f) {
      s.nu = (s.C_down <= 0.f) ? 2.f : 10.f;
      s.phase = PH_DONE;
    } else {
      s.j = 0;
      s.nu_new = (s.nu_up + s.nu_down) / 2.f;
      if ((s.nu_up - s.nu_down) < 1.f) { s.nu = s.nu_new; s.phase = PH_DONE; }
      else s.phase = PH_NU_BISECT;
    }
  } else if (s.phase == PH_NU_BISECT) {
    float fn = (float)tot[5];
    float f_new = ((float)tot[0] - (float)tot[1]) / fn;
    float C_new = c_nu_float(s.nu_new, f_new);
    if (C_new * s.C_up > 0.f) { s.C_up = C_new; s.nu_up = s.nu_new; }
    else { s.C_down = C_new; s.nu_down = s.nu_new; }
    ++s.j;
    if (s.j >= 5) { s.nu = s.nu_new; s.phase = PH_DONE; }
    else {
      s.nu_new = (s.nu_up + s.nu_down) / 2.f;
      if ((s.nu_up - s.nu_down) < 1.f) { s.nu = s.nu_new; s.phase = PH_DONE; }
    }
  }
}

struct ScaleShared {
  double warp_part[kScaleThreads / 32][kScaleVals];
  double gath[2][kScaleCluster][kScaleVals];  // every CTA's totals, pushed by the peers through DSMEM; double-buffered
  double total[kScaleVals];
  ScaleSlot slots[2];  // the two slots' state, advanced by warps 0 and 1 after every round
};

// Runs all rounds for the two slots.  samples0/1: this CTA's slice of each residual vector (shared or global
// memory), n_local entries each.  Every thread of every CTA of the cluster must call this.
__device__ __forceinline__ void scale_rounds(cg::cluster_group& cluster, ScaleShared& sh, ScaleSlot& s0, ScaleSlot& s1,
                                             const float* samples0, const float* samples1, int n_local)
{
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int parity = 0;
#if RGBID_SCALE_PROBE
  long long seg[7] = {0, 0, 0, 0, 0, 0, 0};
  int rounds_done = 0;
#define PROBE(k) { long long now = clock64(); seg[k] += now - tprev; tprev = now; }
#else
#define PROBE(k)
#endif
  // The full state of the two slots lives in shared memory (it is only needed by the two warps that run the control
  // flow); every thread keeps just the fields the sample loops read.
  if (tid == 0) { sh.slots[0] = s0; sh.slots[1] = s1; }
  __syncthreads();
  ScaleSlot v0, v1;
  auto load_view = [](ScaleSlot& v, const ScaleSlot& src) {
    v.phase = src.phase; v.op = src.op; v.mest = src.mest; v.lsq = src.lsq;
    v.bias = src.bias; v.sigma = src.sigma; v.nu_new = src.nu_new;
  };
  load_view(v0, sh.slots[0]);
  load_view(v1, sh.slots[1]);
  // bounded: <= 10 IRLS + 1 + 5 bisection rounds per slot (they advance concurrently)
  for (int round = 0; round < 20; ++round) {
    if (v0.phase == PH_DONE && v1.phase == PH_DONE) break;
#if RGBID_SCALE_PROBE
    long long tprev = clock64();
    ++rounds_done;
#endif
    float acc[kScaleVals];
#pragma unroll
    for (int k = 0; k < kScaleVals; ++k) acc[k] = 0.f;
    if (v0.phase != PH_DONE) slot_accumulate_all(v0, samples0, n_local, acc);
    if (v1.phase != PH_DONE) slot_accumulate_all(v1, samples1, n_local, acc + 6);
    PROBE(0)
    // float inside the warp (<= ~10 samples per lane; the reference sums in float throughout), double above.
    // The 12 sums are reduced together: after the exchange with lane ^ 16 a lane keeps only half of the values, then a
    // quarter, ... -- 8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 12 x 5; lane l ends with the total of value (l >> 1) & 15.
    {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = (k < kScaleVals) ? acc[k] : 0.f;
#pragma unroll
      for (int h = 8, m = 16; h >= 1; h >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = up ? v[i] : v[i + h];
          const float keep = up ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      const int idx = (lane >> 1) & 15;
      if (!(lane & 1) && idx < kScaleVals) sh.warp_part[wid][idx] = (double)v[0];
    }
    PROBE(1)
    __syncthreads();
    PROBE(2)
    if (tid < kScaleVals) {
      const double v = sum4<kScaleThreads / 32>([&](int w) { return sh.warp_part[w][tid]; });
      // push this CTA's total into every CTA of the cluster (remote shared-memory stores: fire and forget, completed
      // by the cluster barrier), instead of every CTA pulling eight remote values after the barrier
      const unsigned me = cluster.block_rank();
#pragma unroll
      for (int r = 0; r < kScaleCluster; ++r) cluster.map_shared_rank(&sh.gath[parity][me][0], r)[tid] = v;
    }
    PROBE(3)
    cluster.sync();
    PROBE(4)
    // the host control flow runs in two warps only, one per residual slot (in every CTA of the cluster: identical
    // inputs, identical decisions); the other warps pick up the new state after the barrier they need anyway
    if (wid < 2) {
      if (lane < 6) {
        const int k = 6 * wid + lane;  // fixed order: identical in every CTA
        sh.total[k] = sum4<kScaleCluster>([&](int r) { return sh.gath[parity][r][k]; });
      }
      __syncwarp();
      ScaleSlot st = sh.slots[wid];
      slot_advance(st, &sh.total[6 * wid]);
      __syncwarp();
      if (lane == 0) sh.slots[wid] = st;
    }
    PROBE(5)
    __syncthreads();
    PROBE(6)
    load_view(v0, sh.slots[0]);
    load_view(v1, sh.slots[1]);
    parity ^= 1;
    // no barrier here: total[] and slots[] are rewritten only after the next round's barriers, warp_part[] after this
    // round's readers have passed the cluster barrier, and gath[] is double-buffered (a CTA is at most one round ahead)
  }
  // peers may still be writing into this CTA's shared memory: do not exit before everybody is done
  cluster.sync();
  s0 = sh.slots[0];
  s1 = sh.slots[1];
#if RGBID_SCALE_PROBE
  if ((blockIdx.x == 0 || blockIdx.x == 131) && (tid == 0 || tid == 200))
    printf("probe cta %d tid %d n_local %d rounds %d | samples %lld reduce %lld sync1 %lld sum+push %lld cluster %lld advance %lld sync2 %lld\n",
           blockIdx.x, tid, n_local, rounds_done, seg[0], seg[1], seg[2], seg[3], seg[4], seg[5], seg[6]);
#endif
}

}  // namespace rgbid
