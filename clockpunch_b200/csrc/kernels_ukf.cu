// K4: per-target unscented Kalman filter predict + update, fused with the target's
// truth propagation.
//
// Reference: UnscentedKalmanFilter.predictAndUpdate (punchclock/estimation/ukf_v2.py:185-375)
// for the filter ezUKF builds (6-D state, identity measurement, ez_ukf.py:17-95), called once
// per target per step from Target.updateNonPhysical (common/agents.py:144-171), plus
// Agent.propagate for the truth state (agents.py:59-71).
//
// Mapping: 16 lanes per target, two targets per warp, and NO block-level synchronisation --
// every warp owns its two targets from load to store (only __syncwarp), so the 16-20 resident
// warps of an SM drift through the phases independently and hide each other's latencies:
//   load     the warp's est_p tile (2*36 contiguous doubles) + SoA state columns -> shared memory
//   factor   upper Cholesky of est_p, evaluated in every lane of the half-warp (same cost as
//            one lane: the two targets of the warp are factored in one SIMD pass)
//   sigma    lanes 0..12 = the 13 sigma points, lane 13 = the truth state: each lane integrates
//            its own state with every DOP853 stage in registers
//   unscent  mean from differences to the centre point, residuals, and the 21 unique entries
//            of pred_p spread over the 16 lanes (13-term sums through shared memory)
//   update   one lane per tasked target: measurement update (second Cholesky, S, C, gain via
//            triangular solves, est_x, est_p, NIS)
//   store    coalesced 2*36-double tiles for est_p / pred_p, SoA columns for the states
//
// Numerics: the reference forms pred_x = sum_i w_i X_i with w_0 = -2e6, w_i = +1.67e5, which
// loses ~1e-6 km to cancellation.  Here the same quantity is evaluated as
//   X_0 + [w_i sum_i (X_i - X_0) + (sum_i w_i - 1) X_0]
// which is algebraically identical (the weights are the reference's own float64 values and
// eps_w = sum w - 1 is computed exactly from them) but carries no cancellation noise.  The
// resampled measurement sigma points of forecast() (ukf_v2.py:211-228) are symmetric about
// pred_x, so their weighted sums are expanded analytically in the same exact-weights form
// (derivation in DESIGN.md "UKF update").
#include "propagate.cuh"

namespace pc {

struct UkfArgs {
  double* est_x;
  double* est_p;
  double* truth_x;
  const double* z;
  const uint8_t* tasked;
  int64_t n, ld;
  double dt, t_now;
  int substeps;
  uint64_t rng_seed, rng_step;
  double* out_pred_x;
  double* out_pred_p;
  double* out_nis;
  uint32_t* out_flags;
  double* out_z;
  int64_t* num_tasked;
  float* last_time_tasked;
  float* tr_pos;
  float* tr_vel;
  const pc_clock* clock;
  pc_ukf_params p;
};

struct TargetSmem {
  double X[14][6];  // propagated sigma points (0..12) and truth (13); 0..12 become residuals
  double U[21];     // packed upper Cholesky factor (row-major, j >= i)
  double x0[6];     // in: est_x; out: est_x
  double xc[6];     // propagated centre point X_0
  double px[6];     // pred_x
  double m[6];      // pred_x - X_0
  double z[6];
  double nis;
  int tasked;
  uint32_t flags;
  double* P;        // -> this target's 36 doubles of WarpSmem::P  (in: est_p, then pred_p)
  double* PE;       // -> WarpSmem::PE (out: est_p)
};

struct WarpSmem {
  double P[2][36];   // contiguous pair of covariance tiles: coalesced 576-byte load/store
  double PE[2][36];
  TargetSmem t[2];
};

__host__ __device__ constexpr int uidx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

// Upper Cholesky A = U^T U of a 6x6 symmetric matrix (scipy.linalg.cholesky default,
// ukf_v2.py:168).  Returns false if a pivot is not positive (reference: LinAlgError).
__device__ __forceinline__ bool chol_upper6(const double* __restrict__ A, double* __restrict__ U) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = A[i * 6 + i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-U[uidx(k, i)], U[uidx(k, i)], s);
    ok = ok && (s > 0.0);
    const double d = rsqrt(s);
    U[uidx(i, i)] = s * d;
#pragma unroll
    for (int j = i + 1; j < 6; ++j) {
      double a = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < i; ++k) a = fma(-U[uidx(k, i)], U[uidx(k, j)], a);
      U[uidx(i, j)] = a * d;
    }
  }
  return ok;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: no state to keep per target.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Six independent N(0,1) draws keyed by (seed, step, target).
__device__ __forceinline__ void normal6(uint64_t seed, uint64_t step, uint64_t target,
                                        double (&nrm)[6]) {
  uint32_t u[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t c[4] = {(uint32_t)target, (uint32_t)(target >> 32), (uint32_t)step,
                     (uint32_t)(step >> 32) ^ (0x5bd1e995u * (uint32_t)h + (uint32_t)h)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int q = 0; q < 4; ++q) u[4 * h + q] = c[q];
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double u1 = ((double)u[2 * q] + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)u[2 * q + 1] + 0.5) * (1.0 / 4294967296.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    nrm[2 * q] = rad * cs;
    nrm[2 * q + 1] = rad * sn;
  }
}

// Measurement update for one tasked target, everything addressed in shared memory.
// forecast / calculateKalmanGain / updateCovariance / calculateInnovations /
// updateStateEstimate (ukf_v2.py:211-228, 294-375).
__device__ __noinline__ void measurement_update(TargetSmem* __restrict__ S, const pc_ukf_params& p) {
  uint32_t flags = 0;
  // STEP 1: re-sample around (pred_x, pred_p): U2 = chol_upper(pred_p)
  double* U2 = S->U;
  if (!chol_upper6(S->P, U2)) flags |= PC_FLAG_NONPD_PRED;

  const double c3 = p.wc0 - p.wm0;
  const double cuu = 2.0 * p.wi * p.gamma * p.gamma;
  const double cdd = p.wc0 + 12.0 * p.wi;
  const double wg = p.wi * p.gamma;
  double delta[6], nu[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    delta[c] = p.eps_w * S->px[c];                  // mean_pred_y - pred_x
    nu[c] = S->z[c] - (S->px[c] + delta[c]);        // innovation (ukf_v2.py:341-343)
  }
  // Innovation covariance S = Yres Wc Yres^T + R  (ukf_v2.py:326-329), upper part.
  double Sm[36];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int j = b; j < 6; ++j) acc = fma(U2[uidx(a, j)], U2[uidx(b, j)], acc);
      Sm[a * 6 + b] = fma(cuu, acc, fma(cdd * delta[a], delta[b], p.R[a * 6 + b]));
    }
  double Us[21];
  if (!chol_upper6(Sm, Us)) flags |= PC_FLAG_NONPD_INNOV;
  double dinv[6];
#pragma unroll
  for (int b = 0; b < 6; ++b) dinv[b] = 1.0 / Us[uidx(b, b)];

  // y = Us^-T nu ; nis = |y|^2  (ukf_v2.py:344-346)
  double y[6];
  double nis = 0.0;
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    double acc = nu[b];
#pragma unroll
    for (int k = 0; k < b; ++k) acc = fma(-Us[uidx(k, b)], y[k], acc);
    y[b] = acc * dinv[b];
    nis = fma(y[b], y[b], nis);
  }
  S->nis = nis;

  // Row by row: cross covariance C = Xres Wc Yres^T (ukf_v2.py:330-332), G = C Us^-1.
  // K = G Us^-T, so K S K^T = G G^T and K nu = G y.
#pragma unroll 1
  for (int a = 0; a < 6; ++a) {
    double D[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) D[j] = S->X[1 + j][a] - S->X[7 + j][a];
    const double sa = delta[a] + c3 * S->m[a];
    double g[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int j = b; j < 6; ++j) acc = fma(D[j], U2[uidx(b, j)], acc);
      double cab = fma(wg, acc, sa * delta[b]);
#pragma unroll
      for (int k = 0; k < b; ++k) cab = fma(-g[k], Us[uidx(k, b)], cab);
      g[b] = cab * dinv[b];
    }
    double xe = S->px[a];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      xe = fma(g[b], y[b], xe);
      S->PE[a * 6 + b] = g[b];  // stash G row; est_p assembled below
    }
    S->x0[a] = xe;  // est_x = pred_x + K nu  (ukf_v2.py:373-375)
  }
  // est_p = pred_p - K S K^T = pred_p - G G^T  (ukf_v2.py:367-371)
  double Gm[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) Gm[k] = S->PE[k];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double acc = S->P[a * 6 + b];
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma(-Gm[a * 6 + k], Gm[b * 6 + k], acc);
      S->PE[a * 6 + b] = acc;
      S->PE[b * 6 + a] = acc;
    }
  S->flags |= flags;
}

// (a, b) of the e-th unique entry of a symmetric 6x6 (row-major upper), 3 bits each.
__device__ __forceinline__ void sym_index(int e, int& ra, int& rb) {
  constexpr unsigned long long RA = 0ull | (0ull << 0) | (0ull << 3) | (0ull << 6) | (0ull << 9) |
      (0ull << 12) | (0ull << 15) | (1ull << 18) | (1ull << 21) | (1ull << 24) | (1ull << 27) |
      (1ull << 30) | (2ull << 33) | (2ull << 36) | (2ull << 39) | (2ull << 42) | (3ull << 45) |
      (3ull << 48) | (3ull << 51) | (4ull << 54) | (4ull << 57) | (5ull << 60);
  constexpr unsigned long long RB = 0ull | (0ull << 0) | (1ull << 3) | (2ull << 6) | (3ull << 9) |
      (4ull << 12) | (5ull << 15) | (1ull << 18) | (2ull << 21) | (3ull << 24) | (4ull << 27) |
      (5ull << 30) | (2ull << 33) | (3ull << 36) | (4ull << 39) | (5ull << 42) | (3ull << 45) |
      (4ull << 48) | (5ull << 51) | (4ull << 54) | (5ull << 57) | (5ull << 60);
  ra = (int)((RA >> (3 * e)) & 7ull);
  rb = (int)((RB >> (3 * e)) & 7ull);
}

template <int WARPS, int MINB, bool J2>
__global__ void __launch_bounds__(WARPS * 32, MINB)
ukf_step_kernel(const __grid_constant__ UkfArgs a) {
  __shared__ WarpSmem wsm[WARPS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int h = lane >> 4, l = lane & 15;
  const int64_t pair = (int64_t)blockIdx.x * WARPS + wib;  // this warp's pair of targets
  const int64_t tg = 2 * pair + h;
  if (2 * pair >= a.n) return;  // whole warp idle (warp-uniform)
  const bool live = tg < a.n;
  const int nlive = (2 * pair + 1 < a.n) ? 2 : 1;
  WarpSmem& W = wsm[wib];
  TargetSmem* S = &W.t[h];
  const double t_now = a.clock ? a.clock->t_now + a.dt : a.t_now;
  const uint64_t rng_step = a.clock ? a.clock->step : a.rng_step;

  // ---- load ---------------------------------------------------------------------------
  {
    const double* gp = a.est_p + 2 * pair * 36;
    double* sp = &W.P[0][0];
    for (int k = lane; k < nlive * 36; k += 32) sp[k] = gp[k];
    if (live && l < 6) S->x0[l] = a.est_x[(int64_t)l * a.ld + tg];
    if (l == 6) {
      S->flags = 0;
      S->nis = __longlong_as_double(0x7ff8000000000000LL);
      S->tasked = (live && a.tasked) ? (a.tasked[tg] != 0) : 0;
      S->P = W.P[h];
      S->PE = W.PE[h];
    }
  }
  __syncwarp();

  // ---- factor + sigma points (generateSigmaPoints, ukf_v2.py:158-183) ---------------------
  double r[3], v[3];
  bool work = false;
  {
    double U[21];
    const bool ok = chol_upper6(W.P[h], U);
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < 21; ++k) S->U[k] = U[k];
      if (live && !ok) S->flags |= PC_FLAG_NONPD_EST;
    }
  }
  __syncwarp();
  if (live && l < 13) {
    // x + gamma * [0, U, -U]: COLUMNS of the upper factor
    const int j = (l == 0) ? 0 : (l - 1) % 6;
    const double sg = (l == 0) ? 0.0 : (l <= 6 ? a.p.gamma : -a.p.gamma);
    double x[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double u = (c <= j) ? S->U[uidx(c, j)] : 0.0;
      x[c] = fma(sg, u, S->x0[c]);
    }
    r[0] = x[0]; r[1] = x[1]; r[2] = x[2];
    v[0] = x[3]; v[1] = x[4]; v[2] = x[5];
    work = true;
  } else if (live && l == 13 && a.truth_x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = a.truth_x[(int64_t)c * a.ld + tg];
      v[c] = a.truth_x[(int64_t)(c + 3) * a.ld + tg];
    }
    work = true;
  } else {
    r[0] = 7000.0; r[1] = 0.0; r[2] = 0.0;
    v[0] = 0.0; v[1] = 7.5; v[2] = 0.0;
  }
  // substeps: one SIMD evaluation for all lanes; the 13 sigma points share the centre's count
  int ns = a.substeps;
  if (ns <= 0) {
    bool cl;
    ns = auto_substeps(r, v, a.dt, &cl);
    const int nc = __shfl_sync(0xffffffffu, ns, h << 4);
    const bool clc = __shfl_sync(0xffffffffu, (int)cl, h << 4) != 0;
    if (l < 13) ns = nc;
    if (live && l == 0 && clc) S->flags |= PC_FLAG_SUBSTEP_CLAMP;
  }
  if (work) {
    propagate_state<J2>(r, v, a.dt, ns);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      S->X[l][c] = r[c];
      S->X[l][c + 3] = v[c];
    }
  }
  __syncwarp();

  // ---- unscented transform (predictStateEstimate / predictCovariance, ukf_v2.py:261-292) ----
  double X0[6], d[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    X0[c] = S->X[0][c];
    d[c] = (l >= 1 && l < 13) ? S->X[l][c] - X0[c] : 0.0;
  }
  __syncwarp();
  if (live && l < 13) {
#pragma unroll
    for (int c = 0; c < 6; ++c) S->X[l][c] = d[c];
  }
  __syncwarp();
  if (live && l < 6) {
    double s = 0.0;
#pragma unroll
    for (int i = 1; i < 13; ++i) s += S->X[i][l];
    const double m = fma(a.p.wi, s, a.p.eps_w * X0[l]);
    S->xc[l] = X0[l];
    S->m[l] = m;
    S->px[l] = X0[l] + m;  // pred_x = sigma_points . mean_weight
  }
  __syncwarp();
  if (live && l < 13) {
#pragma unroll
    for (int c = 0; c < 6; ++c) S->X[l][c] = d[c] - S->m[c];  // sigma_x_res
  }
  __syncwarp();
  if (live) {
    // pred_p = Xres Wc Xres^T + Q: 21 unique entries over 16 lanes
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int e = l + 16 * rep;
      if (e < 21) {
        int ra, rb;
        sym_index(e, ra, rb);
        double acc = 0.0;
#pragma unroll
        for (int i = 1; i < 13; ++i) acc = fma(S->X[i][ra], S->X[i][rb], acc);
        acc = fma(a.p.wi, acc, (a.p.wc0 * S->X[0][ra]) * S->X[0][rb]) + a.p.Q[ra * 6 + rb];
        S->P[ra * 6 + rb] = acc;
        S->P[rb * 6 + ra] = acc;
      }
    }
  }
  __syncwarp();

  // ---- measurement / no-measurement update, one lane per target ----------------------------
  if (live && l == 0) {
    if (S->tasked) {
      if (a.z) {
#pragma unroll
        for (int c = 0; c < 6; ++c) S->z[c] = a.z[(int64_t)c * a.ld + tg];
      } else {
        // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
        double nrm[6];
        normal6(a.rng_seed, rng_step, (uint64_t)tg, nrm);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double acc = S->X[13][c];
#pragma unroll
          for (int k = 0; k <= c; ++k) acc = fma(a.p.Lr[c * 6 + k], nrm[k], acc);
          S->z[c] = acc;
        }
      }
      measurement_update(S, a.p);
    } else {
      // update(None) (ukf_v2.py:238-244): est_x = propagated centre point, est_p = pred_p
#pragma unroll
      for (int c = 0; c < 6; ++c) S->x0[c] = S->xc[c];
    }
  }
  __syncwarp();
  if (live && !S->tasked) {
    for (int k = l; k < 36; k += 16) S->PE[k] = S->P[k];
  }
  __syncwarp();
  if (live && l == 0) {
    // flags + scheduler summaries (agents.py:157-171; env.py:446-452,599-609)
    bool fin = true;
#pragma unroll
    for (int c = 0; c < 6; ++c) fin = fin && isfinite(S->x0[c]);
#pragma unroll
    for (int k = 0; k < 36; k += 7) fin = fin && isfinite(S->PE[k]);
    if (!fin) S->flags |= PC_FLAG_NONFINITE;
    if (a.out_flags) a.out_flags[tg] = S->flags;
    if (a.out_nis) a.out_nis[tg] = S->nis;
    if (a.tr_pos) a.tr_pos[tg] = (float)S->PE[0] + (float)S->PE[7] + (float)S->PE[14];
    if (a.tr_vel) a.tr_vel[tg] = (float)S->PE[21] + (float)S->PE[28] + (float)S->PE[35];
    if (S->tasked) {
      if (a.num_tasked) a.num_tasked[tg] += 1;
      if (a.last_time_tasked) a.last_time_tasked[tg] = (float)t_now;
    }
  }

  // ---- store ---------------------------------------------------------------------------
  {
    double* ge = a.est_p + 2 * pair * 36;
    const double* se = &W.PE[0][0];
    for (int k = lane; k < nlive * 36; k += 32) ge[k] = se[k];
    if (a.out_pred_p) {
      double* gq = a.out_pred_p + 2 * pair * 36;
      const double* sq = &W.P[0][0];
      for (int k = lane; k < nlive * 36; k += 32) gq[k] = sq[k];
    }
    if (live && l < 6) {
      const int64_t g = (int64_t)l * a.ld + tg;
      a.est_x[g] = S->x0[l];
      if (a.truth_x) a.truth_x[g] = S->X[13][l];
      if (a.out_pred_x) a.out_pred_x[g] = S->px[l];
      if (a.out_z && S->tasked) a.out_z[g] = S->z[l];
    }
  }
}

cudaError_t launch_ukf(double* est_x, double* est_p, double* truth_x, const double* z,
                       const uint8_t* tasked, int64_t n, int64_t ld, double t0, double tf,
                       int substeps, int force_model, const pc_ukf_params* p, uint64_t seed,
                       uint64_t step, double* pred_x, double* pred_p, double* nis, uint32_t* flags,
                       double* out_z, int64_t* num_tasked, float* last_time_tasked, float* tr_pos,
                       float* tr_vel, const pc_clock* clock, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  UkfArgs a;
  a.est_x = est_x;
  a.est_p = est_p;
  a.truth_x = truth_x;
  a.z = z;
  a.tasked = tasked;
  a.n = n;
  a.ld = ld;
  a.dt = tf - t0;
  a.t_now = tf;
  a.substeps = substeps;
  a.rng_seed = seed;
  a.rng_step = step;
  a.out_pred_x = pred_x;
  a.out_pred_p = pred_p;
  a.out_nis = nis;
  a.out_flags = flags;
  a.out_z = out_z;
  a.num_tasked = num_tasked;
  a.last_time_tasked = last_time_tasked;
  a.tr_pos = tr_pos;
  a.tr_vel = tr_vel;
  a.clock = clock;
  a.p = *p;
  constexpr int WARPS = 2, MINB = 8;  // 2 warps = 4 targets per CTA; 16 warps resident per SM
  const int64_t pairs = (n + 1) / 2;
  const unsigned blocks = (unsigned)((pairs + WARPS - 1) / WARPS);
  if (force_model == PC_FORCE_TWOBODY_J2)
    ukf_step_kernel<WARPS, MINB, true><<<blocks, WARPS * 32, 0, stream>>>(a);
  else
    ukf_step_kernel<WARPS, MINB, false><<<blocks, WARPS * 32, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pc
