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
#include "ukf_args.cuh"
#include "vis.cuh"

namespace pc {

struct TargetSmem {
  double P[36];     // in: est_p, then pred_p
  double PE[36];    // est_p after a measurement update
  double X[14][6];  // propagated sigma points (0..12) and truth (13); 0..12 become residuals
  double U[21];     // packed upper Cholesky factor (row-major, j >= i)
  double x0[6];     // in: est_x; out: est_x
  double xc[6];     // propagated centre point X_0
  double px[6];     // pred_x
  double m[6];      // pred_x - X_0
  double z[6];
  double tr[6];     // truth state in
  double vq[2];     // horizon terms of (truth', est_x') for the fused visibility test
  double nis;
  int tasked;
  uint32_t flags;
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
  constexpr unsigned long long RA = (0ull << 0) | (0ull << 3) | (0ull << 6) | (0ull << 9) |
      (0ull << 12) | (0ull << 15) | (1ull << 18) | (1ull << 21) | (1ull << 24) | (1ull << 27) |
      (1ull << 30) | (2ull << 33) | (2ull << 36) | (2ull << 39) | (2ull << 42) | (3ull << 45) |
      (3ull << 48) | (3ull << 51) | (4ull << 54) | (4ull << 57) | (5ull << 60);
  constexpr unsigned long long RB = (0ull << 0) | (1ull << 3) | (2ull << 6) | (3ull << 9) |
      (4ull << 12) | (5ull << 15) | (1ull << 18) | (2ull << 21) | (3ull << 24) | (4ull << 27) |
      (5ull << 30) | (2ull << 33) | (3ull << 36) | (4ull << 39) | (5ull << 42) | (3ull << 45) |
      (4ull << 48) | (5ull << 51) | (4ull << 54) | (5ull << 57) | (5ull << 60);
  ra = (int)((RA >> (3 * e)) & 7ull);
  rb = (int)((RB >> (3 * e)) & 7ull);
}

template <int WARPS, int MAXREG, bool J2>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(MAXREG)
ukf_step_kernel(const __grid_constant__ UkfArgs a) {
  __shared__ __align__(16) TargetSmem tsm[WARPS][2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int h = lane >> 4, l = lane & 15;
  const int64_t pair = (int64_t)blockIdx.x * WARPS + wib;  // this warp's pair of targets
  const int64_t tg = 2 * pair + h;
  if (2 * pair >= a.n) return;  // whole warp idle (warp-uniform)
  const bool live = tg < a.n;
  TargetSmem* S = &tsm[wib][h];

  // ---- load: every global input of the warp is requested before anything waits -------------
  {
    const double* gp = a.est_p + tg * 36;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, xin = 0.0, tin = 0.0;
    int tk = 0;
    if (live) {
      p0 = gp[l];
      p1 = gp[l + 16];
      if (l < 4) p2 = gp[l + 32];
      if (l < 6) xin = a.est_x[(int64_t)l * a.ld + tg];
      if (l >= 8 && l < 14 && a.truth_x) tin = a.truth_x[(int64_t)(l - 8) * a.ld + tg];
      if (l == 6 && a.tasked) {
        tk = a.tasked[tg] != 0;
        if (tk && a.clear_tasked) a.tasked[tg] = 0;
      }
    }
    S->P[l] = p0;
    S->P[l + 16] = p1;
    if (l < 4) S->P[l + 32] = p2;
    if (l < 6) S->x0[l] = xin;
    if (l >= 8 && l < 14) S->tr[l - 8] = tin;
    if (l == 6) {
      S->tasked = tk;
      S->flags = 0;
      S->nis = __longlong_as_double(0x7ff8000000000000LL);
    }
  }
  const double t_now = a.clock ? a.clock->t_cur + a.dt : a.t_now;
  const uint64_t rng_step = a.clock ? a.clock->step_cur : a.rng_step;
  __syncwarp();

  // ---- factor + sigma points (generateSigmaPoints, ukf_v2.py:158-183) ---------------------
  {
    double U[21];
    const bool ok = chol_upper6(S->P, U);
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < 21; ++k) S->U[k] = U[k];
      if (live && !ok) S->flags |= PC_FLAG_NONPD_EST;
    }
  }
  __syncwarp();
  double r[3], v[3];
  bool work = false;
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
      r[c] = S->tr[c];
      v[c] = S->tr[c + 3];
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
    d[c] = (l >= 1 && l < 13) ? (work ? (c < 3 ? r[c] : v[c - 3]) : 0.0) - X0[c] : 0.0;
  }
  __syncwarp();
  if (live && l < 13) {
#pragma unroll
    for (int c = 0; c < 6; ++c) S->X[l][c] = d[c];
  }
  __syncwarp();
  if (live && l < 6) {
    double s0 = S->X[1][l] + S->X[2][l], s1 = S->X[3][l] + S->X[4][l];
#pragma unroll
    for (int i = 5; i < 13; i += 2) {
      s0 += S->X[i][l];
      s1 += S->X[i + 1][l];
    }
    const double m = fma(a.p.wi, s0 + s1, a.p.eps_w * X0[l]);
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
        double acc0 = S->X[1][ra] * S->X[1][rb], acc1 = S->X[2][ra] * S->X[2][rb];
#pragma unroll
        for (int i = 3; i < 13; i += 2) {
          acc0 = fma(S->X[i][ra], S->X[i][rb], acc0);
          acc1 = fma(S->X[i + 1][ra], S->X[i + 1][rb], acc1);
        }
        const double acc = fma(a.p.wi, acc0 + acc1, (a.p.wc0 * S->X[0][ra]) * S->X[0][rb]) +
                           a.p.Q[ra * 6 + rb];
        S->P[ra * 6 + rb] = acc;
        S->P[rb * 6 + ra] = acc;
      }
    }
  }
  __syncwarp();

  // ---- measurement / no-measurement update, one lane per target ----------------------------
  const int tasked = S->tasked;
  if (live && l == 0) {
    if (tasked) {
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
  const double* Pout = tasked ? S->PE : S->P;  // est_p of this target

  // ---- outputs ---------------------------------------------------------------------------------
  if (live) {
    // flags + scheduler summaries (agents.py:157-171; env.py:446-452,599-609)
    double chk = 0.0, dg = 0.0;
    if (l < 6) {
      dg = Pout[l * 7];
      chk = S->x0[l] + dg;
    }
    const unsigned hm = 0xffffu << (h << 4);
    const unsigned bad = __ballot_sync(hm, !isfinite(chk)) & hm;
    const float dgf = (float)dg;
    const float t1 = __shfl_down_sync(hm, dgf, 1, 16), t2 = __shfl_down_sync(hm, dgf, 2, 16);
    const float tr3 = (dgf + t1) + t2;  // lane 0: trace of the position block, lane 3: velocity block
    if (l == 0) {
      uint32_t f = S->flags;
      if (bad) f |= PC_FLAG_NONFINITE;
      if (a.out_flags) a.out_flags[tg] = f;
      if (a.out_nis) a.out_nis[tg] = S->nis;
      if (a.tr_pos) a.tr_pos[tg] = tr3;
      if (tasked) {
        if (a.num_tasked) a.num_tasked[tg] += 1;
        if (a.last_time_tasked) a.last_time_tasked[tg] = (float)t_now;
      }
    }
    if (l == 3 && a.tr_vel) a.tr_vel[tg] = tr3;
    // covariance tiles: 36 contiguous doubles per target, 16 lanes
    double* ge = a.est_p + tg * 36;
    ge[l] = Pout[l];
    ge[l + 16] = Pout[l + 16];
    if (l < 4) ge[l + 32] = Pout[l + 32];
    if (a.out_pred_p) {
      double* gq = a.out_pred_p + tg * 36;
      gq[l] = S->P[l];
      gq[l + 16] = S->P[l + 16];
      if (l < 4) gq[l + 32] = S->P[l + 32];
    }
    if (l < 6) {
      const int64_t g = (int64_t)l * a.ld + tg;
      a.est_x[g] = S->x0[l];
      if (a.truth_x) a.truth_x[g] = S->X[13][l];
      if (a.out_pred_x) a.out_pred_x[g] = S->px[l];
      if (a.out_z && tasked) a.out_z[g] = S->z[l];
    }
  }

  // ---- fused visibility: both packed maps for the next step (env.py:455-462) -----------------
  // lanes = sensors, one warp ballot = one 32-bit word of a target's row.
  if (a.sens_q) {
    if (live && l < 2) {
      const double* x = (l == 0) ? S->X[13] : S->x0;
      S->vq[l] = horizon_q(x[0], x[1], x[2], a.body_radius, a.vis_tol);
    }
    __syncwarp();
    const TargetSmem* T0 = &tsm[wib][0];
    const TargetSmem* T1 = &tsm[wib][1];
    const bool two = 2 * pair + 1 < a.n;
    const double RE2 = a.body_radius * a.body_radius;
    uint32_t keep[4] = {0u, 0u, 0u, 0u};  // word `lane` of (t0 truth, t0 est, t1 truth, t1 est)
    for (int j0 = 0, w = 0; j0 < a.M; j0 += 32, ++w) {
      const int j = j0 + lane;
      const bool on = j < a.M;
      double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
      if (on) {
        const double2* q = reinterpret_cast<const double2*>(a.sens_q + j);
        const double2 lo = __ldg(q), hi = __ldg(q + 1);
        s = make_double4(lo.x, lo.y, hi.x, hi.y);
      }
      const uint32_t b0 = __ballot_sync(0xffffffffu, on && sees(s, T0->X[13][0], T0->X[13][1], T0->X[13][2], T0->vq[0], RE2));
      const uint32_t b1 = __ballot_sync(0xffffffffu, on && sees(s, T0->x0[0], T0->x0[1], T0->x0[2], T0->vq[1], RE2));
      uint32_t b2 = 0u, b3 = 0u;
      if (two) {
        b2 = __ballot_sync(0xffffffffu, on && sees(s, T1->X[13][0], T1->X[13][1], T1->X[13][2], T1->vq[0], RE2));
        b3 = __ballot_sync(0xffffffffu, on && sees(s, T1->x0[0], T1->x0[1], T1->x0[2], T1->vq[1], RE2));
      }
      if ((w & 31) == lane) {
        keep[0] = b0; keep[1] = b1; keep[2] = b2; keep[3] = b3;
      }
      if ((w & 31) == 31 || j0 + 32 >= a.M) {  // flush up to 32 words per row, coalesced
        const int wbase = w & ~31;
        if (wbase + lane <= w) {
          const int64_t o0 = (2 * pair) * a.wpr + wbase + lane;
          a.vis_truth[o0] = keep[0];
          a.vis_est[o0] = keep[1];
          if (two) {
            a.vis_truth[o0 + a.wpr] = keep[2];
            a.vis_est[o0 + a.wpr] = keep[3];
          }
        }
      }
    }
  }

  if (a.clock && a.advance_clock && blockIdx.x == 0 && threadIdx.x == 0) {
    // nobody reads t_now / step during this launch (kernels read the t_cur / step_cur copies)
    a.clock->t_now += a.dt;
    a.clock->step += 1;
  }
}

cudaError_t launch_ukf(const UkfArgs& a, int force_model, cudaStream_t stream) {
  if (a.n <= 0) return cudaSuccess;
  // 1 warp = 2 targets per CTA, 17 CTAs per SM (120 registers): a 10 000-target catalog is
  // exactly two resident rounds on 148 SMs (148 * 17 * 2 * 2 = 10 064).
  constexpr int WARPS = 1, MINB = 120;
  const int64_t pairs = (a.n + 1) / 2;
  const unsigned blocks = (unsigned)((pairs + WARPS - 1) / WARPS);
  if (force_model == PC_FORCE_TWOBODY_J2)
    ukf_step_kernel<WARPS, MINB, true><<<blocks, WARPS * 32, 0, stream>>>(a);
  else
    ukf_step_kernel<WARPS, MINB, false><<<blocks, WARPS * 32, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pc
