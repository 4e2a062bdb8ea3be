// K4: per-target unscented Kalman filter predict + update, fused with the target's
// truth propagation.
//
// Reference: UnscentedKalmanFilter.predictAndUpdate (punchclock/estimation/ukf_v2.py:185-375)
// for the filter ezUKF builds (6-D state, identity measurement, ez_ukf.py:17-95), called once
// per target per step from Target.updateNonPhysical (common/agents.py:144-171), plus
// Agent.propagate for the truth state (agents.py:59-71).
//
// Mapping: FOUR lanes per target ("quad"), eight targets per warp, one warp per CTA, nothing
// but warp-level synchronisation.  The 13 sigma points are handled as 6 (+,-) pairs plus the
// centre point; lanes 0..2 of a quad integrate two pairs each and lane 3 integrates the centre
// point together with the target's TRUTH state.  A pair is advanced by one interleaved
// two-state DOP853 (same step size), which gives every lane two independent dependency chains
// and keeps every stage in registers.  Because x+ and x- of a pair live in the same lane the
// unscented transform needs no data exchange until the very end:
//   load     coalesced CTA-wide copy of the 8 covariance tiles (8*36 contiguous doubles) and
//            the SoA state columns into shared memory
//   factor   upper Cholesky of est_p, evaluated by all four lanes of the quad (SIMD: the eight
//            targets of the warp are factored in one pass)
//   sigma    per pair: s = x+ + x-,  D = x+ - x-   (kept in registers)
//   unscent  pred_x from sum(s) and the centre point (quad shuffle all-reduce of 6 values);
//            pred_p = wc0 m m^T + w sum_j [2 a_j a_j^T + D_j D_j^T / 2] + Q with
//            a_j = s_j/2 - pred_x, reduced over the quad (21 values)
//   update   tasked targets only: one lane per target runs the measurement update on a
//            shared-memory copy (second Cholesky, S, C, gain via triangular solves, NIS)
//   vis      lanes = sensors: each warp ballot is one 32-bit word of a packed vis-map row
//   store    coalesced copies of the est_p / pred_p tiles, SoA columns for the states
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

constexpr int kTPW = 8;  // targets per warp

struct TargetSmem {   // scratch of the measurement update (tasked targets)
  double D[6][6];     // D[j][a] = (X_{1+j} - X_{7+j})[a]: difference of the propagated +/- pair j
  double U[21];       // packed upper Cholesky factor (row-major, j >= i)
  double px[6];       // pred_x
  double m[6];        // pred_x - X_0
  double z[6];
  double xo[6];       // est_x out
  double nis;
  uint32_t flags;
  int pad;
  double* P;          // pred_p tile (36)
  double* PE;         // est_p tile out (36)
};

struct WarpSmem {
  double P[kTPW][36];    // in: est_p tiles; out: pred_p tiles       (contiguous: coalesced copies)
  double PE[kTPW][36];   // out: est_p tiles
  double xs[kTPW][6];    // est_x in / out
  double tr[kTPW][6];    // truth in / out
  double4 tv[kTPW][2];   // (truth', est_x') position + horizon term for the fused vis test
  TargetSmem t[kTPW];
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
    const double d = rsqrt_fast(s);
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
    for (int j = 0; j < 6; ++j) D[j] = S->D[j][a];
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
    S->xo[a] = xe;  // est_x = pred_x + K nu  (ukf_v2.py:373-375)
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

__device__ __forceinline__ double quad_sum(double x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}

// One (+,-) sigma pair: x +- col, both members advanced in lock-step; returns sum and
// difference of the propagated members.
template <bool J2>
__device__ __forceinline__ void propagate_pair(const double (&x)[6], const double (&col)[6], double dt,
                                               int ns, double (&sum)[6], double (&dif)[6]) {
  double r[2][3], v[2][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r[0][c] = x[c] + col[c];
    r[1][c] = x[c] - col[c];
    v[0][c] = x[c + 3] + col[c + 3];
    v[1][c] = x[c + 3] - col[c + 3];
  }
  propagate_states<2, J2>(r, v, dt, ns);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    sum[c] = r[0][c] + r[1][c];
    dif[c] = r[0][c] - r[1][c];
    sum[c + 3] = v[0][c] + v[1][c];
    dif[c + 3] = v[0][c] - v[1][c];
  }
}

template <int MAXREG, bool J2>
__global__ void __launch_bounds__(32) __maxnreg__(MAXREG)
ukf_step_kernel(const __grid_constant__ UkfArgs a) {
  __shared__ __align__(16) WarpSmem W;
  const int lane = threadIdx.x;
  const int tl = lane >> 2, q = lane & 3;       // target within the warp, lane within the quad
  const int64_t first = (int64_t)blockIdx.x * kTPW;
  const int64_t tg = first + tl;
  const int nlive = (int)min((int64_t)kTPW, a.n - first);
  const bool live = tl < nlive;
  const bool have_truth = a.truth_x != nullptr;

  // ---- load: coalesced, everything requested before anything waits ---------------------------
  {
    const double* gp = a.est_p + first * 36;
    double* sp = &W.P[0][0];
    double pv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) pv[i] = (lane + 32 * i < nlive * 36) ? gp[lane + 32 * i] : 0.0;
    double xv[2] = {0.0, 0.0}, tv[2] = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = lane + 32 * i, c = k >> 3, t = k & 7;
      if (k < 48 && t < nlive) {
        xv[i] = a.est_x[(int64_t)c * a.ld + first + t];
        if (have_truth) tv[i] = a.truth_x[(int64_t)c * a.ld + first + t];
      }
    }
    int tk = 0;
    if (live && q == 0 && a.tasked) {
      tk = a.tasked[tg] != 0;
      if (tk && a.clear_tasked) a.tasked[tg] = 0;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) sp[lane + 32 * i] = pv[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = lane + 32 * i, c = k >> 3, t = k & 7;
      if (k < 48) {
        W.xs[t][c] = xv[i];
        W.tr[t][c] = tv[i];
      }
    }
    if (q == 0) {
      W.t[tl].flags = (uint32_t)tk << 31;  // bit 31: tasked (internal)
      W.t[tl].nis = __longlong_as_double(0x7ff8000000000000LL);
      W.t[tl].P = W.P[tl];
      W.t[tl].PE = W.PE[tl];
    }
  }
  const double t_now = a.clock ? a.clock->t_cur + a.dt : a.t_now;
  const uint64_t rng_step = a.clock ? a.clock->step_cur : a.rng_step;
  __syncwarp();

  // ---- factor (generateSigmaPoints, ukf_v2.py:158-183) ---------------------------------------
  double x[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) x[c] = W.xs[tl][c];
  double col0[6], col1[6];  // gamma * columns (2q) and (2q+1) of the upper factor
  bool chol_ok;
  {
    double U[21];
    chol_ok = chol_upper6(W.P[tl], U);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      // column j has entries U[c][j] for c <= j
      const double u0 = (c <= 0) ? U[uidx(c, c <= 0 ? 0 : c)] : 0.0;
      const double u1 = (c <= 1) ? U[uidx(c, c <= 1 ? 1 : c)] : 0.0;
      const double u2 = (c <= 2) ? U[uidx(c, c <= 2 ? 2 : c)] : 0.0;
      const double u3 = (c <= 3) ? U[uidx(c, c <= 3 ? 3 : c)] : 0.0;
      const double u4 = (c <= 4) ? U[uidx(c, c <= 4 ? 4 : c)] : 0.0;
      const double u5 = U[uidx(c, 5)];
      col0[c] = a.p.gamma * (q == 0 ? u0 : (q == 1 ? u2 : u4));
      col1[c] = a.p.gamma * (q == 0 ? u1 : (q == 1 ? u3 : u5));
    }
  }

  // ---- sigma points --------------------------------------------------------------------------
  int ns = a.substeps;
  bool clamp = false;
  {
    const double r0[3] = {x[0], x[1], x[2]}, v0[3] = {x[3], x[4], x[5]};
    if (ns <= 0) ns = auto_substeps(r0, v0, a.dt, &clamp);  // identical in the four lanes
  }
  double s0[6], d0[6], s1[6], d1[6];  // lanes 0..2: two pairs; lane 3: s0 = centre, s1 = truth
  if (q < 3) {
    propagate_pair<J2>(x, col0, a.dt, ns, s0, d0);
    propagate_pair<J2>(x, col1, a.dt, ns, s1, d1);
  } else {
    double tr[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) tr[c] = have_truth ? W.tr[tl][c] : x[c];
    const double rt[3] = {tr[0], tr[1], tr[2]}, vt[3] = {tr[3], tr[4], tr[5]};
    const int nt = a.substeps > 0 ? a.substeps : auto_substeps(rt, vt, a.dt, nullptr);
    if (nt == ns) {
      double r[2][3] = {{x[0], x[1], x[2]}, {tr[0], tr[1], tr[2]}};
      double v[2][3] = {{x[3], x[4], x[5]}, {tr[3], tr[4], tr[5]}};
      propagate_states<2, J2>(r, v, a.dt, ns);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s0[c] = r[0][c]; s0[c + 3] = v[0][c];
        s1[c] = r[1][c]; s1[c + 3] = v[1][c];
      }
    } else {  // centre and truth want different step counts: integrate them one after the other
      double r[3] = {x[0], x[1], x[2]}, v[3] = {x[3], x[4], x[5]};
      propagate_state<J2>(r, v, a.dt, ns);
      double r2[3] = {tr[0], tr[1], tr[2]}, v2[3] = {tr[3], tr[4], tr[5]};
      propagate_state<J2>(r2, v2, a.dt, nt);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s0[c] = r[c]; s0[c + 3] = v[c];
        s1[c] = r2[c]; s1[c + 3] = v2[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      d0[c] = 0.0;
      d1[c] = 0.0;
    }
  }
  __syncwarp();

  // ---- unscented transform (predictStateEstimate / predictCovariance, ukf_v2.py:261-292) ------
  // X_0 from lane 3; sum over the 12 outer points of (X_i - X_0) = sum_j s_j - 12 X_0
  double X0[6], m[6], px[6];
  const int l3 = lane | 3;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    X0[c] = __shfl_sync(0xffffffffu, s0[c], l3);
    // each pair contributes (s_j - 2 X_0): differences first, then the (small) sum
    const double part = (q < 3) ? (s0[c] - 2.0 * X0[c]) + (s1[c] - 2.0 * X0[c]) : 0.0;
    m[c] = fma(a.p.wi, quad_sum(part), a.p.eps_w * X0[c]);
    px[c] = X0[c] + m[c];  // pred_x = sigma_points . mean_weight
  }
  // pred_p = Xres Wc Xres^T + Q; for a pair  Xres+ Xres+^T + Xres- Xres-^T = 2 a a^T + D D^T / 2
  double E[21];
  {
    double a0[6], a1[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      a0[c] = 0.5 * (s0[c] - 2.0 * X0[c]) - m[c];
      a1[c] = 0.5 * (s1[c] - 2.0 * X0[c]) - m[c];
    }
#pragma unroll
    for (int ra = 0; ra < 6; ++ra)
#pragma unroll
      for (int rb = ra; rb < 6; ++rb) {
        double e;
        if (q < 3) {
          e = fma(0.5 * d0[ra], d0[rb], 2.0 * a0[ra] * a0[rb]);
          e = fma(0.5 * d1[ra], d1[rb], fma(2.0 * a1[ra], a1[rb], e));
          e *= a.p.wi;
        } else {
          e = fma(a.p.wc0 * m[ra], m[rb], a.p.Q[ra * 6 + rb]);
        }
        E[uidx(ra, rb)] = quad_sum(e);
      }
  }

  // ---- no-measurement result (update(None), ukf_v2.py:238-244) ----------------------------------
  const unsigned fl = W.t[tl].flags;
  const bool tasked = (fl >> 31) != 0;
  const unsigned any_tasked = __ballot_sync(0xffffffffu, live && tasked);
  if (q == 0) {
    // pred_p tile (and est_p = pred_p for untasked targets)
#pragma unroll
    for (int ra = 0; ra < 6; ++ra)
#pragma unroll
      for (int rb = 0; rb < 6; ++rb) {
        const double e = E[ra <= rb ? uidx(ra, rb) : uidx(rb, ra)];
        W.P[tl][ra * 6 + rb] = e;
        W.PE[tl][ra * 6 + rb] = e;
      }
  }
  if (q == 3) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      W.xs[tl][c] = X0[c];   // est_x = propagated centre point
      W.tr[tl][c] = s1[c];   // truth'
    }
  }

  // ---- measurement update (tasked targets only) ------------------------------------------------
  if (any_tasked) {
    TargetSmem* S = &W.t[tl];
    if (tasked) {
      if (q < 3) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          S->D[2 * q][c] = d0[c];
          S->D[2 * q + 1][c] = d1[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          S->px[c] = px[c];
          S->m[c] = m[c];
        }
      }
    }
    __syncwarp();
    if (live && tasked && q == 0) {
      if (a.z) {
#pragma unroll
        for (int c = 0; c < 6; ++c) S->z[c] = a.z[(int64_t)c * a.ld + tg];
      } else {
        // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
        double nrm[6];
        normal6(a.rng_seed, rng_step, (uint64_t)tg, nrm);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double acc = W.tr[tl][c];
#pragma unroll
          for (int k = 0; k <= c; ++k) acc = fma(a.p.Lr[c * 6 + k], nrm[k], acc);
          S->z[c] = acc;
        }
      }
      measurement_update(S, a.p);
#pragma unroll
      for (int c = 0; c < 6; ++c) W.xs[tl][c] = S->xo[c];
    }
  }
  __syncwarp();

  // ---- flags + scheduler summaries (agents.py:157-171; env.py:446-452,599-609) -------------------
  if (live && q == 0) {
    const double* Pout = W.PE[tl];
    double chk = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) chk += W.xs[tl][c] + Pout[c * 7];
    uint32_t f = W.t[tl].flags & 0x7fffffffu;
    if (!chol_ok) f |= PC_FLAG_NONPD_EST;
    if (clamp) f |= PC_FLAG_SUBSTEP_CLAMP;
    if (!isfinite(chk)) f |= PC_FLAG_NONFINITE;
    if (a.out_flags) a.out_flags[tg] = f;
    if (a.out_nis) a.out_nis[tg] = W.t[tl].nis;
    if (a.tr_pos) a.tr_pos[tg] = ((float)Pout[0] + (float)Pout[7]) + (float)Pout[14];
    if (a.tr_vel) a.tr_vel[tg] = ((float)Pout[21] + (float)Pout[28]) + (float)Pout[35];
    if (tasked) {
      if (a.num_tasked) a.num_tasked[tg] += 1;
      if (a.last_time_tasked) a.last_time_tasked[tg] = (float)t_now;
    }
  }

  // ---- store: coalesced tiles + SoA columns -------------------------------------------------------
  {
    double* ge = a.est_p + first * 36;
    const double* se = &W.PE[0][0];
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (lane + 32 * i < nlive * 36) ge[lane + 32 * i] = se[lane + 32 * i];
    if (a.out_pred_p) {
      double* gq = a.out_pred_p + first * 36;
      const double* sq = &W.P[0][0];
#pragma unroll
      for (int i = 0; i < 9; ++i)
        if (lane + 32 * i < nlive * 36) gq[lane + 32 * i] = sq[lane + 32 * i];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = lane + 32 * i, c = k >> 3, t = k & 7;
      if (k < 48 && t < nlive) {
        const int64_t g = (int64_t)c * a.ld + first + t;
        a.est_x[g] = W.xs[t][c];
        if (have_truth) a.truth_x[g] = W.tr[t][c];
        const bool tk = (W.t[t].flags >> 31) != 0;
        if (a.out_z && tk) a.out_z[g] = W.t[t].z[c];
      }
    }
    if (a.out_pred_x && live && q == 1) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.out_pred_x[(int64_t)c * a.ld + tg] = px[c];
    }
  }

  // ---- fused visibility: both packed maps for the next step (env.py:455-462) ---------------------
  // lanes = sensors, one warp ballot = one 32-bit word of a target's row.
  if (a.sens_q) {
    if (live && q < 2) {
      const double* xx = (q == 0) ? W.tr[tl] : W.xs[tl];
      W.tv[tl][q] = make_double4(xx[0], xx[1], xx[2], horizon_q(xx[0], xx[1], xx[2], a.body_radius, a.vis_tol));
    }
    __syncwarp();
    const double RE2 = a.body_radius * a.body_radius;
    for (int w0 = 0; w0 < a.wpr; w0 += 32) {      // 32 words (1024 sensors) per pass
      uint32_t keep[2 * kTPW];
#pragma unroll
      for (int k = 0; k < 2 * kTPW; ++k) keep[k] = 0u;
      const int wend = min(a.wpr, w0 + 32);
      for (int w = w0; w < wend; ++w) {
        const int j = w * 32 + lane;
        const bool on = j < a.M;
        double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
        if (on) {
          const double2* qq = reinterpret_cast<const double2*>(a.sens_q + j);
          const double2 lo = __ldg(qq), hi = __ldg(qq + 1);
          s = make_double4(lo.x, lo.y, hi.x, hi.y);
        }
#pragma unroll
        for (int k = 0; k < 2 * kTPW; ++k) {
          const double4 t = W.tv[k >> 1][k & 1];
          const uint32_t b = __ballot_sync(0xffffffffu, on && sees(s, t.x, t.y, t.z, t.w, RE2));
          if ((w - w0) == lane) keep[k] = b;
        }
      }
      if (w0 + lane < wend) {
#pragma unroll
        for (int k = 0; k < 2 * kTPW; ++k) {
          if ((k >> 1) < nlive) {
            uint32_t* dst = (k & 1) ? a.vis_est : a.vis_truth;
            dst[(first + (k >> 1)) * a.wpr + w0 + lane] = keep[k];
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
  // one warp (8 targets) per CTA; 224 registers -> 9 warps resident per SM, each lane running
  // two interleaved integrations.
  constexpr int MAXREG = 224;
  const unsigned blocks = (unsigned)((a.n + kTPW - 1) / kTPW);
  if (force_model == PC_FORCE_TWOBODY_J2)
    ukf_step_kernel<MAXREG, true><<<blocks, 32, 0, stream>>>(a);
  else
    ukf_step_kernel<MAXREG, false><<<blocks, 32, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pc
