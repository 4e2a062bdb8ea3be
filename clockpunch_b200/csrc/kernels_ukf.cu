// K4: per-target unscented Kalman filter predict + update, fused with the target's
// truth propagation.
//
// Reference: UnscentedKalmanFilter.predictAndUpdate (punchclock/estimation/ukf_v2.py:185-375)
// for the filter ezUKF builds (6-D state, identity measurement, ez_ukf.py:17-95), called once
// per target per step from Target.updateNonPhysical (common/agents.py:144-171), plus
// Agent.propagate for the truth state (agents.py:59-71).
//
// Mapping (one CTA = T targets, 16 lanes per target, two targets per warp):
//   phase 0  CTA-wide coalesced load of the est_p tile (T*36 contiguous doubles) and the
//            SoA state rows into shared memory.
//   phase 1  one THREAD per target: upper Cholesky of est_p (serial 6x6 algebra is packed
//            into the lanes of one warp instead of being replicated 16x) + substep count.
//   phase 2  one LANE per sigma point (lanes 0..12) + the truth state (lane 13): each lane
//            integrates its own state with every DOP853 stage in registers.
//   phase 3  unscented transform through shared memory: mean from differences to the centre
//            point, residuals, and the 21 unique entries of pred_p spread over the 16 lanes.
//   phase 4  one THREAD per tasked target: measurement update (second Cholesky, S, C, gain
//            via triangular solves, est_x, est_p, NIS).
//   phase 5  CTA-wide coalesced stores.
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
  pc_ukf_params p;
};

struct TargetSmem {
  double P[36];     // in: est_p; after phase 3: pred_p
  double PE[36];    // out: est_p
  double U[21];     // packed upper Cholesky factor (row-major, j >= i)
  double X[14][6];  // propagated sigma points (0..12) and truth (13); 0..12 become residuals
  double x0[6];     // in: est_x; out: est_x
  double xc[6];     // propagated centre point X_0
  double px[6];     // pred_x
  double m[6];      // pred_x - X_0
  double z[6];
  double nis;
  int nsub;
  int tasked;
  uint32_t flags;
  int pad;
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

template <int T, bool J2>
__global__ void __launch_bounds__(T * 16, 512 / (T * 16))
ukf_step_kernel(const __grid_constant__ UkfArgs a) {
  __shared__ TargetSmem sm[T];
  constexpr int NT = T * 16;
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * T;
  const int nact = (int)min((int64_t)T, a.n - base);  // live targets in this CTA

  // ---- phase 0: coalesced loads -------------------------------------------------
  {
    const double* gp = a.est_p + base * 36;
    for (int k = tid; k < nact * 36; k += NT) sm[k / 36].P[k % 36] = gp[k];
    for (int k = tid; k < 6 * T; k += NT) {
      const int c = k / T, t = k % T;
      if (t < nact) sm[t].x0[c] = a.est_x[(int64_t)c * a.ld + base + t];
    }
    if (tid < T) {
      sm[tid].flags = 0;
      sm[tid].nis = __longlong_as_double(0x7ff8000000000000LL);
      sm[tid].tasked = (tid < nact && a.tasked) ? (a.tasked[base + tid] != 0) : 0;
    }
  }
  __syncthreads();

  // ---- phase 1: one thread per target: Cholesky of est_p, substep count ----------
  if (tid < nact) {
    TargetSmem* S = &sm[tid];
    if (!chol_upper6(S->P, S->U)) S->flags |= PC_FLAG_NONPD_EST;
    int ns = a.substeps;
    if (ns <= 0) {
      const double r[3] = {S->x0[0], S->x0[1], S->x0[2]};
      const double v[3] = {S->x0[3], S->x0[4], S->x0[5]};
      bool cl;
      ns = auto_substeps(r, v, a.dt, &cl);
      if (cl) S->flags |= PC_FLAG_SUBSTEP_CLAMP;
    }
    S->nsub = ns;
  }
  __syncthreads();

  // ---- phase 2: one lane per sigma point / truth state ---------------------------
  const int t = tid >> 4, l = tid & 15;
  const bool live = t < nact;
  TargetSmem* S = &sm[live ? t : 0];
  double X0[6];  // propagated centre point (for the transform below)
  {
    double r[3], v[3];
    bool work = false;
    int ns = 1;
    if (live && l < 13) {
      // generateSigmaPoints (ukf_v2.py:158-183): x + gamma * [0, U, -U], COLUMNS of upper U
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
      ns = S->nsub;
      work = true;
    } else if (live && l == 13 && a.truth_x) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        r[c] = a.truth_x[(int64_t)c * a.ld + base + t];
        v[c] = a.truth_x[(int64_t)(c + 3) * a.ld + base + t];
      }
      ns = a.substeps > 0 ? a.substeps : auto_substeps(r, v, a.dt, nullptr);
      work = true;
    }
    if (work) {
      propagate_state<J2>(r, v, a.dt, ns);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        S->X[l][c] = r[c];
        S->X[l][c + 3] = v[c];
      }
    }
  }
  __syncwarp();

  // ---- phase 3: unscented transform (predictStateEstimate / predictCovariance,
  //      ukf_v2.py:261-292) -----------------------------------------------------------
  double d[6];
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
        int ra = 0, e0 = e;
        while (e0 >= 6 - ra) { e0 -= 6 - ra; ++ra; }
        const int rb = ra + e0;
        double acc = a.p.wc0 * S->X[0][ra] * S->X[0][rb];
#pragma unroll
        for (int i = 1; i < 13; ++i) acc = fma(a.p.wi * S->X[i][ra], S->X[i][rb], acc);
        acc += a.p.Q[ra * 6 + rb];
        S->P[ra * 6 + rb] = acc;
        S->P[rb * 6 + ra] = acc;
      }
    }
  }
  __syncthreads();

  // ---- phase 4: measurement / no-measurement update, one thread per target ----------
  if (tid < nact) {
    TargetSmem* Q = &sm[tid];
    if (Q->tasked) {
      if (a.z) {
#pragma unroll
        for (int c = 0; c < 6; ++c) Q->z[c] = a.z[(int64_t)c * a.ld + base + tid];
      } else {
        // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
        double nrm[6];
        normal6(a.rng_seed, a.rng_step, (uint64_t)(base + tid), nrm);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double acc = Q->X[13][c];
#pragma unroll
          for (int k = 0; k <= c; ++k) acc = fma(a.p.Lr[c * 6 + k], nrm[k], acc);
          Q->z[c] = acc;
        }
      }
      measurement_update(Q, a.p);
    } else {
      // update(None) (ukf_v2.py:238-244): est_x = propagated centre point, est_p = pred_p
#pragma unroll
      for (int c = 0; c < 6; ++c) Q->x0[c] = Q->xc[c];
#pragma unroll
      for (int k = 0; k < 36; ++k) Q->PE[k] = Q->P[k];
    }
    // flags + scheduler summaries (agents.py:157-171; env.py:446-452,599-609)
    bool fin = true;
#pragma unroll
    for (int c = 0; c < 6; ++c) fin = fin && isfinite(Q->x0[c]);
#pragma unroll
    for (int k = 0; k < 36; k += 7) fin = fin && isfinite(Q->PE[k]);
    if (!fin) Q->flags |= PC_FLAG_NONFINITE;
    const int64_t gi = base + tid;
    if (a.out_flags) a.out_flags[gi] = Q->flags;
    if (a.out_nis) a.out_nis[gi] = Q->nis;
    if (a.tr_pos) a.tr_pos[gi] = (float)Q->PE[0] + (float)Q->PE[7] + (float)Q->PE[14];
    if (a.tr_vel) a.tr_vel[gi] = (float)Q->PE[21] + (float)Q->PE[28] + (float)Q->PE[35];
    if (Q->tasked) {
      if (a.num_tasked) a.num_tasked[gi] += 1;
      if (a.last_time_tasked) a.last_time_tasked[gi] = (float)a.t_now;
    }
  }
  __syncthreads();

  // ---- phase 5: coalesced stores -----------------------------------------------------
  {
    double* ge = a.est_p + base * 36;
    for (int k = tid; k < nact * 36; k += NT) ge[k] = sm[k / 36].PE[k % 36];
    if (a.out_pred_p) {
      double* gq = a.out_pred_p + base * 36;
      for (int k = tid; k < nact * 36; k += NT) gq[k] = sm[k / 36].P[k % 36];
    }
    for (int k = tid; k < 6 * T; k += NT) {
      const int c = k / T, tt = k % T;
      if (tt < nact) {
        const int64_t g = (int64_t)c * a.ld + base + tt;
        a.est_x[g] = sm[tt].x0[c];
        if (a.truth_x) a.truth_x[g] = sm[tt].X[13][c];
        if (a.out_pred_x) a.out_pred_x[g] = sm[tt].px[c];
        if (a.out_z && sm[tt].tasked) a.out_z[g] = sm[tt].z[c];
      }
    }
  }
}

cudaError_t launch_ukf(double* est_x, double* est_p, double* truth_x, const double* z,
                       const uint8_t* tasked, int64_t n, int64_t ld, double t0, double tf,
                       int substeps, int force_model, const pc_ukf_params* p, uint64_t seed,
                       uint64_t step, double* pred_x, double* pred_p, double* nis, uint32_t* flags,
                       double* out_z, int64_t* num_tasked, float* last_time_tasked, float* tr_pos,
                       float* tr_vel, cudaStream_t stream) {
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
  a.p = *p;
  constexpr int T = 8;
  const unsigned blocks = (unsigned)((n + T - 1) / T);
  if (force_model == PC_FORCE_TWOBODY_J2)
    ukf_step_kernel<T, true><<<blocks, T * 16, 0, stream>>>(a);
  else
    ukf_step_kernel<T, false><<<blocks, T * 16, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pc
