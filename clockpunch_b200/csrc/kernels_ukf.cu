// K4: per-target unscented Kalman filter predict + update, fused with the target's
// truth propagation.
//
// Reference: UnscentedKalmanFilter.predictAndUpdate (punchclock/estimation/ukf_v2.py:185-375)
// for the filter ezUKF builds (6-D state, identity measurement, ez_ukf.py:17-95), called once
// per target per step from Target.updateNonPhysical (common/agents.py:144-171), plus
// Agent.propagate for the truth state (agents.py:59-71).
//
// Mapping -- a staged pipeline, each stage with the thread mapping that suits it (measured:
// profiles/README.md; single fused kernels with 16 or 4 lanes per target spent 2/3 of their
// issue slots on cross-lane scaffolding and thrashed the instruction cache):
//   sigma_gen_kernel      one THREAD per target: upper Cholesky of est_p, 13 sigma points written
//                         to the sigma buffer SB (14 blocks of (6, ld) SoA: block 0 = centre =
//                         est_x, 1..6 = +gamma U[:, j], 7..12 = -gamma U[:, j], 13 = truth state)
//   propagate_sigma_kernel (kernels_propagate.cu) one THREAD per state: all 14 N states advanced
//                         in place, every DOP853 stage in registers, fully coalesced
//   ukf_finish_kernel     one THREAD per target: unscented transform streamed over the 6 (+,-)
//                         pairs, measurement update for tasked targets, outputs, and -- so that
//                         the next step starts at the propagation stage -- the Cholesky factor and
//                         sigma points of the NEW estimate (gen_next); fused with the target's row
//                         of both visibility maps
//   ukf_finish_quad_kernel  the same stage for catalogs up to 32 k targets, where one thread per
//                         target cannot fill the machine: FOUR lanes per target, and (with
//                         pc_step_args.task_of) one extra CTA per sensor whose first warp does the
//                         whole stage of the target that sensor tasks (coop_part1/2/3) while the
//                         other three pre-run that code for the instruction caches
// In steady state a step is therefore propagate -> finish; est_x IS block 0 and the truth state IS
// block 13 of SB, so nothing is copied.
//
// Numerics: the reference forms pred_x = sum_i w_i X_i with w_0 = -2e6, w_i = +1.67e5, which
// loses ~1e-6 km to cancellation.  Here the same quantity is evaluated as
//   X_0 + [w_i sum_i (X_i - X_0) + (sum_i w_i - 1) X_0]
// which is algebraically identical (the weights are the reference's own float64 values and
// eps_w = sum w - 1 is computed exactly from them) but carries no cancellation noise.  The
// resampled measurement sigma points of forecast() (ukf_v2.py:211-228) are symmetric about
// pred_x, so their weighted sums are expanded analytically in the same exact-weights form
// (derivation in DESIGN.md "UKF update").  With x+ and x- of pair j in hand,
//   Xres+ Xres+^T + Xres- Xres-^T = 2 a_j a_j^T + D_j D_j^T / 2,
//   a_j = (x+ + x-)/2 - pred_x,  D_j = x+ - x-,
// which is how pred_p is accumulated (again algebraically identical to ukf_v2.py:286-292).
#include <cstdlib>
#include "propagate.cuh"
#include "ukf_args.cuh"
#include "vis.cuh"

#ifdef PC_STAMPS_HEADER   // variant builds only (tools/build_variant.sh); the product library has no stamps
#define PC_STAMPS_UKF
#include PC_STAMPS_HEADER
#endif
#ifndef PC_PHASE_CLOCK
#define PC_PHASE_CLOCK(i)
#endif
#ifndef PC_COOP_CLOCK
#define PC_COOP_CLOCK(i)
#endif

namespace pc {

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

__device__ bool repair_chol6(const double* __restrict__ Ein, double* __restrict__ Uout);  // below

// Call wrapper that keeps the caller's register arrays out of local memory: only the two
// temporaries below have their address taken.
__device__ __forceinline__ void repair_chol6_regs(const double (&E)[21], double (&U)[21]) {
  double Ein[21], Uo[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    Ein[k] = E[k];
    Uo[k] = U[k];
  }
  repair_chol6(Ein, Uo);
#pragma unroll
  for (int k = 0; k < 21; ++k) U[k] = Uo[k];
}


// Upper Cholesky of a symmetric 6x6 given by its packed upper triangle E (uidx order).
__device__ __forceinline__ bool chol_upper6_packed(const double (&E)[21], double (&U)[21]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = E[uidx(i, i)];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-U[uidx(k, i)], U[uidx(k, i)], s);
    ok = ok && (s > 0.0);
    const double d = rsqrt_fast(s);
    U[uidx(i, i)] = s * d;
#pragma unroll
    for (int j = i + 1; j < 6; ++j) {
      double a = E[uidx(i, j)];
#pragma unroll
      for (int k = 0; k < i; ++k) a = fma(-U[uidx(k, i)], U[uidx(k, j)], a);
      U[uidx(i, j)] = a * d;
    }
  }
  return ok;
}

// Measurement update for one tasked target: forecast / calculateKalmanGain / updateCovariance /
// calculateInnovations / updateStateEstimate (ukf_v2.py:211-228, 294-375), every operand in registers
// (inlined).  Used by the four-lane kernel and by the thread-per-target kernels alike: an out-of-line
// version with its scratch in thread-local memory (round 1) kept the thread-per-target kernels at 168
// registers either way and cost 33 us per step at 1024 environments (8 % of the targets measured: nearly
// every warp holds one) and 16 us at 100 000 targets all measured (profiles/README.md).
// Innovation covariance S = Yres Wc Yres^T + R (ukf_v2.py:326-329); y = Us^-T nu, nis = |y|^2 (:344-346);
// cross covariance C = Xres Wc Yres^T (:330-332), G = C Us^-1, K = G Us^-T, so K S K^T = G G^T and
// K nu = G y; est_x = pred_x + K nu (:373-375), est_p = pred_p - K S K^T (:367-371).
__device__ __forceinline__ uint32_t measurement_update_reg(
    const double* __restrict__ sb, int64_t ld, int64_t blk, int64_t t, const double (&px)[6],
    const double (&m)[6], const double (&E)[21], const double (&z)[6], const pc_ukf_params& p,
    double (&xe)[6], double (&PE)[21], double& nis_out) {
  uint32_t flags = 0;
  double U2[21];
  if (!chol_upper6_packed(E, U2)) {
    flags |= PC_FLAG_NONPD_PRED;
    repair_chol6_regs(E, U2);
  }
  const double c3 = p.wc0 - p.wm0;
  const double cuu = 2.0 * p.wi * p.gamma * p.gamma;
  const double cdd = p.wc0 + 12.0 * p.wi;
  const double wg = p.wi * p.gamma;
  double delta[6], nu[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    delta[c] = p.eps_w * px[c];
    nu[c] = z[c] - (px[c] + delta[c]);
  }
  double Sm[21];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int j = b; j < 6; ++j) acc = fma(U2[uidx(a, j)], U2[uidx(b, j)], acc);
      Sm[uidx(a, b)] = fma(cuu, acc, fma(cdd * delta[a], delta[b], p.R[a * 6 + b]));
    }
  double Us[21];
  if (!chol_upper6_packed(Sm, Us)) flags |= PC_FLAG_NONPD_INNOV;
  double dinv[6];
#pragma unroll
  for (int b = 0; b < 6; ++b) dinv[b] = 1.0 / Us[uidx(b, b)];
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
  nis_out = nis;
  double G[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double D[6];
#pragma unroll
    for (int j = 0; j < 6; ++j)
      D[j] = sb[(1 + j) * blk + (int64_t)a * ld + t] - sb[(7 + j) * blk + (int64_t)a * ld + t];
    const double sa = delta[a] + c3 * m[a];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int j = b; j < 6; ++j) acc = fma(D[j], U2[uidx(b, j)], acc);
      double cab = fma(wg, acc, sa * delta[b]);
#pragma unroll
      for (int k = 0; k < b; ++k) cab = fma(-G[a][k], Us[uidx(k, b)], cab);
      G[a][b] = cab * dinv[b];
    }
    double x = px[a];
#pragma unroll
    for (int b = 0; b < 6; ++b) x = fma(G[a][b], y[b], x);
    xe[a] = x;
  }
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double acc = E[uidx(a, b)];
#pragma unroll
      for (int k = 0; k < 6; ++k) acc = fma(-G[a][k], G[b][k], acc);
      PE[uidx(a, b)] = acc;
    }
  return flags;
}

// Nearest-positive-definite repair of a symmetric 6x6 whose Cholesky factorisation failed
// (findNearestPositiveDefiniteMatrix / nearestPD, estimation/ukf_utils.py:155-198,242-256, called
// from generateSigmaPoints on LinAlgError, ukf_v2.py:170-176).  The reference takes the symmetric
// polar factor H of B = (A + A^T)/2 from an SVD, forms (B + H)/2 -- for symmetric B that is
// Q max(Lambda, 0) Q^T, the projection onto the PSD cone -- and then adds multiples of
// spacing(|A|_F) to the diagonal until Cholesky succeeds.  Here: cyclic Jacobi eigen-decomposition,
// clamp, reconstruct, and the same kind of diagonal shift (k^2 spacing at attempt k).  The shift
// sequence is not bit-reproducible between LAPACK and this code; the resulting factors agree to
// ~1e-7 in the repaired (null) directions, i.e. 1e-10 km in the sigma points.  Rare path: noinline,
// arrays in local memory.
__device__ __noinline__ bool repair_chol6(const double* __restrict__ Ein, double* __restrict__ Uout) {
  double A[6][6], V[6][6];
  double fro = 0.0;
#pragma unroll 1
  for (int i = 0; i < 6; ++i)
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
      A[i][j] = Ein[i <= j ? uidx(i, j) : uidx(j, i)];
      V[i][j] = i == j ? 1.0 : 0.0;
      fro = fma(A[i][j], A[i][j], fro);
    }
  fro = sqrt(fro);
  if (!isfinite(fro)) return false;
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll 1
    for (int p = 0; p < 5; ++p)
#pragma unroll 1
      for (int q = p + 1; q < 6; ++q) off = fma(A[p][q], A[p][q], off);
    if (off <= 1e-32 * fro * fro) break;
#pragma unroll 1
    for (int p = 0; p < 5; ++p)
#pragma unroll 1
      for (int q = p + 1; q < 6; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
        const double c = 1.0 / sqrt(fma(t, t, 1.0)), sn = t * c;
#pragma unroll 1
        for (int k = 0; k < 6; ++k) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - sn * akq;
          A[k][q] = sn * akp + c * akq;
        }
#pragma unroll 1
        for (int k = 0; k < 6; ++k) {  // A <- J^T A ; V <- V J
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - sn * aqk;
          A[q][k] = sn * apk + c * aqk;
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  double lam[6];
#pragma unroll 1
  for (int i = 0; i < 6; ++i) lam[i] = A[i][i] > 0.0 ? A[i][i] : 0.0;
  double P[21];
#pragma unroll 1
  for (int i = 0; i < 6; ++i)
#pragma unroll 1
    for (int j = i; j < 6; ++j) {
      double acc = 0.0;
#pragma unroll 1
      for (int k = 0; k < 6; ++k) acc = fma(V[i][k] * lam[k], V[j][k], acc);
      P[uidx(i, j)] = acc;
    }
  const double spacing = __longlong_as_double(__double_as_longlong(fro) + 1) - fro;
  double shift = 0.0;
#pragma unroll 1
  for (int k = 0; k <= 64; ++k) {
    double Es[21], U[21];
#pragma unroll
    for (int e = 0; e < 21; ++e) Es[e] = P[e];
#pragma unroll
    for (int i = 0; i < 6; ++i) Es[uidx(i, i)] += shift;
    bool ok = chol_upper6_packed(Es, U);
#pragma unroll
    for (int e = 0; e < 21; ++e) ok = ok && isfinite(U[e]);
    if (ok) {
#pragma unroll
      for (int e = 0; e < 21; ++e) Uout[e] = U[e];
      return true;
    }
    shift += spacing * (double)((k + 1) * (k + 1));
  }
  return false;
}

// generateSigmaPoints (ukf_v2.py:158-183): x + gamma * [0, U, -U], COLUMNS of the upper factor.
__device__ __forceinline__ void write_sigma_blocks(double* __restrict__ sb, int64_t ld, int64_t t,
                                                   const double (&x)[6], const double (&U)[21],
                                                   double gamma) {
  const int64_t blk = 6 * ld;
#pragma unroll
  for (int c = 0; c < 6; ++c) sb[(int64_t)c * ld + t] = x[c];
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double g = (c <= j) ? gamma * U[uidx(c <= j ? c : j, j)] : 0.0;
      sb[(1 + j) * blk + (int64_t)c * ld + t] = x[c] + g;
      sb[(7 + j) * blk + (int64_t)c * ld + t] = x[c] - g;
    }
}

__global__ void __launch_bounds__(128)
sigma_gen_kernel(const double* __restrict__ est_x, const double* __restrict__ est_p,
                 const double* __restrict__ truth_x, double* __restrict__ sb, int64_t n, int64_t ld,
                 double gamma, uint32_t* __restrict__ sig_flags) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= n) return;
  double x[6], E[21], U[21];
#pragma unroll
  for (int c = 0; c < 6; ++c) x[c] = est_x[(int64_t)c * ld + t];
  const double* P = est_p + t * 36;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) E[uidx(i, j)] = P[i * 6 + j];
  const bool ok = chol_upper6_packed(E, U);
  if (!ok) repair_chol6_regs(E, U);  // reference: nearest-PD fallback; the target stays flagged
  write_sigma_blocks(sb, ld, t, x, U, gamma);
  if (truth_x && truth_x != sb + 13 * 6 * ld) {
#pragma unroll
    for (int c = 0; c < 6; ++c) sb[13 * 6 * ld + (int64_t)c * ld + t] = truth_x[(int64_t)c * ld + t];
  }
  sig_flags[t] = ok ? 0u : (uint32_t)PC_FLAG_NONPD_EST;
  const double r0[3] = {x[0], x[1], x[2]}, v0[3] = {x[3], x[4], x[5]};
  sig_flags[n + t] = orbit_rates_packed(r0, v0);  // substep rule input for all 13 points
}

// Coalesced store of one warp's 32 symmetric 6x6 matrices into an (N, 36) row-major array.
// Each lane owns one matrix (packed upper triangle E); a direct store would be 36 STG.64 per lane
// at a 288-byte lane stride (32 sectors per instruction).  Instead the warp transposes through a
// padded shared-memory tile (row stride 37 doubles: conflict-free for 64-bit lanes) and writes its
// 32 x 288 B = 9216 contiguous bytes with fully coalesced 256-byte rows.
constexpr int kStageRow = 37;
constexpr int kStageDoubles = 32 * kStageRow;

__device__ __forceinline__ void store_sym36_warp(double* __restrict__ g, int64_t warp_t0, int64_t n,
                                                 const double (&E)[21], double* __restrict__ stage,
                                                 int lane) {
  double* mine = stage + lane * kStageRow;
#pragma unroll
  for (int ra = 0; ra < 6; ++ra)
#pragma unroll
    for (int rb = 0; rb < 6; ++rb) mine[ra * 6 + rb] = E[ra <= rb ? uidx(ra, rb) : uidx(rb, ra)];
  __syncwarp();
  const int64_t rows = n - warp_t0 < 32 ? n - warp_t0 : 32;
  const int valid = (int)rows * 36;
  double* out = g + warp_t0 * 36;
#pragma unroll 4
  for (int k = lane; k < 32 * 36; k += 32) {
    const int row = k / 36, e = k - row * 36;
    if (k < valid) out[k] = stage[row * kStageRow + e];
  }
  __syncwarp();
}

constexpr int kVisTile = 256;  // sensors staged per pass of the fused visibility phase (8 KB)
// Each 32-sensor word occupies 33 double4 slots: lanes that work on different words of the same
// tile (the quad kernel) then read shared-memory addresses 33 * 32 B apart, i.e. 8 banks apart,
// and a 128-bit LDS with four distinct addresses is served in one wavefront instead of four.
constexpr int kVisRow = 33;
constexpr int kVisTileSlots = (kVisTile / 32) * kVisRow;

// Asynchronous copy of one chunk of sensor ROWS into shared memory with the bulk-copy engine
// (cp.async.bulk + mbarrier transaction count; UBLKCP in SASS): ONE thread issues one copy per row
// (a row's sensors are contiguous in global memory) at kernel entry, nobody touches the LSU, and
// the tile is waited for only when the visibility phase starts, so its latency hides behind the
// filter algebra.  A row is one 32-sensor word of one environment: global row r belongs to
// environment r / wpr and holds its sensors (r % wpr) * 32 .. + 31 (single catalog: one
// environment).  Slots past the environment's sensor count are filled with a NaN horizon term
// (never visible) by the other threads through the generic proxy.
// Environment of a target / of a sensor row.  A single catalog (env_targets == 0) skips the 64-bit
// divisions, which cost more than the whole visibility phase of a small tile.
// (32-bit divisions: target and row indices are below 2^32, and a 64-bit division is ~100 instructions
// at every inlined site of kernels whose cost at small catalogs is instruction fetch)
__device__ __forceinline__ int64_t env_of_target(int64_t t, int64_t env_targets) {
  return env_targets > 0 ? (int64_t)((uint32_t)t / (uint32_t)env_targets) : 0;
}
__device__ __forceinline__ void row_split(int64_t r, int64_t wpr, bool single, int64_t& e, int64_t& w) {
  if (single) {
    e = 0;
    w = r;
  } else {
    const uint32_t eq = (uint32_t)r / (uint32_t)wpr;
    e = eq;
    w = (int64_t)((uint32_t)r - eq * (uint32_t)wpr);
  }
}

__device__ __forceinline__ void sensor_tile_init(uint64_t* mbar) {
  if (threadIdx.x == 0) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// (rolled loops, 32-bit index arithmetic: unrolled over the 8 rows with 64-bit divisions this was 600
// instructions, inlined twice, on the path of kernels that are instruction-fetch bound at small N)
__device__ __forceinline__ void load_sensor_tile(double4* __restrict__ tile, uint64_t* mbar,
                                                 const double4* __restrict__ sens_q, int64_t r0,
                                                 int64_t row_end, int64_t Me64, int64_t wpr64, bool single,
                                                 int tpb) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const int nrows = (int)(row_end - r0 < kVisTile / 32 ? row_end - r0 : kVisTile / 32);
  const uint32_t Me = (uint32_t)Me64, wpr = (uint32_t)wpr64, rbase = (uint32_t)r0;
  if (threadIdx.x == 0) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    // earlier generic-proxy reads of the tile (previous chunk) are ordered before the async writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned total = 0;
#pragma unroll 1
    for (int rr = 0; rr < nrows; ++rr) {
      const uint32_t r = rbase + (uint32_t)rr;
      const uint32_t w = single ? r : r % wpr;
      const uint32_t left = Me - w * 32u;
      total += (left < 32u ? left : 32u) * 32u;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(total) : "memory");
#pragma unroll 1
    for (int rr = 0; rr < nrows; ++rr) {
      const uint32_t r = rbase + (uint32_t)rr;
      const uint32_t e = single ? 0u : r / wpr;
      const uint32_t w = r - e * wpr;
      const uint32_t left = Me - w * 32u;
      const unsigned bytes = (left < 32u ? left : 32u) * 32u;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + rr * kVisRow);
      const double4* src = sens_q + ((uint64_t)e * Me + w * 32u);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
  }
#pragma unroll 1
  for (int j = threadIdx.x; j < kVisTile; j += tpb) {
    const int rr = j >> 5;
    const uint32_t r = rbase + (uint32_t)rr;
    const uint32_t w = single ? r : r % wpr;
    const uint32_t sl = w * 32u + (uint32_t)(j & 31);
    if (!(rr < nrows && sl < Me)) tile[rr * kVisRow + (j & 31)] = make_double4(0.0, 0.0, 0.0, nan);
  }
}

// Wait for the chunk whose copies were issued `phase` chunks ago modulo 2 (mbarrier phase parity),
// then make the NaN padding written by the other threads visible to everybody.
__device__ __forceinline__ void wait_sensor_tile(uint64_t* mbar, unsigned phase) {
  const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
  unsigned done = 0;
  for (int spin = 0; !done && spin < (1 << 22); ++spin) {   // bounded: a lost transaction must not hang the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mb), "r"(phase) : "memory");
  }
  __syncthreads();
}

// One 32-sensor word of both visibility rows of a target.  Sensors are fetched from shared memory
// eight at a time before they are tested, so that the LDS latency is paid once per batch.
// (ROLLED: the small-catalog kernel is bound by instruction fetch and takes the loop; the large-catalog
// kernels are bound by issue slots and take the straight-line form)
template <bool ROLLED>
__device__ __forceinline__ void vis_word(const double4* __restrict__ row, const double (&tr)[3], double tq,
                                         const double (&xe)[6], double eq, double RE2, uint32_t& wa,
                                         uint32_t& wb) {
  wa = 0;
  wb = 0;
#pragma unroll(ROLLED ? 1 : 4)
  for (int b0 = 0; b0 < 32; b0 += 8) {
    double4 s[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) s[b] = row[b0 + b];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      if (sees(s[b], tr[0], tr[1], tr[2], tq, RE2)) wa |= 1u << (b0 + b);
      if (sees(s[b], xe[0], xe[1], xe[2], eq, RE2)) wb |= 1u << (b0 + b);
    }
  }
}

// ---- stand-alone peer exchange (pc_allgather_step, pc_peer_wait): see PeerPush in ukf_args.cuh -------
// Inside the step the push and the wait are part of the sensor kernel (kernels_env.cu); these two
// kernels are the same operations for a caller that exchanges a summary buffer outside a step.
__global__ void __launch_bounds__(128)
peer_push_kernel(PeerPush pp) {
  peer_push_block(pp, (int)blockIdx.x, (int)gridDim.x);
}

// One small CTA: hold the stream until every peer has published push number *epoch + 1 (bounded spin: a
// lost peer raises *err instead of hanging the GPU), then count the push (advance != 0).
__global__ void peer_wait_kernel(const uint64_t* __restrict__ table, int world, int rank,
                                 unsigned long long* __restrict__ epoch, unsigned int* __restrict__ err,
                                 int advance) {
  if ((int)threadIdx.x < world) peer_wait_thread(table, world, rank, epoch, err, (int)threadIdx.x);
  __syncthreads();
  if (advance && threadIdx.x == 0) *epoch += 1;
}

cudaError_t launch_peer_push(const PeerPush& pp, cudaStream_t stream) {
  peer_push_kernel<<<peer_push_blocks(pp.bytes), 128, 0, stream>>>(pp);
  return cudaGetLastError();
}

cudaError_t launch_peer_wait(const uint64_t* table, int world, int rank, unsigned long long* epoch,
                             unsigned int* err, int advance, cudaStream_t stream) {
  peer_wait_kernel<<<1, 64, 0, stream>>>(table, world, rank, epoch, err, advance);
  return cudaGetLastError();
}

// Stage 3 of the step, one THREAD per target: unscented transform, measurement update, outputs,
// next step's sigma points and -- fused, because the new estimate and the truth position are in
// registers and the kernel is otherwise HBM-latency bound -- the target's row of BOTH visibility
// maps (env.py:455-462) against the sensor tile staged in shared memory.
// MINB (CTAs per SM the register allocation is held to): 6 x 64 or 3 x 128 threads = 168 registers for
// catalogs where few targets are measured (occupancy is what counts: 1 M targets 0.99 vs 1.01 ms per step);
// 4 x 64 = 255 registers where many are (the update keeps ~200 doubles live: 652 -> 128 bytes of spills;
// 1024 environments 0.164 -> 0.152 ms per step, 100 000 targets all measured 0.144 -> 0.137).
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
ukf_finish_kernel(const __grid_constant__ UkfArgs a) {
  extern __shared__ __align__(32) double smem[];
  pdl_wait();  // launched programmatically behind the propagation kernel: everything below reads its output
  pdl_trigger();  // a programmatic successor (the observation pack) may be scheduled behind our last wave
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = smem + warp * kStageDoubles;
  double4* tile = reinterpret_cast<double4*>(smem + (TPB / 32) * kStageDoubles);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tile + kVisTileSlots);
  const int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (a.clock && a.advance_clock && t == 0) {
    // nobody reads t_now / step during this launch (kernels read the t_cur / step_cur copies)
    a.clock->t_now += a.dt;
    a.clock->step += 1;
  }
  const bool live = t < a.n;
  const int64_t tc = live ? t : a.n - 1;  // dead lanes shadow the last target and store nothing
  const int64_t warp_t0 = t - lane;
  const int64_t ld = a.ld, blk = 6 * a.ld;
  const double* __restrict__ sb = a.sb;
  const bool vis = a.vis_truth != nullptr;
  // sensor rows this CTA's targets need (one environment unless the catalog is a batch of envs)
  const int64_t cta_t0 = (int64_t)blockIdx.x * TPB;
  const int64_t cta_t1 = cta_t0 + TPB - 1 < a.n - 1 ? cta_t0 + TPB - 1 : a.n - 1;
  const bool single = a.env_targets <= 0;
  const int64_t row_first = vis ? env_of_target(cta_t0, a.env_targets) * a.wpr : 0;
  const int64_t row_end = vis ? (env_of_target(cta_t1, a.env_targets) + 1) * a.wpr : 0;
  if (vis) {
    sensor_tile_init(mbar);
    load_sensor_tile(tile, mbar, a.sens_q, row_first, row_end, a.env_sensors, a.wpr, single, TPB);
  }

  // ---- unscented transform (predictStateEstimate / predictCovariance, ukf_v2.py:261-292) ------
  double X0[6], tr[3];
#pragma unroll
  for (int c = 0; c < 6; ++c) X0[c] = sb[(int64_t)c * ld + tc];
  if (vis) {
#pragma unroll
    for (int c = 0; c < 3; ++c) tr[c] = sb[13 * blk + (int64_t)c * ld + tc];
  }
  double Sb[6], F[21];  // F = sum_j 2 b_j b_j^T + D_j D_j^T / 2
#pragma unroll
  for (int c = 0; c < 6; ++c) Sb[c] = 0.0;
#pragma unroll
  for (int k = 0; k < 21; ++k) F[k] = 0.0;
#pragma unroll 3
  for (int j = 0; j < 6; ++j) {
    double b[6], D[6], b2[6], Dh[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double xp = sb[(1 + j) * blk + (int64_t)c * ld + tc];
      const double xm = sb[(7 + j) * blk + (int64_t)c * ld + tc];
      b[c] = 0.5 * ((xp - X0[c]) + (xm - X0[c]));  // (x+ + x-)/2 - X_0
      D[c] = xp - xm;
      Sb[c] += b[c];
      b2[c] = b[c] + b[c];
      Dh[c] = 0.5 * D[c];
    }
#pragma unroll
    for (int ra = 0; ra < 6; ++ra)
#pragma unroll
      for (int rb = ra; rb < 6; ++rb)
        F[uidx(ra, rb)] = fma(Dh[ra], D[rb], fma(b2[ra], b[rb], F[uidx(ra, rb)]));
  }
  double m[6], px[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    m[c] = fma(2.0 * a.p.wi, Sb[c], a.p.eps_w * X0[c]);  // sum_i (X_i - X_0) = 2 sum_j b_j
    px[c] = X0[c] + m[c];                                  // pred_x = sigma_points . mean_weight
  }
  double E[21];  // pred_p = Xres Wc Xres^T + Q
#pragma unroll
  for (int ra = 0; ra < 6; ++ra)
#pragma unroll
    for (int rb = ra; rb < 6; ++rb) {
      const int k = uidx(ra, rb);
      // sum_j a_j a_j^T with a_j = b_j - m:  Ebb - (Sb m^T + m Sb^T) + 6 m m^T
      const double cross = fma(6.0 * m[ra], m[rb], -(Sb[ra] * m[rb] + m[ra] * Sb[rb]));
      E[k] = fma(a.p.wi, fma(2.0, cross, F[k]), fma(a.p.wc0 * m[ra], m[rb], a.p.Q[ra * 6 + rb]));
    }

  // ---- update ----------------------------------------------------------------------------------
  int tasked = 0;
  if (a.tasked && live) {
    tasked = a.tasked[t] != 0;
    if (tasked && a.clear_tasked) a.tasked[t] = 0;
  }
  uint32_t flags = a.sig_flags ? a.sig_flags[tc] : 0u;
  double xe[6], PE[21];
  double nis = __longlong_as_double(0x7ff8000000000000LL);
  if (a.out_pred_p) store_sym36_warp(a.out_pred_p, warp_t0, a.n, E, stage, lane);
  if (tasked) {
    const double t_now = a.clock ? a.clock->t_cur + a.dt : a.t_now;
    const uint64_t rng_step = a.clock ? a.clock->step_cur : a.rng_step;
    double z[6];
    if (a.z) {
#pragma unroll
      for (int c = 0; c < 6; ++c) z[c] = a.z[(int64_t)c * ld + t];
    } else {
      // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
      double nrm[6];
      normal6(a.rng_seed, rng_step, (uint64_t)t, nrm);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double acc = sb[13 * blk + (int64_t)c * ld + t];
#pragma unroll
        for (int k = 0; k <= c; ++k) acc = fma(a.p.Lr[c * 6 + k], nrm[k], acc);
        z[c] = acc;
      }
    }
    flags |= measurement_update_reg(sb, ld, blk, t, px, m, E, z, a.p, xe, PE, nis);
    if (a.out_z) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.out_z[(int64_t)c * ld + t] = z[c];
    }
    if (a.num_tasked) a.num_tasked[t] += 1;
    if (a.last_time_tasked) a.last_time_tasked[t] = (float)t_now;
  } else {
    // update(None) (ukf_v2.py:238-244): est_x = propagated centre point, est_p = pred_p
#pragma unroll
    for (int c = 0; c < 6; ++c) xe[c] = X0[c];
#pragma unroll
    for (int k = 0; k < 21; ++k) PE[k] = E[k];
  }

  // ---- outputs -----------------------------------------------------------------------------------
  store_sym36_warp(a.est_p, warp_t0, a.n, PE, stage, lane);
  double chk = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) chk += xe[c] + PE[uidx(c, c)];
  if (!isfinite(chk)) flags |= PC_FLAG_NONFINITE;
  if (live) {
    if (a.out_flags) a.out_flags[t] = flags;
    if (a.out_nis) a.out_nis[t] = nis;
    // scheduler summaries (env.py:599-609: traces of the float32 est_cov blocks)
    const float trp = ((float)PE[uidx(0, 0)] + (float)PE[uidx(1, 1)]) + (float)PE[uidx(2, 2)];
    const float trv = ((float)PE[uidx(3, 3)] + (float)PE[uidx(4, 4)]) + (float)PE[uidx(5, 5)];
    if (a.tr_pos) a.tr_pos[t] = trp;
    if (a.tr_vel) a.tr_vel[t] = trv;
    if (a.out_pred_x) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.out_pred_x[(int64_t)c * ld + t] = px[c];
    }
    if (a.truth_out) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.truth_out[(int64_t)c * ld + t] = sb[13 * blk + (int64_t)c * ld + t];
    }
    if (a.gen_next) {
      // factor the NEW estimate and lay out next step's sigma points (block 0 = est_x)
      double U[21];
      const bool ok = chol_upper6_packed(PE, U);
      if (!ok) repair_chol6_regs(PE, U);
      write_sigma_blocks(a.sb, ld, t, xe, U, a.p.gamma);
      a.sig_flags[t] = ok ? 0u : (uint32_t)PC_FLAG_NONPD_EST;
      const double r0[3] = {xe[0], xe[1], xe[2]}, v0[3] = {xe[3], xe[4], xe[5]};
      a.sig_flags[a.n + t] = orbit_rates_packed(r0, v0);
    } else if (a.est_x_out) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.est_x_out[(int64_t)c * ld + t] = xe[c];
    }
  }

  // ---- both visibility rows of this target for the next step (calcVisMap, visibility.py:56-116)
  if (vis) {
    const double RE2 = a.body_radius * a.body_radius;
    const double tq = horizon_q(tr[0], tr[1], tr[2], a.body_radius, a.vis_tol);
    const double eq = horizon_q(xe[0], xe[1], xe[2], a.body_radius, a.vis_tol);
    const int64_t wpr = a.wpr;
    const int64_t my_row0 = env_of_target(tc, a.env_targets) * wpr;
    unsigned phase = 0;
    for (int64_t r0 = row_first; r0 < row_end; r0 += kVisTile / 32, phase ^= 1u) {
      if (r0 != row_first) {
        __syncthreads();
        load_sensor_tile(tile, mbar, a.sens_q, r0, row_end, a.env_sensors, wpr, single, TPB);
      }
      wait_sensor_tile(mbar, phase);
      const int nrows = (int)(row_end - r0 < kVisTile / 32 ? row_end - r0 : kVisTile / 32);
      for (int64_t w = 0; w < wpr; ++w) {
        const int64_t rr = my_row0 + w - r0;
        if (rr >= 0 && rr < nrows) {
          uint32_t wa, wb;
          vis_word<false>(tile + rr * kVisRow, tr, tq, xe, eq, RE2, wa, wb);
          if (live) {
            a.vis_truth[t * wpr + w] = wa;
            a.vis_est[t * wpr + w] = wb;
          }
        }
      }
    }
  }
  if (a.peer.world && blockIdx.x == gridDim.x - 1) peer_wait_and_count(a.peer);
}

cudaError_t launch_sigma_gen(const double* est_x, const double* est_p, const double* truth_x, double* sb,
                             int64_t n, int64_t ld, double gamma, uint32_t* sig_flags,
                             cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  sigma_gen_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(est_x, est_p, truth_x, sb, n, ld,
                                                                   gamma, sig_flags);
  return cudaGetLastError();
}

// The same stage for SMALL catalogs, four lanes per target.  With one thread per target a 10k
// catalog is 313 warps of ~6000 dependent instructions each on a 592-scheduler machine: pure
// latency.  Here a quad shares one target: the six (+,-) sigma pairs are split over the lanes and
// their partial sums combined with two xor-shuffles (every lane ends up with bit-identical sums,
// so the 6x6 algebra that follows is simply replicated), stores and sigma-point writes are
// distributed by lane, and each lane packs every fourth 32-sensor word of both visibility rows.
constexpr int kQuadThreads = 128;
constexpr int64_t kQuadMaxTargets = 32 * 1024;  // above this one thread per target fills the machine

// ---- cooperative measurement update: one WARP per tasked target ------------------------------------
// At catalogs of ~10^4 targets the per-target kernel is one wave of latency-bound warps that execute
// straight-line code once: after the benchmark's L2 flush it is INSTRUCTION-FETCH bound (removing the
// never-executed in-line update from the kernel made the untasked path 5 us faster, profiles/README.md),
// and the few targets that received a measurement (at most one per sensor) lengthened the warps they
// sit in by more than half the untasked path.  With pc_step_args.task_of the quad kernel is launched
// with ceil(M / 2) extra CTAs in front of the grid; warp w of extra CTA c looks at sensor 2 c + w,
// and if that sensor validly tasks a target (and no lower-numbered sensor of the same environment
// tasks the same one) it does that target's WHOLE per-target stage -- transform, update, outputs,
// next sigma points, estimated visibility row.  The target's own quad then only writes what does not
// depend on the estimate (true visibility row).
//
// The update warp's code is written for SIZE, not for issue rate (one warp per SM runs it, so every
// instruction line is a cold miss): all operands live in shared memory, loops stay rolled, the
// Cholesky factorisation is one out-of-line routine called three times, and independent matrix
// entries are spread over the lanes (lane k < 21 owns entry (ra, rb) of an upper triangle).  Same
// equations as measurement_update_reg; the factorisations perform the same operations in the same
// order, the sums of the unscented transform are taken in a different order (round-off level).
struct CoopSmem {
  double X[14][6];                       // propagated sigma points of the target: block, component
  double P[36], W[36], U2[36], Us[36];   // pred_p; scratch; chol(pred_p); chol(innovation cvr)
  double G[36], PE[36], Un[36];          // gain factor; est_p; chol(est_p)
  double Sb[6], m[6], px[6], delta[6], nu[6], y[6], xe[6], dinv[6], nrm[6], z[6];
  long long nt;                          // num_tasked[t], fetched with the sigma points
};
constexpr int kCoopMaxSensors = 256;

__device__ __forceinline__ int tri_row(int k) { return (k >= 6) + (k >= 11) + (k >= 15) + (k >= 18) + (k >= 20); }

// Upper Cholesky of the symmetric 6x6 whose upper triangle is in W (row-major full-matrix layout),
// factor written to U's upper triangle.  Every lane factors the whole matrix in registers (the
// factorisation is one dependent chain: spreading it over lanes only adds shared-memory round trips
// -- 2.4 k cycles against 0.8 k) with chol_upper6_packed itself, so the factor is bit-identical to
// the in-line update's.  One copy of the code, called three times per update.
__device__ __noinline__ bool chol6_warp(const double* __restrict__ W, double* __restrict__ U, int lane) {
  double E[21], Ur[21];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) E[uidx(i, j)] = W[i * 6 + j];
  const bool ok = chol_upper6_packed(E, Ur);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) U[i * 6 + j] = Ur[uidx(i, j)];
  }
  __syncwarp();
  return ok;
}

// Rare: nearest-PD repair of the full symmetric matrix A, factor written to U (upper).
__device__ __noinline__ void repair6_warp(const double* __restrict__ A, double* __restrict__ U, int lane) {
  if (lane == 0) {
    double Ein[21], Uo[21];
#pragma unroll 1
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) {
        Ein[uidx(i, j)] = A[i * 6 + j];
        Uo[uidx(i, j)] = 0.0;
      }
    repair_chol6(Ein, Uo);
#pragma unroll 1
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) U[i * 6 + j] = Uo[uidx(i, j)];
  }
  __syncwarp();
}

// Benign operands for the dry runs: identity factors, a state well above the surface.
__device__ __forceinline__ void coop_scratch_init(CoopSmem& S) {
  const int lane = threadIdx.x & 31;
  double* f = reinterpret_cast<double*>(&S);
  for (int i = lane; i < (int)(sizeof(CoopSmem) / sizeof(double)); i += 32) f[i] = 0.0;
  __syncwarp();
  if (lane < 6) {
    S.U2[lane * 7] = 1.0;
    S.Us[lane * 7] = 1.0;
    S.P[lane * 7] = 1.0;
    S.dinv[lane] = 1.0;
    S.px[lane] = 7000.0;
    S.X[0][lane] = 7000.0;
  }
  __syncwarp();
}

// First half: loads, measurement, unscented transform, both factorisations.  Returns the flags so far.
__device__ __noinline__ uint32_t coop_part1(const UkfArgs& a, CoopSmem& S, int64_t t, int dry) {
  const int lane = threadIdx.x & 31;
  const bool wr = dry == 0;
  const int64_t ld = a.ld, blk = 6 * a.ld;
  const pc_ukf_params& p = a.p;
  const int k = lane < 21 ? lane : 20;
  const int ra = tri_row(k), rb = k - uidx(ra, ra) + ra;
  double* Xf = &S.X[0][0];
  // every global load of the update is issued here; the measurement draw below runs in their shadow
  double xin[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = lane + 32 * u;
    const int b = i / 6, c = i - 6 * b;
    if (wr && i < 84) xin[u] = a.sb[(int64_t)b * blk + (int64_t)c * ld + t];
  }
  long long nt_in = 0;
  if (wr && lane == 0 && a.num_tasked) nt_in = a.num_tasked[t];
  uint32_t flags = a.sig_flags && wr ? a.sig_flags[t] : 0u;
  const uint64_t rng_step = a.clock ? a.clock->step_cur : a.rng_step;
  if (a.z) {
    if (lane < 6 && wr) S.z[lane] = a.z[(int64_t)lane * ld + t];
  } else if (lane < 3 && wr) {
    // the three Box-Muller pairs of normal6(), one per lane (same counters, same arithmetic); runs in
    // the shadow of the loads above, so the warm-up warp skips it and starts ahead at the transform
    uint32_t c[4] = {(uint32_t)t, (uint32_t)((uint64_t)t >> 32), (uint32_t)rng_step, (uint32_t)(rng_step >> 32)};
    if (lane == 2) c[3] ^= 0x5bd1e995u + 1u;
    uint32_t k0 = (uint32_t)a.rng_seed, k1 = (uint32_t)(a.rng_seed >> 32);
#pragma unroll 1
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
    const uint32_t ua = lane == 1 ? c[2] : c[0], ub = lane == 1 ? c[3] : c[1];
    const double u1 = ((double)ua + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)ub + 0.5) * (1.0 / 4294967296.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    S.nrm[2 * lane] = rad * cs;
    S.nrm[2 * lane + 1] = rad * sn;
  }
  if (wr) PC_COOP_CLOCK(1);
#pragma unroll
  for (int u = 0; u < 3; ++u)
    if (wr && lane + 32 * u < 84) Xf[lane + 32 * u] = xin[u];
  if (lane == 0) S.nt = nt_in;
  __syncwarp();
  if (wr) PC_COOP_CLOCK(2);

  // ---- unscented transform ----------------------------------------------------------------------
  if (lane < 6) {
    const double x0 = S.X[0][lane];
    double sb_ = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) sb_ += 0.5 * ((S.X[1 + j][lane] - x0) + (S.X[7 + j][lane] - x0));
    const double mm = fma(2.0 * p.wi, sb_, p.eps_w * x0);
    const double pxx = x0 + mm;
    S.Sb[lane] = sb_;
    S.m[lane] = mm;
    S.px[lane] = pxx;
    S.delta[lane] = p.eps_w * pxx;
    if (!a.z) {   // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
      double acc = S.X[13][lane];
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (i <= lane) acc = fma(p.Lr[lane * 6 + i], S.nrm[i], acc);
      S.z[lane] = acc;
    }
  }
  double F = 0.0;
  {
    const double x0a = S.X[0][ra], x0b = S.X[0][rb];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double pa = S.X[1 + j][ra], ma = S.X[7 + j][ra], pb = S.X[1 + j][rb], mb = S.X[7 + j][rb];
      const double ba = 0.5 * ((pa - x0a) + (ma - x0a)), bb = 0.5 * ((pb - x0b) + (mb - x0b));
      F = fma(0.5 * (pa - ma), pb - mb, fma(ba + ba, bb, F));
    }
  }
  __syncwarp();
  if (lane < 21) {
    const double ma = S.m[ra], mb = S.m[rb];
    const double cross = fma(6.0 * ma, mb, -(S.Sb[ra] * mb + ma * S.Sb[rb]));
    const double e = fma(p.wi, fma(2.0, cross, F), fma(p.wc0 * ma, mb, p.Q[ra * 6 + rb]));
    S.P[ra * 6 + rb] = e;
    S.P[rb * 6 + ra] = e;
  }
  if (lane < 6) S.nu[lane] = S.z[lane] - (S.px[lane] + S.delta[lane]);
  __syncwarp();
  if (a.out_pred_p && wr) {
#pragma unroll
    for (int i = lane; i < 36; i += 32) a.out_pred_p[t * 36 + i] = S.P[i];
  }
  if (wr) PC_COOP_CLOCK(3);
  if (!chol6_warp(S.P, S.U2, lane) && wr) {
    flags |= PC_FLAG_NONPD_PRED;
    repair6_warp(S.P, S.U2, lane);
  }
  if (wr) PC_COOP_CLOCK(4);

  // ---- innovation covariance and its factor -------------------------------------------------------
  const double cuu = 2.0 * p.wi * p.gamma * p.gamma;
  const double cdd = p.wc0 + 12.0 * p.wi;
  if (lane < 21) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j >= rb) acc = fma(S.U2[ra * 6 + j], S.U2[rb * 6 + j], acc);
    S.W[ra * 6 + rb] = fma(cuu, acc, fma(cdd * S.delta[ra], S.delta[rb], p.R[ra * 6 + rb]));
  }
  __syncwarp();
  if (!chol6_warp(S.W, S.Us, lane)) flags |= PC_FLAG_NONPD_INNOV;
  if (lane < 6) S.dinv[lane] = 1.0 / S.Us[lane * 7];
  __syncwarp();
  if (wr) PC_COOP_CLOCK(5);
  return flags;
}

// Second half: gain, new estimate, every output of the target, next sigma points, estimated
// visibility row.  With dry != 0 nothing is written to global memory: the second warp of an update
// CTA runs this half on scratch operands WHILE the first warp is in the first half, so that the
// instruction lines are already in the SM's instruction cache when the real data arrives (the update
// warp is alone on its code: every line it touches is otherwise a cold miss).
template <bool HOST>
__device__ __noinline__ void coop_part2(const UkfArgs& a, CoopSmem& S, int64_t t, uint32_t flags, int dry) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t ld = a.ld, blk = 6 * a.ld;
  const pc_ukf_params& p = a.p;
  const int k = lane < 21 ? lane : 20;
  const int ra = tri_row(k), rb = k - uidx(ra, ra) + ra;
  const bool wr = dry == 0;
  const double t_now = a.clock ? a.clock->t_cur + a.dt : a.t_now;
  const double c3 = p.wc0 - p.wm0;
  const double wg = p.wi * p.gamma;

  // ---- gain factor G = C Us^-1 (lane a < 6: row a) beside y = Us^-T nu (lane 6) -------------------
  double nis = 0.0;
  if (lane < 6) {
    const double sa = S.delta[lane] + c3 * S.m[lane];
    double D[6], g[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) D[j] = S.X[1 + j][lane] - S.X[7 + j][lane];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int j = b; j < 6; ++j) acc = fma(D[j], S.U2[b * 6 + j], acc);
      double cab = fma(wg, acc, sa * S.delta[b]);
#pragma unroll
      for (int i = 0; i < b; ++i) cab = fma(-g[i], S.Us[i * 6 + b], cab);
      g[b] = cab * S.dinv[b];
      S.G[lane * 6 + b] = g[b];
    }
  } else if (lane == 6) {
    double y[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      double acc = S.nu[b];
#pragma unroll
      for (int i = 0; i < b; ++i) acc = fma(-S.Us[i * 6 + b], y[i], acc);
      y[b] = acc * S.dinv[b];
      S.y[b] = y[b];
      nis = fma(y[b], y[b], nis);
    }
  }
  nis = __shfl_sync(full, nis, 6);
  __syncwarp();
  if (lane < 6) {
    double x = S.px[lane];
#pragma unroll
    for (int b = 0; b < 6; ++b) x = fma(S.G[lane * 6 + b], S.y[b], x);
    S.xe[lane] = x;
  }
  if (lane < 21) {
    double acc = S.P[ra * 6 + rb];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc = fma(-S.G[ra * 6 + i], S.G[rb * 6 + i], acc);
    S.PE[ra * 6 + rb] = acc;
    S.PE[rb * 6 + ra] = acc;
  }
  __syncwarp();

  // ---- outputs ---------------------------------------------------------------------------------------
  if (wr) PC_COOP_CLOCK(6);
  double chk = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) chk += S.xe[c] + S.PE[c * 7];
  if (!isfinite(chk)) flags |= PC_FLAG_NONFINITE;
#pragma unroll
  for (int i = lane; i < 36; i += 32)
    if (wr) {
      a.est_p[t * 36 + i] = S.PE[i];
      if (HOST && a.h_cov) a.h_cov[t * 36 + i] = (float)S.PE[i];
    }
  if (lane == 0 && wr) {
    if (a.out_flags) a.out_flags[t] = flags;
    if (a.out_nis) a.out_nis[t] = nis;
    const float trp = ((float)S.PE[0] + (float)S.PE[7]) + (float)S.PE[14];
    const float trv = ((float)S.PE[21] + (float)S.PE[28]) + (float)S.PE[35];
    if (a.tr_pos) a.tr_pos[t] = trp;
    if (a.tr_vel) a.tr_vel[t] = trv;
    const int64_t nt = S.nt + 1;
    if (a.num_tasked) a.num_tasked[t] = nt;
    if (a.last_time_tasked) a.last_time_tasked[t] = (float)t_now;
    if (HOST) {
      if (a.h_tr_pos) a.h_tr_pos[t] = trp;
      if (a.h_tr_vel) a.h_tr_vel[t] = trv;
      if (a.h_num_tasked) a.h_num_tasked[t] = nt;
      if (a.h_ltt) a.h_ltt[t] = (float)t_now;
    }
  }
  if (lane < 6 && wr) {
    if (a.out_z) a.out_z[(int64_t)lane * ld + t] = S.z[lane];
    if (a.out_pred_x) a.out_pred_x[(int64_t)lane * ld + t] = S.px[lane];
    if (a.truth_out) a.truth_out[(int64_t)lane * ld + t] = S.X[13][lane];
  }
  if (wr) PC_COOP_CLOCK(7);
}

// Third piece: factor of the new covariance, next step's sigma points, estimated visibility row.
template <bool HOST>
__device__ __noinline__ void coop_part3(const UkfArgs& a, CoopSmem& S, int64_t t, int dry) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t ld = a.ld, blk = 6 * a.ld;
  const pc_ukf_params& p = a.p;
  const bool wr = dry == 0;
  const int64_t env = a.vis_est ? env_of_target(t, a.env_targets) : 0;
  if (a.gen_next) {
    const bool ok = chol6_warp(S.PE, S.Un, lane);
    if (!ok && wr) repair6_warp(S.PE, S.Un, lane);
    if (lane == 0) {
      const double r0[3] = {S.xe[0], S.xe[1], S.xe[2]}, v0[3] = {S.xe[3], S.xe[4], S.xe[5]};
      const uint32_t rate = orbit_rates_packed(r0, v0);
      if (wr) {
        a.sig_flags[t] = ok ? 0u : (uint32_t)PC_FLAG_NONPD_EST;
        a.sig_flags[a.n + t] = rate;
      }
    }
#pragma unroll
    for (int i = lane; i < 78; i += 32) {
      const int b = i / 6, c = i - 6 * b;
      double v = S.xe[c];
      if (b > 0) {
        const int j = b <= 6 ? b - 1 : b - 7;
        const double gq = c <= j ? p.gamma * S.Un[c * 6 + j] : 0.0;
        v = b <= 6 ? v + gq : v - gq;
      }
      if (wr) a.sb[(int64_t)b * blk + (int64_t)c * ld + t] = v;
    }
  } else if (a.est_x_out && lane < 6 && wr) {
    a.est_x_out[(int64_t)lane * ld + t] = S.xe[lane];
  }
  if (wr) PC_COOP_CLOCK(8);
  if (a.vis_est) {
    const double RE2 = a.body_radius * a.body_radius;
    const double ex = S.xe[0], ey = S.xe[1], ez = S.xe[2];
    const double eq = horizon_q(ex, ey, ez, a.body_radius, a.vis_tol);
#pragma unroll 1
    for (int64_t w0 = 0; w0 < a.wpr; w0 += 4) {   // four words' sensors in flight at a time
      double4 sv[4];
      bool have[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t sl = (w0 + u) * 32 + lane;
        have[u] = w0 + u < a.wpr && sl < a.env_sensors;
        if (have[u]) sv[u] = a.sens_q[env * a.env_sensors + sl];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool v = have[u] && sees(sv[u], ex, ey, ez, eq, RE2);
        const uint32_t word = __ballot_sync(full, v);
        if (lane == 0 && wr && w0 + u < a.wpr) {
          a.vis_est[t * a.wpr + w0 + u] = word;
          if (HOST && a.h_vis_est) a.h_vis_est[t * a.wpr + w0 + u] = word;
        }
      }
    }
  }
  if (wr) PC_COOP_CLOCK(9);
}

__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

__device__ __forceinline__ void store_sym36_quad(double* __restrict__ g, const double (&E)[21], int q) {
  double2* row = reinterpret_cast<double2*>(g);
#pragma unroll
  for (int c = 0; c < 18; ++c) {
    const int ra = (2 * c) / 6, rb = (2 * c) % 6;  // elements (ra, rb) and (ra, rb + 1)
    if ((c & 3) == q)
      row[c] = make_double2(E[ra <= rb ? uidx(ra, rb) : uidx(rb, ra)],
                            E[ra <= rb + 1 ? uidx(ra, rb + 1) : uidx(rb + 1, ra)]);
  }
}

// regular CTAs: sensor tile, then (host-mirror form) a 1 152-byte float32 covariance stage per warp;
// update CTAs: four CoopSmem
constexpr size_t kQuadTileBytes = sizeof(double4) * kVisTileSlots;
constexpr size_t kQuadStageBytes = (kQuadThreads / 32) * 8 * 36 * sizeof(float);
constexpr size_t kQuadSmem = kQuadTileBytes + kQuadStageBytes > 4 * sizeof(CoopSmem) ? kQuadTileBytes + kQuadStageBytes
                                                                                   : 4 * sizeof(CoopSmem);
// float32 copy of the same matrix (the observation's est_cov) as nine 16-byte chunks over the quad
__device__ __forceinline__ void store_sym36_quad_f32(float* __restrict__ g, const double (&E)[21], int q) {
  float4* row = reinterpret_cast<float4*>(g);
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    if ((c & 3) == q) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = 4 * c + u, r = e / 6, cc = e % 6;
        v[u] = (float)E[r <= cc ? uidx(r, cc) : uidx(cc, r)];
      }
      row[c] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// HOST: also store what the caller reads on the host (float32 est_cov, covariance traces, estimated
// visibility words) into mapped host memory as it is produced (pc_step_host's zero-copy form): the
// PCIe transfer then runs under this kernel instead of behind it.  A separate instantiation, so that the
// device-only step does not carry the code.
template <bool COOP, bool HOST>
__global__ void __launch_bounds__(kQuadThreads, COOP ? 3 : 2)   // in-line update: 255 registers (168: 732 bytes of spills)
ukf_finish_quad_kernel(const __grid_constant__ UkfArgs a) {
  constexpr bool ROLE = COOP;
  __shared__ __align__(32) unsigned char smem_raw[kQuadSmem];
  __shared__ uint64_t mbar_store;
  static_assert(kQuadThreads == 128, "an update CTA is one real warp and three warm-up warps");
  double4* tile = reinterpret_cast<double4*>(smem_raw);
  uint64_t* mbar = &mbar_store;
  if (ROLE && (int64_t)blockIdx.x < a.task_sensors && threadIdx.x >= 32) {
    // Warps 1..3 of an update CTA, BEFORE the dependency wait: these CTAs are the first of the grid, so
    // they become resident while the propagation kernel's last wave is still running.  Each warp runs
    // one piece of the update on scratch operands (nothing is written to global memory, whatever it
    // reads may be stale): the instruction lines of that piece are then in this SM's instruction
    // cache when warp 0 gets there with the real data.  (If the CTA only becomes resident after the
    // propagation has finished, the three pieces are fetched in parallel beside warp 0's first one.)
    CoopSmem& S = reinterpret_cast<CoopSmem*>(smem_raw)[threadIdx.x >> 5];
    coop_scratch_init(S);
    if (threadIdx.x < 64) coop_part1(a, S, 0, 1);
    else if (threadIdx.x < 96) coop_part2<HOST>(a, S, 0, 0u, 1);
    else coop_part3<HOST>(a, S, 0, 1);
  }
  pdl_wait();  // launched programmatically behind the propagation kernel: everything below reads its output
  pdl_trigger();  // a programmatic successor (the observation pack) may be scheduled behind our last wave
  const int q = threadIdx.x & 3;
  if (a.clock && a.advance_clock && blockIdx.x == 0 && threadIdx.x == 0) {
    a.clock->t_now += a.dt;
    a.clock->step += 1;
  }
  // the first M CTAs are update CTAs (one per sensor, see above), the rest take 16 targets each
  const int64_t nb_upd = ROLE ? a.task_sensors : 0;
  if (ROLE && (int64_t)blockIdx.x < nb_upd) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x;
    int64_t tt = a.task_of[j];
    const int j0 = (int)(((uint32_t)j / (uint32_t)a.task_env_sensors) * (uint32_t)a.task_env_sensors);
    bool dup = false;   // a lower-numbered sensor of the same environment already owns this target
#pragma unroll 1
    for (int jj = j0 + lane; jj < (int)j; jj += 32) dup |= a.task_of[jj] == tt;
    if (__any_sync(0xffffffffu, dup)) tt = -1;
    if (tt >= 0 && tt < a.n && threadIdx.x < 32) {
      CoopSmem& S = reinterpret_cast<CoopSmem*>(smem_raw)[0];
      PC_COOP_CLOCK(0);
      const uint32_t fl = coop_part1(a, S, tt, 0);
      coop_part2<HOST>(a, S, tt, fl, 0);
      coop_part3<HOST>(a, S, tt, 0);
    }
    return;
  }
  const int64_t cta = (int64_t)blockIdx.x - nb_upd;
  const int64_t t = (cta * kQuadThreads + threadIdx.x) >> 2;
  PC_PHASE_CLOCK(0);
  const bool live = t < a.n;
  const int64_t tc = live ? t : a.n - 1;
  const int64_t ld = a.ld, blk = 6 * a.ld;
  const double* __restrict__ sb = a.sb;
  const bool vis = a.vis_truth != nullptr;
  const int64_t cta_t0 = cta * (kQuadThreads / 4);
  const int64_t cta_t1 = cta_t0 + kQuadThreads / 4 - 1 < a.n - 1 ? cta_t0 + kQuadThreads / 4 - 1 : a.n - 1;
  const bool single = a.env_targets <= 0;
  const int64_t row_first = vis ? env_of_target(cta_t0, a.env_targets) * a.wpr : 0;
  const int64_t row_end = vis ? (env_of_target(cta_t1, a.env_targets) + 1) * a.wpr : 0;

  // ---- every global load of the untasked path is issued here, before any arithmetic -------------
  // (one round trip to HBM instead of four: at this occupancy latency is the whole cost)
  double X0[6], tr[3], xp[2][6], xm[2][6];
#pragma unroll
  for (int c = 0; c < 6; ++c) X0[c] = sb[(int64_t)c * ld + tc];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = (q + 4 * h < 6) ? q + 4 * h : 5;  // lanes 2, 3 have no second pair: shadow pair 5
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      xp[h][c] = sb[(1 + j) * blk + (int64_t)c * ld + tc];
      xm[h][c] = sb[(7 + j) * blk + (int64_t)c * ld + tc];
    }
  }
  if (vis) {
#pragma unroll
    for (int c = 0; c < 3; ++c) tr[c] = sb[13 * blk + (int64_t)c * ld + tc];
  }
  const int tflag = a.tasked ? (int)a.tasked[tc] : 0;
  uint32_t flags = a.sig_flags ? a.sig_flags[tc] : 0u;
  // host mirror: the bookkeeping of an unmeasured target does not change, but the host block is
  // rewritten in full every step (no state to keep consistent across resets)
  int64_t nt0 = 0;
  float ltt0 = 0.0f;
  if (HOST && q == 0) {
    if (a.h_num_tasked) nt0 = a.num_tasked[tc];
    if (a.h_ltt) ltt0 = a.last_time_tasked[tc];
  }
  // the sensor tile is requested AFTER the register loads are in flight (one thread issues the bulk
  // copies; everybody else would otherwise wait for it at the barrier before issuing anything)
  if (vis) {
    sensor_tile_init(mbar);
    load_sensor_tile(tile, mbar, a.sens_q, row_first, row_end, a.env_sensors, a.wpr, single, kQuadThreads);
  }
  PC_PHASE_CLOCK(1);

  // ---- unscented transform: lane q takes pairs q and q + 4 ------------------------------------
  double Sb[6], F[21];
#pragma unroll
  for (int c = 0; c < 6; ++c) Sb[c] = 0.0;
#pragma unroll
  for (int k = 0; k < 21; ++k) F[k] = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (q + 4 * h < 6) {
      double b[6], D[6], b2[6], Dh[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        b[c] = 0.5 * ((xp[h][c] - X0[c]) + (xm[h][c] - X0[c]));
        D[c] = xp[h][c] - xm[h][c];
        Sb[c] += b[c];
        b2[c] = b[c] + b[c];
        Dh[c] = 0.5 * D[c];
      }
#pragma unroll
      for (int ra = 0; ra < 6; ++ra)
#pragma unroll
        for (int rb = ra; rb < 6; ++rb)
          F[uidx(ra, rb)] = fma(Dh[ra], D[rb], fma(b2[ra], b[rb], F[uidx(ra, rb)]));
    }
  }
  PC_PHASE_CLOCK(2);
#pragma unroll
  for (int c = 0; c < 6; ++c) Sb[c] = quad_sum(Sb[c]);
#pragma unroll
  for (int k = 0; k < 21; ++k) F[k] = quad_sum(F[k]);

  PC_PHASE_CLOCK(3);
  double m[6], px[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    m[c] = fma(2.0 * a.p.wi, Sb[c], a.p.eps_w * X0[c]);
    px[c] = X0[c] + m[c];
  }
  double E[21];
#pragma unroll
  for (int ra = 0; ra < 6; ++ra)
#pragma unroll
    for (int rb = ra; rb < 6; ++rb) {
      const int k = uidx(ra, rb);
      const double cross = fma(6.0 * m[ra], m[rb], -(Sb[ra] * m[rb] + m[ra] * Sb[rb]));
      E[k] = fma(a.p.wi, fma(2.0, cross, F[k]), fma(a.p.wc0 * m[ra], m[rb], a.p.Q[ra * 6 + rb]));
    }

  // ---- update (replicated over the quad: identical inputs, identical results) -------------------
  PC_PHASE_CLOCK(4);
  // every lane has read the flag (dead lanes shadow the last target) before lane 0 of a quad clears it
  __syncwarp();
  const int tasked = live && tflag != 0;
  if (tasked && a.clear_tasked && q == 0) a.tasked[t] = 0;
  // with update warps, a tasked target's quad leaves everything that depends on the estimate to its
  // update warp (which may already be overwriting the target's sigma points: nothing loaded above
  // is used for it) and only packs the true visibility row
  const bool mine = live && !(COOP && tasked);
  double xe[6], PE[21];
  double nis = __longlong_as_double(0x7ff8000000000000LL);
  if (a.out_pred_p && mine) store_sym36_quad(a.out_pred_p + t * 36, E, q);
  if (!COOP && tasked) {
    const double t_now = a.clock ? a.clock->t_cur + a.dt : a.t_now;
    const uint64_t rng_step = a.clock ? a.clock->step_cur : a.rng_step;
    double z[6];
    if (a.z) {
#pragma unroll
      for (int c = 0; c < 6; ++c) z[c] = a.z[(int64_t)c * ld + t];
    } else {
      // Target.getMeasurement (agents.py:173-183): z ~ N(truth', R)
      double nrm[6];
      normal6(a.rng_seed, rng_step, (uint64_t)t, nrm);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double acc = sb[13 * blk + (int64_t)c * ld + t];
#pragma unroll
        for (int k = 0; k <= c; ++k) acc = fma(a.p.Lr[c * 6 + k], nrm[k], acc);
        z[c] = acc;
      }
    }
    flags |= measurement_update_reg(sb, ld, blk, t, px, m, E, z, a.p, xe, PE, nis);
    if (q == 0) {
      if (a.out_z) {
#pragma unroll
        for (int c = 0; c < 6; ++c) a.out_z[(int64_t)c * ld + t] = z[c];
      }
      if (a.num_tasked) a.num_tasked[t] += 1;
      if (a.last_time_tasked) a.last_time_tasked[t] = (float)t_now;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c) xe[c] = X0[c];
#pragma unroll
    for (int k = 0; k < 21; ++k) PE[k] = E[k];
  }

  PC_PHASE_CLOCK(5);
  // ---- outputs, distributed over the quad ----------------------------------------------------------
  double chk = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) chk += xe[c] + PE[uidx(c, c)];
  if (!isfinite(chk)) flags |= PC_FLAG_NONFINITE;
  if (live && !mine && a.truth_out && q == 3) {
#pragma unroll
    for (int c = 0; c < 6; ++c) a.truth_out[(int64_t)c * ld + t] = sb[13 * blk + (int64_t)c * ld + t];
  }
  if (HOST && a.h_cov) {
    // float32 est_cov of the warp's 8 targets: staged in shared memory, then written to the host block
    // as 16-byte stores at consecutive addresses (scattered 16-byte stores become one PCIe packet each:
    // measured 17 GB/s against ~50 for full lines)
    const int lane = threadIdx.x & 31;
    float* stg = reinterpret_cast<float*>(smem_raw + kQuadTileBytes) + (threadIdx.x >> 5) * (8 * 36);
    store_sym36_quad_f32(stg + (lane >> 2) * 36, PE, q);
    const unsigned mm = __ballot_sync(0xffffffffu, mine);
    __syncwarp();
    const int64_t tw0 = t - (lane >> 2);
    float* dst = a.h_cov + tw0 * 36;     // (4-byte stores: the block's covariance section is not 16-byte aligned in general)
#pragma unroll
    for (int i = lane; i < 288; i += 32)
      if ((mm >> (4 * (i / 36))) & 1u) dst[i] = stg[i];
  }
  if (mine) {
    store_sym36_quad(a.est_p + t * 36, PE, q);
    if (q == 1) {
      if (a.out_flags) a.out_flags[t] = flags;
      if (a.out_nis) a.out_nis[t] = nis;
      const float trp = ((float)PE[uidx(0, 0)] + (float)PE[uidx(1, 1)]) + (float)PE[uidx(2, 2)];
      const float trv = ((float)PE[uidx(3, 3)] + (float)PE[uidx(4, 4)]) + (float)PE[uidx(5, 5)];
      if (a.tr_pos) a.tr_pos[t] = trp;
      if (a.tr_vel) a.tr_vel[t] = trv;
      if (HOST) {
        if (a.h_tr_pos) a.h_tr_pos[t] = trp;
        if (a.h_tr_vel) a.h_tr_vel[t] = trv;
      }
    }
    if (HOST && q == 0) {
      if (a.h_num_tasked) a.h_num_tasked[t] = nt0;
      if (a.h_ltt) a.h_ltt[t] = ltt0;
    }
    if (a.out_pred_x && q == 2) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.out_pred_x[(int64_t)c * ld + t] = px[c];
    }
    if (a.truth_out && q == 3) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.truth_out[(int64_t)c * ld + t] = sb[13 * blk + (int64_t)c * ld + t];
    }
    PC_PHASE_CLOCK(6);
    if (a.gen_next) {
      double U[21];
      const bool ok = chol_upper6_packed(PE, U);
      if (!ok) repair_chol6_regs(PE, U);
      // block 0 (centre) and the 12 offset blocks, block i written by lane i & 3
      if (q == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) a.sb[(int64_t)c * ld + t] = xe[c];
        a.sig_flags[t] = ok ? 0u : (uint32_t)PC_FLAG_NONPD_EST;
        const double r0[3] = {xe[0], xe[1], xe[2]}, v0[3] = {xe[3], xe[4], xe[5]};
        a.sig_flags[a.n + t] = orbit_rates_packed(r0, v0);
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        // blocks 1 + j and 7 + j: x +- gamma U[:, j]
        if (((1 + j) & 3) == q) {
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double gq = (c <= j) ? a.p.gamma * U[uidx(c <= j ? c : j, j)] : 0.0;
            a.sb[(1 + j) * blk + (int64_t)c * ld + t] = xe[c] + gq;
          }
        }
        if (((7 + j) & 3) == q) {
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double gq = (c <= j) ? a.p.gamma * U[uidx(c <= j ? c : j, j)] : 0.0;
            a.sb[(7 + j) * blk + (int64_t)c * ld + t] = xe[c] - gq;
          }
        }
      }
    } else if (a.est_x_out && q == 0) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a.est_x_out[(int64_t)c * ld + t] = xe[c];
    }
  }

  PC_PHASE_CLOCK(7);
  // ---- visibility: lane q packs words q, q + 4, ... of both rows ---------------------------------
  if (vis) {
    const double RE2 = a.body_radius * a.body_radius;
    const double tq = horizon_q(tr[0], tr[1], tr[2], a.body_radius, a.vis_tol);
    const double eq = horizon_q(xe[0], xe[1], xe[2], a.body_radius, a.vis_tol);
    const int64_t wpr = a.wpr;
    const int64_t my_row0 = env_of_target(tc, a.env_targets) * wpr;
    unsigned phase = 0;
    for (int64_t r0 = row_first; r0 < row_end; r0 += kVisTile / 32, phase ^= 1u) {
      if (r0 != row_first) {
        __syncthreads();
        load_sensor_tile(tile, mbar, a.sens_q, r0, row_end, a.env_sensors, wpr, single, kQuadThreads);
      }
      wait_sensor_tile(mbar, phase);
      PC_PHASE_CLOCK(8);
      const int nrows = (int)(row_end - r0 < kVisTile / 32 ? row_end - r0 : kVisTile / 32);
      // lane q packs words q, q + 4, ... of ITS target: the four lanes of a quad work on four
      // different tile rows at the same time (hence the 33-slot rows)
      for (int64_t w = q; w < wpr; w += 4) {
        const int64_t rr = my_row0 + w - r0;
        if (rr >= 0 && rr < nrows) {
          uint32_t wa, wb;
          vis_word<true>(tile + rr * kVisRow, tr, tq, xe, eq, RE2, wa, wb);
          if (live) a.vis_truth[t * wpr + w] = wa;
          if (mine) {
            a.vis_est[t * wpr + w] = wb;
            if (HOST && a.h_vis_est) a.h_vis_est[t * wpr + w] = wb;
          }
        }
      }
    }
  }
  PC_PHASE_CLOCK(9);
  if (a.peer.world && blockIdx.x == gridDim.x - 1) peer_wait_and_count(a.peer);
}

template <int TPB, int MINB>
static cudaError_t launch_finish_tpb(const UkfArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)(TPB / 32) * kStageDoubles * sizeof(double) + kVisTileSlots * sizeof(double4) + 16;
  return launch_pdl(ukf_finish_kernel<TPB, MINB>, dim3((unsigned)((a.n + TPB - 1) / TPB)), dim3(TPB), smem, stream, true, a);
}

// Small catalogs cannot fill the machine with one thread per target: they take the quad kernel
// (four lanes per target); large ones get one thread per target with 64- or 128-thread CTAs
// (sensor tile amortised over more warps).  `variant` (pc_option PC_OPT_FINISH_KERNEL) pins the choice:
// 1 quad + update warps (when task_of is given), 2 quad with the in-line update, 3 / 4 one thread per
// target with 64- / 128-thread CTAs; 0 = the size rule.
// (mask_only: the tasking comes as a caller's mask without an actions vector, so ANY number of targets may
// be measured -- BASELINE.json configs[4] measures all of them -- and the four-lane kernel, which replicates
// the update algebra over its lanes, loses to one thread per target from ~10 k targets on: measured with
// every target measured, us per step, four lanes (255-register build) vs one thread (255-register build):
// 6 250 targets 30.7 vs 37.3, 12 500 targets 49.3 vs 41.0, 16 384 targets 52.5 vs 43.9.
// With an actions vector at most one target per sensor is measured and the crossover stays at 32 k.)
static int finish_mapping(int64_t n, int variant, bool mask_only = false) {
  if (variant >= 1 && variant <= 4) return variant;
  if (n <= (mask_only ? kQuadMaxTargets / 4 : kQuadMaxTargets)) return 1;
  return n <= 128 * 1024 ? 3 : 4;
}

// Does launch_ukf_finish take the update-warp kernel (the only one that fills the host mirrors)?
bool finish_fills_host_mirrors(int64_t n, int64_t sensors, int variant) {
  return n > 0 && finish_mapping(n, variant) == 1 && sensors > 0 && sensors <= kCoopMaxSensors;
}

cudaError_t launch_ukf_finish(const UkfArgs& a, cudaStream_t stream) {
  if (a.n <= 0) return cudaSuccess;
  const int mapping = finish_mapping(a.n, a.variant, a.tasked != nullptr && a.task_of == nullptr && !a.clear_tasked);
  if (mapping <= 2) {
    const unsigned ctas = (unsigned)((4 * a.n + kQuadThreads - 1) / kQuadThreads);
    if (mapping == 1 && a.task_of && a.task_sensors > 0 && a.task_sensors <= kCoopMaxSensors) {
      const unsigned upd = (unsigned)a.task_sensors;   // one update CTA per sensor
      if (a.h_cov || a.h_vis_est || a.h_tr_pos)
        return launch_pdl(ukf_finish_quad_kernel<true, true>, dim3(upd + ctas), dim3(kQuadThreads), 0, stream, true, a);
      return launch_pdl(ukf_finish_quad_kernel<true, false>, dim3(upd + ctas), dim3(kQuadThreads), 0, stream, true, a);
    }
    return launch_pdl(ukf_finish_quad_kernel<false, false>, dim3(ctas), dim3(kQuadThreads), 0, stream, true, a);
  }
  // many measured targets per step: a tasking mask without actions (any number may be set), or at least one
  // sensor per 16 targets (a batch of environments)
  const bool update_heavy = (a.tasked != nullptr && a.task_of == nullptr) || a.task_sensors * 16 >= a.n;
  if (mapping == 3) return update_heavy ? launch_finish_tpb<64, 4>(a, stream) : launch_finish_tpb<64, 6>(a, stream);
  if (update_heavy && a.variant == 0) return launch_finish_tpb<64, 4>(a, stream);   // also above 128 k targets
  return launch_finish_tpb<128, 3>(a, stream);
}

}  // namespace pc
