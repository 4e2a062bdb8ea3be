// K3: Earth-occlusion visibility between every (sensor, target) pair.
//
// Reference: calcVisMap (punchclock/common/visibility.py:56-116) loops over all pairs and
// calls satvis.isVis / satvis.visibilityFunc (satvis==0.2.4, not vendored: parity unpinned,
// restated in oracle/visibility.py):
//     V = acos(RE/|r1|) + acos(RE/|r2|) - acos(r1.r2 / (|r1||r2|)),   visible <=> V > 0.
// Binary mode here uses the equivalent acos-free predicate
//     r1.r2 > RE^2 - sqrt((|r1|^2 - RE^2)(|r2|^2 - RE^2))
// (both sides are cos() of angles in [0, pi], where cos is monotone), which differs from
// the acos form only inside |V| <~ 1e-9 rad of grazing.
//
// Mapping: sensors (position + horizon term q = sqrt(|r|^2 - RE^2)) are staged once per
// CTA in shared memory; one thread owns one 32-bit output word = (target i, 32 sensors) of
// the truth map and, when a second target set (the filter estimates) is given, of the
// estimate map, so N * ceil(M/32) threads fill the machine even for small catalogs and the
// packed words are written fully coalesced.  (The fused step kernel in kernels_ukf.cu packs
// the same words with warp ballots, lanes = sensors.)
#include "vis.cuh"

namespace pc {

constexpr int kVisThreads = 128;
constexpr int kSensTile = 1024;  // sensors staged per pass: 1024 * 32 B = 32 KB
// 33 slots per 32-sensor word: threads of a warp own different words, and without the pad their
// 128-bit LDS addresses are 1 KB apart (same banks, 4- to 32-way conflicts)
constexpr int kSensRow = 33;
constexpr int kSensSlots = (kSensTile / 32) * kSensRow;

// One thread per output word: (target i, sensors 32w .. 32w+31), both maps.
template <bool TWO>
__global__ void __launch_bounds__(kVisThreads)
vismap_packed_kernel(const double* __restrict__ sens, int64_t M, int64_t ld_s,
                     const uint8_t* __restrict__ sens_kind, double cs, double sn,
                     const double* __restrict__ ta, const double* __restrict__ tb, int64_t N,
                     int64_t ld_t, double RE, double tol, uint32_t* __restrict__ out_a,
                     uint32_t* __restrict__ out_b, int64_t wpr) {
  __shared__ double4 tile[kSensSlots];
  const int64_t idx = (int64_t)blockIdx.x * kVisThreads + threadIdx.x;
  const bool live = idx < N * wpr;
  const int64_t i = live ? idx / wpr : 0;
  const int64_t w = live ? idx - i * wpr : 0;
  const double RE2 = RE * RE;

  const double ax = ta[i], ay = ta[ld_t + i], az = ta[2 * ld_t + i];
  const double aq = horizon_q(ax, ay, az, RE, tol);
  double bx = 0, by = 0, bz = 0, bq = 0;
  if (TWO) {
    bx = tb[i];
    by = tb[ld_t + i];
    bz = tb[2 * ld_t + i];
    bq = horizon_q(bx, by, bz, RE, tol);
  }

  uint32_t wa = 0, wb = 0;
  for (int64_t j0 = 0; j0 < M; j0 += kSensTile) {
    const int cnt = (int)min((int64_t)kSensTile, M - j0);
    if (j0) __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += kVisThreads) {
      double x = sens[j0 + j], y = sens[ld_s + j0 + j];
      const double z = sens[2 * ld_s + j0 + j];
      if (sens_kind && sens_kind[j0 + j] == PC_KIND_TERRESTRIAL) {  // fused K2: Earth-fixed rotation
        const double xr = cs * x - sn * y;
        y = sn * x + cs * y;
        x = xr;
      }
      tile[(j >> 5) * kSensRow + (j & 31)] = make_double4(x, y, z, horizon_q(x, y, z, RE, tol));
    }
    __syncthreads();
    const int64_t s0 = w * 32 - j0;  // first sensor of this word, relative to the tile
    if (live && s0 >= 0 && s0 < cnt) {
      const int lim = min(32, cnt - (int)s0);
      const double4* __restrict__ row = tile + (s0 >> 5) * kSensRow;
#pragma unroll 4
      for (int b = 0; b < lim; ++b) {
        const double4 s = row[b];
        const double da = fma(s.z, az, fma(s.y, ay, s.x * ax));
        wa |= (uint32_t)(da > fma(-s.w, aq, RE2)) << b;
        if (TWO) {
          const double db = fma(s.z, bz, fma(s.y, by, s.x * bx));
          wb |= (uint32_t)(db > fma(-s.w, bq, RE2)) << b;
        }
      }
    }
  }
  if (live) {
    out_a[idx] = wa;
    if (TWO) out_b[idx] = wb;
  }
}

// Continuous visibility value (calcVisMap(binary=False) -> visibilityFunc(...)[0]).
__global__ void __launch_bounds__(kVisThreads)
vismap_value_kernel(const double* __restrict__ sens, int64_t M, int64_t ld_s,
                    const double* __restrict__ targ, int64_t N, int64_t ld_t, double RE, double tol,
                    double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * kVisThreads + threadIdx.x;
  if (idx >= N * M) return;
  const int64_t i = idx / M, j = idx - i * M;
  const double sx = sens[j], sy = sens[ld_s + j], sz = sens[2 * ld_s + j];
  const double tx = targ[i], ty = targ[ld_t + i], tz = targ[2 * ld_t + i];
  double m1 = sqrt(fma(sz, sz, fma(sy, sy, sx * sx)));
  double m2 = sqrt(fma(tz, tz, fma(ty, ty, tx * tx)));
  if (m1 < RE && RE - m1 <= tol) m1 = RE;
  if (m2 < RE && RE - m2 <= tol) m2 = RE;
  double c = fma(sz, tz, fma(sy, ty, sx * tx)) / (m1 * m2);
  c = fmin(1.0, fmax(-1.0, c));
  out[idx] = acos(RE / m1) + acos(RE / m2) - acos(c);
}

// Visibility value AND its time derivative (calcVisMapAndDerivative, visibility.py:119-180 ->
// satvis.calcVisAndDerVis, restated in oracle/visibility.py; parity unpinned):
//   dV/dt = da1/dt + da2/dt - dphi/dt,
//   da/dt = RE |r|' / (|r|^2 sqrt(1 - (RE/|r|)^2)),
//   dphi/dt = -[(r1'.r2 + r1.r2') / (|r1||r2|) - cos(phi) (|r1|'/|r1| + |r2|'/|r2|)] / sin(phi).
// The radial rate |r|' is r.v/|r|; the reference reaches the same number through the orbit-element
// chain of orbits.py:39-69 (max difference 5e-14 km/s over tests/golden/orbits.npz).
// Aligned position vectors give +-Inf / NaN exactly as the reference documents (visibility.py:141-143).
__global__ void __launch_bounds__(kVisThreads)
vismap_der_kernel(const double* __restrict__ sens, int64_t M, int64_t ld_s,
                  const double* __restrict__ targ, int64_t N, int64_t ld_t, double RE, double tol,
                  double* __restrict__ out_v, double* __restrict__ out_dv) {
  const int64_t idx = (int64_t)blockIdx.x * kVisThreads + threadIdx.x;
  if (idx >= N * M) return;
  const int64_t i = idx / M, j = idx - i * M;
  double s[6], t[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    s[c] = sens[(int64_t)c * ld_s + j];
    t[c] = targ[(int64_t)c * ld_t + i];
  }
  const double n1 = sqrt(fma(s[2], s[2], fma(s[1], s[1], s[0] * s[0])));
  const double n2 = sqrt(fma(t[2], t[2], fma(t[1], t[1], t[0] * t[0])));
  double m1 = n1, m2 = n2;  // magnitudes as the visibility function sees them (surface clamp)
  if (m1 < RE && RE - m1 <= tol) m1 = RE;
  if (m2 < RE && RE - m2 <= tol) m2 = RE;
  const double dot12 = fma(s[2], t[2], fma(s[1], t[1], s[0] * t[0]));
  double c = dot12 / (m1 * m2);
  c = fmin(1.0, fmax(-1.0, c));
  out_v[idx] = acos(RE / m1) + acos(RE / m2) - acos(c);
  const double rd1 = fma(s[2], s[5], fma(s[1], s[4], s[0] * s[3])) / n1;
  const double rd2 = fma(t[2], t[5], fma(t[1], t[4], t[0] * t[3])) / n2;
  const double q1 = RE / n1, q2 = RE / n2;
  const double a1dot = RE * rd1 / (n1 * n1 * sqrt(1.0 - q1 * q1));
  const double a2dot = RE * rd2 / (n2 * n2 * sqrt(1.0 - q2 * q2));
  const double cu = dot12 / (n1 * n2);
  const double cross = fma(s[5], t[2], fma(s[4], t[1], s[3] * t[0])) + fma(s[2], t[5], fma(s[1], t[4], s[0] * t[3]));
  const double cdot = cross / (n1 * n2) - cu * (rd1 / n1 + rd2 / n2);
  const double phidot = -cdot / sqrt(1.0 - cu * cu);
  out_dv[idx] = a1dot + a2dot - phidot;
}

cudaError_t launch_vismap_der(const double* sens, int64_t M, int64_t ld_s, const double* targ,
                              int64_t N, int64_t ld_t, double RE, double tol, double* out_v,
                              double* out_dv, cudaStream_t stream) {
  if (N <= 0 || M <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((N * M + kVisThreads - 1) / kVisThreads);
  vismap_der_kernel<<<blocks, kVisThreads, 0, stream>>>(sens, M, ld_s, targ, N, ld_t, RE, tol, out_v, out_dv);
  return cudaGetLastError();
}

// Block-diagonal map of a batch of E environments (reset path of the vectorised env): target i
// belongs to environment i / Ne and is tested against that environment's Me sensors only.  One
// thread per output word; sensors are read straight from global memory (L1-resident: Me is small).
template <bool TWO>
__global__ void __launch_bounds__(kVisThreads)
vismap_envs_kernel(const double* __restrict__ sens, int64_t ld_s, const double* __restrict__ ta,
                   const double* __restrict__ tb, int64_t N, int64_t ld_t, int64_t Ne, int64_t Me,
                   double RE, double tol, uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_b,
                   int64_t wpr) {
  const int64_t idx = (int64_t)blockIdx.x * kVisThreads + threadIdx.x;
  if (idx >= N * wpr) return;
  const int64_t i = idx / wpr, w = idx - i * wpr, e = i / Ne;
  const double RE2 = RE * RE;
  const double ax = ta[i], ay = ta[ld_t + i], az = ta[2 * ld_t + i];
  const double aq = horizon_q(ax, ay, az, RE, tol);
  double bx = 0, by = 0, bz = 0, bq = 0;
  if (TWO) {
    bx = tb[i];
    by = tb[ld_t + i];
    bz = tb[2 * ld_t + i];
    bq = horizon_q(bx, by, bz, RE, tol);
  }
  uint32_t wa = 0, wb = 0;
  const int lim = (int)min((int64_t)32, Me - w * 32);
  for (int b = 0; b < lim; ++b) {
    const int64_t j = e * Me + w * 32 + b;
    const double x = sens[j], y = sens[ld_s + j], z = sens[2 * ld_s + j];
    const double4 s = make_double4(x, y, z, horizon_q(x, y, z, RE, tol));
    if (sees(s, ax, ay, az, aq, RE2)) wa |= 1u << b;
    if (TWO && sees(s, bx, by, bz, bq, RE2)) wb |= 1u << b;
  }
  out_a[idx] = wa;
  if (TWO) out_b[idx] = wb;
}

cudaError_t launch_vismap_envs(const double* sens, int64_t ld_s, const double* ta, const double* tb,
                               int64_t N, int64_t ld_t, int64_t Ne, int64_t Me, double RE, double tol,
                               uint32_t* out_a, uint32_t* out_b, cudaStream_t stream) {
  const int64_t wpr = (Me + 31) / 32;
  if (N <= 0 || Me <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((N * wpr + kVisThreads - 1) / kVisThreads);
  if (tb)
    vismap_envs_kernel<true><<<blocks, kVisThreads, 0, stream>>>(sens, ld_s, ta, tb, N, ld_t, Ne, Me, RE, tol, out_a, out_b, wpr);
  else
    vismap_envs_kernel<false><<<blocks, kVisThreads, 0, stream>>>(sens, ld_s, ta, nullptr, N, ld_t, Ne, Me, RE, tol, out_a, nullptr, wpr);
  return cudaGetLastError();
}

// Expand packed words to the reference's (N, M) int64 map (visibility.py:112-114).
__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint32_t* __restrict__ words, int64_t N, int64_t M, int64_t wpr,
                   int64_t* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= N * M) return;
  const int64_t i = idx / M, j = idx - i * M;
  out[idx] = (words[i * wpr + (j >> 5)] >> (j & 31)) & 1u;
}

cudaError_t launch_unpack_bits(const uint32_t* words, int64_t N, int64_t M, int64_t wpr,
                               int64_t* out, cudaStream_t stream) {
  const int64_t total = N * M;
  if (total <= 0) return cudaSuccess;
  unpack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(words, N, M, wpr, out);
  return cudaGetLastError();
}

cudaError_t launch_vismap_packed(const double* sens, int64_t M, int64_t ld_s,
                                 const uint8_t* sens_kind, double rot_angle, const double* ta,
                                 const double* tb, int64_t N, int64_t ld_t, double RE, double tol,
                                 uint32_t* out_a, uint32_t* out_b, int64_t wpr,
                                 cudaStream_t stream) {
  if (N <= 0 || M <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((N * wpr + kVisThreads - 1) / kVisThreads);
  const double cs = cos(rot_angle), sn = sin(rot_angle);
  if (tb)
    vismap_packed_kernel<true><<<blocks, kVisThreads, 0, stream>>>(
        sens, M, ld_s, sens_kind, cs, sn, ta, tb, N, ld_t, RE, tol, out_a, out_b, wpr);
  else
    vismap_packed_kernel<false><<<blocks, kVisThreads, 0, stream>>>(
        sens, M, ld_s, sens_kind, cs, sn, ta, nullptr, N, ld_t, RE, tol, out_a, nullptr, wpr);
  return cudaGetLastError();
}

cudaError_t launch_vismap_value(const double* sens, int64_t M, int64_t ld_s, const double* targ,
                                int64_t N, int64_t ld_t, double RE, double tol, double* out,
                                cudaStream_t stream) {
  if (N <= 0 || M <= 0) return cudaSuccess;
  const int64_t total = N * M;
  const unsigned blocks = (unsigned)((total + kVisThreads - 1) / kVisThreads);
  vismap_value_kernel<<<blocks, kVisThreads, 0, stream>>>(sens, M, ld_s, targ, N, ld_t, RE, tol,
                                                         out);
  return cudaGetLastError();
}

}  // namespace pc
