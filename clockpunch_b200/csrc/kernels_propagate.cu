// K1 / K2: batched agent propagation, one thread per state vector.
// Layout: SoA (6, ld) FP64 -> each component row is read/written fully coalesced.
#include <cstdlib>
#include "propagate.cuh"

namespace pc {

template <bool J2>
__global__ void __launch_bounds__(128, 4)
propagate_agents_kernel(const double* __restrict__ x_in, double* __restrict__ x_out,
                        const uint8_t* __restrict__ kind, int default_kind, int64_t n, int64_t ld,
                        double dt, int substeps, double cs, double sn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[3], v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    r[c] = x_in[(int64_t)c * ld + i];
    v[c] = x_in[(int64_t)(c + 3) * ld + i];
  }
  const int k = kind ? (int)kind[i] : default_kind;
  if (k == PC_KIND_TERRESTRIAL) {
    rotate_z(r, v, cs, sn);
  } else {
    const int nsub = substeps > 0 ? substeps : auto_substeps(r, v, dt, nullptr);
    propagate_state<J2>(r, v, dt, nsub);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    x_out[(int64_t)c * ld + i] = r[c];
    x_out[(int64_t)(c + 3) * ld + i] = v[c];
  }
}

cudaError_t launch_propagate_agents(const double* x_in, double* x_out, const uint8_t* kind,
                                    int default_kind, int64_t n, int64_t ld, double t0, double tf,
                                    int substeps, int force_model, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const double dt = tf - t0;
  const double ang = kOmegaEarth * dt;
  const double cs = cos(ang), sn = sin(ang);
  const int threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (force_model == PC_FORCE_TWOBODY_J2)
    propagate_agents_kernel<true><<<blocks, threads, 0, stream>>>(x_in, x_out, kind, default_kind, n,
                                                                  ld, dt, substeps, cs, sn);
  else
    propagate_agents_kernel<false><<<blocks, threads, 0, stream>>>(x_in, x_out, kind, default_kind,
                                                                   n, ld, dt, substeps, cs, sn);
  return cudaGetLastError();
}

// K_b of the UKF pipeline: every state of the sigma buffer (14 blocks of (6, ld): 13 sigma points
// + the truth state per target) advanced in place, one thread per state, blockIdx.y = block.
// All 13 sigma points of a target must take the SAME number of substeps (their differences are
// multiplied by weights of order 1e5..1e6), so blocks 0..12 derive the count from one orbit rate
// per target, stored in sig_aux[n + t] when the sigma points were generated.
//
// Substep counts differ between neighbouring targets (a mixed LEO/MEO/GEO catalog, estimates of
// differing quality), and a warp runs as long as its slowest lane.  So each CTA first does a
// stable counting sort of its 128 targets by substep count (ballot + popc, one smem round trip)
// and every thread then takes the target at its sorted position: warps become homogeneous
// (at most one mixed warp per count boundary) while all accesses stay inside the CTA's own
// 128-target window of each component row, so DRAM traffic is unchanged.
#ifdef PC_STAMPS_HEADER   // variant builds only (tools/build_variant.sh); the product library has no stamps
#define PC_STAMPS_PROP
#include PC_STAMPS_HEADER
#else
#define PC_PROP_STAMP(i, op)
#define PC_PROP_CTA_ENTRY()
#define PC_PROP_CTA_LOADED(h, r0, v0, ns)
#define PC_PROP_CTA_EXIT()
#endif
constexpr int kPropThreads = 128;
constexpr int kSortBuckets = 6;  // 0 = no work (past n), 1..4 = that many substeps, 5 = more

// One state per thread, 128-target windows: used for catalogs of at most a few waves (see
// launch_propagate_sigma), where CTAs that exit early matter more than balance inside a CTA.
template <bool J2>
__global__ void __launch_bounds__(kPropThreads, 4)
propagate_sigma_single_kernel(double* __restrict__ sb, int64_t n, int64_t ld, double dt, int substeps,
                       uint32_t* __restrict__ sig_aux) {
  __shared__ int s_wcnt[kPropThreads / 32][kSortBuckets];
  __shared__ uint16_t s_perm[kPropThreads];
  __shared__ int s_nsub[kPropThreads];
  const int tid = threadIdx.x;
  // Programmatic dependent launch: this kernel needs nothing from its stream predecessor (the
  // sensor / tasking kernel), so it runs beside it; its successor (the per-target stage) may be
  // scheduled as soon as every CTA of this grid is resident.  One thread of the grid does wait for
  // the predecessor, so that "this grid is done" implies "the sensor kernel is done" for the
  // successor, which needs both -- thread 0 of the LAST CTA, AFTER its own work: a CTA that waited
  // first and worked afterwards ended late whenever the sensor kernel is long (it carries the
  // multi-GPU exchange) and stretched this kernel by that much (4.7 us per step at 2 GPUs).
  pdl_trigger();
  const int64_t t0 = (int64_t)blockIdx.x * kPropThreads;
  const int i = blockIdx.y;
  double* blk = sb + (int64_t)i * 6 * ld;
  int slot = tid, nsub = substeps;
  if (substeps <= 0) {
    // blocks 0..12 use the orbit rate stored with the sigma points (computed from the centre point
    // when they were generated); the truth state (13) uses its own
    const int64_t t = t0 + tid;
    int mine = 0;
    if (t < n) {
      float om;
      if (i == 13) {
        const double r[3] = {blk[t], blk[ld + t], blk[2 * ld + t]};
        const double v[3] = {blk[3 * ld + t], blk[4 * ld + t], blk[5 * ld + t]};
        om = orbit_rate(r, v, dt);
      } else {
        om = packed_rate_for_step(sig_aux[n + t], dt);
      }
      bool cl = false;
      mine = substeps_from_rate(om, dt, &cl);
      if (i == 0 && cl) sig_aux[t] |= PC_FLAG_SUBSTEP_CLAMP;
    }
    const int key = mine < kSortBuckets - 1 ? mine : kSortBuckets - 1;
    const int w = tid >> 5, lane = tid & 31;
    int rank = 0;
#pragma unroll
    for (int b = 0; b < kSortBuckets; ++b) {
      const unsigned m = __ballot_sync(0xffffffffu, key == b);
      if (key == b) rank = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) s_wcnt[w][b] = __popc(m);
    }
    s_nsub[tid] = mine;
    __syncthreads();
    int pos = rank;
#pragma unroll
    for (int b = 0; b < kSortBuckets; ++b)
#pragma unroll
      for (int ww = 0; ww < kPropThreads / 32; ++ww)
        if (b < key || (b == key && ww < w)) pos += s_wcnt[ww][b];
    s_perm[pos] = (uint16_t)tid;
    __syncthreads();
    slot = s_perm[tid];
    nsub = s_nsub[slot];
  } else if (t0 + tid >= n) {
    nsub = 0;
  }
  if (nsub != 0) {  // 0: past the end of the catalog
    double* base = blk + t0 + slot;
    double r[3], v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = base[(int64_t)c * ld];
      v[c] = base[(int64_t)(c + 3) * ld];
    }
    propagate_state<J2>(r, v, dt, nsub);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      base[(int64_t)c * ld] = r[c];
      base[(int64_t)(c + 3) * ld] = v[c];
    }
  }
  if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && threadIdx.x == 0) pdl_wait();
}

constexpr int kPropWindow = 2 * kPropThreads;   // targets per CTA: every thread advances two states
constexpr int64_t kFoldMinTargets = 16 * 1024;  // below: one state per thread (see launch_propagate_sigma)

template <bool J2>
__global__ void __launch_bounds__(kPropThreads, 4)
propagate_sigma_kernel(double* __restrict__ sb, int64_t n, int64_t ld, double dt, int substeps,
                       uint32_t* __restrict__ sig_aux) {
  __shared__ int s_cnt[2][kPropThreads / 32][kSortBuckets];
  __shared__ uint16_t s_perm[kPropWindow];
  __shared__ uint8_t s_nsub[kPropWindow];
  __shared__ double s_x[6][kPropWindow];   // the window's states, copied asynchronously while the sort runs
  const int tid = threadIdx.x;
  // Programmatic dependent launch: this kernel needs nothing from its stream predecessor (the
  // sensor / tasking kernel), so it runs beside it; its successor (the per-target stage) may be
  // scheduled as soon as every CTA of this grid is resident.  One thread of the grid does wait for
  // the predecessor, so that "this grid is done" implies "the sensor kernel is done" for the
  // successor, which needs both -- thread 0 of the LAST CTA, AFTER its own work: a CTA that waited
  // first and worked afterwards ended late whenever the sensor kernel is long (it carries the
  // multi-GPU exchange) and stretched this kernel by that much (4.7 us per step at 2 GPUs).
  pdl_trigger();
  PC_PROP_STAMP(0, atomicMin);
  PC_PROP_CTA_ENTRY();
  const int64_t t0 = (int64_t)blockIdx.x * kPropWindow;
  const int i = blockIdx.y;
  double* blk = sb + (int64_t)i * 6 * ld;
  int slot[2] = {tid, kPropThreads + tid};
  int nsub[2] = {substeps, substeps};
  // Both states at this thread's own positions are requested first (global -> shared, no registers):
  // they travel while the rates are fetched and the window is sorted -- one trip to HBM instead of
  // two (rates, then the sorted slots' states) -- and change hands through shared memory afterwards.
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t t = t0 + h * kPropThreads + tid;
    if (t < n) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_x[c][h * kPropThreads + tid]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(blk + (int64_t)c * ld + t) : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (substeps <= 0) {
    // blocks 0..12 use the orbit rate stored with the sigma points (computed from the centre point
    // when they were generated); the truth state (13) uses its own
    int mine[2], key[2];
    if (i == 13) asm volatile("cp.async.wait_group 0;" ::: "memory");   // own copies: visible to this thread
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = t0 + h * kPropThreads + tid;
      mine[h] = 0;
      if (t < n) {
        float om;
        if (i == 13) {
          const int o = h * kPropThreads + tid;
          const double r[3] = {s_x[0][o], s_x[1][o], s_x[2][o]};
          const double v[3] = {s_x[3][o], s_x[4][o], s_x[5][o]};
          om = orbit_rate(r, v, dt);
        } else {
          om = packed_rate_for_step(sig_aux[n + t], dt);
        }
        bool cl = false;
        mine[h] = substeps_from_rate(om, dt, &cl);
        if (i == 0 && cl) sig_aux[t] |= PC_FLAG_SUBSTEP_CLAMP;
      }
      key[h] = mine[h] < kSortBuckets - 1 ? mine[h] : kSortBuckets - 1;
    }
    // Stable counting sort of the window's 256 states by substep count, HEAVIEST FIRST (ballot + popc,
    // one shared-memory round trip).  Thread j then takes sorted positions j and 255 - j: the j-th
    // heaviest state and the j-th lightest.  Neighbouring lanes get equal counts in both rounds (no
    // divergence inside a warp), and every thread of the CTA does about the same total work, so no
    // warp sits idle holding registers while the CTA's slowest one finishes (with one state per
    // thread the CTA lasted 3 substeps where the mean is 1.5: half the FP64 issue slots of a
    // 10^4-target catalog were empty).  All accesses stay inside the CTA's own window of each
    // component row, so DRAM traffic is unchanged.
    const int w = tid >> 5, lane = tid & 31;
    int rank[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int b = 0; b < kSortBuckets; ++b) {
        const unsigned m = __ballot_sync(0xffffffffu, key[h] == b);
        if (key[h] == b) rank[h] = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[h][w][b] = __popc(m);
      }
    s_nsub[tid] = (uint8_t)(mine[0] < 255 ? mine[0] : 255);
    s_nsub[kPropThreads + tid] = (uint8_t)(mine[1] < 255 ? mine[1] : 255);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    int pos[2] = {rank[0], rank[1]};
#pragma unroll
    for (int b = 0; b < kSortBuckets; ++b)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int ww = 0; ww < kPropThreads / 32; ++ww) {
          const int c = s_cnt[hh][ww][b];
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (b > key[h] || (b == key[h] && (hh < h || (hh == h && ww < w)))) pos[h] += c;
        }
    s_perm[pos[0]] = (uint16_t)tid;
    s_perm[pos[1]] = (uint16_t)(kPropThreads + tid);
    __syncthreads();
    slot[0] = s_perm[tid];
    slot[1] = s_perm[kPropWindow - 1 - tid];
    nsub[0] = s_nsub[slot[0]];
    nsub[1] = s_nsub[slot[1]];
    // counts above 254 (clamped orbit rates, extreme steps) are rare: take them from the rule again
    if (nsub[0] == 255 || nsub[1] == 255) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (nsub[h] == 255) {
          const int64_t t = t0 + slot[h];
          float om;
          if (i == 13) {
            const double r[3] = {s_x[0][slot[h]], s_x[1][slot[h]], s_x[2][slot[h]]};
            const double v[3] = {s_x[3][slot[h]], s_x[4][slot[h]], s_x[5][slot[h]]};
            om = orbit_rate(r, v, dt);
          } else {
            om = packed_rate_for_step(sig_aux[n + t], dt);
          }
          nsub[h] = substeps_from_rate(om, dt, nullptr);
        }
    }
  } else {
    if (t0 + tid >= n) nsub[0] = 0;
    if (t0 + kPropThreads + tid >= n) nsub[1] = 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    if (nsub[h] == 0) continue;  // past the end of the catalog
    double* base = blk + t0 + slot[h];
    double r[3], v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[c] = s_x[c][slot[h]];
      v[c] = s_x[c + 3][slot[h]];
    }
    PC_PROP_CTA_LOADED(h, r[0], v[0], nsub[0]);
    propagate_state<J2>(r, v, dt, nsub[h]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      base[(int64_t)c * ld] = r[c];
      base[(int64_t)(c + 3) * ld] = v[c];
    }
  }
  PC_PROP_STAMP(1, atomicMax);
  PC_PROP_CTA_EXIT();
  if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && tid == 0) pdl_wait();
}

cudaError_t launch_propagate_sigma(double* sb, int64_t n, int64_t ld, int have_truth, double t0,
                                   double tf, int substeps, int force_model, uint32_t* sig_flags,
                                   bool beside_predecessor, int mapping, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const unsigned blocks = have_truth ? 14 : 13;
  const bool j2 = force_model == PC_FORCE_TWOBODY_J2;
  // Two mappings.  Large catalogs: two states per thread out of a sorted 256-target window (every thread
  // of a CTA does about the same work: 93 % of the measured FP64 peak at 10^6 targets against 81 % with
  // one state per thread).  Catalogs of a wave or two of CTAs: one state per thread -- the grid is two
  // waves of short CTAs, the second of which frees SM slots one by one, and the per-target kernel
  // launched programmatically behind this one uses exactly that window to bring its update warps'
  // code into the instruction caches (36.5 vs 33.9 us per step at 10^4 targets).
  // (mapping: pc_option PC_OPT_PROPAGATE_KERNEL -- 1 / 2 pin the choice, 0 = the size rule)
  const bool fold = mapping ? mapping == 2 : n > kFoldMinTargets;
  if (!fold) {
    const dim3 grid((unsigned)((n + kPropThreads - 1) / kPropThreads), blocks);
    if (j2)
      return launch_pdl(propagate_sigma_single_kernel<true>, grid, dim3(kPropThreads), 0, stream, beside_predecessor,
                        sb, n, ld, tf - t0, substeps, sig_flags);
    return launch_pdl(propagate_sigma_single_kernel<false>, grid, dim3(kPropThreads), 0, stream, beside_predecessor,
                      sb, n, ld, tf - t0, substeps, sig_flags);
  }
  const dim3 grid((unsigned)((n + kPropWindow - 1) / kPropWindow), blocks);
  if (j2)
    return launch_pdl(propagate_sigma_kernel<true>, grid, dim3(kPropThreads), 0, stream, beside_predecessor, sb, n,
                      ld, tf - t0, substeps, sig_flags);
  return launch_pdl(propagate_sigma_kernel<false>, grid, dim3(kPropThreads), 0, stream, beside_predecessor, sb, n, ld,
                    tf - t0, substeps, sig_flags);
}

// ---- frame transforms / initial-condition generation (transforms.py:34-292) -----------

// r1 = R3(angle) r0 ; v1 = R3(angle) v0 + (0,0,omega) x r1   (_transport_theorem_posvel,
// transforms.py:143-178; ecef2eci: angle = +w JD, omega = +w; eci2ecef: both negated)
__global__ void __launch_bounds__(128)
frame_change_kernel(const double* __restrict__ x_in, double* __restrict__ x_out, int64_t n,
                    int64_t ld, double cs, double sn, double omega) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = x_in[i], y = x_in[ld + i], z = x_in[2 * ld + i];
  const double vx = x_in[3 * ld + i], vy = x_in[4 * ld + i], vz = x_in[5 * ld + i];
  const double rx = cs * x - sn * y, ry = sn * x + cs * y;
  x_out[i] = rx;
  x_out[ld + i] = ry;
  x_out[2 * ld + i] = z;
  x_out[3 * ld + i] = (cs * vx - sn * vy) - omega * ry;
  x_out[4 * ld + i] = (sn * vx + cs * vy) + omega * rx;
  x_out[5 * ld + i] = vz;
}

// coe2eci (transforms.py:240-292): COE -> PQW -> ECI, rot3(-raan) rot1(-inc) rot3(-argp).
__global__ void __launch_bounds__(128)
coe2eci_kernel(const double* __restrict__ coe, double* __restrict__ x_out, int64_t n, int64_t ld,
               double mu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sma = coe[i], ecc = coe[ld + i], inc = coe[2 * ld + i];
  const double raan = coe[3 * ld + i], argp = coe[4 * ld + i], ta = coe[5 * ld + i];
  double sa, ca, so, co, si, ci, sw, cw;
  sincos(ta, &sa, &ca);
  sincos(raan, &so, &co);
  sincos(inc, &si, &ci);
  sincos(argp, &sw, &cw);
  const double p = sma * (1.0 - ecc * ecc);
  const double rr = p / (1.0 + ecc * ca);
  const double vv = sqrt(mu / p);
  const double rp = rr * ca, rq = rr * sa;
  const double vp = -vv * sa, vq = vv * (ecc + ca);
  // columns 1 and 2 of rot3(-raan) rot1(-inc) rot3(-argp) with the reference's rot1/rot3
  // (transforms.py:186-237: active-rotation matrices, so this is NOT the textbook PQW->ECI
  // matrix but its counterpart with the three angles negated -- the reference is the spec).
  const double Px = co * cw - so * ci * sw, Py = -so * cw - co * ci * sw, Pz = si * sw;
  const double Qx = co * sw + so * ci * cw, Qy = -so * sw + co * ci * cw, Qz = -si * cw;
  x_out[i] = Px * rp + Qx * rq;
  x_out[ld + i] = Py * rp + Qy * rq;
  x_out[2 * ld + i] = Pz * rp + Qz * rq;
  x_out[3 * ld + i] = Px * vp + Qx * vq;
  x_out[4 * ld + i] = Py * vp + Qy * vq;
  x_out[5 * ld + i] = Pz * vp + Qz * vq;
}

// lla2ecef (transforms.py:34-52, spherical Earth, zero ECEF velocity) then ecef2eci(., JD).
__global__ void __launch_bounds__(128)
lla2eci_kernel(const double* __restrict__ lla, double* __restrict__ x_out, int64_t n, int64_t ld,
               double cs, double sn, double omega) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lat = lla[i], lon = lla[ld + i], alt = lla[2 * ld + i];
  double slat, clat, slon, clon;
  sincos(lat, &slat, &clat);
  sincos(lon, &slon, &clon);
  const double rd = (kRE + alt) * clat;
  const double x = rd * clon, y = rd * slon, z = (kRE + alt) * slat;
  const double rx = cs * x - sn * y, ry = sn * x + cs * y;
  x_out[i] = rx;
  x_out[ld + i] = ry;
  x_out[2 * ld + i] = z;
  x_out[3 * ld + i] = -omega * ry;
  x_out[4 * ld + i] = omega * rx;
  x_out[5 * ld + i] = 0.0;
}

cudaError_t launch_frame_change(const double* x_in, double* x_out, int64_t n, int64_t ld,
                                double angle, double omega, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  frame_change_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(x_in, x_out, n, ld, cos(angle),
                                                                      sin(angle), omega);
  return cudaGetLastError();
}
cudaError_t launch_coe2eci(const double* coe, double* x_out, int64_t n, int64_t ld, double mu,
                           cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  coe2eci_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(coe, x_out, n, ld, mu);
  return cudaGetLastError();
}
cudaError_t launch_lla2eci(const double* lla, double* x_out, int64_t n, int64_t ld, double jd,
                           cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const double ang = kOmegaEarth * jd;
  lla2eci_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(lla, x_out, n, ld, cos(ang),
                                                                 sin(ang), kOmegaEarth);
  return cudaGetLastError();
}

}  // namespace pc
