// Single-molecule trial in ONE launch: MoleculeInter (src/CalculateEnergy.cpp:581-686)
// + Ewald::MolReciprocal / SwapDestRecip / SwapSourceRecip (src/Ewald.cpp:409-531,
// :657-710) + SwapCorrection / SwapSelf (:1311-1391).
//
// These calls are latency-bound (a few thousand pair candidates and len*nk sincos),
// so the design removes round trips instead of arithmetic:
//   * the trial molecule travels in the kernel parameters (no H2D copy),
//   * probe blocks (pair part, each probe split over kProbeSplit blocks) and k blocks
//     (reciprocal part) run side by side in one grid,
//   * the last block to finish (ticket counter) sums every partial in a fixed order
//     and writes the scalars straight into mapped pinned host memory, then publishes a
//     sequence number the host spins on (no D2H copy, no stream synchronisation).
#pragma once
#include "pair.cuh"
#include "recip.cuh"

namespace gb {

constexpr int kTrialMaxAtoms = 32;  // larger molecules use the staged path
constexpr int kProbeSplit = 4;      // blocks per probe: 4 * 8 warps >= 27 neighbour cells

struct TrialArgs {
  int len, mode;  // mode as k_mol_recip: 0 move, 1 insert, 2 delete
  int excludeMol, nk;
  int nProbeBlocks;  // 2 * len * kProbeSplit (MoleculeInter part) or 0
  int doCorrection;  // SwapCorrection + SwapSelf of the new coordinates
  unsigned long long seq;
  int kind[kTrialMaxAtoms];
  double q[kTrialMaxAtoms];   // true charges (pair part)
  double qr[kTrialMaxAtoms];  // charge * lambdaCoef of the molecule (Ewald part)
  double nx[kTrialMaxAtoms], ny[kTrialMaxAtoms], nz[kTrialMaxAtoms];
  double ox[kTrialMaxAtoms], oy[kTrialMaxAtoms], oz[kTrialMaxAtoms];
};

// hostOut: {lj, real, overlap, recipNew, correction, self}
template <int VDW, int MODE = MODE_ENERGY>
__global__ void __launch_bounds__(kPairThreads)
    k_mol_trial(BoxParams p, CellGrid g, const __grid_constant__ TrialArgs a,
                const int *__restrict__ cellStart, const double *__restrict__ sx,
                const double *__restrict__ sy, const double *__restrict__ sz,
                const double *__restrict__ sq, const int2 *__restrict__ skm,
                const double *__restrict__ kx, const double *__restrict__ ky,
                const double *__restrict__ kz, const double *__restrict__ prefact,
                const double *__restrict__ sumRref, const double *__restrict__ sumIref,
                double *__restrict__ sumRnew, double *__restrict__ sumInew,
                double *__restrict__ probePart, double *__restrict__ blockEnergy,
                unsigned *__restrict__ ticket, double *hostOut,
                volatile unsigned long long *hostFlag) {
  __shared__ JRange ranges[27];
  __shared__ WarpQueue queues[kPairWarps];
  __shared__ double red[2][kPairWarps];
  __shared__ int ovl[kPairWarps];
  __shared__ int isLast;
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if ((int)blockIdx.x < a.nProbeBlocks) {
    // ---- pair part: probe t = 2*atom + (0 old | 1 new), ranges split over blocks
    const int t = blockIdx.x / kProbeSplit, sub = blockIdx.x % kProbeSplit;
    const int at = t >> 1;
    const bool isNew = t & 1;
    const double px = isNew ? a.nx[at] : a.ox[at], py = isNew ? a.ny[at] : a.oy[at],
                 pz = isNew ? a.nz[at] : a.oz[at];
    const int nRangesSh =
        build_ranges(g, p, position_to_cell(g, px, py, pz), false, cellStart, ranges);
    __syncthreads();
    PairAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0};
    JArrays ja = {sx, sy, sz, sq, skm, 0u, 0u, 0u, 0u, 0u};
    warp_probe<VDW, MODE, SWEEP_PROBE, false, true>(
        p, g.generic, px, py, pz, a.kind[at], a.q[at], a.excludeMol, -1, -1,
        isNew ? 1.0 : -1.0, isNew, ranges, nRangesSh, kPairWarps * kProbeSplit,
        sub * kPairWarps + warp, ja, smem_u32(&queues[warp]), acc);
    double e0 = warp_sum(acc.lj), e1 = warp_sum(acc.real);
    int ov = __any_sync(0xffffffffu, acc.overlap);
    if (lane == 0) {
      red[0][warp] = e0;
      red[1][warp] = e1;
      ovl[warp] = ov;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s0 = 0.0, s1 = 0.0;
      int o = 0;
      for (int w = 0; w < kPairWarps; ++w) {
        s0 += red[0][w];
        s1 += red[1][w];
        o |= ovl[w];
      }
      probePart[3 * blockIdx.x + 0] = s0;
      probePart[3 * blockIdx.x + 1] = s1;
      probePart[3 * blockIdx.x + 2] = (double)o;
    }
  } else {
    // ---- reciprocal part: one k per thread
    const int k = (blockIdx.x - a.nProbeBlocks) * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (k < a.nk) {
      const double kxv = kx[k], kyv = ky[k], kzv = kz[k];
      double rn = 0.0, in = 0.0, ro = 0.0, io = 0.0;
      for (int at = 0; at < a.len; ++at) {
        const double q = a.qr[at];
        if (fabs(a.q[at]) < 0.000000001) continue;  // particleHasNoCharge
        double dn = __dadd_rn(__dadd_rn(__dmul_rn(a.nx[at], kxv), __dmul_rn(a.ny[at], kyv)),
                              __dmul_rn(a.nz[at], kzv));
        double s, c;
        sincos(dn, &s, &c);
        rn += q * c;
        in += q * s;
        if (a.mode == 0) {
          double d0 = __dadd_rn(__dadd_rn(__dmul_rn(a.ox[at], kxv), __dmul_rn(a.oy[at], kyv)),
                                __dmul_rn(a.oz[at], kzv));
          sincos(d0, &s, &c);
          ro += q * c;
          io += q * s;
        }
      }
      double r, i;
      if (a.mode == 0) {
        r = sumRref[k] + (rn - ro);
        i = sumIref[k] + (in - io);
      } else if (a.mode == 1) {
        r = sumRref[k] + rn;
        i = sumIref[k] + in;
      } else {
        r = sumRref[k] - rn;
        i = sumIref[k] - in;
      }
      sumRnew[k] = r;
      sumInew[k] = i;
      e = (r * r + i * i) * prefact[k];
    }
    double s = block_sum(e, scratch);
    if (threadIdx.x == 0) blockEnergy[blockIdx.x - a.nProbeBlocks] = s;
  }

  // ---- last block finalises (fixed summation order => deterministic)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    isLast = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!isLast) return;
  __threadfence();
  const int nProbes = a.nProbeBlocks / kProbeSplit;
  double lj = 0.0, re = 0.0, ov = 0.0;
  for (int t = threadIdx.x; t < nProbes; t += blockDim.x)
    for (int s = 0; s < kProbeSplit; ++s) {
      const double *pp = probePart + 3 * (t * kProbeSplit + s);
      lj += __ldcg(pp);
      re += __ldcg(pp + 1);
      ov += __ldcg(pp + 2);
    }
  double rc = 0.0;
  const int nRecipBlocks = gridDim.x - a.nProbeBlocks;
  for (int t = threadIdx.x; t < nRecipBlocks; t += blockDim.x) rc += __ldcg(blockEnergy + t);
  double corr = 0.0, self = 0.0;
  if (a.doCorrection) {
    const int nPairs = a.len * (a.len - 1) / 2;
    for (int t = threadIdx.x; t < nPairs; t += blockDim.x) {
      int i = 0, rem = t;
      while (rem >= a.len - 1 - i) {
        rem -= a.len - 1 - i;
        ++i;
      }
      const int j = i + 1 + rem;
      double dx = a.nx[i] - a.nx[j], dy = a.ny[i] - a.ny[j], dz = a.nz[i] - a.nz[j];
      min_image_vec(p, dx, dy, dz);
      double dist = sqrt(dx * dx + dy * dy + dz * dz);
      corr -= a.qr[i] * a.qr[j] * erf(p.alpha * dist) / dist;
    }
    for (int i = threadIdx.x; i < a.len; i += blockDim.x) self -= a.q[i] * a.q[i];
  }
  lj = block_sum(lj, scratch);
  re = block_sum(re, scratch);
  ov = block_sum(ov, scratch);
  rc = block_sum(rc, scratch);
  if (a.doCorrection) {
    corr = block_sum(corr, scratch);
    self = block_sum(self, scratch);
  }
  if (threadIdx.x == 0) {
    hostOut[0] = lj;
    hostOut[1] = re;
    hostOut[2] = ov;
    hostOut[3] = rc;
    hostOut[4] = kQQFact * corr;
    hostOut[5] = self * p.alpha * kQQFact * 1.12837916709551257390 * 0.5;
    *ticket = 0u;
    __threadfence_system();
    *hostFlag = a.seq;
  }
}

// Accepted single-molecule move (the reference's CellList::RemoveMol/AddMol +
// coordinate commit, src/moves/Translate.h:116-140): the new coordinates travel in the
// kernel parameters; when no atom changed cell the cell-sorted copy is patched in place.
struct AcceptArgs {
  int len, first, molIndex, hasCom, inPlace;
  double x[kTrialMaxAtoms], y[kTrialMaxAtoms], z[kTrialMaxAtoms];
  double com[3];
};

__global__ void k_accept_mol(const __grid_constant__ AcceptArgs a, double *x, double *y,
                             double *z, double *comx, double *comy, double *comz,
                             const int *__restrict__ sortedPos, double *sx, double *sy,
                             double *sz) {
  const int i = threadIdx.x;
  if (i < a.len) {
    x[a.first + i] = a.x[i];
    y[a.first + i] = a.y[i];
    z[a.first + i] = a.z[i];
    if (a.inPlace) {
      const int pos = sortedPos[a.first + i];
      sx[pos] = a.x[i];
      sy[pos] = a.y[i];
      sz[pos] = a.z[i];
    }
  }
  if (i == 0 && a.hasCom) {
    comx[a.molIndex] = a.com[0];
    comy[a.molIndex] = a.com[1];
    comz[a.molIndex] = a.com[2];
  }
}

// CalculateEnergy::ParticleNonbonded, src/CalculateEnergy.cpp:689-725: one thread per trial
// position, partners in the caller's (sortedNB) order.  buf = {px, py, pz, q}[nPartners] then
// {tx, ty, tz}[trials]; kinds in pk.  out[t] = the increment of inter[t].
template <int VDW>
__global__ void k_particle_nonbonded(BoxParams p, int kindI, double qI, int nPartners,
                                     const int *__restrict__ pk,
                                     const double *__restrict__ buf, int trials,
                                     double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= trials) return;
  const double *tr = buf + 4 * (size_t)nPartners;
  const double tx = tr[t], ty = tr[trials + t], tz = tr[2 * trials + t];
  double en = 0.0;
  for (int k = 0; k < nPartners; ++k) {
    double dx = tx - buf[k], dy = ty - buf[nPartners + k], dz = tz - buf[2 * nPartners + k];
    min_image_vec(p, dx, dy, dz);
    const double r2 = dist_sq(dx, dy, dz);
    if (!(p.boxRcutSq > r2)) continue;  // BoxDimensions::InRcut, strict
    en += calc_en<VDW>(p, r2, kindI + pk[k] * p.kindCount);
    if (p.electrostatic) {
      const double qq = qI * buf[3 * nPartners + k] * kQQFact;
      // FFParticle::CalcCoulombAdd_1_4, NB = true (src/FFParticle.cpp:281-292)
      if (qq != 0.0 && !(p.rCutSq < r2)) en += qq / sqrt(r2);
    }
  }
  out[t] = en;
}

}  // namespace gb
