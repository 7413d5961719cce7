// recip.cuh -- Ewald reciprocal-space kernels.
//
// Replaces BoxReciprocalSumsGPU / BoxReciprocalGPU / MolReciprocalGPU /
// SwapReciprocalGPU / BoxForceReciprocalGPU
// (src/GPU/CalculateEwaldCUDAKernel.cu:192-761) behind Ewald::BoxReciprocalSetup
// / BoxReciprocalSums / BoxReciprocal / MolReciprocal / Swap*Recip /
// BoxForceReciprocal (src/Ewald.cpp).
//
// Two structure-factor algorithms:
//  * direct     : one sincos per (atom, k), the reference's own algorithm
//                 (kept as an independent cross-check and for tiny boxes);
//  * factorised : for an orthogonal box k = 2pi (a/Lx, b/Ly, c/Lz), so
//                 exp(ik.r) = X^a Y^b Z^c with per-axis phases built by
//                 recurrence.  S(a,b,+-c) then is a (rows = (a,b) pairs) x
//                 (cols = c) contraction over atoms with 2 FP64 FMA per
//                 (atom, k) -- the FP64-pipe roofline of this path
//                  -- done with a 4x4x4 register tile
//                 per thread, operands staged in shared memory.
// All reductions are fixed-order (split-K partials are summed by a second
// kernel in slab order): no atomics, bit-reproducible.
#pragma once
#include "common.cuh"

namespace gb {

// ---------------------------------------------------------------------------
// Compact per-box list of charged atoms: pb[t] = {x, y, z, q}.
__global__ void k_pack_charged(int n, const int *__restrict__ chargedAtoms,
                               const double *__restrict__ x,
                               const double *__restrict__ y,
                               const double *__restrict__ z,
                               const double *__restrict__ q, double4 *pb) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int a = chargedAtoms[t];
  pb[t] = make_double4(x[a], y[a], z[a], q[a]);
}

// ---------------------------------------------------------------------------
// Direct algorithm.  grid = (ceil(nk/256), nSlabs); partial layout
// part[(slab*2 + {0,1})*nkStride + k].
constexpr int kDirectTile = 256;
__global__ void __launch_bounds__(256)
    k_recip_direct(int k0, int nk, int nkStride, int nAtoms, int atomsPerSlab,
                   const double4 *__restrict__ pb, const double *__restrict__ kx,
                   const double *__restrict__ ky, const double *__restrict__ kz,
                   double *__restrict__ part) {
  __shared__ double4 tile[kDirectTile];
  int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;  // this rank's k range is [k0, nk)
  int slab = blockIdx.y;
  int a0 = slab * atomsPerSlab;
  int a1 = min(nAtoms, a0 + atomsPerSlab);
  double kxv = 0.0, kyv = 0.0, kzv = 0.0;
  if (k < nk) {
    kxv = kx[k];
    kyv = ky[k];
    kzv = kz[k];
  }
  double sr = 0.0, si = 0.0;
  for (int base = a0; base < a1; base += kDirectTile) {
    int m = min(kDirectTile, a1 - base);
    __syncthreads();
    if (threadIdx.x < m) tile[threadIdx.x] = pb[base + threadIdx.x];
    __syncthreads();
    if (k < nk) {
      for (int t = 0; t < m; ++t) {
        double4 a = tile[t];
        // geom::Dot order, lib/GeomLib.h:73-76
        double dot = __dadd_rn(__dadd_rn(__dmul_rn(a.x, kxv), __dmul_rn(a.y, kyv)),
                               __dmul_rn(a.z, kzv));
        double s, c;
        sincos(dot, &s, &c);
        sr += a.w * c;
        si += a.w * s;
      }
    }
  }
  if (k < nk) {
    part[(size_t)(slab * 2 + 0) * nkStride + k] = sr;
    part[(size_t)(slab * 2 + 1) * nkStride + k] = si;
  }
}

// sumR[k] = sum_slab part, sumI likewise (slab order), and the per-block
// partial of sum_k prefact (R^2 + I^2)  (Ewald::BoxReciprocal).
__global__ void __launch_bounds__(256)
    k_recip_finish(int nk, int nkStride, int nSlabs,
                   const double *__restrict__ part,
                   const double *__restrict__ prefact, double *__restrict__ sumR,
                   double *__restrict__ sumI, double *__restrict__ blockEnergy) {
  __shared__ double scratch[32];
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (k < nk) {
    double r = 0.0, i = 0.0;
    for (int s = 0; s < nSlabs; ++s) {
      r += part[(size_t)(s * 2 + 0) * nkStride + k];
      i += part[(size_t)(s * 2 + 1) * nkStride + k];
    }
    sumR[k] = r;
    sumI[k] = i;
    e = (r * r + i * i) * prefact[k];
  }
  double s = block_sum(e, scratch);
  if (threadIdx.x == 0) blockEnergy[blockIdx.x] = s;
}

// Ewald::BoxReciprocal from existing sums.
__global__ void __launch_bounds__(256)
    k_recip_energy(int nk, const double *__restrict__ sumR,
                   const double *__restrict__ sumI,
                   const double *__restrict__ prefact,
                   double *__restrict__ blockEnergy) {
  __shared__ double scratch[32];
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (k < nk) e = (sumR[k] * sumR[k] + sumI[k] * sumI[k]) * prefact[k];
  double s = block_sum(e, scratch);
  if (threadIdx.x == 0) blockEnergy[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------
// Single-molecule deltas.  mode 0: MolReciprocal  new = ref + (sumNew - sumOld)
//                          mode 1: SwapDestRecip  new = ref + sumNew
//                          mode 2: SwapSourceRecip new = ref - sumNew
// mol = {n, then per atom q, newx, newy, newz, oldx, oldy, oldz} in a small
// device buffer (7 doubles per atom after the header double).
__global__ void __launch_bounds__(256)
    k_mol_recip(int nk, int mode, const double *__restrict__ molBuf,
                const double *__restrict__ kx, const double *__restrict__ ky,
                const double *__restrict__ kz,
                const double *__restrict__ prefact,
                const double *__restrict__ sumRref,
                const double *__restrict__ sumIref, double *__restrict__ sumRnew,
                double *__restrict__ sumInew, double *__restrict__ blockEnergy) {
  __shared__ double scratch[32];
  extern __shared__ double molSm[];
  int n = (int)molBuf[0];
  for (int t = threadIdx.x; t < 7 * n; t += blockDim.x) molSm[t] = molBuf[1 + t];
  __syncthreads();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (k < nk) {
    double kxv = kx[k], kyv = ky[k], kzv = kz[k];
    double rn = 0.0, in = 0.0, ro = 0.0, io = 0.0;
    for (int a = 0; a < n; ++a) {
      const double *m = molSm + 7 * a;
      double q = m[0];
      if (fabs(q) < 0.000000001) continue;  // particleHasNoCharge
      double dn = __dadd_rn(__dadd_rn(__dmul_rn(m[1], kxv), __dmul_rn(m[2], kyv)),
                            __dmul_rn(m[3], kzv));
      double s, c;
      sincos(dn, &s, &c);
      rn += q * c;
      in += q * s;
      if (mode == 0) {
        double d0 = __dadd_rn(__dadd_rn(__dmul_rn(m[4], kxv), __dmul_rn(m[5], kyv)),
                              __dmul_rn(m[6], kzv));
        sincos(d0, &s, &c);
        ro += q * c;
        io += q * s;
      }
    }
    double r, i;
    if (mode == 0) {
      r = sumRref[k] + (rn - ro);
      i = sumIref[k] + (in - io);
    } else if (mode == 1) {
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
  if (threadIdx.x == 0) blockEnergy[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------
// Ewald::VirialReciprocal, src/Ewald.cpp:1168-1305.
// k part (:1229-1244): part[c * gridDim.x + block], c = 0..2.
__global__ void __launch_bounds__(256)
    k_virial_recip_k(int nk, double constVal, const double *__restrict__ kx,
                     const double *__restrict__ ky, const double *__restrict__ kz,
                     const double *__restrict__ hsqr, const double *__restrict__ prefact,
                     const double *__restrict__ sumR, const double *__restrict__ sumI,
                     double *__restrict__ part) {
  __shared__ double scratch[32];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double w[3] = {0.0, 0.0, 0.0};
  if (k < nk) {
    const double factor = prefact[k] * (sumR[k] * sumR[k] + sumI[k] * sumI[k]);
    const double c = 2.0 * (constVal + 1.0 / hsqr[k]);
    w[0] = factor * (1.0 - c * kx[k] * kx[k]);
    w[1] = factor * (1.0 - c * ky[k] * ky[k]);
    w[2] = factor * (1.0 - c * kz[k] * kz[k]);
  }
  for (int c = 0; c < 3; ++c) {
    double s = block_sum(w[c], scratch);
    if (threadIdx.x == 0) part[(size_t)c * gridDim.x + blockIdx.x] = s;
  }
}
// Intramolecular part (:1247-1285).  Its per-atom k sum is minus the k-space
// reciprocal force on the atom (same factor as :1575-1586 with the sign flipped), so
// it is contracted from the force kernel's output: w_c -= F_c * (unwrap(r) - com)_c.
__global__ void __launch_bounds__(256)
    k_virial_recip_intra(BoxParams p, int nBoxAtoms, const int *__restrict__ atomList,
                         const int *__restrict__ mol, const double *__restrict__ x,
                         const double *__restrict__ y, const double *__restrict__ z,
                         const double *__restrict__ q, const double *__restrict__ fkx,
                         const double *__restrict__ fky, const double *__restrict__ fkz,
                         double *__restrict__ part) {
  __shared__ double scratch[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double w[3] = {0.0, 0.0, 0.0};
  if (t < nBoxAtoms) {
    const int a = atomList[t];
    if (!(fabs(q[a]) < 0.000000001)) {
      const int m = mol[a];
      const double cx = p.comx[m], cy = p.comy[m], cz = p.comz[m];
      double ux = x[a], uy = y[a], uz = z[a];
      unwrap_vec(p, ux, uy, uz, cx, cy, cz);
      w[0] = -fkx[a] * (ux - cx);
      w[1] = -fky[a] * (uy - cy);
      w[2] = -fkz[a] * (uz - cz);
    }
  }
  for (int c = 0; c < 3; ++c) {
    double s = block_sum(w[c], scratch);
    if (threadIdx.x == 0) part[(size_t)c * gridDim.x + blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------
// Weighted point charges: new = base + scale * sum_p w_p (cos, sin)(k.r_p).
// Ewald::MolExchangeReciprocal (src/Ewald.cpp:714-826; base = ref on the first
// call, the new sums afterwards) and Ewald::ChangeLambdaRecip (:534-585).
// buf = {w[n], x[n], y[n], z[n]}.
__global__ void __launch_bounds__(256)
    k_recip_weighted(int nk, int n, const double *__restrict__ buf, double scale,
                     const double *__restrict__ kx, const double *__restrict__ ky,
                     const double *__restrict__ kz, const double *__restrict__ prefact,
                     const double *baseR, const double *baseI, double *sumRnew,
                     double *sumInew, double *__restrict__ blockEnergy) {
  __shared__ double scratch[32];
  __shared__ double sm[4][256];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < nk;
  const double kxv = live ? kx[k] : 0.0, kyv = live ? ky[k] : 0.0, kzv = live ? kz[k] : 0.0;
  double sR = 0.0, sI = 0.0;
  for (int c0 = 0; c0 < n; c0 += 256) {
    const int cn = min(256, n - c0);
    __syncthreads();
    if ((int)threadIdx.x < cn)
      for (int f = 0; f < 4; ++f) sm[f][threadIdx.x] = buf[(size_t)f * n + c0 + threadIdx.x];
    __syncthreads();
    if (live)
      for (int a = 0; a < cn; ++a) {
        double d = __dadd_rn(__dadd_rn(__dmul_rn(sm[1][a], kxv), __dmul_rn(sm[2][a], kyv)),
                             __dmul_rn(sm[3][a], kzv));
        double sn, cs;
        sincos(d, &sn, &cs);
        sR += sm[0][a] * cs;
        sI += sm[0][a] * sn;
      }
  }
  double e = 0.0;
  if (live) {
    double r = baseR[k] + scale * sR, i = baseI[k] + scale * sI;
    sumRnew[k] = r;
    sumInew[k] = i;
    e = (r * r + i * i) * prefact[k];
  }
  double s = block_sum(e, scratch);
  if (threadIdx.x == 0) blockEnergy[blockIdx.x] = s;
}

// Ewald::ChangeRecip, src/Ewald.cpp:589-642: E_recip of every lambda state with the
// molecule's charges scaled by sqrt(lambda_s) - sqrt(lambda_iState); resident
// coordinates.  blockEnergy[s * gridDim.x + block].
constexpr int kMaxLambdaStates = 64;
__global__ void __launch_bounds__(256)
    k_change_recip(int nk, int first, int len, const double *__restrict__ x,
                   const double *__restrict__ y, const double *__restrict__ z,
                   const double *__restrict__ q, int nStates,
                   const double *__restrict__ coefDiff, const double *__restrict__ kx,
                   const double *__restrict__ ky, const double *__restrict__ kz,
                   const double *__restrict__ prefact, const double *__restrict__ sumRref,
                   const double *__restrict__ sumIref, double *__restrict__ blockEnergy) {
  __shared__ double scratch[32];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double sR = 0.0, sI = 0.0, rr = 0.0, ri = 0.0, pf = 0.0;
  if (k < nk) {
    const double kxv = kx[k], kyv = ky[k], kzv = kz[k];
    for (int a = first; a < first + len; ++a) {
      const double qa = q[a];
      if (fabs(qa) < 0.000000001) continue;
      double d = __dadd_rn(__dadd_rn(__dmul_rn(x[a], kxv), __dmul_rn(y[a], kyv)),
                           __dmul_rn(z[a], kzv));
      double sn, cs;
      sincos(d, &sn, &cs);
      sR += qa * cs;
      sI += qa * sn;
    }
    rr = sumRref[k];
    ri = sumIref[k];
    pf = prefact[k];
  }
  for (int s = 0; s < nStates; ++s) {
    const double c = coefDiff[s];
    const double a = rr + c * sR, b = ri + c * sI;
    double e = block_sum(pf * (a * a + b * b), scratch);
    if (threadIdx.x == 0) blockEnergy[(size_t)s * gridDim.x + blockIdx.x] = e;
  }
}

// ---------------------------------------------------------------------------
// Reciprocal force, direct algorithm (Ewald::BoxForceReciprocal CPU branch,
// src/Ewald.cpp:1541-1592).  One thread per box atom, k staged in tiles.
constexpr int kForceTile = 128;
__global__ void __launch_bounds__(128)
    k_force_recip_direct(BoxParams p, int nBoxAtoms,
                         const int *__restrict__ atomList,
                         const int *__restrict__ mol,
                         const int *__restrict__ molStart,
                         const double *__restrict__ x,
                         const double *__restrict__ y,
                         const double *__restrict__ z,
                         const double *__restrict__ q, int nk,
                         const double *__restrict__ kx,
                         const double *__restrict__ ky,
                         const double *__restrict__ kz,
                         const double *__restrict__ prefact,
                         const double *__restrict__ sumR,
                         const double *__restrict__ sumI, double *rfx,
                         double *rfy, double *rfz, int withIntra = 1) {
  __shared__ double tk[5][kForceTile];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int a = t < nBoxAtoms ? atomList[t] : -1;
  double X = 0.0, Y = 0.0, Z = 0.0;
  double xa = 0.0, ya = 0.0, za = 0.0, qa = 0.0;
  bool charged = false;
  if (a >= 0) {
    xa = x[a];
    ya = y[a];
    za = z[a];
    qa = q[a];
    charged = !(fabs(qa) < 0.000000001);
    if (charged && withIntra) {  // intramolecular correction force, :1556-1569
      int m = mol[a];
      double constValue = p.alpha * kTwoOverSqrtPi;
      for (int j = molStart[m]; j < molStart[m + 1]; ++j) {
        if (j == a) continue;
        double dx = xa - x[j], dy = ya - y[j], dz = za - z[j];
        min_image_vec(p, dx, dy, dz);
        double r2 = dx * dx + dy * dy + dz * dz;
        double dist = sqrt(r2);
        double ex = exp(-1.0 * p.alphaSq * r2);
        double qiqj = qa * q[j] * kQQFact;
        double f = qiqj / r2;
        f *= (erf(p.alpha * dist) / dist) - constValue * ex;
        X -= f * dx;
        Y -= f * dy;
        Z -= f * dz;
      }
    }
  }
  for (int base = 0; base < nk; base += kForceTile) {
    int m = min(kForceTile, nk - base);
    __syncthreads();
    if (threadIdx.x < m) {
      int k = base + threadIdx.x;
      double pf = prefact[k];
      tk[0][threadIdx.x] = kx[k];
      tk[1][threadIdx.x] = ky[k];
      tk[2][threadIdx.x] = kz[k];
      tk[3][threadIdx.x] = pf * sumR[k];
      tk[4][threadIdx.x] = pf * sumI[k];
    }
    __syncthreads();
    if (charged) {
      for (int i = 0; i < m; ++i) {
        double kxv = tk[0][i], kyv = tk[1][i], kzv = tk[2][i];
        double dot = xa * kxv + ya * kyv + za * kzv;
        double s, c;
        sincos(dot, &s, &c);
        double factor = 2.0 * qa * (s * tk[3][i] - c * tk[4][i]);
        X += factor * kxv;
        Y += factor * kyv;
        Z += factor * kzv;
      }
    }
  }
  if (a >= 0) {
    rfx[a] = X;
    rfy[a] = Y;
    rfz[a] = Z;
  }
}

// Ewald::BoxSelf (src/Ewald.cpp:1125-1163) and the box sum of
// Ewald::MolCorrection (src/Ewald.cpp:1056-1085): per-block partials.
__global__ void __launch_bounds__(256)
    k_self_correction(BoxParams p, int nMolsBox, const int *__restrict__ molList,
                      const int *__restrict__ molStart,
                      const double *__restrict__ x, const double *__restrict__ y,
                      const double *__restrict__ z, const double *__restrict__ q,
                      double *__restrict__ blockSelf,
                      double *__restrict__ blockCorr, int selfScaleMol, double selfScale,
                      double corrScale) {
  __shared__ double scratch[32];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  double self = 0.0, corr = 0.0;
  if (t < nMolsBox) {
    int m = molList[t];
    int s = molStart[m], e = molStart[m + 1];
    for (int i = s; i < e; ++i) {
      self += q[i] * q[i];
      if (fabs(q[i]) < 0.000000001) continue;
      for (int j = i + 1; j < e; ++j) {
        double dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];
        min_image_vec(p, dx, dy, dz);
        double dist = sqrt(dx * dx + dy * dy + dz * dz);
        corr += q[i] * q[j] * erf(p.alpha * dist) / dist;
      }
    }
    // fractional molecule: MolCorrection scales by lambdaCoef^2 (src/Ewald.cpp:1084);
    // BoxSelf by lambda, but only in the case its kind-for-molecule index mix-up lets
    // through (:1140-1155) -- the host passes selfScaleMol = -1 otherwise
    if (m == p.lambdaMol) corr *= corrScale;
    if (m == selfScaleMol) self *= selfScale;
  }
  double a = block_sum(self, scratch);
  double b = block_sum(corr, scratch);
  if (threadIdx.x == 0) {
    blockSelf[blockIdx.x] = a;
    blockCorr[blockIdx.x] = b;
  }
}

// Ewald::SwapCorrection (src/Ewald.cpp:1311-1370) and SwapSelf (:1375-1391) of
// the molecule staged in molBuf {len, per atom: q, x, y, z, ...}.  One block;
// out[0] = correction, out[1] = self.
__global__ void __launch_bounds__(128)
    k_swap_correction(BoxParams p, const double *__restrict__ molBuf, double *__restrict__ out) {
  __shared__ double scratch[32];
  const int len = (int)molBuf[0];
  const int nPairs = len * (len - 1) / 2;
  double corr = 0.0, self = 0.0;
  for (int t = threadIdx.x; t < nPairs; t += blockDim.x) {
    // pair index -> (i, j), i < j, row-major over i
    int i = 0, rem = t;
    while (rem >= len - 1 - i) {
      rem -= len - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const double *a = molBuf + 1 + 7 * i, *b = molBuf + 1 + 7 * j;
    double dx = a[1] - b[1], dy = a[2] - b[2], dz = a[3] - b[3];
    min_image_vec(p, dx, dy, dz);
    double dist = sqrt(dx * dx + dy * dy + dz * dz);
    corr -= a[0] * b[0] * erf(p.alpha * dist) / dist;
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    // lambda = 1 charge (slot 4 of the swap-mode record): SwapSelf is not scaled by the
    // fractional molecule's lambda, the correction above is (src/Ewald.cpp:1340-1391)
    double q = molBuf[1 + 7 * i + 4];
    self -= q * q;
  }
  double c = block_sum(corr, scratch);
  double sf = block_sum(self, scratch);
  if (threadIdx.x == 0) {
    out[0] = kQQFact * c;
    out[1] = sf * p.alpha * kQQFact * 1.12837916709551257390 * 0.5;
  }
}

// ---------------------------------------------------------------------------
// Factorised structure factor.
//
// rows[]  : int4 {a, b, cmax, start}; (a,b) integer k indices of the row,
//           valid c in [-cmax, cmax] (row (0,0): [1, cmax]); `start` = index
//           in the reference's k list of the row's first entry.  Rows are
//           sorted by cmax (descending) and padded per tile with cmax = -1.
// tiles[] : int4 {rowBegin, RG, CG, cmaxTile}; thread (rg, cg) = (tid / CG,
//           tid % CG) owns rows rg + RG*u and cols cg + CG*v, u,v in 0..3.
// grid = (nTiles, nSlabs), block = 256, dynamic smem =
//   AT * (KX1 + KY1 + ZS + RS) * 16 B  (tables X, Y, Z and the A tile).
constexpr int kFactThreads = 256;
constexpr int kTR = 4, kTC = 4;
constexpr int kMaxRG = 32;  // rows per tile <= 128

struct FactArgs {
  const int4 *rows;
  const int4 *tiles;
  int KX1, KY1, KZ1;  // table lengths nkx_max+1, nky_max+1, nkz_max+1
  int ZS;             // Z row stride (>= 4*CG of every tile, zero padded)
  int RS;             // A row stride = max rows per tile
  int AT;             // atoms per chunk
  int nAtoms, atomsPerSlab;
  int nkStride;
  int tile0;  // first tile of this rank (multi-GPU sharding)
  double cvx, cvy, cvz;  // 2pi/L per axis (XYZ::Inverse then *2pi, Ewald.cpp:852-854)
};

__global__ void __launch_bounds__(kFactThreads, 1)
    k_recip_fact(FactArgs fa, const double4 *__restrict__ pb,
                 double *__restrict__ part) {
  extern __shared__ __align__(16) unsigned char dynSmem[];
  double2 *tabX = reinterpret_cast<double2 *>(dynSmem);
  double2 *tabY = tabX + fa.AT * fa.KX1;
  double2 *tabZ = tabY + fa.AT * fa.KY1;
  double2 *tileA = tabZ + fa.AT * fa.ZS;
  __shared__ int2 rowAB[kMaxRG * kTR];

  const int4 tile = fa.tiles[fa.tile0 + blockIdx.x];
  const int rowBegin = tile.x, RG = tile.y, CG = tile.z;
  const int R = RG * kTR;
  const int tid = threadIdx.x;
  const bool active = tid < RG * CG;
  const int rg = tid / CG, cg = tid - rg * CG;
  const int slab = blockIdx.y;
  const int a0 = slab * fa.atomsPerSlab;
  const int a1 = min(fa.nAtoms, a0 + fa.atomsPerSlab);

  for (int r = tid; r < R; r += kFactThreads) {
    int4 rw = fa.rows[rowBegin + r];
    rowAB[r] = make_int2(rw.x, rw.y);
  }

  double acc[kTR][kTC][4];
#pragma unroll
  for (int u = 0; u < kTR; ++u)
#pragma unroll
    for (int v = 0; v < kTC; ++v)
#pragma unroll
      for (int w = 0; w < 4; ++w) acc[u][v][w] = 0.0;

  for (int base = a0; base < a1; base += fa.AT) {
    const int nAt = min(fa.AT, a1 - base);
    __syncthreads();  // previous chunk fully consumed
    // ---- phase 1: per-axis phase tables by recurrence -------------------
    for (int t = tid; t < nAt * 3; t += kFactThreads) {
      int at = t / 3, axis = t - at * 3;
      double4 a = pb[base + at];
      double coord = axis == 0 ? a.x : (axis == 1 ? a.y : a.z);
      double cv = axis == 0 ? fa.cvx : (axis == 1 ? fa.cvy : fa.cvz);
      int len = axis == 0 ? fa.KX1 : (axis == 1 ? fa.KY1 : fa.KZ1);
      double2 *tab = axis == 0 ? tabX + at * fa.KX1
                               : (axis == 1 ? tabY + at * fa.KY1
                                            : tabZ + at * fa.ZS);
      double s1, c1;
      sincos(coord * cv, &s1, &c1);
      double scale = axis == 0 ? a.w : 1.0;  // fold the charge into X
      double cr = 1.0, ci = 0.0;
      for (int n = 0; n < len; ++n) {
        tab[n] = make_double2(scale * cr, scale * ci);
        double nr = cr * c1 - ci * s1;
        double ni = cr * s1 + ci * c1;
        cr = nr;
        ci = ni;
      }
      if (axis == 2)
        for (int n = len; n < fa.ZS; ++n) tab[n] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    // ---- phase 2: A[at][row] = q X^a Y^b ---------------------------------
    for (int t = tid; t < nAt * R; t += kFactThreads) {
      int at = t / R, r = t - at * R;
      int2 ab = rowAB[r];
      double2 xv = tabX[at * fa.KX1 + ab.x];
      int bb = ab.y < 0 ? -ab.y : ab.y;
      double2 yv = tabY[at * fa.KY1 + bb];
      if (ab.y < 0) yv.y = -yv.y;
      tileA[at * fa.RS + r] =
          make_double2(xv.x * yv.x - xv.y * yv.y, xv.x * yv.y + xv.y * yv.x);
    }
    __syncthreads();
    // ---- phase 3: rank-1 updates -----------------------------------------
    if (active) {
      const double2 *ap = tileA + rg;
      const double2 *zp = tabZ + cg;
      for (int at = 0; at < nAt; ++at) {
        double2 a[kTR], zv[kTC];
#pragma unroll
        for (int u = 0; u < kTR; ++u) a[u] = ap[at * fa.RS + RG * u];
#pragma unroll
        for (int v = 0; v < kTC; ++v) zv[v] = zp[at * fa.ZS + CG * v];
#pragma unroll
        for (int u = 0; u < kTR; ++u)
#pragma unroll
          for (int v = 0; v < kTC; ++v) {
            acc[u][v][0] = fma(a[u].x, zv[v].x, acc[u][v][0]);  // Ar*cz
            acc[u][v][1] = fma(a[u].y, zv[v].y, acc[u][v][1]);  // Ai*sz
            acc[u][v][2] = fma(a[u].x, zv[v].y, acc[u][v][2]);  // Ar*sz
            acc[u][v][3] = fma(a[u].y, zv[v].x, acc[u][v][3]);  // Ai*cz
          }
      }
    }
  }
  // ---- epilogue: S(a,b,+c) and S(a,b,-c) into the reference's k order -----
  if (active) {
    double *pr = part + (size_t)(slab * 2 + 0) * fa.nkStride;
    double *pi = part + (size_t)(slab * 2 + 1) * fa.nkStride;
#pragma unroll
    for (int u = 0; u < kTR; ++u) {
      int4 rw = fa.rows[rowBegin + rg + RG * u];
      int cmax = rw.z;
      bool origin = (rw.x == 0 && rw.y == 0);
#pragma unroll
      for (int v = 0; v < kTC; ++v) {
        int c = cg + CG * v;
        if (c > cmax) continue;
        double p1 = acc[u][v][0], p2 = acc[u][v][1], p3 = acc[u][v][2],
               p4 = acc[u][v][3];
        if (origin) {
          if (c >= 1) {
            pr[rw.w + c - 1] = p1 - p2;
            pi[rw.w + c - 1] = p3 + p4;
          }
        } else {
          pr[rw.w + cmax + c] = p1 - p2;
          pi[rw.w + cmax + c] = p3 + p4;
          if (c > 0) {
            pr[rw.w + cmax - c] = p1 + p2;
            pi[rw.w + cmax - c] = p4 - p3;
          }
        }
      }
    }
  }
}

// Regenerates the k list of an orthogonal box from its (a, b) row table: one block per
// row {a, b, cmax, first}; entries first.. hold c = -cmax..cmax (1..cmax for the a = b = 0
// row).  Each product and sum is rounded separately, in the order of Ewald::RecipInitOrth
// (src/Ewald.cpp:860-903), so the values equal the host enumeration bit for bit.
__global__ void k_gen_kvectors(const int4 *__restrict__ rows, double cv0, double cv1, double cv2,
                               double *__restrict__ kx, double *__restrict__ ky,
                               double *__restrict__ kz, double *__restrict__ hsqr) {
  const int4 rw = rows[blockIdx.x];
  if (rw.z < 0) return;
  const int clo = (rw.x == 0 && rw.y == 0) ? 1 : -rw.z;
  const double kX = __dmul_rn(cv0, (double)rw.x), kY = __dmul_rn(cv1, (double)rw.y);
  const double xy = __dadd_rn(__dmul_rn(kX, kX), __dmul_rn(kY, kY));
  for (int c = clo + threadIdx.x; c <= rw.z; c += blockDim.x) {
    const double kZ = __dmul_rn(cv2, (double)c);
    const int i = rw.w + (c - clo);
    kx[i] = kX;
    ky[i] = kY;
    kz[i] = kZ;
    hsqr[i] = __dadd_rn(xy, __dmul_rn(kZ, kZ));
  }
}

}  // namespace gb
