// common.cuh -- shared device helpers for the gomc_b200 engine (sm_100a).
//
// Pair functors reproduce the CPU formulas of the reference (SURVEY.md
// Appendix A), not those of its GPU build: FFParticle::CalcEn/CalcVir/
// CalcCoulomb/CalcCoulombVir (src/FFParticle.cpp:295-446), FF_SHIFT
// (src/FFShift.h:145-295), FF_SWITCH (src/FFSwitch.h:141-310); lambda == 1.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gb {

constexpr double kQQFact = 167103.208067979;  // lib/NumLib.h:22
constexpr double kTwoOverSqrtPi = 1.12837916709551257390;  // M_2_SQRTPI
constexpr double kPi = 3.14159265358979323846;

enum { VDW_STD = 0, VDW_SHIFT = 1, VDW_SWITCH = 2, VDW_EXP6 = 3, VDW_MARTINI = 4 };
constexpr double kBigNum = 1.7976931348623158e+308;  // num::BIGNUM = DBL_MAX, lib/NumLib.h:15,24

// Per-box constants handed to kernels by value.
struct BoxParams {
  double ax[3], half[3];
  double rCut, rCutSq, rCutLowSq, rCutCoulombSq, boxRcutSq;
  double alpha, alphaSq;
  double rOnSq, factor1, factor2;  // FF_SWITCH::Init, src/FFSwitch.h:99-106
  int kindCount, vdwKind, ewald, electrostatic;
  const double *sigmaSq, *epsilon_cn, *n, *shiftConst;
  const int *nHalf;  // n/2 when that is an integer in [1,64], else 0
  // EXP6 tables (FF_EXP6::Init, src/FFExp6.h:99-147), handed in by the host
  const double *rMin, *expConst, *rMaxSq;
  // Martini switch (FF_SWITCH_MARTINI::Init, src/FFSwitchMartini.h:121-213)
  const double *mAn, *mBn, *mCn, *mSign, *mSig6;
  double rOn, A6, B6, C6, A1, B1, C1, diElectric_1;
  // non-orthogonal cell: rows of B = normalised cell vectors, Bi = inverse of B
  // (BoxDimensionsNonOrth, src/BoxDimensionsNonOrth.h:79-109); ax[] = edge lengths
  int nonOrth;
  double B[9], Bi[9];
  // molecule centres of mass (virial sweep only)
  const double *comx, *comy, *comz;
  // fractional molecule of the box (lib/Lambda.h; -1: none) and the soft-core constants
  // (src/Forcefield.cpp:58-75)
  int lambdaMol, scCoul;
  double lambdaVDW, lambdaCoulomb, scAlpha, scSigma6, scPower;
};

// BoxDimensions::UnwrapPBC (scalar), src/BoxDimensions.cpp:297-320
__device__ __forceinline__ double unwrap_scalar(double v, double ref, double ax, double half) {
  if (fabs(ref - v) > half) {
    if (ref < half)
      v -= ax;
    else
      v += ax;
  }
  return v;
}

__device__ __forceinline__ double min_image(double raw, double ax, double half) {
  // BoxDimensions::MinImageSigned, src/BoxDimensions.h:169-175
  if (raw > half)
    raw -= ax;
  else if (raw < -half)
    raw += ax;
  return raw;
}

// BoxDimensions::MinImage / BoxDimensionsNonOrth::MinImage on a difference vector:
// unslant (v * Bi), signed wrap per axis, slant back (v * B).
__device__ __forceinline__ void min_image_vec(const BoxParams &p, double &dx, double &dy,
                                              double &dz) {
  if (p.nonOrth) {
    double ux = dx * p.Bi[0] + dy * p.Bi[3] + dz * p.Bi[6];
    double uy = dx * p.Bi[1] + dy * p.Bi[4] + dz * p.Bi[7];
    double uz = dx * p.Bi[2] + dy * p.Bi[5] + dz * p.Bi[8];
    ux = min_image(ux, p.ax[0], p.half[0]);
    uy = min_image(uy, p.ax[1], p.half[1]);
    uz = min_image(uz, p.ax[2], p.half[2]);
    dx = ux * p.B[0] + uy * p.B[3] + uz * p.B[6];
    dy = ux * p.B[1] + uy * p.B[4] + uz * p.B[7];
    dz = ux * p.B[2] + uy * p.B[5] + uz * p.B[8];
  } else {
    dx = min_image(dx, p.ax[0], p.half[0]);
    dy = min_image(dy, p.ax[1], p.half[1]);
    dz = min_image(dz, p.ax[2], p.half[2]);
  }
}

// BoxDimensions::UnwrapPBC(x,y,z,b,ref) / the non-orthogonal override
// (src/BoxDimensionsNonOrth.cpp:305-319): unslant point and reference, unwrap, slant.
__device__ __forceinline__ void unwrap_vec(const BoxParams &p, double &x, double &y, double &z,
                                           double rx, double ry, double rz) {
  if (p.nonOrth) {
    double ux = x * p.Bi[0] + y * p.Bi[3] + z * p.Bi[6];
    double uy = x * p.Bi[1] + y * p.Bi[4] + z * p.Bi[7];
    double uz = x * p.Bi[2] + y * p.Bi[5] + z * p.Bi[8];
    double vx = rx * p.Bi[0] + ry * p.Bi[3] + rz * p.Bi[6];
    double vy = rx * p.Bi[1] + ry * p.Bi[4] + rz * p.Bi[7];
    double vz = rx * p.Bi[2] + ry * p.Bi[5] + rz * p.Bi[8];
    ux = unwrap_scalar(ux, vx, p.ax[0], p.half[0]);
    uy = unwrap_scalar(uy, vy, p.ax[1], p.half[1]);
    uz = unwrap_scalar(uz, vz, p.ax[2], p.half[2]);
    x = ux * p.B[0] + uy * p.B[3] + uz * p.B[6];
    y = ux * p.B[1] + uy * p.B[4] + uz * p.B[7];
    z = ux * p.B[2] + uy * p.B[5] + uz * p.B[8];
  } else {
    x = unwrap_scalar(x, rx, p.ax[0], p.half[0]);
    y = unwrap_scalar(y, ry, p.ax[1], p.half[1]);
    z = unwrap_scalar(z, rz, p.ax[2], p.half[2]);
  }
}

// One definition of r^2 so that the cut-off test and the functor see the
// same bits wherever it is recomputed.
__device__ __forceinline__ double dist_sq(double dx, double dy, double dz) {
  return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

// (sigma^2/r^2)^(n/2): integer n/2 by repeated squaring (n = 12 is the common
// case and costs one multiply), general Mie exponent by pow().
__device__ __forceinline__ double mie_repulse(double rRat2, double attract,
                                              double n, int nHalf) {
  if (nHalf == 6) return attract * attract;
  if (nHalf > 0) {
    double r = 1.0, b = rRat2;
    int e = nHalf;
    while (e) {
      if (e & 1) r *= b;
      b *= b;
      e >>= 1;
    }
    return r;
  }
  return pow(rRat2, n * 0.5);
}

// CUT = false: the reference's two-argument functor forms (no cut-off test), which
// the soft-core wrappers call with the softened r^2.
template <int VDW, bool CUT = true>
__device__ __forceinline__ double calc_en(const BoxParams &p, double r2,
                                          int idx) {
  if (CUT && p.rCutSq < r2) return 0.0;
  if (VDW == VDW_EXP6) {  // src/FFExp6.h:184-224
    if (r2 < p.rMaxSq[idx]) return kBigNum;
    double dist = sqrt(r2);
    double rm = p.rMin[idx];
    double rRat = rm / dist;
    double rRat2 = rRat * rRat;
    double attract = rRat2 * rRat2 * rRat2;
    double alpha_ij = (double)(unsigned)p.n[idx];  // truncated to uint, FFExp6.h:215
    double repulse = (6.0 / alpha_ij) * exp(alpha_ij * (1.0 - dist / rm));
    return p.expConst[idx] * (repulse - attract);
  }
  if (VDW == VDW_MARTINI) {  // src/FFSwitchMartini.h:279-302
    double r_2 = 1.0 / r2;
    double r_6 = r_2 * r_2 * r_2;
    double n_ij = p.n[idx];
    double r_n = pow(r_2, n_ij * 0.5);
    double rij_ron = sqrt(r2) - p.rOn;
    double rij_ron_2 = rij_ron * rij_ron;
    double rij_ron_3 = rij_ron_2 * rij_ron;
    double rij_ron_4 = rij_ron_2 * rij_ron_2;
    double An = p.mAn[idx], Bn = p.mBn[idx], Cn = p.mCn[idx];
    bool sw = r2 > p.rOnSq;
    double shiftRep = sw ? (-(An / 3.0) * rij_ron_3 - (Bn / 4.0) * rij_ron_4 - Cn) : -Cn;
    double shiftAtt = sw ? (-(p.A6 / 3.0) * rij_ron_3 - (p.B6 / 4.0) * rij_ron_4 - p.C6) : -p.C6;
    return p.epsilon_cn[idx] * (p.mSign[idx] * (r_n + shiftRep) - p.mSig6[idx] * (r_6 + shiftAtt));
  }
  double eps = p.epsilon_cn[idx];
  if (eps == 0.0) return 0.0;  // exact: eps*(finite) == 0
  double rRat2 = p.sigmaSq[idx] / r2;
  double attract = rRat2 * rRat2 * rRat2;
  double repulse = mie_repulse(rRat2, attract, p.n[idx], p.nHalf[idx]);
  double e = eps * (repulse - attract);
  if (VDW == VDW_SHIFT) {
    e -= p.shiftConst[idx];
  } else if (VDW == VDW_SWITCH) {
    double d = p.rCutSq - r2;
    double fE = d * d * p.factor2 * (p.factor1 + 2.0 * r2);
    e *= (r2 > p.rOnSq ? fE : 1.0);
  }
  return e;
}

template <int VDW, bool CUT = true>
__device__ __forceinline__ void calc_en_vir(const BoxParams &p, double r2,
                                            int idx, double &en, double &vir) {
  en = 0.0;
  vir = 0.0;
  if (CUT && p.rCutSq < r2) return;
  if (VDW == VDW_EXP6) {  // src/FFExp6.h:184-257
    if (r2 < p.rMaxSq[idx]) {
      en = kBigNum;
      vir = kBigNum;
      return;
    }
    double dist = sqrt(r2);
    double rm = p.rMin[idx];
    double rRat = rm / dist;
    double rRat2 = rRat * rRat;
    double attract = rRat2 * rRat2 * rRat2;
    double alpha_ij = (double)(unsigned)p.n[idx];
    double ex = exp(alpha_ij * (1.0 - dist / rm));
    en = p.expConst[idx] * ((6.0 / alpha_ij) * ex - attract);
    vir = 6.0 * p.expConst[idx] * ((dist / rm) * ex - attract) / r2;
    return;
  }
  if (VDW == VDW_MARTINI) {  // src/FFSwitchMartini.h:279-348 (r_8 = (r^2)^4 as written there)
    en = calc_en<VDW_MARTINI, CUT>(p, r2, idx);
    double n_ij = p.n[idx];
    double r_1 = 1.0 / sqrt(r2);
    double r_8 = r2 * r2 * r2 * r2;
    double r_n2 = pow(r_1, n_ij + 2.0);
    double rij_ron = sqrt(r2) - p.rOn;
    double rij_ron_2 = rij_ron * rij_ron;
    double rij_ron_3 = rij_ron_2 * rij_ron;
    bool sw = r2 > p.rOnSq;
    double dshiftRep = sw ? (p.mAn[idx] * rij_ron_2 + p.mBn[idx] * rij_ron_3) * r_1 : 0.0;
    double dshiftAtt = sw ? (p.A6 * rij_ron_2 + p.B6 * rij_ron_3) * r_1 : 0.0;
    vir = p.epsilon_cn[idx] *
          (p.mSign[idx] * (n_ij * r_n2 + dshiftRep) - p.mSig6[idx] * (6.0 * r_8 + dshiftAtt));
    return;
  }
  double eps = p.epsilon_cn[idx];
  if (eps == 0.0) {
    if (VDW == VDW_SHIFT) en = -p.shiftConst[idx];
    return;
  }
  double rNeg2 = 1.0 / r2;
  double rRat2 = rNeg2 * p.sigmaSq[idx];
  double attract = rRat2 * rRat2 * rRat2;
  double n = p.n[idx];
  double repulse = mie_repulse(rRat2, attract, n, p.nHalf[idx]);
  double Eij = eps * (repulse - attract);
  double Wij = (eps * 6.0) * ((n / 6.0) * repulse - attract) * rNeg2;
  if (VDW == VDW_SHIFT) {
    en = Eij - p.shiftConst[idx];
    vir = Wij;
  } else if (VDW == VDW_SWITCH) {
    double d = p.rCutSq - r2;
    double fE = d * d * p.factor2 * (p.factor1 + 2.0 * r2);
    double fW = 12.0 * p.factor2 * d * (p.rOnSq - r2);
    bool sw = r2 > p.rOnSq;
    double factE = sw ? fE : 1.0, factW = sw ? fW : 0.0;
    en = Eij * factE;
    vir = Wij * factE - Eij * factW;
  } else {
    en = Eij;
    vir = Wij;
  }
}

template <int VDW, bool CUT = true>
__device__ __forceinline__ double calc_coulomb(const BoxParams &p, double r2,
                                               double qq) {
  if (CUT && p.rCutCoulombSq < r2) return 0.0;
  double dist = sqrt(r2);
  if (p.ewald) return qq * erfc(p.alpha * dist) / dist;
  if (VDW == VDW_MARTINI) {  // src/FFSwitchMartini.h:379-396
    double coul = -(p.A1 / 3.0) * (dist * r2) - (p.B1 / 4.0) * (r2 * r2) - p.C1;
    return qq * p.diElectric_1 * (1.0 / dist + coul);
  }
  if (VDW == VDW_SHIFT) return qq * (1.0 / dist - 1.0 / p.rCut);
  if (VDW == VDW_SWITCH) {
    double s = r2 / p.rCutSq - 1.0;
    s *= s;
    return qq * s / dist;
  }
  return qq / dist;
}

template <int VDW, bool CUT = true>
__device__ __forceinline__ void calc_coulomb_en_vir(const BoxParams &p,
                                                    double r2, double qq,
                                                    double &en, double &vir) {
  en = 0.0;
  vir = 0.0;
  if (CUT && p.rCutCoulombSq < r2) return;
  double dist = sqrt(r2);
  if (p.ewald) {
    double x = p.alpha * dist;
    double ec = erfc(x);
    double ex = exp(-1.0 * p.alphaSq * r2);
    en = qq * ec / dist;
    // STD writes 1-erf, SHIFT/SWITCH erfc (FFParticle.cpp:439, FFShift.h:289);
    // the two agree to rounding, erfc is the better conditioned one.
    vir = qq * (ec / dist + (p.alpha * kTwoOverSqrtPi) * ex) / r2;
    return;
  }
  if (VDW == VDW_MARTINI) {  // src/FFSwitchMartini.h:379-396, :440-447
    double coul = -(p.A1 / 3.0) * (dist * r2) - (p.B1 / 4.0) * (r2 * r2) - p.C1;
    en = qq * p.diElectric_1 * (1.0 / dist + coul);
    double virCoul = p.A1 / r2 + p.B1 / (dist * r2);
    vir = qq * p.diElectric_1 * (1.0 / (dist * r2) + virCoul / dist);
  } else if (VDW == VDW_SHIFT) {
    en = qq * (1.0 / dist - 1.0 / p.rCut);
    vir = qq / (r2 * dist);
  } else if (VDW == VDW_SWITCH) {
    double s0 = r2 / p.rCutSq - 1.0;
    double s = s0 * s0;
    double ds = 2.0 * s0 * 2.0 * dist / p.rCutSq;
    en = qq * s / dist;
    vir = -qq * (ds / r2 - s / (r2 * dist));
  } else {
    en = qq / dist;
    vir = qq / (r2 * dist);
  }
}

// 32-bit shared-window addressing: explicit LDS/STS instead of generic LD/ST.
// ---- lambda-taking functor forms (fractional molecule) -------------------------
// src/FFParticle.cpp:295-315, :327-348, :360-386, :400-429 and the identical text in
// FFShift.h, FFSwitch.h, FFSwitchMartini.h, FFExp6.h (which tests rMaxSq first).
__device__ __forceinline__ double soft_rsq(const BoxParams &p, double r2, int idx, double lambda) {
  double s2 = p.sigmaSq[idx];
  double sigma6 = fmax(s2 * s2 * s2, p.scSigma6);
  double dist6 = r2 * r2 * r2;
  double lambdaCoef = p.scAlpha * pow(1.0 - lambda, p.scPower);
  return cbrt(lambdaCoef * sigma6 + dist6);
}
template <int VDW>
__device__ __forceinline__ double calc_en_l(const BoxParams &p, double r2, int idx, double lambda) {
  if (lambda >= 0.999999) return calc_en<VDW>(p, r2, idx);
  if (p.rCutSq < r2) return 0.0;
  if (VDW == VDW_EXP6 && r2 < p.rMaxSq[idx]) return kBigNum;
  return lambda * calc_en<VDW, false>(p, soft_rsq(p, r2, idx, lambda), idx);
}
template <int VDW>
__device__ __forceinline__ void calc_en_vir_l(const BoxParams &p, double r2, int idx,
                                              double lambda, double &en, double &vir) {
  if (lambda >= 0.999999) {
    calc_en_vir<VDW>(p, r2, idx, en, vir);
    return;
  }
  en = vir = 0.0;
  if (p.rCutSq < r2) return;
  if (VDW == VDW_EXP6 && r2 < p.rMaxSq[idx]) {
    en = vir = kBigNum;
    return;
  }
  const double soft = soft_rsq(p, r2, idx, lambda);
  const double corr = r2 / soft;
  calc_en_vir<VDW, false>(p, soft, idx, en, vir);
  en *= lambda;
  vir *= lambda * corr * corr;
}
template <int VDW>
__device__ __forceinline__ double calc_coulomb_l(const BoxParams &p, double r2, int idx,
                                                 double qq, double lambda) {
  if (lambda >= 0.999999) return calc_coulomb<VDW>(p, r2, qq);
  if (p.rCutCoulombSq < r2) return 0.0;
  const double r = p.scCoul ? soft_rsq(p, r2, idx, lambda) : r2;
  return lambda * calc_coulomb<VDW, false>(p, r, qq);
}
template <int VDW>
__device__ __forceinline__ void calc_coulomb_en_vir_l(const BoxParams &p, double r2, int idx,
                                                      double qq, double lambda, double &en,
                                                      double &vir) {
  if (lambda >= 0.999999) {
    calc_coulomb_en_vir<VDW>(p, r2, qq, en, vir);
    return;
  }
  en = vir = 0.0;
  if (p.rCutCoulombSq < r2) return;
  if (p.scCoul) {
    const double soft = soft_rsq(p, r2, idx, lambda);
    const double corr = r2 / soft;
    calc_coulomb_en_vir<VDW, false>(p, soft, qq, en, vir);
    en *= lambda;
    vir *= lambda * corr * corr;
  } else {
    calc_coulomb_en_vir<VDW, false>(p, r2, qq, en, vir);
    en *= lambda;
    vir *= lambda;
  }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// ---- PTX helpers -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA unit (bytes % 16 == 0).
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes,
                                         unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Fixed-order warp and block reductions (deterministic).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;  // valid in lane 0
}

// blockDim.x must be a multiple of 32 and <= 1024; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *scratch /*>=32*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace gb
