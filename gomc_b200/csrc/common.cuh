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

enum { VDW_STD = 0, VDW_SHIFT = 1, VDW_SWITCH = 2 };

// Per-box constants handed to kernels by value.
struct BoxParams {
  double ax[3], half[3];
  double rCut, rCutSq, rCutLowSq, rCutCoulombSq, boxRcutSq;
  double alpha, alphaSq;
  double rOnSq, factor1, factor2;  // FF_SWITCH::Init, src/FFSwitch.h:99-106
  int kindCount, vdwKind, ewald, electrostatic;
  const double *sigmaSq, *epsilon_cn, *n, *shiftConst;
  const int *nHalf;  // n/2 when that is an integer in [1,64], else 0
};

__device__ __forceinline__ double min_image(double raw, double ax, double half) {
  // BoxDimensions::MinImageSigned, src/BoxDimensions.h:169-175
  if (raw > half)
    raw -= ax;
  else if (raw < -half)
    raw += ax;
  return raw;
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

template <int VDW>
__device__ __forceinline__ double calc_en(const BoxParams &p, double r2,
                                          int idx) {
  if (p.rCutSq < r2) return 0.0;
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

template <int VDW>
__device__ __forceinline__ void calc_en_vir(const BoxParams &p, double r2,
                                            int idx, double &en, double &vir) {
  en = 0.0;
  vir = 0.0;
  if (p.rCutSq < r2) return;
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

template <int VDW>
__device__ __forceinline__ double calc_coulomb(const BoxParams &p, double r2,
                                               double qq) {
  if (p.rCutCoulombSq < r2) return 0.0;
  double dist = sqrt(r2);
  if (p.ewald) return qq * erfc(p.alpha * dist) / dist;
  if (VDW == VDW_SHIFT) return qq * (1.0 / dist - 1.0 / p.rCut);
  if (VDW == VDW_SWITCH) {
    double s = r2 / p.rCutSq - 1.0;
    s *= s;
    return qq * s / dist;
  }
  return qq / dist;
}

template <int VDW>
__device__ __forceinline__ void calc_coulomb_en_vir(const BoxParams &p,
                                                    double r2, double qq,
                                                    double &en, double &vir) {
  en = 0.0;
  vir = 0.0;
  if (p.rCutCoulombSq < r2) return;
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
  if (VDW == VDW_SHIFT) {
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
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
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
