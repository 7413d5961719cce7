/*
 * gomc_oracle.c -- TEST INFRASTRUCTURE ONLY (see gomc_oracle.h).
 *
 * CPU restatement, in plain C99, of the GOMC v2.80 energy/force hot path.
 * Loop order follows the reference's CPU branch so that a single-thread run
 * reproduces the reference's +p1 summation order.  Build:
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC gomc_oracle.c -lm
 * (-ffp-contract=off: the reference x86-64 build has no FMA contraction).
 */
#include "gomc_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_2_SQRTPI
#define M_2_SQRTPI 1.12837916709551257390
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* PBC: BoxDimensions::MinImageSigned, src/BoxDimensions.h:169-175           */
static inline double min_image_signed(double raw, double ax, double halfAx) {
  if (raw > halfAx)
    raw -= ax;
  else if (raw < -halfAx)
    raw += ax;
  return raw;
}

/* BoxDimensions::InRcut, src/BoxDimensions.h:231-237; rCutSq[b] is the square
 * of max(rCut, rCutCoulomb[b]) (src/BoxDimensions.cpp:17-18). */
static inline double box_rcut(const orc_params *p) {
  return p->rCut > p->rCutCoulomb ? p->rCut : p->rCutCoulomb;
}
/* BoxDimensionsNonOrth::TransformUnSlant / TransformSlant,
 * src/BoxDimensionsNonOrth.h:84-109 (row vector times matrix). */
static inline void vec_mat(const double a[3], const double m[9], double out[3]) {
  out[0] = a[0] * m[0] + a[1] * m[3] + a[2] * m[6];
  out[1] = a[0] * m[1] + a[1] * m[4] + a[2] * m[7];
  out[2] = a[0] * m[2] + a[1] * m[5] + a[2] * m[8];
}
/* BoxDimensions::MinImage (src/BoxDimensions.h:61-66) or the non-orthogonal
 * override (src/BoxDimensionsNonOrth.cpp:240-245): unslant, wrap, slant. */
static inline void min_image_vec(const orc_params *p, double d[3]) {
  if (p->nonOrth) {
    double u[3];
    vec_mat(d, p->cellBasisInv, u);
    u[0] = min_image_signed(u[0], p->axis[0], p->axis[0] * 0.5);
    u[1] = min_image_signed(u[1], p->axis[1], p->axis[1] * 0.5);
    u[2] = min_image_signed(u[2], p->axis[2], p->axis[2] * 0.5);
    vec_mat(u, p->cellBasis, d);
  } else {
    d[0] = min_image_signed(d[0], p->axis[0], p->axis[0] * 0.5);
    d[1] = min_image_signed(d[1], p->axis[1], p->axis[1] * 0.5);
    d[2] = min_image_signed(d[2], p->axis[2], p->axis[2] * 0.5);
  }
}
static inline int in_rcut(const orc_params *p, double boxRcutSq, double xi,
                          double yi, double zi, double xj, double yj,
                          double zj, double *distSq, double d[3]) {
  d[0] = xi - xj;
  d[1] = yi - yj;
  d[2] = zi - zj;
  min_image_vec(p, d);
  *distSq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  return boxRcutSq > *distSq;
}

/* ------------------------------------------------------------------------ */
/* Pair functors (lambda == 1 branch)                                        */

/* FFParticle::CalcEn(distSq,index) src/FFParticle.cpp:317-325 */
static inline double mie_en(const orc_params *p, double distSq, int idx) {
  double rRat2 = p->sigmaSq[idx] / distSq;
  double rRat4 = rRat2 * rRat2;
  double attract = rRat4 * rRat2;
  double repulse = pow(rRat2, p->n[idx] * 0.5);
  return p->epsilon_cn[idx] * (repulse - attract);
}
/* FFParticle::CalcVir(distSq,index) src/FFParticle.cpp:350-359 */
static inline double mie_vir(const orc_params *p, double distSq, int idx) {
  double rNeg2 = 1.0 / distSq;
  double rRat2 = rNeg2 * p->sigmaSq[idx];
  double rRat4 = rRat2 * rRat2;
  double attract = rRat4 * rRat2;
  double repulse = pow(rRat2, p->n[idx] * 0.5);
  double epsilon_cn_6 = p->epsilon_cn[idx] * 6; /* FFParticle.cpp:191 */
  double nOver6 = p->n[idx] / 6;                /* FFParticle.cpp:193 */
  return epsilon_cn_6 * (nOver6 * repulse - attract) * rNeg2;
}
/* FF_SHIFT::Init shiftConst, src/FFShift.h:100-116 */
static inline double shift_const(const orc_params *p, int idx) {
  double rCutSq = p->rCut * p->rCut;
  double rRat2 = p->sigmaSq[idx] / rCutSq;
  double rRat4 = rRat2 * rRat2;
  double attract = rRat4 * rRat2;
  double repulse = pow(sqrt(rRat2), p->n[idx]);
  return p->epsilon_cn[idx] * (repulse - attract);
}

/* FF_SWITCH_MARTINI::Init constants, src/FFSwitchMartini.h:121-213 */
typedef struct {
  double A6, B6, C6, A1, B1, C1, An, Bn, Cn, sign, sig6;
} martini_consts;
static martini_consts martini_init(const orc_params *p, int idx) {
  martini_consts m;
  double rCut = p->rCut, rOn = p->rOn, rOnCoul = 0.0;
  m.A6 = 6.0 * (7.0 * rOn - 10.0 * rCut) / (pow(rCut, 8.0) * (rCut - rOn) * (rCut - rOn));
  m.B6 = -6.0 * (7.0 * rOn - 9.0 * rCut) /
         (pow(rCut, 8.0) * (rCut - rOn) * (rCut - rOn) * (rCut - rOn));
  m.C6 = pow(rCut, -6.0) - m.A6 / 3.0 * (rCut - rOn) * (rCut - rOn) * (rCut - rOn) -
         m.B6 / 4.0 * (rCut - rOn) * (rCut - rOn) * (rCut - rOn) * (rCut - rOn);
  m.A1 = (2.0 * rOnCoul - 5.0 * rCut) /
         (rCut * rCut * rCut * (rCut - rOnCoul) * (rCut - rOnCoul));
  m.B1 = -1.0 * (2.0 * rOnCoul - 4.0 * rCut) /
         (rCut * rCut * rCut * (rCut - rOnCoul) * (rCut - rOnCoul) * (rCut - rOnCoul));
  m.C1 = 1.0 / rCut - m.A1 / 3.0 * (rCut - rOnCoul) * (rCut - rOnCoul) * (rCut - rOnCoul) -
         m.B1 / 4.0 * (rCut - rOnCoul) * (rCut - rOnCoul) * (rCut - rOnCoul) *
             (rCut - rOnCoul);
  double pn = p->n[idx];
  m.An = pn * ((pn + 1.0) * rOn - (pn + 4.0) * rCut) /
         (pow(rCut, pn + 2.0) * (rCut - rOn) * (rCut - rOn));
  m.Bn = -pn * ((pn + 1.0) * rOn - (pn + 3.0) * rCut) /
         (pow(rCut, pn + 2.0) * (rCut - rOn) * (rCut - rOn) * (rCut - rOn));
  m.Cn = 1.0 / pow(rCut, pn) - m.An / 3.0 * (rCut - rOn) * (rCut - rOn) * (rCut - rOn) -
         m.Bn / 4.0 * (rCut - rOn) * (rCut - rOn) * (rCut - rOn) * (rCut - rOn);
  double sigma = sqrt(p->sigmaSq[idx]);
  m.sig6 = pow(sigma, 6.0);
  m.sign = pow(sigma, pn);
  return m;
}

#ifndef DBL_MAX
#define DBL_MAX 1.7976931348623158e+308 /* num::BIGNUM, lib/NumLib.h:15,24 */
#endif

/* ---- fractional molecule (free energy / NeMTMC): lib/Lambda.h state of one box
 * plus the soft-core constants of src/Forcefield.cpp:58-75.  Test-infrastructure
 * global: set with orc_set_lambda before calling the functions below. */
static struct {
  int mol; /* -1: no fractional molecule */
  double vdw, coulomb, sc_alpha, sc_sigma_6;
  int sc_power, sc_coul, molKind;
} g_lambda = {-1, 1.0, 1.0, 0.0, 0.0, 0, 0, -1};
void orc_set_lambda(int mol, double lambdaVDW, double lambdaCoulomb, double sc_alpha,
                    double sc_sigma_6, int sc_power, int sc_coul, int molKind) {
  g_lambda.mol = mol;
  g_lambda.molKind = molKind;
  g_lambda.vdw = lambdaVDW;
  g_lambda.coulomb = lambdaCoulomb;
  g_lambda.sc_alpha = sc_alpha;
  g_lambda.sc_sigma_6 = sc_sigma_6;
  g_lambda.sc_power = sc_power;
  g_lambda.sc_coul = sc_coul;
}
/* CalculateEnergy::GetLambdaVDW / GetLambdaCoulomb, src/CalculateEnergy.cpp:1558-1572 */
static inline double pair_lambda_vdw(int molA, int molB) {
  double l = 1.0;
  l *= (molA == g_lambda.mol ? g_lambda.vdw : 1.0);
  l *= (molB == g_lambda.mol ? g_lambda.vdw : 1.0);
  return l;
}
static inline double pair_lambda_coulomb(int molA, int molB) {
  double l = 1.0;
  l *= (molA == g_lambda.mol ? g_lambda.coulomb : 1.0);
  l *= (molB == g_lambda.mol ? g_lambda.coulomb : 1.0);
  return l;
}
/* Ewald::GetLambdaCoef, src/Ewald.cpp:1598-1602 */
static inline double mol_lambda_coef(int mol) {
  return sqrt(mol == g_lambda.mol ? g_lambda.coulomb : 1.0);
}
/* soft-core r^2, identical text in every functor (e.g. src/FFParticle.cpp:306-311) */
static inline double soft_rsq(const orc_params *p, double distSq, int idx, double lambda) {
  double sigma6 = p->sigmaSq[idx] * p->sigmaSq[idx] * p->sigmaSq[idx];
  sigma6 = sigma6 > g_lambda.sc_sigma_6 ? sigma6 : g_lambda.sc_sigma_6;
  double dist6 = distSq * distSq * distSq;
  double lambdaCoef = g_lambda.sc_alpha * pow((1.0 - lambda), (double)g_lambda.sc_power);
  double softDist6 = lambdaCoef * sigma6 + dist6;
  return cbrt(softDist6);
}
static _Thread_local int g_nocut = 0; /* the two-argument functor forms have no cut-off test */

double orc_calc_en(const orc_params *p, double distSq, int kind1, int kind2) {
  double rCutSq = p->rCut * p->rCut;
  if (!g_nocut && rCutSq < distSq) return 0.0; /* FFParticle.cpp:297 */
  int idx = kind1 + kind2 * p->kindCount; /* FFParticle.h:111 */
  if (p->vdwKind == ORC_VDW_SWITCH && p->isMartini) { /* FFSwitchMartini.h:279-302 */
    martini_consts m = martini_init(p, idx);
    double rOnSq = p->rOn * p->rOn;
    double r_2 = 1.0 / distSq;
    double r_4 = r_2 * r_2;
    double r_6 = r_4 * r_2;
    double n_ij = p->n[idx];
    double r_n = pow(r_2, (n_ij * 0.5));
    double rij_ron = sqrt(distSq) - p->rOn;
    double rij_ron_2 = rij_ron * rij_ron;
    double rij_ron_3 = rij_ron_2 * rij_ron;
    double rij_ron_4 = rij_ron_2 * rij_ron_2;
    double shifttempRep = -(m.An / 3.0) * rij_ron_3 - (m.Bn / 4.0) * rij_ron_4 - m.Cn;
    double shifttempAtt = -(m.A6 / 3.0) * rij_ron_3 - (m.B6 / 4.0) * rij_ron_4 - m.C6;
    double shiftRep = (distSq > rOnSq ? shifttempRep : -m.Cn);
    double shiftAtt = (distSq > rOnSq ? shifttempAtt : -m.C6);
    return p->epsilon_cn[idx] * (m.sign * (r_n + shiftRep) - m.sig6 * (r_6 + shiftAtt));
  }
  if (p->vdwKind == ORC_VDW_EXP6) { /* FFExp6.h:184-224 */
    if (distSq < p->rMaxSq[idx]) return DBL_MAX;
    double dist = sqrt(distSq);
    double rRat = p->rMin[idx] / dist;
    double rRat2 = rRat * rRat;
    double attract = rRat2 * rRat2 * rRat2;
    unsigned alpha_ij = (unsigned)p->n[idx]; /* truncated to uint, FFExp6.h:215 */
    double repulse = (6.0 / alpha_ij) * exp(alpha_ij * (1.0 - dist / p->rMin[idx]));
    return p->expConst[idx] * (repulse - attract);
  }
  switch (p->vdwKind) {
  case ORC_VDW_SHIFT: { /* FFShift.h:167-175 */
    double rRat2 = p->sigmaSq[idx] / distSq;
    double rRat4 = rRat2 * rRat2;
    double attract = rRat4 * rRat2;
    double repulse = pow(rRat2, p->n[idx] * 0.5);
    return p->epsilon_cn[idx] * (repulse - attract) - shift_const(p, idx);
  }
  case ORC_VDW_SWITCH: { /* FFSwitch.h:163-174, factors :101-105 */
    double rOnSq = p->rOn * p->rOn;
    double factor1 = rCutSq - 3 * rOnSq;
    double factor2 =
        1.0 / ((rCutSq - rOnSq) * (rCutSq - rOnSq) * (rCutSq - rOnSq));
    double rCutSq_rijSq = rCutSq - distSq;
    double rCutSq_rijSq_Sq = rCutSq_rijSq * rCutSq_rijSq;
    double rRat2 = p->sigmaSq[idx] / distSq;
    double attract = rRat2 * rRat2 * rRat2;
    double repulse = pow(rRat2, p->n[idx] * 0.5);
    double fE = rCutSq_rijSq_Sq * factor2 * (factor1 + 2.0 * distSq);
    double factE = (distSq > rOnSq ? fE : 1.0);
    return (p->epsilon_cn[idx] * (repulse - attract)) * factE;
  }
  default:
    return mie_en(p, distSq, idx);
  }
}

double orc_calc_vir(const orc_params *p, double distSq, int kind1, int kind2) {
  double rCutSq = p->rCut * p->rCut;
  if (!g_nocut && rCutSq < distSq) return 0.0;
  int idx = kind1 + kind2 * p->kindCount;
  if (p->vdwKind == ORC_VDW_SWITCH && p->isMartini) { /* FFSwitchMartini.h:328-348 */
    martini_consts m = martini_init(p, idx);
    double rOnSq = p->rOn * p->rOn;
    double n_ij = p->n[idx];
    double r_1 = 1.0 / sqrt(distSq);
    double r_8 = distSq * distSq * distSq * distSq; /* sic: (r^2)^4, not r^-8 */
    double r_n2 = pow(r_1, n_ij + 2.0);
    double rij_ron = sqrt(distSq) - p->rOn;
    double rij_ron_2 = rij_ron * rij_ron;
    double rij_ron_3 = rij_ron_2 * rij_ron;
    double dshifttempRep = m.An * rij_ron_2 + m.Bn * rij_ron_3;
    double dshifttempAtt = m.A6 * rij_ron_2 + m.B6 * rij_ron_3;
    double dshiftRep = (distSq > rOnSq ? dshifttempRep * r_1 : 0);
    double dshiftAtt = (distSq > rOnSq ? dshifttempAtt * r_1 : 0);
    return p->epsilon_cn[idx] *
           (m.sign * (n_ij * r_n2 + dshiftRep) - m.sig6 * (6.0 * r_8 + dshiftAtt));
  }
  if (p->vdwKind == ORC_VDW_EXP6) { /* FFExp6.h:226-257 */
    if (distSq < p->rMaxSq[idx]) return DBL_MAX;
    double dist = sqrt(distSq);
    double rRat = p->rMin[idx] / dist;
    double rRat2 = rRat * rRat;
    double attract = rRat2 * rRat2 * rRat2;
    unsigned alpha_ij = (unsigned)p->n[idx];
    double repulse = (dist / p->rMin[idx]) * exp(alpha_ij * (1.0 - dist / p->rMin[idx]));
    return 6.0 * p->expConst[idx] * (repulse - attract) / distSq;
  }
  if (p->vdwKind == ORC_VDW_SWITCH) { /* FFSwitch.h:199-218 */
    double rOnSq = p->rOn * p->rOn;
    double factor1 = rCutSq - 3 * rOnSq;
    double factor2 =
        1.0 / ((rCutSq - rOnSq) * (rCutSq - rOnSq) * (rCutSq - rOnSq));
    double rCutSq_rijSq = rCutSq - distSq;
    double rCutSq_rijSq_Sq = rCutSq_rijSq * rCutSq_rijSq;
    double rNeg2 = 1.0 / distSq;
    double rRat2 = rNeg2 * p->sigmaSq[idx];
    double attract = rRat2 * rRat2 * rRat2;
    double repulse = pow(rRat2, p->n[idx] * 0.5);
    double fE = rCutSq_rijSq_Sq * factor2 * (factor1 + 2.0 * distSq);
    double fW = 12.0 * factor2 * rCutSq_rijSq * (rOnSq - distSq);
    double factE = (distSq > rOnSq ? fE : 1.0);
    double factW = (distSq > rOnSq ? fW : 0.0);
    double Wij = (p->epsilon_cn[idx] * 6) *
                 ((p->n[idx] / 6) * repulse - attract) * rNeg2;
    double Eij = p->epsilon_cn[idx] * (repulse - attract);
    return Wij * factE - Eij * factW;
  }
  /* STD and SHIFT share the virial (FFShift.h:200-210) */
  return mie_vir(p, distSq, idx);
}

double orc_calc_coulomb(const orc_params *p, double distSq,
                        double qi_qj_fact) {
  double rcc2 = p->rCutCoulomb * p->rCutCoulomb;
  if (!g_nocut && rcc2 < distSq) return 0.0; /* FFParticle.cpp:364 */
  double dist = sqrt(distSq);
  if (p->ewald) { /* FFParticle.cpp:388-394, FFShift.h:239-244, FFSwitch.h:247-252 */
    double val = p->alpha * dist;
    return qi_qj_fact * erfc(val) / dist;
  }
  if (p->vdwKind == ORC_VDW_SWITCH && p->isMartini) { /* FFSwitchMartini.h:379-396 */
    martini_consts m = martini_init(p, 0);
    double rij_ronCoul_3 = dist * distSq;
    double rij_ronCoul_4 = distSq * distSq;
    double coul = -(m.A1 / 3.0) * rij_ronCoul_3 - (m.B1 / 4.0) * rij_ronCoul_4 - m.C1;
    return qi_qj_fact * p->diElectric_1 * (1.0 / dist + coul);
  }
  switch (p->vdwKind) {
  case ORC_VDW_SHIFT: /* FFShift.h:245-249 */
    return qi_qj_fact * (1.0 / dist - 1.0 / p->rCut);
  case ORC_VDW_SWITCH: { /* FFSwitch.h:253-259 */
    double switchVal = distSq / (p->rCut * p->rCut) - 1.0;
    switchVal *= switchVal;
    return qi_qj_fact * switchVal / dist;
  }
  default: /* FFParticle.cpp:395-398 */
    return qi_qj_fact / dist;
  }
}

double orc_calc_coulomb_vir(const orc_params *p, double distSq, double qi_qj) {
  double rcc2 = p->rCutCoulomb * p->rCutCoulomb;
  if (!g_nocut && rcc2 < distSq) return 0.0;
  double dist = sqrt(distSq);
  if (p->ewald) {
    double constValue = p->alpha * M_2_SQRTPI;
    double expConstValue = exp(-1.0 * (p->alpha * p->alpha) * distSq);
    double temp;
    if (p->vdwKind == ORC_VDW_STD)
      temp = 1.0 - erf(p->alpha * dist); /* FFParticle.cpp:439 */
    else
      temp = erfc(p->alpha * dist); /* FFShift.h:289, FFSwitch.h:299 */
    return qi_qj * (temp / dist + constValue * expConstValue) / distSq;
  }
  if (p->vdwKind == ORC_VDW_SWITCH && p->isMartini) { /* FFSwitchMartini.h:440-447 */
    martini_consts m = martini_init(p, 0);
    double rij_ronCoul_2 = distSq;
    double rij_ronCoul_3 = dist * distSq;
    double virCoul = m.A1 / rij_ronCoul_2 + m.B1 / rij_ronCoul_3;
    return qi_qj * p->diElectric_1 * (1.0 / (dist * distSq) + virCoul / dist);
  }
  if (p->vdwKind == ORC_VDW_SWITCH) { /* FFSwitch.h:302-307 */
    double rCutSq = p->rCut * p->rCut;
    double switchVal = distSq / rCutSq - 1.0;
    switchVal *= switchVal;
    double dSwitchVal = 2.0 * (distSq / rCutSq - 1.0) * 2.0 * dist / rCutSq;
    return -qi_qj * (dSwitchVal / distSq - switchVal / (distSq * dist));
  }
  return qi_qj / (distSq * dist);
}

/* The lambda-taking functor forms (src/FFParticle.cpp:295-315, :327-348, :360-386,
 * :400-429; the subclasses repeat them, FF_EXP6 tests rMaxSq first, src/FFExp6.h:184-211,
 * :226-252). */
static double calc_en_l(const orc_params *p, double distSq, int k1, int k2, double lambda) {
  if (p->rCut * p->rCut < distSq) return 0.0;
  int idx = k1 + k2 * p->kindCount;
  if (p->vdwKind == ORC_VDW_EXP6 && distSq < p->rMaxSq[idx]) return DBL_MAX;
  double r;
  g_nocut = 1;
  if (lambda >= 0.999999)
    r = orc_calc_en(p, distSq, k1, k2);
  else
    r = lambda * orc_calc_en(p, soft_rsq(p, distSq, idx, lambda), k1, k2);
  g_nocut = 0;
  return r;
}
static double calc_vir_l(const orc_params *p, double distSq, int k1, int k2, double lambda) {
  if (p->rCut * p->rCut < distSq) return 0.0;
  int idx = k1 + k2 * p->kindCount;
  if (p->vdwKind == ORC_VDW_EXP6 && distSq < p->rMaxSq[idx]) return DBL_MAX;
  double r;
  g_nocut = 1;
  if (lambda >= 0.999999) {
    r = orc_calc_vir(p, distSq, k1, k2);
  } else {
    double softRsq = soft_rsq(p, distSq, idx, lambda);
    double correction = distSq / softRsq;
    r = lambda * correction * correction * orc_calc_vir(p, softRsq, k1, k2);
  }
  g_nocut = 0;
  return r;
}
static double calc_coulomb_l(const orc_params *p, double distSq, int k1, int k2,
                             double qi_qj_fact, double lambda) {
  if (p->rCutCoulomb * p->rCutCoulomb < distSq) return 0.0;
  double r;
  g_nocut = 1;
  if (lambda >= 0.999999)
    r = orc_calc_coulomb(p, distSq, qi_qj_fact);
  else if (g_lambda.sc_coul)
    r = lambda * orc_calc_coulomb(p, soft_rsq(p, distSq, k1 + k2 * p->kindCount, lambda),
                                  qi_qj_fact);
  else
    r = lambda * orc_calc_coulomb(p, distSq, qi_qj_fact);
  g_nocut = 0;
  return r;
}
static double calc_coulomb_vir_l(const orc_params *p, double distSq, int k1, int k2,
                                 double qi_qj, double lambda) {
  if (p->rCutCoulomb * p->rCutCoulomb < distSq) return 0.0;
  double r;
  g_nocut = 1;
  if (lambda >= 0.999999) {
    r = orc_calc_coulomb_vir(p, distSq, qi_qj);
  } else if (g_lambda.sc_coul) {
    double softRsq = soft_rsq(p, distSq, k1 + k2 * p->kindCount, lambda);
    double correction = distSq / softRsq;
    r = lambda * correction * correction * orc_calc_coulomb_vir(p, softRsq, qi_qj);
  } else {
    r = lambda * orc_calc_coulomb_vir(p, distSq, qi_qj);
  }
  g_nocut = 0;
  return r;
}

/* ------------------------------------------------------------------------ */
/* Cell list                                                                 */

int orc_cell_edges(const orc_params *p, int edge[3]) {
  double cutoff = box_rcut(p); /* CellList::SetCutoff, CellList.cpp:61-65 */
  for (int d = 0; d < 3; ++d) { /* CellList::ResizeGrid, CellList.cpp:138-163 */
    int e = (int)floor(p->axis[d] / cutoff);
    edge[d] = e > 3 ? e : 3;
  }
  return edge[0] * edge[1] * edge[2];
}

/* CellList::PositionToCell, src/CellList.h:88-101 (orthogonal: unslant = id) */
static const orc_params *g_cellParams = 0; /* set by csr_make / orc_cell_list_build */
static inline int position_to_cell(const double cellSize[3], const int edge[3],
                                   double px, double py, double pz) {
  if (g_cellParams && g_cellParams->nonOrth) { /* TransformUnSlant first */
    double a[3] = {px, py, pz}, u[3];
    vec_mat(a, g_cellParams->cellBasisInv, u);
    px = u[0];
    py = u[1];
    pz = u[2];
  }
  int cx = (int)(px / cellSize[0]);
  int cy = (int)(py / cellSize[1]);
  int cz = (int)(pz / cellSize[2]);
  cx -= (cx == edge[0] ? 1 : 0);
  cy -= (cy == edge[1] ? 1 : 0);
  cz -= (cz == edge[2] ? 1 : 0);
  return cx * edge[1] * edge[2] + cy * edge[2] + cz;
}

/* CellList::RebuildNeighbors, src/CellList.cpp:191-218 */
static void build_neighbors(const int edge[3], int *neighborList) {
  for (int cx = 0; cx < edge[0]; ++cx)
    for (int cy = 0; cy < edge[1]; ++cy)
      for (int cz = 0; cz < edge[2]; ++cz) {
        int cell = cx * edge[2] * edge[1] + cy * edge[2] + cz;
        int k = 0;
        for (int dx = -1; dx <= 1; ++dx)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz)
              neighborList[cell * 27 + k++] =
                  ((cx + dx + edge[0]) % edge[0]) * edge[2] * edge[1] +
                  ((cy + dy + edge[1]) % edge[1]) * edge[2] +
                  ((cz + dz + edge[2]) % edge[2]);
      }
}

int orc_cell_list_build(const orc_params *p, int nAtomsTotal, const double *x,
                        const double *y, const double *z, const int *boxAtoms,
                        int nBox, int *cellVector, int *cellStart,
                        int *mapParticleToCell, int *neighborList) {
  int edge[3];
  int nCells = orc_cell_edges(p, edge);
  g_cellParams = p;
  double cellSize[3] = {p->axis[0] / edge[0], p->axis[1] / edge[1],
                        p->axis[2] / edge[2]};
  /* counting sort by cell, scanning atoms in ascending index order: gives
   * the ascending-within-cell order of GetCellListNeighbor's std::sort
   * (CellList.cpp:258-285) whatever the linked-list insertion order was. */
  int *sorted = (int *)malloc(sizeof(int) * (size_t)(nBox > 0 ? nBox : 1));
  memcpy(sorted, boxAtoms, sizeof(int) * (size_t)nBox);
  /* boxAtoms need not be sorted: sort ascending (insertion into CSR) */
  for (int i = 1; i < nBox; ++i) { /* usually already sorted: O(n) */
    int v = sorted[i], j = i - 1;
    while (j >= 0 && sorted[j] > v) {
      sorted[j + 1] = sorted[j];
      --j;
    }
    sorted[j + 1] = v;
  }
  for (int i = 0; i < nAtomsTotal; ++i) mapParticleToCell[i] = -1;
  for (int c = 0; c <= nCells; ++c) cellStart[c] = 0;
  for (int i = 0; i < nBox; ++i) {
    int a = sorted[i];
    int c = position_to_cell(cellSize, edge, x[a], y[a], z[a]);
    mapParticleToCell[a] = c;
    cellStart[c + 1]++;
  }
  for (int c = 0; c < nCells; ++c) cellStart[c + 1] += cellStart[c];
  int *fill = (int *)calloc((size_t)nCells, sizeof(int));
  for (int i = 0; i < nBox; ++i) {
    int a = sorted[i];
    int c = mapParticleToCell[a];
    cellVector[cellStart[c] + fill[c]++] = a;
  }
  free(fill);
  free(sorted);
  build_neighbors(edge, neighborList);
  return nCells;
}

typedef struct {
  int nCells;
  int *cellVector, *cellStart, *map, *nbr;
} cell_csr;

static cell_csr csr_make(const orc_params *p, int nAtomsTotal, const double *x,
                         const double *y, const double *z, const int *boxAtoms,
                         int nBox) {
  cell_csr c;
  int edge[3];
  int nCells = orc_cell_edges(p, edge);
  c.cellVector = (int *)malloc(sizeof(int) * (size_t)(nBox + 1));
  c.cellStart = (int *)malloc(sizeof(int) * (size_t)(nCells + 1));
  c.map = (int *)malloc(sizeof(int) * (size_t)(nAtomsTotal + 1));
  c.nbr = (int *)malloc(sizeof(int) * (size_t)nCells * 27);
  c.nCells = orc_cell_list_build(p, nAtomsTotal, x, y, z, boxAtoms, nBox,
                                 c.cellVector, c.cellStart, c.map, c.nbr);
  return c;
}
static void csr_free(cell_csr *c) {
  free(c->cellVector);
  free(c->cellStart);
  free(c->map);
  free(c->nbr);
}

/* ------------------------------------------------------------------------ */
/* BoxInter                                                                  */

int orc_box_inter(const orc_params *p, int nAtomsTotal, const double *x,
                  const double *y, const double *z, const int *kind,
                  const int *mol, const double *charge, const int *boxAtoms,
                  int nBox, double *ljEn, double *realEn) {
  cell_csr c = csr_make(p, nAtomsTotal, x, y, z, boxAtoms, nBox);
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double tempREn = 0.0, tempLJEn = 0.0;
  /* src/CalculateEnergy.cpp:199-250 */
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : tempREn, tempLJEn) schedule(static)
#endif
  for (int ci = 0; ci < nBox; ++ci) {
    int cur = c.cellVector[ci];
    int curCell = c.map[cur];
    for (int nc = 0; nc < 27; ++nc) {
      int nbrCell = c.nbr[curCell * 27 + nc];
      int end = c.cellStart[nbrCell + 1];
      for (int ni = c.cellStart[nbrCell]; ni < end; ++ni) {
        int nb = c.cellVector[ni];
        if (cur < nb && mol[cur] != mol[nb]) {
          double distSq, d[3];
          if (in_rcut(p, boxRcutSq, x[cur], y[cur], z[cur], x[nb], y[nb],
                      z[nb], &distSq, d)) {
            if (p->electrostatic) {
              double qi_qj_fact = charge[cur] * charge[nb] * ORC_QQFACT;
              if (qi_qj_fact != 0.0)
                tempREn += calc_coulomb_l(p, distSq, kind[cur], kind[nb], qi_qj_fact,
                                          pair_lambda_coulomb(mol[cur], mol[nb]));
            }
            tempLJEn += calc_en_l(p, distSq, kind[cur], kind[nb],
                                  pair_lambda_vdw(mol[cur], mol[nb]));
          }
        }
      }
    }
  }
  *ljEn = tempLJEn;
  *realEn = tempREn;
  csr_free(&c);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* BoxForce                                                                  */

int orc_box_force(const orc_params *p, int nAtomsTotal, int nMols,
                  const double *x, const double *y, const double *z,
                  const int *kind, const int *mol, const double *charge,
                  const int *boxAtoms, int nBox, double *ljEn, double *realEn,
                  double *aFx, double *aFy, double *aFz, double *mFx,
                  double *mFy, double *mFz) {
  (void)nMols;
  cell_csr c = csr_make(p, nAtomsTotal, x, y, z, boxAtoms, nBox);
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double tempREn = 0.0, tempLJEn = 0.0;
  /* ResetForce, src/CalculateEnergy.cpp:1408-1428 */
  for (int i = 0; i < nBox; ++i) {
    int a = boxAtoms[i];
    aFx[a] = aFy[a] = aFz[a] = 0.0;
    mFx[mol[a]] = mFy[mol[a]] = mFz[mol[a]] = 0.0;
  }
  /* src/CalculateEnergy.cpp:336-394; serial so that the force accumulation
   * order is the reference's +p1 order. */
  for (int ci = 0; ci < nBox; ++ci) {
    int cur = c.cellVector[ci];
    int curCell = c.map[cur];
    for (int nc = 0; nc < 27; ++nc) {
      int nbrCell = c.nbr[curCell * 27 + nc];
      int end = c.cellStart[nbrCell + 1];
      for (int ni = c.cellStart[nbrCell]; ni < end; ++ni) {
        int nb = c.cellVector[ni];
        if (cur < nb && mol[cur] != mol[nb]) {
          double distSq, d[3];
          if (in_rcut(p, boxRcutSq, x[cur], y[cur], z[cur], x[nb], y[nb],
                      z[nb], &distSq, d)) {
            double fR[3] = {0.0, 0.0, 0.0}, fL[3];
            if (p->electrostatic) {
              double qi_qj_fact = charge[cur] * charge[nb] * ORC_QQFACT;
              if (qi_qj_fact != 0.0) {
                double lc = pair_lambda_coulomb(mol[cur], mol[nb]);
                tempREn += calc_coulomb_l(p, distSq, kind[cur], kind[nb], qi_qj_fact, lc);
                double v = calc_coulomb_vir_l(p, distSq, kind[cur], kind[nb], qi_qj_fact, lc);
                fR[0] = d[0] * v;
                fR[1] = d[1] * v;
                fR[2] = d[2] * v;
              }
            }
            double lv = pair_lambda_vdw(mol[cur], mol[nb]);
            tempLJEn += calc_en_l(p, distSq, kind[cur], kind[nb], lv);
            double w = calc_vir_l(p, distSq, kind[cur], kind[nb], lv);
            fL[0] = d[0] * w;
            fL[1] = d[1] * w;
            fL[2] = d[2] * w;
            aFx[cur] += fL[0] + fR[0];
            aFy[cur] += fL[1] + fR[1];
            aFz[cur] += fL[2] + fR[2];
            aFx[nb] += -(fL[0] + fR[0]);
            aFy[nb] += -(fL[1] + fR[1]);
            aFz[nb] += -(fL[2] + fR[2]);
            mFx[mol[cur]] += (fL[0] + fR[0]);
            mFy[mol[cur]] += (fL[1] + fR[1]);
            mFz[mol[cur]] += (fL[2] + fR[2]);
            mFx[mol[nb]] += -(fL[0] + fR[0]);
            mFy[mol[nb]] += -(fL[1] + fR[1]);
            mFz[mol[nb]] += -(fL[2] + fR[2]);
          }
        }
      }
    }
  }
  *ljEn = tempLJEn;
  *realEn = tempREn;
  csr_free(&c);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Virial                                                                    */

/* BoxDimensions::UnwrapPBC (scalar), src/BoxDimensions.cpp:297-320 */
static inline double unwrap_scalar(double v, double ref, double ax, double halfAx) {
  if (fabs(ref - v) > halfAx) {
    if (ref < halfAx)
      v -= ax;
    else
      v += ax;
  }
  return v;
}
/* BoxDimensions::UnwrapPBC(x,y,z,b,ref) and the non-orthogonal override
 * (src/BoxDimensionsNonOrth.cpp:305-319). */
static inline void unwrap_vec(const orc_params *p, double v[3], const double ref[3]) {
  if (p->nonOrth) {
    double u[3], ur[3];
    vec_mat(v, p->cellBasisInv, u);
    vec_mat(ref, p->cellBasisInv, ur);
    for (int d = 0; d < 3; ++d)
      u[d] = unwrap_scalar(u[d], ur[d], p->axis[d], p->axis[d] * 0.5);
    vec_mat(u, p->cellBasis, v);
  } else {
    for (int d = 0; d < 3; ++d)
      v[d] = unwrap_scalar(v[d], ref[d], p->axis[d], p->axis[d] * 0.5);
  }
}

int orc_virial_calc(const orc_params *p, int nAtomsTotal, const double *x,
                    const double *y, const double *z, const int *kind,
                    const int *mol, const double *charge, const double *comX,
                    const double *comY, const double *comZ,
                    const int *boxAtoms, int nBox, double vT[3], double rT[3]) {
  cell_csr c = csr_make(p, nAtomsTotal, x, y, z, boxAtoms, nBox);
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double vT11 = 0.0, vT22 = 0.0, vT33 = 0.0, rT11 = 0.0, rT22 = 0.0, rT33 = 0.0;
  for (int ci = 0; ci < nBox; ++ci) { /* src/CalculateEnergy.cpp:464-531 */
    int cur = c.cellVector[ci];
    int curCell = c.map[cur];
    for (int nc = 0; nc < 27; ++nc) {
      int nbrCell = c.nbr[curCell * 27 + nc];
      int end = c.cellStart[nbrCell + 1];
      for (int ni = c.cellStart[nbrCell]; ni < end; ++ni) {
        int nb = c.cellVector[ni];
        if (cur < nb && mol[cur] != mol[nb]) {
          double distSq, d[3];
          if (in_rcut(p, boxRcutSq, x[cur], y[cur], z[cur], x[nb], y[nb],
                      z[nb], &distSq, d)) {
            double cc[3] = {comX[mol[cur]] - comX[mol[nb]],
                            comY[mol[cur]] - comY[mol[nb]],
                            comZ[mol[cur]] - comZ[mol[nb]]};
            min_image_vec(p, cc);
            if (p->electrostatic) {
              double qi_qj = charge[cur] * charge[nb];
              if (qi_qj != 0.0) {
                double pRF = calc_coulomb_vir_l(p, distSq, kind[cur], kind[nb], qi_qj,
                                                pair_lambda_coulomb(mol[cur], mol[nb]));
                rT11 += pRF * (d[0] * cc[0]);
                rT22 += pRF * (d[1] * cc[1]);
                rT33 += pRF * (d[2] * cc[2]);
              }
            }
            double pVF = calc_vir_l(p, distSq, kind[cur], kind[nb],
                                    pair_lambda_vdw(mol[cur], mol[nb]));
            vT11 += pVF * (d[0] * cc[0]);
            vT22 += pVF * (d[1] * cc[1]);
            vT33 += pVF * (d[2] * cc[2]);
          }
        }
      }
    }
  }
  vT[0] = vT11;
  vT[1] = vT22;
  vT[2] = vT33;
  rT[0] = rT11 * ORC_QQFACT; /* :555-567 */
  rT[1] = rT22 * ORC_QQFACT;
  rT[2] = rT33 * ORC_QQFACT;
  csr_free(&c);
  return 0;
}

int orc_virial_reciprocal(const orc_params *p, int nBoxMols, const int *boxMols,
                          const int *molStart, const double *x, const double *y,
                          const double *z, const double *charge,
                          const double *comX, const double *comY,
                          const double *comZ, int nk, const double *kx,
                          const double *ky, const double *kz,
                          const double *hsqr, const double *prefact,
                          const double *sumRref, const double *sumIref,
                          double wT[3]) {
  double wT11 = 0.0, wT22 = 0.0, wT33 = 0.0;
  double constVal = 1.0 / (4.0 * (p->alpha * p->alpha)); /* src/Ewald.cpp:1177 */
  for (int i = 0; i < nk; ++i) {                         /* :1229-1244 */
    double factor =
        prefact[i] * (sumRref[i] * sumRref[i] + sumIref[i] * sumIref[i]);
    wT11 += factor * (1.0 - 2.0 * (constVal + 1.0 / hsqr[i]) * kx[i] * kx[i]);
    wT22 += factor * (1.0 - 2.0 * (constVal + 1.0 / hsqr[i]) * ky[i] * ky[i]);
    wT33 += factor * (1.0 - 2.0 * (constVal + 1.0 / hsqr[i]) * kz[i] * kz[i]);
  }
  for (int mi = 0; mi < nBoxMols; ++mi) { /* intramolecular part, :1247-1285 */
    int m = boxMols[mi];
    double com[3] = {comX[m], comY[m], comZ[m]};
    for (int a = molStart[m]; a < molStart[m + 1]; ++a) {
      if (fabs(charge[a]) < 0.000000001) continue;
      double at[3] = {x[a], y[a], z[a]};
      unwrap_vec(p, at, com);
      double diff[3] = {at[0] - com[0], at[1] - com[1], at[2] - com[2]};
      double q = charge[a] * mol_lambda_coef(m);
      /* one OpenMP reduction region per atom (:1272-1284): the k sum goes
       * into a zero-initialised private copy that is then added to wT */
      double p11 = 0.0, p22 = 0.0, p33 = 0.0;
      for (int i = 0; i < nk; ++i) {
        double arg = x[a] * kx[i] + y[a] * ky[i] + z[a] * kz[i];
        double factor =
            prefact[i] * 2.0 * (sumIref[i] * cos(arg) - sumRref[i] * sin(arg)) * q;
        p11 += factor * (kx[i] * diff[0]);
        p22 += factor * (ky[i] * diff[1]);
        p33 += factor * (kz[i] * diff[2]);
      }
      wT11 += p11;
      wT22 += p22;
      wT33 += p33;
    }
  }
  wT[0] = wT11;
  wT[1] = wT22;
  wT[2] = wT33;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* MultiParticle move: counter-based RNG, trial transform, acceptance weight  */

/* Philox4x64-10 (Salmon, Moraes, Dror, Shaw, SC'11), as instantiated by
 * lib/Random123/philox.h:229-251,275 (multipliers, Weyl key increments, 10 rounds). */
static inline void philox_mulhilo(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo) {
  unsigned __int128 pr = (unsigned __int128)a * b;
  *hi = (uint64_t)(pr >> 64);
  *lo = (uint64_t)pr;
}
void orc_philox4x64_10(const uint64_t ctrIn[4], const uint64_t keyIn[2], uint64_t out[4]) {
  uint64_t c[4] = {ctrIn[0], ctrIn[1], ctrIn[2], ctrIn[3]};
  uint64_t k[2] = {keyIn[0], keyIn[1]};
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { /* bumpkey */
      k[0] += 0x9E3779B97F4A7C15ULL;
      k[1] += 0xBB67AE8584CAA73BULL;
    }
    uint64_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2E7470EE14C6C93ULL, c[0], &hi0, &lo0);
    philox_mulhilo(0xCA5A826395121157ULL, c[2], &hi1, &lo1);
    uint64_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
/* r123::u01<double>(uint64) and r123::uneg11<double>(uint64),
 * lib/Random123/uniform.hpp:175-184, :206-215 */
static inline double r123_u01(uint64_t in) {
  const double factor = 1.0 / (18446744073709551615.0 + 1.0);
  return (double)in * factor + 0.5 * factor;
}
static inline double r123_uneg11(uint64_t in) {
  const double factor = 1.0 / (9223372036854775807.0 + 1.0);
  return (double)(int64_t)in * factor + 0.5 * factor;
}
/* Random123Wrapper::getRNG, src/Random123Wrapper.cpp:16-22 */
static inline void mp_rng(uint64_t counter, uint64_t keyValue, uint64_t step,
                          uint64_t seed, uint64_t r[4]) {
  uint64_t c[4] = {counter, keyValue, 0, 0}, k[2] = {step, seed};
  orc_philox4x64_10(c, k, r);
}

/* BoxDimensions::WrapPBC (scalar), src/BoxDimensions.cpp:261-295, and the
 * non-orthogonal override, src/BoxDimensionsNonOrth.cpp:268-281 */
static inline double wrap_scalar(double v, double ax) {
  if (v >= ax)
    v -= ax;
  else if (v < 0)
    v += ax;
  return v;
}
static inline void wrap_vec(const orc_params *p, double v[3]) {
  if (p->nonOrth) {
    double u[3];
    vec_mat(v, p->cellBasisInv, u);
    for (int d = 0; d < 3; ++d) u[d] = wrap_scalar(u[d], p->axis[d]);
    vec_mat(u, p->cellBasis, v);
  } else {
    for (int d = 0; d < 3; ++d) v[d] = wrap_scalar(v[d], p->axis[d]);
  }
}

/* TransformMatrix::FromAxisAngle(theta, CrossProduct(axis), TensorProduct(axis)),
 * src/TransformMatrix.h:165-214 */
static void axis_angle(double theta, const double ax[3], double m[3][3]) {
  double cross[3][3] = {{0.0, -(ax[2]), ax[1]}, {ax[2], 0.0, -(ax[0])}, {-(ax[1]), ax[0], 0.0}};
  double tens[3][3];
  for (int i = 0; i < 3; ++i) {
    tens[0][i] = ax[0];
    tens[1][i] = ax[1];
    tens[2][i] = ax[2];
  }
  for (int i = 0; i < 3; ++i) {
    tens[i][0] *= ax[0];
    tens[i][1] *= ax[1];
    tens[i][2] *= ax[2];
  }
  double c = cos(theta), s = sin(theta);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      m[i][j] = (i == j) ? c : 0.0;
      m[i][j] += s * cross[i][j] + (1 - c) * tens[i][j];
    }
}

/* rotate the atoms of molecule [s,e) about its COM: MultiParticle::RotateForceBiased
 * / RotateRandom, src/moves/MultiParticle.h:615-643, :667-692 */
static void mp_rotate_mol(const orc_params *p, int s, int e, const double m[3][3],
                          const double com[3], double *nx, double *ny, double *nz) {
  for (int a = s; a < e; ++a) {
    double t[3] = {nx[a], ny[a], nz[a]};
    unwrap_vec(p, t, com);
    t[0] += -com[0];
    t[1] += -com[1];
    t[2] += -com[2];
    double r[3] = {m[0][0] * t[0] + m[0][1] * t[1] + m[0][2] * t[2],
                   m[1][0] * t[0] + m[1][1] * t[1] + m[1][2] * t[2],
                   m[2][0] * t[0] + m[2][1] * t[1] + m[2][2] * t[2]};
    r[0] += com[0];
    r[1] += com[1];
    r[2] += com[2];
    wrap_vec(p, r);
    nx[a] = r[0];
    ny[a] = r[1];
    nz[a] = r[2];
  }
}
static void mp_translate_mol(const orc_params *p, int s, int e, const double sh[3], int m,
                             double *nx, double *ny, double *nz, double *ncx, double *ncy,
                             double *ncz) { /* :645-665, :694-715 */
  for (int a = s; a < e; ++a) {
    double t[3] = {nx[a] + sh[0], ny[a] + sh[1], nz[a] + sh[2]};
    wrap_vec(p, t);
    nx[a] = t[0];
    ny[a] = t[1];
    nz[a] = t[2];
  }
  double c[3] = {ncx[m] + sh[0], ncy[m] + sh[1], ncz[m] + sh[2]};
  wrap_vec(p, c);
  ncx[m] = c[0];
  ncy[m] = c[1];
  ncz[m] = c[2];
}

int orc_mp_transform(const orc_params *p, int moveType, int nBoxMols, const int *boxMols,
                     const int *molStart, const double *fx, const double *fy,
                     const double *fz, const double *rfx, const double *rfy,
                     const double *rfz, double max, double lambda, double beta,
                     uint64_t step, uint64_t seed, uint64_t keyValue, double *kX,
                     double *kY, double *kZ, int *inForceRange, double *nx, double *ny,
                     double *nz, double *ncx, double *ncy, double *ncz) {
  for (int mi = 0; mi < nBoxMols; ++mi) { /* CalculateTrialDistRot, :566-613 */
    int m = boxMols[mi];
    double f[3] = {fx[m], fy[m], fz[m]};
    if (rfx) { /* displace: molForceRef + molForceRecRef */
      f[0] += rfx[m];
      f[1] += rfy[m];
      f[2] += rfz[m];
    }
    double lb[3] = {f[0] * lambda * beta, f[1] * lambda * beta, f[2] * lambda * beta};
    /* CalcRandomTransform, :545-564 */
    double lbmax[3] = {lb[0] * max, lb[1] * max, lb[2] * max};
    double val[3] = {0.0, 0.0, 0.0};
    int inRange = fabs(lbmax[0]) > 1E-12 && fabs(lbmax[0]) < 30 && fabs(lbmax[1]) > 1E-12 &&
                  fabs(lbmax[1]) < 30 && fabs(lbmax[2]) > 1E-12 && fabs(lbmax[2]) < 30;
    uint64_t r[4];
    mp_rng((uint64_t)m, keyValue, step, seed, r);
    if (inRange) {
      for (int d = 0; d < 3; ++d)
        val[d] = log(exp(-1.0 * lbmax[d]) + 2.0 * r123_u01(r[d]) * sinh(lbmax[d])) / lb[d];
    }
    kX[m] = val[0];
    kY[m] = val[1];
    kZ[m] = val[2];
    inForceRange[m] = inRange;
    double com[3] = {ncx[m], ncy[m], ncz[m]};
    if (moveType == 1) { /* mp::MPROTATE */
      double mat[3][3];
      if (inRange) {
        double rotLen = sqrt(val[0] * val[0] + val[1] * val[1] + val[2] * val[2]);
        double inv = 1.0 / rotLen;
        double axis[3] = {val[0] * inv, val[1] * inv, val[2] * inv};
        axis_angle(rotLen, axis, mat);
      } else {
        double symRand = max * r123_uneg11(r[0]);
        double u = r123_uneg11(r[1]);
        double theta = 2.0 * M_PI * r123_u01(r[2]);
        double rootTerm = sqrt(1.0 - u * u);
        double axis[3] = {rootTerm * cos(theta), rootTerm * sin(theta), u};
        axis_angle(symRand, axis, mat);
      }
      mp_rotate_mol(p, molStart[m], molStart[m + 1], mat, com, nx, ny, nz);
    } else {
      double sh[3] = {val[0], val[1], val[2]};
      if (!inRange)
        for (int d = 0; d < 3; ++d) sh[d] = max * r123_uneg11(r[d]);
      mp_translate_mol(p, molStart[m], molStart[m + 1], sh, m, nx, ny, nz, ncx, ncy, ncz);
    }
  }
  return 0;
}

/* MultiParticleBrownian::CalculateTrialDistRot, src/moves/MultiParticleBrownianMotion.h:
 * val = (force or torque) * BETA * max + N(0, sqrt(2 max)) with Random123Wrapper::
 * GetGaussianCoords (two Box-Muller pairs, lib/Random123/boxmuller.hpp:126-137), then the
 * same rigid translate / rotate as above; no force range test. */
int orc_bm_transform(const orc_params *p, int moveType, int nBoxMols, const int *boxMols,
                     const int *molStart, const double *fx, const double *fy,
                     const double *fz, const double *rfx, const double *rfy,
                     const double *rfz, double max, double beta, uint64_t step,
                     uint64_t seed, uint64_t keyValue, double *kX, double *kY, double *kZ,
                     double *nx, double *ny, double *nz, double *ncx, double *ncy,
                     double *ncz) {
  const double PI = 3.1415926535897932; /* boxmuller.hpp:104 */
  for (int mi = 0; mi < nBoxMols; ++mi) {
    int m = boxMols[mi];
    double f[3] = {fx[m], fy[m], fz[m]};
    if (rfx) {
      f[0] += rfx[m];
      f[1] += rfy[m];
      f[2] += rfz[m];
    }
    double lb[3] = {f[0] * beta, f[1] * beta, f[2] * beta};
    double stdDev = sqrt(2.0 * max);
    uint64_t r[4];
    mp_rng((uint64_t)m, keyValue, step, seed, r);
    double a0 = PI * r123_uneg11(r[0]), a1 = PI * r123_uneg11(r[2]);
    double rad0 = sqrt(-2. * log(r123_u01(r[1]))), rad1 = sqrt(-2. * log(r123_u01(r[3])));
    double g[3] = {sin(a0) * rad0, cos(a0) * rad0, sin(a1) * rad1};
    double val[3];
    for (int d = 0; d < 3; ++d) val[d] = lb[d] * max + (0.0 + g[d] * stdDev);
    kX[m] = val[0];
    kY[m] = val[1];
    kZ[m] = val[2];
    double com[3] = {ncx[m], ncy[m], ncz[m]};
    if (moveType == 1) {
      double rotLen = sqrt(val[0] * val[0] + val[1] * val[1] + val[2] * val[2]);
      double inv = 1.0 / rotLen;
      double axis[3] = {val[0] * inv, val[1] * inv, val[2] * inv};
      double mat[3][3];
      axis_angle(rotLen, axis, mat);
      mp_rotate_mol(p, molStart[m], molStart[m + 1], mat, com, nx, ny, nz);
    } else {
      mp_translate_mol(p, molStart[m], molStart[m + 1], val, m, nx, ny, nz, ncx, ncy, ncz);
    }
  }
  return 0;
}

/* MultiParticleBrownian::GetCoeff / CalculateWRatio (same file, :415-472): a sum (the
 * logarithm of the weight ratio), one OpenMP + reduction. */
double orc_bm_coeff(int nBoxMols, const int *boxMols, const double *ofx, const double *ofy,
                    const double *ofz, const double *orx, const double *ory,
                    const double *orz, const double *nfx, const double *nfy,
                    const double *nfz, const double *nrx, const double *nry,
                    const double *nrz, const double *kX, const double *kY,
                    const double *kZ, double max, double beta) {
  double max4 = 4.0 * max;
  double priv = 0.0;
  for (int mi = 0; mi < nBoxMols; ++mi) {
    int m = boxMols[mi];
    double o[3] = {ofx[m], ofy[m], ofz[m]}, n[3] = {nfx[m], nfy[m], nfz[m]};
    if (orx) {
      o[0] += orx[m]; o[1] += ory[m]; o[2] += orz[m];
      n[0] += nrx[m]; n[1] += nry[m]; n[2] += nrz[m];
    }
    double k[3] = {kX[m], kY[m], kZ[m]};
    double ov[3], nv[3];
    for (int d = 0; d < 3; ++d) {
      ov[d] = o[d] * beta * max - k[d];
      nv[d] = n[d] * beta * max + k[d];
    }
    double w = 0.0;
    w -= ((nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]) / max4);
    w += ((ov[0] * ov[0] + ov[1] * ov[1] + ov[2] * ov[2]) / max4);
    priv += w;
  }
  double w_ratio = 0.0;
  w_ratio += priv;
  return w_ratio;
}

/* MultiParticle::CalculateWRatio / GetCoeff, src/moves/MultiParticle.h:443-513
 * (one OpenMP product reduction: private copy starts at 1, then multiplies w_ratio) */
double orc_mp_coeff(int nBoxMols, const int *boxMols, const int *inForceRange,
                    const double *ofx, const double *ofy, const double *ofz,
                    const double *orx, const double *ory, const double *orz,
                    const double *nfx, const double *nfy, const double *nfz,
                    const double *nrx, const double *nry, const double *nrz,
                    const double *kX, const double *kY, const double *kZ, double max,
                    double lambda, double beta) {
  double lBeta = lambda * beta;
  double priv = 1.0;
  for (int mi = 0; mi < nBoxMols; ++mi) {
    int m = boxMols[mi];
    if (!inForceRange[m]) continue;
    double o[3] = {ofx[m], ofy[m], ofz[m]}, n[3] = {nfx[m], nfy[m], nfz[m]};
    if (orx) {
      o[0] += orx[m]; o[1] += ory[m]; o[2] += orz[m];
      n[0] += nrx[m]; n[1] += nry[m]; n[2] += nrz[m];
    }
    double k[3] = {kX[m], kY[m], kZ[m]};
    double w = 1.0;
    for (int d = 0; d < 3; ++d) {
      double lbn = n[d] * lBeta, lbo = o[d] * lBeta;
      w *= lbn * exp(-lbn * k[d]) / (2.0 * sinh(lbn * max));
      w /= lbo * exp(lbo * k[d]) / (2.0 * sinh(lbo * max));
    }
    priv *= w;
  }
  double w_ratio = 1.0;
  w_ratio *= priv;
  if (!isfinite(w_ratio)) w_ratio = 0.0;
  return w_ratio;
}

/* ------------------------------------------------------------------------ */
/* MoleculeInter / ParticleInter                                             */

/* energy of one probe position against the 27 cells around it; sign = -1 for
 * the "subtract old" sweep.  Mirrors one EnumerateLocal sweep of
 * src/CalculateEnergy.cpp:593-678 (sum order: neighbour-cell order, ascending
 * atom index inside a cell -- the reference walks its linked list instead,
 * which is the same set in a history-dependent order). */
static void probe_sweep(const orc_params *p, const cell_csr *c,
                        const double cellSize[3], const int edge[3],
                        const double *x, const double *y, const double *z,
                        const int *kind, const double *charge, double px,
                        double py, double pz, int kindI, double qI,
                        double sign, int checkOverlap, double *lj,
                        double *real, int *overlap, const int *mol, int molI) {
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double rCutLowSq = p->rCutLow * p->rCutLow;
  int cell = position_to_cell(cellSize, edge, px, py, pz);
  for (int nc = 0; nc < 27; ++nc) {
    int nbrCell = c->nbr[cell * 27 + nc];
    for (int ni = c->cellStart[nbrCell]; ni < c->cellStart[nbrCell + 1];
         ++ni) {
      int nb = c->cellVector[ni];
      double distSq, d[3];
      if (in_rcut(p, boxRcutSq, px, py, pz, x[nb], y[nb], z[nb], &distSq, d)) {
        if (checkOverlap && distSq < rCutLowSq) *overlap |= 1;
        if (p->electrostatic) {
          double qi_qj_fact = qI * charge[nb] * ORC_QQFACT;
          if (qi_qj_fact != 0.0)
            *real += sign * calc_coulomb_l(p, distSq, kindI, kind[nb], qi_qj_fact,
                                           pair_lambda_coulomb(molI, mol[nb]));
        }
        *lj += sign * calc_en_l(p, distSq, kindI, kind[nb], pair_lambda_vdw(molI, mol[nb]));
      }
    }
  }
}

int orc_molecule_inter(const orc_params *p, int nAtomsTotal, const double *x,
                       const double *y, const double *z, const int *kind,
                       const int *mol, const double *charge,
                       const int *boxAtoms, int nBox, int molIndex,
                       int molStart, int molLen, const double *newX,
                       const double *newY, const double *newZ, double *dLJ,
                       double *dReal) {
  cell_csr c = csr_make(p, nAtomsTotal, x, y, z, boxAtoms, nBox);
  int edge[3];
  orc_cell_edges(p, edge);
  double cellSize[3] = {p->axis[0] / edge[0], p->axis[1] / edge[1],
                        p->axis[2] / edge[2]};
  double lj = 0.0, real = 0.0;
  int overlap = 0;
  for (int a = 0; a < molLen; ++a) {
    int atom = molStart + a;
    /* subtract old energy, src/CalculateEnergy.cpp:593-634 */
    probe_sweep(p, &c, cellSize, edge, x, y, z, kind, charge, x[atom], y[atom],
                z[atom], kind[atom], charge[atom], -1.0, 0, &lj, &real,
                &overlap, mol, molIndex);
    /* add new energy, :637-678 */
    probe_sweep(p, &c, cellSize, edge, x, y, z, kind, charge, newX[a], newY[a],
                newZ[a], kind[atom], charge[atom], 1.0, 1, &lj, &real,
                &overlap, mol, molIndex);
  }
  *dLJ = lj;
  *dReal = real;
  csr_free(&c);
  return overlap;
}

int orc_particle_inter(const orc_params *p, int nAtomsTotal, const double *x,
                       const double *y, const double *z, const int *kind,
                       const int *mol, const double *charge,
                       const int *boxAtoms, int nBox, int molIndex, int kindI,
                       double qI, int trials, const double *tx,
                       const double *ty, const double *tz, double *en,
                       double *real, int *overlap) {
  cell_csr c = csr_make(p, nAtomsTotal, x, y, z, boxAtoms, nBox);
  int edge[3];
  orc_cell_edges(p, edge);
  double cellSize[3] = {p->axis[0] / edge[0], p->axis[1] / edge[1],
                        p->axis[2] / edge[2]};
  for (int t = 0; t < trials; ++t) { /* src/CalculateEnergy.cpp:741-782 */
    double lj = 0.0, re = 0.0;
    int ov = 0;
    probe_sweep(p, &c, cellSize, edge, x, y, z, kind, charge, tx[t], ty[t],
                tz[t], kindI, qI, 1.0, 1, &lj, &re, &ov, mol, molIndex);
    en[t] += lj;
    real[t] += re;
    overlap[t] |= ov;
  }
  csr_free(&c);
  return 0;
}

/* CalculateEnergy::ParticleNonbonded, src/CalculateEnergy.cpp:689-725: the partners are the
 * entries of kind.sortedNB(partIndex) whose site already exists in the trial molecule, in
 * that order (the caller filters); Coulomb part = FFParticle::CalcCoulombAdd_1_4 with
 * NB = true (src/FFParticle.cpp:281-292; the same text in every FF class). */
int orc_particle_nonbonded(const orc_params *p, int kindI, double qI, int nPartners,
                           const int *partnerKind, const double *partnerCharge,
                           const double *px, const double *py, const double *pz, int trials,
                           const double *tx, const double *ty, const double *tz,
                           double *inter) {
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double rCutSq = p->rCut * p->rCut;
  for (int k = 0; k < nPartners; ++k) {
    for (int t = 0; t < trials; ++t) {
      double distSq, d[3];
      if (in_rcut(p, boxRcutSq, tx[t], ty[t], tz[t], px[k], py[k], pz[k], &distSq, d)) {
        inter[t] += orc_calc_en(p, distSq, kindI, partnerKind[k]);
        if (p->electrostatic) {
          double qi_qj_fact = qI * partnerCharge[k] * ORC_QQFACT;
          if (qi_qj_fact != 0.0 && !(rCutSq < distSq)) inter[t] += qi_qj_fact / sqrt(distSq);
        }
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Torque                                                                    */

int orc_calculate_torque(const orc_params *p, int nBoxMols, const int *boxMols,
                         const int *molStart, const double *x, const double *y,
                         const double *z, const double *comX,
                         const double *comY, const double *comZ,
                         const double *aFx, const double *aFy,
                         const double *aFz, const double *rFx,
                         const double *rFy, const double *rFz, double *tx,
                         double *ty, double *tz) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int mi = 0; mi < nBoxMols; ++mi) { /* CalculateEnergy.cpp:1384-1403 */
    int m = boxMols[mi];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int a = molStart[m]; a < molStart[m + 1]; ++a) {
      double dd[3] = {x[a] - comX[m], y[a] - comY[m], z[a] - comZ[m]};
      min_image_vec(p, dd);
      double dx = dd[0], dy = dd[1], dz = dd[2];
      double fx = aFx[a] + rFx[a], fy = aFy[a] + rFy[a], fz = aFz[a] + rFz[a];
      /* geom::Cross, lib/GeomLib.h */
      sx += dy * fz - dz * fy;
      sy += dz * fx - dx * fz;
      sz += dx * fy - dy * fx;
    }
    tx[m] = sx;
    ty[m] = sy;
    tz[m] = sz;
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* LRC                                                                       */

static double energy_lrc_pair(const orc_params *p, int kind1, int kind2) {
  /* FFParticle::EnergyLRC, src/FFParticle.cpp:96-116 */
  int idx = kind1 + kind2 * p->kindCount;
  double tc = 1.0;
  double sigma = sqrt(p->sigmaSq[idx]);
  double rRat = sigma / p->rCut;
  double N = p->n[idx];
  tc *= sigma * p->sigmaSq[idx];
  tc *= 2.0 * M_PI * p->epsilon_cn[idx] / (N - 3.0);
  tc *= (pow(rRat, p->n[idx] - 3.0) - ((N - 3.0) / 3.0) * rRat * rRat * rRat);
  return tc;
}

double orc_energy_lrc(const orc_params *p, int nMolKinds,
                      const int *molKindStart, const int *molKindAtomKinds,
                      const int *numKindInBox) {
  /* SHIFT/SWITCH have zero LRC (FFShift.h / FFSwitch.h EnergyLRC return 0) */
  if (p->vdwKind != ORC_VDW_STD) return 0.0;
  double volume = p->axis[0] * p->axis[1] * p->axis[2];
  if (p->nonOrth) { /* |A . (B x C)|, BoxDimensionsNonOrth::Init */
    double a[3], b[3], c[3];
    for (int k = 0; k < 3; ++k) {
      a[k] = p->cellBasis[k] * p->axis[0];
      b[k] = p->cellBasis[3 + k] * p->axis[1];
      c[k] = p->cellBasis[6 + k] * p->axis[2];
    }
    volume = fabs(a[0] * (b[1] * c[2] - b[2] * c[1]) + a[1] * (b[2] * c[0] - b[0] * c[2]) +
                  a[2] * (b[0] * c[1] - b[1] * c[0]));
  }
  double volInv = 1.0 / volume;
  double en = 0.0;
  /* pairEnCorrections, src/Molecules.cpp (sum over atom pairs of the two
   * molecule kinds), then CalculateEnergy::EnergyCorrection :1261-1269 */
  for (int i = 0; i < nMolKinds; ++i)
    for (int j = 0; j < nMolKinds; ++j) {
      double pairCorr = 0.0;
      for (int pI = molKindStart[i]; pI < molKindStart[i + 1]; ++pI)
        for (int pJ = molKindStart[j]; pJ < molKindStart[j + 1]; ++pJ)
          pairCorr +=
              energy_lrc_pair(p, molKindAtomKinds[pI], molKindAtomKinds[pJ]);
      en += pairCorr * numKindInBox[i] * numKindInBox[j] * volInv;
    }
  return en;
}

/* ------------------------------------------------------------------------ */
/* Reciprocal space                                                          */

int orc_recip_init_orth(const orc_params *p, double *kx, double *ky,
                        double *kz, double *hsqr, double *prefact,
                        int *kmaxOut) {
  /* src/Ewald.cpp:847-903 */
  int counter = 0;
  double alpsqr4 = 1.0 / (4.0 * (p->alpha * p->alpha));
  double cv[3] = {2.0 * M_PI * (1.0 / p->axis[0]),
                  2.0 * M_PI * (1.0 / p->axis[1]),
                  2.0 * M_PI * (1.0 / p->axis[2])};
  /* XYZ::Inverse then *= 2pi (lib/BasicTypes.h): x = 1.0/x; x *= 2pi */
  cv[0] = (1.0 / p->axis[0]) * (2.0 * M_PI);
  cv[1] = (1.0 / p->axis[1]) * (2.0 * M_PI);
  cv[2] = (1.0 / p->axis[2]) * (2.0 * M_PI);
  double volume = p->volume > 0.0 ? p->volume : p->axis[0] * p->axis[1] * p->axis[2];
  double vol = volume / (4.0 * M_PI);
  double rr2 = p->recip_rcut * p->recip_rcut;
  int nkx_max = (int)(p->recip_rcut * p->axis[0] / (2.0 * M_PI)) + 1;
  int nky_max = (int)(p->recip_rcut * p->axis[1] / (2.0 * M_PI)) + 1;
  int nkz_max = (int)(p->recip_rcut * p->axis[2] / (2.0 * M_PI)) + 1;
  if (kmaxOut) {
    int m = nkx_max > nky_max ? nkx_max : nky_max;
    *kmaxOut = m > nkz_max ? m : nkz_max;
  }
  for (int ix = 0; ix <= nkx_max; ix++) {
    int nky_min = (ix == 0) ? 0 : -nky_max;
    for (int iy = nky_min; iy <= nky_max; iy++) {
      int nkz_min = (ix == 0 && iy == 0) ? 1 : -nkz_max;
      for (int iz = nkz_min; iz <= nkz_max; iz++) {
        double kX = cv[0] * ix;
        double kY = cv[1] * iy;
        double kZ = cv[2] * iz;
        double ksqr = kX * kX + kY * kY + kZ * kZ;
        if (ksqr < rr2) {
          if (kx) {
            kx[counter] = kX;
            ky[counter] = kY;
            kz[counter] = kZ;
            hsqr[counter] = ksqr;
            prefact[counter] =
                ORC_QQFACT * exp(-ksqr * alpsqr4) / (ksqr * vol);
          }
          counter++;
        }
      }
    }
  }
  return counter;
}

int orc_recip_init_nonorth(const orc_params *p, double *kx, double *ky,
                           double *kz, double *hsqr, double *prefact,
                           int *kmaxOut) {
  /* src/Ewald.cpp:905-965: cellB = normalised basis scaled by the edge lengths,
   * reciprocal rows = adjoint * 2 pi / det */
  int counter = 0;
  double alpsqr4 = 1.0 / (4.0 * (p->alpha * p->alpha));
  double cb[9], inv[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cb[3 * r + c] = p->cellBasis[3 * r + c] * p->axis[r];
  /* XYZArray::AdjointMatrix, src/XYZArray.h:507-522 (x[i] = cb[3i], ...) */
#define X(i) cb[3 * (i)]
#define Y(i) cb[3 * (i) + 1]
#define Z(i) cb[3 * (i) + 2]
  inv[0] = Y(1) * Z(2) - Y(2) * Z(1);
  inv[1] = Y(2) * Z(0) - Y(0) * Z(2);
  inv[2] = Y(0) * Z(1) - Y(1) * Z(0);
  inv[3] = X(2) * Z(1) - X(1) * Z(2);
  inv[4] = X(0) * Z(2) - X(2) * Z(0);
  inv[5] = X(1) * Z(0) - X(0) * Z(1);
  inv[6] = X(1) * Y(2) - X(2) * Y(1);
  inv[7] = X(2) * Y(0) - X(0) * Y(2);
  inv[8] = X(0) * Y(1) - X(1) * Y(0);
  double det = X(0) * inv[0] + X(1) * inv[1] + X(2) * inv[2];
  /* volume = |A . (B x C)| (BoxDimensionsNonOrth::Init) */
  double bxc[3] = {Y(1) * Z(2) - Z(1) * Y(2), Z(1) * X(2) - X(1) * Z(2),
                   X(1) * Y(2) - Y(1) * X(2)};
  double volume = fabs(X(0) * bxc[0] + Y(0) * bxc[1] + Z(0) * bxc[2]);
#undef X
#undef Y
#undef Z
  for (int i = 0; i < 9; ++i) inv[i] *= (2.0 * M_PI) / det;
  double vol = volume / (4.0 * M_PI);
  double rr2 = p->recip_rcut * p->recip_rcut;
  int nkx_max = (int)(p->recip_rcut * p->axis[0] / (2.0 * M_PI)) + 1;
  int nky_max = (int)(p->recip_rcut * p->axis[1] / (2.0 * M_PI)) + 1;
  int nkz_max = (int)(p->recip_rcut * p->axis[2] / (2.0 * M_PI)) + 1;
  if (kmaxOut) {
    int m = nkx_max > nky_max ? nkx_max : nky_max;
    *kmaxOut = m > nkz_max ? m : nkz_max;
  }
  for (int ix = 0; ix <= nkx_max; ix++) {
    int nky_min = (ix == 0) ? 0 : -nky_max;
    for (int iy = nky_min; iy <= nky_max; iy++) {
      int nkz_min = (ix == 0 && iy == 0) ? 1 : -nkz_max;
      for (int iz = nkz_min; iz <= nkz_max; iz++) {
        /* Dot(cellB_Inv.Get(r), XYZ(x,y,z)) */
        double kX = inv[0] * ix + inv[1] * iy + inv[2] * iz;
        double kY = inv[3] * ix + inv[4] * iy + inv[5] * iz;
        double kZ = inv[6] * ix + inv[7] * iy + inv[8] * iz;
        double ksqr = kX * kX + kY * kY + kZ * kZ;
        if (ksqr < rr2) {
          if (kx) {
            kx[counter] = kX;
            ky[counter] = kY;
            kz[counter] = kZ;
            hsqr[counter] = ksqr;
            prefact[counter] = ORC_QQFACT * exp(-ksqr * alpsqr4) / (ksqr * vol);
          }
          counter++;
        }
      }
    }
  }
  return counter;
}

int orc_box_recip_sums_slab(int nBoxMols, const int *boxMols,
                            const int *molStart, const double *x,
                            const double *y, const double *z,
                            const double *charge, int k0, int k1,
                            const double *kx, const double *ky,
                            const double *kz, double *sumR, double *sumI) {
  /* src/Ewald.cpp:222-269: memset, then molecule-outer / k-inner */
  for (int i = k0; i < k1; ++i) sumR[i] = sumI[i] = 0.0;
  /* k is the parallel axis in the reference too (omp for over i inside the
   * molecule loop); hoisting the parallel region keeps each sumR[i]'s
   * accumulation order (molecule order) unchanged. */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int i = k0; i < k1; ++i) {
    double accR = 0.0, accI = 0.0;
    for (int mi = 0; mi < nBoxMols; ++mi) {
      int m = boxMols[mi];
      double sumReal = 0.0, sumImaginary = 0.0;
      for (int a = molStart[m]; a < molStart[m + 1]; ++a) {
        if (fabs(charge[a]) < 0.000000001) continue; /* Ewald.cpp:107-111 */
        /* geom::Dot, lib/GeomLib.h:73-76 */
        double dot = x[a] * kx[i] + y[a] * ky[i] + z[a] * kz[i];
        sumReal += charge[a] * cos(dot);
        sumImaginary += charge[a] * sin(dot);
      }
      double lambdaCoef = mol_lambda_coef(m); /* src/Ewald.cpp:240,266-267 */
      accR += (lambdaCoef * sumReal);
      accI += (lambdaCoef * sumImaginary);
    }
    sumR[i] = accR;
    sumI[i] = accI;
  }
  return 0;
}

int orc_box_recip_sums(int nBoxMols, const int *boxMols, const int *molStart,
                       const double *x, const double *y, const double *z,
                       const double *charge, int nk, const double *kx,
                       const double *ky, const double *kz, double *sumR,
                       double *sumI) {
  return orc_box_recip_sums_slab(nBoxMols, boxMols, molStart, x, y, z, charge,
                                 0, nk, kx, ky, kz, sumR, sumI);
}

double orc_box_reciprocal(int nk, const double *sumR, const double *sumI,
                          const double *prefact) {
  double e = 0.0; /* src/Ewald.cpp:396-400 */
  for (int i = 0; i < nk; ++i)
    e += (sumR[i] * sumR[i] + sumI[i] * sumI[i]) * prefact[i];
  return e;
}

double orc_mol_reciprocal(int molLen, const double *q, const double *oldX,
                          const double *oldY, const double *oldZ,
                          const double *newX, const double *newY,
                          const double *newZ, int nk, const double *kx,
                          const double *ky, const double *kz,
                          const double *prefact, const double *sumRref,
                          const double *sumIref, double *sumRnew,
                          double *sumInew) {
  return orc_mol_reciprocal_l(molLen, q, oldX, oldY, oldZ, newX, newY, newZ, nk, kx, ky, kz,
                              prefact, sumRref, sumIref, sumRnew, sumInew, 1.0);
}

double orc_mol_reciprocal_l(int molLen, const double *q, const double *oldX,
                            const double *oldY, const double *oldZ,
                            const double *newX, const double *newY,
                            const double *newZ, int nk, const double *kx,
                            const double *ky, const double *kz,
                            const double *prefact, const double *sumRref,
                            const double *sumIref, double *sumRnew,
                            double *sumInew, double lambdaCoef) {
  double eNew = 0.0; /* src/Ewald.cpp:434-466 */
  for (int i = 0; i < nk; ++i) {
    double sRn = 0.0, sIn = 0.0, sRo = 0.0, sIo = 0.0;
    for (int a = 0; a < molLen; ++a) {
      if (fabs(q[a]) < 0.000000001) continue;
      double dotNew = newX[a] * kx[i] + newY[a] * ky[i] + newZ[a] * kz[i];
      double dotOld = oldX[a] * kx[i] + oldY[a] * ky[i] + oldZ[a] * kz[i];
      sRn += q[a] * cos(dotNew);
      sIn += q[a] * sin(dotNew);
      sRo += q[a] * cos(dotOld);
      sIo += q[a] * sin(dotOld);
    }
    sumRnew[i] = sumRref[i] + lambdaCoef * (sRn - sRo);
    sumInew[i] = sumIref[i] + lambdaCoef * (sIn - sIo);
    eNew += (sumRnew[i] * sumRnew[i] + sumInew[i] * sumInew[i]) * prefact[i];
  }
  return eNew;
}

double orc_recip_weighted(int n, const double *w, const double *x,
                          const double *y, const double *z, int nk,
                          const double *kx, const double *ky, const double *kz,
                          const double *prefact, const double *baseR,
                          const double *baseI, double scale, double *sumRnew,
                          double *sumInew) {
  double eNew = 0.0; /* src/Ewald.cpp:744-807 (scale 1) / :557-580 */
  for (int i = 0; i < nk; ++i) {
    double sR = 0.0, sI = 0.0;
    for (int a = 0; a < n; ++a) {
      double dot = x[a] * kx[i] + y[a] * ky[i] + z[a] * kz[i];
      sR += w[a] * cos(dot);
      sI += w[a] * sin(dot);
    }
    sumRnew[i] = baseR[i] + scale * sR;
    sumInew[i] = baseI[i] + scale * sI;
    eNew += (sumRnew[i] * sumRnew[i] + sumInew[i] * sumInew[i]) * prefact[i];
  }
  return eNew;
}

void orc_change_recip(int molLen, const double *q, const double *mx,
                      const double *my, const double *mz, int nk,
                      const double *kx, const double *ky, const double *kz,
                      const double *prefact, const double *sumRref,
                      const double *sumIref, int nStates,
                      const double *lambdaCoul, int iState,
                      double *energyRecip) {
  for (int s = 0; s < nStates; ++s) energyRecip[s] = 0.0;
  for (int i = 0; i < nk; ++i) { /* src/Ewald.cpp:603-630 */
    double sR = 0.0, sI = 0.0;
    for (int a = 0; a < molLen; ++a) {
      if (fabs(q[a]) < 0.000000001) continue;
      double dot = mx[a] * kx[i] + my[a] * ky[i] + mz[a] * kz[i];
      sR += q[a] * cos(dot);
      sI += q[a] * sin(dot);
    }
    for (int s = 0; s < nStates; ++s) {
      double coefDiff = sqrt(lambdaCoul[s]) - sqrt(lambdaCoul[iState]);
      energyRecip[s] += prefact[i] * ((sumRref[i] + coefDiff * sR) *
                                          (sumRref[i] + coefDiff * sR) +
                                      (sumIref[i] + coefDiff * sI) *
                                          (sumIref[i] + coefDiff * sI));
    }
  }
}

double orc_swap_recip(int insert, int molLen, const double *q,
                      const double *mx, const double *my, const double *mz,
                      int nk, const double *kx, const double *ky,
                      const double *kz, const double *prefact,
                      const double *sumRref, const double *sumIref,
                      double *sumRnew, double *sumInew) {
  double eNew = 0.0; /* src/Ewald.cpp:501-524 / :681-703 */
  for (int i = 0; i < nk; ++i) {
    double sR = 0.0, sI = 0.0;
    for (int a = 0; a < molLen; ++a) {
      if (fabs(q[a]) < 0.000000001) continue;
      double dot = mx[a] * kx[i] + my[a] * ky[i] + mz[a] * kz[i];
      sR += q[a] * cos(dot);
      sI += q[a] * sin(dot);
    }
    if (insert) {
      sumRnew[i] = sumRref[i] + sR;
      sumInew[i] = sumIref[i] + sI;
    } else {
      sumRnew[i] = sumRref[i] - sR;
      sumInew[i] = sumIref[i] - sI;
    }
    eNew += (sumRnew[i] * sumRnew[i] + sumInew[i] * sumInew[i]) * prefact[i];
  }
  return eNew;
}

int orc_box_force_reciprocal(const orc_params *p, int nBoxMols,
                             const int *boxMols, const int *molStart,
                             const double *x, const double *y, const double *z,
                             const double *charge, int nk, const double *kx,
                             const double *ky, const double *kz,
                             const double *prefact, const double *sumR,
                             const double *sumI, double *rFx, double *rFy,
                             double *rFz, double *mFx, double *mFy,
                             double *mFz) {
  double constValue = p->alpha * M_2_SQRTPI; /* src/Ewald.cpp:1502 */
  double alphaSq = p->alpha * p->alpha;
  double boxRcutSq = box_rcut(p) * box_rcut(p);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int mi = 0; mi < nBoxMols; ++mi) { /* :1541-1592 */
    int m = boxMols[mi];
    double lambdaCoef = mol_lambda_coef(m);
    double msx = 0.0, msy = 0.0, msz = 0.0;
    for (int a = molStart[m]; a < molStart[m + 1]; ++a) {
      double X = 0.0, Y = 0.0, Z = 0.0;
      if (!(fabs(charge[a]) < 0.000000001)) {
        for (int j = molStart[m]; j < molStart[m + 1]; ++j) {
          if (a != j) { /* intra (correction) force, :1556-1569 */
            double distSq, d[3];
            in_rcut(p, boxRcutSq, x[a], y[a], z[a], x[j], y[j], z[j], &distSq,
                    d);
            double dist = sqrt(distSq);
            double expConstValue = exp(-1.0 * alphaSq * distSq);
            double qiqj = charge[a] * charge[j] * ORC_QQFACT;
            double intraForce = qiqj * lambdaCoef * lambdaCoef / distSq;
            intraForce *=
                ((erf(p->alpha * dist) / dist) - constValue * expConstValue);
            X -= intraForce * d[0];
            Y -= intraForce * d[1];
            Z -= intraForce * d[2];
          }
        }
        for (int i = 0; i < nk; ++i) { /* :1575-1586 */
          double dot = x[a] * kx[i] + y[a] * ky[i] + z[a] * kz[i];
          double factor = 2.0 * charge[a] * prefact[i] * lambdaCoef *
                          (sin(dot) * sumR[i] - cos(dot) * sumI[i]);
          X += factor * kx[i];
          Y += factor * ky[i];
          Z += factor * kz[i];
        }
      }
      rFx[a] = X;
      rFy[a] = Y;
      rFz[a] = Z;
      msx += X;
      msy += Y;
      msz += Z;
    }
    /* molForceRec.Set(0) then Add per atom (:1547, :1589) */
    mFx[m] = msx;
    mFy[m] = msy;
    mFz[m] = msz;
  }
  return 0;
}

static double mol_correction(const orc_params *p, int len, const double *q,
                             const double *mx, const double *my,
                             const double *mz, int skipUncharged) {
  double boxRcutSq = box_rcut(p) * box_rcut(p);
  double correction = 0.0;
  for (int i = 0; i < len; ++i) {
    if (skipUncharged && fabs(q[i]) < 0.000000001) continue;
    for (int j = i + 1; j < len; ++j) {
      double distSq, d[3];
      in_rcut(p, boxRcutSq, mx[i], my[i], mz[i], mx[j], my[j], mz[j], &distSq,
              d);
      double dist = sqrt(distSq);
      correction += (q[i] * q[j] * erf(p->alpha * dist) / dist);
    }
  }
  return correction;
}

double orc_box_correction(const orc_params *p, int nBoxMols,
                          const int *boxMols, const int *molStart,
                          const double *x, const double *y, const double *z,
                          const double *charge) {
  double total = 0.0; /* summed as src/CalculateEnergy.cpp:102-113 */
  for (int mi = 0; mi < nBoxMols; ++mi) {
    int m = boxMols[mi];
    int s = molStart[m], len = molStart[m + 1] - s;
    /* Ewald::MolCorrection, src/Ewald.cpp:1056-1085 */
    double c = mol_correction(p, len, charge + s, x + s, y + s, z + s, 1);
    double lambdaCoef = mol_lambda_coef(m);
    total += -1.0 * ORC_QQFACT * c * lambdaCoef * lambdaCoef;
  }
  return total;
}

double orc_box_self(const orc_params *p, int nBoxMols, const int *boxMols,
                    const int *molStart, const double *charge) {
  /* src/Ewald.cpp:1125-1163 groups by molecule kind (molSelfEnergy * molNum);
   * here every molecule is summed individually, which is the same value up
   * to rounding. */
  double self = 0.0;
  for (int mi = 0; mi < nBoxMols; ++mi) {
    int m = boxMols[mi];
    double molSelf = 0.0;
    for (int a = molStart[m]; a < molStart[m + 1]; ++a)
      molSelf += charge[a] * charge[a];
    /* src/Ewald.cpp:1140-1155 takes the fractional molecule out of its kind's count and
     * adds it back times lambdaRef.GetLambdaCoulomb(i, box) -- called with the KIND index
     * i where a molecule index is expected, so the factor is lambda only when the
     * fractional molecule's index equals its kind index, and 1 otherwise.  As written. */
    self += molSelf * ((m == g_lambda.mol && g_lambda.mol == g_lambda.molKind)
                           ? g_lambda.coulomb
                           : 1.0);
  }
  self *= -1.0 * p->alpha * ORC_QQFACT * M_2_SQRTPI * 0.5;
  return self;
}

double orc_swap_correction(const orc_params *p, int molLen, const double *q,
                           const double *mx, const double *my,
                           const double *mz) {
  /* src/Ewald.cpp:1311-1335: correction -= ...; return qqFact*correction */
  double c = mol_correction(p, molLen, q, mx, my, mz, 0);
  return ORC_QQFACT * (-c);
}

/* Ewald::ChangeSelf (src/Ewald.cpp:1395-1417) and ChangeCorrection (:1089-1122): lambda = 1
 * self / correction of one molecule (true charges, resident coordinates); the caller
 * scales by lambda_s - lambda_iState. */
void orc_change_self_correction(const orc_params *p, int molLen, const double *q,
                                const double *mx, const double *my, const double *mz,
                                double *enSelf, double *correction) {
  double en_self = 0.0;
  for (int i = 0; i < molLen; ++i) en_self += (q[i] * q[i]);
  en_self *= -1.0 * p->alpha * ORC_QQFACT * M_2_SQRTPI * 0.5;
  double c = mol_correction(p, molLen, q, mx, my, mz, 1);
  c *= -1.0 * ORC_QQFACT;
  *enSelf = en_self;
  *correction = c;
}

double orc_swap_self(const orc_params *p, int molLen, const double *q) {
  double en_self = 0.0; /* src/Ewald.cpp:1375-1391 */
  for (int i = 0; i < molLen; ++i) en_self -= q[i] * q[i];
  return en_self * p->alpha * ORC_QQFACT * M_2_SQRTPI * 0.5;
}
