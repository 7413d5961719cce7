/*
 * gomc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the GOMC (v2.80) energy/force hot path.  It is
 * the parity checker for the CUDA engine in gomc_b200/csrc and the "port"
 * CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * path (libgomc_b200.so) never links or calls anything in oracle/.
 *
 * Pinning: the reference's own tests hold NO golden energies for this path
 * (SURVEY.md section 8c).  The restatement is pinned instead against outputs
 * of the reference itself, compiled unmodified from /root/reference by
 * oracle/ref_build.mk and driven by oracle/ref_probe.cpp; the dumps are
 * committed under tests/golden/ with the generating script
 * (oracle/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the GOMC tree).  lambda == 1 (the case of all five BASELINE configs) unless a
 * fractional molecule is declared with orc_set_lambda; four lambda fixtures pin
 * that state too.
 */
#ifndef GOMC_ORACLE_H
#define GOMC_ORACLE_H

#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* lib/NumLib.h:22 */
#define ORC_QQFACT 167103.208067979

enum { ORC_VDW_STD = 0, ORC_VDW_SHIFT = 1, ORC_VDW_SWITCH = 2, ORC_VDW_EXP6 = 3 };

/* Force-field + per-box constants.  Mirrors what Forcefield::Init
 * (src/Forcefield.cpp:29-84), FFParticle::Blend (src/FFParticle.cpp:155-199)
 * and BoxDimensions::Init (src/BoxDimensions.cpp:17-18) derive. */
typedef struct {
  int vdwKind;       /* ORC_VDW_*                       (Forcefield.cpp:94-108) */
  int ewald;         /* forcefield.ewald                                        */
  int electrostatic; /* forcefield.electrostatic                                */
  int kindCount;     /* FFParticle::count                                       */
  double rCut;       /* forcefield.rCut (LJ cut-off)                            */
  double rCutLow;    /* forcefield.rCutLow                                      */
  double rOn;        /* forcefield.rswitch (SWITCH only)                        */
  double rCutCoulomb; /* forcefield.rCutCoulomb[box]                            */
  double alpha;       /* forcefield.alpha[box]      (Forcefield.cpp:80)         */
  double recip_rcut;  /* forcefield.recip_rcut[box] (Forcefield.cpp:82)         */
  double axis[3];     /* BoxDimensions::axis (orthogonal box)                   */
  const double *sigmaSq;    /* [kindCount^2], index kind1 + kind2*count         */
  const double *epsilon_cn; /* [kindCount^2]                                    */
  const double *n;          /* [kindCount^2]                                    */
  /* ---- EXP6 (src/FFExp6.h:99-147): tables the reference derives with Brent's
   * method at init; handed in, as the reference hands them to its GPU build
   * (InitExp6VariablesCUDA).  NULL unless vdwKind == ORC_VDW_EXP6. */
  const double *rMin, *expConst, *rMaxSq;
  /* ---- Martini switch (src/FFSwitchMartini.h): vdwKind == ORC_VDW_SWITCH with
   * isMartini != 0; constants are derived from rCut, rOn, n, sigmaSq. */
  int isMartini;
  double diElectric_1;      /* 1 / forcefield.dielectric                        */
  /* ---- non-orthogonal cell (src/BoxDimensionsNonOrth.{h,cpp}): row-major 3x3,
   * rows = the NORMALISED cell basis vectors and the inverse of that matrix;
   * axis[] then holds the cell edge lengths.  nonOrth == 0: ignored. */
  int nonOrth;
  double cellBasis[9], cellBasisInv[9];
  /* BoxDimensions::volume[box] when it is not the product of the axes: after
   * BoxDimensions::SetVolume (src/BoxDimensions.cpp:214-226, every volume trial) the
   * stored volume is oldVolume + delta and the axes are cbrt-scaled, and
   * RecipInitOrth's prefactor divides by THAT volume (src/Ewald.cpp:857).  0: product. */
  double volume;
} orc_params;

/* ---- cell list (src/CellList.cpp:138-285, src/CellList.h:88-101) ---------- */
/* edge[d] = max(floor(axis[d]/cutoff),3); returns number of cells. */
int orc_cell_edges(const orc_params *p, int edge[3]);
/* CSR of the atoms listed in boxAtoms[nBox] (sorted ascending inside each
 * cell, as GetCellListNeighbor does).  cellVector[nBox], cellStart[nCells+1],
 * mapParticleToCell[nAtomsTotal] (-1 for atoms not in the box),
 * neighborList[nCells*27]. */
int orc_cell_list_build(const orc_params *p, int nAtomsTotal, const double *x,
                        const double *y, const double *z, const int *boxAtoms,
                        int nBox, int *cellVector, int *cellStart,
                        int *mapParticleToCell, int *neighborList);

/* ---- pair path ---------------------------------------------------------- */
/* CalculateEnergy::BoxInter, src/CalculateEnergy.cpp:157-266 (CPU branch). */
int orc_box_inter(const orc_params *p, int nAtomsTotal, const double *x,
                  const double *y, const double *z, const int *kind,
                  const int *mol, const double *charge, const int *boxAtoms,
                  int nBox, double *ljEn, double *realEn);

/* CalculateEnergy::BoxForce, src/CalculateEnergy.cpp:268-406.  Force arrays
 * are indexed by global atom / molecule index and are zeroed for the atoms
 * and molecules of the box first (ResetForce, :1408-1428). */
int orc_box_force(const orc_params *p, int nAtomsTotal, int nMols,
                  const double *x, const double *y, const double *z,
                  const int *kind, const int *mol, const double *charge,
                  const int *boxAtoms, int nBox, double *ljEn, double *realEn,
                  double *aFx, double *aFy, double *aFz, double *mFx,
                  double *mFy, double *mFz);

/* CalculateEnergy::MoleculeInter, src/CalculateEnergy.cpp:581-686.  The moved
 * molecule (atoms molStart..molStart+len) must NOT be in boxAtoms (the caller
 * does CellList::RemoveMol first, src/moves/Translate.h:84-89).  Returns the
 * overlap flag. */
int orc_molecule_inter(const orc_params *p, int nAtomsTotal, const double *x,
                       const double *y, const double *z, const int *kind,
                       const int *mol, const double *charge,
                       const int *boxAtoms, int nBox, int molIndex,
                       int molStart, int molLen, const double *newX,
                       const double *newY, const double *newZ, double *dLJ,
                       double *dReal);

/* CalculateEnergy::ParticleInter, src/CalculateEnergy.cpp:727-785: energies of
 * `trials` candidate positions of one atom of kind kindI / charge qI that
 * belongs to molecule molIndex (which is not in boxAtoms).  en/real are
 * incremented, overlap or-ed, as in the reference. */
int orc_particle_inter(const orc_params *p, int nAtomsTotal, const double *x,
                       const double *y, const double *z, const int *kind,
                       const int *mol, const double *charge,
                       const int *boxAtoms, int nBox, int molIndex, int kindI,
                       double qI, int trials, const double *tx,
                       const double *ty, const double *tz, double *en,
                       double *real, int *overlap);

/* CalculateEnergy::ParticleNonbonded, src/CalculateEnergy.cpp:689-725 (1-N intramolecular
 * non-bonded energy of CBMC trial positions against the already-built partner sites). */
int orc_particle_nonbonded(const orc_params *p, int kindI, double qI, int nPartners,
                           const int *partnerKind, const double *partnerCharge,
                           const double *px, const double *py, const double *pz, int trials,
                           const double *tx, const double *ty, const double *tz,
                           double *inter);

/* CalculateEnergy::CalculateTorque, src/CalculateEnergy.cpp:1365-1406. */
int orc_calculate_torque(const orc_params *p, int nBoxMols, const int *boxMols,
                         const int *molStart, const double *x, const double *y,
                         const double *z, const double *comX,
                         const double *comY, const double *comZ,
                         const double *aFx, const double *aFy,
                         const double *aFz, const double *rFx,
                         const double *rFy, const double *rFz, double *tx,
                         double *ty, double *tz);

/* FFParticle::EnergyLRC (src/FFParticle.cpp:96-116) summed as
 * CalculateEnergy::EnergyCorrection (src/CalculateEnergy.cpp:1254-1270).
 * molKindAtomKinds: CSR of atom kinds per molecule kind. */
double orc_energy_lrc(const orc_params *p, int nMolKinds,
                      const int *molKindStart, const int *molKindAtomKinds,
                      const int *numKindInBox);

/* pair functors exposed for unit tests (src/FFParticle.cpp:295-446,
 * src/FFShift.h, src/FFSwitch.h) */
double orc_calc_en(const orc_params *p, double distSq, int kind1, int kind2);
double orc_calc_vir(const orc_params *p, double distSq, int kind1, int kind2);
double orc_calc_coulomb(const orc_params *p, double distSq, double qi_qj_fact);
double orc_calc_coulomb_vir(const orc_params *p, double distSq,
                            double qi_qj_fact);

/* ---- Ewald reciprocal path ---------------------------------------------- */
/* Ewald::RecipInitOrth, src/Ewald.cpp:847-903.  Arrays sized >= the return
 * value of a first call with kx == NULL (count only).  kmaxOut optional. */
int orc_recip_init_orth(const orc_params *p, double *kx, double *ky,
                        double *kz, double *hsqr, double *prefact,
                        int *kmaxOut);
/* Ewald::RecipInitNonOrth, src/Ewald.cpp:905-965 (same calling convention). */
int orc_recip_init_nonorth(const orc_params *p, double *kx, double *ky,
                           double *kz, double *hsqr, double *prefact,
                           int *kmaxOut);

/* Ewald::BoxReciprocalSetup / BoxReciprocalSums, src/Ewald.cpp:193-361:
 * molecule-outer, k-inner, skipping |q|<1e-9 atoms (src/Ewald.cpp:107-111). */
int orc_box_recip_sums(int nBoxMols, const int *boxMols, const int *molStart,
                       const double *x, const double *y, const double *z,
                       const double *charge, int nk, const double *kx,
                       const double *ky, const double *kz, double *sumR,
                       double *sumI);
/* Same values for a k-slab [k0,k1) only (bounded CPU-baseline sample). */
int orc_box_recip_sums_slab(int nBoxMols, const int *boxMols,
                            const int *molStart, const double *x,
                            const double *y, const double *z,
                            const double *charge, int k0, int k1,
                            const double *kx, const double *ky,
                            const double *kz, double *sumR, double *sumI);

/* Ewald::BoxReciprocal, src/Ewald.cpp:375-406. */
double orc_box_reciprocal(int nk, const double *sumR, const double *sumI,
                          const double *prefact);

/* Ewald::MolReciprocal, src/Ewald.cpp:409-473 (non-cached; the cached variant
 * src/EwaldCached.cpp:245-295 returns the same value up to rounding).  Writes
 * sumRnew/sumInew, returns E_new (the caller subtracts sysPotRef recip). */
double orc_mol_reciprocal(int molLen, const double *q, const double *oldX,
                          const double *oldY, const double *oldZ,
                          const double *newX, const double *newY,
                          const double *newZ, int nk, const double *kx,
                          const double *ky, const double *kz,
                          const double *prefact, const double *sumRref,
                          const double *sumIref, double *sumRnew,
                          double *sumInew);

/* Ewald::SwapDestRecip (insert=1, src/Ewald.cpp:478-531) and SwapSourceRecip
 * (insert=0, :657-710).  Returns E_new. */
double orc_swap_recip(int insert, int molLen, const double *q,
                      const double *mx, const double *my, const double *mz,
                      int nk, const double *kx, const double *ky,
                      const double *kz, const double *prefact,
                      const double *sumRref, const double *sumIref,
                      double *sumRnew, double *sumInew);

/* Ewald::BoxForceReciprocal, src/Ewald.cpp:1496-1596 (CPU branch), including
 * the intramolecular erf-correction force (:1556-1569). */
int orc_box_force_reciprocal(const orc_params *p, int nBoxMols,
                             const int *boxMols, const int *molStart,
                             const double *x, const double *y, const double *z,
                             const double *charge, int nk, const double *kx,
                             const double *ky, const double *kz,
                             const double *prefact, const double *sumR,
                             const double *sumI, double *rFx, double *rFy,
                             double *rFz, double *mFx, double *mFy,
                             double *mFz);

/* Fractional molecule of the box (lib/Lambda.h) and soft-core constants
 * (src/Forcefield.cpp:58-75) for every function below: pair lambdas as
 * CalculateEnergy::GetLambdaVDW/GetLambdaCoulomb (src/CalculateEnergy.cpp:1558-1572),
 * reciprocal coefficient sqrt(lambdaCoulomb) as Ewald::GetLambdaCoef
 * (src/Ewald.cpp:1598-1602).  mol = -1 switches it off (the default).  Global state
 * of the test oracle, not thread-safe across callers. */
void orc_set_lambda(int mol, double lambdaVDW, double lambdaCoulomb, double sc_alpha,
                    double sc_sigma_6, int sc_power, int sc_coul, int molKind);
/* orc_mol_reciprocal with the molecule's lambdaCoef (src/Ewald.cpp:419,459-462). */
double orc_mol_reciprocal_l(int molLen, const double *q, const double *oldX,
                            const double *oldY, const double *oldZ,
                            const double *newX, const double *newY,
                            const double *newZ, int nk, const double *kx,
                            const double *ky, const double *kz,
                            const double *prefact, const double *sumRref,
                            const double *sumIref, double *sumRnew,
                            double *sumInew, double lambdaCoef);

/* Philox4x64-10 as instantiated by lib/Random123/philox.h (the generator behind
 * Random123Wrapper, src/Random123Wrapper.cpp:16-22). */
void orc_philox4x64_10(const uint64_t ctr[4], const uint64_t key[2], uint64_t out[4]);
/* MultiParticle::CalculateTrialDistRot (src/moves/MultiParticle.h:566-715):
 * moveType 0 displace (f = molForceRef, rf = molForceRecRef), 1 rotate
 * (f = molTorqueRef, rf = NULL).  nx/ny/nz and ncx/ncy/ncz hold the current
 * coordinates / COMs on entry and the trial ones on return; k = t_k or r_k. */
int orc_mp_transform(const orc_params *p, int moveType, int nBoxMols, const int *boxMols,
                     const int *molStart, const double *fx, const double *fy,
                     const double *fz, const double *rfx, const double *rfy,
                     const double *rfz, double max, double lambda, double beta,
                     uint64_t step, uint64_t seed, uint64_t keyValue, double *kX,
                     double *kY, double *kZ, int *inForceRange, double *nx, double *ny,
                     double *nz, double *ncx, double *ncy, double *ncz);
/* MultiParticleBrownian (src/moves/MultiParticleBrownianMotion.h): trial transform
 * (:CalculateTrialDistRot, Gaussian variates) and GetCoeff (log of the weight ratio). */
int orc_bm_transform(const orc_params *p, int moveType, int nBoxMols, const int *boxMols,
                     const int *molStart, const double *fx, const double *fy,
                     const double *fz, const double *rfx, const double *rfy,
                     const double *rfz, double max, double beta, uint64_t step,
                     uint64_t seed, uint64_t keyValue, double *kX, double *kY, double *kZ,
                     double *nx, double *ny, double *nz, double *ncx, double *ncy,
                     double *ncz);
double orc_bm_coeff(int nBoxMols, const int *boxMols, const double *ofx, const double *ofy,
                    const double *ofz, const double *orx, const double *ory,
                    const double *orz, const double *nfx, const double *nfy,
                    const double *nfz, const double *nrx, const double *nry,
                    const double *nrz, const double *kX, const double *kY,
                    const double *kZ, double max, double beta);
/* MultiParticle::GetCoeff, src/moves/MultiParticle.h:460-513 (o* old, n* new
 * forces or torques; *r* the reciprocal parts, NULL for rotation). */
double orc_mp_coeff(int nBoxMols, const int *boxMols, const int *inForceRange,
                    const double *ofx, const double *ofy, const double *ofz,
                    const double *orx, const double *ory, const double *orz,
                    const double *nfx, const double *nfy, const double *nfz,
                    const double *nrx, const double *nry, const double *nrz,
                    const double *kX, const double *kY, const double *kZ, double max,
                    double lambda, double beta);

/* CalculateEnergy::VirialCalc, pair part (src/CalculateEnergy.cpp:411-567):
 * diagonal of the LJ tensor vT and of the real-space Coulomb tensor rT (already
 * multiplied by qqFact).  Tail correction and reciprocal part are separate. */
int orc_virial_calc(const orc_params *p, int nAtomsTotal, const double *x,
                    const double *y, const double *z, const int *kind,
                    const int *mol, const double *charge, const double *comX,
                    const double *comY, const double *comZ,
                    const int *boxAtoms, int nBox, double vT[3], double rT[3]);
/* Ewald::VirialReciprocal, src/Ewald.cpp:1168-1305 (CPU branch): diagonal wT. */
int orc_virial_reciprocal(const orc_params *p, int nBoxMols, const int *boxMols,
                          const int *molStart, const double *x, const double *y,
                          const double *z, const double *charge,
                          const double *comX, const double *comY,
                          const double *comZ, int nk, const double *kx,
                          const double *ky, const double *kz,
                          const double *hsqr, const double *prefact,
                          const double *sumRref, const double *sumIref,
                          double wT[3]);

/* Ewald::MolExchangeReciprocal (src/Ewald.cpp:714-826: w = +q*lambdaCoef for the
 * inserted atoms in order, then -(q*lambdaCoef) for the removed ones, charged
 * atoms only, scale = 1, base = ref sums on the first call else the new sums)
 * and Ewald::ChangeLambdaRecip (:534-585: w = q, scale = sqrt(lNew)-sqrt(lOld)).
 * Returns E_new. */
double orc_recip_weighted(int n, const double *w, const double *x,
                          const double *y, const double *z, int nk,
                          const double *kx, const double *ky, const double *kz,
                          const double *prefact, const double *baseR,
                          const double *baseI, double scale, double *sumRnew,
                          double *sumInew);
/* Ewald::ChangeRecip, src/Ewald.cpp:589-642: energyRecip[s] for every lambda
 * state (caller subtracts sysPotRef recip). */
void orc_change_recip(int molLen, const double *q, const double *mx,
                      const double *my, const double *mz, int nk,
                      const double *kx, const double *ky, const double *kz,
                      const double *prefact, const double *sumRref,
                      const double *sumIref, int nStates,
                      const double *lambdaCoul, int iState,
                      double *energyRecip);

/* Ewald::MolCorrection (src/Ewald.cpp:1056-1085), summed over boxMols. */
double orc_box_correction(const orc_params *p, int nBoxMols,
                          const int *boxMols, const int *molStart,
                          const double *x, const double *y, const double *z,
                          const double *charge);
/* Ewald::BoxSelf, src/Ewald.cpp:1125-1163. */
double orc_box_self(const orc_params *p, int nBoxMols, const int *boxMols,
                    const int *molStart, const double *charge);
/* Ewald::SwapCorrection(trialMol), src/Ewald.cpp:1311-1335. */
double orc_swap_correction(const orc_params *p, int molLen, const double *q,
                           const double *mx, const double *my,
                           const double *mz);
/* Ewald::ChangeSelf (:1395-1417) / ChangeCorrection (:1089-1122): the lambda = 1 self and
 * correction energies of one molecule that the callers scale by lambda differences. */
void orc_change_self_correction(const orc_params *p, int molLen, const double *q,
                                const double *mx, const double *my, const double *mz,
                                double *enSelf, double *correction);
/* Ewald::SwapSelf, src/Ewald.cpp:1375-1391. */
double orc_swap_self(const orc_params *p, int molLen, const double *q);

int orc_max_threads(void);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
