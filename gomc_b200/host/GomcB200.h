// GomcB200.h -- C++ host mirror of GOMC's energy interfaces over the C ABI.
//
// The reference reaches its hot path through two objects held by every move
// (src/moves/MoveBase.h:83-84): `CalculateEnergy &calcEnRef` and
// `Ewald *calcEwald` (Ewald / EwaldCached / NoEwald, chosen in
// src/System.cpp:132-148).  The classes below keep those names, method names,
// argument meaning and the reference's error behaviour (print + exit, as
// gpuAssert does in src/GPU/VariablesCUDA.cuh:20-39), and forward to
// include/gomc_b200.h.  Header-only; link with libgomc_b200.so.
//
// Coordinates are passed as SoA views (what XYZArray is, src/XYZArray.h:25-70).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/gomc_b200.h"

namespace gomc_b200 {

struct XYZView {           // XYZArray: three parallel arrays
  const double *x, *y, *z;
  int count;
};
struct XYZ {
  double x, y, z;
};
struct Intermolecular {    // src/EnergyTypes.h:40-73
  double virial = 0.0, energy = 0.0;
};
struct Energy {            // the members of src/EnergyTypes.h:75-129 this path fills
  double inter = 0.0, real = 0.0, recip = 0.0, self = 0.0, correction = 0.0,
         tailCorrection = 0.0;
};

struct Virial {            // src/EnergyTypes.h:176-327 (members this path fills)
  double inter = 0.0, real = 0.0, recip = 0.0, tailCorrection = 0.0;
  double interTens[3][3] = {}, realTens[3][3] = {}, recipTens[3][3] = {};
};

inline void check(int rc, const char *what) {
  if (rc != GOMCB200_OK) {
    fprintf(stderr, "GPUassert: %s failed (%d): %s\n", what, rc, gomcb200_last_error());
    exit(EXIT_FAILURE);  // same convention as the reference's gpuAssert
  }
}

// Owns the device engine (the role of VariablesCUDA, src/FFParticle.cpp:56).
class EngineB200 {
public:
  EngineB200(int device, int boxTotal) { check(gomcb200_create(&e_, device, boxTotal), "create"); }
  ~EngineB200() { gomcb200_destroy(e_); }
  EngineB200(const EngineB200 &) = delete;
  EngineB200 &operator=(const EngineB200 &) = delete;
  gomcb200_engine *get() const { return e_; }

  // FFParticle::Init -> InitGPUForceField (src/FFParticle.cpp:87)
  void InitForceField(const double *sigmaSq, const double *epsilon_cn, const double *n,
                      int vdwKind, int isMartini, int count, double rCut,
                      const double *rCutCoulomb, double rCutLow, double rOn,
                      const double *alpha, bool ewald, bool electrostatic,
                      double diElectric_1 = 1.0) {
    check(gomcb200_init_forcefield(e_, sigmaSq, epsilon_cn, n, vdwKind, isMartini, count, rCut,
                                   rCutCoulomb, rCutLow, rOn, alpha, ewald, electrostatic,
                                   diElectric_1),
          "InitGPUForceField");
  }
  // FF_EXP6::Init -> InitExp6VariablesCUDA (src/FFExp6.h:145)
  void InitExp6(const double *rMin, const double *expConst, const double *rMaxSq, int size) {
    check(gomcb200_init_exp6(e_, rMin, expConst, rMaxSq, size), "InitExp6VariablesCUDA");
  }
  // forcefield.sc_alpha / sc_sigma_6 / sc_power / sc_coul (src/Forcefield.cpp:58-75)
  void InitSoftcore(double sc_alpha, double sc_sigma_6, int sc_power, bool sc_coul) {
    check(gomcb200_init_softcore(e_, sc_alpha, sc_sigma_6, sc_power, sc_coul), "InitSoftcore");
  }
  // Lambda::Set / UnSet -> UpdateGPULambda (lib/Lambda.h:65-88)
  void SetLambda(double vdw, double coulomb, int mol, int kind, int box) {
    check(gomcb200_update_lambda(e_, box, mol, kind, vdw, coulomb, 1), "UpdateGPULambda");
  }
  void UnSetLambda(int box) {
    check(gomcb200_update_lambda(e_, box, 0, 0, 1.0, 1.0, 0), "UpdateGPULambda");
  }
  // CalculateEnergy::Init (src/CalculateEnergy.cpp:60-81)
  void InitTopology(const std::vector<int> &particleKind, const std::vector<int> &particleMol,
                    const std::vector<double> &particleCharge, const std::vector<int> &molStart) {
    molStart_ = molStart;
    charge_ = particleCharge;
    check(gomcb200_init_topology(e_, (int)particleKind.size(), (int)molStart.size() - 1,
                                 particleKind.data(), particleMol.data(), particleCharge.data(),
                                 molStart.data()),
          "InitCoordinatesCUDA");
  }
  void SetBoxMolecules(int box, const std::vector<int> &mols) {
    check(gomcb200_set_box_molecules(e_, box, mols.data(), (int)mols.size()), "SetBoxMolecules");
  }
  void SetBoxAxes(int box, const XYZ &axis) {  // UpdateCellBasisCUDA
    double a[3] = {axis.x, axis.y, axis.z};
    check(gomcb200_set_box_axes(e_, box, a), "UpdateCellBasisCUDA");
  }
  // UpdateCellBasisCUDA + UpdateInvCellBasisCUDA for a BoxDimensionsNonOrth box
  // (src/CalculateEnergy.cpp:177-190): normalised cell vectors, their inverse, edge lengths
  void SetBoxCellBasis(int box, const double cellBasis[9], const double cellBasisInv[9],
                       const XYZ &axis) {
    double a[3] = {axis.x, axis.y, axis.z};
    check(gomcb200_set_box_cell_basis(e_, box, cellBasis, cellBasisInv, a),
          "UpdateInvCellBasisCUDA");
  }
  void SetCoordinates(const XYZView &c) {
    check(gomcb200_set_coords(e_, c.x, c.y, c.z, 0, c.count), "SetCoordinates");
  }
  void SetCOM(const XYZView &c) { check(gomcb200_set_com(e_, c.x, c.y, c.z, 0, c.count), "SetCOM"); }
  // what Translate::Accept / Rotate::Accept copy (src/moves/Translate.h:106-113)
  void AcceptMolecule(int molIndex, const XYZView &molCoords, const XYZ &com) {
    double c[3] = {com.x, com.y, com.z};
    check(gomcb200_set_molecule_coords(e_, molIndex, molCoords.x, molCoords.y, molCoords.z, c),
          "AcceptMolecule");
  }
  int MolStart(int m) const { return molStart_[m]; }
  int MolLength(int m) const { return molStart_[m + 1] - molStart_[m]; }
  double Charge(int atom) const { return charge_[atom]; }

private:
  gomcb200_engine *e_ = nullptr;
  std::vector<int> molStart_;
  std::vector<double> charge_;
};

// ---------------------------------------------------------------------------
class CalculateEnergy {
public:
  explicit CalculateEnergy(EngineB200 &eng) : eng_(eng) {}

  // src/CalculateEnergy.cpp:157-266 (pair sums; the LRC term stays on the host)
  Energy BoxInter(const XYZView &coords, const XYZ &boxAxes, int box) {
    Energy en;
    double a[3] = {boxAxes.x, boxAxes.y, boxAxes.z};
    check(gomcb200_call_box_inter(eng_.get(), box, coords.x, coords.y, coords.z, a, &en.real,
                                  &en.inter),
          "CallBoxInterGPU");
    return en;
  }
  // src/CalculateEnergy.cpp:268-406; force arrays may be null (stay on the device)
  Energy BoxForce(const XYZView &coords, double *aFx, double *aFy, double *aFz, double *mFx,
                  double *mFy, double *mFz, const XYZ &boxAxes, int box) {
    Energy en;
    double a[3] = {boxAxes.x, boxAxes.y, boxAxes.z};
    check(gomcb200_call_box_force(eng_.get(), box, coords.x, coords.y, coords.z, a, &en.real,
                                  &en.inter, aFx, aFy, aFz, mFx, mFy, mFz),
          "CallBoxForceGPU");
    return en;
  }
  // src/CalculateEnergy.cpp:411-579: pair tensors (CallBoxInterForceGPU) plus the
  // reciprocal part through calcEwald (CallVirialReciprocalGPU); resident coordinates
  // and COMs.  Like the reference only the tensor diagonals are computed; the tail
  // correction (VirialCorrection) is left to the caller's host formula.
  template <class EwaldT>
  Virial VirialCalc(EwaldT &calcEwald, int box) {
    Virial v;
    double vT[3], rT[3];
    check(gomcb200_box_inter_virial(eng_.get(), box, vT, rT), "CallBoxInterForceGPU");
    for (int c = 0; c < 3; ++c) {
      v.interTens[c][c] = vT[c];
      v.realTens[c][c] = rT[c];
    }
    v.inter = vT[0] + vT[1] + vT[2];
    v.real = rT[0] + rT[1] + rT[2];
    return calcEwald.VirialReciprocal(v, box);
  }
  // src/CalculateEnergy.cpp:581-686; returns the overlap flag
  bool MoleculeInter(Intermolecular &inter_LJ, Intermolecular &inter_coulomb,
                     const XYZView &molCoords, int molIndex, int box) const {
    int overlap = 0;
    check(gomcb200_molecule_inter(eng_.get(), box, molIndex, molCoords.x, molCoords.y,
                                  molCoords.z, &inter_LJ.energy, &inter_coulomb.energy,
                                  &overlap),
          "MoleculeInter");
    return overlap != 0;
  }
  // MoleculeInter + Ewald::MolReciprocal of one trial with a single synchronisation
  // (the pair Translate::CalcEn issues, src/moves/Translate.h:82-95)
  bool MoleculeTrial(Intermolecular &inter_LJ, Intermolecular &inter_coulomb, double &recipNew,
                     const XYZView &molCoords, int molIndex, int box) const {
    int overlap = 0;
    check(gomcb200_molecule_trial(eng_.get(), box, molIndex, molCoords.x, molCoords.y,
                                  molCoords.z, &inter_LJ.energy, &inter_coulomb.energy, &overlap,
                                  &recipNew),
          "MoleculeTrial");
    return overlap != 0;
  }
  // src/CalculateEnergy.cpp:727-785
  void ParticleInter(double *en, double *real, const XYZView &trialPos, bool *overlap,
                     int partIndex, int molIndex, int box, int trials) const {
    std::vector<int> ov(trials, 0);
    check(gomcb200_particle_inter(eng_.get(), box, molIndex, partIndex, trials, trialPos.x,
                                  trialPos.y, trialPos.z, en, real, ov.data()),
          "ParticleInter");
    for (int t = 0; t < trials; ++t) overlap[t] |= (ov[t] != 0);
  }
  // src/CalculateEnergy.cpp:689-725.  The reference walks kind.sortedNB(partIndex) and tests
  // trialMol.AtomExists(partner); here the caller hands in that filtered partner list (kinds,
  // charges and trialMol.GetCoords() of the partners) in the same order.
  void ParticleNonbonded(double *inter, int kindOfPart, double chargeOfPart,
                         const std::vector<int> &partnerKind,
                         const std::vector<double> &partnerCharge, const XYZView &partnerCoords,
                         const XYZView &trialPos, int box, int trials) const {
    check(gomcb200_particle_nonbonded(eng_.get(), box, kindOfPart, chargeOfPart,
                                      (int)partnerKind.size(), partnerKind.data(),
                                      partnerCharge.data(), partnerCoords.x, partnerCoords.y,
                                      partnerCoords.z, trials, trialPos.x, trialPos.y, trialPos.z,
                                      inter),
          "ParticleNonbonded");
  }
  // src/CalculateEnergy.cpp:1365-1406 (uses the resident force buffers and COM)
  void CalculateTorque(double *tx, double *ty, double *tz, int first, int count, int box) {
    check(gomcb200_calculate_torque(eng_.get(), box), "CalculateTorque");
    if (tx || ty || tz)
      check(gomcb200_get_forces(eng_.get(), GOMCB200_MOL_TORQUE, tx, ty, tz, first, count),
            "CalculateTorque download");
  }
  void ResetForce(int) {}  // implicit: BoxForce overwrites every atom of the box

private:
  EngineB200 &eng_;
};

// ---------------------------------------------------------------------------
// Ewald: the ~30 virtuals of src/Ewald.h:46-179 that are on the hot path.
class Ewald {
public:
  Ewald(EngineB200 &eng, const double *alpha, const double *recip_rcut, int boxTotal)
      : eng_(eng), alpha_(alpha, alpha + boxTotal), recipRcut_(recip_rcut, recip_rcut + boxTotal),
        sysPotRecip_(boxTotal, 0.0) {}
  virtual ~Ewald() {}

  // Ewald::AllocMem (src/Ewald.cpp:141-188): excess = 1.0 NVT/GCMC, 1.25 GEMC, 1.5 NPT
  virtual void AllocMem(const std::vector<XYZ> &axes, double excess) {
    check(gomcb200_init_ewald(eng_.get(), 0, recipRcut_.data()), "InitEwaldVariablesCUDA");
    int imageTotal = 0;
    for (size_t b = 0; b < axes.size(); ++b) {
      double a[3] = {axes[b].x, axes[b].y, axes[b].z};
      int n = 0;
      check(gomcb200_recip_count(eng_.get(), (int)b, a, excess, &n), "RecipCountInit");
      imageTotal = n > imageTotal ? n : imageTotal;
    }
    check(gomcb200_init_ewald(eng_.get(), imageTotal, recipRcut_.data()), "InitEwaldVariablesCUDA");
  }
  // volume = boxAxes.volume[box] of a volume trial's newDim (0: product of the axes)
  virtual void RecipInit(int box, const XYZ &axis, double volume = 0.0) {  // src/Ewald.cpp:644 -> :847
    double a[3] = {axis.x, axis.y, axis.z};
    int n = 0, kmax = 0;
    check(gomcb200_recip_init_volume(eng_.get(), box, a, volume, &n, &kmax), "RecipInit");
  }
  virtual void BoxReciprocalSetup(int box, const XYZView &molCoords) {  // :193
    eng_.SetCoordinates(molCoords);
    check(gomcb200_box_reciprocal_setup(eng_.get(), box, &currentEnergyRecip_), "BoxReciprocalSetup");
  }
  virtual void BoxReciprocalSums(int box, const XYZView &molCoords) {  // :281
    eng_.SetCoordinates(molCoords);
    check(gomcb200_box_reciprocal_sums(eng_.get(), box, &currentEnergyRecip_), "BoxReciprocalSums");
  }
  virtual double BoxReciprocal(int box, bool isNewVolume) const {  // :375
    double e = 0.0;
    check(gomcb200_box_reciprocal(eng_.get(), box, isNewVolume, &e), "BoxReciprocal");
    return e;
  }
  // returns E_new - sysPotRef.boxEnergy[box].recip, src/Ewald.cpp:409-473
  virtual double MolReciprocal(const XYZView &molCoords, int molIndex, int box) {
    double e = 0.0;
    check(gomcb200_mol_reciprocal(eng_.get(), box, molIndex, molCoords.x, molCoords.y,
                                  molCoords.z, &e),
          "CallMolReciprocalGPU");
    return e - sysPotRecip_[box];
  }
  virtual double SwapDestRecip(const XYZView &newMolCoords, int box, int molIndex) {  // :478
    double e = 0.0;
    check(gomcb200_swap_reciprocal(eng_.get(), box, molIndex, newMolCoords.x, newMolCoords.y,
                                   newMolCoords.z, 1, &e),
          "CallSwapReciprocalGPU");
    return e - sysPotRecip_[box];
  }
  virtual double SwapSourceRecip(const XYZView &oldMolCoords, int box, int molIndex) {  // :657
    double e = 0.0;
    check(gomcb200_swap_reciprocal(eng_.get(), box, molIndex, oldMolCoords.x, oldMolCoords.y,
                                   oldMolCoords.z, 0, &e),
          "CallSwapReciprocalGPU");
    return e - sysPotRecip_[box];
  }
  // src/Ewald.cpp:1311-1335 / :1340-1370 (one-block kernel; min-image on the device)
  virtual double SwapCorrection(const XYZView &molCoords, int molIndex, int box,
                                const XYZ & /*axis*/) const {
    double c = 0.0;
    check(gomcb200_swap_correction(eng_.get(), box, molIndex, molCoords.x, molCoords.y,
                                   molCoords.z, &c, nullptr),
          "SwapCorrection");
    return c;
  }
  virtual double SwapSelf(int molIndex, int box) const {  // src/Ewald.cpp:1375-1391
    const int len = eng_.MolLength(molIndex);
    std::vector<double> zero(len, 0.0);
    double s = 0.0;
    check(gomcb200_swap_correction(eng_.get(), box, molIndex, zero.data(), zero.data(),
                                   zero.data(), nullptr, &s),
          "SwapSelf");
    return s;
  }
  // Swap{Dest,Source}Recip + SwapCorrection + SwapSelf of one box with one sync
  virtual void SwapTrial(const XYZView &molCoords, int box, int molIndex, bool insert,
                         double &recipDelta, double &correction, double &self) {
    double e = 0.0;
    check(gomcb200_swap_trial(eng_.get(), box, molIndex, molCoords.x, molCoords.y, molCoords.z,
                              insert ? 1 : 0, &e, &correction, &self),
          "SwapTrial");
    recipDelta = e - sysPotRecip_[box];
  }
  // src/Ewald.cpp:714-826.  newMols/oldMols: coordinates of the inserted / removed
  // molecules (molecule indices give the charges); lambdaCoef = 1 (no fractional
  // molecule).  Returns E_new - sysPotRef recip.
  virtual double MolExchangeReciprocal(const std::vector<XYZView> &newMols,
                                       const std::vector<XYZView> &oldMols,
                                       const std::vector<int> &molIndexNew,
                                       const std::vector<int> &molIndexOld, int box,
                                       bool first_call) {
    std::vector<double> w, x, y, z;
    auto add = [&](const XYZView &c, int m, double sign) {
      const int start = eng_.MolStart(m);
      for (int p = 0; p < eng_.MolLength(m); ++p) {
        const double q = eng_.Charge(start + p);
        if (std::fabs(q) < 0.000000001) continue;  // particleHasNoCharge
        w.push_back(sign * q);
        x.push_back(c.x[p]);
        y.push_back(c.y[p]);
        z.push_back(c.z[p]);
      }
    };
    for (size_t m = 0; m < newMols.size(); ++m) add(newMols[m], molIndexNew[m], 1.0);
    for (size_t m = 0; m < oldMols.size(); ++m) add(oldMols[m], molIndexOld[m], -1.0);
    double e = 0.0;
    check(gomcb200_mol_exchange_reciprocal(eng_.get(), box, (int)w.size(), w.data(), x.data(),
                                           y.data(), z.data(), first_call ? 1 : 0, 1.0, &e),
          "CallMolExchangeReciprocalGPU");
    return e - sysPotRecip_[box];
  }
  // src/Ewald.cpp:534-585
  virtual double ChangeLambdaRecip(const XYZView &molCoords, double lambdaOld, double lambdaNew,
                                   int molIndex, int box) {
    double e = 0.0;
    check(gomcb200_change_lambda_mol_reciprocal(eng_.get(), box, molIndex, molCoords.x,
                                                molCoords.y, molCoords.z,
                                                std::sqrt(lambdaNew) - std::sqrt(lambdaOld), &e),
          "CallChangeLambdaMolReciprocalGPU");
    return e - sysPotRecip_[box];
  }
  // src/Ewald.cpp:589-642: energyDiff[s] = E_recip(lambda_s) - sysPotRef recip
  virtual void ChangeRecip(double *energyDiffRecip, double &dUdL_CoulRecip,
                           const std::vector<double> &lambda_Coul, int iState, int molIndex,
                           int box) const {
    check(gomcb200_change_recip(eng_.get(), box, molIndex, (int)lambda_Coul.size(),
                                lambda_Coul.data(), iState, energyDiffRecip),
          "ChangeRecip");
    for (size_t s = 0; s < lambda_Coul.size(); ++s) energyDiffRecip[s] -= sysPotRecip_[box];
    dUdL_CoulRecip += energyDiffRecip[lambda_Coul.size() - 1] - energyDiffRecip[0];
  }
  // src/Ewald.cpp:1395-1417 and :1089-1122: energyDiff[s] += (lambda_s - lambda_iState) * E
  virtual void ChangeSelfAndCorrection(double *energyDiffSelf, double *energyDiffCorrection,
                                       double &dUdL_self, double &dUdL_correction,
                                       const std::vector<double> &lambda_Coul, int iState,
                                       int molIndex, int box) const {
    double en_self = 0.0, correction = 0.0;
    check(gomcb200_change_self_correction(eng_.get(), box, molIndex, &en_self, &correction),
          "ChangeSelf/ChangeCorrection");
    for (size_t s = 0; s < lambda_Coul.size(); ++s) {
      const double coefDiff = lambda_Coul[s] - lambda_Coul[iState];
      energyDiffSelf[s] += coefDiff * en_self;
      energyDiffCorrection[s] += coefDiff * correction;
    }
    dUdL_self += en_self;
    dUdL_correction += correction;
  }
  virtual Virial VirialReciprocal(const Virial &virial, int box) const {  // :1168-1305
    Virial v = virial;
    double wT[3];
    check(gomcb200_virial_reciprocal(eng_.get(), box, wT), "CallVirialReciprocalGPU");
    for (int c = 0; c < 3; ++c) v.recipTens[c][c] = wT[c];
    v.recip = wT[0] + wT[1] + wT[2];
    return v;
  }
  virtual void BoxSelfAndCorrection(int box, double &self, double &correction) const {
    check(gomcb200_box_self_correction(eng_.get(), box, &self, &correction), "BoxSelf");
  }
  virtual void BoxForceReciprocal(double *rFx, double *rFy, double *rFz, double *mFx,
                                  double *mFy, double *mFz, int nAtoms, int nMols, int box) {
    check(gomcb200_box_force_reciprocal(eng_.get(), box), "CallBoxForceReciprocalGPU");
    if (rFx || rFy || rFz)
      check(gomcb200_get_forces(eng_.get(), GOMCB200_ATOM_FORCE_REC, rFx, rFy, rFz, 0, nAtoms),
            "BoxForceReciprocal download");
    if (mFx || mFy || mFz)
      check(gomcb200_get_forces(eng_.get(), GOMCB200_MOL_FORCE_REC, mFx, mFy, mFz, 0, nMols),
            "BoxForceReciprocal download");
  }
  // accept / reject state machine, src/Ewald.cpp:1021-1053, :1420-1487
  virtual void SetRecipRef(int box) { check(gomcb200_set_recip_ref(eng_.get(), box), "SetRecipRef"); }
  virtual void UpdateRecip(int box) { check(gomcb200_update_recip(eng_.get(), box), "UpdateRecip"); }
  virtual void CopyRecip(int box) { check(gomcb200_copy_recip(eng_.get(), box), "CopyRecip"); }
  virtual void UpdateRecipVec(int box) {
    check(gomcb200_update_recip_vec(eng_.get(), box), "UpdateRecipVec");
  }
  virtual void RestoreMol(int) {}      // non-cached Ewald: nothing to restore
  virtual void exgMolCache() {}
  virtual void backupMolCache() {}
  // sysPotRef.boxEnergy[box].recip, kept by the caller's SystemPotential
  void SetSysPotRecip(int box, double recip) { sysPotRecip_[box] = recip; }

protected:
  static double MinImageSigned(double raw, double ax) {  // src/BoxDimensions.h:169-175
    double half = ax * 0.5;
    if (raw > half) raw -= ax;
    else if (raw < -half) raw += ax;
    return raw;
  }
  EngineB200 &eng_;
  std::vector<double> alpha_, recipRcut_, sysPotRecip_;
  mutable double currentEnergyRecip_ = 0.0;
};

// EwaldCached (src/EwaldCached.h:11): the 4*M*imageTotal*8 B per-molecule cos/sin
// cache is a CPU optimisation.  On B200 recomputing the old position's a*nk sincos
// is cheaper than streaming the cache, so the cached class shares every kernel
// with Ewald and its cache-maintenance entry points are no-ops; results agree
// with the reference's cached path to rounding (tests/test_host_mirror_gpu.py).
class EwaldCached : public Ewald {
public:
  using Ewald::Ewald;
  void RestoreMol(int) override {}
  void exgMolCache() override {}
  void backupMolCache() override {}
};

// NoEwald (src/NoEwald.h): every reciprocal term is zero.
class NoEwald : public Ewald {
public:
  using Ewald::Ewald;
  void AllocMem(const std::vector<XYZ> &, double) override {}
  void RecipInit(int, const XYZ &, double = 0.0) override {}
  void BoxReciprocalSetup(int, const XYZView &) override {}
  void BoxReciprocalSums(int, const XYZView &) override {}
  double BoxReciprocal(int, bool) const override { return 0.0; }
  double MolReciprocal(const XYZView &, int, int) override { return 0.0; }
  double SwapDestRecip(const XYZView &, int, int) override { return 0.0; }
  double SwapSourceRecip(const XYZView &, int, int) override { return 0.0; }
  double SwapCorrection(const XYZView &, int, int, const XYZ &) const override { return 0.0; }
  double MolExchangeReciprocal(const std::vector<XYZView> &, const std::vector<XYZView> &,
                               const std::vector<int> &, const std::vector<int> &, int,
                               bool) override { return 0.0; }
  double ChangeLambdaRecip(const XYZView &, double, double, int, int) override { return 0.0; }
  void ChangeRecip(double *d, double &, const std::vector<double> &l, int, int,
                   int) const override { for (size_t s = 0; s < l.size(); ++s) d[s] = 0.0; }
  double SwapSelf(int, int) const override { return 0.0; }
  Virial VirialReciprocal(const Virial &virial, int) const override { return virial; }
  void ChangeSelfAndCorrection(double *, double *, double &, double &,
                               const std::vector<double> &, int, int, int) const override {}
  void BoxSelfAndCorrection(int, double &self, double &correction) const override {
    self = correction = 0.0;
  }
  void BoxForceReciprocal(double *, double *, double *, double *, double *, double *, int, int,
                          int) override {}
  void SetRecipRef(int) override {}
  void UpdateRecip(int) override {}
  void CopyRecip(int) override {}
  void UpdateRecipVec(int) override {}
};

// MultiParticle (src/moves/MultiParticle.h): the device-resident move.  The move
// loop above it (box pick, move-type draw, acceptance test, move-settings update) is
// the caller's and unchanged; this class mirrors the methods that touch coordinates
// and forces.  lambda = 0.5 as in the reference constructor (:95).
class MultiParticle {
public:
  MultiParticle(EngineB200 &eng, CalculateEnergy &calcEn, Ewald &calcEwald, double BETA)
      : eng_(eng), calcEn_(calcEn), calcEwald_(calcEwald), BETA_(BETA) {}
  enum { MPDISPLACE = 0, MPROTATE = 1 };

  // Prep (:164-247): reference forces / torques of the current positions
  void Prep(int box, int moveType) {
    bPick_ = box;
    moveType_ = moveType;
    calcEwald_.CopyRecip(box);
    calcEwald_.BoxForceReciprocal(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, box);
    check(gomcb200_box_force(eng_.get(), box, &refInter_, &refReal_), "BoxForce");
    calcEn_.CalculateTorque(nullptr, nullptr, nullptr, 0, 0, box);
  }
  // Transform (:329-411) == CallTranslateParticlesGPU / CallRotateParticlesGPU
  void Transform(double t_max, double r_max, unsigned long long step, unsigned int key,
                 unsigned long long seed, const signed char *isMoleculeInvolved = nullptr) {
    max_ = moveType_ == MPROTATE ? r_max : t_max;
    check(gomcb200_mp_transform(eng_.get(), bPick_, moveType_, max_, lambda_ * BETA_, step, key,
                                seed, isMoleculeInvolved),
          moveType_ == MPROTATE ? "CallRotateParticlesGPU" : "CallTranslateParticlesGPU");
  }
  // CalcEn (:414-441): energies and forces of the trial positions
  Energy CalcEn() {
    check(gomcb200_mp_select(eng_.get(), 1), "mp_select");
    Energy en;
    check(gomcb200_box_reciprocal_sums(eng_.get(), bPick_, &en.recip), "BoxReciprocalSums");
    check(gomcb200_box_force(eng_.get(), bPick_, &en.inter, &en.real), "BoxForce");
    calcEwald_.BoxForceReciprocal(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0,
                                  bPick_);
    calcEn_.CalculateTorque(nullptr, nullptr, nullptr, 0, 0, bPick_);
    return en;
  }
  // GetCoeff (:460-513)
  double GetCoeff() const {
    double w = 1.0;
    check(gomcb200_mp_coeff(eng_.get(), bPick_, moveType_, max_, lambda_ * BETA_, &w), "GetCoeff");
    return w;
  }
  // Accept (:515-541): result decided by the caller (pr < MPCoeff * uBoltz)
  void Accept(bool result) {
    if (result)
      check(gomcb200_mp_accept(eng_.get(), bPick_), "mp_accept");
    else
      check(gomcb200_mp_select(eng_.get(), 0), "mp_select");
  }

protected:
  EngineB200 &eng_;
  CalculateEnergy &calcEn_;
  Ewald &calcEwald_;
  double BETA_, lambda_ = 0.5, max_ = 0.0, refInter_ = 0.0, refReal_ = 0.0;
  int bPick_ = 0, moveType_ = 0;
};

// MultiParticleBrownian (src/moves/MultiParticleBrownianMotion.h): Gaussian trial
// transform, GetCoeff returns the logarithm of the weight ratio (accept with
// exp(-BETA * dU + MPCoeff), :480-484).
class MultiParticleBrownian : public MultiParticle {
public:
  using MultiParticle::MultiParticle;
  void Transform(double t_max, double r_max, unsigned long long step, unsigned int key,
                 unsigned long long seed, const signed char *isMoleculeInvolved = nullptr) {
    max_ = moveType_ == MPROTATE ? r_max : t_max;
    check(gomcb200_bm_transform(eng_.get(), bPick_, moveType_, max_, BETA_, step, key, seed,
                                isMoleculeInvolved),
          moveType_ == MPROTATE ? "BrownianMotionRotateParticlesGPU"
                                : "BrownianMotionTranslateParticlesGPU");
  }
  double GetCoeff() const {
    double w = 0.0;
    check(gomcb200_bm_coeff(eng_.get(), bPick_, moveType_, max_, BETA_, &w), "GetCoeff");
    return w;
  }
};

}  // namespace gomc_b200
