// gomc_shim.cu -- the reference's own GPU seam, implemented on libgomc_b200.so.
//
// GOMC's GPU build reaches its kernels through the free functions declared in
// src/GPU/*.cuh.  This file DEFINES those functions -- same names, same
// C++ signatures (the reference's headers are included for the declarations) -- on top of
// the C ABI of include/gomc_b200.h, so that the UNMODIFIED reference sources, compiled with
// -DGOMC_CUDA, link against it instead of src/GPU/*.cu (integration/Makefile builds
// oracle/_ref/GOMC_B200_<ENS> that way).  Nothing of the reference is copied or patched:
// the move loop, CBMC, PRNG, ensembles and all I/O are the reference's own objects.
//
// It is the literal drop-in: like the functions it replaces it receives host arrays on
// every call and returns host arrays; the engine's resident state only caches topology and
// the k-space state machine.  (INTEGRATION.md shows the faster wiring that keeps
// coordinates resident and uses the fused single-molecule entry points.)
//
// Compiled by nvcc only because the reference headers declare __global__ kernels; there is
// no device code here.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "BoxDimensions.h"
#include "BoxDimensionsNonOrth.h"
#include "CUDAMemoryManager.cuh"
#include "CalculateEnergyCUDAKernel.cuh"
#include "CalculateEwaldCUDAKernel.cuh"
#include "CalculateForceCUDAKernel.cuh"
#include "ConstantDefinitionsCUDAKernel.cuh"
#include "TransformParticlesCUDAKernel.cuh"
#include "XYZArray.h"

#include "gomc_b200.h"

namespace {

struct Shim {
  gomcb200_engine *e = nullptr;
  // force field as handed to InitGPUForceField (kept for a re-init when `electrostatic`,
  // which only arrives with the pair calls, differs from the first guess)
  std::vector<double> sigmaSq, epsCn, n, rCutCoulomb, alpha;
  int vdwKind = 0, isMartini = 0, count = 0, ewald = 0, electrostatic = -1;
  double rCut = 0, rCutLow = 0, rOn = 0, diElectric_1 = 1.0;
  std::vector<double> rMin, expConst, rMaxSq;
  int imageTotal = 0;
  // topology (first pair call) and box membership (every pair call)
  bool haveTopo = false;
  int nAtoms = 0, nMols = 0;
  std::vector<int> molStart, particleMol;
  std::vector<int> boxMols[BOX_TOTAL];
  bool haveBox[BOX_TOTAL] = {};
  // UpdateGPULambda before the topology is known
  bool lambdaPending = false;
  int lMol[BOX_TOTAL];
  double lVdw[BOX_TOTAL], lCoul[BOX_TOTAL];
  bool lFrac[BOX_TOTAL];
  std::vector<double> tx, ty, tz;  // scratch
} g;

[[noreturn]] void die(const char *what) {
  // the reference's convention: print and exit (src/GPU/VariablesCUDA.cuh:20-39)
  fprintf(stderr, "GPUassert (gomc_b200): %s: %s\n", what, gomcb200_last_error());
  exit(EXIT_FAILURE);
}
#define CKS(call)                \
  do {                           \
    if ((call) != 0) die(#call); \
  } while (0)

void init_forcefield(int electrostatic) {
  std::vector<double> zero(BOX_TOTAL, 0.0);
  CKS(gomcb200_init_forcefield(g.e, g.sigmaSq.data(), g.epsCn.data(), g.n.data(), g.vdwKind,
                               g.isMartini, g.count, g.rCut, g.rCutCoulomb.data(), g.rCutLow,
                               g.rOn, g.alpha.data(), g.ewald, electrostatic, g.diElectric_1));
  if (!g.rMin.empty())
    CKS(gomcb200_init_exp6(g.e, g.rMin.data(), g.expConst.data(), g.rMaxSq.data(),
                           (int)g.rMin.size()));
  if (g.imageTotal) CKS(gomcb200_init_ewald(g.e, g.imageTotal, zero.data()));
  g.electrostatic = electrostatic;
}

void apply_lambda() {
  for (int b = 0; b < (int)BOX_TOTAL; ++b)
    CKS(gomcb200_update_lambda(g.e, b, g.lMol[b], -1, g.lVdw[b], g.lCoul[b], g.lFrac[b] ? 1 : 0));
  g.lambdaPending = false;
}

void ensure_topology(const std::vector<int> &particleKind, const std::vector<int> &particleMol,
                     const std::vector<double> &particleCharge) {
  if (g.haveTopo) return;
  g.nAtoms = (int)particleMol.size();
  g.nMols = g.nAtoms ? particleMol.back() + 1 : 0;
  g.molStart.assign(g.nMols + 1, g.nAtoms);
  for (int a = g.nAtoms - 1; a >= 0; --a) g.molStart[particleMol[a]] = a;  // atoms are contiguous
  g.particleMol = particleMol;
  CKS(gomcb200_init_topology(g.e, g.nAtoms, g.nMols, particleKind.data(), particleMol.data(),
                             particleCharge.data(), g.molStart.data()));
  g.haveTopo = true;
  if (g.lambdaPending) apply_lambda();
}

// molecules of the box = molecules of the atoms in its cell list
void ensure_membership(uint box, const std::vector<int> &cellVector) {
  std::vector<char> in(g.nMols, 0);
  for (int a : cellVector) in[g.particleMol[a]] = 1;
  std::vector<int> mols;
  for (int m = 0; m < g.nMols; ++m)
    if (in[m]) mols.push_back(m);
  if (g.haveBox[box] && mols == g.boxMols[box]) return;
  CKS(gomcb200_set_box_molecules(g.e, (int)box, mols.data(), (int)mols.size()));
  g.boxMols[box] = mols;
  g.haveBox[box] = true;
}

void set_dims(BoxDimensions const &boxAxes, uint box) {
  XYZ ax = boxAxes.GetAxis(box);
  double a3[3] = {ax.x, ax.y, ax.z};
  if (boxAxes.orthogonal[box]) {
    CKS(gomcb200_set_box_cell_basis(g.e, (int)box, nullptr, nullptr, a3));
    return;
  }
  const BoxDimensionsNonOrth &no = static_cast<const BoxDimensionsNonOrth &>(boxAxes);
  double cb[9], ci[9];
  for (int r = 0; r < 3; ++r) {
    XYZ v = no.cellBasis[box].Get(r), w = no.cellBasis_Inv[box].Get(r);
    cb[3 * r] = v.x; cb[3 * r + 1] = v.y; cb[3 * r + 2] = v.z;
    ci[3 * r] = w.x; ci[3 * r + 1] = w.y; ci[3 * r + 2] = w.z;
  }
  CKS(gomcb200_set_box_cell_basis(g.e, (int)box, cb, ci, a3));
}

void pair_prologue(const std::vector<int> &cellVector, XYZArray const &coords,
                   BoxDimensions const &boxAxes, bool electrostatic,
                   const std::vector<double> &particleCharge, const std::vector<int> &particleKind,
                   const std::vector<int> &particleMol, bool sc_coul, double sc_sigma_6,
                   double sc_alpha, uint sc_power, uint box) {
  if ((int)electrostatic != g.electrostatic) init_forcefield(electrostatic ? 1 : 0);
  ensure_topology(particleKind, particleMol, particleCharge);
  ensure_membership(box, cellVector);
  set_dims(boxAxes, box);
  CKS(gomcb200_init_softcore(g.e, sc_alpha, sc_sigma_6, (int)sc_power, sc_coul ? 1 : 0));
  CKS(gomcb200_set_coords(g.e, coords.x, coords.y, coords.z, 0, (int)coords.Count()));
}

}  // namespace

// ---- ConstantDefinitionsCUDAKernel.cuh ------------------------------------------------
void InitGPUForceField(VariablesCUDA &vars, double const *sigmaSq, double const *epsilon_Cn,
                       double const *n, int VDW_Kind, int isMartini, int count, double Rcut,
                       double RcutSq, double const *rCutCoulomb, double const *rCutCoulombSq,
                       double RcutLow, double Ron, double const *alpha, double const *alphaSq,
                       int ewald, double diElectric_1) {
  (void)vars; (void)RcutSq; (void)rCutCoulombSq; (void)alphaSq;
  if (!g.e && gomcb200_create(&g.e, -1, (int)BOX_TOTAL) != 0) die("gomcb200_create");
  const size_t sz = (size_t)count * count;
  g.sigmaSq.assign(sigmaSq, sigmaSq + sz);
  g.epsCn.assign(epsilon_Cn, epsilon_Cn + sz);
  g.n.assign(n, n + sz);
  g.rCutCoulomb.assign(rCutCoulomb, rCutCoulomb + BOX_TOTAL);
  g.alpha.assign(alpha, alpha + BOX_TOTAL);
  g.vdwKind = VDW_Kind; g.isMartini = isMartini; g.count = count; g.ewald = ewald;
  g.rCut = Rcut; g.rCutLow = RcutLow; g.rOn = Ron; g.diElectric_1 = diElectric_1;
  init_forcefield(1);  // `electrostatic` arrives with the first pair call
}

void InitCoordinatesCUDA(VariablesCUDA *, uint, uint, uint) {}

void InitEwaldVariablesCUDA(VariablesCUDA *, uint imageTotal) {
  g.imageTotal = (int)imageTotal;
  std::vector<double> zero(BOX_TOTAL, 0.0);  // the k lists come from the host (set_kvectors)
  CKS(gomcb200_init_ewald(g.e, g.imageTotal, zero.data()));
}

void InitExp6VariablesCUDA(VariablesCUDA *, double *rMin, double *expConst, double *rMaxSq,
                           uint size) {
  g.rMin.assign(rMin, rMin + size);
  g.expConst.assign(expConst, expConst + size);
  g.rMaxSq.assign(rMaxSq, rMaxSq + size);
  CKS(gomcb200_init_exp6(g.e, rMin, expConst, rMaxSq, (int)size));
}

void UpdateGPULambda(VariablesCUDA *, int *molIndex, double *lambdaVDW, double *lambdaCoulomb,
                     bool *isFraction) {
  for (int b = 0; b < (int)BOX_TOTAL; ++b) {
    g.lMol[b] = molIndex[b];
    g.lVdw[b] = lambdaVDW[b];
    g.lCoul[b] = lambdaCoulomb[b];
    g.lFrac[b] = isFraction[b];
  }
  g.lambdaPending = true;
  if (g.haveTopo) apply_lambda();
}

void CopyCurrentToRefCUDA(VariablesCUDA *, uint box, uint) { CKS(gomcb200_set_recip_ref(g.e, (int)box)); }
void CopyRefToNewCUDA(VariablesCUDA *, uint box, uint) { CKS(gomcb200_copy_recip(g.e, (int)box)); }
void UpdateRecipVecCUDA(VariablesCUDA *, uint box) { CKS(gomcb200_update_recip_vec(g.e, (int)box)); }
void UpdateRecipCUDA(VariablesCUDA *, uint box) { CKS(gomcb200_update_recip(g.e, (int)box)); }
// the pair calls read the box from their BoxDimensions argument
void UpdateCellBasisCUDA(VariablesCUDA *, uint, double *, double *, double *) {}
void UpdateInvCellBasisCUDA(VariablesCUDA *, uint, double *, double *, double *) {}
void DestroyEwaldCUDAVars(VariablesCUDA *) {}
void DestroyExp6CUDAVars(VariablesCUDA *) {}
void DestroyCUDAVars(VariablesCUDA *) {
  gomcb200_destroy(g.e);
  g.e = nullptr;
}

// ---- CUDAMemoryManager.cuh (MoleculeLookup.cpp:119-125 allocates one array itself) ------
long long CUDAMemoryManager::totalAllocatedBytes = 0;
std::unordered_map<void *, std::pair<unsigned int, std::string>> CUDAMemoryManager::allocatedPointers;
cudaError_t CUDAMemoryManager::mallocMemory(void **address, unsigned int size, std::string) {
  return cudaMalloc(address, size);
}
cudaError_t CUDAMemoryManager::freeMemory(void *address, std::string) { return cudaFree(address); }
bool CUDAMemoryManager::isFreed() { return true; }

// ---- CalculateEnergyCUDAKernel.cuh / CalculateForceCUDAKernel.cuh -----------------------
void CallBoxInterGPU(VariablesCUDA *, const std::vector<int> &cellVector, const std::vector<int> &,
                     const std::vector<std::vector<int>> &, XYZArray const &coords,
                     BoxDimensions const &boxAxes, bool electrostatic,
                     const std::vector<double> &particleCharge,
                     const std::vector<int> &particleKind, const std::vector<int> &particleMol,
                     double &REn, double &LJEn, bool sc_coul, double sc_sigma_6, double sc_alpha,
                     uint sc_power, uint const box) {
  pair_prologue(cellVector, coords, boxAxes, electrostatic, particleCharge, particleKind,
                particleMol, sc_coul, sc_sigma_6, sc_alpha, sc_power, box);
  CKS(gomcb200_box_inter(g.e, (int)box, &LJEn, &REn));
}

void CallBoxForceGPU(VariablesCUDA *, const std::vector<int> &cellVector, const std::vector<int> &,
                     const std::vector<std::vector<int>> &, const std::vector<int> &,
                     XYZArray const &coords, BoxDimensions const &boxAxes, bool electrostatic,
                     const std::vector<double> &particleCharge,
                     const std::vector<int> &particleKind, const std::vector<int> &particleMol,
                     double &REn, double &LJEn, double *aForcex, double *aForcey, double *aForcez,
                     double *mForcex, double *mForcey, double *mForcez, int atomCount,
                     int molCount, bool sc_coul, double sc_sigma_6, double sc_alpha,
                     uint sc_power, uint const box) {
  pair_prologue(cellVector, coords, boxAxes, electrostatic, particleCharge, particleKind,
                particleMol, sc_coul, sc_sigma_6, sc_alpha, sc_power, box);
  CKS(gomcb200_box_force(g.e, (int)box, &LJEn, &REn));
  // only this box's entries change (the host arrays hold both boxes)
  g.tx.resize(std::max(atomCount, molCount)); g.ty.resize(g.tx.size()); g.tz.resize(g.tx.size());
  CKS(gomcb200_get_forces(g.e, GOMCB200_ATOM_FORCE, g.tx.data(), g.ty.data(), g.tz.data(), 0, atomCount));
  for (int a : cellVector) { aForcex[a] = g.tx[a]; aForcey[a] = g.ty[a]; aForcez[a] = g.tz[a]; }
  CKS(gomcb200_get_forces(g.e, GOMCB200_MOL_FORCE, g.tx.data(), g.ty.data(), g.tz.data(), 0, molCount));
  for (int m : g.boxMols[box]) { mForcex[m] = g.tx[m]; mForcey[m] = g.ty[m]; mForcez[m] = g.tz[m]; }
}

void CallBoxInterForceGPU(VariablesCUDA *, const std::vector<int> &cellVector,
                          const std::vector<int> &, const std::vector<std::vector<int>> &,
                          const std::vector<int> &, XYZArray const &currentCoords,
                          XYZArray const &currentCOM, BoxDimensions const &boxAxes,
                          bool electrostatic, const std::vector<double> &particleCharge,
                          const std::vector<int> &particleKind, const std::vector<int> &particleMol,
                          double &rT11, double &rT12, double &rT13, double &rT22, double &rT23,
                          double &rT33, double &vT11, double &vT12, double &vT13, double &vT22,
                          double &vT23, double &vT33, bool sc_coul, double sc_sigma_6,
                          double sc_alpha, uint sc_power, uint const box) {
  pair_prologue(cellVector, currentCoords, boxAxes, electrostatic, particleCharge, particleKind,
                particleMol, sc_coul, sc_sigma_6, sc_alpha, sc_power, box);
  CKS(gomcb200_set_com(g.e, currentCOM.x, currentCOM.y, currentCOM.z, 0, (int)currentCOM.Count()));
  double vT[3], rT[3];
  CKS(gomcb200_box_inter_virial(g.e, (int)box, vT, rT));
  // the caller multiplies the Coulomb tensor by qqFact (src/CalculateEnergy.cpp:549-567); the
  // reference sums the diagonal only
  vT11 = vT[0]; vT22 = vT[1]; vT33 = vT[2];
  rT11 = rT[0] / num::qqFact; rT22 = rT[1] / num::qqFact; rT33 = rT[2] / num::qqFact;
  vT12 = vT13 = vT23 = rT12 = rT13 = rT23 = 0.0;
}

void CallVirialReciprocalGPU(VariablesCUDA *, XYZArray const &, XYZArray const &,
                             const std::vector<double> &, double &rT11, double &rT12,
                             double &rT13, double &rT22, double &rT23, double &rT33, uint, double,
                             uint box) {
  // coordinates and centres of mass are the ones CallBoxInterForceGPU just received
  // (CalculateEnergy::VirialCalc calls the two back to back, src/CalculateEnergy.cpp:449-575)
  double wT[3];
  CKS(gomcb200_virial_reciprocal(g.e, (int)box, wT));
  rT11 = wT[0]; rT22 = wT[1]; rT33 = wT[2];
  rT12 = rT13 = rT23 = 0.0;
}

// ---- CalculateEwaldCUDAKernel.cuh -------------------------------------------------------
void CallBoxReciprocalSetupGPU(VariablesCUDA *, XYZArray const &coords, double const *kx,
                               double const *ky, double const *kz,
                               const std::vector<double> &particleCharge, uint imageSize,
                               double *sumRnew, double *sumInew, double *prefact, double *hsqr,
                               double &energyRecip, uint box) {
  CKS(gomcb200_set_kvectors(g.e, (int)box, (int)imageSize, kx, ky, kz, hsqr, prefact));
  CKS(gomcb200_call_box_reciprocal_points(g.e, (int)box, 1, (int)coords.Count(), coords.x, coords.y,
                                          coords.z, particleCharge.data(), sumRnew, sumInew,
                                          &energyRecip));
}

void CallBoxReciprocalSumsGPU(VariablesCUDA *, XYZArray const &coords,
                              const std::vector<double> &particleCharge, uint, double *sumRnew,
                              double *sumInew, double &energyRecip, uint box) {
  CKS(gomcb200_call_box_reciprocal_points(g.e, (int)box, 0, (int)coords.Count(), coords.x, coords.y,
                                          coords.z, particleCharge.data(), sumRnew, sumInew,
                                          &energyRecip));
}

void CallMolReciprocalGPU(VariablesCUDA *, XYZArray const &currentCoords, XYZArray const &newCoords,
                          const std::vector<double> &particleCharge, uint, double *sumRnew,
                          double *sumInew, double &energyRecipNew, uint box) {
  CKS(gomcb200_call_mol_reciprocal(g.e, (int)box, (int)newCoords.Count(), particleCharge.data(),
                                   currentCoords.x, currentCoords.y, currentCoords.z, newCoords.x,
                                   newCoords.y, newCoords.z, sumRnew, sumInew, &energyRecipNew));
}

void CallChangeLambdaMolReciprocalGPU(VariablesCUDA *, XYZArray const &coords,
                                      const std::vector<double> &particleCharge, uint imageSize,
                                      double *sumRnew, double *sumInew, double &energyRecipNew,
                                      const double lambdaCoef, uint box) {
  CKS(gomcb200_mol_exchange_reciprocal(g.e, (int)box, (int)coords.Count(), particleCharge.data(),
                                       coords.x, coords.y, coords.z, 1, lambdaCoef,
                                       &energyRecipNew));
  CKS(gomcb200_get_recip_sums(g.e, (int)box, GOMCB200_SUM_NEW, sumRnew, sumInew, (int)imageSize));
}

void CallSwapReciprocalGPU(VariablesCUDA *, XYZArray const &coords,
                           const std::vector<double> &particleCharge, uint, double *sumRnew,
                           double *sumInew, const bool insert, double &energyRecipNew, uint box) {
  CKS(gomcb200_call_swap_reciprocal(g.e, (int)box, (int)coords.Count(), particleCharge.data(),
                                    coords.x, coords.y, coords.z, insert ? 1 : 0, sumRnew, sumInew,
                                    &energyRecipNew));
}

void CallMolExchangeReciprocalGPU(VariablesCUDA *, uint imageSize, double *sumRnew,
                                  double *sumInew, uint box) {
  CKS(gomcb200_set_recip_sums(g.e, (int)box, GOMCB200_SUM_NEW, sumRnew, sumInew, (int)imageSize));
}

void CallBoxForceReciprocalGPU(VariablesCUDA *, XYZArray &atomForceRec, XYZArray &molForceRec,
                               const std::vector<double> &, const std::vector<int> &particleMol,
                               const std::vector<bool> &, const bool *particleUsed,
                               const std::vector<int> &, const std::vector<int> &, double, uint,
                               XYZArray const &molCoords, BoxDimensions const &boxAxes, int box) {
  set_dims(boxAxes, (uint)box);
  CKS(gomcb200_set_coords(g.e, molCoords.x, molCoords.y, molCoords.z, 0, (int)molCoords.Count()));
  CKS(gomcb200_box_force_reciprocal(g.e, box));
  const int nA = (int)atomForceRec.Count(), nM = (int)molForceRec.Count();
  g.tx.resize(std::max(nA, nM)); g.ty.resize(g.tx.size()); g.tz.resize(g.tx.size());
  CKS(gomcb200_get_forces(g.e, GOMCB200_ATOM_FORCE_REC, g.tx.data(), g.ty.data(), g.tz.data(), 0, nA));
  for (int a = 0; a < nA; ++a)
    if (particleUsed[a]) atomForceRec.Set(a, g.tx[a], g.ty[a], g.tz[a]);
  CKS(gomcb200_get_forces(g.e, GOMCB200_MOL_FORCE_REC, g.tx.data(), g.ty.data(), g.tz.data(), 0, nM));
  std::vector<char> done(nM, 0);
  for (int a = 0; a < nA; ++a) {
    const int m = particleMol[a];
    if (particleUsed[a] && !done[m]) {
      molForceRec.Set(m, g.tx[m], g.ty[m], g.tz[m]);
      done[m] = 1;
    }
  }
}

// ---- TransformParticlesCUDAKernel.cuh ---------------------------------------------------
namespace {
// trial coordinates / COMs / displacement of the transform just run -> the caller's arrays
void fetch_trial(XYZArray &newMolPos, XYZArray &newCOMs, XYZArray &k, std::vector<int> *inForceRange) {
  CKS(gomcb200_mp_select(g.e, 1));
  CKS(gomcb200_get_coords(g.e, newMolPos.x, newMolPos.y, newMolPos.z, 0, (int)newMolPos.Count()));
  CKS(gomcb200_get_com(g.e, newCOMs.x, newCOMs.y, newCOMs.z, 0, (int)newCOMs.Count()));
  CKS(gomcb200_mp_select(g.e, 0));
  std::vector<int> ifr(g.nMols);
  CKS(gomcb200_mp_get_trial(g.e, k.x, k.y, k.z, ifr.data()));
  if (inForceRange)
    for (int m = 0; m < g.nMols && m < (int)inForceRange->size(); ++m) (*inForceRange)[m] = ifr[m];
}
void push_state(XYZArray &newMolPos, XYZArray &newCOMs) {
  CKS(gomcb200_set_coords(g.e, newMolPos.x, newMolPos.y, newMolPos.z, 0, (int)newMolPos.Count()));
  CKS(gomcb200_set_com(g.e, newCOMs.x, newCOMs.y, newCOMs.z, 0, (int)newCOMs.Count()));
}
}  // namespace

void CallTranslateParticlesGPU(VariablesCUDA *, const std::vector<int8_t> &isMoleculeInvolved,
                               int box, double t_max, double *mForcex, double *mForcey,
                               double *mForcez, std::vector<int> &inForceRange, ulong step,
                               unsigned int key, ulong seed, const std::vector<int> &, int,
                               int molCount, double, double, double, XYZArray &newMolPos,
                               XYZArray &newCOMs, double lambdaBETA, XYZArray &t_k,
                               XYZArray &molForceRecRef) {
  push_state(newMolPos, newCOMs);
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_FORCE, mForcex, mForcey, mForcez, 0, molCount));
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_FORCE_REC, molForceRecRef.x, molForceRecRef.y,
                          molForceRecRef.z, 0, molCount));
  CKS(gomcb200_mp_transform(g.e, box, 0, t_max, lambdaBETA, step, key, seed,
                            reinterpret_cast<const signed char *>(isMoleculeInvolved.data())));
  fetch_trial(newMolPos, newCOMs, t_k, &inForceRange);
}

void CallRotateParticlesGPU(VariablesCUDA *, const std::vector<int8_t> &isMoleculeInvolved, int box,
                            double r_max, double *mTorquex, double *mTorquey, double *mTorquez,
                            std::vector<int> &inForceRange, ulong step, unsigned int key,
                            ulong seed, const std::vector<int> &, int, int molCount, double,
                            double, double, XYZArray &newMolPos, XYZArray &newCOMs,
                            double lambdaBETA, XYZArray &r_k) {
  push_state(newMolPos, newCOMs);
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_TORQUE, mTorquex, mTorquey, mTorquez, 0, molCount));
  CKS(gomcb200_mp_transform(g.e, box, 1, r_max, lambdaBETA, step, key, seed,
                            reinterpret_cast<const signed char *>(isMoleculeInvolved.data())));
  fetch_trial(newMolPos, newCOMs, r_k, &inForceRange);
}

void BrownianMotionRotateParticlesGPU(VariablesCUDA *, const std::vector<unsigned int> &moleculeInvolved,
                                      XYZArray &mTorque, XYZArray &newMolPos, XYZArray &newCOMs,
                                      XYZArray &r_k, const XYZ &, const double BETA,
                                      const double r_max, ulong step, unsigned int key, ulong seed,
                                      const int box, const bool) {
  push_state(newMolPos, newCOMs);
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_TORQUE, mTorque.x, mTorque.y, mTorque.z, 0,
                          (int)mTorque.Count()));
  std::vector<signed char> inv(g.nMols, 0);
  for (unsigned m : moleculeInvolved) inv[m] = 1;
  CKS(gomcb200_bm_transform(g.e, box, 1, r_max, BETA, step, key, seed, inv.data()));
  fetch_trial(newMolPos, newCOMs, r_k, nullptr);
}

void BrownianMotionTranslateParticlesGPU(VariablesCUDA *, const std::vector<unsigned int> &moleculeInvolved,
                                         XYZArray &mForce, XYZArray &mForceRec, XYZArray &newMolPos,
                                         XYZArray &newCOMs, XYZArray &t_k, const XYZ &,
                                         const double BETA, const double t_max, ulong step,
                                         unsigned int key, ulong seed, const int box, const bool) {
  push_state(newMolPos, newCOMs);
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_FORCE, mForce.x, mForce.y, mForce.z, 0, (int)mForce.Count()));
  CKS(gomcb200_set_forces(g.e, GOMCB200_MOL_FORCE_REC, mForceRec.x, mForceRec.y, mForceRec.z, 0,
                          (int)mForceRec.Count()));
  std::vector<signed char> inv(g.nMols, 0);
  for (unsigned m : moleculeInvolved) inv[m] = 1;
  CKS(gomcb200_bm_transform(g.e, box, 0, t_max, BETA, step, key, seed, inv.data()));
  fetch_trial(newMolPos, newCOMs, t_k, nullptr);
}
