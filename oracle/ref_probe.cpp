/*
 * ref_probe.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Harness that drives the UNMODIFIED GOMC reference (compiled from
 * /root/reference by oracle/ref_build.mk) through its own L4 interfaces
 * (CalculateEnergy / Ewald) and dumps inputs and outputs
 * as named binary arrays.  The dumps pin oracle/gomc_oracle.c and become the
 * golden vectors in tests/golden/ (see oracle/make_golden.py).
 *
 *   gomc_probe_<ENS> golden <in.conf> <out.bin> [nMoves] [seed]
 *       full Simulation construction; every function of the hot path.
 *   gomc_probe_<ENS> slab <in.conf> <out.bin> <nSel> <seed> [coords]
 *       light initialisation; BoxInter over the full box and BoxReciprocalSums on nSel
 *       k-vectors spread over the whole list (full-size parity pins, tests/golden/full_*).
 *   gomc_probe_<ENS> time <in.conf> <out.bin> <kFraction> <reps> [force]
 *       light initialisation (no full structure-factor build, no virial) and
 *       wall-clock timing of BoxInter + BoxReciprocalSums(k-slab) +
 *       BoxReciprocal; with "force" also BoxForce + BoxForceReciprocal.
 *
 * Thread count comes from OMP_NUM_THREADS (the reference's "+pN").
 * Private members are reached with the usual "#define private public" trick;
 * no reference source is modified or copied.
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#define private public
#define protected public
#include "Simulation.h"
#include "Setup.h"
#include "CalculateEnergy.h"
#include "Ewald.h"
#include "EwaldCached.h"
#include "NoEwald.h"
#include "FFParticle.h"
#include "FFExp6.h"
#include "BoxDimensionsNonOrth.h"
#include "TrialMol.h"
#include "MultiParticle.h"
#include "MultiParticleBrownianMotion.h"
#if ENSEMBLE == NPT || ENSEMBLE == GEMC
#include "VolumeTransfer.h"
#endif
#undef private
#undef protected

namespace {

struct Dump {
  FILE *f;
  explicit Dump(const char *path) {
    f = fopen(path, "wb");
    if (!f) {
      perror(path);
      exit(2);
    }
    fwrite("GOMCDUMP", 1, 8, f);
  }
  ~Dump() { fclose(f); }
  void raw(const std::string &name, unsigned char dtype, uint64_t count,
           const void *data, size_t elem) {
    uint32_t nl = (uint32_t)name.size();
    fwrite(&nl, 4, 1, f);
    fwrite(name.data(), 1, nl, f);
    fwrite(&dtype, 1, 1, f);
    fwrite(&count, 8, 1, f);
    fwrite(data, elem, count, f);
  }
  void f64(const std::string &n, const double *d, size_t c) {
    raw(n, 0, c, d, 8);
  }
  void f64(const std::string &n, const std::vector<double> &v) {
    raw(n, 0, v.size(), v.data(), 8);
  }
  void f64(const std::string &n, double v) { raw(n, 0, 1, &v, 8); }
  void i32(const std::string &n, const std::vector<int> &v) {
    raw(n, 1, v.size(), v.data(), 4);
  }
  void i32(const std::string &n, int v) { raw(n, 1, 1, &v, 4); }
  void xyz(const std::string &n, const XYZArray &a) {
    f64(n + ".x", a.x, a.Count());
    f64(n + ".y", a.y, a.Count());
    f64(n + ".z", a.z, a.Count());
  }
};

double now() {
  return std::chrono::duration<double>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

std::string bname(const char *s, uint b) {
  return std::string("box") + std::to_string(b) + "." + s;
}

void dump_static(Dump &out, StaticVals &sv, System &sys) {
  Forcefield &ff = sv.forcefield;
  FFParticle &fp = *ff.particles;
  Molecules &mols = sv.mol;
  uint nAtoms = sys.coordinates.Count();
  out.i32("nAtoms", (int)nAtoms);
  out.i32("nMols", (int)mols.count);
  out.i32("boxTotal", (int)BOX_TOTAL);
  out.i32("boxesWithU", (int)BOXES_WITH_U_NB);
  out.i32("ff.vdwKind", (int)ff.vdwKind);
  out.i32("ff.isMartini", (int)ff.isMartini);
  out.i32("ff.exp6", (int)ff.exp6);
  out.i32("ff.ewald", (int)ff.ewald);
  out.i32("ff.electrostatic", (int)ff.electrostatic);
  out.i32("ff.useLRC", (int)ff.useLRC);
  out.i32("ff.kindCount", (int)fp.count);
  out.f64("ff.rCut", ff.rCut);
  out.f64("ff.rCutLow", ff.rCutLow);
  out.f64("ff.rswitch", ff.rswitch);
  out.f64("ff.tolerance", ff.tolerance);
  out.f64("ff.rCutCoulomb", ff.rCutCoulomb, BOX_TOTAL);
  out.f64("ff.alpha", ff.alpha, BOX_TOTAL);
  out.f64("ff.recip_rcut", ff.recip_rcut, BOX_TOTAL);
  out.f64("ff.sigmaSq", fp.sigmaSq, fp.count * fp.count);
  out.f64("ff.epsilon_cn", fp.epsilon_cn, fp.count * fp.count);
  out.f64("ff.n", fp.n, fp.count * fp.count);
  out.f64("ff.dielectric", ff.dielectric);
  if (ff.exp6) {
    FF_EXP6 &fe = static_cast<FF_EXP6 &>(fp);
    out.f64("ff.rMin", fe.rMin, fp.count * fp.count);
    out.f64("ff.expConst", fe.expConst, fp.count * fp.count);
    out.f64("ff.rMaxSq", fe.rMaxSq, fp.count * fp.count);
  }
  out.xyz("coords", sys.coordinates);
  out.xyz("com", sys.com);
  out.i32("particleKind", sys.calcEnergy.particleKind);
  out.i32("particleMol", sys.calcEnergy.particleMol);
  out.f64("particleCharge", sys.calcEnergy.particleCharge);
  std::vector<int> start(mols.count + 1), kidx(mols.count);
  for (uint m = 0; m <= mols.count; ++m) start[m] = (int)mols.start[m];
  for (uint m = 0; m < mols.count; ++m) kidx[m] = (int)mols.kIndex[m];
  out.i32("molStart", start);
  out.i32("molKindIndex", kidx);
  out.i32("nMolKinds", (int)mols.kindsCount);
  std::vector<int> mkStart(1, 0), mkAtomKinds;
  for (uint k = 0; k < mols.kindsCount; ++k) {
    for (uint a = 0; a < mols.kinds[k].NumAtoms(); ++a)
      mkAtomKinds.push_back((int)mols.kinds[k].AtomKind(a));
    mkStart.push_back((int)mkAtomKinds.size());
  }
  out.i32("molKindStart", mkStart);
  out.i32("molKindAtomKinds", mkAtomKinds);
  out.f64("pairEnCorrections", mols.pairEnCorrections,
          mols.kindsCount * mols.kindsCount);
  for (uint b = 0; b < BOX_TOTAL; ++b) {
    XYZ ax = sys.boxDimRef.axis.Get(b);
    double a3[3] = {ax.x, ax.y, ax.z};
    out.f64(bname("axis", b), a3, 3);
    out.i32(bname("orthogonal", b), (int)sys.boxDimRef.orthogonal[b]);
    out.f64(bname("boxRcut", b), sys.boxDimRef.rCut[b]);
    out.f64(bname("volume", b), sys.boxDimRef.volume[b]);
    if (!sys.boxDimRef.orthogonal[b]) {
      BoxDimensionsNonOrth &no = static_cast<BoxDimensionsNonOrth &>(sys.boxDimRef);
      double cb[9], ci[9];
      for (int r = 0; r < 3; ++r) {
        XYZ v = no.cellBasis[b].Get(r), w = no.cellBasis_Inv[b].Get(r);
        cb[3 * r] = v.x; cb[3 * r + 1] = v.y; cb[3 * r + 2] = v.z;
        ci[3 * r] = w.x; ci[3 * r + 1] = w.y; ci[3 * r + 2] = w.z;
      }
      out.f64(bname("cellBasis", b), cb, 9);
      out.f64(bname("cellBasisInv", b), ci, 9);
    }
    std::vector<int> molsInBox, numKind;
    MoleculeLookup::box_iterator it = sys.molLookupRef.BoxBegin(b),
                                 end = sys.molLookupRef.BoxEnd(b);
    while (it != end) {
      molsInBox.push_back((int)*it);
      ++it;
    }
    for (uint k = 0; k < mols.kindsCount; ++k)
      numKind.push_back((int)sys.molLookupRef.NumKindInBox(k, b));
    out.i32(bname("mols", b), molsInBox);
    out.i32(bname("numKindInBox", b), numKind);
  }
}

void dump_kvectors(Dump &out, Ewald &ew, uint b, bool sums) {
  uint nk = ew.imageSizeRef[b];
  out.i32(bname("nk", b), (int)nk);
  out.i32(bname("kmax", b), (int)ew.kmax[b]);
  out.i32(bname("imageTotal", b), (int)ew.imageTotal);
  out.f64(bname("kx", b), ew.kxRef[b], nk);
  out.f64(bname("ky", b), ew.kyRef[b], nk);
  out.f64(bname("kz", b), ew.kzRef[b], nk);
  out.f64(bname("hsqr", b), ew.hsqrRef[b], nk);
  out.f64(bname("prefact", b), ew.prefactRef[b], nk);
  if (sums) {
    out.f64(bname("sumRref", b), ew.sumRref[b], nk);
    out.f64(bname("sumIref", b), ew.sumIref[b], nk);
  }
}

int run_golden(int argc, char **argv) {
  const char *conf = argv[2];
  const char *outPath = argv[3];
  int nMoves = argc > 4 ? atoi(argv[4]) : 4;
  unsigned seed = argc > 5 ? (unsigned)atoi(argv[5]) : 7u;
  Simulation sim(conf);
  System &sys = *sim.system;
  StaticVals &sv = *sim.staticValues;
  Forcefield &ff = sv.forcefield;
  Molecules &mols = sv.mol;
  Ewald &ew = *sys.calcEwald;
  CalculateEnergy &ce = sys.calcEnergy;
  Dump out(outPath);
  out.i32("threads", omp_get_max_threads());
  // optional fractional molecule in box 0 (free energy / NeMTMC state):
  //   <pick> <lambdaVDW> <lambdaCoulomb> <sc_alpha> <sc_sigma> <sc_power> <sc_coul>
  if (argc > 12) {
    std::vector<uint> inBox0;
    for (MoleculeLookup::box_iterator it = sys.molLookupRef.BoxBegin(0);
         it != sys.molLookupRef.BoxEnd(0); ++it)
      inBox0.push_back(*it);
    uint lm = inBox0[(size_t)atoi(argv[6]) % inBox0.size()];
    double lv = atof(argv[7]), lc = atof(argv[8]);
    Forcefield &ffw = const_cast<Forcefield &>(sv.forcefield);
    ffw.sc_alpha = atof(argv[9]);
    ffw.sc_sigma = atof(argv[10]);
    ffw.sc_sigma_6 = pow(ffw.sc_sigma, 6.0); // src/Forcefield.cpp:75
    ffw.sc_power = (uint)atoi(argv[11]);
    ffw.sc_coul = atoi(argv[12]) != 0;
    sys.lambdaRef.Set(lv, lc, lm, mols.GetMolKind(lm), 0);
    // the reference state System::Init built with lambda = 1 is rebuilt
    sys.calcEwald->UpdateVectorsAndRecipTerms(false);
    sys.potential = sys.calcEnergy.SystemTotal();
    double lp[8] = {(double)lm, lv, lc, ffw.sc_alpha, ffw.sc_sigma_6,
                    (double)ffw.sc_power, (double)ffw.sc_coul, (double)mols.GetMolKind(lm)};
    out.f64("lambda.params", lp, 8);
  }
  dump_static(out, sv, sys);
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);

  for (uint b = 0; b < BOXES_WITH_U_NB; ++b) {
    MoleculeLookup::box_iterator it = sys.molLookupRef.BoxBegin(b),
                                 end = sys.molLookupRef.BoxEnd(b);
    std::vector<uint> molsInBox;
    while (it != end) {
      molsInBox.push_back(*it);
      ++it;
    }
    if (molsInBox.empty()) continue;
    // ---- SystemTotal pieces as computed by System::Init ------------------
    const Energy &e0 = sys.potential.boxEnergy[b];
    double sysEn[8] = {e0.inter, e0.real,       e0.recip,      e0.self,
                       e0.correction, e0.tailCorrection, e0.intraBond,
                       e0.intraNonbond};
    out.f64(bname("systemTotal", b), sysEn, 8);
    // ---- BoxInter --------------------------------------------------------
    SystemPotential pot =
        ce.BoxInter(SystemPotential(), sys.coordinates, sys.boxDimRef, b);
    out.f64(bname("BoxInter.inter", b), pot.boxEnergy[b].inter);
    out.f64(bname("BoxInter.real", b), pot.boxEnergy[b].real);
    out.f64(bname("BoxInter.tailCorrection", b),
            pot.boxEnergy[b].tailCorrection);
    // ---- VirialCalc (pair tensors + tail correction + VirialReciprocal) ----
    {
      Virial v = ce.VirialCalc(b);
      double it[3] = {v.interTens[0][0], v.interTens[1][1], v.interTens[2][2]};
      double rt[3] = {v.realTens[0][0], v.realTens[1][1], v.realTens[2][2]};
      double wt[3] = {v.recipTens[0][0], v.recipTens[1][1], v.recipTens[2][2]};
      out.f64(bname("Virial.interTens", b), it, 3);
      out.f64(bname("Virial.realTens", b), rt, 3);
      out.f64(bname("Virial.recipTens", b), wt, 3);
      double sc[5] = {v.inter, v.real, v.recip, v.tailCorrection, v.total};
      out.f64(bname("Virial.scalars", b), sc, 5);
    }
    // ---- BoxForce (needs multiParticleEnabled for ResetForce) -------------
    {
      XYZArray aF(sys.coordinates.Count()), mF(mols.count);
      aF.Reset();
      mF.Reset();
      SystemPotential pf = ce.BoxForce(SystemPotential(), sys.coordinates, aF,
                                       mF, sys.boxDimRef, b);
      out.f64(bname("BoxForce.inter", b), pf.boxEnergy[b].inter);
      out.f64(bname("BoxForce.real", b), pf.boxEnergy[b].real);
      out.xyz(bname("BoxForce.atomForce", b), aF);
      out.xyz(bname("BoxForce.molForce", b), mF);
      if (ff.ewald) {
        // ---- reciprocal force + torque -------------------------------------
        XYZArray aR(sys.coordinates.Count()), mR(mols.count), tq(mols.count);
        aR.Reset();
        mR.Reset();
        tq.Reset();
        ew.CopyRecip(b);
        ew.BoxForceReciprocal(sys.coordinates, aR, mR, b);
        out.xyz(bname("BoxForceReciprocal.atomForceRec", b), aR);
        out.xyz(bname("BoxForceReciprocal.molForceRec", b), mR);
        ce.CalculateTorque(molsInBox, sys.coordinates, sys.com, aF, aR, tq, b);
        out.xyz(bname("CalculateTorque.molTorque", b), tq);
      }
    }
    // ---- Ewald terms -------------------------------------------------------
    if (ff.ewald) {
      dump_kvectors(out, ew, b, true);
      ew.CopyRecip(b);
      out.f64(bname("BoxReciprocal", b), ew.BoxReciprocal(b, false));
      double corr = 0.0;
      for (uint m : molsInBox) corr += ew.MolCorrection(m, b);
      out.f64(bname("MolCorrection.sum", b), corr);
      out.f64(bname("BoxSelf", b), ew.BoxSelf(b));
      // BoxReciprocalSums on the current coordinates must reproduce sumRref
      ew.BoxReciprocalSums(b, sys.coordinates);
      out.f64(bname("BoxReciprocalSums.sumRnew", b), ew.sumRnew[b],
              ew.imageSizeRef[b]);
      out.f64(bname("BoxReciprocalSums.sumInew", b), ew.sumInew[b],
              ew.imageSizeRef[b]);
    }
    // ---- single-molecule moves: MoleculeInter + MolReciprocal -------------
    XYZ ax = sys.boxDimRef.axis.Get(b);
    std::vector<int> mvMol;
    std::vector<double> mvX, mvY, mvZ, mvLJ, mvReal, mvRecip;
    std::vector<int> mvOverlap, mvStart;
    for (int t = 0; t < nMoves; ++t) {
      uint m = molsInBox[(size_t)(rng() % molsInBox.size())];
      uint start = mols.MolStart(m), len = mols.GetKind(m).NumAtoms();
      XYZArray nc(len);
      // rigid displacement, large for odd t (tests PBC wrap), small otherwise;
      // every 4th move also puts the molecule on top of a neighbour to
      // exercise the overlap flag.
      // (slanted cells: keep the unslant displacement below one cell length, the
      // reference's WrapPBC wraps only once)
      double big = sys.boxDimRef.orthogonal[b] ? 0.45 : 0.2;
      double amp = (t % 2) ? big * std::min(ax.x, std::min(ax.y, ax.z)) : 0.4;
      XYZ shift(U(rng) * amp, U(rng) * amp, U(rng) * amp);
      if (t % 4 == 3 && molsInBox.size() > 1) {
        uint other = molsInBox[(size_t)(rng() % molsInBox.size())];
        if (other != m) {
          XYZ d = sys.coordinates.Get(mols.MolStart(other)) -
                  sys.coordinates.Get(start);
          shift = d + XYZ(0.3, 0.2, 0.1);
        }
      }
      for (uint a = 0; a < len; ++a) {
        XYZ p = sys.coordinates.Get(start + a) + shift;
        p = sys.boxDimRef.WrapPBC(p, b);
        nc.Set(a, p);
      }
      Intermolecular iLJ, iReal;
      sys.cellList.RemoveMol(m, b, sys.coordinates);
      bool overlap = ce.MoleculeInter(iLJ, iReal, nc, m, b);
      double dRecip = 0.0;
      if (ff.ewald) dRecip = ew.MolReciprocal(nc, m, b);
      if (ff.ewald && t == 0) {
        out.f64(bname("MolReciprocal0.sumRnew", b), ew.sumRnew[b],
                ew.imageSizeRef[b]);
        out.f64(bname("MolReciprocal0.sumInew", b), ew.sumInew[b],
                ew.imageSizeRef[b]);
      }
      // reject: state stays as it was (non-cached: nothing; cached: RestoreMol)
      ew.RestoreMol(m);
      sys.cellList.AddMol(m, b, sys.coordinates);
      mvMol.push_back((int)m);
      mvStart.push_back((int)mvX.size());
      for (uint a = 0; a < len; ++a) {
        mvX.push_back(nc.x[a]);
        mvY.push_back(nc.y[a]);
        mvZ.push_back(nc.z[a]);
      }
      mvLJ.push_back(iLJ.energy);
      mvReal.push_back(iReal.energy);
      mvRecip.push_back(dRecip);
      mvOverlap.push_back((int)overlap);
    }
    mvStart.push_back((int)mvX.size());
    out.i32(bname("move.mol", b), mvMol);
    out.i32(bname("move.start", b), mvStart);
    out.f64(bname("move.x", b), mvX);
    out.f64(bname("move.y", b), mvY);
    out.f64(bname("move.z", b), mvZ);
    out.f64(bname("move.dLJ", b), mvLJ);
    out.f64(bname("move.dReal", b), mvReal);
    out.f64(bname("move.dRecip", b), mvRecip);
    out.i32(bname("move.overlap", b), mvOverlap);
    out.f64(bname("sysPotRef.recip", b), sys.potential.boxEnergy[b].recip);

    // ---- swap-type deltas: SwapDestRecip / SwapSourceRecip / corrections ---
    {
      uint m = molsInBox[(size_t)(rng() % molsInBox.size())];
      uint start = mols.MolStart(m), len = mols.GetKind(m).NumAtoms();
      cbmc::TrialMol newMol(mols.GetKind(m), sys.boxDimRef, b);
      cbmc::TrialMol oldMol(mols.GetKind(m), sys.boxDimRef, b);
      XYZArray nc(len);
      XYZ shift(U(rng) * 3.0, U(rng) * 3.0, U(rng) * 3.0);
      for (uint a = 0; a < len; ++a) {
        XYZ p = sys.coordinates.Get(start + a) + shift;
        nc.Set(a, sys.boxDimRef.WrapPBC(p, b));
      }
      newMol.SetCoords(nc, 0);
      oldMol.SetCoords(sys.coordinates, start);
      out.i32(bname("swap.mol", b), (int)m);
      out.xyz(bname("swap.newCoords", b), nc);
      if (ff.ewald) {
        double dDest = ew.SwapDestRecip(newMol, b, m);
        out.f64(bname("SwapDestRecip", b), dDest);
        out.f64(bname("SwapDestRecip.sumRnew", b), ew.sumRnew[b],
                ew.imageSizeRef[b]);
        // source right after destination, as MoleculeTransfer::CalcEn calls them
        // (src/moves/MoleculeTransfer.h:127-129): the cached class hands the molecule's
        // old cos/sin rows from one to the other; the rejection (RestoreMol) comes after
        double dSrc = ew.SwapSourceRecip(oldMol, b, m);
        out.f64(bname("SwapSourceRecip", b), dSrc);
        out.f64(bname("SwapSourceRecip.sumInew", b), ew.sumInew[b],
                ew.imageSizeRef[b]);
        ew.RestoreMol(m);
        out.f64(bname("SwapCorrection.new", b), ew.SwapCorrection(newMol));
        out.f64(bname("SwapCorrection.old", b), ew.SwapCorrection(oldMol));
        out.f64(bname("SwapSelf", b), ew.SwapSelf(newMol));
      }
      // ---- ParticleInter: trial positions of the first atom of molecule m --
      const uint trials = 6;
      XYZArray tp(trials);
      std::vector<double> en(trials, 0.0), re(trials, 0.0);
      bool *ov = new bool[trials];
      for (uint t = 0; t < trials; ++t) {
        ov[t] = false;
        XYZ p(0.5 * (U(rng) + 1.0) * ax.x, 0.5 * (U(rng) + 1.0) * ax.y,
              0.5 * (U(rng) + 1.0) * ax.z);
        if (t == trials - 1) // on top of another atom: overlap
          p = sys.boxDimRef.WrapPBC(
              sys.coordinates.Get(
                  mols.MolStart(molsInBox[(m == molsInBox[0]) ? 1 : 0])) +
                  XYZ(0.2, 0.1, 0.05),
              b);
        tp.Set(t, p);
      }
      sys.cellList.RemoveMol(m, b, sys.coordinates);
      ce.ParticleInter(en.data(), re.data(), tp, ov, 0, m, b, trials);
      sys.cellList.AddMol(m, b, sys.coordinates);
      std::vector<int> ovi(trials);
      for (uint t = 0; t < trials; ++t) ovi[t] = ov[t];
      delete[] ov;
      out.xyz(bname("ParticleInter.trialPos", b), tp);
      out.f64(bname("ParticleInter.en", b), en);
      out.f64(bname("ParticleInter.real", b), re);
      out.i32(bname("ParticleInter.overlap", b), ovi);
      // ---- CBMC growth of the LAST site of a chain molecule (DCSingle / DCLinkNoDih call
      // ParticleInter and ParticleNonbonded back to back on the same trial positions):
      // sites 0 .. len-2 of molecule m exist, the last one is tried at `trials2` positions
      // one bond length from its neighbour
      if (len >= 5) {
        const uint part = len - 1, trials2 = 10;
        cbmc::TrialMol grow(mols.GetKind(m), sys.boxDimRef, b);
        for (uint a = 0; a + 1 < len; ++a) grow.AddAtom(a, sys.coordinates.Get(start + a));
        XYZArray tp2(trials2);
        XYZ anchor = sys.coordinates.Get(start + len - 2);
        for (uint t = 0; t < trials2; ++t) {
          XYZ d(U(rng), U(rng), U(rng));
          d *= 1.54 / d.Length();
          if (t == 0)   // fold the chain back: close to site 0, inside its cut-off for sure
            d = (sys.boxDimRef.MinImage(sys.coordinates.Get(start) - anchor, b)) * 0.45;
          tp2.Set(t, sys.boxDimRef.WrapPBC(anchor + d, b));
        }
        std::vector<double> nb(trials2, 0.0), en2(trials2, 0.0), re2(trials2, 0.0);
        bool *ov2 = new bool[trials2];
        for (uint t = 0; t < trials2; ++t) ov2[t] = false;
        ce.ParticleNonbonded(nb.data(), grow, tp2, part, b, trials2);
        sys.cellList.RemoveMol(m, b, sys.coordinates);
        ce.ParticleInter(en2.data(), re2.data(), tp2, ov2, part, m, b, trials2);
        sys.cellList.AddMol(m, b, sys.coordinates);
        std::vector<int> ovi2(trials2);
        for (uint t = 0; t < trials2; ++t) ovi2[t] = ov2[t];
        delete[] ov2;
        {  // the partner list ParticleNonbonded walks: sortedNB(part) filtered by AtomExists
          std::vector<int> partners;
          const MoleculeKind &mkd = mols.GetKind(m);
          for (const uint *pp = mkd.sortedNB.Begin(part); pp != mkd.sortedNB.End(part); ++pp)
            if (grow.AtomExists(*pp)) partners.push_back((int)*pp);
          out.i32(bname("grow.partners", b), partners);
        }
        out.i32(bname("grow.mol", b), (int)m);
        out.i32(bname("grow.part", b), (int)part);
        out.xyz(bname("grow.trialPos", b), tp2);
        out.f64(bname("grow.ParticleNonbonded", b), nb);
        out.f64(bname("grow.ParticleInter.en", b), en2);
        out.f64(bname("grow.ParticleInter.real", b), re2);
        out.i32(bname("grow.ParticleInter.overlap", b), ovi2);
      }
    }

    // ---- MultiParticle move: trial transform, CalcEn, acceptance weight ----
    // Drives the move object's own methods (private members opened above) the way
    // MultiParticle::Prep / Transform / CalcEn / Accept do, with a fixed box and
    // move type instead of the PRNG draws.
    if (b == 0 && sys.moves[mv::MULTIPARTICLE] != NULL) {
      MultiParticle *mp = static_cast<MultiParticle *>(sys.moves[mv::MULTIPARTICLE]);
      const int nTypes = mp->allTranslate ? 1 : 2;
      for (int type = 0; type < nTypes; ++type) {
        const char *tn = type == mp::MPROTATE ? "mpRotate" : "mpDisplace";
        std::string pre = std::string(tn) + ".";
        ulong step = 4242 + 17 * type;
        sys.r123wrapper.SetStep(step);
        mp->bPick = b;
        mp->moveType = type;
        mp->SetMolInBox(b);
        std::fill(mp->inForceRange.begin(), mp->inForceRange.end(), false);
        // reference forces / torques of the current positions (Prep, :219-236)
        if (ff.ewald) {
          ew.CopyRecip(b);
          ew.BoxForceReciprocal(sys.coordinates, sys.atomForceRecRef,
                                sys.molForceRecRef, b);
        }
        ce.BoxForce(sys.potential, sys.coordinates, sys.atomForceRef,
                    sys.molForceRef, sys.boxDimRef, b);
        ce.CalculateTorque(mp->moleculeIndex, sys.coordinates, sys.com,
                           sys.atomForceRef, sys.atomForceRecRef,
                           mp->molTorqueRef, b);
        sys.coordinates.CopyRange(mp->newMolsPos, 0, 0, sys.coordinates.Count());
        sys.com.CopyRange(mp->newCOMs, 0, 0, sys.com.Count());
        mp->CalculateTrialDistRot();
        double par[6] = {mp->moveSetRef.GetTMAX(b), mp->moveSetRef.GetRMAX(b),
                         mp->lambda * mp->BETA, (double)step,
                         (double)sys.r123wrapper.GetSeedValue(),
                         (double)sys.r123wrapper.GetKeyValue()};
        out.f64(bname((pre + "params").c_str(), b), par, 6);
        out.xyz(bname((pre + "molForceRef").c_str(), b), sys.molForceRef);
        out.xyz(bname((pre + "molForceRecRef").c_str(), b), sys.molForceRecRef);
        out.xyz(bname((pre + "molTorqueRef").c_str(), b), mp->molTorqueRef);
        out.xyz(bname((pre + "k").c_str(), b), type == mp::MPROTATE ? mp->r_k : mp->t_k);
        out.i32(bname((pre + "inForceRange").c_str(), b), mp->inForceRange);
        out.xyz(bname((pre + "newMolsPos").c_str(), b), mp->newMolsPos);
        out.xyz(bname((pre + "newCOMs").c_str(), b), mp->newCOMs);
        mp->CalcEn();
        double w = mp->GetCoeff();
        out.f64(bname((pre + "wRatio").c_str(), b), w);
        double en[3] = {mp->sysPotNew.boxEnergy[b].inter,
                        mp->sysPotNew.boxEnergy[b].real,
                        mp->sysPotNew.boxEnergy[b].recip};
        out.f64(bname((pre + "newEnergy").c_str(), b), en, 3);
        out.xyz(bname((pre + "molForceNew").c_str(), b), mp->molForceNew);
        out.xyz(bname((pre + "molForceRecNew").c_str(), b), mp->molForceRecNew);
        out.xyz(bname((pre + "molTorqueNew").c_str(), b), mp->molTorqueNew);
        // reject: restore the cell list (Accept's else branch, :537-539)
        sys.cellList.GridAll(sys.boxDimRef, sys.coordinates, sys.molLookupRef);
        ew.exgMolCache();
      }
    }

    // ---- MultiParticleBrownian move (same driving pattern) --------------------
    if (b == 0 && sys.moves[mv::MULTIPARTICLE_BM] != NULL) {
      MultiParticleBrownian *bm =
          static_cast<MultiParticleBrownian *>(sys.moves[mv::MULTIPARTICLE_BM]);
      const int nTypes = bm->allTranslate ? 1 : 2;
      for (int type = 0; type < nTypes; ++type) {
        const char *tn = type == mp::MPROTATE ? "bmRotate" : "bmDisplace";
        std::string pre = std::string(tn) + ".";
        ulong step = 5151 + 13 * type;
        sys.r123wrapper.SetStep(step);
        bm->bPick = b;
        bm->moveType = type;
        bm->SetMolInBox(b);
        if (ff.ewald) {
          ew.CopyRecip(b);
          ew.BoxForceReciprocal(sys.coordinates, sys.atomForceRecRef,
                                sys.molForceRecRef, b);
        }
        ce.BoxForce(sys.potential, sys.coordinates, sys.atomForceRef,
                    sys.molForceRef, sys.boxDimRef, b);
        ce.CalculateTorque(bm->moleculeIndex, sys.coordinates, sys.com,
                           sys.atomForceRef, sys.atomForceRecRef,
                           bm->molTorqueRef, b);
        sys.coordinates.CopyRange(bm->newMolsPos, 0, 0, sys.coordinates.Count());
        sys.com.CopyRange(bm->newCOMs, 0, 0, sys.com.Count());
        bm->CalculateTrialDistRot();
        double par[6] = {bm->moveSetRef.GetTMAX(b), bm->moveSetRef.GetRMAX(b),
                         bm->BETA, (double)step,
                         (double)sys.r123wrapper.GetSeedValue(),
                         (double)sys.r123wrapper.GetKeyValue()};
        out.f64(bname((pre + "params").c_str(), b), par, 6);
        out.xyz(bname((pre + "molTorqueRef").c_str(), b), bm->molTorqueRef);
        out.xyz(bname((pre + "k").c_str(), b), type == mp::MPROTATE ? bm->r_k : bm->t_k);
        out.xyz(bname((pre + "newMolsPos").c_str(), b), bm->newMolsPos);
        out.xyz(bname((pre + "newCOMs").c_str(), b), bm->newCOMs);
        // without a force-range test a badly overlapping start configuration can
        // throw a molecule further than one box length (WrapPBC wraps once); the
        // reference's cell list aborts on that, so the energy part is skipped then
        bool inside = true;
        XYZ ax2 = sys.boxDimRef.GetAxis(b);
        for (uint a = 0; a < bm->newMolsPos.Count() && inside; ++a) {
          XYZ u = bm->newMolsPos.Get(a);
          if (!sys.boxDimRef.orthogonal[b])
            u = static_cast<BoxDimensionsNonOrth &>(sys.boxDimRef).TransformUnSlant(u, b);
          inside = u.x >= 0 && u.x < ax2.x && u.y >= 0 && u.y < ax2.y && u.z >= 0 && u.z < ax2.z;
        }
        if (inside) {
          bm->CalcEn();
          out.f64(bname((pre + "wRatio").c_str(), b), bm->GetCoeff());
          out.xyz(bname((pre + "molForceNew").c_str(), b), bm->molForceNew);
          out.xyz(bname((pre + "molForceRecNew").c_str(), b), bm->molForceRecNew);
          out.xyz(bname((pre + "molTorqueNew").c_str(), b), bm->molTorqueNew);
          sys.cellList.GridAll(sys.boxDimRef, sys.coordinates, sys.molLookupRef);
          ew.exgMolCache();
        }
      }
    }

    // ---- MEMC / NeMTMC / free-energy reciprocal deltas (own RNG stream so
    // that the entries above keep their values) ------------------------------
    // (EwaldCached refuses these moves outright, src/EwaldCached.cpp:392-440)
    if (ff.ewald && molsInBox.size() > 3 && dynamic_cast<EwaldCached *>(&ew) == NULL) {
      std::mt19937_64 rng2(seed * 7919ULL + 17ULL * b + 1ULL);
      std::uniform_real_distribution<double> U2(-1.0, 1.0);
      uint pick[4];
      for (int i = 0; i < 4; ++i)
        pick[i] = molsInBox[(size_t)(rng2() % molsInBox.size())];
      // two exchange calls: (insert copy of pick0 at shifted place, remove
      // pick1) with first_call = true, then (pick2 in, pick3 out) on top
      std::vector<int> exMol;
      std::vector<double> exE;
      for (int call = 0; call < 2; ++call) {
        uint mN = pick[2 * call], mO = pick[2 * call + 1];
        uint lenN = mols.GetKind(mN).NumAtoms();
        cbmc::TrialMol tN(mols.GetKind(mN), sys.boxDimRef, b);
        cbmc::TrialMol tO(mols.GetKind(mO), sys.boxDimRef, b);
        XYZArray nc(lenN);
        XYZ shift(U2(rng2) * 4.0, U2(rng2) * 4.0, U2(rng2) * 4.0);
        for (uint a = 0; a < lenN; ++a)
          nc.Set(a, sys.boxDimRef.WrapPBC(
                        sys.coordinates.Get(mols.MolStart(mN) + a) + shift, b));
        tN.SetCoords(nc, 0);
        tO.SetCoords(sys.coordinates, mols.MolStart(mO));
        std::vector<cbmc::TrialMol> vN(1, tN), vO(1, tO);
        std::vector<uint> iN(1, mN), iO(1, mO);
        double d = ew.MolExchangeReciprocal(vN, vO, iN, iO, call == 0);
        exMol.push_back((int)mN);
        exMol.push_back((int)mO);
        exE.push_back(d);
        out.xyz(bname(call ? "exchange1.newCoords" : "exchange0.newCoords", b), nc);
      }
      out.i32(bname("exchange.mols", b), exMol);
      out.f64(bname("exchange.dRecip", b), exE);
      out.f64(bname("exchange.sumRnew", b), ew.sumRnew[b], ew.imageSizeRef[b]);
      out.f64(bname("exchange.sumInew", b), ew.sumInew[b], ew.imageSizeRef[b]);
      // ChangeLambdaRecip: molecule pick0 at its current coordinates
      {
        uint m = pick[0];
        uint len = mols.GetKind(m).NumAtoms();
        XYZArray mc(len);
        for (uint a = 0; a < len; ++a)
          mc.Set(a, sys.coordinates.Get(mols.MolStart(m) + a));
        double d = ew.ChangeLambdaRecip(mc, 0.3, 0.85, m, b);
        out.i32(bname("changeLambda.mol", b), (int)m);
        out.f64(bname("changeLambda.dRecip", b), d);
        out.f64(bname("changeLambda.sumRnew", b), ew.sumRnew[b], ew.imageSizeRef[b]);
        std::vector<double> lam = {0.0, 0.2, 0.5, 0.75, 1.0};
        std::vector<Energy> ediff(lam.size());
        Energy dUdL;
        ew.ChangeRecip(ediff.data(), dUdL, lam, 2, m, b);
        std::vector<double> er;
        for (auto &e : ediff) er.push_back(e.recip);
        out.f64(bname("changeRecip.lambda", b), lam);
        out.f64(bname("changeRecip.dRecip", b), er);
        out.f64(bname("changeRecip.dUdL", b), dUdL.recip);
        // ChangeSelf / ChangeCorrection on the same molecule and lambda states
        std::vector<Energy> ed2(lam.size());
        Energy dUdL2;
        ew.ChangeSelf(ed2.data(), dUdL2, lam, 2, m, b);
        ew.ChangeCorrection(ed2.data(), dUdL2, lam, 2, m, b);
        std::vector<double> es, ec;
        for (auto &e : ed2) {
          es.push_back(e.self);
          ec.push_back(e.correction);
        }
        out.f64(bname("changeSelf.dSelf", b), es);
        out.f64(bname("changeCorrection.dCorrection", b), ec);
        double dd[2] = {dUdL2.self, dUdL2.correction};
        out.f64(bname("changeSelfCorrection.dUdL", b), dd, 2);
      }
    }
  }
  return 0;
}

// Light initialisation: the statements of Simulation::Simulation
// (src/Simulation.cpp:19-40) and System::Init (src/System.cpp:106-160) except
// the two O(N*nk) start-up sweeps (BoxReciprocalSetup inside Ewald::Init and
// SystemTotal's virial), which a bounded timing run cannot afford.
struct LightSim {
  StaticVals *sv;
  System *sys;
  bool ewaldOn;
};

LightSim light_init(const char *conf) {
  static Setup set;
  set.Init(conf, NULL);
  ulong startStep = 0;
  StaticVals *sv = new StaticVals(set);
  static MultiSim const *msNull = NULL;
  System *sysp = new System(*sv, set, startStep, msNull);
  System &sys = *sysp;
  sv->Init(set, sys);
  // ---- System::Init, minus SystemTotal and the k-space start-up sweep ----
#ifdef VARIABLE_PARTICLE_NUMBER
  sys.molLookup.Init(sv->mol, set.pdb.atoms, sv->forcefield,
                     set.config.in.restart.restartFromCheckpoint);
#endif
  sys.moveSettings.Init(*sv, set.pdb.remarks, sys.molLookupRef.GetNumKind(),
                        set.config.in.restart.restartFromCheckpoint);
  sys.vel.Init(set.pdb.atoms, set.config.in);
  sys.xsc.Init(set.pdb, sys.vel, set.config.in, sys.molLookupRef, sv->mol);
  sys.boxDimensions->Init(set.config.in.restart, set.config.sys.volume,
                          set.pdb.cryst, sv->forcefield);
  sys.coordinates.InitFromPDB(set.pdb.atoms);
  sys.com.CalcCOM();
  sys.atomForceRef.Init(set.pdb.atoms.beta.size());
  sys.molForceRef.Init(sys.com.Count());
  sys.atomForceRecRef.Init(set.pdb.atoms.beta.size());
  sys.molForceRecRef.Init(sys.com.Count());
  sys.cellList.SetCutoff();
  sys.cellList.GridAll(sys.boxDimRef, sys.coordinates, sys.molLookupRef);
  bool ewaldOn = set.config.sys.elect.ewald;
  if (ewaldOn)
    sys.calcEwald = new Ewald(*sv, sys);
  else
    sys.calcEwald = new NoEwald(*sv, sys);
  sys.InitLambda();
  sys.calcEnergy.Init(sys);
  Ewald &ew = *sys.calcEwald;
  Molecules &mols = sv->mol;
  if (ewaldOn) {
    // Ewald::Init (src/Ewald.cpp:100-127) without BoxReciprocalSetup
    for (uint m = 0; m < mols.count; ++m) {
      const MoleculeKind &molKind = mols.GetKind(m);
      for (uint a = 0; a < molKind.NumAtoms(); ++a) {
        ew.particleKind.push_back(molKind.AtomKind(a));
        ew.particleMol.push_back(m);
        ew.particleCharge.push_back(molKind.AtomCharge(a));
        ew.particleHasNoCharge.push_back(std::abs(molKind.AtomCharge(a)) <
                                         0.000000001);
      }
    }
    ew.startMol.resize(sys.coordinates.Count());
    ew.lengthMol.resize(sys.coordinates.Count());
    for (int atom = 0; atom < (int)sys.coordinates.Count(); atom++) {
      ew.startMol[atom] = mols.MolStart(ew.particleMol[atom]);
      ew.lengthMol[atom] = mols.MolLength(ew.particleMol[atom]);
    }
    ew.AllocMem();
    for (uint b = 0; b < BOXES_WITH_U_NB; ++b) {
      ew.RecipInit(b, sys.boxDimRef);
#ifdef GOMC_CUDA
      // the GPU build uploads the k-vectors to the device inside
      // BoxReciprocalSetup (src/Ewald.cpp:199-228) and can afford the sweep
      ew.BoxReciprocalSetup(b, sys.coordinates);
#else
      std::memset(ew.sumRnew[b], 0, sizeof(double) * ew.imageSize[b]);
      std::memset(ew.sumInew[b], 0, sizeof(double) * ew.imageSize[b]);
#endif
      ew.SetRecipRef(b);
    }
  }
  LightSim ls = {sv, sysp, ewaldOn};
  return ls;
}

int run_time(int argc, char **argv) {
  const char *conf = argv[2];
  const char *outPath = argv[3];
  int kFrac = argc > 4 ? atoi(argv[4]) : 1;
  int reps = argc > 5 ? atoi(argv[5]) : 3;
  bool doForce = argc > 6 && std::string(argv[6]) == "force";
  LightSim ls = light_init(conf);
  StaticVals *sv = ls.sv;
  System &sys = *ls.sys;
  const bool ewaldOn = ls.ewaldOn;
  Ewald &ew = *sys.calcEwald;
  Dump out(outPath);
  out.i32("threads", omp_get_max_threads());
  dump_static(out, *sv, sys);
  const uint b = 0;
  uint nkFull = ewaldOn ? ew.imageSizeRef[b] : 0;
  uint nkSlab = ewaldOn ? std::max(1u, nkFull / (uint)std::max(1, kFrac)) : 0;
  if (ewaldOn) dump_kvectors(out, ew, b, false);
  out.i32("time.nkFull", (int)nkFull);
  out.i32("time.nkSlab", (int)nkSlab);
  std::vector<double> tInter, tSums, tRecip, tForce, tForceRec;
  double lj = 0, real = 0, recip = 0;
  for (int r = 0; r < reps + 1; ++r) { // first repetition is the warm-up
    double t0 = now();
    SystemPotential pot = sys.calcEnergy.BoxInter(
        SystemPotential(), sys.coordinates, sys.boxDimRef, b);
    double t1 = now();
    if (ewaldOn) {
      ew.imageSizeRef[b] = nkSlab; // bounded k-slab sample
      ew.BoxReciprocalSums(b, sys.coordinates);
    }
    double t2 = now();
    if (ewaldOn) recip = ew.BoxReciprocal(b, false);
    double t3 = now();
    if (ewaldOn) ew.imageSizeRef[b] = nkFull;
    lj = pot.boxEnergy[b].inter;
    real = pot.boxEnergy[b].real;
    if (r > 0) {
      tInter.push_back(t1 - t0);
      tSums.push_back(t2 - t1);
      tRecip.push_back(t3 - t2);
    }
    if (doForce) {
      double f0 = now();
      sys.calcEnergy.BoxForce(SystemPotential(), sys.coordinates,
                              sys.atomForceRef, sys.molForceRef, sys.boxDimRef,
                              b);
      double f1 = now();
      if (ewaldOn) {
#ifndef GOMC_CUDA
        ew.imageSizeRef[b] = std::max(1u, nkSlab / 8);
#endif
        // BoxForceReciprocal opens one parallel region per atom; sample the
        // first molecules only by shrinking nothing else -- cost is linear in
        // the slab size, reported as such.
        ew.BoxForceReciprocal(sys.coordinates, sys.atomForceRecRef,
                              sys.molForceRecRef, b);
        ew.imageSizeRef[b] = nkFull;
      }
      double f2 = now();
      if (r > 0) {
        tForce.push_back(f1 - f0);
        tForceRec.push_back(f2 - f1);
      }
    }
  }
  out.f64("time.BoxInter", tInter);
  out.f64("time.BoxReciprocalSums.slab", tSums);
  out.f64("time.BoxReciprocal.slab", tRecip);
  out.f64("time.BoxForce", tForce);
  out.f64("time.BoxForceReciprocal.slab8", tForceRec);
  out.f64("time.lj", lj);
  out.f64("time.real", real);
  out.f64("time.recipSlab", recip);
  if (ewaldOn) {
    out.f64("time.sumRnew.slab", ew.sumRnew[b], nkSlab);
    out.f64("time.sumInew.slab", ew.sumInew[b], nkSlab);
  }
  return 0;
}


// Full-size parity pin: BoxInter over the whole box and the structure factor on a chosen
// subset of the k list (the selected k-vectors are moved to the front of the Ref arrays and
// imageSizeRef shortened, so Ewald::BoxReciprocalSums itself produces the values).
//   gomc_probe_<ENS> slab <in.conf> <out.bin> <nSel> <seed>
int run_slab(int argc, char **argv) {
  const char *conf = argv[2];
  const char *outPath = argv[3];
  int nSel = argc > 4 ? atoi(argv[4]) : 256;
  unsigned seed = argc > 5 ? (unsigned)atoi(argv[5]) : 7u;
  LightSim ls = light_init(conf);
  StaticVals *sv = ls.sv;
  System &sys = *ls.sys;
  const bool ewaldOn = ls.ewaldOn;
  Ewald &ew = *sys.calcEwald;
  Dump out(outPath);
  out.i32("threads", omp_get_max_threads());
  const uint b = 0;
  const uint nAtoms = sys.coordinates.Count();
  out.i32("nAtoms", (int)nAtoms);
  XYZ ax = sys.boxDimRef.axis.Get(b);
  double a3[3] = {ax.x, ax.y, ax.z};
  out.f64("box0.axis", a3, 3);
  // inputs as parsed: checksums (and the coordinates themselves for small boxes)
  double cs[6] = {0, 0, 0, 0, 0, 0};
  for (uint i = 0; i < nAtoms; ++i) {
    cs[0] += sys.coordinates.x[i];
    cs[1] += sys.coordinates.y[i];
    cs[2] += sys.coordinates.z[i];
    cs[3] += sys.coordinates.x[i] * (double)(i % 97 + 1);
    cs[4] += sys.coordinates.y[i] * (double)(i % 89 + 1);
    cs[5] += sys.coordinates.z[i] * (double)(i % 83 + 1);
  }
  out.f64("coords.checksum", cs, 6);
  if (argc > 6 && std::string(argv[6]) == "coords") out.xyz("coords", sys.coordinates);
  SystemPotential pot =
      sys.calcEnergy.BoxInter(SystemPotential(), sys.coordinates, sys.boxDimRef, b);
  out.f64("BoxInter.inter", pot.boxEnergy[b].inter);
  out.f64("BoxInter.real", pot.boxEnergy[b].real);
  if (ewaldOn) {
    const uint nk = ew.imageSizeRef[b];
    out.i32("box0.nk", (int)nk);
    out.i32("box0.kmax", (int)ew.kmax[b]);
    // selection: evenly spaced through the list (every x slab, every column range) plus
    // random picks and both ends
    std::vector<int> sel;
    std::mt19937_64 rng(seed);
    nSel = std::min<int>(nSel, (int)nk);
    for (int i = 0; i < nSel / 2; ++i) sel.push_back((int)((uint64_t)i * nk / (nSel / 2)));
    sel.push_back((int)nk - 1);
    while ((int)sel.size() < nSel) sel.push_back((int)(rng() % nk));
    std::sort(sel.begin(), sel.end());
    sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
    nSel = (int)sel.size();
    std::vector<double> kx(nSel), ky(nSel), kz(nSel), pf(nSel);
    for (int i = 0; i < nSel; ++i) {
      kx[i] = ew.kxRef[b][sel[i]];
      ky[i] = ew.kyRef[b][sel[i]];
      kz[i] = ew.kzRef[b][sel[i]];
      pf[i] = ew.prefactRef[b][sel[i]];
    }
    for (int i = 0; i < nSel; ++i) {
      ew.kxRef[b][i] = kx[i];
      ew.kyRef[b][i] = ky[i];
      ew.kzRef[b][i] = kz[i];
      ew.prefactRef[b][i] = pf[i];
    }
    ew.imageSizeRef[b] = nSel;
    ew.BoxReciprocalSums(b, sys.coordinates);
    out.i32("slab.index", sel);
    out.f64("slab.kx", kx);
    out.f64("slab.ky", ky);
    out.f64("slab.kz", kz);
    out.f64("slab.prefact", pf);
    out.f64("slab.sumRnew", ew.sumRnew[b], nSel);
    out.f64("slab.sumInew", ew.sumInew[b], nSel);
  }
  (void)sv;
  return 0;
}

#if ENSEMBLE == NPT
// Volume trials of the NPT ensemble driven through the reference's own VolumeTransfer
// object (src/moves/VolumeTransfer.h): Prep and Transform with a chosen volume change,
// the real CalcEn (GridBox, RecipInit(newDim), BoxReciprocalSetup, BoxInter,
// BoxReciprocal(box, true)), then the accept branch (UpdateRecip + UpdateRecipVec, :255-260)
// for the first trial and the reject branch (:263-269) for the second.  After each, a
// single-molecule MolReciprocal / MoleculeInter shows which state the Ewald object is in.
//   gomc_probe_NPT volume <in.conf> <out.bin> <delta1> <delta2>
int run_volume(int argc, char **argv) {
  const char *conf = argv[2];
  const char *outPath = argv[3];
  const double deltas[2] = {argc > 4 ? atof(argv[4]) : 150.0, argc > 5 ? atof(argv[5]) : -220.0};
  Simulation sim(conf);
  System &sys = *sim.system;
  StaticVals &sv = *sim.staticValues;
  Molecules &mols = sv.mol;
  Ewald &ew = *sys.calcEwald;
  CalculateEnergy &ce = sys.calcEnergy;
  VolumeTransfer *vt = static_cast<VolumeTransfer *>(sys.moves[mv::VOL_TRANSFER]);
  Dump out(outPath);
  out.i32("threads", omp_get_max_threads());
  dump_static(out, sv, sys);
  const uint b = 0;
  dump_kvectors(out, ew, b, true);
  out.f64("box0.sysPotRef.recip", sys.potential.boxEnergy[b].recip);
  out.f64("box0.sysPotRef.inter", sys.potential.boxEnergy[b].inter);
  out.f64("box0.sysPotRef.real", sys.potential.boxEnergy[b].real);
  std::vector<uint> molsInBox;
  for (MoleculeLookup::box_iterator it = sys.molLookupRef.BoxBegin(b);
       it != sys.molLookupRef.BoxEnd(b); ++it)
    molsInBox.push_back(*it);
  // one fixed single-molecule displacement evaluated in whatever state is current
  auto probe_move = [&](const std::string &tag, uint m) {
    uint len = mols.GetKind(m).NumAtoms(), st = mols.MolStart(m);
    XYZArray newPos(len);
    XYZ ax = sys.boxDimRef.axis.Get(b);
    for (uint a = 0; a < len; ++a) {
      XYZ p = sys.coordinates.Get(st + a);
      p.x += 0.37; p.y -= 0.21; p.z += 0.13;
      newPos.Set(a, p);
    }
    sys.boxDimRef.WrapPBC(newPos, b);
    out.xyz(tag + ".newPos", newPos);
    out.i32(tag + ".mol", (int)m);
    double recip = ew.MolReciprocal(newPos, m, b);
    out.f64(tag + ".MolReciprocal", recip);
    sys.cellList.RemoveMol(m, b, sys.coordinates);
    Intermolecular iLJ, iReal;
    bool ov = ce.MoleculeInter(iLJ, iReal, newPos, m, b);
    sys.cellList.AddMol(m, b, sys.coordinates);
    out.f64(tag + ".dLJ", iLJ.energy);
    out.f64(tag + ".dReal", iReal.energy);
    out.i32(tag + ".overlap", (int)ov);
    out.f64(tag + ".sysPotRef.recip", sys.potential.boxEnergy[b].recip);
  };
  probe_move("state0", molsInBox[molsInBox.size() / 3]);
  for (int t = 0; t < 2; ++t) {
    const std::string tag = std::string("trial") + std::to_string(t);
    // ---- Prep (:72-97) with box 0, Transform (:115-133) with a chosen delta ----
    vt->box = b;
    vt->newDim = sys.boxDimRef;
    sys.coordinates.CopyRange(vt->newMolsPos, 0, 0, sys.coordinates.Count());
    sys.com.CopyRange(vt->newCOMs, 0, 0, sys.com.Count());
    XYZ scale;
    uint state = sys.boxDimRef.ShiftVolume(vt->newDim, scale, b, deltas[t]);
    if (state != mv::fail_state::NO_FAIL) return 3;
    sys.coordinates.TranslateOneBox(vt->newMolsPos, vt->newCOMs, sys.com, vt->newDim, b, scale);
    XYZ nax = vt->newDim.axis.Get(b);
    double a3[3] = {nax.x, nax.y, nax.z};
    out.f64(tag + ".delta", deltas[t]);
    out.f64(tag + ".newAxis", a3, 3);
    out.f64(tag + ".newVolume", vt->newDim.volume[b]);
    out.xyz(tag + ".newCoords", vt->newMolsPos);
    out.xyz(tag + ".newCOM", vt->newCOMs);
    // ---- the reference's CalcEn (:139-198) ----
    vt->CalcEn();
    const Energy &en = vt->sysPotNew.boxEnergy[b];
    out.f64(tag + ".inter", en.inter);
    out.f64(tag + ".real", en.real);
    out.f64(tag + ".recip", en.recip);
    out.f64(tag + ".tailCorrection", en.tailCorrection);
    out.f64(tag + ".total", vt->sysPotNew.Total());
    out.f64(tag + ".coeff", vt->GetCoeff());
    uint nkNew = ew.imageSize[b];
    out.i32(tag + ".nk", (int)nkNew);
    out.f64(tag + ".kx", ew.kx[b], nkNew);
    out.f64(tag + ".ky", ew.ky[b], nkNew);
    out.f64(tag + ".kz", ew.kz[b], nkNew);
    out.f64(tag + ".hsqr", ew.hsqr[b], nkNew);
    out.f64(tag + ".prefact", ew.prefact[b], nkNew);
    out.f64(tag + ".sumRnew", ew.sumRnew[b], nkNew);
    out.f64(tag + ".sumInew", ew.sumInew[b], nkNew);
    if (t == 0) {
      // ---- accept branch of VolumeTransfer::Accept (:243-260) ----
      sys.potential = vt->sysPotNew;
      swap(sys.coordinates, vt->newMolsPos);
      swap(sys.com, vt->newCOMs);
      sys.boxDimRef = vt->newDim;
      ew.UpdateRecip(b);
      ew.UpdateRecipVec(b);
      out.i32(tag + ".accepted", 1);
      out.f64(tag + ".after.BoxReciprocal", ew.BoxReciprocal(b, false));
    } else {
      // ---- reject branch (:263-269) ----
      sys.cellList.GridBox(sys.boxDimRef, sys.coordinates, sys.molLookupRef, b);
      ew.exgMolCache();
      out.i32(tag + ".accepted", 0);
    }
    probe_move(tag + ".after", molsInBox[(molsInBox.size() * (t + 2)) / 5]);
  }
  return 0;
}
#endif

} // namespace

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr,
            "usage: %s golden <in.conf> <out.bin> [nMoves] [seed]\n"
            "       %s time   <in.conf> <out.bin> <kFraction> <reps> [force]\n",
            argv[0], argv[0]);
    return 1;
  }
  std::string mode = argv[1];
  if (mode == "golden") return run_golden(argc, argv);
  if (mode == "time") return run_time(argc, argv);
  if (mode == "slab") return run_slab(argc, argv);
#if ENSEMBLE == NPT
  if (mode == "volume") return run_volume(argc, argv);
#endif
  fprintf(stderr, "unknown mode %s\n", argv[1]);
  return 1;
}
