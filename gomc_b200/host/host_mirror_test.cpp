// host_mirror_test.cpp -- drives the C++ host mirror (GomcB200.h) the way GOMC's
// System::Init and Translate::CalcEn/Accept drive the reference classes, on a
// system read from a flat binary file, and prints the results as JSON for
// tests/test_host_mirror_gpu.py to compare with the oracle.
//
// input file (little endian): int32 header {nAtoms, nMols, kindCount, vdwKind, ewald, nMoves}
//   then doubles {rCut, rCutCoulomb, rCutLow, rOn, alpha, recip_rcut, axis[3]},
//   tables sigmaSq/epsilon_cn/n [kindCount^2], x,y,z,charge [nAtoms], comx,comy,comz [nMols],
//   int32 kind, mol [nAtoms], molStart [nMols+1], per move: int32 mol, then 3*len doubles.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "GomcB200.h"

using namespace gomc_b200;

template <typename T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  return v;
}

int main(int argc, char **argv) {
  if (argc < 2) return 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  auto h = rd<int>(f, 6);
  const int nAtoms = h[0], nMols = h[1], K = h[2], vdwKind = h[3], ewaldOn = h[4], nMoves = h[5];
  auto d = rd<double>(f, 9);
  auto sig = rd<double>(f, K * K), eps = rd<double>(f, K * K), nn = rd<double>(f, K * K);
  auto x = rd<double>(f, nAtoms), y = rd<double>(f, nAtoms), z = rd<double>(f, nAtoms),
       q = rd<double>(f, nAtoms);
  auto cx = rd<double>(f, nMols), cy = rd<double>(f, nMols), cz = rd<double>(f, nMols);
  auto kind = rd<int>(f, nAtoms), mol = rd<int>(f, nAtoms), molStart = rd<int>(f, nMols + 1);

  EngineB200 eng(0, 1);
  double rcc[1] = {d[1]}, alpha[1] = {d[4]}, rr[1] = {d[5]};
  eng.InitForceField(sig.data(), eps.data(), nn.data(), vdwKind, 0, K, d[0], rcc, d[2], d[3], alpha,
                     ewaldOn, true);
  eng.InitTopology(kind, mol, q, molStart);
  std::vector<int> all(nMols);
  for (int m = 0; m < nMols; ++m) all[m] = m;
  eng.SetBoxMolecules(0, all);
  XYZ axis = {d[6], d[7], d[8]};
  eng.SetBoxAxes(0, axis);
  XYZView coords = {x.data(), y.data(), z.data(), nAtoms};
  XYZView com = {cx.data(), cy.data(), cz.data(), nMols};
  eng.SetCOM(com);

  CalculateEnergy calcEnergy(eng);
  Ewald *calcEwald = ewaldOn ? static_cast<Ewald *>(new EwaldCached(eng, alpha, rr, 1))
                             : static_cast<Ewald *>(new NoEwald(eng, alpha, rr, 1));
  // System::Init order (src/System.cpp:106-160): Ewald::Init then SystemTotal
  calcEwald->AllocMem({axis}, 1.0);
  calcEwald->RecipInit(0, axis);
  calcEwald->BoxReciprocalSetup(0, coords);
  calcEwald->SetRecipRef(0);
  Energy pot = calcEnergy.BoxInter(coords, axis, 0);
  pot.recip = calcEwald->BoxReciprocal(0, false);
  calcEwald->BoxSelfAndCorrection(0, pot.self, pot.correction);
  calcEwald->SetSysPotRecip(0, pot.recip);
  printf("{\"inter\": %.17g, \"real\": %.17g, \"recip\": %.17g, \"self\": %.17g, "
         "\"correction\": %.17g, \"moves\": [",
         pot.inter, pot.real, pot.recip, pot.self, pot.correction);
  // Translate::CalcEn / Accept (src/moves/Translate.h:82-113), every move accepted
  for (int t = 0; t < nMoves; ++t) {
    int m = rd<int>(f, 1)[0];
    int len = molStart[m + 1] - molStart[m];
    auto nx = rd<double>(f, len), ny = rd<double>(f, len), nz = rd<double>(f, len);
    XYZView mc = {nx.data(), ny.data(), nz.data(), len};
    Intermolecular iLJ, iReal;
    bool overlap = calcEnergy.MoleculeInter(iLJ, iReal, mc, m, 0);
    double dRecip = overlap ? 0.0 : calcEwald->MolReciprocal(mc, m, 0);
    double swapCorr = calcEwald->SwapCorrection(mc, m, 0, axis);
    printf("%s{\"mol\": %d, \"dLJ\": %.17g, \"dReal\": %.17g, \"dRecip\": %.17g, "
           "\"overlap\": %d, \"swapCorr\": %.17g, \"swapSelf\": %.17g}",
           t ? ", " : "", m, iLJ.energy, iReal.energy, dRecip, (int)overlap, swapCorr,
           calcEwald->SwapSelf(m, 0));
    if (!overlap) {  // accept
      XYZ c = {nx[0], ny[0], nz[0]};
      eng.AcceptMolecule(m, mc, c);
      calcEwald->UpdateRecip(0);
      pot.inter += iLJ.energy;
      pot.real += iReal.energy;
      pot.recip += dRecip;
      calcEwald->SetSysPotRecip(0, pot.recip);
      for (int a = 0; a < len; ++a) {
        x[molStart[m] + a] = nx[a];
        y[molStart[m] + a] = ny[a];
        z[molStart[m] + a] = nz[a];
      }
    } else {
      calcEwald->RestoreMol(m);
    }
  }
  // Simulation::RecalculateAndCheck (src/Simulation.cpp:170-223): running sums
  // versus a recomputation from scratch
  Energy chk = calcEnergy.BoxInter(coords, axis, 0);
  calcEwald->BoxReciprocalSums(0, coords);
  chk.recip = calcEwald->BoxReciprocal(0, false);
  printf("], \"running\": {\"inter\": %.17g, \"real\": %.17g, \"recip\": %.17g}, "
         "\"recomputed\": {\"inter\": %.17g, \"real\": %.17g, \"recip\": %.17g}",
         pot.inter, pot.real, pot.recip, chk.inter, chk.real, chk.recip);
  // one MultiParticle displacement step (src/moves/MultiParticle.h: Prep, Transform,
  // CalcEn, GetCoeff, Accept(false)) through the mirror class
  {
    calcEwald->SetRecipRef(0);
    MultiParticle mp(eng, calcEnergy, *calcEwald, 1.0 / 300.0);
    mp.Prep(0, MultiParticle::MPDISPLACE);
    mp.Transform(0.02, 0.03, 77, 0, 123);
    Energy en = mp.CalcEn();
    double w = mp.GetCoeff();
    mp.Accept(false);
    Energy back = calcEnergy.BoxInter(coords, axis, 0);
    printf(", \"mp\": {\"inter\": %.17g, \"real\": %.17g, \"recip\": %.17g, \"w\": %.17g, "
           "\"inter_after_reject\": %.17g}",
           en.inter, en.real, en.recip, w, back.inter);
  }
  printf("}\n");
  delete calcEwald;
  fclose(f);
  return 0;
}
