// multi_gpu_test.cpp -- ONE host process, N GPUs, through the C ABI only.
//
// GOMC is a single process (src/Main.cpp:318-326 picks one device).  This driver shows how
// such a process reaches every GPU of a box: one engine per device, one host thread per
// engine, gomcb200_set_comm() with a unique id made by thread 0.  Every thread then makes
// the very same calls GOMC's System::Init makes (RecipInit, BoxReciprocalSetup ... here the
// fused gomcb200_call_full_box_energy) and every thread gets the COMPLETE energies back: the
// cell slabs and FFT slabs are sharded inside, the exchange is NCCL on the engines' streams.
// Prints one JSON object: the single-GPU energies and the energies every rank returned.
//
// usage: multi_gpu_test <system file of tests/test_host_mirror_gpu.py> <nDevices>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
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

struct SystemData {
  int nAtoms, nMols, K, vdwKind, ewaldOn;
  std::vector<double> d, sig, eps, nn, x, y, z, q;
  std::vector<int> kind, mol, molStart;
};

static void evaluate(const SystemData &s, int device, const void *id, int rank, int world,
                     double out[3]) {
  EngineB200 eng(device, 1);
  double rcc[1] = {s.d[1]}, alpha[1] = {s.d[4]}, rr[1] = {s.d[5]};
  eng.InitForceField(s.sig.data(), s.eps.data(), s.nn.data(), s.vdwKind, 0, s.K, s.d[0], rcc,
                     s.d[2], s.d[3], alpha, s.ewaldOn, true);
  eng.InitTopology(s.kind, s.mol, s.q, s.molStart);
  std::vector<int> all(s.nMols);
  for (int m = 0; m < s.nMols; ++m) all[m] = m;
  eng.SetBoxMolecules(0, all);
  XYZ axis = {s.d[6], s.d[7], s.d[8]};
  eng.SetBoxAxes(0, axis);
  if (world > 1 && gomcb200_set_comm(eng.get(), id, rank, world)) {
    fprintf(stderr, "set_comm: %s\n", gomcb200_last_error());
    exit(3);
  }
  if (s.ewaldOn) {
    Ewald ew(eng, alpha, rr, 1);
    ew.AllocMem({axis}, 1.0);
    ew.RecipInit(0, axis);
    // BoxReciprocalSetup + SetRecipRef, then the evaluation on the reference k set
    XYZView coords = {s.x.data(), s.y.data(), s.z.data(), s.nAtoms};
    ew.BoxReciprocalSetup(0, coords);
    ew.SetRecipRef(0);
    if (gomcb200_call_full_box_energy(eng.get(), 0, s.x.data(), s.y.data(), s.z.data(), &out[0],
                                      &out[1], &out[2])) {
      fprintf(stderr, "full_box_energy: %s\n", gomcb200_last_error());
      exit(3);
    }
  } else if (gomcb200_call_full_box_energy(eng.get(), 0, s.x.data(), s.y.data(), s.z.data(),
                                           &out[0], &out[1], &out[2])) {
    fprintf(stderr, "full_box_energy: %s\n", gomcb200_last_error());
    exit(3);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) return 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  const int world = atoi(argv[2]);
  SystemData s;
  auto h = rd<int>(f, 6);
  s.nAtoms = h[0], s.nMols = h[1], s.K = h[2], s.vdwKind = h[3], s.ewaldOn = h[4];
  s.d = rd<double>(f, 9);
  s.sig = rd<double>(f, s.K * s.K), s.eps = rd<double>(f, s.K * s.K), s.nn = rd<double>(f, s.K * s.K);
  s.x = rd<double>(f, s.nAtoms), s.y = rd<double>(f, s.nAtoms), s.z = rd<double>(f, s.nAtoms);
  s.q = rd<double>(f, s.nAtoms);
  rd<double>(f, 3 * (size_t)s.nMols);  // centres of mass: not needed here
  s.kind = rd<int>(f, s.nAtoms), s.mol = rd<int>(f, s.nAtoms), s.molStart = rd<int>(f, s.nMols + 1);
  fclose(f);

  double single[3];
  evaluate(s, 0, nullptr, 0, 1, single);

  char id[128];
  if (gomcb200_comm_unique_id(id)) {
    fprintf(stderr, "unique id: %s\n", gomcb200_last_error());
    return 3;
  }
  std::vector<double> res(3 * (size_t)world);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r)
    th.emplace_back([&, r] { evaluate(s, r, id, r, world, &res[3 * (size_t)r]); });
  for (auto &t : th) t.join();

  printf("{\"world\": %d, \"single\": [%.17g, %.17g, %.17g], \"ranks\": [", world, single[0],
         single[1], single[2]);
  for (int r = 0; r < world; ++r)
    printf("%s[%.17g, %.17g, %.17g]", r ? ", " : "", res[3 * r], res[3 * r + 1], res[3 * r + 2]);
  printf("]}\n");
  return 0;
}
