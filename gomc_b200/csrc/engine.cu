// engine.cu -- device-resident state + the C ABI of include/gomc_b200.h.
//
// Host side of the B200 engine: owns device memory, bins atoms into cells,
// plans the reciprocal-space tiling, launches the kernels of pair.cuh and
// recip.cuh on one stream and returns scalars through pinned memory.
// No CPU fallback: every entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "../../include/gomc_b200.h"
#include "common.cuh"
#include "pair.cuh"
#include "recip.cuh"
#include "trial.cuh"
#include "mp.cuh"
#include "recip_mma.cuh"
#include "recip_i8.cuh"
#include "force_mma.cuh"
#include "pair2.cuh"
#include "nufft.h"
#include "comm.h"

// NVTX ranges under the reference's own profiling event names (src/GOMCEventsProfileDef.h:
// 105-132), so that an Nsight timeline of GOMC on this engine reads like one of GOMC built
// with GOMC_NVTX_ENABLED.  nvtx3 is header-only; without a profiler attached a range costs a
// few nanoseconds.
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define GB_RANGE(name) NvtxRange gbNvtxRange_(name)

using namespace gb;

namespace {

thread_local std::string g_lastError;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_lastError = buf;
  return code;
}

#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess)                                                    \
      return fail(GOMCB200_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,  \
                  cudaGetErrorString(_e));                                    \
  } while (0)

// Owning device buffer: freed when the engine (or the KSet / BoxState holding it) goes away.
// bumped by every (re)allocation or release of a DevBuf: invalidates captured CUDA graphs
std::atomic<long long> g_devBufGeneration{0};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap) {
    o.p = nullptr;
    o.cap = 0;
  }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      cap = o.cap;
      o.p = nullptr;
      o.cap = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    ++g_devBufGeneration;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) ++g_devBufGeneration;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct KSet {
  DevBuf<double> kx, ky, kz, hsqr, prefact;
  std::vector<double> hkx, hky, hkz, hhsqr, hprefact;
  int n = 0, kmax = 0;
  int nmax[3] = {0, 0, 0};
  double cv[3] = {0, 0, 0};
  double L[3] = {0, 0, 0};  // box edges this k set was built for (non-uniform FFT grid mapping)
  // factorised-sum plan
  DevBuf<int4> rows, tiles;
  int nTiles = 0, maxRows = 0, nRowsPadded = 0;
  bool planValid = false;
  // DMMA plan (recip_mma.cuh)
  DevBuf<int4> mmaRows, mmaTiles, mmaSegs;
  DevBuf<int> mmaCtaSeg;
  std::vector<int4> hMmaTiles, hRowsSorted, hSegs;
  // int8 tensor-core path (recip_i8.cuh): tiles of 64 rows x 32 c values
  DevBuf<int4> i8Tiles, i8Rows;
  int i8NTiles = 0, i8NTilesAll = 0, i8NCB = 0, i8NSL = 6, i8CPer = 32, i8Nb = 64;
  // DMMA reciprocal-force plan (force_mma.cuh)
  DevBuf<int4> fmRows;
  DevBuf<FmTile> fmTiles;
  DevBuf<FmBlock> fmBlocks;
  int fmNBlocks = 0;
  DevBuf<double> fmW;
  int fmNTiles = 0;
  bool fmValid = false;
  std::vector<int> hCtaSeg;
  int tilesForShard = -1;
  int mmaZS = 0;
  bool mmaValid = false;
  int itemsForAtoms = -1, itemsForShard = -1, nCtas = 0, maxSlabs = 0, itemsAT = 0;
  // non-uniform FFT path (nufft.cu): fine grid + window of this k set
  gbn::NufftGrid ng;
  bool ngValid = false;
};

struct BoxState {
  double axis[3] = {0, 0, 0};
  bool haveAxes = false;
  bool nonOrth = false;
  double B[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Bi[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::vector<int> hMols, hAtoms, hCharged;
  double qMaxAbs = 0.0;  // largest |charge| in the box (fixed-point scale of the int8 kernel)
  DevBuf<int> molList, atomList, chargedList;
  int nMols = 0, nAtoms = 0, nCharged = 0;
  // cell-sorted copy
  bool cellsDirty = true;
  CellGrid grid;
  DevBuf<int> keys, keysSorted, vals, sortedAtoms, cellStart;
  DevBuf<int> sortedPos;  // global atom index -> position in the cell-sorted copy
  DevBuf<double> sx, sy, sz, sq;
  DevBuf<int2> skm;
  DevBuf<int> maxCellPop;  // largest cell population of the current binning (device)
  bool sumsComplete = true;  // sharded engine: every rank holds all of sumRnew/sumInew
  // erfc(alpha r)/r and Coulomb virial factor as piecewise polynomials in r^2 (pair2.cuh)
  DevBuf<double> coulTab;
  double tabAlpha = -1.0, tabRc2 = -1.0;
  int tabN = 0, tabHi0 = 0;
  // reciprocal space
  KSet kset[2];  // index with cur / 1-cur
  int cur = 0;   // kset[cur] = "new" k set (kx[]), kset[1-cur] = Ref
  DevBuf<double> sum[4];  // Rnew, Inew, Rref, Iref through the idx below
  int iRnew = 0, iInew = 1, iRref = 2, iIref = 3;
  DevBuf<double4> packed;
  bool packedDirty = true;
  // fractional molecule of this box (lib/Lambda.h)
  int lambdaMol = -1, lambdaMolKind = -1;
  double lambdaVDW = 1.0, lambdaCoulomb = 1.0;
};

}  // namespace

struct gomcb200_engine {
  int device = 0, numSMs = 148, nBoxes = 1;
  size_t smemOptin = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  // force field
  bool haveFF = false;
  int vdwKind = 0, ewald = 0, electrostatic = 0, kindCount = 0;
  double rCut = 0, rCutLow = 0, rOn = 0;
  std::vector<double> rCutCoulomb, alpha, recipRcut;
  DevBuf<double> sigmaSq, epsilon_cn, nTab, shiftConst;
  DevBuf<double> rMin, expConst, rMaxSq, mAn, mBn, mCn, mSign, mSig6;
  double mA6 = 0, mB6 = 0, mC6 = 0, mA1 = 0, mB1 = 0, mC1 = 0, diElectric_1 = 1.0;
  bool haveExp6 = false;
  DevBuf<int> nHalf;
  // topology
  bool haveTopo = false;
  int nAtoms = 0, nMols = 0, maxMolLen = 0;
  std::vector<int> hKind, hMol, hMolStart;
  std::vector<int> hMolBox;  // box of each molecule (-1: in no box)
  std::vector<double> hCharge;
  std::vector<double> hChargeEff;  // charge * sqrt(lambdaCoulomb) of its molecule (Ewald terms)
  DevBuf<double> qEff;
  double scAlpha = 0.0, scSigma6 = 0.0, scPower = 0.0;
  int scCoul = 0;
  // lazily maintained host mirror of the coordinates (single-molecule moves read
  // the old positions from it instead of a D2H round trip)
  std::vector<double> hx, hy, hz;
  bool mirrorValid = false;
  DevBuf<int> kind, mol, molStart;
  DevBuf<double> x, y, z, q, comx, comy, comz;
  DevBuf<double> force[5][3];
  DevBuf<double> scratchF[3];  // k-space forces for VirialReciprocal
  // MultiParticle move: the other coordinate / COM / force set (trial while the
  // reference one is active and vice versa), t_k or r_k, in-range flags
  DevBuf<double> xT, yT, zT, comxT, comyT, comzT, forceT[5][3], mpK[3];
  DevBuf<int> mpInRange;
  DevBuf<signed char> mpInvolved;
  bool trialActive = false;
  std::vector<BoxState> box;
  int imageTotal = 0;
  // 0 direct, 1 factorised SIMT, 2 factorised DMMA, 3 int8 tensor cores, 5 = non-uniform FFT,
  // 4 = automatic: the non-uniform FFT for orthogonal boxes, the direct kernels otherwise
  int recipAlgo = 4;
  int pairAlgo = 1;  // 1: k_pair_box2 (pair2.cuh) for orthogonal boxes, 0: k_pair_box
  gbc::Comm *comm = nullptr;  // NCCL communicator of a sharded engine (gomcb200_set_comm)
  DevBuf<double> energy3;     // LJ, real, recip of a full-box evaluation (all-reduced in place)
  gbn::Nufft *nufft = nullptr;
  double recipAutoWork = 1e11;
  int shardRank = 0, shardWorld = 1;
  // scratch
  DevBuf<double> part, blockA, blockB, result, molBuf, probeOut;
  DevBuf<double2> phaseTables;
  DevBuf<unsigned char> zPlanes;  // int8 path: byte planes of the Z phases
  DevBuf<Probe> probes;
  DevBuf<unsigned char> cubTemp;
  double *hRes = nullptr;       // pinned, 64 doubles
  // fused single-molecule trial: mapped pinned {6 results, flag}, ticket, partials
  double *hTrial = nullptr, *dTrial = nullptr;
  unsigned long long trialSeq = 0;
  DevBuf<unsigned> ticket;
  DevBuf<double> trialPart;
  double *hStage = nullptr;     // pinned staging for small uploads
  size_t hStageCap = 0;
  // timing
  bool timing = false;
  cudaEvent_t ev[8];
  float lastTotalMs = 0, lastDominantMs = 0;
  // second stream of the full-box evaluation: the structure factor (and, sharded, its
  // all-gather) runs next to the cell binning and the pair sweep; scratch of its own
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  DevBuf<double> blockA2, result2;
  bool overlap = true;
  // the whole full-box evaluation as a CUDA graph (both streams, the collectives included):
  // captured once the state it bakes in has been identical for a few calls, replayed while
  // it stays so (StepKey), dropped and re-captured otherwise
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<unsigned char> key;
    long long launches = 0;
    int warm = 0;
    bool sumsComplete = true;
  } stepGraph[2];
  bool useGraph = true, capturing = false;
};

namespace {

BoxParams make_params(const gomcb200_engine *e, int b) {
  const BoxState &bx = e->box[b];
  BoxParams p;
  std::memset(&p, 0, sizeof(p));  // padding included: the struct is compared bytewise (step_key)
  for (int d = 0; d < 3; ++d) {
    p.ax[d] = bx.axis[d];
    p.half[d] = bx.axis[d] * 0.5;  // BoxDimensions halfAx
  }
  p.rCut = e->rCut;
  p.rCutSq = e->rCut * e->rCut;
  p.rCutLowSq = e->rCutLow * e->rCutLow;
  double rcc = e->rCutCoulomb[b];
  p.rCutCoulombSq = rcc * rcc;
  double br = std::max(e->rCut, rcc);  // src/BoxDimensions.cpp:17-18
  p.boxRcutSq = br * br;
  p.alpha = e->alpha[b];
  p.alphaSq = p.alpha * p.alpha;
  p.rOnSq = e->rOn * e->rOn;
  p.factor1 = p.rCutSq - 3 * p.rOnSq;  // src/FFSwitch.h:103-105
  double d3 = (p.rCutSq - p.rOnSq);
  p.factor2 = 1.0 / (d3 * d3 * d3);
  p.kindCount = e->kindCount;
  p.vdwKind = e->vdwKind;
  p.ewald = e->ewald;
  p.electrostatic = e->electrostatic;
  p.sigmaSq = e->sigmaSq.p;
  p.epsilon_cn = e->epsilon_cn.p;
  p.n = e->nTab.p;
  p.shiftConst = e->shiftConst.p;
  p.nHalf = e->nHalf.p;
  p.rMin = e->rMin.p;
  p.expConst = e->expConst.p;
  p.rMaxSq = e->rMaxSq.p;
  p.mAn = e->mAn.p;
  p.mBn = e->mBn.p;
  p.mCn = e->mCn.p;
  p.mSign = e->mSign.p;
  p.mSig6 = e->mSig6.p;
  p.rOn = e->rOn;
  p.A6 = e->mA6; p.B6 = e->mB6; p.C6 = e->mC6;
  p.A1 = e->mA1; p.B1 = e->mB1; p.C1 = e->mC1;
  p.diElectric_1 = e->diElectric_1;
  p.nonOrth = bx.nonOrth;
  for (int i = 0; i < 9; ++i) {
    p.B[i] = bx.B[i];
    p.Bi[i] = bx.Bi[i];
  }
  p.comx = e->comx.p;
  p.comy = e->comy.p;
  p.comz = e->comz.p;
  p.lambdaMol = bx.lambdaMol;
  p.lambdaVDW = bx.lambdaVDW;
  p.lambdaCoulomb = bx.lambdaCoulomb;
  p.scAlpha = e->scAlpha;
  p.scSigma6 = e->scSigma6;
  p.scPower = e->scPower;
  p.scCoul = e->scCoul;
  return p;
}

int check_box(const gomcb200_engine *e, int b, bool needAxes = true, bool needTopo = true) {
  if (!e) return fail(GOMCB200_EINVAL, "null engine");
  if (b < 0 || b >= e->nBoxes) return fail(GOMCB200_EINVAL, "box %d out of range", b);
  if (!e->haveFF) return fail(GOMCB200_EINVAL, "gomcb200_init_forcefield not called");
  if (needTopo && !e->haveTopo) return fail(GOMCB200_EINVAL, "gomcb200_init_topology not called");
  if (e->vdwKind == VDW_EXP6 && !e->haveExp6)
    return fail(GOMCB200_EINVAL, "Potential EXP6: gomcb200_init_exp6 not called");
  if (needAxes && !e->box[b].haveAxes)
    return fail(GOMCB200_EINVAL, "gomcb200_set_box_axes not called for box %d", b);
  return 0;
}

// Entry points that read or update the whole structure factor, or the forces of every atom,
// are only valid on an unsharded engine: with gomcb200_set_shard(world > 1) the sums of
// k-vectors owned by other ranks are zero and forces exist only for this rank's cell slab.
// withComm: the entry point is also valid on a sharded engine that owns a communicator
// (gomcb200_set_comm): there the forces are all-reduced and every rank holds complete sums.
int check_unsharded(const gomcb200_engine *e, const char *what, bool withComm = false) {
  if (e && e->shardWorld > 1 && !(withComm && e->comm))
    return fail(GOMCB200_EINVAL,
                "%s is not available on a sharded engine (gomcb200_set_shard world %d): "
                "only the full-box sweeps are sharded",
                what, e->shardWorld);
  return 0;
}

int stage_reserve(gomcb200_engine *e, size_t bytes) {
  if (bytes <= e->hStageCap) return 0;
  if (e->hStage) cudaFreeHost(e->hStage);
  e->hStage = nullptr;
  e->hStageCap = 0;
  CK(cudaHostAlloc(&e->hStage, bytes * 2, cudaHostAllocDefault));
  e->hStageCap = bytes * 2;
  return 0;
}

// ---- cell binning ---------------------------------------------------------
int ensure_cells(gomcb200_engine *e, int b) {
  BoxState &bx = e->box[b];
  if (!bx.cellsDirty) return 0;
  const int n = bx.nAtoms;
  double br = std::max(e->rCut, e->rCutCoulomb[b]);
  CellGrid g;
  for (int d = 0; d < 3; ++d) {  // CellList::ResizeGrid, src/CellList.cpp:138-163
    int ed = (int)std::floor(bx.axis[d] / br);
    g.edge[d] = std::max(ed, 3);
    g.cellSize[d] = bx.axis[d] / g.edge[d];
    g.generic[d] = g.edge[d] < 4 || bx.nonOrth;
  }
  g.nonOrth = bx.nonOrth;
  for (int i = 0; i < 9; ++i) g.Bi[i] = bx.Bi[i];
  g.nCells = g.edge[0] * g.edge[1] * g.edge[2];
  bx.grid = g;
  CK(bx.keys.reserve(n + 1));
  CK(bx.keysSorted.reserve(n + 1));
  CK(bx.vals.reserve(n + 1));
  CK(bx.sortedAtoms.reserve(n + 1));
  CK(bx.cellStart.reserve(g.nCells + 2));
  CK(bx.sx.reserve(n + 2));
  CK(bx.sy.reserve(n + 2));
  CK(bx.sz.reserve(n + 2));
  CK(bx.sq.reserve(n + 2));
  CK(bx.skm.reserve(n + 2));
  CK(bx.sortedPos.reserve(e->nAtoms + 1));
  CK(bx.maxCellPop.reserve(4));
  CK(cudaMemsetAsync(bx.maxCellPop.p, 0, sizeof(int), e->stream));
  if (n > 0) {
    int blocks = (n + 255) / 256;
    k_cell_keys<<<blocks, 256, 0, e->stream>>>(g, n, bx.atomList.p, e->x.p, e->y.p,
                                              e->z.p, bx.keys.p, bx.vals.p);
    int bits = 1;
    while ((1 << bits) < g.nCells + 1) ++bits;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, bx.keys.p, bx.keysSorted.p,
                                    bx.vals.p, bx.sortedAtoms.p, n, 0, bits,
                                    e->stream);
    CK(e->cubTemp.reserve(tmp + 16));
    CK(cub::DeviceRadixSort::SortPairs(e->cubTemp.p, tmp, bx.keys.p, bx.keysSorted.p,
                                       bx.vals.p, bx.sortedAtoms.p, n, 0, bits,
                                       e->stream));
    k_gather_sorted<<<blocks, 256, 0, e->stream>>>(n, bx.sortedAtoms.p, e->x.p, e->y.p,
                                                  e->z.p, e->q.p, e->kind.p, e->mol.p,
                                                  bx.sx.p, bx.sy.p, bx.sz.p, bx.sq.p,
                                                  bx.skm.p, bx.sortedPos.p);
    e->launches += 4;
  }
  k_cell_bounds<<<(n + 1 + 255) / 256, 256, 0, e->stream>>>(
      g.nCells, n, bx.keysSorted.p, bx.cellStart.p, bx.maxCellPop.p);
  e->launches += 1;
  CK(cudaGetLastError());
  bx.cellsDirty = false;
  return 0;
}

int fetch_result(gomcb200_engine *e, int n) {
  CK(cudaMemcpyAsync(e->hRes, e->result.p, sizeof(double) * n, cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- Coulomb real-space table of k_pair_box2 (pair2.cuh) -------------------
// f(s) = erfc(alpha sqrt(s)) / sqrt(s) and g(s) = (f(s) + 2 alpha/sqrt(pi) exp(-alpha^2 s)) / s
// (FFParticle::CalcCoulomb / CalcCoulombVir with Ewald on, src/FFParticle.cpp:400-446) on
// s = r^2 in [2^eMin, 2^eMax): 32 intervals per octave, degree-7 interpolant at the Chebyshev
// nodes of each interval, computed in long double and stored as monomial coefficients in
// u in [-1, 1], coefficient-major.
int ensure_coul_table(gomcb200_engine *e, int b) {
  BoxState &bx = e->box[b];
  const double alpha = e->alpha[b], rc = e->rCutCoulomb[b];
  const double rc2 = rc * rc;
  if (bx.tabN > 0 && bx.tabAlpha == alpha && bx.tabRc2 == rc2) return 0;
  const int eMin = -2;  // r >= 0.5 A; closer pairs take the library path
  int eMax = 1;
  while (std::ldexp(1.0, eMax) <= rc2 * (1.0 + 1e-12)) ++eMax;
  const int perOct = 1 << kCtBits, n = (eMax - eMin) * perOct;
  std::vector<double> tab((size_t)2 * kCtWords * n);
  const long double al = alpha, PI = 3.14159265358979323846264338327950288L;
  const long double c2 = 2.0L * al / sqrtl(PI);
  // Chebyshev polynomials T_k as monomial coefficients
  long double T[kCtCoef][kCtCoef] = {};
  T[0][0] = 1.0L;
  T[1][1] = 1.0L;
  for (int k = 2; k < kCtCoef; ++k)
    for (int j = 0; j < kCtCoef; ++j)
      T[k][j] = (j > 0 ? 2.0L * T[k - 1][j - 1] : 0.0L) - T[k - 2][j];
  long double node[kCtCoef];
  for (int j = 0; j < kCtCoef; ++j) node[j] = cosl(PI * (2 * j + 1) / (2.0L * kCtCoef));
  for (int i = 0; i < n; ++i) {
    const int oct = eMin + i / perOct, sub = i % perOct;
    const long double s0 = ldexpl(1.0L + (long double)sub / perOct, oct);
    const long double w = ldexpl(1.0L / perOct, oct);
    long double fv[2][kCtCoef];
    for (int j = 0; j < kCtCoef; ++j) {
      const long double sv = s0 + w * (node[j] + 1.0L) * 0.5L, r = sqrtl(sv);
      const long double f = erfcl(al * r) / r;
      fv[0][j] = f;
      fv[1][j] = (f + c2 * expl(-al * al * sv)) / sv;
    }
    for (int which = 0; which < 2; ++which) {
      long double mono[kCtCoef] = {};
      for (int k = 0; k < kCtCoef; ++k) {
        long double a = 0.0L;
        for (int j = 0; j < kCtCoef; ++j) {
          // T_k(node_j) = cos(k (2j+1) pi / (2n))
          a += fv[which][j] * cosl(PI * k * (2 * j + 1) / (2.0L * kCtCoef));
        }
        a *= (k == 0 ? 1.0L : 2.0L) / kCtCoef;
        for (int j = 0; j <= k; ++j) mono[j] += a * T[k][j];
      }
      double *tw = tab.data() + (size_t)which * kCtWords * n;
      for (int c = 0; c < 4; ++c) tw[(size_t)c * n + i] = (double)mono[c];
      float *hi = reinterpret_cast<float *>(tw + (size_t)4 * n);  // (c4,c5)[n], (c6,c7)[n]
      hi[2 * i] = (float)mono[4];
      hi[2 * i + 1] = (float)mono[5];
      hi[2 * n + 2 * i] = (float)mono[6];
      hi[2 * n + 2 * i + 1] = (float)mono[7];
    }
  }
  CK(bx.coulTab.reserve(tab.size()));
  // the stream may still read the previous table
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(bx.coulTab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  bx.tabAlpha = alpha;
  bx.tabRc2 = rc2;
  bx.tabN = n;
  bx.tabHi0 = (1023 + eMin) << kCtBits;
  return 0;
}

// warps per CTA of the box sweep: the energy kernel fits 32 warps in the
// register file (<= 64 regs/thread), the force kernel 20 (<= 102 regs/thread)
constexpr int kWarpsEnergy = 32, kWarpsForce = 20, kWarpsVirial = 16;

template <int FORCE>  // MODE_ENERGY, MODE_FORCE or MODE_VIRIAL
void launch_pair(gomcb200_engine *e, int b, const BoxParams &p, int slices, int useSmem,
                 int smemAtoms, size_t smemBytes, int grid, int cell0,
                 const int *gatePop = nullptr, int gateCap = 0) {
  BoxState &bx = e->box[b];
  constexpr int NW = FORCE == MODE_ENERGY ? kWarpsEnergy
                                          : (FORCE == MODE_VIRIAL ? kWarpsVirial : kWarpsForce);
  double *fx = e->force[GOMCB200_ATOM_FORCE][0].p;
  double *fy = e->force[GOMCB200_ATOM_FORCE][1].p;
  double *fz = e->force[GOMCB200_ATOM_FORCE][2].p;
#define LAUNCH_M(V, M)                                                              \
  do {                                                                              \
    cudaFuncSetAttribute(k_pair_box<V, M, NW>,                                      \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes); \
    k_pair_box<V, M, NW><<<grid, NW * 32, smemBytes, e->stream>>>(                  \
        p, bx.grid, slices, cell0, useSmem, smemAtoms, bx.cellStart.p, bx.sx.p, bx.sy.p,  \
        bx.sz.p, bx.sq.p, bx.skm.p, bx.sortedAtoms.p, e->blockA.p, e->blockB.p, fx, \
        fy, fz, gatePop, gateCap);                                                  \
  } while (0)
  // boxes with a fractional molecule run the instantiation that carries the soft-core branch
#define LAUNCH(V)                                 \
  do {                                            \
    if (bx.lambdaMol >= 0)                        \
      LAUNCH_M(V, FORCE | MODE_LAMBDA);           \
    else                                          \
      LAUNCH_M(V, FORCE);                         \
  } while (0)
  if (e->vdwKind == VDW_SHIFT)
    LAUNCH(VDW_SHIFT);
  else if (e->vdwKind == VDW_SWITCH)
    LAUNCH(VDW_SWITCH);
  else if (e->vdwKind == VDW_EXP6)
    LAUNCH(VDW_EXP6);
  else if (e->vdwKind == VDW_MARTINI)
    LAUNCH(VDW_MARTINI);
  else
    LAUNCH(VDW_STD);
#undef LAUNCH
#undef LAUNCH_M
  e->launches += 1;
}

// k_pair_box2 (pair2.cuh): orthogonal boxes.  FAST (Ewald real-space terms from the table)
// whenever the box has Ewald electrostatics and no fractional molecule.
template <int MODE>
int launch_pair2(gomcb200_engine *e, int b, const BoxParams &p, int slices, int grid, int cell0,
                 int *capOut) {
  BoxState &bx = e->box[b];
  // warps per CTA (one CTA per SM): 80 / 96 / 128 registers per thread
  constexpr int NW = MODE == MODE_ENERGY ? 24 : (MODE == MODE_VIRIAL ? 16 : 20);
  const bool lam = bx.lambdaMol >= 0;
  const bool fast = !lam && e->ewald && e->electrostatic;
  int tabN = 0;
  if (fast) {
    int rc = ensure_coul_table(e, b);
    if (rc) return rc;
    tabN = bx.tabN;
  }
  // staging capacity: everything the SM's shared memory leaves (one CTA per SM)
  const size_t fixed = pair2_smem_bytes(0, NW, MODE, tabN) + 6 * 1024 +
                       (MODE == MODE_VIRIAL ? 8 * 1024 : 0);
  if (e->smemOptin <= fixed + 56 * 256) return fail(GOMCB200_ECUDA, "shared memory too small");
  int cap = (int)((e->smemOptin - fixed) / 56) & ~1;
  // no more than the sweep can use: all neighbour cells of an average cell, with slack
  const double avg = (double)bx.nAtoms / bx.grid.nCells;
  const long long want = (long long)((MODE == MODE_FORCE ? 27.0 : 14.0) * (avg + 2.0) * 1.5) + 128;
  if (MODE != MODE_FORCE || want <= cap) cap = (int)std::min<long long>(cap, want) & ~1;
  cap = std::min(cap, 60000);
  Pair2Args A;
  A.p = p;
  A.g = bx.grid;
  A.slices = slices;
  A.cell0 = cell0;
  A.cap = cap;
  A.gateCap = MODE == MODE_ENERGY ? cap / 2 : cap;  // energy passes re-stage the self range
  // work items per CTA = i-atoms x candidate segments.  Whole candidate lists (one segment)
  // measured fastest as soon as every warp gets an item or two (an item's fixed cost
  // outweighs the better balance of smaller ones: 0.42 -> 0.40 ms on the 100k-atom box);
  // cells with few atoms are cut further
  const double nI = std::max(1.0, avg / slices);
  A.nSeg = std::max(1, std::min(kP2MaxSeg, (int)std::ceil(1.5 * NW / nI)));
  const double brs = p.boxRcutSq;
  A.cutF = (float)(brs * (1.0 + 1e-4) + 1e-6);
  A.tabN = tabN;
  A.tabHi0 = bx.tabHi0;
  A.tabF = bx.coulTab.p;
  A.tabG = bx.coulTab.p ? bx.coulTab.p + (size_t)kCtWords * tabN : nullptr;
  A.cellStart = bx.cellStart.p;
  A.sx = bx.sx.p;
  A.sy = bx.sy.p;
  A.sz = bx.sz.p;
  A.sq = bx.sq.p;
  A.skm = bx.skm.p;
  A.sortedAtoms = bx.sortedAtoms.p;
  A.maxCellPop = bx.maxCellPop.p;
  A.partLJ = e->blockA.p;
  A.partReal = e->blockB.p;
  A.fx = e->force[GOMCB200_ATOM_FORCE][0].p;
  A.fy = e->force[GOMCB200_ATOM_FORCE][1].p;
  A.fz = e->force[GOMCB200_ATOM_FORCE][2].p;
  const size_t smemBytes = pair2_smem_bytes(cap, NW, MODE, tabN);
#define LAUNCH2(V, M, F)                                                                   \
  do {                                                                                     \
    cudaFuncSetAttribute(k_pair_box2<V, M, NW, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smemBytes);                                                  \
    k_pair_box2<V, M, NW, F><<<grid, NW * 32, smemBytes, e->stream>>>(A);                  \
  } while (0)
#define LAUNCH2V(V)                          \
  do {                                       \
    if (lam)                                 \
      LAUNCH2(V, MODE | MODE_LAMBDA, false); \
    else if (fast)                           \
      LAUNCH2(V, MODE, true);                \
    else                                     \
      LAUNCH2(V, MODE, false);               \
  } while (0)
  if (e->vdwKind == VDW_SHIFT)
    LAUNCH2V(VDW_SHIFT);
  else if (e->vdwKind == VDW_SWITCH)
    LAUNCH2V(VDW_SWITCH);
  else if (e->vdwKind == VDW_EXP6)
    LAUNCH2V(VDW_EXP6);
  else if (e->vdwKind == VDW_MARTINI)
    LAUNCH2V(VDW_MARTINI);
  else
    LAUNCH2V(VDW_STD);
#undef LAUNCH2V
#undef LAUNCH2
  e->launches += 1;
  *capOut = A.gateCap;
  return 0;
}

int allreduce_energies(gomcb200_engine *e, double *buf, int n);

// pair sweep; results (LJ, real) land in e->result[0..1]; mode MODE_VIRIAL: the six
// tensor sums (LJ 11/22/33, Coulomb 11/22/33 without qqFact) in e->result[0..5]
int run_pair(gomcb200_engine *e, int b, int mode) {
  const bool force = mode == MODE_FORCE;
  int rc = ensure_cells(e, b);
  if (rc) return rc;
  BoxState &bx = e->box[b];
  BoxParams p = make_params(e, b);
  const int nCells = bx.grid.nCells;
  // multi-GPU: this rank owns cells [cell0, cell1) (x-major slab)
  const int cell0 = (int)(((long long)nCells * e->shardRank) / e->shardWorld);
  const int cell1 = (int)(((long long)nCells * (e->shardRank + 1)) / e->shardWorld);
  // CTAs per cell ("slices" of a cell's i-atoms): every slice re-stages the whole
  // neighbourhood (~8 % of a full cell's work), and CTAs run in waves of one per SM; take the
  // slice count with the least  waves x (staging + work / slices)
  const int owned = std::max(1, cell1 - cell0);
  int slices = 1;
  {
    double best = 1e300;
    for (int sl = 1; sl <= 16; ++sl) {
      const double waves = std::ceil((double)owned * sl / e->numSMs);
      const double cost = waves * (0.08 + 1.0 / sl);
      if (cost < best - 1e-12) {
        best = cost;
        slices = sl;
      }
    }
  }
  int grid = (cell1 - cell0) * slices;
  // shared-memory staging of the neighbour cells (40 B per atom)
  const int nWarps = mode == MODE_ENERGY ? kWarpsEnergy
                                         : (mode == MODE_VIRIAL ? kWarpsVirial : kWarpsForce);
  const size_t queueBytes = sizeof(WarpQueue) * nWarps;
  size_t staticSmem = (mode == MODE_VIRIAL ? 28 : 12) * 1024 + queueBytes;
  size_t capAtoms = (e->smemOptin > staticSmem ? (e->smemOptin - staticSmem) : 0) / 40;
  double avg = (double)bx.nAtoms / nCells;
  size_t want = (size_t)((force ? 27.0 : 14.0) * avg * 1.4) + 96;
  int smemAtoms = (int)std::min(capAtoms, want);
  smemAtoms &= ~1;
  int useSmem = smemAtoms >= 64;
  size_t smemBytes = queueBytes + (useSmem ? (size_t)smemAtoms * 40 : 0);
  CK(e->blockA.reserve((mode == MODE_VIRIAL ? 6 : 1) * (size_t)grid + 1024));
  CK(e->blockB.reserve(grid + 1024));
  if (grid == 0) {
    CK(cudaMemsetAsync(e->result.p, 0, 8 * sizeof(double), e->stream));
    return 0;
  }
  // orthogonal boxes: k_pair_box2, with the first kernel queued behind it as the fallback
  // for cells too full to stage (both test the same device-side population count)
  const int *gate = nullptr;
  int gateCap = 0;
  const bool reduceForces = force && e->comm && e->shardWorld > 1;
  if (reduceForces) {
    // this rank writes the atoms of its cell slab (complete forces: full shell); the rest
    // stays zero and the three arrays are summed over the ranks below
    if (e->nBoxes > 1)
      return fail(GOMCB200_EINVAL, "sharded BoxForce supports single-box engines (the force "
                                   "arrays are summed over the ranks as a whole)");
    for (int c = 0; c < 3; ++c)
      CK(cudaMemsetAsync(e->force[GOMCB200_ATOM_FORCE][c].p, 0,
                         sizeof(double) * (size_t)e->nAtoms, e->stream));
  }
  if (e->pairAlgo == 1 && !bx.nonOrth) {
    if (mode == MODE_VIRIAL)
      rc = launch_pair2<MODE_VIRIAL>(e, b, p, slices, grid, cell0, &gateCap);
    else if (force)
      rc = launch_pair2<MODE_FORCE>(e, b, p, slices, grid, cell0, &gateCap);
    else
      rc = launch_pair2<MODE_ENERGY>(e, b, p, slices, grid, cell0, &gateCap);
    if (rc) return rc;
    CK(cudaGetLastError());
    gate = bx.maxCellPop.p;
  }
  if (mode == MODE_VIRIAL) {
    launch_pair<MODE_VIRIAL>(e, b, p, slices, useSmem, smemAtoms, smemBytes, grid, cell0, gate,
                             gateCap);
    CK(cudaGetLastError());
    const double *a = e->blockA.p;
    k_final_reduce<<<1, 1024, 0, e->stream>>>(grid, 4, a, e->blockB.p, a + 2 * (size_t)grid,
                                             a + 3 * (size_t)grid, e->result.p);
    k_final_reduce<<<1, 1024, 0, e->stream>>>(grid, 2, a + 4 * (size_t)grid,
                                             a + 5 * (size_t)grid, nullptr, nullptr,
                                             e->result.p + 4);
    e->launches += 2;
    CK(cudaGetLastError());
    return 0;
  }
  if (force) {
    // ResetForce (src/CalculateEnergy.cpp:1408-1428) is implicit: every atom
    // and molecule of the box is overwritten below.
    launch_pair<MODE_FORCE>(e, b, p, slices, useSmem, smemAtoms, smemBytes, grid, cell0, gate,
                            gateCap);
  } else {
    launch_pair<MODE_ENERGY>(e, b, p, slices, useSmem, smemAtoms, smemBytes, grid, cell0, gate,
                             gateCap);
  }
  CK(cudaGetLastError());
  k_final_reduce<<<1, 1024, 0, e->stream>>>(grid, 2, e->blockA.p, e->blockB.p, nullptr,
                                           nullptr, e->result.p);
  e->launches += 1;
  if (reduceForces)
    for (int c = 0; c < 3; ++c) {
      rc = allreduce_energies(e, e->force[GOMCB200_ATOM_FORCE][c].p, e->nAtoms);
      if (rc) return rc;
    }
  if (force && bx.nMols > 0) {
    k_mol_force<<<(bx.nMols + 255) / 256, 256, 0, e->stream>>>(
        bx.nMols, bx.molList.p, e->molStart.p, e->force[GOMCB200_ATOM_FORCE][0].p,
        e->force[GOMCB200_ATOM_FORCE][1].p, e->force[GOMCB200_ATOM_FORCE][2].p,
        e->force[GOMCB200_MOL_FORCE][0].p, e->force[GOMCB200_MOL_FORCE][1].p,
        e->force[GOMCB200_MOL_FORCE][2].p);
    e->launches += 1;
  }
  CK(cudaGetLastError());
  return 0;
}

// ---- reciprocal space -------------------------------------------------------
// Ewald::RecipInitOrth (src/Ewald.cpp:847-903): same loop nest, same order,
// same floating-point expressions, so the k list is index-compatible with the
// host arrays of an unmodified GOMC.  Also derives the (a,b)-row table the
// factorised kernel needs.
struct RowRec {
  int a, b, cmax, start;
};

// Enumerates the half-space k list.  ks == nullptr: count only.
// Ewald::RecipInitNonOrth / RecipCountInit for a slanted cell (src/Ewald.cpp:905-1018):
// reciprocal rows = adjoint(cell) * 2 pi / det; no (a,b)-row table (the valid c
// range of a row is not symmetric), so such boxes use the direct kernels.
int recip_enumerate_nonorth(const gomcb200_engine *e, int b, const double ax[3], KSet *ks) {
  const BoxState &bx = e->box[b];
  const double alpha = e->alpha[b];
  const double recip_rcut = e->recipRcut[b];
  const double rr2 = recip_rcut * recip_rcut;
  const double alpsqr4 = 1.0 / (4.0 * (alpha * alpha));
  double cb[9], inv[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cb[3 * r + c] = bx.B[3 * r + c] * ax[r];
  auto X = [&](int i) { return cb[3 * i]; };
  auto Y = [&](int i) { return cb[3 * i + 1]; };
  auto Z = [&](int i) { return cb[3 * i + 2]; };
  inv[0] = Y(1) * Z(2) - Y(2) * Z(1);   // XYZArray::AdjointMatrix, src/XYZArray.h:507-522
  inv[1] = Y(2) * Z(0) - Y(0) * Z(2);
  inv[2] = Y(0) * Z(1) - Y(1) * Z(0);
  inv[3] = X(2) * Z(1) - X(1) * Z(2);
  inv[4] = X(0) * Z(2) - X(2) * Z(0);
  inv[5] = X(1) * Z(0) - X(0) * Z(1);
  inv[6] = X(1) * Y(2) - X(2) * Y(1);
  inv[7] = X(2) * Y(0) - X(0) * Y(2);
  inv[8] = X(0) * Y(1) - X(1) * Y(0);
  const double det = X(0) * inv[0] + X(1) * inv[1] + X(2) * inv[2];
  const double bxc[3] = {Y(1) * Z(2) - Z(1) * Y(2), Z(1) * X(2) - X(1) * Z(2),
                         X(1) * Y(2) - Y(1) * X(2)};
  const double volume = std::fabs(X(0) * bxc[0] + Y(0) * bxc[1] + Z(0) * bxc[2]);
  for (int i = 0; i < 9; ++i) inv[i] *= (2.0 * M_PI) / det;
  const double vol = volume / (4.0 * M_PI);
  int nmax[3];
  for (int d = 0; d < 3; ++d) nmax[d] = int(recip_rcut * ax[d] / (2.0 * M_PI)) + 1;
  if (ks) {
    ks->hkx.clear(); ks->hky.clear(); ks->hkz.clear(); ks->hhsqr.clear();
    ks->hprefact.clear();
    for (int d = 0; d < 3; ++d) { ks->nmax[d] = nmax[d]; ks->cv[d] = 0.0; }
    ks->kmax = std::max(std::max(nmax[0], nmax[1]), std::max(nmax[1], nmax[2]));
  }
  int counter = 0;
  for (int ix = 0; ix <= nmax[0]; ix++) {
    int nky_min = (ix == 0) ? 0 : -nmax[1];
    for (int iy = nky_min; iy <= nmax[1]; iy++) {
      int nkz_min = (ix == 0 && iy == 0) ? 1 : -nmax[2];
      for (int iz = nkz_min; iz <= nmax[2]; iz++) {
        double kX = inv[0] * ix + inv[1] * iy + inv[2] * iz;
        double kY = inv[3] * ix + inv[4] * iy + inv[5] * iz;
        double kZ = inv[6] * ix + inv[7] * iy + inv[8] * iz;
        double ksqr = kX * kX + kY * kY + kZ * kZ;
        if (ksqr < rr2) {
          if (ks) {
            ks->hkx.push_back(kX);
            ks->hky.push_back(kY);
            ks->hkz.push_back(kZ);
            ks->hhsqr.push_back(ksqr);
            ks->hprefact.push_back(kQQFact * exp(-ksqr * alpsqr4) / (ksqr * vol));
          }
          counter++;
        }
      }
    }
  }
  return counter;
}

int recip_enumerate(const gomcb200_engine *e, int b, const double ax[3], KSet *ks,
                    std::vector<RowRec> *rowsOut, double volume = 0.0) {
  if (e->box[b].nonOrth) {
    if (rowsOut) rowsOut->clear();
    return recip_enumerate_nonorth(e, b, ax, ks);
  }
  const double alpha = e->alpha[b];
  const double recip_rcut = e->recipRcut[b];
  const double rr2 = recip_rcut * recip_rcut;
  const double alpsqr4 = 1.0 / (4.0 * (alpha * alpha));
  double cv[3];
  for (int d = 0; d < 3; ++d) cv[d] = (1.0 / ax[d]) * (2.0 * M_PI);
  // boxAxes.volume[box]: the product of the axes, except after BoxDimensions::SetVolume
  // (a volume trial), where the caller passes the stored value (src/Ewald.cpp:857)
  const double vol = (volume > 0.0 ? volume : ax[0] * ax[1] * ax[2]) / (4.0 * M_PI);
  int nmax[3];
  for (int d = 0; d < 3; ++d) nmax[d] = int(recip_rcut * ax[d] / (2.0 * M_PI)) + 1;
  if (ks) {
    for (int d = 0; d < 3; ++d) { ks->nmax[d] = nmax[d]; ks->cv[d] = cv[d]; ks->L[d] = ax[d]; }
    ks->kmax = std::max(std::max(nmax[0], nmax[1]), std::max(nmax[1], nmax[2]));
  }
  // Same loop nest, order and floating-point expressions as the reference; for large
  // boxes (1e6 k-vectors per volume trial) the x slabs are enumerated by a few host
  // threads in two passes (count, then fill at the prefix offsets), which leaves the k
  // list and its order bit-identical to the serial enumeration.
  const int nX = nmax[0] + 1;
  struct RowScan {
    int iy, zlo, zhi, cnt;
  };
  std::vector<std::vector<RowScan>> slabRows(nX);
  std::vector<int> slabCount(nX, 0), slabStart(nX + 1, 0);
  auto scan_slab = [&](int ix) {
    int nky_min = (ix == 0) ? 0 : -nmax[1];
    int total = 0;
    for (int iy = nky_min; iy <= nmax[1]; iy++) {
      int nkz_min = (ix == 0 && iy == 0) ? 1 : -nmax[2];
      RowScan r = {iy, 0, 0, 0};
      for (int iz = nkz_min; iz <= nmax[2]; iz++) {
        double kX = cv[0] * ix, kY = cv[1] * iy, kZ = cv[2] * iz;
        double ksqr = kX * kX + kY * kY + kZ * kZ;
        if (ksqr < rr2) {
          if (r.cnt == 0) r.zlo = iz;
          r.zhi = iz;
          ++r.cnt;
        }
      }
      if (r.cnt > 0) slabRows[ix].push_back(r);
      total += r.cnt;
    }
    slabCount[ix] = total;
  };
  auto fill_slab = [&](int ix) {
    int pos = slabStart[ix];
    for (const RowScan &r : slabRows[ix]) {
      int nkz_min = (ix == 0 && r.iy == 0) ? 1 : -nmax[2];
      for (int iz = nkz_min; iz <= nmax[2]; iz++) {
        double kX = cv[0] * ix, kY = cv[1] * r.iy, kZ = cv[2] * iz;
        double ksqr = kX * kX + kY * kY + kZ * kZ;
        if (ksqr < rr2) {
          ks->hkx[pos] = kX;
          ks->hky[pos] = kY;
          ks->hkz[pos] = kZ;
          ks->hhsqr[pos] = ksqr;
          ks->hprefact[pos] = kQQFact * exp(-ksqr * alpsqr4) / (ksqr * vol);
          ++pos;
        }
      }
    }
  };
  auto run_slabs = [&](auto &&fn) {
    const long long cube = (long long)nX * (2 * nmax[1] + 1) * (2 * nmax[2] + 1);
    int nThreads = cube > 200000 ? (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (nThreads <= 1) {
      for (int ix = 0; ix < nX; ++ix) fn(ix);
      return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t)
      pool.emplace_back([&]() {
        for (int ix = next.fetch_add(1); ix < nX; ix = next.fetch_add(1)) fn(ix);
      });
    for (auto &th : pool) th.join();
  };
  run_slabs(scan_slab);
  for (int ix = 0; ix < nX; ++ix) slabStart[ix + 1] = slabStart[ix] + slabCount[ix];
  const int counter = slabStart[nX];
  if (ks) {
    ks->hkx.resize(counter); ks->hky.resize(counter); ks->hkz.resize(counter);
    ks->hhsqr.resize(counter); ks->hprefact.resize(counter);
    run_slabs(fill_slab);
  }
  if (rowsOut) {
    for (int ix = 0; ix < nX; ++ix) {
      int first = slabStart[ix];
      for (const RowScan &r : slabRows[ix]) {
        // valid c form one run symmetric about 0 (ksqr is even and monotone
        // in |c|); anything else would break the factorised indexing.
        bool origin = (ix == 0 && r.iy == 0);
        bool ok = origin ? (r.zlo == 1 && r.cnt == r.zhi)
                         : (r.zlo == -r.zhi && r.cnt == 2 * r.zhi + 1);
        if (!ok) return -1;
        rowsOut->push_back({ix, r.iy, r.zhi, first});
        first += r.cnt;
      }
    }
  }
  return counter;
}

int build_plan(gomcb200_engine *e, KSet &ks, std::vector<RowRec> &rows) {
  std::stable_sort(rows.begin(), rows.end(),
                   [](const RowRec &x, const RowRec &y) { return x.cmax > y.cmax; });
  std::vector<int4> hrows, htiles;
  size_t i = 0;
  int maxRows = 0;
  while (i < rows.size()) {
    int cmaxT = rows[i].cmax;
    int CG = (cmaxT + 1 + kTC - 1) / kTC;
    int RG = std::min(kFactThreads / CG, kMaxRG);
    int R = RG * kTR;
    int4 t = make_int4((int)hrows.size(), RG, CG, cmaxT);
    for (int r = 0; r < R; ++r) {
      if (i + r < rows.size()) {
        const RowRec &rw = rows[i + r];
        hrows.push_back(make_int4(rw.a, rw.b, rw.cmax, rw.start));
      } else {
        hrows.push_back(make_int4(0, 0, -1, 0));
      }
    }
    htiles.push_back(t);
    maxRows = std::max(maxRows, R);
    i += R;
  }
  ks.nTiles = (int)htiles.size();
  ks.nRowsPadded = (int)hrows.size();
  ks.maxRows = maxRows;
  CK(ks.rows.reserve(hrows.size() + 1));
  CK(ks.tiles.reserve(htiles.size() + 1));
  if (!hrows.empty()) {
    CK(cudaMemcpyAsync(ks.rows.p, hrows.data(), hrows.size() * sizeof(int4),
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(ks.tiles.p, htiles.data(), htiles.size() * sizeof(int4),
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  ks.planValid = true;
  // the DMMA plan is derived from the sorted rows per shard (build_mma_tiles)
  ks.hRowsSorted.clear();
  for (const RowRec &rw : rows) ks.hRowsSorted.push_back(make_int4(rw.a, rw.b, rw.cmax, rw.start));
  ks.mmaValid = !ks.hRowsSorted.empty();
  ks.itemsForAtoms = -1;
  ks.tilesForShard = -1;
  // reciprocal-force plan: 16-row tiles of the same sorted rows
  {
    std::vector<int4> frows(ks.hRowsSorted);
    while (frows.size() % kFmRows) frows.push_back(make_int4(0, 0, -1, 0));
    std::vector<FmTile> ft;
    size_t wOff = 0;
    for (size_t rb = 0; rb < frows.size(); rb += kFmRows) {
      int cmaxT = frows[rb].z;
      FmTile t;
      t.rowBegin = (int)rb;
      t.KT = (2 * (cmaxT + 1) + 3) & ~3;
      t.wOff = (int)wOff;
      t.pad = 0;
      wOff += (size_t)t.KT * kFmWS;
      ft.push_back(t);
    }
    std::vector<FmBlock> fb;
    for (const FmTile &t : ft)
      for (int k0 = 0; k0 < t.KT; k0 += kFmKB) {
        FmBlock b;
        b.rowBegin = t.rowBegin;
        b.k0 = k0;
        b.kLen = std::min(kFmKB, t.KT - k0);
        b.wOff = t.wOff + k0 * kFmWS;
        b.first = k0 == 0;
        b.last = k0 + kFmKB >= t.KT;
        b.pad0 = b.pad1 = 0;
        fb.push_back(b);
      }
    CK(ks.fmBlocks.reserve(fb.size() + 1));
    if (!fb.empty())
      CK(cudaMemcpyAsync(ks.fmBlocks.p, fb.data(), fb.size() * sizeof(FmBlock),
                         cudaMemcpyHostToDevice, e->stream));
    ks.fmNBlocks = (int)fb.size();
    CK(ks.fmRows.reserve(frows.size() + 1));
    CK(ks.fmTiles.reserve(ft.size() + 1));
    CK(ks.fmW.reserve(wOff + 16));
    if (!frows.empty()) {
      CK(cudaMemcpyAsync(ks.fmRows.p, frows.data(), frows.size() * sizeof(int4),
                         cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(ks.fmTiles.p, ft.data(), ft.size() * sizeof(FmTile),
                         cudaMemcpyHostToDevice, e->stream));
      CK(cudaStreamSynchronize(e->stream));
    }
    ks.fmNTiles = (int)ft.size();
    ks.fmValid = !ft.empty();
  }
  return 0;
}

// Tiles with few columns are bound by their table loads, not by the DMMAs: a
// chunk of an NT <= kMmaMinCostNT tile costs about as much as one of NT == kMmaMinCostNT.
constexpr int kMmaMinCostNT = 4;

// DMMA row tiles of THIS rank.  Multi-GPU: a rank owns a contiguous range of the
// cmax-sorted (a,b) rows, balanced by the number of k-vectors in them, so that
// every S(k) is complete on exactly one GPU and only scalar energies are
// exchanged.  Rows are cut into 128-row tiles, columns into blocks of 40.
int build_mma_tiles(gomcb200_engine *e, KSet &ks) {
  const int shardKey = e->shardRank * 1024 + e->shardWorld;
  if (ks.tilesForShard == shardKey) return 0;
  const size_t nRows = ks.hRowsSorted.size();
  std::vector<long long> cum(nRows + 1, 0);
  for (size_t r = 0; r < nRows; ++r) cum[r + 1] = cum[r] + std::max(ks.hRowsSorted[r].z, kMmaMinCostNT * 4) + 4;
  auto cut = [&](int rk) {
    long long w = cum[nRows] * rk / e->shardWorld;
    return (size_t)(std::lower_bound(cum.begin(), cum.end(), w) - cum.begin());
  };
  const size_t r0 = e->shardRank == 0 ? 0 : cut(e->shardRank);
  const size_t r1 = e->shardRank == e->shardWorld - 1 ? nRows : cut(e->shardRank + 1);
  std::vector<int4> mrows(ks.hRowsSorted.begin() + r0, ks.hRowsSorted.begin() + r1);
  while (mrows.size() % kMmaRows) mrows.push_back(make_int4(0, 0, -1, 0));
  const int KZ1 = ks.nmax[2] + 1;
  const int colBlocks = (KZ1 + 4 * kMmaMaxNT - 1) / (4 * kMmaMaxNT);
  ks.hMmaTiles.clear();
  for (int cb = 0; cb < colBlocks; ++cb) {
    const int c0 = cb * 4 * kMmaMaxNT;
    for (size_t rb = 0; rb < mrows.size(); rb += kMmaRows) {
      int cmaxT = mrows[rb].z;  // rows are sorted: first row has the largest cmax
      if (cmaxT < c0) break;
      int need = std::min(cmaxT - c0 + 1, 4 * kMmaMaxNT);
      int NT = std::min(kMmaMaxNT, (need + 3) / 4);
      ks.hMmaTiles.push_back(make_int4((int)rb, c0, NT, cmaxT));
    }
  }
  {  // int8 path: tiles of 8 a values x 8 b values (the XY table slice a tile needs is then
     // 16 entries instead of all of them), one tile per column block it reaches.  Multi-GPU:
     // the blocks are cut from ALL rows of the box (a rank's cmax band is a thin ring in the
     // (a, b) plane and would fill 8 x 8 blocks badly) and the equal-cost tiles are dealt
     // round-robin, so again every S(k) is complete on exactly one GPU.
    ks.i8NSL = KZ1 <= 32 ? 6 : 5;
    ks.i8CPer = KZ1 <= 32 ? ((KZ1 + 7) / 8) * 8 : (KZ1 <= 40 ? ((KZ1 + 7) / 8) * 8 : 40);
    ks.i8Nb = 2 * ks.i8CPer;
    const int ncb = (KZ1 + ks.i8CPer - 1) / ks.i8CPer;
    const int KX1 = ks.nmax[0] + 1;
    std::map<std::pair<int, int>, std::vector<int4>> blocks;
    for (const int4 &rw : ks.hRowsSorted) {
      if (rw.z < 0) continue;
      const int bb = rw.y >= 0 ? rw.y / 8 : -((-rw.y + 7) / 8);  // floor(b / 8)
      auto &slots = blocks[{rw.x / 8, bb}];
      if (slots.empty()) slots.assign(kI8Pairs, make_int4(0, 0, -1, 0));
      slots[(rw.x & 7) * 8 + (rw.y - 8 * bb)] = rw;
    }
    std::vector<int4> irows, it;
    std::vector<std::pair<int, int4>> order;  // (cmax of the tile, descriptor), big first
    for (auto &kv : blocks) {
      const int ab = kv.first.first, bb = kv.first.second;
      int cmaxT = -1;
      for (const int4 &rw : kv.second) cmaxT = std::max(cmaxT, rw.z);
      const int yFirst = bb >= 0 ? 8 * bb : -(8 * bb + 7);  // smallest |b| of the block
      const int rb = (int)irows.size();
      irows.insert(irows.end(), kv.second.begin(), kv.second.end());
      order.push_back({cmaxT, make_int4(rb, 0, 8 * ab, KX1 + yFirst)});
    }
    std::sort(order.begin(), order.end(),
              [](const std::pair<int, int4> &l, const std::pair<int, int4> &r) {
                return l.first > r.first;
              });
    int nAll = 0;
    for (int cb = 0; cb < ncb; ++cb)
      for (auto &o : order) {
        if (o.first < ks.i8CPer * cb) break;
        int4 d = o.second;
        d.y = cb;
        if (nAll++ % e->shardWorld == e->shardRank) it.push_back(d);
      }
    ks.i8NTilesAll = nAll;
    ks.i8NTiles = (int)it.size();
    ks.i8NCB = ncb;
    CK(ks.i8Rows.reserve(irows.size() + 1));
    CK(ks.i8Tiles.reserve(it.size() + 1));
    if (!irows.empty())
      CK(cudaMemcpyAsync(ks.i8Rows.p, irows.data(), irows.size() * sizeof(int4),
                         cudaMemcpyHostToDevice, e->stream));
    if (!it.empty())
      CK(cudaMemcpyAsync(ks.i8Tiles.p, it.data(), it.size() * sizeof(int4),
                         cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // the vectors are locals
  }
  int ZS = colBlocks * 4 * kMmaMaxNT;
  while (ZS % 8 != 2) ++ZS;  // conflict-free B fragments
  ks.mmaZS = ZS;
  CK(ks.mmaRows.reserve(mrows.size() + 1));
  CK(ks.mmaTiles.reserve(ks.hMmaTiles.size() + 1));
  if (!mrows.empty())
    CK(cudaMemcpyAsync(ks.mmaRows.p, mrows.data(), mrows.size() * sizeof(int4),
                       cudaMemcpyHostToDevice, e->stream));
  if (!ks.hMmaTiles.empty())
    CK(cudaMemcpyAsync(ks.mmaTiles.p, ks.hMmaTiles.data(), ks.hMmaTiles.size() * sizeof(int4),
                       cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  ks.tilesForShard = shardKey;
  ks.itemsForAtoms = -1;
  return 0;
}

// Work split of the DMMA kernel.  The job is a line of (tile, atom chunk) units,
// unit weight = NT of the tile + 1 (A-tile generation); this rank takes its
// contiguous share of the line and cuts it into one equal piece per SM
// (persistent CTAs, single wave).  A piece = 1..3 segments {tile, chunkBegin,
// chunkEnd, slab}; slab numbers the segments of a tile on this rank.
int build_mma_segments(gomcb200_engine *e, KSet &ks, int nAt, int AT) {
  const int shardKey = e->shardRank * 1024 + e->shardWorld;
  int rcT = build_mma_tiles(e, ks);
  if (rcT) return rcT;
  if (ks.itemsForAtoms == nAt && ks.itemsForShard == shardKey && ks.itemsAT == AT) return 0;
  const int nT = (int)ks.hMmaTiles.size();
  const long long nChunks = std::max(1, (nAt + AT - 1) / AT);
  // measured cost model (profiles/r1g_mma_cost_model.txt, clock64 per CTA, AT = 24):
  // cycles = sum_segments [ 44400 + chunks * cost(NT) ]; few-column tiles are
  // bound by their table loads (floor ~3000 cycles per chunk)
  static const long long kChunkCycles[11] = {3000, 3000, 3320, 3785, 4235, 4830,
                                             5547, 6298, 7096, 7806, 8632};
  auto chunkCost = [&](int t) {
    return kChunkCycles[std::min(10, std::max(0, ks.hMmaTiles[t].z))] * AT / 24;
  };
  const long long segCost = 44400;
  long long total = 0;
  for (int t = 0; t < nT; ++t) total += chunkCost(t) * nChunks + segCost;
  const int nCtas = (int)std::max<long long>(1, std::min<long long>(e->numSMs, total / (8 * segCost) + 1));
  const long long target = (total + segCost * nCtas) / nCtas;
  std::vector<int4> segs;
  std::vector<int> ctaSeg(1, 0);
  std::vector<int> slabOfTile(nT, 0);
  long long acc = 0;
  int cta = 0;
  for (int t = 0; t < nT; ++t) {
    long long c = 0;
    const long long w = chunkCost(t);
    while (c < nChunks) {
      acc += segCost;
      long long room = target - acc;
      long long take = std::max(1LL, (room + w / 2) / w);
      if (cta == nCtas - 1) take = nChunks - c;  // the last CTA takes what is left
      take = std::min(take, nChunks - c);
      segs.push_back(make_int4(t, (int)c, (int)(c + take), slabOfTile[t]++));
      acc += take * w;
      c += take;
      if (acc + w / 2 >= target && cta < nCtas - 1) {
        ctaSeg.push_back((int)segs.size());
        ++cta;
        acc = 0;
      }
    }
  }
  while ((int)ctaSeg.size() < nCtas + 1) ctaSeg.push_back((int)segs.size());
  int maxSlabs = 1;
  for (int t = 0; t < nT; ++t) maxSlabs = std::max(maxSlabs, slabOfTile[t]);
  CK(ks.mmaSegs.reserve(segs.size() + 1));
  CK(ks.mmaCtaSeg.reserve(ctaSeg.size() + 1));
  if (!segs.empty())
    CK(cudaMemcpyAsync(ks.mmaSegs.p, segs.data(), segs.size() * sizeof(int4),
                       cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(ks.mmaCtaSeg.p, ctaSeg.data(), ctaSeg.size() * sizeof(int),
                     cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  ks.hSegs = segs;
  ks.hCtaSeg = ctaSeg;
  ks.nCtas = nCtas;
  ks.maxSlabs = maxSlabs;
  ks.itemsForAtoms = nAt;
  ks.itemsForShard = shardKey;
  ks.itemsAT = AT;
  return 0;
}

int upload_kset(gomcb200_engine *e, KSet &ks, bool prefactOnly = false) {
  size_t n = ks.hkx.size();
  size_t cap = std::max<size_t>(n, (size_t)e->imageTotal) + 1;
  CK(ks.kx.reserve(cap));
  CK(ks.ky.reserve(cap));
  CK(ks.kz.reserve(cap));
  CK(ks.hsqr.reserve(cap));
  CK(ks.prefact.reserve(cap));
  if (n) {
    size_t bytes = n * sizeof(double);
    if (!prefactOnly) {
      CK(cudaMemcpyAsync(ks.kx.p, ks.hkx.data(), bytes, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(ks.ky.p, ks.hky.data(), bytes, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(ks.kz.p, ks.hkz.data(), bytes, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(ks.hsqr.p, ks.hhsqr.data(), bytes, cudaMemcpyHostToDevice, e->stream));
    }
    CK(cudaMemcpyAsync(ks.prefact.p, ks.hprefact.data(), bytes, cudaMemcpyHostToDevice,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  ks.n = (int)n;
  return 0;
}

int ensure_sums(gomcb200_engine *e, BoxState &bx, size_t n) {
  size_t cap = std::max<size_t>(n, (size_t)e->imageTotal) + 1;
  for (int i = 0; i < 4; ++i) {
    if (bx.sum[i].cap < cap) {
      CK(bx.sum[i].reserve(cap));
      CK(cudaMemsetAsync(bx.sum[i].p, 0, bx.sum[i].cap * sizeof(double), e->stream));
    }
  }
  return 0;
}

int ensure_packed(gomcb200_engine *e, int b) {
  BoxState &bx = e->box[b];
  if (!bx.packedDirty) return 0;
  CK(bx.packed.reserve(bx.nCharged + 1));
  if (bx.nCharged > 0) {
    k_pack_charged<<<(bx.nCharged + 255) / 256, 256, 0, e->stream>>>(
        bx.nCharged, bx.chargedList.p, e->x.p, e->y.p, e->z.p, e->qEff.p, bx.packed.p);
    e->launches += 1;
  }
  bx.packedDirty = false;
  return 0;
}

// Structure factor of box b on k set `ks` into sumRnew/sumInew; energy into
// result[0].  (BoxReciprocalSetup: ks = new set; BoxReciprocalSums: Ref set.)
int nufft_allgather_cb(void *ctx, void *buf, size_t bytesPerRank, cudaStream_t st) {
  gomcb200_engine *e = static_cast<gomcb200_engine *>(ctx);
  std::string err;
  if (gbc::comm_allgather_inplace(e->comm, buf, bytesPerRank, st, err))
    return fail(GOMCB200_ECUDA, "%s", err.c_str());
  e->launches += 1;
  return 0;
}

// sum of n doubles at buf over the ranks of the communicator (no-op without one)
int allreduce_energies(gomcb200_engine *e, double *buf, int n) {
  if (!e->comm || e->shardWorld <= 1) return 0;
  std::string err;
  if (gbc::comm_allreduce_sum(e->comm, buf, (size_t)n, e->stream, err))
    return fail(GOMCB200_ECUDA, "%s", err.c_str());
  e->launches += 1;
  return 0;
}

int run_recip_sums(gomcb200_engine *e, int b, KSet &ks) {
  BoxState &bx = e->box[b];
  const int nk = ks.n;
  int rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  rc = ensure_packed(e, b);
  if (rc) return rc;
  const int nkStride = (nk + 31) & ~31;
  const int nBlocks = (nk + 255) / 256;
  CK(e->blockA.reserve(nBlocks + 1024));
  if (nk == 0) {
    CK(cudaMemsetAsync(e->result.p, 0, 8 * sizeof(double), e->stream));
    return 0;
  }
  const int nAt = bx.nCharged;
  int nSlabs = 1;
  if (e->timing && !e->capturing) cudaEventRecord(e->ev[2], e->stream);
  bool i8Done = false, nufftDone = false;
  const bool wantI8 = e->recipAlgo == 3;
  if ((e->recipAlgo == 4 || e->recipAlgo == 5) && ks.planValid && ks.ngValid && nAt > 0) {
    // non-uniform FFT: the whole structure factor at O(N w^3 + n^3 log n).  Multi-GPU: it is
    // cheap enough to be replicated (every rank then holds the complete sums, which the
    // single-molecule deltas need); rank 0 alone reports the energy.
    CK(e->part.reserve((size_t)2 * nkStride + 64));
    // With a communicator the spread and the first two FFT passes are sharded by x slabs and
    // the pruned slabs all-gathered (nufft.h); without one (set_shard only) it is replicated.
    gbn::NufftShard sh = {e->shardRank, e->shardWorld, nufft_allgather_cb, e};
    rc = gbn::nufft_type1(e->nufft, e->stream, ks.ng, ks.L, bx.packed.p, nAt, ks.rows.p,
                          ks.nRowsPadded, e->part.p, e->part.p + nkStride, &e->launches,
                          (e->comm && e->shardWorld > 1) ? &sh : nullptr);
    if (rc) return fail(GOMCB200_ECUDA, "nufft_type1: %s", gbn::nufft_last_error(e->nufft));
    nSlabs = 1;
    nufftDone = true;
  }
  if (!nufftDone && wantI8 && ks.mmaValid && nAt > 0) {
    rc = build_mma_tiles(e, ks);
    if (rc) return rc;
    I8Args ia;
    ia.KX1 = ks.nmax[0] + 1;
    const int KY1 = ks.nmax[1] + 1;
    ia.XYS = ia.KX1 + KY1;
    ia.NCB = ks.i8NCB;
    ia.cPer = ks.i8CPer;
    ia.nb = ks.i8Nb;
    ia.nTiles = ks.i8NTiles;
    const int NSL = ks.i8NSL;
    ia.nSteps = (nAt + kI8StepAtoms - 1) / kI8StepAtoms;
    ia.nChunks = (ia.nSteps + kI8ChunkSteps - 1) / kI8ChunkSteps;
    ia.nkStride = nkStride;
    const size_t planeB = (size_t)ia.nb * 32;
    const int TR = i8_tr(NSL), BR = i8_br(NSL);
    const size_t smem = (size_t)TR * kI8TabEntries * 32 * 16 + kI8AR * (size_t)NSL * kI8PlaneA +
                        BR * (size_t)NSL * planeB + (2 * TR + 2 * BR + 2 * kI8AR + 1) * 8 + 16 +
                        34 * 16;  // + the issuer's instruction list (16 B aligned)
    if (smem + 1024 <= e->smemOptin && ks.i8NTilesAll > 0 && ks.i8NTiles == 0) {
      // more ranks than tiles: this rank owns no k-vector
      CK(e->part.reserve((size_t)2 * nkStride + 64));
      CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)2 * nkStride, e->stream));
      i8Done = true;
    } else if (smem + 1024 <= e->smemOptin && ks.i8NTiles > 0) {
      // |A| < 1 needs q / qScale with qScale a power of two above every |q| of the box
      // (|qEff| = |q| sqrt(lambda) <= |q|: the bound taken when the box was filled holds)
      const double qmax = bx.qMaxAbs;
      int ex = 0;
      std::frexp(qmax, &ex);  // qmax = m * 2^ex, m in [0.5, 1)
      const double qScale = std::ldexp(1.0, ex);
      ia.outScale = std::ldexp(qScale, 8 * i8_min_group(NSL) - i8_frac_a(NSL) - i8_frac_b(NSL));
      ia.sumScale = qScale;
      const int nPad = ia.nSteps * kI8StepAtoms;
      CK(e->phaseTables.reserve((size_t)nPad * ia.XYS + 1024));  // slices may overrun by < 16 entries
      CK(e->zPlanes.reserve((size_t)ia.nSteps * ia.NCB * NSL * planeB + 16));
      const long long total = (long long)nPad * (ia.XYS + ia.cPer * ia.NCB);
      const unsigned tg = (unsigned)((total + 255) / 256);
      if (NSL == 6)
        k_i8_tables<6><<<tg, 256, 0, e->stream>>>(nAt, nPad, ia.KX1, KY1, ia.NCB, ia.cPer,
                                                  ks.cv[0], ks.cv[1], ks.cv[2], 1.0 / qScale,
                                                  bx.packed.p, e->phaseTables.p, e->zPlanes.p);
      else
        k_i8_tables<5><<<tg, 256, 0, e->stream>>>(nAt, nPad, ia.KX1, KY1, ia.NCB, ia.cPer,
                                                  ks.cv[0], ks.cv[1], ks.cv[2], 1.0 / qScale,
                                                  bx.packed.p, e->phaseTables.p, e->zPlanes.p);
      nSlabs = ia.nChunks;
      CK(e->part.reserve((size_t)nSlabs * 2 * nkStride + 64));
      CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)nSlabs * 2 * nkStride, e->stream));
      ia.rows = ks.i8Rows.p;
      ia.tiles = ks.i8Tiles.p;
      ia.tabXY = e->phaseTables.p;
      ia.zPlanes = e->zPlanes.p;
      ia.part = e->part.p;
      const int grid = ks.i8NTiles * ia.nChunks;
      if (NSL == 6) {
        CK(cudaFuncSetAttribute(k_recip_i8<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_recip_i8<6><<<grid, kI8Threads, smem, e->stream>>>(ia);
      } else {
        CK(cudaFuncSetAttribute(k_recip_i8<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_recip_i8<5><<<grid, kI8Threads, smem, e->stream>>>(ia);
      }
      e->launches += 2;
      i8Done = true;
    }
  }
  if (i8Done || nufftDone) {
  } else if (e->recipAlgo >= 2 && ks.mmaValid && nAt > 0) {
    rc = build_mma_tiles(e, ks);  // (re)derives tiles and ZS for the current shard
    if (rc) return rc;
    MmaArgs ma;
    ma.rows = ks.mmaRows.p;
    ma.tiles = ks.mmaTiles.p;
    ma.KX1 = ks.nmax[0] + 1;
    const int KY1 = ks.nmax[1] + 1, KZ1 = ks.nmax[2] + 1;
    ma.XYS = ma.KX1 + KY1;
    ma.ZS = ks.mmaZS;
    size_t budget = e->smemOptin > 8192 ? e->smemOptin - 2048 : 0;
    size_t perAtom = 2 * (size_t)(ma.XYS + ma.ZS) * sizeof(double2) +
                     2 * (size_t)kMmaWarps * kMmaAWS;
    int AT = (int)std::min<size_t>(32, budget / perAtom) & ~3;
    if (AT < 4) return fail(GOMCB200_EINVAL, "k range too large for the DMMA kernel");
    ma.AT = AT;
    ma.nkStride = nkStride;
    rc = build_mma_segments(e, ks, nAt, AT);
    if (rc) return rc;
    ma.segs = ks.mmaSegs.p;
    ma.ctaSeg = ks.mmaCtaSeg.p;
    const int nChunks = (nAt + AT - 1) / AT;
    const int nPad = nChunks * AT;
    const size_t nXY = (size_t)nPad * ma.XYS, nZ = (size_t)nPad * ma.ZS;
    CK(e->phaseTables.reserve(nXY + nZ + 16));
    ma.tabXY = e->phaseTables.p;
    ma.tabZ = e->phaseTables.p + nXY;
    {
      long long total = (long long)nPad * (ma.XYS + ma.ZS);
      k_phase_tables<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(
          nAt, nPad, ma.KX1, KY1, KZ1, ma.ZS, ks.cv[0], ks.cv[1], ks.cv[2], bx.packed.p,
          e->phaseTables.p, e->phaseTables.p + nXY);
    }
    nSlabs = ks.maxSlabs;
    CK(e->part.reserve((size_t)nSlabs * 2 * nkStride + 64));
    CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)nSlabs * 2 * nkStride, e->stream));
    size_t smem = perAtom * AT;
    CK(cudaFuncSetAttribute(k_recip_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const bool dbg = getenv("GOMCB200_DEBUG_MMA") != nullptr;
    DevBuf<long long> dbgBuf;
    ma.ctaCycles = nullptr;
    if (dbg) {
      CK(dbgBuf.reserve(ks.nCtas + 1));
      ma.ctaCycles = dbgBuf.p;
    }
    k_recip_mma<<<ks.nCtas, kMmaThreads, smem, e->stream>>>(ma, e->part.p);
    e->launches += 2;
    if (dbg) {
      std::vector<long long> cyc(ks.nCtas);
      std::vector<int4> hs(ks.hSegs);
      CK(cudaStreamSynchronize(e->stream));
      CK(cudaMemcpy(cyc.data(), dbgBuf.p, sizeof(long long) * ks.nCtas, cudaMemcpyDeviceToHost));
      for (int c = 0; c < ks.nCtas; ++c) {
        fprintf(stderr, "MMADBG cta %d cycles %lld segs", c, cyc[c]);
        for (int sg = ks.hCtaSeg[c]; sg < ks.hCtaSeg[c + 1]; ++sg)
          fprintf(stderr, " [NT=%d chunks=%d]", ks.hMmaTiles[hs[sg].x].z, hs[sg].z - hs[sg].y);
        fprintf(stderr, "\n");
      }
      dbgBuf.release();
    }
  } else if (e->recipAlgo >= 1 && ks.planValid && ks.nTiles > 0 && nAt > 0) {
    FactArgs fa;
    fa.rows = ks.rows.p;
    fa.tiles = ks.tiles.p;
    fa.KX1 = ks.nmax[0] + 1;
    fa.KY1 = ks.nmax[1] + 1;
    fa.KZ1 = ks.nmax[2] + 1;
    fa.ZS = ((fa.KZ1 + kTC - 1) / kTC) * kTC;
    fa.RS = ks.maxRows;
    size_t perAtom = (size_t)(fa.KX1 + fa.KY1 + fa.ZS + fa.RS) * sizeof(double2);
    size_t budget = e->smemOptin > 8192 ? e->smemOptin - 4096 : 0;
    int AT = (int)std::min<size_t>(32, budget / perAtom);
    if (AT < 1)
      return fail(GOMCB200_EINVAL, "k range too large for the factorised kernel");
    fa.AT = AT;
    // multi-GPU: this rank owns tiles [tile0, tile1) (tiles have equal work)
    const int tile0 = (int)(((long long)ks.nTiles * e->shardRank) / e->shardWorld);
    const int tile1 = (int)(((long long)ks.nTiles * (e->shardRank + 1)) / e->shardWorld);
    const int myTiles = std::max(1, tile1 - tile0);
    nSlabs = std::max(1, (e->numSMs + myTiles / 2) / myTiles);
    nSlabs = std::min(nSlabs, std::max(1, nAt / (2 * AT)));
    int per = (nAt + nSlabs - 1) / nSlabs;
    per = ((per + AT - 1) / AT) * AT;
    nSlabs = (nAt + per - 1) / per;
    fa.nAtoms = nAt;
    fa.atomsPerSlab = per;
    fa.nkStride = nkStride;
    fa.cvx = ks.cv[0];
    fa.cvy = ks.cv[1];
    fa.cvz = ks.cv[2];
    CK(e->part.reserve((size_t)nSlabs * 2 * nkStride + 64));
    size_t smem = perAtom * AT;
    CK(cudaFuncSetAttribute(k_recip_fact, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)smem));
    if (e->shardWorld > 1)
      CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)nSlabs * 2 * nkStride,
                         e->stream));
    fa.tile0 = tile0;
    if (tile1 > tile0) {
      dim3 grid(tile1 - tile0, nSlabs);
      k_recip_fact<<<grid, kFactThreads, smem, e->stream>>>(fa, bx.packed.p, e->part.p);
    }
    e->launches += 1;
  } else {
    nSlabs = std::max(1, std::min(64, (4 * e->numSMs + nBlocks - 1) / nBlocks));
    nSlabs = std::min(nSlabs, std::max(1, nAt / 64));
    int per = nAt > 0 ? (nAt + nSlabs - 1) / nSlabs : 1;
    nSlabs = nAt > 0 ? (nAt + per - 1) / per : 1;
    CK(e->part.reserve((size_t)nSlabs * 2 * nkStride + 64));
    // multi-GPU: a contiguous share of the k list per rank, zero elsewhere
    const int k0 = (int)(((long long)nk * e->shardRank) / e->shardWorld);
    const int k1 = (int)(((long long)nk * (e->shardRank + 1)) / e->shardWorld);
    if (e->shardWorld > 1)
      CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)nSlabs * 2 * nkStride,
                         e->stream));
    if (k1 > k0 && nAt > 0) {
      dim3 grid((k1 - k0 + 255) / 256, nSlabs);
      k_recip_direct<<<grid, 256, 0, e->stream>>>(k0, k1, nkStride, nAt, per, bx.packed.p,
                                                 ks.kx.p, ks.ky.p, ks.kz.p, e->part.p);
      e->launches += 1;
    } else if (e->shardWorld == 1) {
      CK(cudaMemsetAsync(e->part.p, 0, sizeof(double) * (size_t)nSlabs * 2 * nkStride,
                         e->stream));
    }
  }
  CK(cudaGetLastError());
  if (e->timing && !e->capturing) cudaEventRecord(e->ev[3], e->stream);
  k_recip_finish<<<nBlocks, 256, 0, e->stream>>>(nk, nkStride, nSlabs, e->part.p,
                                                ks.prefact.p, bx.sum[bx.iRnew].p,
                                                bx.sum[bx.iInew].p, e->blockA.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, 1, e->blockA.p, nullptr, nullptr,
                                           nullptr, e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  if (nufftDone && e->shardWorld > 1 && e->shardRank != 0)  // replicated: counted once
    CK(cudaMemsetAsync(e->result.p, 0, sizeof(double), e->stream));
  bx.sumsComplete = e->shardWorld == 1 || nufftDone;
  return 0;
}

int ensure_mirror(gomcb200_engine *e) {
  if (e->mirrorValid) return 0;
  e->hx.resize(e->nAtoms);
  e->hy.resize(e->nAtoms);
  e->hz.resize(e->nAtoms);
  CK(cudaStreamSynchronize(e->stream));
  size_t bytes = sizeof(double) * (size_t)e->nAtoms;
  CK(cudaMemcpy(e->hx.data(), e->x.p, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(e->hy.data(), e->y.p, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(e->hz.data(), e->z.p, bytes, cudaMemcpyDeviceToHost));
  e->mirrorValid = true;
  return 0;
}

// Stage {len, per atom: qEff, new xyz, old xyz (mode 0) | lambda = 1 charge (swap modes)} into
// the pinned area (at stageOffset) and queue its upload to molBuf.
int stage_molbuf_raw(gomcb200_engine *e, int len, const double *qEff, const double *qTrue,
                     const double *nx, const double *ny, const double *nz, const double *ox,
                     const double *oy, const double *oz, int mode, size_t stageOffset) {
  size_t nd = 1 + 7 * (size_t)len;
  int rc = stage_reserve(e, stageOffset + nd * sizeof(double));
  if (rc) return rc;
  CK(e->molBuf.reserve(nd + 8));
  double *h = reinterpret_cast<double *>(reinterpret_cast<char *>(e->hStage) + stageOffset);
  h[0] = (double)len;
  for (int a = 0; a < len; ++a) {
    double *m = h + 1 + 7 * a;
    m[0] = qEff[a];  // Ewald terms see q * lambdaCoef
    m[1] = nx[a];
    m[2] = ny[a];
    m[3] = nz[a];
    // old coordinates; the swap modes carry the lambda = 1 charge there instead (SwapSelf
    // uses it, src/Ewald.cpp:1375-1391)
    m[4] = mode == 0 ? ox[a] : qTrue[a];
    m[5] = mode == 0 ? oy[a] : 0.0;
    m[6] = mode == 0 ? oz[a] : 0.0;
  }
  CK(cudaMemcpyAsync(e->molBuf.p, h, nd * sizeof(double), cudaMemcpyHostToDevice,
                     e->stream));
  return 0;
}

// the same for the resident molecule molIndex (old coordinates from the host mirror)
int stage_molbuf(gomcb200_engine *e, int molIndex, const double *nx, const double *ny,
                 const double *nz, int mode, size_t stageOffset) {
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  if (mode == 0) {
    int rc = ensure_mirror(e);
    if (rc) return rc;
  }
  return stage_molbuf_raw(e, len, e->hChargeEff.data() + s, e->hCharge.data() + s, nx, ny, nz,
                          mode == 0 ? e->hx.data() + s : nullptr,
                          mode == 0 ? e->hy.data() + s : nullptr,
                          mode == 0 ? e->hz.data() + s : nullptr, mode, stageOffset);
}

// k-space delta of the molecule staged in molBuf (k_mol_recip) on the Ref set; result[0]
int launch_mol_recip(gomcb200_engine *e, int b, int len, int mode) {
  BoxState &bx = e->box[b];
  KSet &ks = bx.kset[1 - bx.cur];
  const int nk = ks.n;
  const int nBlocks = (nk + 255) / 256;
  CK(e->blockA.reserve(nBlocks + 1024));
  k_mol_recip<<<nBlocks, 256, 7 * len * sizeof(double), e->stream>>>(
      nk, mode, e->molBuf.p, ks.kx.p, ks.ky.p, ks.kz.p, ks.prefact.p, bx.sum[bx.iRref].p,
      bx.sum[bx.iIref].p, bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p, e->blockA.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, 1, e->blockA.p, nullptr, nullptr,
                                           nullptr, e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int run_mol_recip(gomcb200_engine *e, int b, int molIndex, const double *nx,
                  const double *ny, const double *nz, int mode, double *out,
                  size_t stageOffset = 0, bool sync = true) {
  BoxState &bx = e->box[b];
  KSet &ks = bx.kset[1 - bx.cur];  // Ref set
  const int nk = ks.n;
  if (nk == 0) {
    *out = 0.0;
    return 0;
  }
  int rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  const int len = e->hMolStart[molIndex + 1] - e->hMolStart[molIndex];
  rc = stage_molbuf(e, molIndex, nx, ny, nz, mode, stageOffset);
  if (rc) return rc;
  rc = launch_mol_recip(e, b, len, mode);
  if (rc) return rc;
  if (!sync) {  // caller synchronises once for several queued operations
    CK(cudaMemcpyAsync(e->hRes, e->result.p, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    return 0;
  }
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *out = e->hRes[0];
  return 0;
}

// sync = false: the caller itself synchronises the stream before it returns to the user
// (whose buffers must stay untouched only that long), so the kernels can be queued behind
// the copies at once.
int upload3(gomcb200_engine *e, DevBuf<double> &dx, DevBuf<double> &dy, DevBuf<double> &dz,
            const double *x, const double *y, const double *z, int first, int count,
            int limit, bool sync = true) {
  if (first < 0 || count < 0 || first + count > limit)
    return fail(GOMCB200_EINVAL, "range [%d,%d) outside [0,%d)", first, first + count, limit);
  if (count == 0) return 0;
  size_t bytes = sizeof(double) * (size_t)count;
  CK(cudaMemcpyAsync(dx.p + first, x, bytes, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(dy.p + first, y, bytes, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(dz.p + first, z, bytes, cudaMemcpyHostToDevice, e->stream));
  if (sync) CK(cudaStreamSynchronize(e->stream));  // caller may reuse pageable buffers
  return 0;
}

void mark_coords_dirty(gomcb200_engine *e) {
  for (auto &bx : e->box) {
    bx.cellsDirty = true;
    bx.packedDirty = true;
  }
}

template <int VDW>
void launch_probe(gomcb200_engine *e, int b, const BoxParams &p, int excludeMol, int n) {
  BoxState &bx = e->box[b];
  if (bx.lambdaMol >= 0)
    k_probe<VDW, MODE_LAMBDA><<<n, kPairThreads, 0, e->stream>>>(
        p, bx.grid, excludeMol, e->probes.p, bx.cellStart.p, bx.sx.p, bx.sy.p, bx.sz.p, bx.sq.p,
        bx.skm.p, e->probeOut.p);
  else
    k_probe<VDW><<<n, kPairThreads, 0, e->stream>>>(p, bx.grid, excludeMol, e->probes.p,
                                                   bx.cellStart.p, bx.sx.p, bx.sy.p, bx.sz.p,
                                                   bx.sq.p, bx.skm.p, e->probeOut.p);
}

// probes staged in e->hStage as Probe[n]; results in e->hStage after the call
int run_probes(gomcb200_engine *e, int b, int excludeMol, int n, std::vector<double> &out,
               bool sync = true) {
  int rc = ensure_cells(e, b);
  if (rc) return rc;
  BoxParams p = make_params(e, b);
  CK(e->probes.reserve(n + 1));
  CK(e->probeOut.reserve(3 * (size_t)n + 8));
  CK(cudaMemcpyAsync(e->probes.p, e->hStage, sizeof(Probe) * (size_t)n,
                     cudaMemcpyHostToDevice, e->stream));
  if (e->vdwKind == VDW_SHIFT)
    launch_probe<VDW_SHIFT>(e, b, p, excludeMol, n);
  else if (e->vdwKind == VDW_SWITCH)
    launch_probe<VDW_SWITCH>(e, b, p, excludeMol, n);
  else if (e->vdwKind == VDW_EXP6)
    launch_probe<VDW_EXP6>(e, b, p, excludeMol, n);
  else if (e->vdwKind == VDW_MARTINI)
    launch_probe<VDW_MARTINI>(e, b, p, excludeMol, n);
  else
    launch_probe<VDW_STD>(e, b, p, excludeMol, n);
  e->launches += 1;
  CK(cudaGetLastError());
  // results come back through the pinned staging area, right after the probes
  double *hOut = reinterpret_cast<double *>(reinterpret_cast<char *>(e->hStage) +
                                            sizeof(Probe) * (size_t)n);
  CK(cudaMemcpyAsync(hOut, e->probeOut.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost,
                     e->stream));
  if (!sync) return 0;
  CK(cudaStreamSynchronize(e->stream));
  out.assign(hOut, hOut + 3 * (size_t)n);
  return 0;
}

template <int VDW>
void launch_trial(gomcb200_engine *e, int b, const BoxParams &p, const TrialArgs &a, int nBlocks) {
  BoxState &bx = e->box[b];
  KSet &ks = bx.kset[1 - bx.cur];
  volatile unsigned long long *flag =
      reinterpret_cast<volatile unsigned long long *>(e->dTrial + 8);
  if (bx.lambdaMol >= 0)
    k_mol_trial<VDW, MODE_LAMBDA><<<nBlocks, kPairThreads, 0, e->stream>>>(
        p, bx.grid, a, bx.cellStart.p, bx.sx.p, bx.sy.p, bx.sz.p, bx.sq.p, bx.skm.p, ks.kx.p,
        ks.ky.p, ks.kz.p, ks.prefact.p, bx.sum[bx.iRref].p, bx.sum[bx.iIref].p,
        bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p, e->trialPart.p, e->blockA.p, e->ticket.p,
        e->dTrial, flag);
  else
    k_mol_trial<VDW><<<nBlocks, kPairThreads, 0, e->stream>>>(
        p, bx.grid, a, bx.cellStart.p, bx.sx.p, bx.sy.p, bx.sz.p, bx.sq.p, bx.skm.p, ks.kx.p,
        ks.ky.p, ks.kz.p, ks.prefact.p, bx.sum[bx.iRref].p, bx.sum[bx.iIref].p,
        bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p, e->trialPart.p, e->blockA.p, e->ticket.p,
        e->dTrial, flag);
}

// One-launch single-molecule trial (trial.cuh), molecules of <= kTrialMaxAtoms atoms.
// out = {lj, real, overlap, recipNew, correction, self}
int run_trial_fused(gomcb200_engine *e, int b, int molIndex, const double *nx, const double *ny,
                    const double *nz, int mode, bool probes, bool recip, bool correction,
                    double out[6]) {
  BoxState &bx = e->box[b];
  int rc = 0;
  if (probes) {
    rc = ensure_cells(e, b);
    if (rc) return rc;
  }
  const int nk = recip ? bx.kset[1 - bx.cur].n : 0;
  if (nk) {
    rc = ensure_sums(e, bx, nk);
    if (rc) return rc;
  }
  if (mode == 0) {
    rc = ensure_mirror(e);
    if (rc) return rc;
  }
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  TrialArgs a;
  a.len = len;
  a.mode = mode;
  a.excludeMol = molIndex;
  a.nk = nk;
  a.nProbeBlocks = probes ? 2 * len * kProbeSplit : 0;
  a.doCorrection = correction ? 1 : 0;
  a.seq = ++e->trialSeq;
  for (int i = 0; i < len; ++i) {
    a.kind[i] = e->hKind[s + i];
    a.q[i] = e->hCharge[s + i];
    a.qr[i] = e->hChargeEff[s + i];
    a.nx[i] = nx[i];
    a.ny[i] = ny[i];
    a.nz[i] = nz[i];
    a.ox[i] = mode == 0 ? e->hx[s + i] : 0.0;
    a.oy[i] = mode == 0 ? e->hy[s + i] : 0.0;
    a.oz[i] = mode == 0 ? e->hz[s + i] : 0.0;
  }
  int nRecipBlocks = (nk + kPairThreads - 1) / kPairThreads;
  if (a.nProbeBlocks + nRecipBlocks == 0) nRecipBlocks = 1;  // the finalising block
  CK(e->blockA.reserve(nRecipBlocks + 1024));
  const int nBlocks = a.nProbeBlocks + nRecipBlocks;
  BoxParams p = make_params(e, b);
  if (e->vdwKind == VDW_SHIFT)
    launch_trial<VDW_SHIFT>(e, b, p, a, nBlocks);
  else if (e->vdwKind == VDW_SWITCH)
    launch_trial<VDW_SWITCH>(e, b, p, a, nBlocks);
  else if (e->vdwKind == VDW_EXP6)
    launch_trial<VDW_EXP6>(e, b, p, a, nBlocks);
  else if (e->vdwKind == VDW_MARTINI)
    launch_trial<VDW_MARTINI>(e, b, p, a, nBlocks);
  else
    launch_trial<VDW_STD>(e, b, p, a, nBlocks);
  e->launches += 1;
  CK(cudaGetLastError());
  // the kernel publishes a.seq in mapped host memory after its results
  volatile unsigned long long *flag =
      reinterpret_cast<volatile unsigned long long *>(e->hTrial + 8);
  for (unsigned spin = 1; *flag != a.seq; ++spin) {
    if ((spin & 0xfffffu) == 0) {  // every ~1M polls make sure the launch is still alive
      cudaError_t q = cudaStreamQuery(e->stream);
      if (q != cudaErrorNotReady) {
        if (q == cudaSuccess && *flag == a.seq) break;
        return fail(GOMCB200_ECUDA, "single-molecule trial kernel: %s",
                    q == cudaSuccess ? "finished without publishing its result"
                                     : cudaGetErrorString(q));
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  for (int i = 0; i < 6; ++i) out[i] = e->hTrial[i];
  return 0;
}

void timing_begin(gomcb200_engine *e) {
  if (e->timing) cudaEventRecord(e->ev[0], e->stream);
}
void timing_end(gomcb200_engine *e, bool recip) {
  if (!e->timing) return;
  cudaEventRecord(e->ev[1], e->stream);
  cudaEventSynchronize(e->ev[1]);
  cudaEventElapsedTime(&e->lastTotalMs, e->ev[0], e->ev[1]);
  e->lastDominantMs = 0;
  if (recip) cudaEventElapsedTime(&e->lastDominantMs, e->ev[2], e->ev[3]);
}

// Everything one full-box evaluation queues (both streams, the collectives, the D2H copy of
// the three energies into hRes[8..10]) -- the unit that is captured as a CUDA graph.
int enqueue_full_box(gomcb200_engine *e, int box, bool recipOn) {
  BoxState &bx = e->box[box];
  int rc = 0;
  const bool fork = recipOn && e->overlap;
  if (fork) {
    // structure factor on the second stream, behind the coordinates that are on the first:
    // its small set-up kernels, and on a sharded engine the all-gather of the FFT slabs, run
    // under the cell binning and the pair sweep.  The stream and scratch members are
    // exchanged for the duration of the call so that every helper below queues there.
    CK(cudaEventRecord(e->evFork, e->stream));
    CK(cudaStreamWaitEvent(e->stream2, e->evFork, 0));
    std::swap(e->stream, e->stream2);
    std::swap(e->blockA, e->blockA2);
    std::swap(e->result, e->result2);
    rc = run_recip_sums(e, box, bx.kset[1 - bx.cur]);
    cudaError_t ce = cudaSuccess;
    if (!rc) ce = cudaEventRecord(e->evJoin, e->stream);
    std::swap(e->stream, e->stream2);
    std::swap(e->blockA, e->blockA2);
    std::swap(e->result, e->result2);
    if (rc) return rc;
    CK(ce);
  }
  rc = run_pair(e, box, false);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->energy3.p, e->result.p, 2 * sizeof(double), cudaMemcpyDeviceToDevice,
                     e->stream));
  if (fork) {
    CK(cudaStreamWaitEvent(e->stream, e->evJoin, 0));
    CK(cudaMemcpyAsync(e->energy3.p + 2, e->result2.p, sizeof(double), cudaMemcpyDeviceToDevice,
                       e->stream));
  } else if (recipOn) {
    rc = run_recip_sums(e, box, bx.kset[1 - bx.cur]);
    if (rc) return rc;
    CK(cudaMemcpyAsync(e->energy3.p + 2, e->result.p, sizeof(double), cudaMemcpyDeviceToDevice,
                       e->stream));
  } else {
    CK(cudaMemsetAsync(e->energy3.p + 2, 0, sizeof(double), e->stream));
  }
  // sharded engine with a communicator: the path's only reduction, three doubles, on the
  // engine's stream; every rank returns the complete energies
  rc = allreduce_energies(e, e->energy3.p, 3);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->hRes + 8, e->energy3.p, 3 * sizeof(double), cudaMemcpyDeviceToHost,
                     e->stream));
  return 0;
}

// What a captured step bakes in besides the (stable) buffer addresses: if any of it differs,
// the graph is dropped and the step re-captured.
void step_key(gomcb200_engine *e, int box, std::vector<unsigned char> &key) {
  BoxState &bx = e->box[box];
  const BoxParams p = make_params(e, box);
  const KSet &ks = bx.kset[1 - bx.cur];
  const long long ints[] = {box, bx.nAtoms, bx.nMols, bx.nCharged, bx.cur, bx.iRnew, bx.iInew,
                            ks.n, (long long)(size_t)ks.kx.p, (long long)(size_t)ks.prefact.p,
                            e->recipAlgo, e->pairAlgo, e->shardRank, e->shardWorld,
                            (long long)(size_t)e->comm, e->overlap, e->timing, bx.nonOrth,
                            ks.planValid, ks.ngValid, ks.mmaValid,
                            g_devBufGeneration.load(), gbn::nufft_alloc_generation()};
  key.resize(sizeof(p) + sizeof(ints) + sizeof(bx.axis));
  std::memcpy(key.data(), &p, sizeof(p));
  std::memcpy(key.data() + sizeof(p), ints, sizeof(ints));
  std::memcpy(key.data() + sizeof(p) + sizeof(ints), bx.axis, sizeof(bx.axis));
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

const char *gomcb200_last_error(void) { return g_lastError.c_str(); }
int gomcb200_version(void) { return 100; }

int gomcb200_create(gomcb200_engine **out, int device, int nBoxes) {
  if (!out || nBoxes < 1 || nBoxes > 2) return fail(GOMCB200_EINVAL, "bad arguments");
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GOMCB200_ENODEV, "no CUDA device: %s (this engine has no CPU fallback)",
                ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
  if (device < 0) {
    CK(cudaGetDevice(&device));
  }
  if (device >= count) return fail(GOMCB200_ENODEV, "device %d of %d", device, count);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(GOMCB200_ENODEV, "device %d is sm_%d%d; this build is sm_100a only", device,
                prop.major, prop.minor);
  gomcb200_engine *e = new gomcb200_engine();
  e->device = device;
  e->numSMs = prop.multiProcessorCount;
  e->smemOptin = prop.sharedMemPerBlockOptin;
  e->nBoxes = nBoxes;
  e->box.resize(nBoxes);
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CK(cudaHostAlloc(&e->hRes, 64 * sizeof(double), cudaHostAllocDefault));
  CK(e->result.reserve(64));
  CK(cudaHostAlloc(&e->hTrial, 16 * sizeof(double), cudaHostAllocMapped));
  memset(e->hTrial, 0, 16 * sizeof(double));
  CK(cudaHostGetDevicePointer(&e->dTrial, e->hTrial, 0));
  CK(e->ticket.reserve(4));
  CK(cudaMemset(e->ticket.p, 0, 4 * sizeof(unsigned)));
  CK(e->trialPart.reserve(3 * 2 * kTrialMaxAtoms * kProbeSplit + 8));
  for (auto &ev : e->ev) CK(cudaEventCreate(&ev));
  CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&e->evJoin, cudaEventDisableTiming));
  CK(e->result2.reserve(64));
  if (const char *ev = getenv("GOMCB200_OVERLAP")) e->overlap = atoi(ev) != 0;
  if (const char *ev = getenv("GOMCB200_GRAPH")) e->useGraph = atoi(ev) != 0;
  e->nufft = gbn::nufft_create();
  *out = e;
  return 0;
}

int gomcb200_destroy(gomcb200_engine *e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  // every DevBuf member releases its memory in the engine's destructor (delete below)
  if (e->hRes) cudaFreeHost(e->hRes);
  if (e->hTrial) cudaFreeHost(e->hTrial);
  if (e->hStage) cudaFreeHost(e->hStage);
  for (auto &ev : e->ev) cudaEventDestroy(ev);
  gbn::nufft_destroy(e->nufft);
  gbc::comm_destroy(e->comm);
  for (auto &g : e->stepGraph)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->evFork) cudaEventDestroy(e->evFork);
  if (e->evJoin) cudaEventDestroy(e->evJoin);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  cudaStreamDestroy(e->stream);
  delete e;
  return 0;
}

long long gomcb200_launch_count(const gomcb200_engine *e) { return e ? e->launches : 0; }

int gomcb200_init_forcefield(gomcb200_engine *e, const double *sigmaSq,
                             const double *epsilon_cn, const double *n, int vdwKind,
                             int isMartini, int count, double rCut,
                             const double *rCutCoulomb, double rCutLow, double rOn,
                             const double *alpha, int ewald, int electrostatic,
                             double diElectric_1) {
  if (!e || !sigmaSq || !epsilon_cn || !n || count < 1)
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (vdwKind < 0 || vdwKind > 3) return fail(GOMCB200_EINVAL, "unknown vdwKind %d", vdwKind);
  if (isMartini && vdwKind != VDW_SWITCH)
    return fail(GOMCB200_EINVAL, "Martini parameters need Potential SWITCH");
  CK(cudaSetDevice(e->device));
  // Forcefield::Init picks FF_SWITCH_MARTINI for SWITCH + Martini (src/Forcefield.cpp:100-103)
  e->vdwKind = (isMartini && vdwKind == VDW_SWITCH) ? VDW_MARTINI : vdwKind;
  vdwKind = e->vdwKind;
  e->diElectric_1 = diElectric_1;
  e->haveExp6 = false;
  e->ewald = ewald;
  e->electrostatic = electrostatic;
  e->kindCount = count;
  e->rCut = rCut;
  e->rCutLow = rCutLow;
  e->rOn = rOn;
  e->rCutCoulomb.assign(e->nBoxes, rCut);
  e->alpha.assign(e->nBoxes, 0.0);
  e->recipRcut.assign(e->nBoxes, 0.0);
  for (int b = 0; b < e->nBoxes; ++b) {
    if (rCutCoulomb) e->rCutCoulomb[b] = rCutCoulomb[b];
    if (alpha) e->alpha[b] = alpha[b];
  }
  size_t sz = (size_t)count * count;
  std::vector<double> shift(sz, 0.0);
  std::vector<int> nHalf(sz, 0);
  for (size_t i = 0; i < sz; ++i) {
    double h = n[i] * 0.5;
    if (h == std::floor(h) && h >= 1.0 && h <= 64.0) nHalf[i] = (int)h;
    if (vdwKind == VDW_SHIFT) {  // FF_SHIFT::Init, src/FFShift.h:100-116
      double rRat2 = sigmaSq[i] / (rCut * rCut);
      double rRat4 = rRat2 * rRat2;
      double attract = rRat4 * rRat2;
      double repulse = pow(sqrt(rRat2), n[i]);
      shift[i] = epsilon_cn[i] * (repulse - attract);
    }
  }
  if (vdwKind == VDW_MARTINI) {  // FF_SWITCH_MARTINI::Init, src/FFSwitchMartini.h:121-213
    const double rOnCoul = 0.0;
    const double d = rCut - rOn, dc = rCut - rOnCoul;
    e->mA6 = 6.0 * (7.0 * rOn - 10.0 * rCut) / (pow(rCut, 8.0) * d * d);
    e->mB6 = -6.0 * (7.0 * rOn - 9.0 * rCut) / (pow(rCut, 8.0) * d * d * d);
    e->mC6 = pow(rCut, -6.0) - e->mA6 / 3.0 * d * d * d - e->mB6 / 4.0 * d * d * d * d;
    e->mA1 = (2.0 * rOnCoul - 5.0 * rCut) / (rCut * rCut * rCut * dc * dc);
    e->mB1 = -1.0 * (2.0 * rOnCoul - 4.0 * rCut) / (rCut * rCut * rCut * dc * dc * dc);
    e->mC1 = 1.0 / rCut - e->mA1 / 3.0 * dc * dc * dc - e->mB1 / 4.0 * dc * dc * dc * dc;
    std::vector<double> An(sz), Bn(sz), Cn(sz), sg(sz), s6(sz);
    for (size_t i = 0; i < sz; ++i) {
      double pn = n[i];
      An[i] = pn * ((pn + 1.0) * rOn - (pn + 4.0) * rCut) / (pow(rCut, pn + 2.0) * d * d);
      Bn[i] = -pn * ((pn + 1.0) * rOn - (pn + 3.0) * rCut) / (pow(rCut, pn + 2.0) * d * d * d);
      Cn[i] = 1.0 / pow(rCut, pn) - An[i] / 3.0 * d * d * d - Bn[i] / 4.0 * d * d * d * d;
      double sigma = sqrt(sigmaSq[i]);
      s6[i] = pow(sigma, 6.0);
      sg[i] = pow(sigma, pn);
    }
    CK(e->mAn.reserve(sz)); CK(e->mBn.reserve(sz)); CK(e->mCn.reserve(sz));
    CK(e->mSign.reserve(sz)); CK(e->mSig6.reserve(sz));
    CK(cudaMemcpy(e->mAn.p, An.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->mBn.p, Bn.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->mCn.p, Cn.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->mSign.p, sg.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->mSig6.p, s6.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
  }
  CK(e->sigmaSq.reserve(sz)); CK(e->epsilon_cn.reserve(sz)); CK(e->nTab.reserve(sz));
  CK(e->shiftConst.reserve(sz)); CK(e->nHalf.reserve(sz));
  CK(cudaMemcpy(e->sigmaSq.p, sigmaSq, sz * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->epsilon_cn.p, epsilon_cn, sz * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->nTab.p, n, sz * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->shiftConst.p, shift.data(), sz * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->nHalf.p, nHalf.data(), sz * sizeof(int), cudaMemcpyHostToDevice));
  e->haveFF = true;
  return 0;
}

int gomcb200_init_softcore(gomcb200_engine *e, double sc_alpha, double sc_sigma_6, int sc_power,
                           int sc_coul) {
  if (!e) return fail(GOMCB200_EINVAL, "bad arguments");
  e->scAlpha = sc_alpha;
  e->scSigma6 = sc_sigma_6;
  e->scPower = (double)sc_power;
  e->scCoul = sc_coul ? 1 : 0;
  return 0;
}

int gomcb200_update_lambda(gomcb200_engine *e, int box, int molIndex, int molKind,
                           double lambdaVDW, double lambdaCoulomb, int isFraction) {
  if (!e || box < 0 || box >= e->nBoxes || !e->haveTopo)
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (isFraction && (molIndex < 0 || molIndex >= e->nMols))
    return fail(GOMCB200_EINVAL, "molecule index %d out of range", molIndex);
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  auto set_eff = [&](int m, double coef) -> int {
    const int s = e->hMolStart[m], len = e->hMolStart[m + 1] - s;
    for (int a = s; a < s + len; ++a) e->hChargeEff[a] = e->hCharge[a] * coef;
    CK(cudaMemcpyAsync(e->qEff.p + s, e->hChargeEff.data() + s, sizeof(double) * (size_t)len,
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
  };
  if (bx.lambdaMol >= 0) {
    int rc = set_eff(bx.lambdaMol, 1.0);
    if (rc) return rc;
  }
  if (isFraction) {
    bx.lambdaMol = molIndex;
    bx.lambdaMolKind = molKind;
    bx.lambdaVDW = lambdaVDW;
    bx.lambdaCoulomb = lambdaCoulomb;
    int rc = set_eff(molIndex, std::sqrt(lambdaCoulomb));  // Ewald::GetLambdaCoef
    if (rc) return rc;
  } else {
    bx.lambdaMol = bx.lambdaMolKind = -1;
    bx.lambdaVDW = bx.lambdaCoulomb = 1.0;
  }
  bx.packedDirty = true;
  return 0;
}

int gomcb200_init_exp6(gomcb200_engine *e, const double *rMin, const double *expConst,
                       const double *rMaxSq, int size) {
  if (!e || !rMin || !expConst || !rMaxSq) return fail(GOMCB200_EINVAL, "bad arguments");
  if (!e->haveFF) return fail(GOMCB200_EINVAL, "gomcb200_init_forcefield not called");
  if (size != e->kindCount * e->kindCount)
    return fail(GOMCB200_EINVAL, "size %d != count^2 %d", size, e->kindCount * e->kindCount);
  CK(cudaSetDevice(e->device));
  CK(e->rMin.reserve(size)); CK(e->expConst.reserve(size)); CK(e->rMaxSq.reserve(size));
  CK(cudaMemcpy(e->rMin.p, rMin, size * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->expConst.p, expConst, size * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->rMaxSq.p, rMaxSq, size * sizeof(double), cudaMemcpyHostToDevice));
  e->haveExp6 = true;
  return 0;
}

int gomcb200_init_topology(gomcb200_engine *e, int nAtoms, int nMols, const int *particleKind,
                           const int *particleMol, const double *particleCharge,
                           const int *molStart) {
  if (!e || nAtoms < 0 || nMols < 0 || !particleKind || !particleMol || !particleCharge ||
      !molStart)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  e->nAtoms = nAtoms;
  e->nMols = nMols;
  e->hKind.assign(particleKind, particleKind + nAtoms);
  e->hMol.assign(particleMol, particleMol + nAtoms);
  e->hCharge.assign(particleCharge, particleCharge + nAtoms);
  e->hChargeEff = e->hCharge;
  for (auto &bx : e->box) {
    bx.lambdaMol = bx.lambdaMolKind = -1;
    bx.lambdaVDW = bx.lambdaCoulomb = 1.0;
  }
  e->hMolStart.assign(molStart, molStart + nMols + 1);
  e->maxMolLen = 0;
  for (int m = 0; m < nMols; ++m)
    e->maxMolLen = std::max(e->maxMolLen, molStart[m + 1] - molStart[m]);
  size_t na = (size_t)nAtoms + 1, nm = (size_t)nMols + 1;
  CK(e->kind.reserve(na)); CK(e->mol.reserve(na)); CK(e->q.reserve(na));
  CK(e->molStart.reserve(nm + 1));
  CK(e->x.reserve(na)); CK(e->y.reserve(na)); CK(e->z.reserve(na));
  CK(e->comx.reserve(nm)); CK(e->comy.reserve(nm)); CK(e->comz.reserve(nm));
  for (int w = 0; w < 5; ++w) {
    size_t n = (w == GOMCB200_ATOM_FORCE || w == GOMCB200_ATOM_FORCE_REC) ? na : nm;
    for (int c = 0; c < 3; ++c) {
      CK(e->force[w][c].reserve(n));
      CK(cudaMemset(e->force[w][c].p, 0, e->force[w][c].cap * sizeof(double)));
    }
  }
  e->trialActive = false;
  CK(cudaMemcpy(e->kind.p, particleKind, nAtoms * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->mol.p, particleMol, nAtoms * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->q.p, particleCharge, nAtoms * sizeof(double), cudaMemcpyHostToDevice));
  CK(e->qEff.reserve(na));
  CK(cudaMemcpy(e->qEff.p, particleCharge, nAtoms * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->molStart.p, molStart, (nMols + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemset(e->comx.p, 0, e->comx.cap * sizeof(double)));
  CK(cudaMemset(e->comy.p, 0, e->comy.cap * sizeof(double)));
  CK(cudaMemset(e->comz.p, 0, e->comz.cap * sizeof(double)));
  e->haveTopo = true;
  return 0;
}

int gomcb200_set_box_molecules(gomcb200_engine *e, int box, const int *molIndices,
                               int nMolsInBox) {
  if (!e || box < 0 || box >= e->nBoxes || nMolsInBox < 0 || (!molIndices && nMolsInBox))
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (!e->haveTopo) return fail(GOMCB200_EINVAL, "gomcb200_init_topology not called");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  bx.hMols.assign(molIndices, molIndices + nMolsInBox);
  if ((int)e->hMolBox.size() != e->nMols) e->hMolBox.assign(e->nMols, -1);
  for (int &mb : e->hMolBox)
    if (mb == box) mb = -1;
  bx.hAtoms.clear();
  bx.hCharged.clear();
  bx.qMaxAbs = 0.0;
  for (int m : bx.hMols) {
    if (m < 0 || m >= e->nMols) return fail(GOMCB200_EINVAL, "molecule index %d out of range", m);
    e->hMolBox[m] = box;
    for (int a = e->hMolStart[m]; a < e->hMolStart[m + 1]; ++a) {
      bx.hAtoms.push_back(a);
      // Ewald::Init particleHasNoCharge, src/Ewald.cpp:107-111
      if (!(std::fabs(e->hCharge[a]) < 0.000000001)) {
        bx.hCharged.push_back(a);
        bx.qMaxAbs = std::max(bx.qMaxAbs, std::fabs(e->hCharge[a]));
      }
    }
  }
  // ascending atom order: stable radix sort then yields ascending-within-cell
  // order, the order GetCellListNeighbor sorts to (src/CellList.cpp:276)
  std::sort(bx.hAtoms.begin(), bx.hAtoms.end());
  bx.nMols = nMolsInBox;
  bx.nAtoms = (int)bx.hAtoms.size();
  bx.nCharged = (int)bx.hCharged.size();
  CK(bx.molList.reserve(bx.nMols + 1));
  CK(bx.atomList.reserve(bx.nAtoms + 1));
  CK(bx.chargedList.reserve(bx.nCharged + 1));
  if (bx.nMols)
    CK(cudaMemcpy(bx.molList.p, bx.hMols.data(), bx.nMols * sizeof(int), cudaMemcpyHostToDevice));
  if (bx.nAtoms)
    CK(cudaMemcpy(bx.atomList.p, bx.hAtoms.data(), bx.nAtoms * sizeof(int),
                  cudaMemcpyHostToDevice));
  if (bx.nCharged)
    CK(cudaMemcpy(bx.chargedList.p, bx.hCharged.data(), bx.nCharged * sizeof(int),
                  cudaMemcpyHostToDevice));
  bx.cellsDirty = true;
  bx.packedDirty = true;
  return 0;
}

int gomcb200_set_box_axes(gomcb200_engine *e, int box, const double axis[3]) {
  if (!e || box < 0 || box >= e->nBoxes || !axis) return fail(GOMCB200_EINVAL, "bad arguments");
  for (int d = 0; d < 3; ++d) {
    if (!(axis[d] > 0.0)) return fail(GOMCB200_EINVAL, "axis %d is not positive", d);
    e->box[box].axis[d] = axis[d];
  }
  e->box[box].haveAxes = true;
  e->box[box].cellsDirty = true;
  return 0;
}

int gomcb200_set_box_cell_basis(gomcb200_engine *e, int box, const double cellBasis[9],
                                const double cellBasisInv[9], const double axis[3]) {
  if (!e || box < 0 || box >= e->nBoxes || !axis) return fail(GOMCB200_EINVAL, "bad arguments");
  int rc = gomcb200_set_box_axes(e, box, axis);
  if (rc) return rc;
  BoxState &bx = e->box[box];
  bx.nonOrth = cellBasis != nullptr;
  if (cellBasis) {
    if (!cellBasisInv) return fail(GOMCB200_EINVAL, "cellBasisInv missing");
    for (int i = 0; i < 9; ++i) {
      bx.B[i] = cellBasis[i];
      bx.Bi[i] = cellBasisInv[i];
    }
  }
  return 0;
}

static int set_coords_impl(gomcb200_engine *e, const double *x, const double *y, const double *z,
                           int first, int count, bool sync) {
  if (!e || !e->haveTopo) return fail(GOMCB200_EINVAL, "topology not initialised");
  CK(cudaSetDevice(e->device));
  int rc = upload3(e, e->x, e->y, e->z, x, y, z, first, count, e->nAtoms, sync);
  if (rc) return rc;
  if (count > 4096) {
    e->mirrorValid = false;  // refreshed lazily by the next single-molecule call
  } else if (e->mirrorValid) {
    memcpy(e->hx.data() + first, x, sizeof(double) * (size_t)count);
    memcpy(e->hy.data() + first, y, sizeof(double) * (size_t)count);
    memcpy(e->hz.data() + first, z, sizeof(double) * (size_t)count);
  }
  mark_coords_dirty(e);
  return 0;
}

int gomcb200_set_coords(gomcb200_engine *e, const double *x, const double *y, const double *z,
                        int first, int count) {
  return set_coords_impl(e, x, y, z, first, count, true);
}

int gomcb200_get_coords(gomcb200_engine *e, double *x, double *y, double *z, int first,
                        int count) {
  if (!e || !e->haveTopo || first < 0 || count < 0 || first + count > e->nAtoms)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  size_t bytes = sizeof(double) * (size_t)count;
  CK(cudaMemcpy(x, e->x.p + first, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(y, e->y.p + first, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(z, e->z.p + first, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int gomcb200_set_com(gomcb200_engine *e, const double *x, const double *y, const double *z,
                     int first, int count) {
  if (!e || !e->haveTopo) return fail(GOMCB200_EINVAL, "topology not initialised");
  CK(cudaSetDevice(e->device));
  return upload3(e, e->comx, e->comy, e->comz, x, y, z, first, count, e->nMols);
}

int gomcb200_get_com(gomcb200_engine *e, double *x, double *y, double *z, int first, int count) {
  if (!e || !e->haveTopo || first < 0 || count < 0 || first + count > e->nMols)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  size_t bytes = sizeof(double) * (size_t)count;
  if (x) CK(cudaMemcpy(x, e->comx.p + first, bytes, cudaMemcpyDeviceToHost));
  if (y) CK(cudaMemcpy(y, e->comy.p + first, bytes, cudaMemcpyDeviceToHost));
  if (z) CK(cudaMemcpy(z, e->comz.p + first, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int gomcb200_set_molecule_coords(gomcb200_engine *e, int molIndex, const double *x,
                                 const double *y, const double *z, const double com[3]) {
  if (!e || !e->haveTopo || molIndex < 0 || molIndex >= e->nMols)
    return fail(GOMCB200_EINVAL, "bad arguments");
  int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  if (len <= kTrialMaxAtoms && x && y && z) {
    // one parameter-carried launch, no copy, no synchronisation (trial.cuh)
    CK(cudaSetDevice(e->device));
    const int b = (int)e->hMolBox.size() == e->nMols ? e->hMolBox[molIndex] : -1;
    AcceptArgs a;
    a.len = len;
    a.first = s;
    a.molIndex = molIndex;
    a.hasCom = com ? 1 : 0;
    a.inPlace = 0;
    for (int d = 0; d < 3; ++d) a.com[d] = com ? com[d] : 0.0;
    if (b >= 0 && !e->box[b].cellsDirty && !e->box[b].nonOrth && e->mirrorValid) {
      const CellGrid &g = e->box[b].grid;
      a.inPlace = 1;
      for (int i = 0; i < len && a.inPlace; ++i)
        if (position_to_cell(g, x[i], y[i], z[i]) !=
            position_to_cell(g, e->hx[s + i], e->hy[s + i], e->hz[s + i]))
          a.inPlace = 0;
    }
    for (int i = 0; i < len; ++i) {
      a.x[i] = x[i];
      a.y[i] = y[i];
      a.z[i] = z[i];
    }
    BoxState &bs = e->box[b >= 0 ? b : 0];
    k_accept_mol<<<1, kTrialMaxAtoms, 0, e->stream>>>(a, e->x.p, e->y.p, e->z.p, e->comx.p,
                                                     e->comy.p, e->comz.p, bs.sortedPos.p,
                                                     bs.sx.p, bs.sy.p, bs.sz.p);
    e->launches += 1;
    CK(cudaGetLastError());
    if (e->mirrorValid) {
      memcpy(e->hx.data() + s, x, sizeof(double) * (size_t)len);
      memcpy(e->hy.data() + s, y, sizeof(double) * (size_t)len);
      memcpy(e->hz.data() + s, z, sizeof(double) * (size_t)len);
    }
    if (b >= 0) {
      e->box[b].packedDirty = true;
      if (!a.inPlace) e->box[b].cellsDirty = true;
    } else {
      mark_coords_dirty(e);
    }
    return 0;
  }
  int rc = gomcb200_set_coords(e, x, y, z, s, len);
  if (rc) return rc;
  if (com) rc = gomcb200_set_com(e, com, com + 1, com + 2, molIndex, 1);
  return rc;
}

// ---- pair path --------------------------------------------------------------
int gomcb200_box_inter(gomcb200_engine *e, int box, double *LJEn, double *REn) {
  GB_RANGE("energy_box_inter");
  int rc = check_box(e, box);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  timing_begin(e);
  rc = run_pair(e, box, false);
  if (rc) return rc;
  rc = allreduce_energies(e, e->result.p, 2);
  if (rc) return rc;
  rc = fetch_result(e, 2);
  if (rc) return rc;
  timing_end(e, false);
  if (LJEn) *LJEn = e->hRes[0];
  if (REn) *REn = e->hRes[1];
  return 0;
}

int gomcb200_box_force(gomcb200_engine *e, int box, double *LJEn, double *REn) {
  GB_RANGE("energy_box_force");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_box_force", true);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  timing_begin(e);
  rc = run_pair(e, box, true);
  if (rc) return rc;
  rc = allreduce_energies(e, e->result.p, 2);
  if (rc) return rc;
  rc = fetch_result(e, 2);
  if (rc) return rc;
  timing_end(e, false);
  if (LJEn) *LJEn = e->hRes[0];
  if (REn) *REn = e->hRes[1];
  return 0;
}

int gomcb200_molecule_inter(gomcb200_engine *e, int box, int molIndex, const double *newX,
                            const double *newY, const double *newZ, double *dLJ,
                            double *dReal, int *overlap) {
  GB_RANGE("energy_molecule_inter");
  int rc = check_box(e, box);
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !newX || !newY || !newZ)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  if (len <= kTrialMaxAtoms) {
    double r[6];
    rc = run_trial_fused(e, box, molIndex, newX, newY, newZ, 0, true, false, false, r);
    if (rc) return rc;
    if (dLJ) *dLJ = r[0];
    if (dReal) *dReal = r[1];
    if (overlap) *overlap = r[2] != 0.0;
    return 0;
  }
  rc = stage_reserve(e, (sizeof(Probe) + 3 * sizeof(double)) * 2 * (size_t)len + 64);
  if (rc) return rc;
  rc = ensure_mirror(e);  // old coordinates of the molecule
  if (rc) return rc;
  const double *ox = e->hx.data() + s, *oy = e->hy.data() + s, *oz = e->hz.data() + s;
  Probe *pr = reinterpret_cast<Probe *>(e->hStage);
  for (int a = 0; a < len; ++a) {  // order of src/CalculateEnergy.cpp:590-678
    Probe o = {ox[a], oy[a], oz[a], e->hCharge[s + a], -1.0, e->hKind[s + a], 0, 0};
    Probe n = {newX[a], newY[a], newZ[a], e->hCharge[s + a], 1.0, e->hKind[s + a], 1, 0};
    pr[2 * a] = o;
    pr[2 * a + 1] = n;
  }
  std::vector<double> out;
  rc = run_probes(e, box, molIndex, 2 * len, out);
  if (rc) return rc;
  double lj = 0.0, re = 0.0;
  int ov = 0;
  for (int t = 0; t < 2 * len; ++t) {
    lj += out[3 * t];
    re += out[3 * t + 1];
    if (out[3 * t + 2] != 0.0) ov = 1;
  }
  if (dLJ) *dLJ = lj;
  if (dReal) *dReal = re;
  if (overlap) *overlap = ov;
  return 0;
}

int gomcb200_molecule_trial(gomcb200_engine *e, int box, int molIndex, const double *newX,
                            const double *newY, const double *newZ, double *dLJ, double *dReal,
                            int *overlap, double *energyRecipNew) {
  GB_RANGE("energy_molecule_inter");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_molecule_trial");
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !newX || !newY || !newZ)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  if (len <= kTrialMaxAtoms) {
    double r[6];
    rc = run_trial_fused(e, box, molIndex, newX, newY, newZ, 0, true,
                         e->ewald && e->electrostatic, false, r);
    if (rc) return rc;
    if (dLJ) *dLJ = r[0];
    if (dReal) *dReal = r[1];
    if (overlap) *overlap = r[2] != 0.0;
    if (energyRecipNew) *energyRecipNew = r[3];
    return 0;
  }
  const size_t probeBytes = (sizeof(Probe) + 3 * sizeof(double)) * 2 * (size_t)len;
  const size_t molOff = (probeBytes + 63) & ~(size_t)63;
  rc = stage_reserve(e, molOff + (1 + 7 * (size_t)len) * sizeof(double) + 64);
  if (rc) return rc;
  rc = ensure_mirror(e);
  if (rc) return rc;
  Probe *pr = reinterpret_cast<Probe *>(e->hStage);
  for (int a = 0; a < len; ++a) {
    Probe o = {e->hx[s + a], e->hy[s + a], e->hz[s + a], e->hCharge[s + a], -1.0,
               e->hKind[s + a], 0, 0};
    Probe n = {newX[a], newY[a], newZ[a], e->hCharge[s + a], 1.0, e->hKind[s + a], 1, 0};
    pr[2 * a] = o;
    pr[2 * a + 1] = n;
  }
  std::vector<double> unused;
  rc = run_probes(e, box, molIndex, 2 * len, unused, /*sync=*/false);
  if (rc) return rc;
  BoxState &bx = e->box[box];
  const bool recip = e->ewald && e->electrostatic && bx.kset[1 - bx.cur].n > 0;
  double dummy = 0.0;
  if (recip) {
    rc = run_mol_recip(e, box, molIndex, newX, newY, newZ, 0, &dummy, molOff, /*sync=*/false);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(e->stream));  // the only synchronisation of the trial
  const double *hOut = reinterpret_cast<const double *>(reinterpret_cast<const char *>(e->hStage) +
                                                        sizeof(Probe) * 2 * (size_t)len);
  double lj = 0.0, re = 0.0;
  int ov = 0;
  for (int t = 0; t < 2 * len; ++t) {
    lj += hOut[3 * t];
    re += hOut[3 * t + 1];
    if (hOut[3 * t + 2] != 0.0) ov = 1;
  }
  if (dLJ) *dLJ = lj;
  if (dReal) *dReal = re;
  if (overlap) *overlap = ov;
  if (energyRecipNew) *energyRecipNew = recip ? e->hRes[0] : 0.0;
  return 0;
}

int gomcb200_swap_correction(gomcb200_engine *e, int box, int molIndex, const double *x,
                             const double *y, const double *z, double *correction,
                             double *self) {
  GB_RANGE("ewald_molecule_swap_correction_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !x || !y || !z)
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (!(e->ewald && e->electrostatic)) {
    if (correction) *correction = 0.0;
    if (self) *self = 0.0;
    return 0;
  }
  CK(cudaSetDevice(e->device));
  rc = stage_molbuf(e, molIndex, x, y, z, 1, 0);
  if (rc) return rc;
  k_swap_correction<<<1, 128, 0, e->stream>>>(make_params(e, box), e->molBuf.p, e->result.p + 1);
  e->launches += 1;
  CK(cudaGetLastError());
  rc = fetch_result(e, 3);
  if (rc) return rc;
  if (correction) *correction = e->hRes[1];
  if (self) *self = e->hRes[2];
  return 0;
}

int gomcb200_change_self_correction(gomcb200_engine *e, int box, int molIndex, double *enSelf,
                                    double *correction) {
  int rc = check_box(e, box);
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols) return fail(GOMCB200_EINVAL, "bad arguments");
  if (enSelf) *enSelf = 0.0;
  if (correction) *correction = 0.0;
  if (!(e->ewald && e->electrostatic)) return 0;
  CK(cudaSetDevice(e->device));
  rc = ensure_mirror(e);  // resident coordinates of the molecule
  if (rc) return rc;
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  const size_t nd = 1 + 7 * (size_t)len;
  rc = stage_reserve(e, nd * sizeof(double));
  if (rc) return rc;
  CK(e->molBuf.reserve(nd + 8));
  double *h = e->hStage;
  h[0] = (double)len;
  for (int a = 0; a < len; ++a) {
    double *m = h + 1 + 7 * a;
    m[0] = e->hCharge[s + a];  // lambda = 1 charges (src/Ewald.cpp:1099-1114, :1403-1408)
    m[1] = e->hx[s + a];
    m[2] = e->hy[s + a];
    m[3] = e->hz[s + a];
    m[4] = e->hCharge[s + a];
    m[5] = m[6] = 0.0;
  }
  CK(cudaMemcpyAsync(e->molBuf.p, h, nd * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  k_swap_correction<<<1, 128, 0, e->stream>>>(make_params(e, box), e->molBuf.p, e->result.p + 1);
  e->launches += 1;
  CK(cudaGetLastError());
  rc = fetch_result(e, 3);
  if (rc) return rc;
  if (correction) *correction = e->hRes[1];
  if (enSelf) *enSelf = e->hRes[2];
  return 0;
}

int gomcb200_swap_trial(gomcb200_engine *e, int box, int molIndex, const double *x,
                        const double *y, const double *z, int insert, double *energyRecipNew,
                        double *correction, double *self) {
  GB_RANGE("ewald_molecule_swap_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_swap_trial");
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !x || !y || !z)
    return fail(GOMCB200_EINVAL, "bad arguments");
  double en = 0.0, co = 0.0, se = 0.0;
  if (e->ewald && e->electrostatic &&
      e->hMolStart[molIndex + 1] - e->hMolStart[molIndex] <= kTrialMaxAtoms) {
    CK(cudaSetDevice(e->device));
    double r[6];
    rc = run_trial_fused(e, box, molIndex, x, y, z, insert ? 1 : 2, false, true, true, r);
    if (rc) return rc;
    en = r[3];
    co = r[4];
    se = r[5];
  } else if (e->ewald && e->electrostatic) {
    CK(cudaSetDevice(e->device));
    BoxState &bx = e->box[box];
    const bool recip = bx.kset[1 - bx.cur].n > 0;
    if (recip) {
      rc = run_mol_recip(e, box, molIndex, x, y, z, insert ? 1 : 2, &en, 0, /*sync=*/false);
    } else {
      rc = stage_molbuf(e, molIndex, x, y, z, 1, 0);
    }
    if (rc) return rc;
    k_swap_correction<<<1, 128, 0, e->stream>>>(make_params(e, box), e->molBuf.p,
                                                e->result.p + 1);
    e->launches += 1;
    CK(cudaGetLastError());
    rc = fetch_result(e, 3);  // the only synchronisation of the trial
    if (rc) return rc;
    en = recip ? e->hRes[0] : 0.0;
    co = e->hRes[1];
    se = e->hRes[2];
  }
  if (energyRecipNew) *energyRecipNew = en;
  if (correction) *correction = co;
  if (self) *self = se;
  return 0;
}

int gomcb200_particle_inter(gomcb200_engine *e, int box, int molIndex, int partIndex,
                            int trials, const double *tx, const double *ty, const double *tz,
                            double *en, double *real, int *overlap) {
  GB_RANGE("energy_CBMC_inter");
  int rc = check_box(e, box);
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || trials < 0 || !tx || !ty || !tz)
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (trials == 0) return 0;
  CK(cudaSetDevice(e->device));
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  if (partIndex < 0 || partIndex >= len) return fail(GOMCB200_EINVAL, "partIndex out of range");
  rc = stage_reserve(e, (sizeof(Probe) + 3 * sizeof(double)) * (size_t)trials + 64);
  if (rc) return rc;
  Probe *pr = reinterpret_cast<Probe *>(e->hStage);
  for (int t = 0; t < trials; ++t) {
    Probe n = {tx[t], ty[t], tz[t], e->hCharge[s + partIndex], 1.0, e->hKind[s + partIndex],
               1, 0};
    pr[t] = n;
  }
  std::vector<double> out;
  rc = run_probes(e, box, molIndex, trials, out);
  if (rc) return rc;
  for (int t = 0; t < trials; ++t) {
    if (en) en[t] += out[3 * t];
    if (real) real[t] += out[3 * t + 1];
    if (overlap && out[3 * t + 2] != 0.0) overlap[t] |= 1;
  }
  return 0;
}

int gomcb200_particle_nonbonded(gomcb200_engine *e, int box, int kindI, double chargeI,
                                int nPartners, const int *partnerKind,
                                const double *partnerCharge, const double *px, const double *py,
                                const double *pz, int trials, const double *tx, const double *ty,
                                const double *tz, double *inter) {
  GB_RANGE("energy_CBMC_intra_nonbonded");
  int rc = check_box(e, box, true, false);
  if (rc) return rc;
  if (kindI < 0 || kindI >= e->kindCount || nPartners < 0 || trials < 0 || !inter ||
      (nPartners && (!partnerKind || !partnerCharge || !px || !py || !pz)) ||
      (trials && (!tx || !ty || !tz)))
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (trials == 0 || nPartners == 0) return 0;
  for (int k = 0; k < nPartners; ++k)
    if (partnerKind[k] < 0 || partnerKind[k] >= e->kindCount)
      return fail(GOMCB200_EINVAL, "partner kind out of range");
  CK(cudaSetDevice(e->device));
  const size_t nd = 4 * (size_t)nPartners + 3 * (size_t)trials;
  const size_t bytes = nd * sizeof(double) + (size_t)nPartners * sizeof(int);
  rc = stage_reserve(e, bytes + 64);
  if (rc) return rc;
  double *h = reinterpret_cast<double *>(e->hStage);
  std::memcpy(h, px, sizeof(double) * nPartners);
  std::memcpy(h + nPartners, py, sizeof(double) * nPartners);
  std::memcpy(h + 2 * (size_t)nPartners, pz, sizeof(double) * nPartners);
  std::memcpy(h + 3 * (size_t)nPartners, partnerCharge, sizeof(double) * nPartners);
  double *ht = h + 4 * (size_t)nPartners;
  std::memcpy(ht, tx, sizeof(double) * trials);
  std::memcpy(ht + trials, ty, sizeof(double) * trials);
  std::memcpy(ht + 2 * (size_t)trials, tz, sizeof(double) * trials);
  std::memcpy(h + nd, partnerKind, sizeof(int) * nPartners);
  CK(e->molBuf.reserve(nd + (size_t)nPartners + 8));
  CK(e->probeOut.reserve((size_t)trials + 8));
  CK(cudaMemcpyAsync(e->molBuf.p, h, bytes, cudaMemcpyHostToDevice, e->stream));
  const BoxParams p = make_params(e, box);
  const int *dk = reinterpret_cast<const int *>(e->molBuf.p + nd);
  const int blocks = (trials + 63) / 64;
#define PNB(V)                                                                              \
  k_particle_nonbonded<V><<<blocks, 64, 0, e->stream>>>(p, kindI, chargeI, nPartners, dk, \
                                                        e->molBuf.p, trials, e->probeOut.p)
  if (e->vdwKind == VDW_SHIFT)
    PNB(VDW_SHIFT);
  else if (e->vdwKind == VDW_SWITCH)
    PNB(VDW_SWITCH);
  else if (e->vdwKind == VDW_EXP6)
    PNB(VDW_EXP6);
  else if (e->vdwKind == VDW_MARTINI)
    PNB(VDW_MARTINI);
  else
    PNB(VDW_STD);
#undef PNB
  e->launches += 1;
  CK(cudaGetLastError());
  std::vector<double> out(trials);
  CK(cudaMemcpyAsync(out.data(), e->probeOut.p, sizeof(double) * trials, cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  for (int t = 0; t < trials; ++t) inter[t] += out[t];
  return 0;
}

int gomcb200_calculate_torque(gomcb200_engine *e, int box) {
  GB_RANGE("energy_box_torque");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_calculate_torque", true);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  if (bx.nMols == 0) return 0;
  BoxParams p = make_params(e, box);
  k_torque<<<(bx.nMols + 255) / 256, 256, 0, e->stream>>>(
      p, bx.nMols, bx.molList.p, e->molStart.p, e->x.p, e->y.p, e->z.p, e->comx.p,
      e->comy.p, e->comz.p, e->force[0][0].p, e->force[0][1].p, e->force[0][2].p,
      e->force[2][0].p, e->force[2][1].p, e->force[2][2].p, e->force[4][0].p,
      e->force[4][1].p, e->force[4][2].p);
  e->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int gomcb200_get_forces(gomcb200_engine *e, int which, double *x, double *y, double *z,
                        int first, int count) {
  if (!e || !e->haveTopo || which < 0 || which > 4) return fail(GOMCB200_EINVAL, "bad arguments");
  int limit = (which == GOMCB200_ATOM_FORCE || which == GOMCB200_ATOM_FORCE_REC) ? e->nAtoms
                                                                                 : e->nMols;
  if (first < 0 || count < 0 || first + count > limit)
    return fail(GOMCB200_EINVAL, "range out of bounds");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  size_t bytes = sizeof(double) * (size_t)count;
  if (x) CK(cudaMemcpy(x, e->force[which][0].p + first, bytes, cudaMemcpyDeviceToHost));
  if (y) CK(cudaMemcpy(y, e->force[which][1].p + first, bytes, cudaMemcpyDeviceToHost));
  if (z) CK(cudaMemcpy(z, e->force[which][2].p + first, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- reciprocal path --------------------------------------------------------
int gomcb200_init_ewald(gomcb200_engine *e, int imageTotal, const double *recip_rcut) {
  if (!e || imageTotal < 0 || !recip_rcut) return fail(GOMCB200_EINVAL, "bad arguments");
  if (!e->haveFF) return fail(GOMCB200_EINVAL, "gomcb200_init_forcefield not called");
  e->imageTotal = imageTotal;
  for (int b = 0; b < e->nBoxes; ++b) e->recipRcut[b] = recip_rcut[b];
  return 0;
}

int gomcb200_recip_count(gomcb200_engine *e, int box, const double axis[3], double excess,
                         int *imageSize) {
  if (!e || box < 0 || box >= e->nBoxes || !axis || !imageSize || !e->haveFF)
    return fail(GOMCB200_EINVAL, "bad arguments");
  double ax[3] = {axis[0] * excess, axis[1] * excess, axis[2] * excess};
  *imageSize = recip_enumerate(e, box, ax, nullptr, nullptr);
  return 0;
}

int gomcb200_recip_init(gomcb200_engine *e, int box, const double axis[3], int *imageSize,
                        int *kmax) {
  GB_RANGE("ewald_box_recip_initialization");
  return gomcb200_recip_init_volume(e, box, axis, 0.0, imageSize, kmax);
}

int gomcb200_recip_init_volume(gomcb200_engine *e, int box, const double axis[3], double volume,
                               int *imageSize, int *kmax) {
  GB_RANGE("ewald_box_recip_initialization");
  if (!e || box < 0 || box >= e->nBoxes || !axis || !e->haveFF || !(volume >= 0.0))
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[bx.cur];
  std::vector<RowRec> rows;
  int n = recip_enumerate(e, box, axis, &ks, &rows, volume);
  ks.planValid = false;
  if (n < 0) {  // cannot happen for an orthogonal box; keep the direct kernel usable
    rows.clear();
    n = recip_enumerate(e, box, axis, &ks, nullptr, volume);
  }
  if (e->imageTotal > 0 && n > e->imageTotal)
    return fail(GOMCB200_EKMAX,
                "Kmax exceeded due to large change in system volume (%d > imageTotal %d)", n,
                e->imageTotal);
  // With a row table the device regenerates kx, ky, kz and |k|^2 itself (same products and
  // sums, no contraction, hence the same bits as the host list that get_kvectors returns);
  // only the prefactor, whose exp() must stay the host's, crosses PCIe.
  int rc = upload_kset(e, ks, !rows.empty());
  if (rc) return rc;
  ks.mmaValid = false;
  ks.fmValid = false;
  ks.ngValid = false;
  if (!rows.empty()) {
    rc = build_plan(e, ks, rows);
    if (rc) return rc;
    ks.ngValid = gbn::nufft_choose(ks.nmax, &ks.ng) == 0;
    k_gen_kvectors<<<(unsigned)ks.nRowsPadded, 64, 0, e->stream>>>(
        ks.rows.p, ks.cv[0], ks.cv[1], ks.cv[2], ks.kx.p, ks.ky.p, ks.kz.p, ks.hsqr.p);
    CK(cudaGetLastError());
  }
  if (imageSize) *imageSize = n;
  if (kmax) *kmax = ks.kmax;
  return 0;
}

int gomcb200_get_kvectors(gomcb200_engine *e, int box, int which, double *kx, double *ky,
                          double *kz, double *hsqr, double *prefact, int n) {
  if (!e || box < 0 || box >= e->nBoxes) return fail(GOMCB200_EINVAL, "bad arguments");
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[(which & 1) == GOMCB200_K_NEW ? bx.cur : 1 - bx.cur];
  if (n > ks.n) return fail(GOMCB200_EINVAL, "n %d > imageSize %d", n, ks.n);
  size_t bytes = sizeof(double) * (size_t)n;
  if (which & GOMCB200_K_DEVICE) {  // the copies the kernels read
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    if (kx) CK(cudaMemcpy(kx, ks.kx.p, bytes, cudaMemcpyDeviceToHost));
    if (ky) CK(cudaMemcpy(ky, ks.ky.p, bytes, cudaMemcpyDeviceToHost));
    if (kz) CK(cudaMemcpy(kz, ks.kz.p, bytes, cudaMemcpyDeviceToHost));
    if (hsqr) CK(cudaMemcpy(hsqr, ks.hsqr.p, bytes, cudaMemcpyDeviceToHost));
    if (prefact) CK(cudaMemcpy(prefact, ks.prefact.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (kx) memcpy(kx, ks.hkx.data(), bytes);
  if (ky) memcpy(ky, ks.hky.data(), bytes);
  if (kz) memcpy(kz, ks.hkz.data(), bytes);
  if (hsqr) memcpy(hsqr, ks.hhsqr.data(), bytes);
  if (prefact) memcpy(prefact, ks.hprefact.data(), bytes);
  return 0;
}

static int recip_sums_common(gomcb200_engine *e, int box, bool newSet, double *energyRecip) {
  int rc = check_box(e, box);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  timing_begin(e);
  rc = run_recip_sums(e, box, bx.kset[newSet ? bx.cur : 1 - bx.cur]);
  if (rc) return rc;
  rc = allreduce_energies(e, e->result.p, 1);
  if (rc) return rc;
  rc = fetch_result(e, 1);
  if (rc) return rc;
  timing_end(e, true);
  if (energyRecip) *energyRecip = e->hRes[0];
  return 0;
}

int gomcb200_box_reciprocal_setup(gomcb200_engine *e, int box, double *energyRecip) {
  GB_RANGE("ewald_box_recip_setup");
  return recip_sums_common(e, box, true, energyRecip);
}

int gomcb200_box_reciprocal_sums(gomcb200_engine *e, int box, double *energyRecip) {
  GB_RANGE("ewald_box_recip_energy");
  return recip_sums_common(e, box, false, energyRecip);
}

int gomcb200_box_reciprocal(gomcb200_engine *e, int box, int isNewVolume, double *energyRecip) {
  GB_RANGE("ewald_box_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[isNewVolume ? bx.cur : 1 - bx.cur];
  if (ks.n == 0) {
    if (energyRecip) *energyRecip = 0.0;
    return 0;
  }
  rc = ensure_sums(e, bx, ks.n);
  if (rc) return rc;
  int nBlocks = (ks.n + 255) / 256;
  CK(e->blockA.reserve(nBlocks + 1024));
  k_recip_energy<<<nBlocks, 256, 0, e->stream>>>(ks.n, bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p,
                                                ks.prefact.p, e->blockA.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, 1, e->blockA.p, nullptr, nullptr, nullptr,
                                           e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  rc = fetch_result(e, 1);
  if (rc) return rc;
  if (energyRecip) *energyRecip = e->hRes[0];
  return 0;
}

int gomcb200_mol_reciprocal(gomcb200_engine *e, int box, int molIndex, const double *newX,
                            const double *newY, const double *newZ, double *energyRecipNew) {
  GB_RANGE("ewald_molecule_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_mol_reciprocal");
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !newX || !newY || !newZ || !energyRecipNew)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  return run_mol_recip(e, box, molIndex, newX, newY, newZ, 0, energyRecipNew);
}

int gomcb200_swap_reciprocal(gomcb200_engine *e, int box, int molIndex, const double *x,
                             const double *y, const double *z, int insert,
                             double *energyRecipNew) {
  GB_RANGE("ewald_molecule_swap_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_swap_reciprocal");
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || !x || !y || !z || !energyRecipNew)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  return run_mol_recip(e, box, molIndex, x, y, z, insert ? 1 : 2, energyRecipNew);
}

int gomcb200_mol_exchange_reciprocal(gomcb200_engine *e, int box, int n, const double *w,
                                     const double *x, const double *y, const double *z,
                                     int firstCall, double scale, double *energyRecipNew) {
  GB_RANGE("ewald_molecule_MEMC_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_mol_exchange_reciprocal");
  if (rc) return rc;
  if (n < 0 || (n && (!w || !x || !y || !z)) || !energyRecipNew)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[1 - bx.cur];
  const int nk = ks.n;
  *energyRecipNew = 0.0;
  if (nk == 0 || !(e->ewald && e->electrostatic)) return 0;
  rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  rc = stage_reserve(e, 4 * (size_t)n * sizeof(double) + 64);
  if (rc) return rc;
  CK(e->molBuf.reserve(4 * (size_t)n + 8));
  double *h = e->hStage;
  const double *src[4] = {w, x, y, z};
  for (int f = 0; f < 4; ++f)
    if (n) memcpy(h + (size_t)f * n, src[f], sizeof(double) * (size_t)n);
  if (n)
    CK(cudaMemcpyAsync(e->molBuf.p, h, 4 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                       e->stream));
  const int nBlocks = (nk + 255) / 256;
  CK(e->blockA.reserve(nBlocks + 1024));
  const double *bR = firstCall ? bx.sum[bx.iRref].p : bx.sum[bx.iRnew].p;
  const double *bI = firstCall ? bx.sum[bx.iIref].p : bx.sum[bx.iInew].p;
  k_recip_weighted<<<nBlocks, 256, 0, e->stream>>>(nk, n, e->molBuf.p, scale, ks.kx.p, ks.ky.p,
                                                  ks.kz.p, ks.prefact.p, bR, bI,
                                                  bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p,
                                                  e->blockA.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, 1, e->blockA.p, nullptr, nullptr, nullptr,
                                           e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *energyRecipNew = e->hRes[0];
  return 0;
}

int gomcb200_change_lambda_mol_reciprocal(gomcb200_engine *e, int box, int molIndex,
                                          const double *x, const double *y, const double *z,
                                          double lambdaCoef, double *energyRecipNew) {
  GB_RANGE("ewald_molecule_NEMTMC_recip_energy");
  if (!e || !e->haveTopo || molIndex < 0 || molIndex >= e->nMols || !x || !y || !z)
    return fail(GOMCB200_EINVAL, "bad arguments");
  const int s = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - s;
  std::vector<double> w, cx, cy, cz;
  for (int a = 0; a < len; ++a) {
    if (std::fabs(e->hCharge[s + a]) < 0.000000001) continue;  // particleHasNoCharge
    w.push_back(e->hCharge[s + a]);
    cx.push_back(x[a]);
    cy.push_back(y[a]);
    cz.push_back(z[a]);
  }
  return gomcb200_mol_exchange_reciprocal(e, box, (int)w.size(), w.data(), cx.data(), cy.data(),
                                          cz.data(), 1, lambdaCoef, energyRecipNew);
}

int gomcb200_change_recip(gomcb200_engine *e, int box, int molIndex, int nStates,
                          const double *lambdaCoul, int iState, double *energyRecip) {
  GB_RANGE("ewald_molecule_NEMTMC_recip_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_change_recip");
  if (rc) return rc;
  if (molIndex < 0 || molIndex >= e->nMols || nStates < 1 || nStates > kMaxLambdaStates ||
      !lambdaCoul || iState < 0 || iState >= nStates || !energyRecip)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[1 - bx.cur];
  const int nk = ks.n;
  for (int s = 0; s < nStates; ++s) energyRecip[s] = 0.0;
  if (nk == 0 || !(e->ewald && e->electrostatic)) return 0;
  rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  rc = stage_reserve(e, kMaxLambdaStates * sizeof(double));
  if (rc) return rc;
  CK(e->molBuf.reserve(kMaxLambdaStates + 8));
  for (int s = 0; s < nStates; ++s)
    e->hStage[s] = std::sqrt(lambdaCoul[s]) - std::sqrt(lambdaCoul[iState]);
  CK(cudaMemcpyAsync(e->molBuf.p, e->hStage, nStates * sizeof(double), cudaMemcpyHostToDevice,
                     e->stream));
  const int nBlocks = (nk + 255) / 256;
  CK(e->blockA.reserve((size_t)nBlocks * nStates + 1024));
  const int first = e->hMolStart[molIndex], len = e->hMolStart[molIndex + 1] - first;
  k_change_recip<<<nBlocks, 256, 0, e->stream>>>(nk, first, len, e->x.p, e->y.p, e->z.p, e->q.p,
                                                nStates, e->molBuf.p, ks.kx.p, ks.ky.p, ks.kz.p,
                                                ks.prefact.p, bx.sum[bx.iRref].p,
                                                bx.sum[bx.iIref].p, e->blockA.p);
  e->launches += 1;
  for (int s0 = 0; s0 < nStates; s0 += 4) {
    const int c = std::min(4, nStates - s0);
    const double *a[4];
    for (int j = 0; j < 4; ++j) a[j] = e->blockA.p + (size_t)(s0 + std::min(j, c - 1)) * nBlocks;
    k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, c, a[0], a[1], a[2], a[3],
                                             e->result.p + s0);
    e->launches += 1;
  }
  CK(cudaGetLastError());
  rc = fetch_result(e, nStates);
  if (rc) return rc;
  for (int s = 0; s < nStates; ++s) energyRecip[s] = e->hRes[s];
  return 0;
}

// Reciprocal force of every atom of the box from the sums (sumR, sumI) into rf*;
// withIntra: include the intramolecular correction force (src/Ewald.cpp:1556-1569),
// else the pure k-space part (what VirialReciprocal's per-atom k sum is made of).
static int run_force_recip(gomcb200_engine *e, int box, const double *sumR, const double *sumI,
                           double *rfx, double *rfy, double *rfz, bool withIntra) {
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[1 - bx.cur];
  BoxParams p = make_params(e, box);
  bool done = false;
  int rc;
  if ((e->recipAlgo == 4 || e->recipAlgo == 5) && ks.planValid && ks.ngValid && bx.nCharged > 0 &&
      ks.n > 0) {
    // type-2 non-uniform FFT: potential on the fine grid, analytic window gradient
    rc = ensure_packed(e, box);
    if (rc) return rc;
    if (withIntra) {
      k_force_recip_intra<<<(bx.nAtoms + 127) / 128, 128, 0, e->stream>>>(
          p, bx.nAtoms, bx.atomList.p, e->mol.p, e->molStart.p, e->x.p, e->y.p, e->z.p,
          e->qEff.p, rfx, rfy, rfz);
      e->launches += 1;
    } else {
      const size_t bytes = sizeof(double) * (size_t)e->nAtoms;
      CK(cudaMemsetAsync(rfx, 0, bytes, e->stream));
      CK(cudaMemsetAsync(rfy, 0, bytes, e->stream));
      CK(cudaMemsetAsync(rfz, 0, bytes, e->stream));
    }
    rc = gbn::nufft_type2_force(e->nufft, e->stream, ks.ng, ks.L, bx.packed.p,
                                bx.chargedList.p, bx.nCharged, ks.rows.p, ks.nRowsPadded,
                                ks.prefact.p, sumR, sumI, rfx, rfy, rfz, 0, &e->launches);
    if (rc) return fail(GOMCB200_ECUDA, "nufft_type2_force: %s", gbn::nufft_last_error(e->nufft));
    done = true;
  } else if (e->recipAlgo >= 2 && ks.fmValid && bx.nCharged > 0 && ks.n > 0) {
    rc = ensure_packed(e, box);
    if (rc) return rc;
    FmArgs fa;
    fa.blocks = ks.fmBlocks.p;
    fa.nBlocks = ks.fmNBlocks;
    fa.tiles = ks.fmTiles.p;
    fa.rows = ks.fmRows.p;
    fa.wm = ks.fmW.p;
    fa.pb = bx.packed.p;
    fa.chargedAtoms = bx.chargedList.p;
    fa.nTiles = ks.fmNTiles;
    fa.nAtoms = bx.nCharged;
    fa.KX1 = ks.nmax[0] + 1;
    fa.KY1 = ks.nmax[1] + 1;
    fa.KZ1 = ks.nmax[2] + 1;
    int zfs = (2 * fa.KZ1 + 3) & ~3;
    while (zfs % 16 != 4) zfs += 4;
    fa.ZFS = zfs;
    fa.XYS = fa.KX1 + fa.KY1;
    fa.cvx = ks.cv[0];
    fa.cvy = ks.cv[1];
    fa.cvz = ks.cv[2];
    const size_t wBuf = (size_t)kFmKB * kFmWS * 8;
    auto smemFor = [&](int AB) {
      return (size_t)AB * fa.ZFS * 8 + (size_t)AB * fa.XYS * 16 + 2 * wBuf;
    };
    const size_t budget = e->smemOptin > 4096 ? e->smemOptin - 2048 : 0;
    int AB = smemFor(64) <= budget ? 64 : (smemFor(32) <= budget ? 32 : 0);
    if (AB) {
      k_force_wmat<<<ks.fmNTiles, 256, 0, e->stream>>>(ks.fmNTiles, ks.fmTiles.p, ks.fmRows.p,
                                                      sumR, sumI, ks.prefact.p, ks.fmW.p);
      if (withIntra) {
        k_force_recip_intra<<<(bx.nAtoms + 127) / 128, 128, 0, e->stream>>>(
            p, bx.nAtoms, bx.atomList.p, e->mol.p, e->molStart.p, e->x.p, e->y.p, e->z.p,
            e->qEff.p, rfx, rfy, rfz);
      } else {  // the DMMA kernel accumulates onto the buffers
        const size_t bytes = sizeof(double) * (size_t)e->nAtoms;
        CK(cudaMemsetAsync(rfx, 0, bytes, e->stream));
        CK(cudaMemsetAsync(rfy, 0, bytes, e->stream));
        CK(cudaMemsetAsync(rfz, 0, bytes, e->stream));
      }
      const int grid = (bx.nCharged + AB - 1) / AB;
      const size_t smem = smemFor(AB);
      if (AB == 64) {
        CK(cudaFuncSetAttribute(k_force_recip_mma<64>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_force_recip_mma<64><<<grid, kFmThreads, smem, e->stream>>>(fa, rfx, rfy, rfz);
      } else {
        CK(cudaFuncSetAttribute(k_force_recip_mma<32>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_force_recip_mma<32><<<grid, kFmThreads, smem, e->stream>>>(fa, rfx, rfy, rfz);
      }
      e->launches += 3;
      done = true;
    }
  }
  if (!done) {
    k_force_recip_direct<<<(bx.nAtoms + 127) / 128, 128, 0, e->stream>>>(
        p, bx.nAtoms, bx.atomList.p, e->mol.p, e->molStart.p, e->x.p, e->y.p, e->z.p, e->qEff.p,
        ks.n, ks.kx.p, ks.ky.p, ks.kz.p, ks.prefact.p, sumR, sumI, rfx, rfy, rfz,
        withIntra ? 1 : 0);
    e->launches += 1;
  }
  CK(cudaGetLastError());
  return 0;
}

int gomcb200_box_force_reciprocal(gomcb200_engine *e, int box) {
  GB_RANGE("ewald_box_recip_force");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_box_force_reciprocal", true);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  if (e->shardWorld > 1 && !bx.sumsComplete)
    return fail(GOMCB200_EINVAL,
                "gomcb200_box_force_reciprocal on a sharded engine needs the complete structure "
                "factor on every rank (the non-uniform FFT path of BoxReciprocalSums)");
  KSet &ks = bx.kset[1 - bx.cur];
  rc = ensure_sums(e, bx, ks.n);
  if (rc) return rc;
  if (bx.nAtoms == 0) return 0;
  rc = run_force_recip(e, box, bx.sum[bx.iRnew].p, bx.sum[bx.iInew].p, e->force[2][0].p,
                       e->force[2][1].p, e->force[2][2].p, true);
  if (rc) return rc;
  k_mol_force<<<(bx.nMols + 255) / 256, 256, 0, e->stream>>>(
      bx.nMols, bx.molList.p, e->molStart.p, e->force[2][0].p, e->force[2][1].p,
      e->force[2][2].p, e->force[3][0].p, e->force[3][1].p, e->force[3][2].p);
  e->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---- MultiParticle move -------------------------------------------------------
static int mp_reserve(gomcb200_engine *e) {
  const size_t na = (size_t)e->nAtoms + 1, nm = (size_t)e->nMols + 1;
  CK(e->xT.reserve(na)); CK(e->yT.reserve(na)); CK(e->zT.reserve(na));
  CK(e->comxT.reserve(nm)); CK(e->comyT.reserve(nm)); CK(e->comzT.reserve(nm));
  for (int w = 0; w < 5; ++w) {
    size_t n = (w == GOMCB200_ATOM_FORCE || w == GOMCB200_ATOM_FORCE_REC) ? na : nm;
    for (int c = 0; c < 3; ++c)
      if (e->forceT[w][c].cap < n) {
        CK(e->forceT[w][c].reserve(n));
        CK(cudaMemset(e->forceT[w][c].p, 0, e->forceT[w][c].cap * sizeof(double)));
      }
  }
  for (int c = 0; c < 3; ++c) CK(e->mpK[c].reserve(nm));
  CK(e->mpInRange.reserve(nm));
  return 0;
}

static int mp_transform_impl(gomcb200_engine *e, int box, int brownian, int moveType,
                             double max, double lambdaBETA, unsigned long long step,
                             unsigned int key, unsigned long long seed,
                             const signed char *isMoleculeInvolved);

int gomcb200_mp_transform(gomcb200_engine *e, int box, int moveType, double max,
                          double lambdaBETA, unsigned long long step, unsigned int key,
                          unsigned long long seed, const signed char *isMoleculeInvolved) {
  GB_RANGE("transform_multiParticle_move");
  return mp_transform_impl(e, box, 0, moveType, max, lambdaBETA, step, key, seed,
                           isMoleculeInvolved);
}

int gomcb200_bm_transform(gomcb200_engine *e, int box, int moveType, double max, double BETA,
                          unsigned long long step, unsigned int key, unsigned long long seed,
                          const signed char *isMoleculeInvolved) {
  GB_RANGE("transform_multiParticleBM_move");
  return mp_transform_impl(e, box, 1, moveType, max, BETA, step, key, seed,
                           isMoleculeInvolved);
}

static int mp_transform_impl(gomcb200_engine *e, int box, int brownian, int moveType,
                             double max, double lambdaBETA, unsigned long long step,
                             unsigned int key, unsigned long long seed,
                             const signed char *isMoleculeInvolved) {
  int rc = check_box(e, box);
  if (rc) return rc;
  if (moveType < 0 || moveType > 1) return fail(GOMCB200_EINVAL, "moveType must be 0 or 1");
  if (e->trialActive)
    return fail(GOMCB200_EINVAL, "trial coordinates are active: gomcb200_mp_select(e, 0) first");
  CK(cudaSetDevice(e->device));
  rc = mp_reserve(e);
  if (rc) return rc;
  BoxState &bx = e->box[box];
  const size_t ba = sizeof(double) * (size_t)e->nAtoms, bm = sizeof(double) * (size_t)e->nMols;
  // newMolsPos / newCOMs start as copies of the reference (MultiParticle::Prep, :238-239)
  CK(cudaMemcpyAsync(e->xT.p, e->x.p, ba, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->yT.p, e->y.p, ba, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->zT.p, e->z.p, ba, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->comxT.p, e->comx.p, bm, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->comyT.p, e->comy.p, bm, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->comzT.p, e->comz.p, bm, cudaMemcpyDeviceToDevice, e->stream));
  if (e->nBoxes > 1) {
    // two boxes (GEMC/GCMC): the other box's forces and torques must survive the buffer
    // exchange of an accepted move -- MultiParticle::Prep copies atomForceRef, molForceRef,
    // the Rec forces and the torques into the New sets (src/moves/MultiParticle.h:235-239)
    for (int w = 0; w < 5; ++w) {
      const size_t bytes =
          (w == GOMCB200_ATOM_FORCE || w == GOMCB200_ATOM_FORCE_REC) ? ba : bm;
      for (int c = 0; c < 3; ++c)
        CK(cudaMemcpyAsync(e->forceT[w][c].p, e->force[w][c].p, bytes, cudaMemcpyDeviceToDevice,
                           e->stream));
    }
  }
  CK(cudaMemsetAsync(e->mpInRange.p, 0, sizeof(int) * (size_t)e->nMols, e->stream));
  for (int c = 0; c < 3; ++c) CK(cudaMemsetAsync(e->mpK[c].p, 0, bm, e->stream));
  if (bx.nMols == 0) return 0;
  MpArgs a;
  a.brownian = brownian;
  a.moveType = moveType;
  a.nMolsBox = bx.nMols;
  a.max = max;
  a.lambdaBeta = lambdaBETA;
  a.step = step;
  a.seed = seed;
  a.key = key;
  a.molList = bx.molList.p;
  a.molStart = e->molStart.p;
  a.involved = nullptr;
  if (isMoleculeInvolved) {
    CK(e->mpInvolved.reserve(e->nMols + 1));
    CK(cudaMemcpyAsync(e->mpInvolved.p, isMoleculeInvolved, (size_t)e->nMols,
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // the caller's buffer may be pageable
    a.involved = e->mpInvolved.p;
  }
  a.x = e->x.p; a.y = e->y.p; a.z = e->z.p;
  a.cx = e->comx.p; a.cy = e->comy.p; a.cz = e->comz.p;
  const int fw = moveType == 1 ? GOMCB200_MOL_TORQUE : GOMCB200_MOL_FORCE;
  a.fx = e->force[fw][0].p; a.fy = e->force[fw][1].p; a.fz = e->force[fw][2].p;
  a.rfx = a.rfy = a.rfz = nullptr;
  if (moveType == 0) {
    a.rfx = e->force[GOMCB200_MOL_FORCE_REC][0].p;
    a.rfy = e->force[GOMCB200_MOL_FORCE_REC][1].p;
    a.rfz = e->force[GOMCB200_MOL_FORCE_REC][2].p;
  }
  a.nx = e->xT.p; a.ny = e->yT.p; a.nz = e->zT.p;
  a.ncx = e->comxT.p; a.ncy = e->comyT.p; a.ncz = e->comzT.p;
  a.kx = e->mpK[0].p; a.ky = e->mpK[1].p; a.kz = e->mpK[2].p;
  a.inForceRange = e->mpInRange.p;
  k_mp_transform<<<(bx.nMols + 127) / 128, 128, 0, e->stream>>>(make_params(e, box), a);
  e->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int gomcb200_mp_get_trial(gomcb200_engine *e, double *kx, double *ky, double *kz,
                          int *inForceRange) {
  if (!e || !e->haveTopo || e->mpInRange.cap == 0)
    return fail(GOMCB200_EINVAL, "gomcb200_mp_transform not called");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const size_t bm = sizeof(double) * (size_t)e->nMols;
  if (kx) CK(cudaMemcpy(kx, e->mpK[0].p, bm, cudaMemcpyDeviceToHost));
  if (ky) CK(cudaMemcpy(ky, e->mpK[1].p, bm, cudaMemcpyDeviceToHost));
  if (kz) CK(cudaMemcpy(kz, e->mpK[2].p, bm, cudaMemcpyDeviceToHost));
  if (inForceRange)
    CK(cudaMemcpy(inForceRange, e->mpInRange.p, sizeof(int) * (size_t)e->nMols,
                  cudaMemcpyDeviceToHost));
  return 0;
}

int gomcb200_mp_select(gomcb200_engine *e, int trial) {
  if (!e || !e->haveTopo) return fail(GOMCB200_EINVAL, "topology not initialised");
  if ((trial != 0) == e->trialActive) return 0;
  if (e->xT.cap == 0) return fail(GOMCB200_EINVAL, "gomcb200_mp_transform not called");
  // O(1): the two sets trade places; every kernel reads the active one
  std::swap(e->x, e->xT); std::swap(e->y, e->yT); std::swap(e->z, e->zT);
  std::swap(e->comx, e->comxT); std::swap(e->comy, e->comyT); std::swap(e->comz, e->comzT);
  for (int w = 0; w < 5; ++w)
    for (int c = 0; c < 3; ++c) std::swap(e->force[w][c], e->forceT[w][c]);
  e->trialActive = trial != 0;
  e->mirrorValid = false;
  mark_coords_dirty(e);
  return 0;
}

int gomcb200_mp_accept(gomcb200_engine *e, int box) {
  // MultiParticle::Accept, :522-534: the trial set (active) becomes the reference
  int rc = check_box(e, box);
  if (rc) return rc;
  if (!e->trialActive) return fail(GOMCB200_EINVAL, "no active trial coordinates");
  e->trialActive = false;  // buffers stay swapped: what was the trial is now current
  return gomcb200_update_recip(e, box);
}

int gomcb200_mp_coeff(gomcb200_engine *e, int box, int moveType, double max, double lambdaBETA,
                      double *wRatio) {
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_mp_coeff", true);
  if (rc) return rc;
  if (!wRatio || moveType < 0 || moveType > 1) return fail(GOMCB200_EINVAL, "bad arguments");
  if (!e->trialActive)
    return fail(GOMCB200_EINVAL, "needs the trial set active (new forces computed on it)");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  *wRatio = 1.0;
  if (bx.nMols == 0) return 0;
  const int nb = (bx.nMols + 255) / 256;
  CK(e->blockA.reserve(nb + 1024));
  const int fw = moveType == 1 ? GOMCB200_MOL_TORQUE : GOMCB200_MOL_FORCE;
  const int rw = GOMCB200_MOL_FORCE_REC;
  const bool rec = moveType == 0;
  // while the trial set is active: force[] = new, forceT[] = reference
  k_mp_coeff<<<nb, 256, 0, e->stream>>>(
      bx.nMols, bx.molList.p, e->mpInRange.p, max, lambdaBETA, e->forceT[fw][0].p,
      e->forceT[fw][1].p, e->forceT[fw][2].p, rec ? e->forceT[rw][0].p : nullptr,
      rec ? e->forceT[rw][1].p : nullptr, rec ? e->forceT[rw][2].p : nullptr, e->force[fw][0].p,
      e->force[fw][1].p, e->force[fw][2].p, rec ? e->force[rw][0].p : nullptr,
      rec ? e->force[rw][1].p : nullptr, rec ? e->force[rw][2].p : nullptr, e->mpK[0].p,
      e->mpK[1].p, e->mpK[2].p, e->blockA.p);
  k_mp_coeff_final<<<1, 256, 0, e->stream>>>(nb, e->blockA.p, e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *wRatio = e->hRes[0];
  return 0;
}

int gomcb200_bm_coeff(gomcb200_engine *e, int box, int moveType, double max, double BETA,
                      double *wRatio) {
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_bm_coeff", true);
  if (rc) return rc;
  if (!wRatio || moveType < 0 || moveType > 1) return fail(GOMCB200_EINVAL, "bad arguments");
  if (!e->trialActive)
    return fail(GOMCB200_EINVAL, "needs the trial set active (new forces computed on it)");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  *wRatio = 0.0;
  if (bx.nMols == 0) return 0;
  const int nb = (bx.nMols + 255) / 256;
  CK(e->blockA.reserve(nb + 1024));
  const int fw = moveType == 1 ? GOMCB200_MOL_TORQUE : GOMCB200_MOL_FORCE;
  const int rw = GOMCB200_MOL_FORCE_REC;
  const bool rec = moveType == 0;
  k_bm_coeff<<<nb, 256, 0, e->stream>>>(
      bx.nMols, bx.molList.p, max, BETA, e->forceT[fw][0].p, e->forceT[fw][1].p,
      e->forceT[fw][2].p, rec ? e->forceT[rw][0].p : nullptr, rec ? e->forceT[rw][1].p : nullptr,
      rec ? e->forceT[rw][2].p : nullptr, e->force[fw][0].p, e->force[fw][1].p,
      e->force[fw][2].p, rec ? e->force[rw][0].p : nullptr, rec ? e->force[rw][1].p : nullptr,
      rec ? e->force[rw][2].p : nullptr, e->mpK[0].p, e->mpK[1].p, e->mpK[2].p, e->blockA.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nb, 1, e->blockA.p, nullptr, nullptr, nullptr,
                                           e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *wRatio = e->hRes[0];
  return 0;
}

// ---- virial -----------------------------------------------------------------
int gomcb200_box_inter_virial(gomcb200_engine *e, int box, double vT[3], double rT[3]) {
  GB_RANGE("energy_box_virial");
  int rc = check_box(e, box);
  if (rc) return rc;
  if (!vT || !rT) return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  rc = run_pair(e, box, MODE_VIRIAL);
  if (rc) return rc;
  rc = fetch_result(e, 6);
  if (rc) return rc;
  for (int c = 0; c < 3; ++c) {
    vT[c] = e->hRes[c];
    rT[c] = e->hRes[3 + c] * kQQFact;  // src/CalculateEnergy.cpp:555-567
  }
  return 0;
}

int gomcb200_virial_reciprocal(gomcb200_engine *e, int box, double wT[3]) {
  GB_RANGE("ewald_box_recip_virial");
  int rc = check_box(e, box);
  if (rc) return rc;
  rc = check_unsharded(e, "gomcb200_virial_reciprocal");
  if (rc) return rc;
  if (!wT) return fail(GOMCB200_EINVAL, "bad arguments");
  wT[0] = wT[1] = wT[2] = 0.0;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[1 - bx.cur];
  if (!(e->ewald && e->electrostatic) || ks.n == 0 || bx.nAtoms == 0) return 0;
  rc = ensure_sums(e, bx, ks.n);
  if (rc) return rc;
  const double *sR = bx.sum[bx.iRref].p, *sI = bx.sum[bx.iIref].p;
  const int gk = (ks.n + 255) / 256, ga = (bx.nAtoms + 255) / 256;
  CK(e->blockA.reserve(3 * (size_t)gk + 1024));
  CK(e->blockB.reserve(3 * (size_t)ga + 1024));
  for (int c = 0; c < 3; ++c) CK(e->scratchF[c].reserve(e->nAtoms + 1));
  const double constVal = 1.0 / (4.0 * (e->alpha[box] * e->alpha[box]));
  k_virial_recip_k<<<gk, 256, 0, e->stream>>>(ks.n, constVal, ks.kx.p, ks.ky.p, ks.kz.p,
                                             ks.hsqr.p, ks.prefact.p, sR, sI, e->blockA.p);
  e->launches += 1;
  rc = run_force_recip(e, box, sR, sI, e->scratchF[0].p, e->scratchF[1].p, e->scratchF[2].p,
                       false);
  if (rc) return rc;
  k_virial_recip_intra<<<ga, 256, 0, e->stream>>>(make_params(e, box), bx.nAtoms, bx.atomList.p,
                                                 e->mol.p, e->x.p, e->y.p, e->z.p, e->qEff.p,
                                                 e->scratchF[0].p, e->scratchF[1].p,
                                                 e->scratchF[2].p, e->blockB.p);
  const double *a = e->blockA.p, *b = e->blockB.p;
  k_final_reduce<<<1, 1024, 0, e->stream>>>(gk, 3, a, a + gk, a + 2 * (size_t)gk, nullptr,
                                           e->result.p);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(ga, 3, b, b + ga, b + 2 * (size_t)ga, nullptr,
                                           e->result.p + 3);
  e->launches += 3;
  CK(cudaGetLastError());
  rc = fetch_result(e, 6);
  if (rc) return rc;
  for (int c = 0; c < 3; ++c) wT[c] = e->hRes[c] + e->hRes[3 + c];
  return 0;
}

int gomcb200_get_recip_sums(gomcb200_engine *e, int box, int which, double *sumR, double *sumI,
                            int n) {
  if (!e || box < 0 || box >= e->nBoxes || n < 0) return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  if ((size_t)n > bx.sum[0].cap) return fail(GOMCB200_EINVAL, "n exceeds allocated sums");
  CK(cudaStreamSynchronize(e->stream));
  int ir = which == GOMCB200_SUM_NEW ? bx.iRnew : bx.iRref;
  int ii = which == GOMCB200_SUM_NEW ? bx.iInew : bx.iIref;
  if (sumR) CK(cudaMemcpy(sumR, bx.sum[ir].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (sumI) CK(cudaMemcpy(sumI, bx.sum[ii].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int gomcb200_set_recip_ref(gomcb200_engine *e, int box) {
  // Ewald::SetRecipRef, src/Ewald.cpp:1021-1053: sums new -> ref, k new -> Ref
  if (!e || box < 0 || box >= e->nBoxes) return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &src = bx.kset[bx.cur], &dst = bx.kset[1 - bx.cur];
  int rc = ensure_sums(e, bx, src.n);
  if (rc) return rc;
  size_t bytes = sizeof(double) * (size_t)src.n;
  if (src.n) {
    CK(cudaMemcpyAsync(bx.sum[bx.iRref].p, bx.sum[bx.iRnew].p, bytes, cudaMemcpyDeviceToDevice,
                       e->stream));
    CK(cudaMemcpyAsync(bx.sum[bx.iIref].p, bx.sum[bx.iInew].p, bytes, cudaMemcpyDeviceToDevice,
                       e->stream));
  }
  dst.hkx = src.hkx; dst.hky = src.hky; dst.hkz = src.hkz; dst.hhsqr = src.hhsqr;
  dst.hprefact = src.hprefact;
  for (int d = 0; d < 3; ++d) { dst.nmax[d] = src.nmax[d]; dst.cv[d] = src.cv[d]; dst.L[d] = src.L[d]; }
  dst.kmax = src.kmax;
  rc = upload_kset(e, dst);
  if (rc) return rc;
  // duplicate the plan (small)
  dst.planValid = false;
  if (src.planValid) {
    size_t nr = src.rows.cap, nt = src.tiles.cap;
    CK(dst.rows.reserve(nr));
    CK(dst.tiles.reserve(nt));
    CK(cudaMemcpyAsync(dst.rows.p, src.rows.p, std::min(nr, dst.rows.cap) * sizeof(int4),
                       cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(dst.tiles.p, src.tiles.p, std::min(nt, dst.tiles.cap) * sizeof(int4),
                       cudaMemcpyDeviceToDevice, e->stream));
    dst.nTiles = src.nTiles;
    dst.nRowsPadded = src.nRowsPadded;
    dst.maxRows = src.maxRows;
    dst.planValid = true;
  }
  dst.hRowsSorted = src.hRowsSorted;
  dst.ng = src.ng;
  dst.ngValid = src.ngValid && dst.planValid;
  dst.mmaValid = src.mmaValid;
  dst.tilesForShard = -1;
  dst.itemsForAtoms = -1;
  dst.fmValid = false;
  if (src.fmValid) {
    CK(dst.fmRows.reserve(src.fmRows.cap));
    CK(dst.fmTiles.reserve(src.fmTiles.cap));
    CK(dst.fmW.reserve(src.fmW.cap));
    CK(cudaMemcpyAsync(dst.fmRows.p, src.fmRows.p,
                       std::min(src.fmRows.cap, dst.fmRows.cap) * sizeof(int4),
                       cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(dst.fmTiles.p, src.fmTiles.p,
                       std::min(src.fmTiles.cap, dst.fmTiles.cap) * sizeof(FmTile),
                       cudaMemcpyDeviceToDevice, e->stream));
    CK(dst.fmBlocks.reserve(src.fmBlocks.cap));
    CK(cudaMemcpyAsync(dst.fmBlocks.p, src.fmBlocks.p,
                       std::min(src.fmBlocks.cap, dst.fmBlocks.cap) * sizeof(FmBlock),
                       cudaMemcpyDeviceToDevice, e->stream));
    dst.fmNBlocks = src.fmNBlocks;
    dst.fmNTiles = src.fmNTiles;
    dst.fmValid = true;
  }
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int gomcb200_copy_recip(gomcb200_engine *e, int box) {
  // Ewald::CopyRecip, src/Ewald.cpp:1441-1460: ref -> new
  if (!e || box < 0 || box >= e->nBoxes) return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  int n = bx.kset[1 - bx.cur].n;
  int rc = ensure_sums(e, bx, n);
  if (rc) return rc;
  if (n) {
    size_t bytes = sizeof(double) * (size_t)n;
    CK(cudaMemcpyAsync(bx.sum[bx.iRnew].p, bx.sum[bx.iRref].p, bytes, cudaMemcpyDeviceToDevice,
                       e->stream));
    CK(cudaMemcpyAsync(bx.sum[bx.iInew].p, bx.sum[bx.iIref].p, bytes, cudaMemcpyDeviceToDevice,
                       e->stream));
  }
  return 0;
}

int gomcb200_update_recip(gomcb200_engine *e, int box) {
  // Ewald::UpdateRecip, src/Ewald.cpp:1420-1434: O(1) swap new <-> ref
  if (!e || box < 0 || box >= e->nBoxes) return fail(GOMCB200_EINVAL, "bad arguments");
  BoxState &bx = e->box[box];
  std::swap(bx.iRnew, bx.iRref);
  std::swap(bx.iInew, bx.iIref);
  return 0;
}

int gomcb200_update_recip_vec(gomcb200_engine *e, int box) {
  // Ewald::UpdateRecipVec, src/Ewald.cpp:1462-1487: O(1) swap k <-> kRef
  if (!e || box < 0 || box >= e->nBoxes) return fail(GOMCB200_EINVAL, "bad arguments");
  e->box[box].cur = 1 - e->box[box].cur;
  return 0;
}

int gomcb200_box_self_correction(gomcb200_engine *e, int box, double *self, double *correction) {
  GB_RANGE("ewald_box_self_energy");
  int rc = check_box(e, box);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  if (bx.nMols == 0) {
    if (self) *self = 0.0;
    if (correction) *correction = 0.0;
    return 0;
  }
  BoxParams p = make_params(e, box);
  int nBlocks = (bx.nMols + 255) / 256;
  CK(e->blockA.reserve(nBlocks + 1024));
  CK(e->blockB.reserve(nBlocks + 1024));
  const bool selfQuirk = bx.lambdaMol >= 0 && bx.lambdaMol == bx.lambdaMolKind;
  k_self_correction<<<nBlocks, 256, 0, e->stream>>>(
      p, bx.nMols, bx.molList.p, e->molStart.p, e->x.p, e->y.p, e->z.p, e->q.p, e->blockA.p,
      e->blockB.p, selfQuirk ? bx.lambdaMol : -1, bx.lambdaCoulomb, bx.lambdaCoulomb);
  k_final_reduce<<<1, 1024, 0, e->stream>>>(nBlocks, 2, e->blockA.p, e->blockB.p, nullptr,
                                           nullptr, e->result.p);
  e->launches += 2;
  CK(cudaGetLastError());
  rc = fetch_result(e, 2);
  if (rc) return rc;
  // src/Ewald.cpp:1159 and :1084
  if (self) *self = e->hRes[0] * (-1.0 * p.alpha * kQQFact * kTwoOverSqrtPi * 0.5);
  if (correction) *correction = -1.0 * kQQFact * e->hRes[1];
  return 0;
}

// ---- literal drop-ins --------------------------------------------------------
int gomcb200_call_box_inter(gomcb200_engine *e, int box, const double *x, const double *y,
                            const double *z, const double axis[3], double *REn, double *LJEn) {
  GB_RANGE("energy_box_inter");
  int rc = 0;
  if (axis) rc = gomcb200_set_box_axes(e, box, axis);
  if (rc) return rc;
  rc = gomcb200_set_coords(e, x, y, z, 0, e ? e->nAtoms : 0);
  if (rc) return rc;
  return gomcb200_box_inter(e, box, LJEn, REn);
}

int gomcb200_call_box_reciprocal_sums(gomcb200_engine *e, int box, const double *x,
                                      const double *y, const double *z, double *energyRecip) {
  int rc = gomcb200_set_coords(e, x, y, z, 0, e ? e->nAtoms : 0);
  if (rc) return rc;
  return gomcb200_box_reciprocal_sums(e, box, energyRecip);
}

int gomcb200_call_box_force(gomcb200_engine *e, int box, const double *x, const double *y,
                            const double *z, const double axis[3], double *REn, double *LJEn,
                            double *aForcex, double *aForcey, double *aForcez,
                            double *mForcex, double *mForcey, double *mForcez) {
  GB_RANGE("energy_box_force");
  int rc = 0;
  if (axis) rc = gomcb200_set_box_axes(e, box, axis);
  if (rc) return rc;
  rc = gomcb200_set_coords(e, x, y, z, 0, e ? e->nAtoms : 0);
  if (rc) return rc;
  rc = gomcb200_box_force(e, box, LJEn, REn);
  if (rc) return rc;
  if (aForcex || aForcey || aForcez)
    rc = gomcb200_get_forces(e, GOMCB200_ATOM_FORCE, aForcex, aForcey, aForcez, 0, e->nAtoms);
  if (rc) return rc;
  if (mForcex || mForcey || mForcez)
    rc = gomcb200_get_forces(e, GOMCB200_MOL_FORCE, mForcex, mForcey, mForcez, 0, e->nMols);
  return rc;
}

static int full_box_energy_impl(gomcb200_engine *e, int box, const double *x, const double *y,
                                const double *z, double *LJEn, double *REn,
                                double *energyRecip) {
  int rc = check_box(e, box);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  if (x) {
    // no host wait between the upload and the kernels: this call synchronises before it
    // returns, which is as long as the caller's buffers have to stay as they are
    rc = set_coords_impl(e, x, y, z, 0, e->nAtoms, false);
    if (rc) return rc;
  }
  timing_begin(e);
  CK(e->energy3.reserve(4));
  BoxState &bx = e->box[box];
  const bool recipOn = e->ewald && e->electrostatic;
  // CUDA graph of the step: only the "coordinates changed" case (re-bin, re-pack), the one
  // every MC evaluation is
  gomcb200_engine::StepGraph &sg = e->stepGraph[box];
  bool done = false;
  // (not on a sharded engine: a graph replay of the engine's collectives next to a launcher's
  // own communicator stalled an 8-rank run; the plain path is what was validated there)
  if (e->useGraph && !e->comm && bx.cellsDirty && bx.packedDirty) {
    std::vector<unsigned char> key;
    step_key(e, box, key);
    if (sg.exec && key == sg.key) {
      CK(cudaGraphLaunch(sg.exec, e->stream));
      bx.cellsDirty = bx.packedDirty = false;
      bx.sumsComplete = sg.sumsComplete;
      e->launches += sg.launches;
      done = true;
    } else {
      if (sg.exec) {
        cudaGraphExecDestroy(sg.exec);
        sg.exec = nullptr;
      }
      sg.warm = key == sg.key ? sg.warm + 1 : 0;
      sg.key = key;
      if (sg.warm >= 2) {  // buffers and tables have settled: capture this call
        const long long l0 = e->launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        e->capturing = true;
        rc = enqueue_full_box(e, box, recipOn);
        e->capturing = false;
        cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
        if (!rc && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&sg.exec, graph, 0) == cudaSuccess) {
          sg.launches = e->launches - l0;
          sg.sumsComplete = bx.sumsComplete;
          std::vector<unsigned char> after;
          step_key(e, box, after);
          if (after == key) {  // nothing was (re)allocated while capturing
            cudaGraphDestroy(graph);
            graph = nullptr;
            CK(cudaGraphLaunch(sg.exec, e->stream));
            done = true;
            if (getenv("GOMCB200_GRAPH_VERBOSE"))
              fprintf(stderr, "[gomc_b200] box %d: step captured as a CUDA graph (%lld launches)\n",
                      box, sg.launches);
          } else {
            cudaGraphExecDestroy(sg.exec);
            sg.exec = nullptr;
          }
        }
        if (!done) {  // not capturable here: run the plain path from now on
          if (graph) cudaGraphDestroy(graph);
          const cudaError_t why = cudaGetLastError();
          if (getenv("GOMCB200_GRAPH_VERBOSE"))
            fprintf(stderr, "[gomc_b200] box %d: graph capture failed (rc %d, end %s, last %s)\n",
                    box, rc, cudaGetErrorString(ce), cudaGetErrorString(why));
          e->useGraph = false;
          sg.exec = nullptr;
          e->launches = l0;
          bx.cellsDirty = bx.packedDirty = true;
          rc = 0;
        }
      }
    }
  }
  if (!done) {
    rc = enqueue_full_box(e, box, recipOn);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(e->stream));
  const double recip = e->hRes[10];
  timing_end(e, recipOn);
  if (LJEn) *LJEn = e->hRes[8];
  if (REn) *REn = e->hRes[9];
  if (energyRecip) *energyRecip = recip;
  return 0;
}

int gomcb200_call_full_box_energy(gomcb200_engine *e, int box, const double *x,
                                  const double *y, const double *z, double *LJEn, double *REn,
                                  double *energyRecip) {
  GB_RANGE("energy_system_total(inter,recip)");
  const int rc = full_box_energy_impl(e, box, x, y, z, LJEn, REn, energyRecip);
  // the coordinate upload is asynchronous: on an early error return it may still be reading
  // the caller's buffers
  if (rc && e && x && e->stream) cudaStreamSynchronize(e->stream);
  return rc;
}


// ---- literal drop-ins of the reciprocal seam: explicit charges / k list in ----------
int gomcb200_set_kvectors(gomcb200_engine *e, int box, int n, const double *kx, const double *ky,
                          const double *kz, const double *hsqr, const double *prefact) {
  if (!e || box < 0 || box >= e->nBoxes || n < 0 || (n && (!kx || !ky || !kz || !hsqr || !prefact)))
    return fail(GOMCB200_EINVAL, "bad arguments");
  if (e->imageTotal > 0 && n > e->imageTotal)
    return fail(GOMCB200_EKMAX, "k list of %d entries exceeds imageTotal %d", n, e->imageTotal);
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[bx.cur];
  ks.hkx.assign(kx, kx + n); ks.hky.assign(ky, ky + n); ks.hkz.assign(kz, kz + n);
  ks.hhsqr.assign(hsqr, hsqr + n); ks.hprefact.assign(prefact, prefact + n);
  ks.planValid = ks.mmaValid = ks.fmValid = ks.ngValid = false;
  // Recover the integer structure of Ewald::RecipInitOrth's list (src/Ewald.cpp:865-896):
  // entries are ordered by (a, b, c) with k = (a cvx, b cvy, c cvz); the first entry is
  // (0, 0, 1), the first with ky != 0 is (0, 1, .), the first with kx != 0 is (1, ., .).
  std::vector<RowRec> rows;
  bool ok = !bx.nonOrth && n > 0 && kx[0] == 0.0 && ky[0] == 0.0 && kz[0] > 0.0;
  double cv[3] = {0.0, 0.0, kz ? (n ? kz[0] : 0.0) : 0.0};
  if (ok) {
    for (int i = 0; i < n && (cv[0] == 0.0 || cv[1] == 0.0); ++i) {
      if (cv[1] == 0.0 && kx[i] == 0.0 && ky[i] != 0.0) cv[1] = ky[i];
      if (cv[0] == 0.0 && kx[i] != 0.0) cv[0] = kx[i];
    }
    ok = cv[0] > 0.0 && cv[1] > 0.0;
  }
  int nm[3] = {0, 0, 0};
  if (ok) {
    int pa = 0, pb = 0, pc = 0;
    for (int i = 0; i < n && ok; ++i) {
      const int a = (int)std::lround(kx[i] / cv[0]), b = (int)std::lround(ky[i] / cv[1]),
                c = (int)std::lround(kz[i] / cv[2]);
      // the list must be exactly the products the device regenerates
      ok = kx[i] == cv[0] * a && ky[i] == cv[1] * b && kz[i] == cv[2] * c;
      if (i == 0 || a != pa || b != pb) {
        rows.push_back({a, b, c, i});
      } else if (c != pc + 1) {
        ok = false;
      }
      rows.back().cmax = c;  // last c of the row
      pa = a; pb = b; pc = c;
      nm[0] = std::max(nm[0], std::abs(a));
      nm[1] = std::max(nm[1], std::abs(b));
      nm[2] = std::max(nm[2], std::abs(c));
    }
    for (size_t r = 0; r < rows.size() && ok; ++r) {
      const int first = rows[r].start;
      const int cnt = (r + 1 < rows.size() ? rows[r + 1].start : n) - first;
      const int c0 = (int)std::lround(kz[first] / cv[2]);
      const bool origin = rows[r].a == 0 && rows[r].b == 0;
      ok = origin ? (c0 == 1 && cnt == rows[r].cmax) : (c0 == -rows[r].cmax && cnt == 2 * rows[r].cmax + 1);
    }
  }
  for (int d = 0; d < 3; ++d) {
    ks.cv[d] = ok ? cv[d] : 0.0;
    ks.nmax[d] = ok ? nm[d] : 0;
    ks.L[d] = ok ? (2.0 * M_PI) / cv[d] : bx.axis[d];
  }
  ks.kmax = std::max(nm[0], std::max(nm[1], nm[2]));
  int rc = upload_kset(e, ks, false);  // the host's own values, prefactor included
  if (rc) return rc;
  if (ok && !rows.empty()) {
    rc = build_plan(e, ks, rows);
    if (rc) return rc;
    ks.ngValid = gbn::nufft_choose(ks.nmax, &ks.ng) == 0;
  }
  return 0;
}

int gomcb200_set_recip_sums(gomcb200_engine *e, int box, int which, const double *sumR,
                            const double *sumI, int n) {
  if (!e || box < 0 || box >= e->nBoxes || n < 0 || !sumR || !sumI)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  int rc = ensure_sums(e, bx, n);
  if (rc) return rc;
  const int ir = which == GOMCB200_SUM_NEW ? bx.iRnew : bx.iRref;
  const int ii = which == GOMCB200_SUM_NEW ? bx.iInew : bx.iIref;
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(bx.sum[ir].p, sumR, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(bx.sum[ii].p, sumI, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  return 0;
}

int gomcb200_set_forces(gomcb200_engine *e, int which, const double *x, const double *y,
                        const double *z, int first, int count) {
  if (!e || !e->haveTopo || which < 0 || which > 4) return fail(GOMCB200_EINVAL, "bad arguments");
  const int limit = (which == GOMCB200_ATOM_FORCE || which == GOMCB200_ATOM_FORCE_REC) ? e->nAtoms
                                                                                       : e->nMols;
  CK(cudaSetDevice(e->device));
  return upload3(e, e->force[which][0], e->force[which][1], e->force[which][2], x, y, z, first,
                 count, limit);
}

static int download_sums(gomcb200_engine *e, BoxState &bx, int n, double *sumR, double *sumI) {
  if (sumR) CK(cudaMemcpy(sumR, bx.sum[bx.iRnew].p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  if (sumI) CK(cudaMemcpy(sumI, bx.sum[bx.iInew].p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return 0;
}

int gomcb200_call_box_reciprocal_points(gomcb200_engine *e, int box, int newSet, int n,
                                        const double *x, const double *y, const double *z,
                                        const double *q, double *sumRnew, double *sumInew,
                                        double *energyRecip) {
  int rc = check_box(e, box, false, false);  // explicit charges: no topology needed
  if (rc) return rc;
  if (n < 0 || (n && (!x || !y || !z || !q))) return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  KSet &ks = bx.kset[newSet ? bx.cur : 1 - bx.cur];
  // the charged points as the packed {x, y, z, q} array the structure-factor kernels read
  rc = stage_reserve(e, sizeof(double4) * (size_t)(n + 1));
  if (rc) return rc;
  double4 *h = reinterpret_cast<double4 *>(e->hStage);
  int m = 0;
  double qmax = 0.0;
  for (int i = 0; i < n; ++i) {
    if (std::fabs(q[i]) < 0.000000001) continue;  // particleHasNoCharge, src/Ewald.cpp:107-111
    h[m++] = make_double4(x[i], y[i], z[i], q[i]);
    qmax = std::max(qmax, std::fabs(q[i]));
  }
  CK(bx.packed.reserve(m + 1));
  if (m)
    CK(cudaMemcpyAsync(bx.packed.p, h, sizeof(double4) * (size_t)m, cudaMemcpyHostToDevice,
                       e->stream));
  const int savedCharged = bx.nCharged;
  const double savedQmax = bx.qMaxAbs;
  bx.nCharged = m;
  bx.qMaxAbs = std::max(qmax, 1e-300);
  bx.packedDirty = false;
  rc = run_recip_sums(e, box, ks);
  bx.nCharged = savedCharged;
  bx.qMaxAbs = savedQmax;
  bx.packedDirty = true;  // the resident atoms are packed again by the next resident call
  if (rc) return rc;
  rc = fetch_result(e, 1);
  if (rc) return rc;
  if (energyRecip) *energyRecip = e->hRes[0];
  return download_sums(e, bx, ks.n, sumRnew, sumInew);
}

int gomcb200_call_mol_reciprocal(gomcb200_engine *e, int box, int len, const double *q,
                                 const double *oldX, const double *oldY, const double *oldZ,
                                 const double *newX, const double *newY, const double *newZ,
                                 double *sumRnew, double *sumInew, double *energyRecipNew) {
  int rc = check_box(e, box, false, false);  // explicit charges: no topology needed
  if (rc) return rc;
  if (len < 1 || !q || !oldX || !oldY || !oldZ || !newX || !newY || !newZ || !energyRecipNew)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  const int nk = bx.kset[1 - bx.cur].n;
  *energyRecipNew = 0.0;
  if (nk == 0) return 0;
  rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  rc = stage_molbuf_raw(e, len, q, q, newX, newY, newZ, oldX, oldY, oldZ, 0, 0);
  if (rc) return rc;
  rc = launch_mol_recip(e, box, len, 0);
  if (rc) return rc;
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *energyRecipNew = e->hRes[0];
  return download_sums(e, bx, nk, sumRnew, sumInew);
}

int gomcb200_call_swap_reciprocal(gomcb200_engine *e, int box, int len, const double *q,
                                  const double *x, const double *y, const double *z, int insert,
                                  double *sumRnew, double *sumInew, double *energyRecipNew) {
  int rc = check_box(e, box, false, false);  // explicit charges: no topology needed
  if (rc) return rc;
  if (len < 1 || !q || !x || !y || !z || !energyRecipNew)
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  BoxState &bx = e->box[box];
  const int nk = bx.kset[1 - bx.cur].n;
  *energyRecipNew = 0.0;
  if (nk == 0) return 0;
  rc = ensure_sums(e, bx, nk);
  if (rc) return rc;
  rc = stage_molbuf_raw(e, len, q, q, x, y, z, nullptr, nullptr, nullptr, insert ? 1 : 2, 0);
  if (rc) return rc;
  rc = launch_mol_recip(e, box, len, insert ? 1 : 2);
  if (rc) return rc;
  rc = fetch_result(e, 1);
  if (rc) return rc;
  *energyRecipNew = e->hRes[0];
  return download_sums(e, bx, nk, sumRnew, sumInew);
}

int gomcb200_set_shard(gomcb200_engine *e, int rank, int world) {
  if (!e || world < 1 || rank < 0 || rank >= world) return fail(GOMCB200_EINVAL, "bad arguments");
  e->shardRank = rank;
  e->shardWorld = world;
  return 0;
}

int gomcb200_comm_unique_id(void *id128) {
  if (!id128) return fail(GOMCB200_EINVAL, "bad arguments");
  std::string err;
  if (gbc::comm_unique_id(id128, err)) return fail(GOMCB200_ECUDA, "%s", err.c_str());
  return 0;
}

int gomcb200_set_comm(gomcb200_engine *e, const void *id128, int rank, int world) {
  if (!e || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128))
    return fail(GOMCB200_EINVAL, "bad arguments");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  if (e->comm) {
    gbc::comm_destroy(e->comm);
    e->comm = nullptr;
  }
  if (world > 1) {
    std::string err;
    e->comm = gbc::comm_create(id128, rank, world, err);
    if (!e->comm) return fail(GOMCB200_ECUDA, "%s", err.c_str());
  }
  e->shardRank = rank;
  e->shardWorld = world;
  return 0;
}

int gomcb200_mark_coords_changed(gomcb200_engine *e) {
  if (!e) return fail(GOMCB200_EINVAL, "null engine");
  mark_coords_dirty(e);
  return 0;
}

int gomcb200_set_recip_algo(gomcb200_engine *e, int algo) {
  if (!e || algo < 0 || algo > 5) return fail(GOMCB200_EINVAL, "bad arguments");
  e->recipAlgo = algo;
  return 0;
}

int gomcb200_set_pair_algo(gomcb200_engine *e, int algo) {
  if (!e || algo < 0 || algo > 1) return fail(GOMCB200_EINVAL, "bad arguments");
  e->pairAlgo = algo;
  return 0;
}

int gomcb200_set_recip_auto_work(gomcb200_engine *e, double work) {
  if (!e || !(work >= 0.0)) return fail(GOMCB200_EINVAL, "bad arguments");
  e->recipAutoWork = work;
  return 0;
}

int gomcb200_enable_timing(gomcb200_engine *e, int on) {
  if (!e) return fail(GOMCB200_EINVAL, "null engine");
  e->timing = on != 0;
  return 0;
}

int gomcb200_last_timing(const gomcb200_engine *e, float *totalMs, float *dominantMs) {
  if (!e) return fail(GOMCB200_EINVAL, "null engine");
  if (totalMs) *totalMs = e->lastTotalMs;
  if (dominantMs) *dominantMs = e->lastDominantMs;
  return 0;
}

}  // extern "C"
