// pair.cuh -- cell-list Lennard-Jones/Mie + real-space Coulomb kernels.
//
// Replaces BoxInterGPU / BoxForceGPU (src/GPU/CalculateEnergyCUDAKernel.cu:149,
// src/GPU/CalculateForceCUDAKernel.cu:601) and gives device versions of
// CalculateEnergy::MoleculeInter / ParticleInter (host-only in the reference).
//
// Design (B200): atoms are kept in cell-sorted SoA (x,y,z,q double; kind,mol
// packed int2).  One CTA owns a slice of the i-atoms of one cell and first
// stages the atoms of all 27 neighbour cells into shared memory (~150 KB for a
// water box at rc = 10 A, read once from L2 with coalesced 8-byte loads).  Each
// warp then owns one i-atom at a time: its 32 lanes test 32 different j-atoms
// per round (cheap distance test), hits are compacted with ballot/popc into a
// per-warp ring buffer, and whenever 32 hits are queued every lane evaluates
// one full pair (erfc, Mie) -- so the expensive path always runs with full
// lanes instead of the ~15 % in-range fraction.  Per-atom forces are summed in
// registers and reduced with a fixed shuffle tree: no atomics, bit-reproducible.
#pragma once
#include "common.cuh"

namespace gb {

struct CellGrid {
  int edge[3];
  int nCells;
  double cellSize[3];
  int generic[3];  // 1: fewer than 4 cells on that axis -> per-pair min-image
  int nonOrth;     // cells are assigned in unslant coordinates (src/CellList.h:88-101)
  double Bi[9];
};

// CellList::PositionToCell, src/CellList.h:88-101 (orthogonal box).
__host__ __device__ __forceinline__ int position_to_cell(const CellGrid &g, double x,
                                                double y, double z) {
  if (g.nonOrth) {  // BoxDimensionsNonOrth::TransformUnSlant
    double ux = x * g.Bi[0] + y * g.Bi[3] + z * g.Bi[6];
    double uy = x * g.Bi[1] + y * g.Bi[4] + z * g.Bi[7];
    double uz = x * g.Bi[2] + y * g.Bi[5] + z * g.Bi[8];
    x = ux;
    y = uy;
    z = uz;
  }
  int cx = (int)(x / g.cellSize[0]);
  int cy = (int)(y / g.cellSize[1]);
  int cz = (int)(z / g.cellSize[2]);
  cx -= (cx == g.edge[0] ? 1 : 0);
  cy -= (cy == g.edge[1] ? 1 : 0);
  cz -= (cz == g.edge[2] ? 1 : 0);
  return cx * g.edge[1] * g.edge[2] + cy * g.edge[2] + cz;
}

__global__ void k_cell_keys(CellGrid g, int n, const int *__restrict__ atomList,
                            const double *__restrict__ x,
                            const double *__restrict__ y,
                            const double *__restrict__ z, int *keys,
                            int *vals) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int a = atomList[t];
  keys[t] = position_to_cell(g, x[a], y[a], z[a]);
  vals[t] = a;
}

// cellStart[c] = first sorted position whose key >= c (lower bound); *maxPop (zeroed by the
// caller) = largest cell population.  One thread per sorted position t (and one past the end):
// it fills cellStart for every cell in (key[t-1], key[t]] -- no search, one coalesced pass --
// and, where a cell ends, reports its population.
__global__ void k_cell_bounds(int nCells, int n,
                              const int *__restrict__ sortedKeys,
                              int *cellStart, int *maxPop) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  const int prev = t == 0 ? -1 : sortedKeys[t - 1];
  const int cur = t == n ? nCells : sortedKeys[t];
  for (int c = prev + 1; c <= cur; ++c) cellStart[c] = t;
  if (t > 0 && cur != prev) {
    // the run of key `prev` ends at t; its start = lower bound of that key (binary search,
    // only the ~nCells threads that sit on a run end do it)
    int lo = 0, hi = t - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sortedKeys[mid] < prev)
        lo = mid + 1;
      else
        hi = mid;
    }
    atomicMax(maxPop, t - lo);
  }
}

__global__ void k_gather_sorted(int n, const int *__restrict__ sortedAtoms,
                                const double *__restrict__ x,
                                const double *__restrict__ y,
                                const double *__restrict__ z,
                                const double *__restrict__ q,
                                const int *__restrict__ kind,
                                const int *__restrict__ mol, double *sx,
                                double *sy, double *sz, double *sq, int2 *skm,
                                int *sortedPos) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int a = sortedAtoms[t];
  sortedPos[a] = t;
  sx[t] = x[a];
  sy[t] = y[a];
  sz[t] = z[a];
  sq[t] = q[a];
  skm[t] = make_int2(kind[a], mol[a]);
}

// ---------------------------------------------------------------------------
// A neighbour "range": j-atoms [begin,end) in the j arrays (global sorted
// arrays or their shared-memory staging), the global sorted index of `begin`,
// and the periodic shift to add to (xi - xj) -- exactly what MinImageSigned
// would add when the axis has >= 4 cells; 0 plus a per-pair test otherwise.
struct JRange {
  int begin, end, gbase, isSelf;
  double sx, sy, sz;
};

constexpr int kQueue = 64;  // per-warp ring buffer entries (power of two)

struct WarpQueue {
  int j[kQueue];
  double dx[kQueue], dy[kQueue], dz[kQueue];
};

struct PairAcc {
  double lj, real, fx, fy, fz;
  int overlap;
  // virial mode only: sixth accumulator and the COM of the i atom's molecule
  double ex, cix, ciy, ciz;
};
// pair evaluation modes: energies; energies + forces; virial tensors
// (acc.lj/real/fx = LJ diag 11/22/33, acc.fy/fz/ex = Coulomb diag 11/22/33)
// MODE_LAMBDA is OR-ed in when the box has a fractional molecule: only those
// instantiations carry the soft-core branch, the usual kernels stay as lean as before.
enum { MODE_ENERGY = 0, MODE_FORCE = 1, MODE_VIRIAL = 2, MODE_LAMBDA = 4 };

enum { SWEEP_PROBE = 0, SWEEP_HALF = 1, SWEEP_FULL = 2 };

template <int VDW, int MODEL>
__device__ __forceinline__ void eval_pair(const BoxParams &p, int ki, double qi,
                                          int excludeMol, bool countEnergy,
                                          double sign, bool checkOverlap,
                                          double dx, double dy, double dz,
                                          double qj, int2 kmj, PairAcc &acc) {
  if (kmj.y == excludeMol) return;
  double r2 = dist_sq(dx, dy, dz);
  if (checkOverlap && r2 < p.rCutLowSq) acc.overlap = 1;
  int idx = ki + kmj.x * p.kindCount;
  constexpr int MODE = MODEL & 3;
  constexpr bool LAM = (MODEL & MODE_LAMBDA) != 0;
  // fractional molecule: GetLambdaVDW / GetLambdaCoulomb (src/CalculateEnergy.cpp:1558-1572);
  // excludeMol is the molecule of the i atom / probe in every sweep
  if (LAM && p.lambdaMol >= 0 && (excludeMol == p.lambdaMol || kmj.y == p.lambdaMol)) {
    const double lv = p.lambdaVDW, lc = p.lambdaCoulomb;
    if (MODE == MODE_ENERGY) {
      if (p.electrostatic) {
        double qq = qi * qj * kQQFact;
        if (qq != 0.0) acc.real += sign * calc_coulomb_l<VDW>(p, r2, idx, qq, lc);
      }
      acc.lj += sign * calc_en_l<VDW>(p, r2, idx, lv);
      return;
    }
    double eL, wL, eC = 0.0, wC = 0.0;
    calc_en_vir_l<VDW>(p, r2, idx, lv, eL, wL);
    if (p.electrostatic) {
      double qq = qi * qj * (MODE == MODE_VIRIAL ? 1.0 : kQQFact);
      if (qq != 0.0) calc_coulomb_en_vir_l<VDW>(p, r2, idx, qq, lc, eC, wC);
    }
    if (MODE == MODE_VIRIAL) {
      double cx = acc.cix - p.comx[kmj.y], cy = acc.ciy - p.comy[kmj.y],
             cz = acc.ciz - p.comz[kmj.y];
      min_image_vec(p, cx, cy, cz);
      const double t1 = dx * cx, t2 = dy * cy, t3 = dz * cz;
      acc.lj += wL * t1;
      acc.real += wL * t2;
      acc.fx += wL * t3;
      acc.fy += wC * t1;
      acc.fz += wC * t2;
      acc.ex += wC * t3;
    } else {
      if (countEnergy) {
        acc.lj += eL;
        acc.real += eC;
      }
      double w = wL + wC;
      acc.fx += dx * w;
      acc.fy += dy * w;
      acc.fz += dz * w;
    }
    return;
  }
  if (MODE == MODE_VIRIAL) {
    // CalculateEnergy::VirialCalc, src/CalculateEnergy.cpp:483-527
    double cx = acc.cix - p.comx[kmj.y], cy = acc.ciy - p.comy[kmj.y],
           cz = acc.ciz - p.comz[kmj.y];
    min_image_vec(p, cx, cy, cz);
    double eL, wL, eC = 0.0, wC = 0.0;
    calc_en_vir<VDW>(p, r2, idx, eL, wL);
    if (p.electrostatic) {
      double qq = qi * qj;  // qqFact multiplies the finished tensor (:555-567)
      if (qq != 0.0) calc_coulomb_en_vir<VDW>(p, r2, qq, eC, wC);
    }
    const double t1 = dx * cx, t2 = dy * cy, t3 = dz * cz;
    acc.lj += wL * t1;
    acc.real += wL * t2;
    acc.fx += wL * t3;
    acc.fy += wC * t1;
    acc.fz += wC * t2;
    acc.ex += wC * t3;
  } else if (MODE == MODE_FORCE) {
    double eL, wL, eC = 0.0, wC = 0.0;
    calc_en_vir<VDW>(p, r2, idx, eL, wL);
    if (p.electrostatic) {
      double qq = qi * qj * kQQFact;
      if (qq != 0.0) calc_coulomb_en_vir<VDW>(p, r2, qq, eC, wC);
    }
    if (countEnergy) {
      acc.lj += eL;
      acc.real += eC;
    }
    double w = wL + wC;
    acc.fx += dx * w;
    acc.fy += dy * w;
    acc.fz += dz * w;
  } else {
    if (p.electrostatic) {
      double qq = qi * qj * kQQFact;
      if (qq != 0.0) acc.real += sign * calc_coulomb<VDW>(p, r2, qq);
    }
    acc.lj += sign * calc_en<VDW>(p, r2, idx);
  }
}

// Where the j atoms live: global arrays, or their shared-memory staging given as
// 32-bit shared-window addresses (so that the loads are LDS, not generic LD).
struct JArrays {
  const double *x, *y, *z, *q;
  const int2 *km;
  unsigned sx, sy, sz, sq, skm;
};
template <bool SM>
__device__ __forceinline__ double jload(const double *g, unsigned s, int j) {
  return SM ? lds_f64(s + (unsigned)j * 8u) : g[j];
}
template <bool SM>
__device__ __forceinline__ int2 jload_km(const int2 *g, unsigned s, int j) {
  if (SM) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s + (unsigned)j * 8u));
    return v;
  }
  return g[j];
}

// One warp, one position (xi,yi,zi), ranges[rangeFirst::rangeStride].
//  SWEEP_PROBE: position is not part of the j set; every in-range j counts.
//  SWEEP_HALF : energy-only half shell; in the self range only j > selfIndex.
//  SWEEP_FULL : all neighbours, own force; energy counted for jGlobal > iGlobal.
// selfIndex is in the index space of the self range; iGlobal the global sorted
// index of the atom.  SM: j arrays staged in shared memory.  GEN: at least one
// axis has fewer than 4 cells (per-pair minimum image on that axis).
// qBase: shared address of this warp's WarpQueue.
template <int VDW, int MODE, int SWEEP, bool SM, bool GEN>
__device__ __forceinline__ void warp_probe(
    const BoxParams &p, const int generic[3], double xi, double yi, double zi,
    int ki, double qi, int excludeMol, int selfIndex, int iGlobal, double sign,
    bool checkOverlap, const JRange *ranges, int nRanges, int rangeStride,
    int rangeFirst, const JArrays &ja, unsigned qBase, PairAcc &acc) {
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;
  const unsigned qJ = qBase, qX = qBase + 4u * kQueue, qY = qX + 8u * kQueue,
                 qZ = qY + 8u * kQueue;
  int head = 0, count = 0;

  auto test = [&](const JRange &rg, int j, bool &in, int &jEnc, double &dx, double &dy,
                  double &dz) {
    in = false;
    jEnc = j;
    bool valid = j < rg.end;
    if (SWEEP == SWEEP_HALF) valid = valid && !(rg.isSelf && j <= selfIndex);
    if (SWEEP == SWEEP_FULL) valid = valid && !(rg.isSelf && j == selfIndex);
    if (valid) {
      dx = (xi - jload<SM>(ja.x, ja.sx, j)) + rg.sx;
      dy = (yi - jload<SM>(ja.y, ja.sy, j)) + rg.sy;
      dz = (zi - jload<SM>(ja.z, ja.sz, j)) + rg.sz;
      if (GEN) {
        if (p.nonOrth) {
          min_image_vec(p, dx, dy, dz);
        } else {
          if (generic[0]) dx = min_image(dx, p.ax[0], p.half[0]);
          if (generic[1]) dy = min_image(dy, p.ax[1], p.half[1]);
          if (generic[2]) dz = min_image(dz, p.ax[2], p.half[2]);
        }
      }
      double r2 = dist_sq(dx, dy, dz);
      in = p.boxRcutSq > r2;  // BoxDimensions::InRcut, strict
      if (SWEEP == SWEEP_FULL && (rg.gbase + (j - rg.begin)) > iGlobal)
        jEnc |= 0x40000000;  // this side of the pair owns the energy
    }
  };
  auto push = [&](bool in, int jEnc, double dx, double dy, double dz) {
    unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) {
      unsigned pos = (unsigned)((head + count + __popc(m & ltMask)) & (kQueue - 1));
      asm volatile("st.shared.s32 [%0], %1;" ::"r"(qJ + pos * 4u), "r"(jEnc) : "memory");
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(qX + pos * 8u), "d"(dx) : "memory");
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(qY + pos * 8u), "d"(dy) : "memory");
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(qZ + pos * 8u), "d"(dz) : "memory");
    }
    count += __popc(m);
  };
  auto drain = [&](bool active) {
    if (active) {
      unsigned pos = (unsigned)((head + lane) & (kQueue - 1));
      int je;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(je) : "r"(qJ + pos * 4u));
      int jj = je & 0x3fffffff;
      double ddx = lds_f64(qX + pos * 8u), ddy = lds_f64(qY + pos * 8u),
             ddz = lds_f64(qZ + pos * 8u);
      eval_pair<VDW, MODE>(p, ki, qi, excludeMol, SWEEP != SWEEP_FULL || (je & 0x40000000),
                            sign, checkOverlap, ddx, ddy, ddz, jload<SM>(ja.q, ja.sq, jj),
                            jload_km<SM>(ja.km, ja.skm, jj), acc);
    }
  };

  for (int r = rangeFirst; r < nRanges; r += rangeStride) {
    const JRange rg = ranges[r];
    for (int base = rg.begin; base < rg.end; base += 64) {
      // two candidates per lane per round: halves the loop / queue bookkeeping
      bool in0, in1;
      int e0, e1;
      double ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
      test(rg, base + lane, in0, e0, ax0, ay0, az0);
      test(rg, base + 32 + lane, in1, e1, ax1, ay1, az1);
      push(in0, e0, ax0, ay0, az0);
      if (count >= 32) {
        __syncwarp();
        drain(true);
        head = (head + 32) & (kQueue - 1);
        count -= 32;
        __syncwarp();
      }
      push(in1, e1, ax1, ay1, az1);
      __syncwarp();
      if (count >= 32) {
        drain(true);
        head = (head + 32) & (kQueue - 1);
        count -= 32;
        __syncwarp();
      }
    }
  }
  __syncwarp();
  drain(lane < count);
  __syncwarp();
}

// Neighbour ranges of a cell.  halfShell: own cell + the 13 "forward" cells
// (each unordered cell pair exactly once; needs >= 3 cells per axis, which
// CellList::ResizeGrid guarantees).
// Cooperative: thread t of the CTA fills range t (call from all threads, then
// __syncthreads()); returns the range count.
__device__ __forceinline__ int build_ranges(const CellGrid &g,
                                            const BoxParams &p, int cell,
                                            bool halfShell,
                                            const int *__restrict__ cellStart,
                                            JRange *ranges) {
  const int d0 = halfShell ? 13 : 0;
  const int d = d0 + (int)threadIdx.x;
  if (d < 27) {
    int cz = cell % g.edge[2];
    int cy = (cell / g.edge[2]) % g.edge[1];
    int cx = cell / (g.edge[2] * g.edge[1]);
    int dx = d / 9 - 1, dy = (d / 3) % 3 - 1, dz = d % 3 - 1;
    int nx = cx + dx, ny = cy + dy, nz = cz + dz;
    JRange r;
    r.sx = r.sy = r.sz = 0.0;
    r.isSelf = (d == 13);
    // raw = xi - xj; a j in a wrapped-below cell sits near +L: raw ~ -L -> +ax
    if (nx < 0) { nx += g.edge[0]; if (!g.generic[0]) r.sx = p.ax[0]; }
    else if (nx >= g.edge[0]) { nx -= g.edge[0]; if (!g.generic[0]) r.sx = -p.ax[0]; }
    if (ny < 0) { ny += g.edge[1]; if (!g.generic[1]) r.sy = p.ax[1]; }
    else if (ny >= g.edge[1]) { ny -= g.edge[1]; if (!g.generic[1]) r.sy = -p.ax[1]; }
    if (nz < 0) { nz += g.edge[2]; if (!g.generic[2]) r.sz = p.ax[2]; }
    else if (nz >= g.edge[2]) { nz -= g.edge[2]; if (!g.generic[2]) r.sz = -p.ax[2]; }
    int nc = nx * g.edge[1] * g.edge[2] + ny * g.edge[2] + nz;
    r.begin = cellStart[nc];
    r.end = cellStart[nc + 1];
    r.gbase = r.begin;
    ranges[d - d0] = r;
  }
  return 27 - d0;
}

constexpr int kMaxIPerPass = 512;  // i-atoms of a cell slice handled per pass
constexpr int kPairThreads = 256;
constexpr int kPairWarps = kPairThreads / 32;

// Full-box pair sweep.  grid = nCells * slices.  FORCE: all 27 neighbour
// cells, every atom accumulates its own force, energy counted for j > i.
// !FORCE: half shell (own cell with j > i plus 13 forward cells).
// Dynamic shared memory: staged j atoms (40 B each) when useSmem, else the j
// arrays are read from global memory (cells too full to stage).
template <int VDW, int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1)
    k_pair_box(BoxParams p, CellGrid g, int slices, int cell0, int useSmem,
               int smemAtoms,
               const int *__restrict__ cellStart, const double *__restrict__ sx,
               const double *__restrict__ sy, const double *__restrict__ sz,
               const double *__restrict__ sq, const int2 *__restrict__ skm,
               const int *__restrict__ sortedAtoms, double *__restrict__ partLJ,
               double *__restrict__ partReal, double *__restrict__ fx,
               double *__restrict__ fy, double *__restrict__ fz,
               const int *__restrict__ gatePop, int gateCap) {
  // launched behind k_pair_box2 (pair2.cuh) as its fallback: runs only when a cell is too
  // full for that kernel's staging pass (the same test makes k_pair_box2 return at once)
  if (gatePop && *gatePop + 2 <= gateCap) return;
  constexpr bool FORCE = (MODE & 3) == MODE_FORCE;
  constexpr bool VIRIAL = (MODE & 3) == MODE_VIRIAL;
  extern __shared__ __align__(16) unsigned char dynSmem[];
  __shared__ JRange ranges[27];
  __shared__ JRange stagedRanges[27];
  __shared__ double enLJ[kMaxIPerPass], enReal[kMaxIPerPass];
  __shared__ double enVir[VIRIAL ? 4 : 1][VIRIAL ? kMaxIPerPass : 1];
  __shared__ int nextI;
  // dynamic smem: per-warp hit queues, then the staged neighbour atoms
  WarpQueue *queues = reinterpret_cast<WarpQueue *>(dynSmem);
  unsigned char *stageBase = dynSmem + sizeof(WarpQueue) * NWARPS;

  const int cell = cell0 + blockIdx.x / slices, slice = blockIdx.x % slices;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int iBegin0 = cellStart[cell], iEnd0 = cellStart[cell + 1];
  const int nI = iEnd0 - iBegin0;
  const int iBegin = iBegin0 + (int)(((long long)nI * slice) / slices);
  const int iEnd = iBegin0 + (int)(((long long)nI * (slice + 1)) / slices);

  const int nRanges = build_ranges(g, p, cell, !FORCE, cellStart, ranges);
  __syncthreads();

  JArrays ja = {sx, sy, sz, sq, skm, 0u, 0u, 0u, 0u, 0u};
  int selfOffset = 0;  // self-range index = global sorted index + selfOffset
  const JRange *useRanges = ranges;
  // stage only if this CTA's neighbourhood fits (else the global-memory path)
  bool staged = false;
  if (useSmem && iEnd > iBegin) {
    int tot = 0;
    for (int r = 0; r < nRanges; ++r) tot += ranges[r].end - ranges[r].begin;
    staged = tot <= smemAtoms;
  }
  if (staged) {
    double *smx = reinterpret_cast<double *>(stageBase);
    double *smy = smx + smemAtoms;
    double *smz = smy + smemAtoms;
    double *smq = smz + smemAtoms;
    int2 *smkm = reinterpret_cast<int2 *>(smq + smemAtoms);
    int off = 0;
    for (int r = 0; r < nRanges; ++r) {
      const JRange rg = ranges[r];
      int len = rg.end - rg.begin;
      for (int t = threadIdx.x; t < len; t += blockDim.x) {
        smx[off + t] = sx[rg.begin + t];
        smy[off + t] = sy[rg.begin + t];
        smz[off + t] = sz[rg.begin + t];
        smq[off + t] = sq[rg.begin + t];
        smkm[off + t] = skm[rg.begin + t];
      }
      if (threadIdx.x == 0) {
        JRange s = rg;
        s.begin = off;
        s.end = off + len;
        stagedRanges[r] = s;
      }
      if (rg.isSelf) selfOffset = off - rg.begin;
      off += len;
    }
    ja.sx = smem_u32(smx); ja.sy = smem_u32(smy); ja.sz = smem_u32(smz);
    ja.sq = smem_u32(smq); ja.skm = smem_u32(smkm);
    useRanges = stagedRanges;
  }
  const unsigned qBase = smem_u32(queues + warp);
  const bool anyGeneric = g.generic[0] | g.generic[1] | g.generic[2];
  constexpr int SW = FORCE ? SWEEP_FULL : SWEEP_HALF;

  // i-atoms are handed to the warps dynamically (no tail inside the CTA); per-atom
  // energies go to shared memory and are summed in atom order, so the result does
  // not depend on which warp took which atom.
  double blockLJ = 0.0, blockReal = 0.0;
  double blockVir[4] = {0.0, 0.0, 0.0, 0.0};
  for (int pass0 = iBegin; pass0 < iEnd; pass0 += kMaxIPerPass) {
    const int passEnd = min(iEnd, pass0 + kMaxIPerPass);
    if (threadIdx.x == 0) nextI = pass0;
    __syncthreads();
    for (;;) {
      int i = 0;
      if (lane == 0) i = atomicAdd(&nextI, 1);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (i >= passEnd) break;
      PairAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0};
      double xi = sx[i], yi = sy[i], zi = sz[i], qi = sq[i];
      int2 kmi = skm[i];
      if (VIRIAL) {
        acc.cix = p.comx[kmi.y];
        acc.ciy = p.comy[kmi.y];
        acc.ciz = p.comz[kmi.y];
      }
      if (staged) {
        if (anyGeneric)
          warp_probe<VDW, MODE, SW, true, true>(p, g.generic, xi, yi, zi, kmi.x, qi, kmi.y,
                                                 i + selfOffset, i, 1.0, false, useRanges,
                                                 nRanges, 1, 0, ja, qBase, acc);
        else
          warp_probe<VDW, MODE, SW, true, false>(p, g.generic, xi, yi, zi, kmi.x, qi, kmi.y,
                                                  i + selfOffset, i, 1.0, false, useRanges,
                                                  nRanges, 1, 0, ja, qBase, acc);
      } else {
        warp_probe<VDW, MODE, SW, false, true>(p, g.generic, xi, yi, zi, kmi.x, qi, kmi.y,
                                                i + selfOffset, i, 1.0, false, useRanges,
                                                nRanges, 1, 0, ja, qBase, acc);
      }
      double e0 = warp_sum(acc.lj), e1 = warp_sum(acc.real);
      if (FORCE) {
        double f0 = warp_sum(acc.fx), f1 = warp_sum(acc.fy), f2 = warp_sum(acc.fz);
        if (lane == 0) {
          int a = sortedAtoms[i];
          fx[a] = f0;
          fy[a] = f1;
          fz[a] = f2;
        }
      }
      if (lane == 0) {
        enLJ[i - pass0] = e0;
        enReal[i - pass0] = e1;
      }
      if (VIRIAL) {
        double v2 = warp_sum(acc.fx), v3 = warp_sum(acc.fy), v4 = warp_sum(acc.fz),
               v5 = warp_sum(acc.ex);
        if (lane == 0) {
          enVir[0][i - pass0] = v2;
          enVir[1][i - pass0] = v3;
          enVir[2][i - pass0] = v4;
          enVir[3][i - pass0] = v5;
        }
      }
    }
    __syncthreads();
    // fixed-order sum of the per-atom energies of this pass (warp 0, tree of 32)
    if (warp == 0) {
      double a = 0.0, b = 0.0;
      for (int t = lane; t < passEnd - pass0; t += 32) {
        a += enLJ[t];
        b += enReal[t];
      }
      a = warp_sum(a);
      b = warp_sum(b);
      blockLJ += a;
      blockReal += b;
      if (VIRIAL) {
        for (int v = 0; v < 4; ++v) {
          double c = 0.0;
          for (int t = lane; t < passEnd - pass0; t += 32) c += enVir[v][t];
          blockVir[v] += warp_sum(c);
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partLJ[blockIdx.x] = blockLJ;
    partReal[blockIdx.x] = blockReal;
    if (VIRIAL)  // partLJ holds 6 stripes of gridDim.x partials in this mode
      for (int v = 0; v < 4; ++v) partLJ[(size_t)(v + 2) * gridDim.x + blockIdx.x] = blockVir[v];
  }
}

// Deterministic final reduction of up to 4 partial arrays of length n into
// out[0..nArr).  One block.
__global__ void k_final_reduce(int n, int nArr, const double *a0,
                               const double *a1, const double *a2,
                               const double *a3, double *out) {
  __shared__ double scratch[32];
  const double *arr[4] = {a0, a1, a2, a3};
  for (int k = 0; k < nArr; ++k) {
    double v = 0.0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) v += arr[k][t];
    double s = block_sum(v, scratch);
    if (threadIdx.x == 0) out[k] = s;
    __syncthreads();
  }
}

// molForce[m] = sum of atomForce over the atoms of m (fixed order).
__global__ void k_mol_force(int nMolsBox, const int *__restrict__ molList,
                            const int *__restrict__ molStart,
                            const double *__restrict__ fx,
                            const double *__restrict__ fy,
                            const double *__restrict__ fz, double *mx,
                            double *my, double *mz) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nMolsBox) return;
  int m = molList[t];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = molStart[m]; i < molStart[m + 1]; ++i) {
    a += fx[i];
    b += fy[i];
    c += fz[i];
  }
  mx[m] = a;
  my[m] = b;
  mz[m] = c;
}

// CalculateEnergy::CalculateTorque, src/CalculateEnergy.cpp:1365-1406.
__global__ void k_torque(BoxParams p, int nMolsBox,
                         const int *__restrict__ molList,
                         const int *__restrict__ molStart,
                         const double *__restrict__ x,
                         const double *__restrict__ y,
                         const double *__restrict__ z,
                         const double *__restrict__ cx,
                         const double *__restrict__ cy,
                         const double *__restrict__ cz,
                         const double *__restrict__ afx,
                         const double *__restrict__ afy,
                         const double *__restrict__ afz,
                         const double *__restrict__ rfx,
                         const double *__restrict__ rfy,
                         const double *__restrict__ rfz, double *tx, double *ty,
                         double *tz) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nMolsBox) return;
  int m = molList[t];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = molStart[m]; i < molStart[m + 1]; ++i) {
    double dx = x[i] - cx[m], dy = y[i] - cy[m], dz = z[i] - cz[m];
    min_image_vec(p, dx, dy, dz);
    double fx = afx[i] + rfx[i], fy = afy[i] + rfy[i], fz = afz[i] + rfz[i];
    a += dy * fz - dz * fy;
    b += dz * fx - dx * fz;
    c += dx * fy - dy * fx;
  }
  tx[m] = a;
  ty[m] = b;
  tz[m] = c;
}

// ---------------------------------------------------------------------------
// Probe kernel: energies of a few probe positions against the box, excluding
// one molecule.  Used by MoleculeInter (2*len probes: old with sign -1, new
// with sign +1 and overlap check) and ParticleInter (one probe per trial).
// grid = nProbes, block = 256.  out[probe] = {lj, real, overlap}.
struct Probe {
  double x, y, z, q, sign;
  int kind, checkOverlap, pad;
};

template <int VDW, int MODE = MODE_ENERGY>
__global__ void __launch_bounds__(kPairThreads)
    k_probe(BoxParams p, CellGrid g, int excludeMol,
            const Probe *__restrict__ probes, const int *__restrict__ cellStart,
            const double *__restrict__ sx, const double *__restrict__ sy,
            const double *__restrict__ sz, const double *__restrict__ sq,
            const int2 *__restrict__ skm, double *__restrict__ out) {
  __shared__ JRange ranges[27];
  __shared__ WarpQueue queues[kPairWarps];
  __shared__ double red[2][kPairWarps];
  __shared__ int ovl[kPairWarps];
  const Probe pr = probes[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nRangesSh =
      build_ranges(g, p, position_to_cell(g, pr.x, pr.y, pr.z), false, cellStart, ranges);
  __syncthreads();
  PairAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0};
  JArrays ja = {sx, sy, sz, sq, skm, 0u, 0u, 0u, 0u, 0u};
  warp_probe<VDW, MODE, SWEEP_PROBE, false, true>(
      p, g.generic, pr.x, pr.y, pr.z, pr.kind, pr.q, excludeMol, -1, -1, pr.sign,
      pr.checkOverlap != 0, ranges, nRangesSh, kPairWarps, warp, ja, smem_u32(&queues[warp]),
      acc);
  double a = warp_sum(acc.lj), b = warp_sum(acc.real);
  int ov = __any_sync(0xffffffffu, acc.overlap);
  if (lane == 0) {
    red[0][warp] = a;
    red[1][warp] = b;
    ovl[warp] = ov;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    int o = 0;
    for (int w = 0; w < kPairWarps; ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
      o |= ovl[w];
    }
    out[3 * blockIdx.x + 0] = s0;
    out[3 * blockIdx.x + 1] = s1;
    out[3 * blockIdx.x + 2] = (double)o;
  }
}

}  // namespace gb
