// pair2.cuh -- full-box pair sweep, second generation (orthogonal boxes).
//
// Same contract as k_pair_box (pair.cuh): CalculateEnergy::BoxInter / BoxForce / VirialCalc
// (src/CalculateEnergy.cpp:157-406, :411-579) over the cell-sorted copy, no atomics on the
// results, bit-reproducible.  What changed is where the instructions go.  The first kernel
// spent half of its issue slots on FP64 candidate distance tests (89 % of which fail) and a
// third on library erfc/sqrt/divide; this one
//   * stages the neighbour cells with 1-D TMA bulk copies (cp.async.bulk on an mbarrier) and
//     derives an FP32 copy (float4, origin = centre of the home cell, periodic shift of the
//     cell pair already applied) that only the CANDIDATE FILTER reads: r^2 in FP32 against a
//     cut-off widened by 1e-4 (far above the FP32 error of cell-relative coordinates), one
//     LDS.128 + 7 FP32 instructions per 32 candidates, no bounds tests (padding and the odd
//     border atoms of the 16-byte aligned copies are NaN);
//   * collects hits as bit masks (8 filter rounds per lane), compacts them with one warp scan
//     into a ring of 16-bit staged indices, and evaluates 32 queued pairs at a time.  The
//     evaluation recomputes (xi - xj), BoxDimensions::MinImageSigned and r^2 in FP64 exactly
//     as the reference does and applies the strict InRcut test itself -- the filter never
//     decides anything;
//   * takes erfc(alpha r)/r and the Coulomb virial factor from a piecewise degree-7
//     polynomial in r^2 (32 intervals per octave, indexed by exponent + 5 mantissa bits,
//     relative error <= 4e-14 near the cut-off and 2e-16 elsewhere; built on the host in long
//     double, staged in shared memory): 7 DFMA instead of sqrt + erfc + divide (~150
//     instructions).  Arguments outside the table (r < 0.5 A) take the library path;
//   * splits a cell's work into (i-atom, candidate segment) items so that 32 warps end
//     together, and handles full shells (forces) in as many staging passes as shared memory
//     asks for.
// Triclinic boxes keep the first kernel.
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace gb {

constexpr int kP2Block = 16;    // filter rounds (32 candidates each) per compaction
constexpr int kP2Span = 32 * kP2Block;  // candidates per block; segments are multiples of it
constexpr int kP2Queue = 512;   // per-warp ring entries, >= 31 + kP2Span / 2 (see emit)
constexpr int kP2IChunk = 256;  // i-atoms per accumulator chunk
constexpr int kP2MaxSeg = 4;    // candidate segments per i-atom (work items = atoms x segments)
constexpr int kCtBits = 5;      // Coulomb table: 2^5 intervals per octave of r^2
constexpr int kCtCoef = 8;      // degree 7
// table storage per function: c0..c3 as double[n] each, then (c4,c5) and (c6,c7) as float2[n]
// each (the high-order terms are <= 1e-5 of the value: FP32 coefficients cost < 5e-15)
constexpr int kCtWords = 6;     // 8-byte words per interval

struct Pair2Args {
  BoxParams p;
  CellGrid g;
  int slices, cell0, cap;  // cap: staged atoms per pass (even)
  int gateCap;             // largest cell population (+2) this launch can stage
  int nSeg;
  float cutF;              // filter cut-off (r^2, widened)
  int tabN, tabHi0;        // Coulomb table: intervals, (1023 + eMin) << kCtBits; tabN = 0: none
  const double *tabF, *tabG;
  const int *cellStart;
  const double *sx, *sy, *sz, *sq;
  const int2 *skm;
  const int *sortedAtoms;
  const int *maxCellPop;   // device: largest cell population (gate, see run_pair)
  double *partLJ, *partReal, *fx, *fy, *fz;
};

// mode: MODE_ENERGY / MODE_FORCE / MODE_VIRIAL of pair.cuh
__host__ __device__ inline size_t pair2_smem_bytes(int cap, int nWarps, int mode, int tabN) {
  const int nAcc = mode == 2 ? 6 : (mode == 1 ? 5 : 0);
  size_t b = 16;                                        // mbarrier
  b += (size_t)(cap + kP2Span) * 16;                    // float4 filter copy + NaN tail
  b += (size_t)cap * 40;                                // x y z q (double) + kind/mol
  b += (size_t)nWarps * kP2Queue * (mode == 0 ? 4 : 2); // hit rings
  b += (size_t)tabN * kCtWords * 8 * (mode == 0 ? 1 : 2);  // Coulomb table(s): f, g
  b += (size_t)nAcc * kP2MaxSeg * kP2IChunk * 8;        // per-(atom, segment) partials
  return b;
}

__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}

// Piecewise polynomial in s = r^2.  base: shared address of the table (layout above), n
// intervals.  False: s outside the table.
__device__ __forceinline__ bool coul_tab(unsigned base, int n, int hi0, double s, double &val) {
  const int hi = __double2hiint(s), lo = __double2loint(s);
  const int idx = (hi >> (20 - kCtBits)) - hi0;
  if ((unsigned)idx >= (unsigned)n) return false;
  // the mantissa bits below the interval index, as u in [-1, 1)
  const unsigned mh =
      ((((unsigned)hi << kCtBits) | ((unsigned)lo >> (32 - kCtBits))) & 0x000fffffu) | 0x3ff00000u;
  const double v = __hiloint2double((int)mh, (int)((unsigned)lo << kCtBits));
  const double u = fma(2.0, v, -3.0);
  const unsigned a = base + (unsigned)idx * 8u, st = (unsigned)n * 8u;
  float2 c45, c67;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(c67.x), "=f"(c67.y) : "r"(a + 5u * st));
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(c45.x), "=f"(c45.y) : "r"(a + 4u * st));
  double r = fma((double)c67.y, u, (double)c67.x);
  r = fma(r, u, (double)c45.y);
  r = fma(r, u, (double)c45.x);
#pragma unroll
  for (int c = 3; c >= 0; --c) r = fma(r, u, lds_f64(a + (unsigned)c * st));
  val = r;
  return true;
}

// One block of filter rounds: lane tests candidates a0 + 512 k (k < NK rounds), bit k of the
// result = inside the widened cut-off.  GEN: axes with fewer than 4 cells take the FP32
// minimum image (gL = axis length there, 0 elsewhere).
template <bool GEN, int NK>
__device__ __forceinline__ unsigned filter_rounds(unsigned a0, float xf, float yf, float zf,
                                                  float cutF, float gLx, float gLy, float gLz,
                                                  float iLx, float iLy, float iLz) {
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const float4 c = lds_f4(a0 + (unsigned)k * 512u);
    float dx = xf - c.x, dy = yf - c.y, dz = zf - c.z;
    if (GEN) {
      dx = fmaf(-gLx, rintf(dx * iLx), dx);
      dy = fmaf(-gLy, rintf(dy * iLy), dy);
      dz = fmaf(-gLz, rintf(dz * iLz), dz);
    }
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    mask |= (r2 < cutF) ? (1u << k) : 0u;
  }
  return mask;
}

// the library path of the Coulomb terms for arguments outside the table (kept out of line:
// it is almost never taken and must not cost the hot path registers)
template <int VDW>
__device__ __noinline__ double coul_slow_en(const BoxParams *p, double r2, double qq) {
  return calc_coulomb<VDW>(*p, r2, qq);
}
template <int VDW>
__device__ __noinline__ void coul_slow_en_vir(const BoxParams *p, double r2, double qq,
                                              double *en, double *vir) {
  calc_coulomb_en_vir<VDW>(*p, r2, qq, *en, *vir);
}


// BoxDimensions::MinImageSigned (src/BoxDimensions.h:169-175) as raw - ax * rint(raw / ax).
// |raw| <= ax, so rint gives -1, 0 or +1 and the fma is the same single rounding as the
// reference's raw -/+ ax.  The two forms can only pick different images when |raw| is within
// rounding of ax / 2 -- and GOMC requires rcut < ax / 2, so such a pair fails InRcut on
// either image: every in-range pair gets the reference's bits.
__device__ __forceinline__ double min_image_rint(double raw, double ax, double invAx) {
  return fma(-ax, rint(raw * invAx), raw);
}

// MODEL as in pair.cuh.  FAST: Ewald real-space terms from the table.
// MODE_ENERGY: items are dealt to the warps statically, the hit ring and the per-lane
// energy accumulators live across items (entries carry the i-atom), nothing is reduced per
// atom.  MODE_FORCE / MODE_VIRIAL: items are handed out dynamically and reduced per item
// into (atom, segment) slots that are summed in fixed order.
template <int VDW, int MODEL, int NWARPS, bool FAST>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_pair_box2(const __grid_constant__ Pair2Args A) {
  constexpr int MODE = MODEL & 3;
  constexpr bool FORCE = MODE == MODE_FORCE;
  constexpr bool VIRIAL = MODE == MODE_VIRIAL;
  constexpr bool PERSIST = MODE == MODE_ENERGY;
  constexpr int NACC = VIRIAL ? 6 : (FORCE ? 5 : 0);
  constexpr int NT = NWARPS * 32;
  constexpr unsigned QB = PERSIST ? 4u : 2u;  // bytes per ring entry
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ JRange ranges[27];
  __shared__ int rOff[28];       // staged offset of each range's aligned copy
  __shared__ int rLead[27];      // border atoms in front of the first real one (0/1)
  __shared__ int passFirst[28];  // ranges of staging pass q: [passFirst[q], passFirst[q+1])
  __shared__ int nPasses;
  __shared__ int nextItem;
  __shared__ double enLJ[PERSIST ? NWARPS : kP2IChunk], enReal[PERSIST ? NWARPS : kP2IChunk];
  __shared__ double enVir[VIRIAL ? 4 : 1][VIRIAL ? kP2IChunk : 1];

  const BoxParams &p = A.p;
  const int cap = A.cap;
  // the gate of run_pair: a cell that does not fit one staging pass -> the first kernel runs
  if (*A.maxCellPop + 2 > A.gateCap) return;

  const unsigned sBar = smem_u32(dyn);
  float4 *f4 = reinterpret_cast<float4 *>(dyn + 16);
  double *stx = reinterpret_cast<double *>(dyn + 16 + (size_t)(cap + kP2Span) * 16);
  double *sty = stx + cap, *stz = sty + cap, *stq = stz + cap;
  int2 *stkm = reinterpret_cast<int2 *>(stq + cap);
  unsigned char *queues = reinterpret_cast<unsigned char *>(stkm + cap);
  double *tabs = reinterpret_cast<double *>(queues + (size_t)NWARPS * kP2Queue * QB);
  constexpr bool TWO = MODE != MODE_ENERGY;  // f and g tables
  double *accS = tabs + (size_t)A.tabN * kCtWords * (TWO ? 2 : 1);
  const unsigned aF4 = smem_u32(f4), aX = smem_u32(stx), aY = smem_u32(sty), aZ = smem_u32(stz),
                 aQ = smem_u32(stq), aKM = smem_u32(stkm);
  const unsigned aTabF = smem_u32(tabs);
  const unsigned aTabG = aTabF + (TWO ? (unsigned)A.tabN * kCtWords * 8u : 0u);

  const int cell = A.cell0 + blockIdx.x / A.slices, slice = blockIdx.x % A.slices;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int iBegin0 = A.cellStart[cell], iEnd0 = A.cellStart[cell + 1];
  const int nI0 = iEnd0 - iBegin0;
  const int iBegin = iBegin0 + (int)(((long long)nI0 * slice) / A.slices);
  const int iEnd = iBegin0 + (int)(((long long)nI0 * (slice + 1)) / A.slices);
  const unsigned aQueue = smem_u32(queues + (size_t)warp * kP2Queue * QB);

  const int nRanges = build_ranges(A.g, p, cell, !FORCE, A.cellStart, ranges);
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (!FORCE) {
    // half shell: lay the forward cells out so that both halves (and all four quarters) of
    // the flat candidate list hold a similar number of in-range pairs -- face, edge and
    // corner neighbours alternate -- because work items are contiguous segments of the list
    // and are dealt to the warps statically.  Index = d - 13 of build_ranges.
    // order {0, 2, 5, 1, 4, 6, 7, 3, 8, 11, 10, 9, 12, 13}, one nibble per position
    const unsigned long long perm = 0xDC9AB837641520ULL;
    JRange r;
    if (threadIdx.x < 14) r = ranges[(int)((perm >> (4 * threadIdx.x)) & 15ULL)];
    __syncthreads();
    if (threadIdx.x < 14) ranges[threadIdx.x] = r;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // greedy grouping of the ranges into staging passes of <= cap atoms (aligned copies)
    // MODE_ENERGY keeps the self range (ranges[0], the i-atoms) at the front of every pass:
    // ring entries name their i-atom by staged index
    int q = 0, used = 0;
    passFirst[0] = 0;
    const int selfLenE = ((ranges[0].end + 1) & ~1) - (ranges[0].begin & ~1);
    for (int r = 0; r < nRanges; ++r) {
      const int b = ranges[r].begin & ~1, e = (ranges[r].end + 1) & ~1;
      const int len = ranges[r].end > ranges[r].begin ? e - b : 0;
      if (used + len > cap) {
        passFirst[++q] = r;
        used = PERSIST ? selfLenE : 0;
      }
      rOff[r] = used;
      rLead[r] = ranges[r].begin - b;
      used += len;
    }
    passFirst[++q] = nRanges;
    nPasses = q;
  }
  __syncthreads();
  if (iEnd <= iBegin) {  // no i-atoms in this slice (uniform)
    if (threadIdx.x == 0) {
      A.partLJ[blockIdx.x] = 0.0;
      A.partReal[blockIdx.x] = 0.0;
      if (VIRIAL)
        for (int v = 0; v < 4; ++v) A.partLJ[(size_t)(v + 2) * gridDim.x + blockIdx.x] = 0.0;
    }
    return;
  }

  // home-cell centre = origin of the FP32 copy
  const int hcz = cell % A.g.edge[2], hcy = (cell / A.g.edge[2]) % A.g.edge[1],
            hcx = cell / (A.g.edge[2] * A.g.edge[1]);
  const double ox = (hcx + 0.5) * A.g.cellSize[0], oy = (hcy + 0.5) * A.g.cellSize[1],
               oz = (hcz + 0.5) * A.g.cellSize[2];
  const bool gx = A.g.generic[0] != 0, gy = A.g.generic[1] != 0, gz = A.g.generic[2] != 0;
  const bool anyGen = gx | gy | gz;
  const float gLx = gx ? (float)p.ax[0] : 0.0f, gLy = gy ? (float)p.ax[1] : 0.0f,
              gLz = gz ? (float)p.ax[2] : 0.0f;
  const float iLx = 1.0f / (float)p.ax[0], iLy = 1.0f / (float)p.ax[1],
              iLz = 1.0f / (float)p.ax[2];
  const double invAx = 1.0 / p.ax[0], invAy = 1.0 / p.ax[1], invAz = 1.0 / p.ax[2];
  const float cutF = A.cutF;
  const int nSeg = A.nSeg;
  const float qnan = __int_as_float(0x7fc00000);

  double blockLJ = 0.0, blockReal = 0.0;
  double blockVir[4] = {0.0, 0.0, 0.0, 0.0};
  unsigned phase = 0;
  bool tabLoaded = false;
  PairAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0};  // PERSIST: whole kernel

  // One queued pair: FP64 distance as the reference computes it, strict InRcut, functors.
  auto eval = [&](unsigned jj, int selfIdx, double xi, double yi, double zi, double qi, int2 kmi,
                  PairAcc &ac) {
    const double dx = min_image_rint(xi - lds_f64(aX + jj * 8u), p.ax[0], invAx);
    const double dy = min_image_rint(yi - lds_f64(aY + jj * 8u), p.ax[1], invAy);
    const double dz = min_image_rint(zi - lds_f64(aZ + jj * 8u), p.ax[2], invAz);
    const double r2 = dist_sq(dx, dy, dz);
    if (!(p.boxRcutSq > r2) || (int)jj == selfIdx) return;  // InRcut, strict
    const double qj = lds_f64(aQ + jj * 8u);
    int2 kmj;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];"
                 : "=r"(kmj.x), "=r"(kmj.y)
                 : "r"(aKM + jj * 8u));
    if (!FAST) {
      eval_pair<VDW, MODEL>(p, kmi.x, qi, kmi.y, true, 1.0, false, dx, dy, dz, qj, kmj, ac);
      return;
    }
    if (kmj.y == kmi.y) return;
    const int idx = kmi.x + kmj.x * p.kindCount;
    if (MODE == MODE_ENERGY) {
      const double qq = qi * qj * kQQFact;
      if (qq != 0.0 && !(p.rCutCoulombSq < r2)) {
        double f;
        if (coul_tab(aTabF, A.tabN, A.tabHi0, r2, f))
          ac.real += qq * f;
        else
          ac.real += coul_slow_en<VDW>(&A.p, r2, qq);
      }
      ac.lj += calc_en<VDW>(p, r2, idx);
    } else {
      double eL, wL, eC = 0.0, wC = 0.0;
      calc_en_vir<VDW>(p, r2, idx, eL, wL);
      const double qq = qi * qj * (VIRIAL ? 1.0 : kQQFact);
      if (qq != 0.0 && !(p.rCutCoulombSq < r2)) {
        double f, gg;
        if (coul_tab(aTabF, A.tabN, A.tabHi0, r2, f) &&
            coul_tab(aTabG, A.tabN, A.tabHi0, r2, gg)) {
          eC = qq * f;
          wC = qq * gg;
        } else {
          coul_slow_en_vir<VDW>(&A.p, r2, qq, &eC, &wC);
        }
      }
      if (VIRIAL) {
        double cx = ac.cix - p.comx[kmj.y], cy = ac.ciy - p.comy[kmj.y],
               cz = ac.ciz - p.comz[kmj.y];
        min_image_vec(p, cx, cy, cz);
        const double t1 = dx * cx, t2 = dy * cy, t3 = dz * cz;
        ac.lj += wL * t1;
        ac.real += wL * t2;
        ac.fx += wL * t3;
        ac.fy += wC * t1;
        ac.fz += wC * t2;
        ac.ex += wC * t3;
      } else {
        ac.lj += eL;
        ac.real += eC;
        const double w = wL + wC;
        ac.fx += dx * w;
        ac.fy += dy * w;
        ac.fz += dz * w;
      }
    }
  };

  // MODE_ENERGY ring entry: (staged index of i) << 16 | staged index of j; half shell, so
  // j is never i itself
  auto evalStaged = [&](unsigned ent) {
    const unsigned is = ent >> 16;
    int2 kmi;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(kmi.x), "=r"(kmi.y) : "r"(aKM + is * 8u));
    eval(ent & 0xffffu, -1, lds_f64(aX + is * 8u), lds_f64(aY + is * 8u), lds_f64(aZ + is * 8u),
         lds_f64(aQ + is * 8u), kmi, acc);
  };

  for (int chunk0 = iBegin; chunk0 < iEnd; chunk0 += kP2IChunk) {
    const int nIc = min(kP2IChunk, iEnd - chunk0);
    if (!PERSIST)
      for (int t = threadIdx.x; t < NACC * kP2MaxSeg * kP2IChunk; t += NT) accS[t] = 0.0;
    for (int q = 0; q < nPasses; ++q) {
      const int r0 = passFirst[q], r1 = passFirst[q + 1];
      __syncthreads();  // previous pass / chunk done with the staging buffers
      // ---- stage: TMA bulk copies of the five sorted arrays, range by range ----------
      if (warp == 0) {
        const bool again = PERSIST && q > 0;  // re-stage the self range in front
        const unsigned tb = (unsigned)A.tabN * kCtWords * 8u;
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          unsigned bytes = 0;
          for (int r = again ? -1 : r0; r < r1; r = r < r0 ? r0 : r + 1) {
            const int rr = r < 0 ? 0 : r;
            const int b = ranges[rr].begin & ~1, e = (ranges[rr].end + 1) & ~1;
            if (ranges[rr].end > ranges[rr].begin) bytes += (unsigned)(e - b) * 40u;
          }
          if (FAST && !tabLoaded) bytes += tb * (TWO ? 2u : 1u);
          mbar_expect_tx(sBar, bytes);
        }
        __syncwarp();
        // one lane per range issues its five copies (lane 31: the re-staged self range)
        const int r = lane == 31 ? (again ? 0 : -1) : r0 + lane;
        if (r >= 0 && (lane == 31 || r < r1) && ranges[r].end > ranges[r].begin) {
          const int b = ranges[r].begin & ~1, e = (ranges[r].end + 1) & ~1;
          const unsigned nb = (unsigned)(e - b) * 8u, o = lane == 31 ? 0u : (unsigned)rOff[r] * 8u;
          bulk_g2s(aX + o, A.sx + b, nb, sBar);
          bulk_g2s(aY + o, A.sy + b, nb, sBar);
          bulk_g2s(aZ + o, A.sz + b, nb, sBar);
          bulk_g2s(aQ + o, A.sq + b, nb, sBar);
          bulk_g2s(aKM + o, A.skm + b, nb, sBar);
        }
        if (FAST && !tabLoaded && lane == 30) {
          bulk_g2s(aTabF, A.tabF, tb, sBar);
          if (TWO) bulk_g2s(aTabG, A.tabG, tb, sBar);
        }
      }
      tabLoaded = true;
      while (!mbar_try_wait(sBar, phase)) __nanosleep(64);
      phase ^= 1u;
      // ---- FP32 filter copy (shift of the cell pair applied; border atoms -> NaN) ------
      int flatEnd = 0, selfBase = 0;  // selfBase: staged index of sorted atom 0 of the self range
      bool haveSelf = false;
      if (PERSIST && q > 0) {  // re-staged i-atoms: addressable, but not candidates
        const int lenE = ((ranges[0].end + 1) & ~1) - (ranges[0].begin & ~1);
        for (int t = threadIdx.x; t < lenE; t += NT) f4[t] = make_float4(qnan, qnan, qnan, 0.0f);
        selfBase = rLead[0] - ranges[0].begin;
        flatEnd = lenE;
      }
      for (int r = r0; r < r1; ++r) {
        const JRange rg = ranges[r];
        const int len = rg.end - rg.begin;
        if (len <= 0) continue;
        const int lead = rLead[r], lenE = ((rg.end + 1) & ~1) - (rg.begin & ~1);
        const int off = rOff[r];
        for (int t = threadIdx.x; t < lenE; t += NT) {
          float4 v = make_float4(qnan, qnan, qnan, 0.0f);
          const int u = t - lead;
          if (u >= 0 && u < len) {
            v.x = (float)((stx[off + t] - rg.sx) - ox);
            v.y = (float)((sty[off + t] - rg.sy) - oy);
            v.z = (float)((stz[off + t] - rg.sz) - oz);
          }
          f4[off + t] = v;
        }
        if (rg.isSelf) {
          selfBase = off + lead - rg.begin;
          haveSelf = true;
        }
        flatEnd = off + lenE;
      }
      for (int t = threadIdx.x; t < kP2Span; t += NT)
        f4[flatEnd + t] = make_float4(qnan, qnan, qnan, 0.0f);
      if (threadIdx.x == 0) nextItem = 0;
      __syncthreads();

      // ---- work items: (i-atom, candidate segment); segments are whole filter blocks ----
      const int segLen = ((flatEnd + nSeg - 1) / nSeg + kP2Span - 1) / kP2Span * kP2Span;
      const int nItems = nIc * nSeg;
      int head = 0, count = 0;
      int item = PERSIST ? warp - NWARPS : 0;
      for (;;) {
        if (PERSIST) {
          item += NWARPS;
        } else {
          if (lane == 0) item = atomicAdd(&nextItem, 1);
          item = __shfl_sync(0xffffffffu, item, 0);
        }
        if (item >= nItems) break;
        const int seg = item / nIc, il = item - seg * nIc;
        const int i = chunk0 + il;
        const int selfIdx = haveSelf ? selfBase + i : -1;
        int start = seg * segLen;
        const int end = min(flatEnd, start + segLen);
        // half shell: j > i.  Blocks stay aligned (a block never reaches into the next
        // segment); the candidates in front of i + 1 are masked out below.
        const int firstJ = (!FORCE && haveSelf) ? selfIdx + 1 : 0;
        if (firstJ > start) start = firstJ / kP2Span * kP2Span;
        if (start >= end) continue;
        const double xi = A.sx[i], yi = A.sy[i], zi = A.sz[i], qi = A.sq[i];
        const int2 kmi = A.skm[i];
        const float xf = (float)(xi - ox), yf = (float)(yi - oy), zf = (float)(zi - oz);
        if (!PERSIST) {
          acc = PairAcc{0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0};
          head = count = 0;
          if (VIRIAL) {
            acc.cix = p.comx[kmi.y];
            acc.ciy = p.comy[kmi.y];
            acc.ciz = p.comz[kmi.y];
          }
        }
        const unsigned tag = PERSIST ? (unsigned)(selfBase + i) << 16 : 0u;

        auto drain = [&](bool active) {
          if (!active) return;
          unsigned ent;
          const unsigned qa = aQueue + (unsigned)((head + lane) & (kP2Queue - 1)) * QB;
          if (PERSIST) {
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ent) : "r"(qa));
            evalStaged(ent);
          } else {
            ent = lds_u16(qa);
            eval(ent, selfIdx, xi, yi, zi, qi, kmi, acc);
          }
        };
        // append the hits of `mask` (bit k: candidate t0 + 32 k + lane) to the ring
        auto emit = [&](unsigned mask, int t0) {
          // ordinal-major order: the h-th hits of all lanes sit next to each other, so the 32
          // entries of a drain come from different lanes -> different shared-memory banks
          // (a lane's own hits are 32 candidates = one bank period apart)
          const unsigned lt = (1u << lane) - 1u;
          int pos = head + count, total = 0;
          for (;;) {
            const unsigned bal = __ballot_sync(0xffffffffu, mask != 0);
            if (bal == 0) break;
            if (mask) {
              const int k = __ffs(mask) - 1;
              mask &= mask - 1;
              const unsigned qa =
                  aQueue + (unsigned)((pos + __popc(bal & lt)) & (kP2Queue - 1)) * QB;
              const unsigned v = tag | (unsigned)(t0 + 32 * k + lane);
              if (PERSIST)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(qa), "r"(v) : "memory");
              else
                sts_u16(qa, v);
            }
            const int nb = __popc(bal);
            pos += nb;
            total += nb;
          }
          count += total;
          __syncwarp();
          while (count >= 32) {
            drain(true);
            head = (head + 32) & (kP2Queue - 1);
            count -= 32;
          }
          __syncwarp();
        };

        for (int t0 = start; t0 < end; t0 += kP2Span) {
          const unsigned a0 = aF4 + (unsigned)(t0 + lane) * 16u;
          unsigned mask =
              anyGen ? filter_rounds<true, kP2Block>(a0, xf, yf, zf, cutF, gLx, gLy, gLz, iLx, iLy,
                                                     iLz)
                     : filter_rounds<false, kP2Block>(a0, xf, yf, zf, cutF, gLx, gLy, gLz, iLx,
                                                      iLy, iLz);
          if (t0 < firstJ) {  // rounds of this lane in front of firstJ
            const int d = firstJ - t0 - lane;
            const int nb = d <= 0 ? 0 : min(kP2Block, (d + 31) >> 5);
            mask &= ~((1u << nb) - 1u);
          }
          // the ring holds 31 left-overs + half a block of hits for certain; a block with
          // more hits than that (dense neighbourhoods) goes in two halves
          const int tot = __reduce_add_sync(0xffffffffu, __popc(mask));
          if (count + tot <= kP2Queue) {
            emit(mask, t0);
          } else {
            emit(mask & 0x00ffu, t0);
            emit(mask & 0xff00u, t0);
          }
        }
        if (!PERSIST) {
          drain(lane < count);
          __syncwarp();
          // per-(atom, segment) partials; one warp owns an item, passes run in order
          double s0 = warp_sum(acc.lj), s1 = warp_sum(acc.real);
          double s2 = warp_sum(acc.fx), s3 = warp_sum(acc.fy), s4 = warp_sum(acc.fz), s5 = 0.0;
          if (VIRIAL) s5 = warp_sum(acc.ex);
          if (lane == 0) {
            double *a = accS + (size_t)seg * kP2IChunk + il;
            constexpr int ST = kP2MaxSeg * kP2IChunk;
            a[0] += s0;
            a[ST] += s1;
            a[2 * ST] += s2;
            a[3 * ST] += s3;
            a[4 * ST] += s4;
            if (VIRIAL) a[5 * ST] += s5;
          }
        }
      }
      if (PERSIST) {
        // flush the ring before the staging buffers change
        if (lane < count) {
          unsigned ent;
          asm volatile("ld.shared.u32 %0, [%1];"
                       : "=r"(ent)
                       : "r"(aQueue + (unsigned)((head + lane) & (kP2Queue - 1)) * 4u));
          evalStaged(ent);
        }
        __syncwarp();
      }
    }
    if (PERSIST) continue;
    __syncthreads();
    // ---- chunk epilogue: segments in fixed order; forces out; energies per atom ---------
    constexpr int ST = kP2MaxSeg * kP2IChunk;
    for (int il = threadIdx.x; il < nIc; il += NT) {
      double v[NACC > 0 ? NACC : 1];
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        double s = 0.0;
        for (int g = 0; g < nSeg; ++g) s += accS[(size_t)a * ST + (size_t)g * kP2IChunk + il];
        v[a] = s;
      }
      if (FORCE) {
        // every pair was seen from both sides: half of the energy belongs to this atom
        enLJ[il] = 0.5 * v[0];
        enReal[il] = 0.5 * v[1];
        const int at = A.sortedAtoms[chunk0 + il];
        A.fx[at] = v[2];
        A.fy[at] = v[3];
        A.fz[at] = v[4];
      } else if (VIRIAL) {
        enLJ[il] = v[0];
        enReal[il] = v[1];
        enVir[0][il] = v[2];
        enVir[1][il] = v[3];
        enVir[2][il] = v[4];
        enVir[3][il] = v[5];
      }
    }
    __syncthreads();
    if (warp == 0) {
      double a = 0.0, b = 0.0;
      for (int t = lane; t < nIc; t += 32) {
        a += enLJ[t];
        b += enReal[t];
      }
      blockLJ += warp_sum(a);
      blockReal += warp_sum(b);
      if (VIRIAL) {
        for (int v = 0; v < 4; ++v) {
          double c = 0.0;
          for (int t = lane; t < nIc; t += 32) c += enVir[v][t];
          blockVir[v] += warp_sum(c);
        }
      }
    }
  }
  if (PERSIST) {
    // warp totals, then the warps in order
    const double a = warp_sum(acc.lj), b = warp_sum(acc.real);
    if (lane == 0) {
      enLJ[warp] = a;
      enReal[warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int w = 0; w < NWARPS; ++w) {
        blockLJ += enLJ[w];
        blockReal += enReal[w];
      }
  }
  if (threadIdx.x == 0) {
    A.partLJ[blockIdx.x] = blockLJ;
    A.partReal[blockIdx.x] = blockReal;
    if (VIRIAL)
      for (int v = 0; v < 4; ++v) A.partLJ[(size_t)(v + 2) * gridDim.x + blockIdx.x] = blockVir[v];
  }
}

}  // namespace gb
