// recip_mma.cuh -- structure-factor sums on the FP64 MMA path (DMMA, sm_100a).
//
// Same factorisation as k_recip_fact (recip.cuh): with A_i(a,b) = q_i X_i^a Y_i^b
// and Z_i^c = (cz, sz) the four real products
//     P1 = sum_i Ar cz   P2 = sum_i Ai sz   P3 = sum_i Ar sz   P4 = sum_i Ai cz
// give S(a,b,+c) = (P1-P2, P3+P4) and S(a,b,-c) = (P1+P2, P4-P3).  Written as a
// real GEMM  C[m][n] = sum_atoms A[atom][m] * B[atom][n]  with m = 2*row + {Ar,Ai}
// and n = 2*c + {cz,sz}, every entry of C is one of the P's: no wasted flops,
// 2 FMA per (atom, k) as the roofline assumes.  The GEMM runs on
// mma.sync.m8n8k4.f64 (SASS DMMA): 256 FMA per warp instruction instead of 32,
// 4x less shared-memory operand traffic than the SIMT register tile, and a
// slightly higher measured ceiling (37.2 vs 34.1 TFLOP/s, profiles/r1_fp64_peak.json).
//
// Kernel organisation (one persistent CTA per SM, 16 warps):
//   * k_phase_tables (pre-kernel) writes per-atom phase tables once per
//     evaluation: tabXY[atom] = {q e^{i a tx}, a<=KX ; e^{i b ty}, b<=KY} and
//     tabZ[atom] = {e^{i c tz}, c<ZS}  (1.2e7 sincos for the 100k-atom box);
//   * the host cuts the whole job -- a line of (row tile, atom chunk) units
//     weighted by the tile's column count -- into one equal piece per SM, so
//     there is a single wave and no tail; a piece is 1-3 "segments"
//     (tile, chunk range, slab);
//   * per chunk of AT atoms the XY and Z tables arrive in shared memory by
//     cp.async.bulk (TMA, SASS UBLKCP) on mbarriers, two chunks ahead;
//   * software pipeline: while the warps issue the DMMAs of chunk c they also
//     build the A tile (complex products X^a Y^b) of chunk c+1 into the other
//     buffer -- one element per thread per k4 step -- so the FP64 MMA pipe never
//     waits for operand generation.  The A tile is private to the warp that
//     consumes it and table buffers are recycled by "last warp done refills",
//     so there is no CTA-wide barrier in steady state;
//   * each warp owns 8 (a,b) rows x up to 40 c columns = MT(2) x NT(<=10) m8n8
//     accumulator tiles (80 registers).
// Partials per slab go to part[slab][re/im][k]; k_recip_finish sums them in slab
// order (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace gb {

constexpr int kMmaThreads = 512;
constexpr int kMmaWarps = 16;
constexpr int kMmaMT = 2;                      // m8 tiles per warp: 8 (a,b) rows
constexpr int kMmaRowsPerWarp = kMmaMT * 4;    // 8
constexpr int kMmaRows = kMmaWarps * kMmaRowsPerWarp;  // 128 rows per tile
constexpr int kMmaMaxNT = 10;                  // n8 tiles per column block: 40 c
constexpr int kMmaAWS = 160;  // bytes per atom of a warp's private A tile: 8 rows x 16 B + pad
                              // (stride 20 doubles == 4 mod 16: conflict-free A fragments)

struct MmaArgs {
  const int4 *rows;    // {a, b, cmax, start}, sorted by cmax descending, padded
  const int4 *tiles;   // {rowBegin, c0, NT, cmaxTile}
  const int4 *segs;    // {tile, chunkBegin, chunkEnd, slab}
  const int *ctaSeg;   // [gridDim.x + 1] segment ranges per CTA
  const double2 *tabXY;  // [atom][XYS]
  const double2 *tabZ;   // [atom][ZS]
  int KX1;             // X table length (Y follows at offset KX1)
  int XYS;             // = KX1 + KY1
  int ZS;              // Z row stride (double2), ZS % 8 == 2, >= 40 * colBlocks
  int AT;              // atoms per chunk (multiple of 4)
  int nkStride;
  long long *ctaCycles;  // optional per-CTA clock64 duration (work-split calibration)
};

// ---- phase tables ----------------------------------------------------------
// One thread per (atom, table entry).  cv = 2 pi / L per axis
// (XYZ::Inverse then *= 2 pi, src/Ewald.cpp:852-854).
__global__ void __launch_bounds__(256)
    k_phase_tables(int nAtoms, int nAtomsPadded, int KX1, int KY1, int KZ1, int ZS,
                   double cvx, double cvy, double cvz, const double4 *__restrict__ pb,
                   double2 *__restrict__ tabXY, double2 *__restrict__ tabZ) {
  const int PS = KX1 + KY1 + ZS;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nAtomsPadded * PS;
  if (idx >= total) return;
  int atom = (int)(idx / PS), e = (int)(idx - (long long)atom * PS);
  double2 v = make_double2(0.0, 0.0);
  double coord = 0.0, cv = 0.0, scale = 1.0;
  int n = 0;
  bool valid = atom < nAtoms;
  double2 *dst;
  if (e < KX1 + KY1) {
    dst = tabXY + (size_t)atom * (KX1 + KY1) + e;
  } else {
    dst = tabZ + (size_t)atom * ZS + (e - KX1 - KY1);
  }
  if (valid) {
    double4 a = pb[atom];
    if (e < KX1) {
      coord = a.x; cv = cvx; n = e; scale = a.w;  // charge folded into X
    } else if (e < KX1 + KY1) {
      coord = a.y; cv = cvy; n = e - KX1;
    } else {
      coord = a.z; cv = cvz; n = e - KX1 - KY1;
      valid = n < KZ1;
    }
    if (valid) {
      double s, c;
      sincos((coord * cv) * (double)n, &s, &c);
      v = make_double2(scale * c, scale * s);
    }
  }
  *dst = v;
}

// ---- PTX helpers: mbarrier / bulk copy in common.cuh -----------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64x2(unsigned addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

// A-tile element for this thread's row: (q X^a)(Y^b), conj(Y) when b < 0.
__device__ __forceinline__ void agen_one(unsigned xyAtom, unsigned offX, unsigned offY,
                                         double ysign, unsigned dst) {
  double2 xv = lds_f64x2(xyAtom + offX);
  double2 yv = lds_f64x2(xyAtom + offY);
  yv.y *= ysign;
  sts_f64x2(dst, xv.x * yv.x - xv.y * yv.y, xv.x * yv.y + xv.y * yv.x);
}

// One chunk: nK4 steps of (MT x NT) DMMAs on tileA/Z of this chunk, interleaved
// with building the next chunk's A tile (one element per thread per step).
// All addresses are 32-bit shared-window addresses.
template <int NT>
__device__ __forceinline__ void mma_chunk(double (&acc)[kMmaMT][kMmaMaxNT][2], unsigned aAddr,
                                          unsigned zAddr, int nK4, unsigned aStep,
                                          unsigned zStep, bool genNext, unsigned xyNext,
                                          unsigned xyStep, unsigned offX, unsigned offY,
                                          double ysign, unsigned aNext, unsigned aNextStep) {
  for (int k4 = 0; k4 < nK4; ++k4) {
    double a[kMmaMT], b[NT];
#pragma unroll
    for (int mt = 0; mt < kMmaMT; ++mt) a[mt] = lds_f64(aAddr + mt * 64);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = lds_f64(zAddr + nt * 64);
    if (genNext) agen_one(xyNext, offX, offY, ysign, aNext);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < kMmaMT; ++mt)
        dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    aAddr += aStep;
    zAddr += zStep;
    xyNext += xyStep;
    aNext += aNextStep;
  }
}

__global__ void __launch_bounds__(kMmaThreads, 1)
    k_recip_mma(MmaArgs ma, double *__restrict__ part) {
  extern __shared__ __align__(16) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long mbarStore[4];  // "full": XY0, XY1, Z0, Z1
  __shared__ int doneCnt[4];  // warps finished with XY0, XY1, Z0, Z1 ("empty" side)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tStart = clock64();
  const int AT = ma.AT, XYS = ma.XYS, ZS = ma.ZS;
  const unsigned xyBytes = (unsigned)(AT * XYS * 16), zBytes = (unsigned)(AT * ZS * 16);
  const unsigned aWarpBytes = (unsigned)(AT * kMmaAWS);
  const unsigned aBytes = aWarpBytes * kMmaWarps;
  const unsigned smBase = smem_u32(dynSmem);
  // buffer / barrier addresses are computed, not indexed (keeps them in registers)
  auto xyBufAt = [&](int i) { return smBase + (unsigned)i * xyBytes; };
  auto zBufAt = [&](int i) { return smBase + 2 * xyBytes + (unsigned)i * zBytes; };
  auto aBufAt = [&](int i) { return smBase + 2 * xyBytes + 2 * zBytes + (unsigned)i * aBytes; };
  const unsigned barBase = smem_u32(&mbarStore[0]);
  auto barXYAt = [&](int i) { return barBase + 8u * (unsigned)i; };
  auto barZAt = [&](int i) { return barBase + 16u + 8u * (unsigned)i; };
  unsigned phaseBits = 0;  // bit i: parity to wait for next on barrier i (XY0, XY1, Z0, Z1)

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(barXYAt(i), 1);
      mbar_init(barZAt(i), 1);
    }
    for (int i = 0; i < 4; ++i) doneCnt[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // A-generation role of this lane: one of the warp's own 8 rows, atoms g, g+4, ...
  // (the A tile is private to the warp: no CTA-wide barrier in steady state)
  const int genRow = warp * kMmaRowsPerWarp + (lane & 7), genG = lane >> 3;
  const int nK4 = AT / 4;
  const unsigned aWarpOff = (unsigned)warp * aWarpBytes;

  // "empty" signalling: every warp reports when it is done with a table buffer;
  // the last one to report refills it through the TMA unit.
  auto arriveAndRefill = [&](int which, unsigned dstBuf, unsigned bar, const double2 *src,
                             unsigned bytes, bool more) {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      int old = atomicAdd(&doneCnt[which], 1);
      if (old == kMmaWarps - 1) {
        atomicExch(&doneCnt[which], 0);
        if (more) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(bar, bytes);
          bulk_g2s(dstBuf, src, bytes, bar);
        }
      }
    }
  };

  const int segBegin = ma.ctaSeg[blockIdx.x], segEnd = ma.ctaSeg[blockIdx.x + 1];
  for (int sIdx = segBegin; sIdx < segEnd; ++sIdx) {
    const int4 seg = ma.segs[sIdx];
    const int4 tile = ma.tiles[seg.x];
    const int rowBegin = tile.x, c0 = tile.y, NT = tile.z;
    const int chunk0 = seg.y, nChunks = seg.z - seg.y, slab = seg.w;

    const int4 myRow = ma.rows[rowBegin + genRow];
    const unsigned offX = (unsigned)myRow.x * 16u;
    const unsigned offY = (unsigned)(ma.KX1 + (myRow.y < 0 ? -myRow.y : myRow.y)) * 16u;
    const double ysign = myRow.y < 0 ? -1.0 : 1.0;

    double acc[kMmaMT][kMmaMaxNT][2];
#pragma unroll
    for (int mt = 0; mt < kMmaMT; ++mt)
#pragma unroll
      for (int nt = 0; nt < kMmaMaxNT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // ---- prologue: XY(0), XY(1), Z(0), Z(1) in flight; build A(0) ----------
    const double2 *xySrc = ma.tabXY + (size_t)chunk0 * AT * XYS;
    const double2 *zSrc = ma.tabZ + (size_t)chunk0 * AT * ZS;
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int i = 0; i < 2 && i < nChunks; ++i) {
        mbar_expect_tx(barXYAt(i), xyBytes);
        bulk_g2s(xyBufAt(i), xySrc + (size_t)i * AT * XYS, xyBytes, barXYAt(i));
        mbar_expect_tx(barZAt(i), zBytes);
        bulk_g2s(zBufAt(i), zSrc + (size_t)i * AT * ZS, zBytes, barZAt(i));
      }
    }
    mbar_wait(barXYAt(0), phaseBits & 1u);
    phaseBits ^= 1u;
    for (int j = 0; j < nK4; ++j) {
      int at = 4 * j + genG;
      agen_one(xyBufAt(0) + (unsigned)(at * XYS) * 16u, offX, offY, ysign,
               aBufAt(0) + aWarpOff + (unsigned)(at * kMmaAWS) + (unsigned)((lane & 7) * 16));
    }
    // XY[0] is free once every warp has built its A(0): refill with XY(2)
    arriveAndRefill(0, xyBufAt(0), barXYAt(0), xySrc + (size_t)2 * AT * XYS, xyBytes,
                    2 < nChunks);

    for (int c = 0; c < nChunks; ++c) {
      const int cur = c & 1, nxt = cur ^ 1;
      const bool genNext = c + 1 < nChunks;
      mbar_wait(barZAt(cur), (phaseBits >> (2 + cur)) & 1u);
      phaseBits ^= 1u << (2 + cur);
      if (genNext) {
        mbar_wait(barXYAt(nxt), (phaseBits >> nxt) & 1u);
        phaseBits ^= 1u << nxt;
      }
      // lane fragment addresses: A[k = lane&3][m = lane>>2], B[k = lane&3][n = lane>>2]
      const unsigned aAddr = aBufAt(cur) + aWarpOff + (unsigned)((lane & 3) * kMmaAWS) +
                             (unsigned)((lane >> 2) * 8);
      const unsigned zAddr = zBufAt(cur) + (unsigned)((lane & 3) * ZS * 16) + (unsigned)(c0 * 16) +
                             (unsigned)((lane >> 2) * 8);
      const unsigned aStep = 4u * kMmaAWS, zStep = 4u * (unsigned)ZS * 16u;
      const unsigned xyNext = xyBufAt(nxt) + (unsigned)(genG * XYS) * 16u;
      const unsigned xyStep = 4u * (unsigned)XYS * 16u;
      const unsigned aNext = aBufAt(nxt) + aWarpOff + (unsigned)(genG * kMmaAWS) +
                             (unsigned)((lane & 7) * 16);
      __syncwarp();  // A(c) was written by this warp's lanes in the previous iteration
#define MMA_CASE(N)                                                                       \
  case N:                                                                                 \
    mma_chunk<N>(acc, aAddr, zAddr, nK4, aStep, zStep, genNext, xyNext, xyStep, offX, offY, \
                 ysign, aNext, aStep);                                                    \
    break;
      switch (NT) {
        MMA_CASE(1) MMA_CASE(2) MMA_CASE(3) MMA_CASE(4) MMA_CASE(5)
        MMA_CASE(6) MMA_CASE(7) MMA_CASE(8) MMA_CASE(9)
        default:
          mma_chunk<10>(acc, aAddr, zAddr, nK4, aStep, zStep, genNext, xyNext, xyStep, offX,
                        offY, ysign, aNext, aStep);
      }
#undef MMA_CASE
      // this warp is done with Z[cur] (DMMAs of chunk c) and XY[nxt] (A(c+1) built)
      arriveAndRefill(2 + cur, zBufAt(cur), barZAt(cur), zSrc + (size_t)(c + 2) * AT * ZS, zBytes,
                      c + 2 < nChunks);
      if (genNext)
        arriveAndRefill(nxt, xyBufAt(nxt), barXYAt(nxt), xySrc + (size_t)(c + 3) * AT * XYS,
                        xyBytes, c + 3 < nChunks);
    }

    // ---- epilogue: exchange Ar*/Ai* partners (lane ^ 4), write S(a,b,+-c) ------
    double *pr = part + (size_t)(slab * 2 + 0) * ma.nkStride;
    double *pi = part + (size_t)(slab * 2 + 1) * ma.nkStride;
    const int comp = (lane >> 2) & 1;  // 0: lane holds Ar*{cz,sz}; 1: Ai*{cz,sz}
#pragma unroll
    for (int mt = 0; mt < kMmaMT; ++mt) {
      const int r = warp * kMmaRowsPerWarp + mt * 4 + (lane >> 3);
      const int4 rw = ma.rows[rowBegin + r];
      const int cmax = rw.z;
      const bool origin = (rw.x == 0 && rw.y == 0);
#pragma unroll
      for (int nt = 0; nt < kMmaMaxNT; ++nt) {
        if (nt < NT) {
          double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
          double o0 = __shfl_xor_sync(0xffffffffu, v0, 4);
          double o1 = __shfl_xor_sync(0xffffffffu, v1, 4);
          const int cc = c0 + nt * 4 + (lane & 3);
          if (cc <= cmax) {
            // comp 0: v0 = P1 (Ar cz), v1 = P3 (Ar sz), o0 = P4 (Ai cz), o1 = P2 (Ai sz)
            // comp 1: v0 = P4, v1 = P2, o0 = P1, o1 = P3
            if (comp == 0) {  // writes S(a,b,+c)
              double re = v0 - o1, im = v1 + o0;
              if (origin) {
                if (cc >= 1) {
                  pr[rw.w + cc - 1] = re;
                  pi[rw.w + cc - 1] = im;
                }
              } else {
                pr[rw.w + cmax + cc] = re;
                pi[rw.w + cmax + cc] = im;
              }
            } else if (!origin && cc > 0) {  // writes S(a,b,-c)
              pr[rw.w + cmax - cc] = o0 + v1;  // P1 + P2
              pi[rw.w + cmax - cc] = v0 - o1;  // P4 - P3
            }
          }
        }
      }
    }
    __syncthreads();  // tile A / table buffers are reused by the next segment
  }
  if (ma.ctaCycles && tid == 0) ma.ctaCycles[blockIdx.x] = clock64() - tStart;
}

}  // namespace gb
