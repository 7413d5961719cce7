// Structure factor on the INT8 tensor cores (tcgen05.mma kind::i8) -- recip algorithm 3,
// selected with gomcb200_set_recip_algo(e, 3) (the default path is the non-uniform FFT of
// nufft.cu; this direct sum stays as an independent implementation the parity tests
// cross-check at full size), exact to a stated bound (energy ~1e-12 relative) instead of
// FP64 round-off.
//
// Same real GEMM as recip_mma.cuh -- C[2*row+{r,i}][2*c+{cz,sz}] = sum_atoms A * B with
// A = (q/qs) X^a Y^b and B = Z^c -- but every operand is bounded by 1 in magnitude, so each is
// written as a fixed-point number and cut into NSL byte slices:
//     a  = sum_i d_i 2^(8i - FA)   (two's complement, top slice signed, FA = 8 NSL - 1)
//     b' = b + 1 = sum_j e_j 2^(8j - FB)   (in [0, 2]: every slice unsigned, FB = 8 NSL - 2)
//     sum_atoms a b = sum_{i,j} 2^(8(i+j) - FA - FB) sum_atoms d_i e_j  -  sum_atoms a
// Every slice product is an exact u8/s8 x u8 -> s32 MMA.  Slice pairs with i + j >= NSL - 2 are
// kept and accumulated in TMEM in NSL + 1 groups of equal weight; integer accumulation is exact
// and order-independent, and an atom chunk is at most 4096 atoms so that no group can overflow
// (NSL * 4096 * 255^2 < 2^31).  sum_atoms a of the quantised a is exact in FP64 and removes the
// +1 offset of B in the epilogue.  Chunk partials are converted to FP64 once and summed in a
// fixed order by k_recip_finish, like the DMMA path.
//   NSL = 6 (FA 47, FB 46; 26 products; dropped part < 2^-52 per unit product) when the c range
//           fits 64 accumulator columns per group (7 groups x 64 = 448 of the 512 TMEM columns);
//   NSL = 5 (FA 39, FB 38; 19 products; measured 5e-12 of max |S| on the 100k-atom box) for
//           column blocks of up to 80 columns (6 groups x 80 = 480 TMEM columns).
//
// CTA = (tile of 8 a values x 8 b values = 64 (a,b) rows = 128 real rows, column block, chunk):
//   warp 0   : cp.async.bulk loader of the two 8-entry slices of the XY phase table the tile
//              needs (8 KB per 32-atom step); warp 18: loader of the B byte planes -- separate
//              rings, so the table runs ahead of the producers independently of the MMAs;
//   warp 1   : MMA issuer (one lane): per step a prepared list of <= 12 instructions -- the B
//              planes of consecutive slices are adjacent in shared memory, so one instruction
//              with N = up to 256 covers several slice pairs whose results land in adjacent
//              accumulator groups;
//   warps 2-17: build the A byte planes of the next steps (complex product -> magic-number
//              fixed point -> PRMT byte transposition -> no-swizzle K-major canonical layout);
//   warps 2-5 also run the epilogue (tcgen05.ld, FP64 recombination, S(a,b,+-c) partials).
// Measured (tools/umma_i8_probe.cu): one M128 K32 kind::i8 instruction takes 59 / 74 / 138
// cycles at N = 64 / 128 / 256, i.e. 4.4k / 7.1k / 7.6k MAC per cycle per SM and ~90-100 B per
// cycle of shared-memory operand fetch.  With K = 32 the operands of a step are 80 KB for
// 6.2 MMAC, so this kernel is bound by operand fetch (shared with the producers' own traffic),
// not by the integer math: clock64 instrumentation shows the producer warps waiting for
// A-plane stages to be released 52 % of the time (compute 13 %, stores 23 %).
#pragma once
#include "common.cuh"
#include "recip_mma.cuh"

namespace gb {

constexpr int kI8Threads = 608;          // table loader, issuer, 16 producer warps, B loader
constexpr int kI8Producers = 512;
constexpr int kI8StepAtoms = 32;  // K of one UMMA
constexpr int kI8Pairs = 64;             // (a,b) rows per tile
constexpr int kI8PlaneA = 128 * 32;      // bytes of one A slice plane
constexpr int kI8ChunkSteps = 128;       // 4096 atoms per chunk (int32 overflow bound)
constexpr int kI8TmemCols = 512;
constexpr int kI8AR = 4;                 // A plane ring stages (producers fill two per round)
// ring depths: deep enough that neither loader is ever the pacing item (a bulk copy takes
// about as long as three steps of MMAs); the six-slice planes are larger, so fewer fit
__host__ __device__ constexpr int i8_tr(int nsl) { return nsl == 5 ? 8 : 4; }  // XY table ring
__host__ __device__ constexpr int i8_br(int nsl) { return nsl == 5 ? 6 : 4; }  // B plane ring
constexpr int kI8TabEntries = 16;        // staged XY entries per step: 8 X (a block) + 8 Y (b block)
// NSL byte slices per operand; slice pairs with i + j >= NSL - 2 are kept, in NSL + 1
// accumulator groups.  NSL = 6 (47/46 fractional bits, 26 MMAs, dropped part < 2^-52 per
// product) when a column block fits 64 TMEM columns per group; NSL = 5 (39/38 bits, 19
// MMAs, dropped part < 1e-13 per product) for blocks of up to 80 columns.
__host__ __device__ constexpr int i8_frac_a(int nsl) { return 8 * nsl - 1; }
__host__ __device__ constexpr int i8_frac_b(int nsl) { return 8 * nsl - 2; }
__host__ __device__ constexpr int i8_min_group(int nsl) { return nsl - 2; }

struct I8Args {
  const int4 *rows;      // {a, b, cmax, start}: 64 slots per tile = 8 a values x 8 b values of one
                         // (a block, b block); cmax = -1 marks an empty slot
  const int4 *tiles;     // {rowBegin, colBlock, first X entry, first Y entry (table index)}
  const double2 *tabXY;  // [step][entry < XYS][i < 4][quad < 8] (atom 4*quad + i), X carries q / qScale
  const unsigned char *zPlanes;  // [step][colBlock][slice][nb * 32]
  double *part;          // [chunk][re/im][nkStride]
  int KX1, XYS, NCB, nSteps, nChunks, nkStride, nTiles;
  int nb;                // rows of a B plane = accumulator columns per group (multiple of 16)
  int cPer;              // c values per column block (nb / 2)
  double outScale;       // qScale * 2^(8*minGroup - fracA - fracB)
  double sumScale;       // qScale (row sums of a are in units of q / qScale)
};

// no-swizzle K-major canonical layout: 8 rows x 16 B core matrices; the two K halves of a
// row group are adjacent (LBO 128 B), row groups follow at SBO 256 B
__host__ __device__ constexpr unsigned i8_off(int row, int k) {
  return (unsigned)(((row & 7) + (row >> 3) * 16 + (k >> 4) * 8) * 16 + (k & 15));
}

// ---- tables -----------------------------------------------------------------
// XY: double2 phases in [step][entry][32 atoms] order (conflict-free for the producers);
// Z: byte planes of cos / sin(c tz) with kI8FracB fractional bits.
template <int NSL>
__global__ void __launch_bounds__(256)
    k_i8_tables(int nAtoms, int nPad, int KX1, int KY1, int NCB, int cPer, double cvx,
                double cvy, double cvz, double invQScale, const double4 *__restrict__ pb,
                double2 *__restrict__ tabXY, unsigned char *__restrict__ zPlanes) {
  const int XYS = KX1 + KY1;
  const int PS = XYS + cPer * NCB;
  const int planeB = 2 * cPer * 32;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nPad * PS) return;
  // atom fastest inside a step so that the table stores coalesce
  const int step = (int)(idx / ((long long)PS * 32));
  const int rem = (int)(idx - (long long)step * PS * 32);
  const int e = rem >> 5, ai = rem & 31;
  const int atom = step * 32 + ai;
  const bool valid = atom < nAtoms;
  double4 a = valid ? pb[atom] : make_double4(0.0, 0.0, 0.0, 0.0);
  if (e < XYS) {
    double2 v = make_double2(0.0, 0.0);
    if (valid) {
      const bool isX = e < KX1;
      const double arg = isX ? (double)e * (cvx * a.x) : (double)(e - KX1) * (cvy * a.y);
      double s, c;
      sincos(arg, &s, &c);
      const double sc = isX ? a.w * invQScale : 1.0;
      v = make_double2(c * sc, s * sc);
    }
    // atom 4*quad + i sits at i*8 + quad: the producers' 128-bit loads are conflict-free
    tabXY[((size_t)step * XYS + e) * 32 + (ai & 3) * 8 + (ai >> 2)] = v;
  } else {
    const int cc = e - XYS;  // c value, 0 .. cPer*NCB-1
    const int cb = cc / cPer, cl = cc - cb * cPer;
    double s = -1.0, c = -1.0;  // padding atoms: b' = 0
    if (valid) sincos((double)cc * (cvz * a.z), &s, &c);
    unsigned char *base = zPlanes + ((size_t)step * NCB + cb) * (NSL * planeB);
    const double vals[2] = {c, s};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // magic-number fixed point: the low 48 bits of (v + 1.5 * 2^(52 - frac)) are the
      // two's-complement digits of round(v * 2^frac)
      // b' = b + 1 in [0, 2]: every digit of B is unsigned (the epilogue removes sum_atoms a)
      const double m = (vals[h] + 1.0) + 1.5 * (double)(1ll << (52 - i8_frac_b(NSL)));
      const unsigned lo = (unsigned)__double2loint(m), hi = (unsigned)__double2hiint(m);
      const unsigned off = i8_off(2 * cl + h, ai);
#pragma unroll
      for (int j = 0; j < NSL; ++j)
        base[j * planeB + off] = (unsigned char)((j < 4 ? lo >> (8 * j) : hi >> (8 * (j - 4))) & 0xffu);
    }
  }
}

// ---- PTX helpers ---------------------------------------------------------------
__device__ __forceinline__ unsigned long long i8_desc(unsigned saddr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFF) >> 4);
  d |= (unsigned long long)(128 >> 4) << 16;  // LBO
  d |= (unsigned long long)(256 >> 4) << 32;  // SBO
  d |= 1ull << 46;                            // descriptor version (Blackwell)
  return d;
}
__device__ __forceinline__ unsigned i8_idesc(int aSigned, int bSigned, int n) {
  return (2u << 4) | ((unsigned)aSigned << 7) | ((unsigned)bSigned << 10) |
         ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__device__ __forceinline__ void i8_mma(unsigned tmem, unsigned long long da, unsigned long long db,
                                       unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void i8_commit(unsigned mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
               : "memory");
}
__device__ __forceinline__ void i8_mbar_arrive(unsigned mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

template <int NSL>
__global__ void __launch_bounds__(kI8Threads, 1) k_recip_i8(I8Args ia) {
  constexpr int G = NSL + 1;               // accumulator groups
  constexpr int kI8TR = i8_tr(NSL), kI8BR = i8_br(NSL);
  constexpr int SMIN = i8_min_group(NSL);  // smallest kept i + j
  extern __shared__ __align__(128) unsigned char i8smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int planeB = ia.nb * 32;
  // rings: XY table slices (freed by the producers as soon as they have loaded them),
  // A planes and B planes (freed by the MMAs that read them)
  constexpr int tabBytes = kI8TabEntries * 32 * 16;
  unsigned char *tab0 = i8smem;
  unsigned char *aPl0 = i8smem + kI8TR * tabBytes;
  unsigned char *bPl0 = aPl0 + kI8AR * NSL * kI8PlaneA;
  unsigned long long *bars =
      reinterpret_cast<unsigned long long *>(bPl0 + kI8BR * NSL * planeB);
  // bars: tabFull[TR], tabEmpty[TR], bFull[BR], bDone[BR], aFull[AR], aDone[AR], accFull
  unsigned *tmemSlot = reinterpret_cast<unsigned *>(bars + 2 * kI8TR + 2 * kI8BR + 2 * kI8AR + 1);
  const unsigned bTabFull = smem_u32(bars), bTabEmpty = smem_u32(bars + kI8TR),
                 bBFull = smem_u32(bars + 2 * kI8TR), bBDone = smem_u32(bars + 2 * kI8TR + kI8BR),
                 bAFull = smem_u32(bars + 2 * kI8TR + 2 * kI8BR),
                 bADone = smem_u32(bars + 2 * kI8TR + 2 * kI8BR + kI8AR),
                 bAccFull = smem_u32(bars + 2 * kI8TR + 2 * kI8BR + 2 * kI8AR);

  // chunk-major order: CTAs that run together read the same few MB of tables (L2 resident)
  const int chunk = blockIdx.x / ia.nTiles, tileIdx = blockIdx.x % ia.nTiles;
  const int4 tile = ia.tiles[tileIdx];
  const int rowBegin = tile.x, cb = tile.y;
  const int step0 = chunk * kI8ChunkSteps;
  const int nSteps = min(kI8ChunkSteps, ia.nSteps - step0);

  if (tid == 0) {
    for (int s = 0; s < kI8TR; ++s) {
      mbar_init(bTabFull + 8 * s, 1);
      mbar_init(bTabEmpty + 8 * s, kI8Producers / 32);
    }
    for (int s = 0; s < kI8BR; ++s) {
      mbar_init(bBFull + 8 * s, 1);
      mbar_init(bBDone + 8 * s, 1);
    }
    for (int s = 0; s < kI8AR; ++s) {
      mbar_init(bAFull + 8 * s, kI8Producers / 32);
      mbar_init(bADone + 8 * s, 1);
    }
    mbar_init(bAccFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmemSlot)),
                 "n"(kI8TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *tmemSlot;

  if (warp == 0) {
    // ===== XY table loader (runs ahead of the producers, independent of the MMAs) =====
    if (lane == 0) {
      const double2 *xSrc = ia.tabXY + (size_t)tile.z * 32;
      const double2 *ySrc = ia.tabXY + (size_t)tile.w * 32;
      for (int s = 0; s < nSteps; ++s) {
        const int st = s % kI8TR;
        if (s >= kI8TR) mbar_wait(bTabEmpty + 8 * st, ((s / kI8TR) - 1) & 1);
        mbar_expect_tx(bTabFull + 8 * st, (unsigned)tabBytes);
        const size_t stepOff = (size_t)(step0 + s) * ia.XYS * 32;
        bulk_g2s(smem_u32(tab0 + st * tabBytes), xSrc + stepOff, tabBytes / 2, bTabFull + 8 * st);
        bulk_g2s(smem_u32(tab0 + st * tabBytes + tabBytes / 2), ySrc + stepOff, tabBytes / 2,
                 bTabFull + 8 * st);
      }
    }
  } else if (warp == kI8Threads / 32 - 1) {
    // ===== B plane loader =====
    if (lane == 0) {
      for (int s = 0; s < nSteps; ++s) {
        const int sb = s % kI8BR;
        if (s >= kI8BR) mbar_wait(bBDone + 8 * sb, ((s / kI8BR) - 1) & 1);
        mbar_expect_tx(bBFull + 8 * sb, (unsigned)(NSL * planeB));
        bulk_g2s(smem_u32(bPl0 + sb * NSL * planeB),
                 ia.zPlanes + ((size_t)(step0 + s) * ia.NCB + cb) * (size_t)(NSL * planeB),
                 (unsigned)(NSL * planeB), bBFull + 8 * sb);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const int nb = ia.nb;
      const int maxCnt = 256 / nb;  // B planes per instruction (N <= 256)
      // B planes j = jlo .. NSL-1 of slice i, as many per instruction as fit N <= 256:
      // adjacent planes form one wide operand whose column blocks are adjacent groups
      uint4 *list = reinterpret_cast<uint4 *>(
          (reinterpret_cast<unsigned long long>(tmemSlot) + 31ull) & ~15ull);
      int nList = 0;
      for (int i = 0; i < NSL; ++i) {
        int j = SMIN - i > 0 ? SMIN - i : 0;
        while (j < NSL) {
          const int left = NSL - j;
          const int cnt = left < maxCnt ? left : maxCnt;
          list[nList++] = make_uint4((unsigned)(i * kI8PlaneA) >> 4, (unsigned)(j * planeB) >> 4,
                                     i8_idesc(i == NSL - 1, 0, nb * cnt),
                                     (unsigned)(nb * (i + j - SMIN)));
          j += cnt;
        }
      }
      for (int s = 0; s < nSteps; ++s) {
        const int sa = s % kI8AR, sb = s % kI8BR;
        mbar_wait(bBFull + 8 * sb, (s / kI8BR) & 1);
        mbar_wait(bAFull + 8 * sa, (s / kI8AR) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned aBase = smem_u32(aPl0 + sa * NSL * kI8PlaneA),
                       bBase = smem_u32(bPl0 + sb * NSL * planeB);
        if (s == 0) {
          // first step: one instruction per slice pair so that the first product landing in
          // an accumulator group overwrites it (no separate TMEM clear)
          unsigned touched = 0;
#pragma unroll
          for (int i = 0; i < NSL; ++i) {
            const unsigned long long da = i8_desc(aBase + i * kI8PlaneA);
            for (int j = SMIN - i > 0 ? SMIN - i : 0; j < NSL; ++j) {
              const int g = i + j - SMIN;
              i8_mma(tmem + nb * g, da, i8_desc(bBase + j * planeB),
                     i8_idesc(i == NSL - 1, 0, nb), (touched >> g) & 1u);
              touched |= 1u << g;
            }
          }
        } else {
          // steady state: the instruction list was prepared once (the issuing thread is a
          // single lane; recomputing descriptors per step would make IT the bottleneck)
          const unsigned long long da0 = i8_desc(aBase), db0 = i8_desc(bBase);
          for (int t = 0; t < nList; ++t) {
            const uint4 en = list[t];  // {A offset >> 4, B offset >> 4, idesc, TMEM column}
            i8_mma(tmem + en.w, da0 + en.x, db0 + en.y, en.z, 1u);
          }
        }
        i8_commit(bADone + 8 * sa);
        i8_commit(bBDone + 8 * sb);
      }
      i8_commit(bAccFull);
    }
  } else {
    // ===== producers: A planes of step s from the staged XY table =====
    const int pt = tid - 64;            // 0..511
    const int quad = pt & 7;            // atoms 4*quad .. 4*quad+3 of the step
    const int pr = pt >> 3;             // 0..63: (a,b) pair of the tile
    const int4 rwp = ia.rows[rowBegin + pr];
    // offsets inside the staged slices: X entries tile.z .. +7, then Y entries tile.w .. +7
    const int offX = rwp.z < 0 ? -1 : (rwp.x - tile.z) * 32 * 16;  // empty slot: zeros
    const int offY = (8 + ia.KX1 + (rwp.y < 0 ? -rwp.y : rwp.y) - tile.w) * 32 * 16;
    const double magic = 1.5 * (double)(1ll << (52 - i8_frac_a(NSL)));
    // Re goes to row 2*pr, Im to row 2*pr+1.  A warp holds 4 pairs x 8 quads, and the two K
    // chunks of a row are 128 B apart (same banks), so the quads of the second chunk store
    // their Im word while those of the first store Re, and vice versa: the 32 words of one
    // store instruction then cover all 32 banks.
    const bool swp = quad >= 4;
    const unsigned offRe = i8_off(2 * pr, 4 * quad), offIm = i8_off(2 * pr + 1, 4 * quad);
    const unsigned offS[2] = {swp ? offIm : offRe, swp ? offRe : offIm};
    // sum over this thread's atoms of the QUANTISED a, kept as integers per byte plane (one
    // dp4a per stored word; exact); the epilogue needs it to undo the +1 offset of B
    int accRe[NSL], accIm[NSL];
#pragma unroll
    for (int j = 0; j < NSL; ++j) accRe[j] = accIm[j] = 0;
    const bool conjY = rwp.y < 0;  // same for every row of a tile (b blocks do not straddle 0)
    // two steps per round: twice the independent work per warp, one proxy fence per round
    for (int s = 0; s < nSteps; s += 2) {
      const int n2 = nSteps - s >= 2 ? 2 : 1;
      unsigned lo[2][2][4], hi[2][2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u < n2) {
          const int st = (s + u) % kI8TR;
          mbar_wait(bTabFull + 8 * st, ((s + u) / kI8TR) & 1);
          const unsigned tb = smem_u32(tab0 + st * tabBytes) + quad * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double mr = magic, mi = magic;
            if (offX >= 0) {
              const double2 xv = lds_f64x2(tb + offX + i * 128);
              const double2 yv = lds_f64x2(tb + offY + i * 128);
              // (x * y or x * conj(y)) + magic in four FMAs; the magic addend puts the result
              // on the fixed-point grid (one extra rounding of half a grid step)
              if (conjY) {
                mr = fma(xv.x, yv.x, fma(xv.y, yv.y, magic));
                mi = fma(xv.y, yv.x, fma(-xv.x, yv.y, magic));
              } else {
                mr = fma(xv.x, yv.x, fma(-xv.y, yv.y, magic));
                mi = fma(xv.x, yv.y, fma(xv.y, yv.x, magic));
              }
            }
            const double v0 = swp ? mi : mr, v1 = swp ? mr : mi;  // slot 0 / slot 1
            lo[u][0][i] = (unsigned)__double2loint(v0);
            hi[u][0][i] = (unsigned)__double2hiint(v0);
            lo[u][1][i] = (unsigned)__double2loint(v1);
            hi[u][1][i] = (unsigned)__double2hiint(v1);
          }
          __syncwarp();
          if (lane == 0) i8_mbar_arrive(bTabEmpty + 8 * st);  // table stage can be refilled
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u < n2) {
          const int sa = (s + u) % kI8AR;
          if (s + u >= kI8AR) mbar_wait(bADone + 8 * sa, (((s + u) / kI8AR) - 1) & 1);  // planes free
          unsigned char *ap = aPl0 + sa * NSL * kI8PlaneA;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            // 4x4 byte transpose: word j of the output = byte j of the four atoms
            const unsigned *w = lo[u][c], *h = hi[u][c];
            const unsigned t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[2], w[3], 0x5140);
            const unsigned t2 = __byte_perm(w[0], w[1], 0x7362), t3 = __byte_perm(w[2], w[3], 0x7362);
            const unsigned t4 = __byte_perm(h[0], h[1], 0x5140), t5 = __byte_perm(h[2], h[3], 0x5140);
            unsigned o[6];
            o[0] = __byte_perm(t0, t1, 0x5410);
            o[1] = __byte_perm(t0, t1, 0x7632);
            o[2] = __byte_perm(t2, t3, 0x5410);
            o[3] = __byte_perm(t2, t3, 0x7632);
            o[4] = __byte_perm(t4, t5, 0x5410);
            o[5] = __byte_perm(t4, t5, 0x7632);
            unsigned *dst = reinterpret_cast<unsigned *>(ap + offS[c]);
            int *acc = c ? accIm : accRe;  // slot sums; swapped back after the loop
#pragma unroll
            for (int j = 0; j < NSL; ++j) {
              dst[j * kI8PlaneA / 4] = o[j];
              acc[j] = j == NSL - 1 ? __dp4a((int)o[j], 0x01010101, acc[j])               // signed top
                                    : (int)__dp4a(o[j], 0x01010101u, (unsigned)acc[j]);  // unsigned
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic -> async proxy
      __syncwarp();
      if (lane == 0) {
        i8_mbar_arrive(bAFull + 8 * (s % kI8AR));
        if (n2 == 2) i8_mbar_arrive(bAFull + 8 * ((s + 1) % kI8AR));
      }
    }
    long long totRe = 0, totIm = 0;
#pragma unroll
    for (int j = 0; j < NSL; ++j) {
      totRe += (long long)(swp ? accIm[j] : accRe[j]) << (8 * j);
      totIm += (long long)(swp ? accRe[j] : accIm[j]) << (8 * j);
    }
    const double fscale = 1.0 / (double)(1ll << i8_frac_a(NSL));
    const double sumRe = (double)totRe * fscale, sumIm = (double)totIm * fscale;
    // row sums of a: 8 quads per row, combined through shared memory (table ring is idle now)
    double *rowSum = reinterpret_cast<double *>(tab0);  // [128 rows][8 quads]
    asm volatile("bar.sync 1, %0;" ::"n"(kI8Producers) : "memory");  // every warp left the loop
    rowSum[(2 * pr) * 8 + quad] = sumRe;
    rowSum[(2 * pr + 1) * 8 + quad] = sumIm;
    asm volatile("bar.sync 1, %0;" ::"n"(kI8Producers) : "memory");  // producers only
    // ===== epilogue (warps 2..5: one TMEM lane quarter each) =====
    if (warp < 6) {
      mbar_wait(bAccFull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;                 // TMEM lanes 32q .. 32q+31
      const int row = 32 * q + lane;          // real row: pair row>>1, component row&1
      const int4 rw = ia.rows[rowBegin + (row >> 1)];
      const int cmax = rw.z;
      const bool origin = rw.x == 0 && rw.y == 0;
      double *pre = ia.part + (size_t)(chunk * 2 + 0) * ia.nkStride;
      double *pim = ia.part + (size_t)(chunk * 2 + 1) * ia.nkStride;
      const unsigned taddr = tmem + ((unsigned)(32 * q) << 16);
      const int pieces = ia.nb / 16;
      double aSum = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) aSum += rowSum[row * 8 + t];
      aSum *= ia.sumScale;
#pragma unroll 1
      for (int pc = 0; pc < pieces; ++pc) {  // 16 columns (8 c values) at a time
        double v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = 0.0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          unsigned r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, "
              "%10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr + ia.nb * g + 16 * pc));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const double w = (double)(1ull << (8 * g));
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += (double)(int)r[c] * w;
        }
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) {
          // even lane (Ar row): v0 = P1 (Ar cz), v1 = P3 (Ar sz); odd lane: v0 = P4, v1 = P2
          // sum a (b + 1) - sum a
          const double v0 = v[2 * c2] * ia.outScale - aSum, v1 = v[2 * c2 + 1] * ia.outScale - aSum;
          const double o0 = __shfl_xor_sync(0xffffffffu, v0, 1);
          const double o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
          const int cl = 8 * pc + c2;
          const int cc = ia.cPer * cb + cl;
          if (cmax >= 0 && cl < ia.cPer && cc <= cmax) {
            if ((row & 1) == 0) {  // S(a,b,+c) = (P1 - P2, P3 + P4)
              const double re = v0 - o1, im = v1 + o0;
              if (origin) {
                if (cc >= 1) {
                  pre[rw.w + cc - 1] = re;
                  pim[rw.w + cc - 1] = im;
                }
              } else {
                pre[rw.w + cmax + cc] = re;
                pim[rw.w + cmax + cc] = im;
              }
            } else if (!origin && cc > 0) {  // S(a,b,-c) = (P1 + P2, P4 - P3)
              pre[rw.w + cmax - cc] = o0 + v1;
              pim[rw.w + cmax - cc] = v0 - o1;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "n"(kI8TmemCols)
                 : "memory");
}

}  // namespace gb
