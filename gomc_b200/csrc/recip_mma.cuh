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
// Data flow per CTA (8 warps, one work item = row tile x atom slab):
//   * k_phase_tables (pre-kernel) writes per-atom phase tables
//       [X: q*e^{i a tx}, a=0..KX][Y: e^{i b ty}, b=0..KY][Z: e^{i c tz}, c=0..ZS)
//     once per evaluation (1.16e7 sincos for the 100k-atom box);
//   * chunks of AT atoms of those tables are brought into shared memory with
//     one cp.async.bulk (TMA, SASS UBLKCP) per chunk, double-buffered on an
//     mbarrier, so table traffic overlaps the math;
//   * all threads build the A tile (complex product X^a Y^b) for the chunk;
//   * each warp owns 16 (a,b) rows x up to 40 c columns = MT(4) x NT(<=10)
//     m8n8 accumulator tiles and streams the chunk four atoms per DMMA.
// Partials per atom slab go to part[slab][re/im][k]; k_recip_finish sums them in
// slab order (deterministic).
#pragma once
#include "common.cuh"

namespace gb {

constexpr int kMmaThreads = 256;
constexpr int kMmaWarps = 8;
constexpr int kMmaMT = 4;                      // m8 tiles per warp: 16 (a,b) rows
constexpr int kMmaRowsPerWarp = kMmaMT * 4;    // 16
constexpr int kMmaRows = kMmaWarps * kMmaRowsPerWarp;  // 128 rows per tile
constexpr int kMmaMaxNT = 10;                  // n8 tiles per column block: 40 c

struct MmaArgs {
  const int4 *rows;    // {a, b, cmax, start}, sorted by cmax descending, padded
  const int4 *tiles;   // {rowBegin, c0, NT, unused}
  const int4 *items;   // {tile, atomBegin, atomEnd, slab}
  const double2 *tables;  // [atom][PS]
  int KX1, KY1;        // X and Y table lengths
  int PS;              // per-atom table stride (double2), PS % 8 == 2
  int zOff;            // offset of Z inside an atom's table = KX1 + KY1
  int RS;              // A tile row stride (double2), RS % 8 == 2, >= 128
  int AT;              // atoms per chunk (multiple of 4)
  int nkStride;
};

// ---- phase tables ----------------------------------------------------------
// One thread per (atom, table entry).  cv = 2 pi / L per axis.
__global__ void __launch_bounds__(256)
    k_phase_tables(int nAtoms, int nAtomsPadded, int KX1, int KY1, int ZS, int KZ1,
                   int PS, double cvx, double cvy, double cvz,
                   const double4 *__restrict__ pb, double2 *__restrict__ tables) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nAtomsPadded * PS;
  if (idx >= total) return;
  int atom = (int)(idx / PS), e = (int)(idx - (long long)atom * PS);
  double2 v = make_double2(0.0, 0.0);
  if (atom < nAtoms) {
    double4 a = pb[atom];
    double coord, cv, scale = 1.0;
    int n;
    bool valid = true;
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
  tables[idx] = v;
}

// ---- PTX helpers -------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA unit (bytes % 16 == 0).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int NT>
__device__ __forceinline__ void mma_chunk(double (&acc)[kMmaMT][kMmaMaxNT][2],
                                          const double *__restrict__ aBase,
                                          const double *__restrict__ zBase, int nK4,
                                          int aStride4, int zStride4) {
  // aBase/zBase already include the lane's (k = lane&3, m|n = lane>>2) offsets;
  // *Stride4 = doubles per 4 atoms
  for (int k4 = 0; k4 < nK4; ++k4) {
    double a[kMmaMT], b[NT];
#pragma unroll
    for (int mt = 0; mt < kMmaMT; ++mt) a[mt] = aBase[mt * 8];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b[nt] = zBase[nt * 8];
#pragma unroll
    for (int mt = 0; mt < kMmaMT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    aBase += aStride4;
    zBase += zStride4;
  }
}

__global__ void __launch_bounds__(kMmaThreads, 1)
    k_recip_mma(MmaArgs ma, double *__restrict__ part) {
  extern __shared__ __align__(16) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long mbar[2];
  __shared__ int2 rowAB[kMmaRows];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 item = ma.items[blockIdx.x];
  const int4 tile = ma.tiles[item.x];
  const int rowBegin = tile.x, c0 = tile.y, NT = tile.z;
  const int aBegin = item.y, aEnd = item.z, slab = item.w;
  const int AT = ma.AT, PS = ma.PS, RS = ma.RS;

  double2 *tab[2];
  tab[0] = reinterpret_cast<double2 *>(dynSmem);
  tab[1] = tab[0] + (size_t)AT * PS;
  double2 *tileA = tab[1] + (size_t)AT * PS;

  for (int r = tid; r < kMmaRows; r += kMmaThreads) {
    int4 rw = ma.rows[rowBegin + r];
    rowAB[r] = make_int2(rw.x, rw.y);
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nChunks = (aEnd - aBegin + AT - 1) / AT;
  const unsigned chunkBytes = (unsigned)((size_t)AT * PS * sizeof(double2));
  if (tid == 0 && nChunks > 0) {
    mbar_expect_tx(&mbar[0], chunkBytes);
    bulk_g2s(tab[0], ma.tables + (size_t)aBegin * PS, chunkBytes, &mbar[0]);
  }

  double acc[kMmaMT][kMmaMaxNT][2];
#pragma unroll
  for (int mt = 0; mt < kMmaMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < kMmaMaxNT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  const int RS2 = 2 * RS, PS2 = 2 * PS;
  for (int c = 0; c < nChunks; ++c) {
    const int buf = c & 1;
    if (tid == 0 && c + 1 < nChunks) {
      // the other buffer was last read (generic proxy) before the barrier that
      // ended chunk c-1; order those reads before the async-proxy overwrite
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[buf ^ 1], chunkBytes);
      bulk_g2s(tab[buf ^ 1], ma.tables + (size_t)(aBegin + (c + 1) * AT) * PS, chunkBytes,
               &mbar[buf ^ 1]);
    }
    mbar_wait(&mbar[buf], (unsigned)((c >> 1) & 1));
    const double2 *T = tab[buf];
    // ---- A tile: A[at][row] = (q X^a)(Y^b), conj(Y) for b < 0 --------------
    for (int t = tid; t < AT * kMmaRows; t += kMmaThreads) {
      int at = t >> 7, r = t & (kMmaRows - 1);
      int2 ab = rowAB[r];
      double2 xv = T[at * PS + ab.x];
      int bb = ab.y < 0 ? -ab.y : ab.y;
      double2 yv = T[at * PS + ma.KX1 + bb];
      if (ab.y < 0) yv.y = -yv.y;
      tileA[at * RS + r] =
          make_double2(xv.x * yv.x - xv.y * yv.y, xv.x * yv.y + xv.y * yv.x);
    }
    __syncthreads();
    // ---- DMMA over the chunk, four atoms per instruction ---------------------
    {
      const double *aBase = reinterpret_cast<const double *>(tileA) + (lane & 3) * RS2 +
                            warp * (kMmaRowsPerWarp * 2) + (lane >> 2);
      const double *zBase = reinterpret_cast<const double *>(T) + (lane & 3) * PS2 +
                            2 * (ma.zOff + c0) + (lane >> 2);
      const int nK4 = AT / 4;
      switch (NT) {
        case 2: mma_chunk<2>(acc, aBase, zBase, nK4, 4 * RS2, 4 * PS2); break;
        case 4: mma_chunk<4>(acc, aBase, zBase, nK4, 4 * RS2, 4 * PS2); break;
        case 6: mma_chunk<6>(acc, aBase, zBase, nK4, 4 * RS2, 4 * PS2); break;
        case 8: mma_chunk<8>(acc, aBase, zBase, nK4, 4 * RS2, 4 * PS2); break;
        default: mma_chunk<10>(acc, aBase, zBase, nK4, 4 * RS2, 4 * PS2); break;
      }
    }
    __syncthreads();
  }

  // ---- epilogue: exchange Ar*/Ai* partners (lane ^ 4), write S(a,b,+-c) --------
  double *pr = part + (size_t)(slab * 2 + 0) * ma.nkStride;
  double *pi = part + (size_t)(slab * 2 + 1) * ma.nkStride;
  const int comp = (lane >> 2) & 1;  // 0: this lane holds Ar*{cz,sz}; 1: Ai*{cz,sz}
#pragma unroll
  for (int mt = 0; mt < kMmaMT; ++mt) {
    const int r = warp * kMmaRowsPerWarp + mt * 4 + (lane >> 3);
    const int4 rw = ma.rows[rowBegin + r];
    const int cmax = rw.z;
    const bool origin = (rw.x == 0 && rw.y == 0);
#pragma unroll
    for (int nt = 0; nt < kMmaMaxNT; ++nt) {
      if (nt >= NT) break;
      double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
      double o0 = __shfl_xor_sync(0xffffffffu, v0, 4);
      double o1 = __shfl_xor_sync(0xffffffffu, v1, 4);
      const int cc = c0 + nt * 4 + (lane & 3);
      if (cc > cmax) continue;
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
        pr[rw.w + cmax - cc] = o0 + v1;   // P1 + P2
        pi[rw.w + cmax - cc] = v0 - o1;   // P4 - P3
      }
    }
  }
}

}  // namespace gb
