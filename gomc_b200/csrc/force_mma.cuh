// force_mma.cuh -- reciprocal-space forces on the FP64 MMA path (DMMA, sm_100a).
//
// Ewald::BoxForceReciprocal (src/Ewald.cpp:1496-1596):
//   F_i = 2 q_i sum_k prefact_k ( sin(k.r_i) R_k - cos(k.r_i) I_k ) k  -  intramolecular term
// With W_k = prefact_k conj(S_k) the bracket is Im(e^{ik.r_i} W_k).  For an
// orthogonal box e^{ik.r} = A_i(a,b) Z_i^c, A = X^a Y^b, so per (atom,row(a,b))
//   G0 = sum_c W(a,b,c) Z^c ,   G1 = sum_c c W(a,b,c) Z^c          (c in -cmax..cmax)
//   F_x = 2 q cvx sum_rows a Im(A G0),  F_y likewise with b,  F_z = 2 q cvz sum_rows Im(A G1).
// Pairing +c / -c (Z^-c = conj Z^c) makes G0, G1 a real GEMM
//   G[atom][4 row + o] = sum_{k = 2c+{cos,sin}} Zf[atom][k] * Wm[k][4 row + o]
// with 4 FMA per (atom, k-vector) -- the 8 flop N nk roofline of SURVEY.md section 8d.
// It runs on mma.sync.m8n8k4.f64; the contraction with A is a ~3 % epilogue.
//
// One CTA owns AB atoms (their Z features and X/Y tables stay in shared memory)
// and streams all row tiles of the pre-combined matrix Wm (3.9 MB for the
// 100k-atom box, L2 resident) through a TMA double buffer.  Every atom's force is
// summed in a fixed order inside one CTA: no atomics, bit-reproducible.
#pragma once
#include "common.cuh"
#include "recip_mma.cuh"

namespace gb {

constexpr int kFmThreads = 512;
constexpr int kFmWarps = 16;
constexpr int kFmRows = 16;            // rows per tile -> 64 outputs = 8 n8 tiles
constexpr int kFmN = kFmRows * 4;      // 64
constexpr int kFmWS = kFmN + 4;        // W row stride (doubles), == 4 mod 16

struct FmTile {  // per row tile
  int rowBegin;  // into rows[]
  int KT;        // k extent (multiple of 4): 2*(cmaxTile+1) rounded up
  int wOff;      // offset (doubles) of the tile's Wm block [KT][kFmWS]
  int pad;
};
constexpr int kFmKB = 80;  // k rows per streamed W block (one TMA copy)
struct FmBlock {  // a tile is streamed as ceil(KT / kFmKB) blocks
  int rowBegin;   // of the tile
  int k0, kLen;   // k range of this block (multiples of 4)
  int wOff;       // offset (doubles) of the block inside Wm
  int first, last, pad0, pad1;  // first / last block of its tile
};

// Wm[k][4 r + o] from the structure factor (see header).  One thread per (tile,k,r).
__global__ void __launch_bounds__(256)
    k_force_wmat(int nTiles, const FmTile *__restrict__ tiles, const int4 *__restrict__ rows,
                 const double *__restrict__ sumR, const double *__restrict__ sumI,
                 const double *__restrict__ prefact, double *__restrict__ wm) {
  const int tile = blockIdx.x;
  if (tile >= nTiles) return;
  const FmTile t = tiles[tile];
  for (int e = threadIdx.x; e < t.KT * kFmRows; e += blockDim.x) {
    const int k = e / kFmRows, r = e - k * kFmRows;
    const int c = k >> 1, isSin = k & 1;
    const int4 rw = rows[t.rowBegin + r];
    double q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
    if (c <= rw.z) {  // rw.z = cmax (-1 for padding rows)
      const bool origin = (rw.x == 0 && rw.y == 0);
      double wpr = 0.0, wpi = 0.0, wmr = 0.0, wmi = 0.0;
      if (origin) {
        if (c >= 1) {
          int ip = rw.w + c - 1;
          double pf = prefact[ip];
          wpr = pf * sumR[ip];
          wpi = -pf * sumI[ip];
        }
      } else {
        int ip = rw.w + rw.z + c;
        double pf = prefact[ip];
        wpr = pf * sumR[ip];
        wpi = -pf * sumI[ip];
        if (c > 0) {
          int im = rw.w + rw.z - c;
          double pm = prefact[im];
          wmr = pm * sumR[im];
          wmi = -pm * sumI[im];
        }
      }
      q1 = wpr + wmr;
      q2 = wmi - wpi;
      q3 = wpr - wmr;
      q4 = wpi + wmi;
    }
    const double cc = (double)c;
    double *o = wm + t.wOff + (size_t)k * kFmWS + 4 * r;
    if (!isSin) {  // cos feature
      o[0] = q1;
      o[1] = q4;
      o[2] = cc * q3;
      o[3] = -cc * q2;
    } else {       // sin feature
      o[0] = q2;
      o[1] = q3;
      o[2] = -cc * q4;
      o[3] = cc * q1;
    }
  }
  // zero the 4 padding columns so that the bulk copy moves defined data
  for (int k = threadIdx.x; k < t.KT; k += blockDim.x)
    for (int j = 0; j < 4; ++j) wm[t.wOff + (size_t)k * kFmWS + kFmN + j] = 0.0;
}

struct FmArgs {
  const FmBlock *blocks;
  int nBlocks;
  const FmTile *tiles;
  const int4 *rows;
  const double *wm;
  const double4 *pb;       // packed charged atoms {x,y,z,q}
  const int *chargedAtoms; // global atom index of packed atom t
  int nTiles, nAtoms;
  int KX1, KY1, KZ1;
  int ZFS;                 // Z feature stride per atom (doubles), == 4 mod 16, >= 2*KZ1 (+pad to 4)
  int XYS;                 // KX1 + KY1 (double2 per atom)
  double cvx, cvy, cvz;
};

// AB atoms per CTA (64 or 32).  Warp w: m-tile = w % (AB/8), n-part = w / (AB/8).
template <int AB>
__global__ void __launch_bounds__(kFmThreads, 1)
    k_force_recip_mma(FmArgs fa, double *__restrict__ rfx, double *__restrict__ rfy,
                      double *__restrict__ rfz) {
  constexpr int MTILES = AB / 8;
  constexpr int NSPLIT = kFmWarps / MTILES;
  constexpr int NTW = 8 / NSPLIT;  // n8 tiles per warp
  extern __shared__ __align__(16) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long mbarStore[2];
  __shared__ int2 rowAB[2][kFmRows];
  __shared__ double redF[kFmWarps][8][3];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mtile = warp % MTILES, npart = warp / MTILES;
  const int atom0 = blockIdx.x * AB;

  double *zf = reinterpret_cast<double *>(dynSmem);                 // [AB][ZFS]
  double2 *xy = reinterpret_cast<double2 *>(zf + (size_t)AB * fa.ZFS);  // [AB][XYS]
  double *wbuf0 = reinterpret_cast<double *>(xy + (size_t)AB * fa.XYS);
  const unsigned zfAddr = smem_u32(zf), xyAddr = smem_u32(xy), wAddr0 = smem_u32(wbuf0);
  const unsigned wStrideBuf = (unsigned)(kFmKB * kFmWS * 8);  // bytes per W buffer
  const unsigned bar0 = smem_u32(&mbarStore[0]);

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- per-atom tables: Z features (cos,sin interleaved) and X, Y phases ------
  const int zfEntries = fa.ZFS / 2;  // complex entries per atom incl. padding
  for (int e = tid; e < AB * zfEntries; e += kFmThreads) {
    int at = e / zfEntries, c = e - at * zfEntries;
    double s = 0.0, co = 0.0;
    if (atom0 + at < fa.nAtoms && c < fa.KZ1) {
      double4 a = fa.pb[atom0 + at];
      sincos((a.z * fa.cvz) * (double)c, &s, &co);
    }
    zf[at * fa.ZFS + 2 * c] = co;
    zf[at * fa.ZFS + 2 * c + 1] = s;
  }
  for (int e = tid; e < AB * fa.XYS; e += kFmThreads) {
    int at = e / fa.XYS, n = e - at * fa.XYS;
    double s = 0.0, co = 0.0;
    if (atom0 + at < fa.nAtoms) {
      double4 a = fa.pb[atom0 + at];
      if (n < fa.KX1)
        sincos((a.x * fa.cvx) * (double)n, &s, &co);
      else
        sincos((a.y * fa.cvy) * (double)(n - fa.KX1), &s, &co);
    }
    xy[e] = make_double2(co, s);
  }
  __syncthreads();

  // first W block
  if (tid == 0 && fa.nBlocks > 0) {
    FmBlock b0 = fa.blocks[0];
    unsigned bytes = (unsigned)(b0.kLen * kFmWS * 8);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar0, bytes);
    bulk_g2s(wAddr0, fa.wm + b0.wOff, bytes, bar0);
  }

  double fx = 0.0, fy = 0.0, fz = 0.0;  // for atom (8*mtile + lane>>2), this lane's rows
  const int myAtom = 8 * mtile + (lane >> 2);
  const unsigned aFrag = zfAddr + (unsigned)((myAtom * fa.ZFS + (lane & 3)) * 8);
  const unsigned xyAtom = xyAddr + (unsigned)(myAtom * fa.XYS) * 16u;
  unsigned phase = 0;
  double acc[NTW][2];
#pragma unroll
  for (int nt = 0; nt < NTW; ++nt) acc[nt][0] = acc[nt][1] = 0.0;

  for (int t = 0; t < fa.nBlocks; ++t) {
    const int buf = t & 1;
    const FmBlock bl = fa.blocks[t];
    __syncthreads();  // everyone finished block t-1: its buffer is free
    if (t + 1 < fa.nBlocks && tid == 0) {
      FmBlock bn = fa.blocks[t + 1];
      unsigned bytes = (unsigned)(bn.kLen * kFmWS * 8);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar0 + 8u * (buf ^ 1), bytes);
      bulk_g2s(wAddr0 + (buf ^ 1) * wStrideBuf, fa.wm + bn.wOff, bytes, bar0 + 8u * (buf ^ 1));
    }
    if (bl.first && tid < kFmRows) {  // (a,b) of the tile's rows, read in the contraction
      int4 rw = fa.rows[bl.rowBegin + tid];
      rowAB[0][tid] = make_int2(rw.x, rw.y);
    }
    mbar_wait(bar0 + 8u * buf, (phase >> buf) & 1u);
    phase ^= 1u << buf;

    const unsigned bFrag = wAddr0 + buf * wStrideBuf +
                           (unsigned)(((lane & 3) * kFmWS + 8 * (npart * NTW) + (lane >> 2)) * 8);
    const unsigned aBlk = aFrag + (unsigned)bl.k0 * 8u;
    const int nK4 = bl.kLen >> 2;
    for (int k4 = 0; k4 < nK4; ++k4) {
      double a = lds_f64(aBlk + (unsigned)k4 * 32u);
      double b[NTW];
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt)
        b[nt] = lds_f64(bFrag + (unsigned)k4 * (4u * kFmWS * 8u) + (unsigned)nt * 64u);
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) dmma_m8n8k4(acc[nt][0], acc[nt][1], a, b[nt]);
    }
    if (bl.last) {
      __syncthreads();  // rowAB of this tile visible (written at its first block)
      // contraction with A = X^a Y^b: this lane holds (G_r, G_i) of G0 (lane&1 == 0)
      // or G1 (lane&1 == 1) for row 2*ntGlobal + ((lane&3)>>1) of the tile
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        const int r = 2 * (npart * NTW + nt) + ((lane & 3) >> 1);
        const int2 ab = rowAB[0][r];
        double2 xv = lds_f64x2(xyAtom + (unsigned)ab.x * 16u);
        const int bb = ab.y < 0 ? -ab.y : ab.y;
        double2 yv = lds_f64x2(xyAtom + (unsigned)(fa.KX1 + bb) * 16u);
        if (ab.y < 0) yv.y = -yv.y;
        const double ar = xv.x * yv.x - xv.y * yv.y, ai = xv.x * yv.y + xv.y * yv.x;
        const double im = ar * acc[nt][1] + ai * acc[nt][0];
        if ((lane & 1) == 0) {
          fx = fma((double)ab.x, im, fx);
          fy = fma((double)ab.y, im, fy);
        } else {
          fz += im;
        }
        acc[nt][0] = acc[nt][1] = 0.0;
      }
    }
  }
  // ---- reduce: quad lanes (rows / G0,G1), then the NSPLIT warps of an m-tile -----
  fx += __shfl_xor_sync(0xffffffffu, fx, 1);
  fy += __shfl_xor_sync(0xffffffffu, fy, 1);
  fz += __shfl_xor_sync(0xffffffffu, fz, 1);
  fx += __shfl_xor_sync(0xffffffffu, fx, 2);
  fy += __shfl_xor_sync(0xffffffffu, fy, 2);
  fz += __shfl_xor_sync(0xffffffffu, fz, 2);
  if ((lane & 3) == 0) {
    redF[warp][lane >> 2][0] = fx;
    redF[warp][lane >> 2][1] = fy;
    redF[warp][lane >> 2][2] = fz;
  }
  __syncthreads();
  if (tid < AB) {
    const int mt = tid >> 3, a8 = tid & 7;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int np = 0; np < NSPLIT; ++np) {
      sx += redF[np * MTILES + mt][a8][0];
      sy += redF[np * MTILES + mt][a8][1];
      sz += redF[np * MTILES + mt][a8][2];
    }
    const int at = atom0 + tid;
    if (at < fa.nAtoms) {
      const double q2 = 2.0 * fa.pb[at].w;
      const int g = fa.chargedAtoms[at];
      // added to the intramolecular correction force already stored by
      // k_force_recip_intra (src/Ewald.cpp:1556-1569)
      rfx[g] += q2 * fa.cvx * sx;
      rfy[g] += q2 * fa.cvy * sy;
      rfz[g] += q2 * fa.cvz * sz;
    }
  }
}

// Intramolecular (correction) part of BoxForceReciprocal, src/Ewald.cpp:1556-1569;
// initialises the reciprocal force of every atom of the box (0 for uncharged atoms).
__global__ void __launch_bounds__(128)
    k_force_recip_intra(BoxParams p, int nBoxAtoms, const int *__restrict__ atomList,
                        const int *__restrict__ mol, const int *__restrict__ molStart,
                        const double *__restrict__ x, const double *__restrict__ y,
                        const double *__restrict__ z, const double *__restrict__ q,
                        double *rfx, double *rfy, double *rfz) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nBoxAtoms) return;
  int a = atomList[t];
  double X = 0.0, Y = 0.0, Z = 0.0;
  double qa = q[a];
  if (!(fabs(qa) < 0.000000001)) {
    double xa = x[a], ya = y[a], za = z[a];
    int m = mol[a];
    double constValue = p.alpha * kTwoOverSqrtPi;
    for (int j = molStart[m]; j < molStart[m + 1]; ++j) {
      if (j == a) continue;
      double dx = xa - x[j], dy = ya - y[j], dz = za - z[j];
      min_image_vec(p, dx, dy, dz);
      double r2 = dx * dx + dy * dy + dz * dz;
      double dist = sqrt(r2);
      double ex = exp(-1.0 * p.alphaSq * r2);
      double f = qa * q[j] * kQQFact / r2;
      f *= (erf(p.alpha * dist) / dist) - constValue * ex;
      X -= f * dx;
      Y -= f * dy;
      Z -= f * dz;
    }
  }
  rfx[a] = X;
  rfy[a] = Y;
  rfz[a] = Z;
}

}  // namespace gb
