// nufft.cu -- structure factor and reciprocal force of an orthogonal box as non-uniform FFTs.
//
// What it replaces: the N x nk sums of Ewald::BoxReciprocalSetup / BoxReciprocalSums
// (src/Ewald.cpp:193-361; GPU build: BoxReciprocalSumsGPU, src/GPU/CalculateEwaldCUDAKernel.cu:
// 192-231) and of Ewald::BoxForceReciprocal (:1496-1596).  For an orthogonal box every
// k-vector is 2 pi (a/Lx, b/Ly, c/Lz) with integer (a, b, c), so
//     S(a,b,c) = sum_i q_i exp(2 pi i (a u_i + b v_i + c w_i)),  (u,v,w) = r / L
// is a type-1 non-uniform discrete Fourier transform.  It is evaluated as
//   1. spread: every charge is smeared onto a fine periodic grid with the separable
//      "exponential of semicircle" window psi(t) = exp(beta (sqrt(1 - (2t/w)^2) - 1)),
//      |t| <= w/2 grid spacings (Barnett, Magland, af Klinteberg 2019);
//   2. a pruned 3-D FFT of the real grid (z: real -> half spectrum, then y, then x, each
//      pass keeping only the modes |m| <= nmax of its axis);
//   3. deconvolution: division by the window's Fourier transform per axis.
// With w = 16 and an oversampling n / (2 nmax + 1) >= 1.5 the result agrees with the direct
// sum to ~1e-13 of max |S| (tests hold it to the reference at 1e-9); the cost drops from
// 4 N nk flop to ~N w^3 + n^3 log n (cfg4: 4.1e10 -> ~2e9 issued FMA).
//
// Everything is FP64 and fixed-order: the spreading is a GATHER (each thread owns grid
// points and adds the atoms of the neighbouring bins in sorted order), there are no
// atomics, two calls on the same state return identical bits.
//
// Roofline: the spread is FP64-pipe bound (DFMA; the window tables are broadcast LDS),
// the FFT passes are L2/HBM bound (16.8 MB grid for n = 128: L2 resident).
#include "nufft.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace gbn {

namespace {

constexpr int kBrick = 16;     // bin edge in grid points (>= w/2 + 1)
constexpr int kMaxW = 16;

// bumped by every (re)allocation: a caller that captured CUDA graphs over these buffers
// compares it to know that its pointers are still the ones in the graph
long long g_allocGeneration = 0;

template <typename T>
struct Buf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    ++g_allocGeneration;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  ~Buf() {
    if (p) cudaFree(p);
  }
};

// ---- per-atom set-up --------------------------------------------------------------
struct GridGeom {
  int n[3];
  int nb[3];       // bins per axis = n / kBrick
  double invL[3];  // 1 / axis
};

__device__ __forceinline__ double grid_coord(double x, double invL, int n) {
  double u = x * invL;
  u -= floor(u);  // GOMC keeps coordinates wrapped into [0, L); be safe anyway
  double t = u * (double)n;
  if (t >= (double)n) t -= (double)n;
  return t;
}

// key[i] = home bin of charged atom i, val[i] = i
__global__ void k_nufft_keys(GridGeom g, int nAtoms, const double4 *__restrict__ packed,
                             int *__restrict__ keys, int *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nAtoms) return;
  const double4 a = packed[i];
  const int bx = min((int)grid_coord(a.x, g.invL[0], g.n[0]) / kBrick, g.nb[0] - 1);
  const int by = min((int)grid_coord(a.y, g.invL[1], g.n[1]) / kBrick, g.nb[1] - 1);
  const int bz = min((int)grid_coord(a.z, g.invL[2], g.n[2]) / kBrick, g.nb[2] - 1);
  keys[i] = (bx * g.nb[1] + by) * g.nb[2] + bz;
  vals[i] = i;
}

// binStart[c] = first sorted position whose key >= c: one thread per sorted position (and one
// past the end) fills the entries of the bins in (key[t-1], key[t]].
__global__ void k_nufft_bounds(int nBins, int n, const int *__restrict__ sortedKeys,
                               int *__restrict__ binStart) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  const int prev = t == 0 ? -1 : sortedKeys[t - 1];
  const int cur = t == n ? nBins : sortedKeys[t];
  for (int c = prev + 1; c <= cur; ++c) binStart[c] = t;
}

// Window tables of the atoms in bin-sorted order.  One thread per (atom, axis, j):
// tab[(s*3 + d)*W + j] = psi(x0_d + j - t_d), with q folded into the x table when
// foldQ; start[s] = {x0, y0, z0, original index}.  dtab (type 2 only): d psi / dt.
// binStart != nullptr: tables only for the atoms of the x-bin planes px0 .. px1 (cyclic: the
// planes whose stencils can reach a rank's slab of bricks).  Atoms are sorted x-plane-major,
// so these are one or two contiguous runs of the sorted order; their bounds are read from
// binStart on the device and the (fixed-size) grid strides over them.
template <int W>
__global__ void k_nufft_tables(GridGeom g, int nAtoms, double beta,
                               const double4 *__restrict__ packed,
                               const int *__restrict__ sortedIdx, int foldQ,
                               int4 *__restrict__ start, double *__restrict__ tab,
                               double *__restrict__ dtab,
                               const int *__restrict__ binStart, int px0, int px1) {
  int lo1 = 0, len1 = nAtoms, lo2 = 0, len2 = 0;
  if (binStart) {
    const int perX = g.nb[1] * g.nb[2];
    if (px0 <= px1) {
      lo1 = binStart[px0 * perX];
      len1 = binStart[(px1 + 1) * perX] - lo1;
    } else {  // wraps: planes px0 .. nb-1 and 0 .. px1
      lo1 = binStart[px0 * perX];
      len1 = nAtoms - lo1;
      len2 = binStart[(px1 + 1) * perX];
    }
  }
  const long long total = (long long)(len1 + len2) * 3 * W;
  for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < total;
       tid += (long long)gridDim.x * blockDim.x) {
  const int sl = (int)(tid / (3 * W));
  const int s = sl < len1 ? lo1 + sl : lo2 + (sl - len1);
  const int r = (int)(tid - (long long)sl * 3 * W);
  const int d = r / W, j = r - d * W;
  const int i = sortedIdx[s];
  const double4 a = packed[i];
  const double xd = d == 0 ? a.x : (d == 1 ? a.y : a.z);
  const double t = grid_coord(xd, g.invL[d], g.n[d]);
  const int x0 = (int)ceil(t - 0.5 * W);
  const double z = ((double)(x0 + j) - t) * (2.0 / W);
  const double s2 = 1.0 - z * z;
  double v = 0.0, dv = 0.0;
  if (s2 > 0.0) {
    const double rt = sqrt(s2);
    v = exp(beta * (rt - 1.0));
    // d/dt psi((x0 + j - t) 2/W) with respect to the ATOM coordinate t (grid units)
    dv = v * beta * z / rt * (2.0 / W);
  }
  if (d == 0 && foldQ) {
    v *= a.w;
    dv *= a.w;
  }
  tab[(size_t)(s * 3 + d) * W + j] = v;
  if (dtab) dtab[(size_t)(s * 3 + d) * W + j] = dv;
  if (j == 0) {
    int *st = reinterpret_cast<int *>(start + s);
    st[d] = x0;
    if (d == 0) st[3] = i;
  }
  }
}

// ---- spread (gather) ----------------------------------------------------------------
// One CTA (8 warps) = one 16 x 16 x 16 brick of the fine grid; warp w owns the 8 x 4 patch
// of (x, y) columns (w & 1, w >> 1) with all 16 z values: 512 grid points per warp, held as
// the accumulators of 4 x 2 FP64 tensor-core tiles (mma.sync.m8n8k4.f64, SASS DMMA).
// The atoms of the 27 neighbouring bins are the candidates; warp w scans the bins w, w+8,
// ... in order and contributes up to kTake atoms whose stencil reaches the brick to each
// chunk (its own list segment: no CTA-wide compaction, and the order -- segment 0..7, each
// in scan order -- is fixed, so the floating-point sums are reproducible).  The window
// tables of a chunk are staged in shared memory RESOLVED TO THE BRICK: row (slot, axis)
// holds the window value at each of the brick's 16 points of that axis (zero outside the
// stencil).  Each warp lists the atoms that reach its patch (lane-parallel test, ballot)
// and takes them four at a time: with A[x][k] = BX_k[x] * BY_k[y_t] and B[k][z] = BZ_k[z]
//   C_t,u[x][z] += sum_{k<4} A_t[x][k] B_u[k][z]      (t = 4 y rows, u = 2 z halves)
// i.e. 8 DMMA per 4 atoms fed by 5 shared-memory loads and 4 DMUL per lane -- the SIMT form
// (16 DFMA against 16 broadcast z values per atom) was bound by shared-memory wavefronts.
// Software pipeline, one barrier per round: while chunk k is accumulated, the tables of
// chunk k+1 arrive by cp.async (double buffer) and the lists of chunk k+2 are filled
// (triple buffer), so no global-memory latency sits between two barriers unhidden.
constexpr int kSpreadWarps = 8;
constexpr int kTake = 6;                        // atoms per warp per chunk
constexpr int kChunk = kSpreadWarps * kTake;    // table slots per chunk (48)
constexpr int kSlot = 3 * kBrick + 8;           // doubles per slot: X | Y | Z | pad; 448 B = 64 mod 128
constexpr size_t kSpreadSmem = sizeof(double) * 2 * (kChunk + 1) * kSlot;  // +1: the all-zero slot

__device__ __forceinline__ void cp_async8(double *dstShared, const double *srcGlobal) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dstShared)),
               "l"(srcGlobal)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int W>
__global__ void __launch_bounds__(256, 4)
    k_nufft_spread(GridGeom g, const int *__restrict__ binStart, const int4 *__restrict__ start,
                   const double *__restrict__ tab, double *__restrict__ grid, int brick0) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double *sB = reinterpret_cast<double *>(smemRaw);  // [2][kChunk + 1][kSlot]
  __shared__ int sX0[3][kChunk], sY0[3][kChunk], sZ0[3][kChunk], sSrc[3][kChunk];
  __shared__ int cnt[3][kSpreadWarps];
  __shared__ int hitList[kSpreadWarps][kChunk + 4];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ltMask = (1u << lane) - 1u;
  const int brick = (int)blockIdx.x + brick0;  // x-major; a rank's share is a slab of x
  const int bz = brick % g.nb[2];
  const int by = (brick / g.nb[2]) % g.nb[1];
  const int bx = brick / (g.nb[2] * g.nb[1]);
  const int bx0 = bx * kBrick, by0 = by * kBrick, bz0 = bz * kBrick;
  const int px = 8 * (warp & 1), py = 4 * (warp >> 1);  // patch origin inside the brick
  const int px0 = bx0 + px, py0 = by0 + py;
  const int mx = g.n[0] - 1, my = g.n[1] - 1, mz = g.n[2] - 1;
  const int fr = lane >> 2, fk = lane & 3;  // fragment row / column index, atom of the group

  double acc[4][2][2];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  // the all-zero slot of both table buffers pads the last group of four
  if (threadIdx.x < 2 * kSlot)
    sB[((size_t)(threadIdx.x / kSlot) * (kChunk + 1) + kChunk) * kSlot + threadIdx.x % kSlot] = 0.0;

  // Candidates = the atoms of the 27 neighbouring bins as one sequence of 32-atom batches
  // (bin-major); warp w takes batches w, w+8, ... so that every warp contributes about the
  // same number of atoms per chunk.
  __shared__ int nbBeg[27], nbEnd[27], nbCum[28];
  if (threadIdx.x < 27) {
    const int nbr = threadIdx.x;
    int cx = bx + nbr / 9 - 1, cy = by + (nbr / 3) % 3 - 1, cz = bz + nbr % 3 - 1;
    cx += cx < 0 ? g.nb[0] : 0;
    cx -= cx >= g.nb[0] ? g.nb[0] : 0;
    cy += cy < 0 ? g.nb[1] : 0;
    cy -= cy >= g.nb[1] ? g.nb[1] : 0;
    cz += cz < 0 ? g.nb[2] : 0;
    cz -= cz >= g.nb[2] ? g.nb[2] : 0;
    const int bin = (cx * g.nb[1] + cy) * g.nb[2] + cz;
    nbBeg[nbr] = binStart[bin];
    nbEnd[nbr] = binStart[bin + 1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < 27; ++i) {
      nbCum[i] = c;
      c += (nbEnd[i] - nbBeg[i] + 31) >> 5;
    }
    nbCum[27] = c;
  }
  __syncthreads();
  const int nBatches = nbCum[27];
  int bid = warp, bi = 0;        // next batch of this warp, its bin (warp-uniform)
  unsigned pendMask = 0u;
  int pendBase = 0;
  int pstX = 0, pstY = 0, pstZ = 0;  // this lane's candidate of the pending batch
  // the following batch, already requested from global memory (consumed one round later)
  int4 nxt = make_int4(0, 0, 0, 0);
  int nxtBase = -1, nxtValid = 0;
  auto prefetch = [&]() {
    nxtBase = -1;
    if (bid < nBatches) {
      while (bid >= nbCum[bi + 1]) ++bi;
      nxtBase = nbBeg[bi] + 32 * (bid - nbCum[bi]);
      nxtValid = nxtBase + lane < nbEnd[bi];
      if (nxtValid) nxt = start[nxtBase + lane];
      bid += kSpreadWarps;
    }
  };
  prefetch();

  // up to kTake relevant atoms of this warp's batches into list buffer lb
  auto fill = [&](int lb) {
    int nw = 0;
    while (nw < kTake) {
      if (pendMask == 0u) {
        if (nxtBase < 0) break;  // exhausted
        bool rel = false;
        if (nxtValid) {
          // offset of the brick's first point inside the stencil, unwrapped to
          // [-(kBrick-1), n-kBrick]: brick point p has stencil index o + p, so the stencil
          // reaches the brick iff o < W (n >= 64 > W + kBrick keeps this unambiguous)
          const int ox = ((bx0 - nxt.x + (kBrick - 1)) & mx) - (kBrick - 1);
          const int oy = ((by0 - nxt.y + (kBrick - 1)) & my) - (kBrick - 1);
          const int oz = ((bz0 - nxt.z + (kBrick - 1)) & mz) - (kBrick - 1);
          rel = ox < W && oy < W && oz < W;
          pstX = nxt.x;
          pstY = nxt.y;
          pstZ = nxt.z;
        }
        pendMask = __ballot_sync(0xffffffffu, rel);
        pendBase = nxtBase;
        prefetch();
        if (pendMask == 0u) continue;
      }
      const int take = min(__popc(pendMask), kTake - nw);
      const int myRank = __popc(pendMask & ltMask);
      if (((pendMask >> lane) & 1u) && myRank < take) {
        const int pos = warp * kTake + nw + myRank;
        sX0[lb][pos] = pstX;
        sY0[lb][pos] = pstY;
        sZ0[lb][pos] = pstZ;
        sSrc[lb][pos] = pendBase + lane;
      }
      for (int t = 0; t < take; ++t) pendMask &= pendMask - 1u;  // drop the taken bits
      nw += take;
    }
    if (lane == 0) cnt[lb][warp] = nw;
  };

  // thread = one (slot, axis) row: the 16 brick-resolved window values, by cp.async
  auto stage = [&](int lb, int tb) {
    if (threadIdx.x < kChunk * 3) {
      const int r = threadIdx.x / 3, d = threadIdx.x - 3 * r;
      if ((r % kTake) < cnt[lb][r / kTake]) {
        const int x0 = d == 0 ? sX0[lb][r] : (d == 1 ? sY0[lb][r] : sZ0[lb][r]);
        const int b0 = d == 0 ? bx0 : (d == 1 ? by0 : bz0);
        const int m = d == 0 ? mx : (d == 1 ? my : mz);
        const int o = ((b0 - x0 + (kBrick - 1)) & m) - (kBrick - 1);  // stencil index of point 0
        double *dst = sB + ((size_t)tb * (kChunk + 1) + r) * kSlot + d * kBrick;
        const double *src = tab + ((size_t)sSrc[lb][r] * 3 + d) * W;
#pragma unroll
        for (int pp = 0; pp < kBrick; ++pp) {
          const int p = (pp + lane) & (kBrick - 1);  // rotated per lane: rows sit 128 B apart
          const int j = p + o;
          if ((unsigned)j < (unsigned)W)
            cp_async8(dst + p, src + j);
          else
            dst[p] = 0.0;
        }
      }
    }
  };

  // this warp's hits of the chunk, four at a time through the FP64 tensor cores
  auto accumulate = [&](int lb, int tb) {
    const double *tbl = sB + (size_t)tb * (kChunk + 1) * kSlot;
    int nHit = 0;
#pragma unroll
    for (int h = 0; h < (kChunk + 31) / 32; ++h) {
      const int slot = 32 * h + lane;
      bool hit = false;
      if (slot < kChunk && (slot % kTake) < cnt[lb][slot / kTake]) {
        const int ox = ((px0 - sX0[lb][slot] + 7) & mx) - 7;  // patch: 8 wide, 4 high
        const int oy = ((py0 - sY0[lb][slot] + 3) & my) - 3;
        hit = ox < W && oy < W;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) hitList[warp][nHit + __popc(m & ltMask)] = slot;
      nHit += __popc(m);
    }
    if (lane < 4) hitList[warp][nHit + lane] = kChunk;  // pad with the all-zero slot
    __syncwarp();
    for (int g4 = 0; g4 < nHit; g4 += 4) {
      const double *row = tbl + hitList[warp][g4 + fk] * kSlot;
      const double ax = row[px + fr];
      const double2 y01 = *reinterpret_cast<const double2 *>(row + kBrick + py);
      const double2 y23 = *reinterpret_cast<const double2 *>(row + kBrick + py + 2);
      const double b0 = row[2 * kBrick + fr], b1 = row[2 * kBrick + 8 + fr];
      const double a0 = ax * y01.x, a1 = ax * y01.y, a2 = ax * y23.x, a3 = ax * y23.y;
      dmma_m8n8k4(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma_m8n8k4(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma_m8n8k4(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma_m8n8k4(acc[1][1][0], acc[1][1][1], a1, b1);
      dmma_m8n8k4(acc[2][0][0], acc[2][0][1], a2, b0);
      dmma_m8n8k4(acc[2][1][0], acc[2][1][1], a2, b1);
      dmma_m8n8k4(acc[3][0][0], acc[3][0][1], a3, b0);
      dmma_m8n8k4(acc[3][1][0], acc[3][1][1], a3, b1);
    }
    __syncwarp();
  };

  fill(0);
  __syncthreads();
  stage(0, 0);
  fill(1);
  cp_async_wait_all();
  __syncthreads();
  for (int k = 0;; ++k) {
    const int lb = k % 3;
    int total = 0;
#pragma unroll
    for (int w2 = 0; w2 < kSpreadWarps; ++w2) total += cnt[lb][w2];
    if (total == 0) break;
    stage((k + 1) % 3, (k + 1) & 1);
    accumulate(lb, k & 1);
    fill((k + 2) % 3);
    cp_async_wait_all();
    __syncthreads();
  }
  // C fragment: row = x = lane >> 2, columns z = 8u + 2 (lane & 3) + {0, 1}; tile t = y row
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double *out = grid + ((size_t)(px0 + fr) * g.n[1] + (py0 + t)) * g.n[2] + bz0 + 2 * fk;
#pragma unroll
    for (int u = 0; u < 2; ++u)
      *reinterpret_cast<double2 *>(out + 8 * u) = make_double2(acc[t][u][0], acc[t][u][1]);
  }
}

// ---- FFT passes ---------------------------------------------------------------------
// Radix-2 decimation-in-time on `lines` complex lines of length n held in shared memory
// (bit-reversed on load by the caller); exponent sign +: X[m] = sum_g x[g] e^{+2 pi i m g/n}.
// tw[j] = e^{+2 pi i j / n}, j < n/2 (shared memory).
__device__ __forceinline__ int bit_reverse(int v, int logn) {
  return (int)(__brev((unsigned)v) >> (32 - logn));
}

__device__ __forceinline__ void fft_lines(double2 *buf, int lines, int n, int logn,
                                          const double2 *tw) {
  const int half = n >> 1;
  for (int sft = 1; sft <= logn; ++sft) {
    const int hm = 1 << (sft - 1);
    const int twStep = n >> sft;
    for (int t = threadIdx.x; t < lines * half; t += blockDim.x) {
      const int line = t / half, b = t - line * half;
      const int k = b & (hm - 1);
      const int i0 = ((b >> (sft - 1)) << sft) + k;
      double2 *p = buf + (size_t)line * n;
      const double2 u = p[i0], v = p[i0 + hm], wv = tw[k * twStep];
      const double vr = v.x * wv.x - v.y * wv.y, vi = v.x * wv.y + v.y * wv.x;
      p[i0] = make_double2(u.x + vr, u.y + vi);
      p[i0 + hm] = make_double2(u.x - vr, u.y - vi);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void load_twiddles(double2 *sTw, const double2 *__restrict__ tw,
                                              int n) {
  for (int t = threadIdx.x; t < n / 2; t += blockDim.x) sTw[t] = tw[t];
}

// Pass z (forward): two real lines (x, y) and (x, y+1) as one complex line; keeps the
// modes c = 0..C1-1.  h1[(x*n2 + y)*C1 + c].  LP line pairs per CTA.
__global__ void __launch_bounds__(256)
    k_fft_z_fwd(int n1, int n2, int n3, int logn3, int C1, int LP,
                const double *__restrict__ grid, const double2 *__restrict__ tw,
                double2 *__restrict__ h1) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n3 / 2;
  load_twiddles(sTw, tw, n3);
  const long long pair0 = (long long)blockIdx.x * LP;
  const long long nPairs = (long long)n1 * n2 / 2;
  const int lines = (int)min((long long)LP, nPairs - pair0);
  for (int t = threadIdx.x; t < lines * n3; t += blockDim.x) {
    const int line = t / n3, gz = t - line * n3;
    const size_t row = (size_t)(pair0 + line) * 2;  // (x*n2 + y), y even
    buf[(size_t)line * n3 + bit_reverse(gz, logn3)] =
        make_double2(grid[row * n3 + gz], grid[(row + 1) * n3 + gz]);
  }
  __syncthreads();
  fft_lines(buf, lines, n3, logn3, sTw);
  for (int t = threadIdx.x; t < lines * C1; t += blockDim.x) {
    const int line = t / C1, c = t - line * C1;
    const double2 zc = buf[(size_t)line * n3 + c];
    const double2 zm = buf[(size_t)line * n3 + ((n3 - c) & (n3 - 1))];
    const size_t row = (size_t)(pair0 + line) * 2;
    h1[row * C1 + c] = make_double2(0.5 * (zc.x + zm.x), 0.5 * (zc.y - zm.y));
    h1[(row + 1) * C1 + c] = make_double2(0.5 * (zc.y + zm.y), -0.5 * (zc.x - zm.x));
  }
}

// Generic strided pass: for fixed outer index o and a block of CB inner columns,
// FFT along the middle axis.  in[(o*nIn + g)*C + c] for g < nIn (input length n = nIn)
// -> out[(o*nOut + mi)*C + c] = X[(mi - mOff) mod n], mi < nOut.
// Used as pass y (o = x) with nOut = 2 nmax2 + 1.
__global__ void __launch_bounds__(256)
    k_fft_mid(int n, int logn, int C, int CB, int nOut, int mOff,
              const double2 *__restrict__ in, const double2 *__restrict__ tw,
              double2 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n / 2;
  load_twiddles(sTw, tw, n);
  const int nCB = (C + CB - 1) / CB;
  const int o = blockIdx.x / nCB, c0 = (blockIdx.x % nCB) * CB;
  const int cb = min(CB, C - c0);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) {
    const int g = t / cb, j = t - g * cb;
    buf[(size_t)j * n + bit_reverse(g, logn)] = in[((size_t)o * n + g) * C + c0 + j];
  }
  __syncthreads();
  fft_lines(buf, cb, n, logn, sTw);
  for (int t = threadIdx.x; t < nOut * cb; t += blockDim.x) {
    const int mi = t / cb, j = t - mi * cb;
    out[((size_t)o * nOut + mi) * C + c0 + j] = buf[(size_t)j * n + ((mi - mOff) & (n - 1))];
  }
}

// Pass x (outermost axis): for fixed (bi, column block): in[(x*NB + bi)*C + c], x < n
// -> out[(ai*NB + bi)*C + c] = X[(ai - aOff) mod n], ai < nOut.
__global__ void __launch_bounds__(256)
    k_fft_outer(int n, int logn, int NB, int C, int CB, int nOut, int aOff,
                const double2 *__restrict__ in, const double2 *__restrict__ tw,
                double2 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n / 2;
  load_twiddles(sTw, tw, n);
  const int nCB = (C + CB - 1) / CB;
  const int bi = blockIdx.x / nCB, c0 = (blockIdx.x % nCB) * CB;
  const int cb = min(CB, C - c0);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) {
    const int g = t / cb, j = t - g * cb;
    buf[(size_t)j * n + bit_reverse(g, logn)] = in[((size_t)g * NB + bi) * C + c0 + j];
  }
  __syncthreads();
  fft_lines(buf, cb, n, logn, sTw);
  for (int t = threadIdx.x; t < nOut * cb; t += blockDim.x) {
    const int ai = t / cb, j = t - ai * cb;
    out[((size_t)ai * NB + bi) * C + c0 + j] = buf[(size_t)j * n + ((ai - aOff) & (n - 1))];
  }
}

// Deconvolve and scatter into the reference's k order.  One warp per (a, b) row.
// h3[((a + A)*NB + (b + B))*C + c], c >= 0; S(a, b, -c) = conj h3(-a, -b, c).
__global__ void __launch_bounds__(256)
    k_nufft_finish(int nRows, const int4 *__restrict__ rows, int A, int B, int NB, int C,
                   const double2 *__restrict__ h3, const double *__restrict__ dc0,
                   const double *__restrict__ dc1, const double *__restrict__ dc2,
                   double *__restrict__ outR, double *__restrict__ outI) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nRows) return;
  const int4 rw = rows[row];
  if (rw.z < 0) return;
  const int lane = threadIdx.x & 31;
  const int clo = (rw.x == 0 && rw.y == 0) ? 1 : -rw.z;
  const double dab = dc0[rw.x] * dc1[abs(rw.y)];
  for (int c = clo + lane; c <= rw.z; c += 32) {
    double2 v;
    if (c >= 0) {
      v = h3[((size_t)(rw.x + A) * NB + (rw.y + B)) * C + c];
    } else {
      v = h3[((size_t)(-rw.x + A) * NB + (-rw.y + B)) * C + (-c)];
      v.y = -v.y;
    }
    const double sc = dab * dc2[abs(c)];
    outR[rw.w + (c - clo)] = v.x * sc;
    outI[rw.w + (c - clo)] = v.y * sc;
  }
}

// ---- type 2 ---------------------------------------------------------------------------
// Potential coefficients on the pruned mode box: for every list entry k = (a, b, c):
//   coef = prefact_k conj(S_k) / (psihat(a) psihat(b) psihat(c))
// into slot (a, b, c) when c >= 0 and its conjugate into (-a, -b, -c) when c <= 0, so that
// phi(r) = sum_{all k} coef_k e^{ik.r} = sum_{half} 2 prefact (R cos + I sin)(k.r) is real.
__global__ void __launch_bounds__(256)
    k_nufft_fill(int nRows, const int4 *__restrict__ rows, int A, int B, int NB, int C,
                 const double *__restrict__ prefact, const double *__restrict__ sumR,
                 const double *__restrict__ sumI, const double *__restrict__ dc0,
                 const double *__restrict__ dc1, const double *__restrict__ dc2,
                 double2 *__restrict__ h3) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nRows) return;
  const int4 rw = rows[row];
  if (rw.z < 0) return;
  const int lane = threadIdx.x & 31;
  const int clo = (rw.x == 0 && rw.y == 0) ? 1 : -rw.z;
  const double dab = dc0[rw.x] * dc1[abs(rw.y)];
  for (int c = clo + lane; c <= rw.z; c += 32) {
    const int k = rw.w + (c - clo);
    const double sc = prefact[k] * dab * dc2[abs(c)];
    const double re = sumR[k] * sc, im = -sumI[k] * sc;  // conj(S) * scale
    if (c >= 0) h3[((size_t)(rw.x + A) * NB + (rw.y + B)) * C + c] = make_double2(re, im);
    if (c <= 0) h3[((size_t)(-rw.x + A) * NB + (-rw.y + B)) * C + (-c)] = make_double2(re, -im);
  }
}

// Inverse pass x: in[(ai*NB + bi)*C + c], ai < nIn placed at mode (ai - aOff) mod n, zero
// elsewhere -> out[(x*NB + bi)*C + c], x < n.
__global__ void __launch_bounds__(256)
    k_ifft_outer(int n, int logn, int NB, int C, int CB, int nIn, int aOff,
                 const double2 *__restrict__ in, const double2 *__restrict__ tw,
                 double2 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n / 2;
  load_twiddles(sTw, tw, n);
  const int nCB = (C + CB - 1) / CB;
  const int bi = blockIdx.x / nCB, c0 = (blockIdx.x % nCB) * CB;
  const int cb = min(CB, C - c0);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) buf[t] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int t = threadIdx.x; t < nIn * cb; t += blockDim.x) {
    const int ai = t / cb, j = t - ai * cb;
    buf[(size_t)j * n + bit_reverse((ai - aOff) & (n - 1), logn)] =
        in[((size_t)ai * NB + bi) * C + c0 + j];
  }
  __syncthreads();
  fft_lines(buf, cb, n, logn, sTw);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) {
    const int g = t / cb, j = t - g * cb;
    out[((size_t)g * NB + bi) * C + c0 + j] = buf[(size_t)j * n + g];
  }
}

// Inverse pass y: in[(o*nIn + mi)*C + c] placed at mode (mi - mOff) mod n -> out[(o*n + g)*C + c].
__global__ void __launch_bounds__(256)
    k_ifft_mid(int n, int logn, int C, int CB, int nIn, int mOff,
               const double2 *__restrict__ in, const double2 *__restrict__ tw,
               double2 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n / 2;
  load_twiddles(sTw, tw, n);
  const int nCB = (C + CB - 1) / CB;
  const int o = blockIdx.x / nCB, c0 = (blockIdx.x % nCB) * CB;
  const int cb = min(CB, C - c0);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) buf[t] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int t = threadIdx.x; t < nIn * cb; t += blockDim.x) {
    const int mi = t / cb, j = t - mi * cb;
    buf[(size_t)j * n + bit_reverse((mi - mOff) & (n - 1), logn)] =
        in[((size_t)o * nIn + mi) * C + c0 + j];
  }
  __syncthreads();
  fft_lines(buf, cb, n, logn, sTw);
  for (int t = threadIdx.x; t < n * cb; t += blockDim.x) {
    const int g = t / cb, j = t - g * cb;
    out[((size_t)o * n + g) * C + c0 + j] = buf[(size_t)j * n + g];
  }
}

// Inverse pass z: half spectra (c = 0..C1-1, Hermitian in c) of the lines (x, y) and
// (x, y+1) combined as Z = A + iB; the real / imaginary parts of its transform are the
// two real lines.  grid[(x*n2 + y)*n3 + gz].
__global__ void __launch_bounds__(256)
    k_ifft_z(int n1, int n2, int n3, int logn3, int C1, int LP, const double2 *__restrict__ h1,
             const double2 *__restrict__ tw, double *__restrict__ grid) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  double2 *sTw = reinterpret_cast<double2 *>(smemRaw);
  double2 *buf = sTw + n3 / 2;
  load_twiddles(sTw, tw, n3);
  const long long pair0 = (long long)blockIdx.x * LP;
  const long long nPairs = (long long)n1 * n2 / 2;
  const int lines = (int)min((long long)LP, nPairs - pair0);
  for (int t = threadIdx.x; t < lines * n3; t += blockDim.x) buf[t] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int t = threadIdx.x; t < lines * C1; t += blockDim.x) {
    const int line = t / C1, c = t - line * C1;
    const size_t row = (size_t)(pair0 + line) * 2;
    const double2 a = h1[row * C1 + c], b = h1[(row + 1) * C1 + c];
    double2 *p = buf + (size_t)line * n3;
    // Z_c = A_c + i B_c;  Z_{-c} = conj(A_c) + i conj(B_c)
    if (c == 0) {
      p[0] = make_double2(a.x - b.y, a.y + b.x);
    } else {
      p[bit_reverse(c, logn3)] = make_double2(a.x - b.y, a.y + b.x);
      p[bit_reverse(n3 - c, logn3)] = make_double2(a.x + b.y, -a.y + b.x);
    }
  }
  __syncthreads();
  fft_lines(buf, lines, n3, logn3, sTw);
  for (int t = threadIdx.x; t < lines * n3; t += blockDim.x) {
    const int line = t / n3, gz = t - line * n3;
    const size_t row = (size_t)(pair0 + line) * 2;
    const double2 v = buf[(size_t)line * n3 + gz];
    grid[row * n3 + gz] = v.x;
    grid[(row + 1) * n3 + gz] = v.y;
  }
}

// Interpolation with the analytic window gradient: F = -q grad phi.  One warp per atom in
// bin-sorted order (neighbouring warps read neighbouring grid regions: L1 / L2 hits).  A
// half-warp reads one (x, y) column of the stencil as 16 consecutive z values (coalesced);
// lane = (half h, z index jz) keeps its z weights in registers and accumulates, per x
// plane, P = sum_y psi_y v and Q = sum_y psi'_y v; fixed-order shuffle reduction.
template <int W>
__global__ void __launch_bounds__(256)
    k_nufft_interp_force(GridGeom g, int nAtoms, const int4 *__restrict__ start,
                         const double *__restrict__ tab, const double *__restrict__ dtab,
                         const double *__restrict__ grid, const int *__restrict__ atomIndex,
                         double *__restrict__ fx, double *__restrict__ fy,
                         double *__restrict__ fz) {
  __shared__ double sT[8][6 * W];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  if (s >= nAtoms) return;
  const int4 st = start[s];
  for (int t = lane; t < 3 * W; t += 32) {
    sT[warp][t] = tab[(size_t)s * 3 * W + t];
    sT[warp][3 * W + t] = dtab[(size_t)s * 3 * W + t];
  }
  __syncwarp();
  const double *tx = sT[warp], *ty = tx + W, *tz = ty + W;
  const double *dx = tz + W, *dy = dx + W, *dz = dy + W;
  const int mx = g.n[0] - 1, my = g.n[1] - 1, mz = g.n[2] - 1;
  const int jz = lane & 15, h = lane >> 4;
  const bool zin = jz < W;
  const double tzl = zin ? tz[jz] : 0.0, dzl = zin ? dz[jz] : 0.0;
  double ty8[W / 2], dy8[W / 2];
  int yoff[W / 2];
#pragma unroll
  for (int k = 0; k < W / 2; ++k) {
    ty8[k] = ty[2 * k + h];
    dy8[k] = dy[2 * k + h];
    yoff[k] = ((st.y + 2 * k + h) & my) * g.n[2] + ((st.z + jz) & mz);
  }
  double A = 0.0, B = 0.0, C = 0.0;
#pragma unroll 2
  for (int ix = 0; ix < W; ++ix) {
    const double *plane = grid + (size_t)((st.x + ix) & mx) * g.n[1] * g.n[2];
    double v[W / 2];
#pragma unroll
    for (int k = 0; k < W / 2; ++k) v[k] = zin ? plane[yoff[k]] : 0.0;
    double P = 0.0, Q = 0.0;
#pragma unroll
    for (int k = 0; k < W / 2; ++k) {
      P = fma(ty8[k], v[k], P);
      Q = fma(dy8[k], v[k], Q);
    }
    const double txv = tx[ix];
    A = fma(dx[ix], P, A);
    B = fma(txv, Q, B);
    C = fma(txv, P, C);
  }
  double ax = tzl * A, ay = tzl * B, az = dzl * C;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, o);
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    az += __shfl_xor_sync(0xffffffffu, az, o);
  }
  if (lane == 0) {
    // the x tables carry q; d psi / dt is per grid unit: F = -q grad phi, d/dx = (n / L) d/dt
    const int a = atomIndex[st.w];
    fx[a] += -ax * (double)g.n[0] * g.invL[0];
    fy[a] += -ay * (double)g.n[1] * g.invL[1];
    fz[a] += -az * (double)g.n[2] * g.invL[2];
  }
}

// ---- host side ------------------------------------------------------------------------
double es_window(double z, double beta) {
  if (std::fabs(z) >= 1.0) return 0.0;
  return std::exp(beta * (std::sqrt(1.0 - z * z) - 1.0));
}

// 1 / psihat(a), a = 0..n/2:  psihat(a) = (w/2) int_{-1}^{1} es(z) cos(pi a w z / n) dz
// by Gauss-Legendre quadrature (nodes by Newton iteration on P_m).
std::vector<double> deconv_table(int n, int w, double beta) {
  const int m = 192;
  std::vector<long double> xs(m), ws(m);
  for (int i = 0; i < m; ++i) {
    long double x = std::cos(3.14159265358979323846264338327950288L * (i + 0.75L) / (m + 0.5L));
    long double pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1.0L, p2 = 0.0L;
      for (int j = 0; j < m; ++j) {
        long double p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j + 1.0L) * x * p2 - j * p3) / (j + 1.0L);
      }
      pp = m * (x * p1 - p2) / (x * x - 1.0L);
      long double dx = p1 / pp;
      x -= dx;
      if (std::fabs((double)dx) < 1e-19) break;
    }
    xs[i] = x;
    ws[i] = 2.0L / ((1.0L - x * x) * pp * pp);
  }
  std::vector<double> out(n / 2 + 1);
  for (int a = 0; a <= n / 2; ++a) {
    long double acc = 0.0L;
    for (int i = 0; i < m; ++i) {
      long double z = xs[i];
      long double ev = std::exp((long double)beta * (std::sqrt(1.0L - z * z) - 1.0L));
      acc += ws[i] * ev * std::cos(3.14159265358979323846264338327950288L * a * w * z / n);
    }
    out[a] = (double)(1.0L / (acc * (w * 0.5L)));
  }
  return out;
}

int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

}  // namespace

struct Nufft {
  std::string err;
  // scratch
  Buf<int> keys, keysSorted, vals, sortedIdx, binStart;
  Buf<unsigned char> cubTemp;
  Buf<int4> start;
  Buf<double> tab, dtab, grid;
  Buf<double2> h1, h2, h3;
  // cached tables
  std::map<int, double2 *> twiddle;                            // n -> device e^{2 pi i j/n}
  std::map<std::tuple<int, int, long long>, double *> deconv;  // (n, w, beta bits) -> device
  // binning of the last type-1 call
  int binnedAtoms = -1, binnedW = 0;
  ~Nufft() {
    for (auto &kv : twiddle) cudaFree(kv.second);
    for (auto &kv : deconv) cudaFree(kv.second);
  }
};

Nufft *nufft_create() { return new Nufft(); }
void nufft_destroy(Nufft *p) { delete p; }
const char *nufft_last_error(const Nufft *p) { return p ? p->err.c_str() : ""; }

long long nufft_alloc_generation() { return g_allocGeneration; }

int nufft_choose(const int nmax[3], NufftGrid *g) {
  double sigmaMin = 1e30;
  for (int d = 0; d < 3; ++d) {
    if (nmax[d] < 1) return -1;
    const int modes = 2 * nmax[d] + 1;
    int n = 64;
    while (n < 1.5 * modes) n *= 2;
    g->n[d] = n;
    g->nmax[d] = nmax[d];
    sigmaMin = std::min(sigmaMin, (double)n / modes);
  }
  // window width from the aliasing estimate ln(1/eps) / (pi sqrt(1 - 1/sigma)), measured one
  // point wider.  The reciprocal force (type 2) keeps ~1e-13: its error is judged against the
  // largest force.  The structure factor (type 1) takes 1e-11 of max |S| -- measured 2e-12 on
  // the 100k-atom box, energy 1e-14 relative, against a 1e-9 bar -- which saves two points
  // (a quarter of the spread) wherever the oversampling is above ~1.65.
  double eps = 1e-13, eps1 = 1e-11;
  if (const char *ev = getenv("GOMCB200_NUFFT_EPS")) eps = std::max(1e-14, std::min(1e-6, atof(ev)));
  if (const char *ev = getenv("GOMCB200_NUFFT_EPS1")) eps1 = std::max(1e-14, std::min(1e-6, atof(ev)));
  eps1 = std::max(eps1, eps);
  auto width = [&](double e) {
    int w = (int)std::ceil(std::log(1.0 / e) / (M_PI * std::sqrt(1.0 - 1.0 / sigmaMin))) + 1;
    return std::min(kMaxW, std::max(12, (w + 1) & ~1));
  };
  g->w = width(eps);
  g->w1 = std::min(g->w, width(eps1));
  g->beta = 0.97 * M_PI * g->w * (1.0 - 1.0 / (2.0 * sigmaMin));
  g->beta1 = 0.97 * M_PI * g->w1 * (1.0 - 1.0 / (2.0 * sigmaMin));
  return 0;
}

namespace {

#define NCK(call)                                                                      \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      char _b[256];                                                                    \
      snprintf(_b, sizeof _b, "%s:%d %s: %s", __FILE__, __LINE__, #call,               \
               cudaGetErrorString(_e));                                                \
      nf->err = _b;                                                                    \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)

int get_twiddle(Nufft *nf, int n, const double2 **out) {
  auto it = nf->twiddle.find(n);
  if (it == nf->twiddle.end()) {
    std::vector<double2> h(n / 2);
    for (int j = 0; j < n / 2; ++j) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * j / n;
      h[j] = make_double2((double)std::cos(a), (double)std::sin(a));
    }
    double2 *d = nullptr;
    NCK(cudaMalloc(&d, sizeof(double2) * (n / 2)));
    NCK(cudaMemcpy(d, h.data(), sizeof(double2) * (n / 2), cudaMemcpyHostToDevice));
    it = nf->twiddle.emplace(n, d).first;
  }
  *out = it->second;
  return 0;
}

int get_deconv(Nufft *nf, int n, int w, double beta, const double **out) {
  long long bits;
  static_assert(sizeof bits == sizeof beta, "");
  memcpy(&bits, &beta, sizeof bits);
  auto key = std::make_tuple(n, w, bits);
  auto it = nf->deconv.find(key);
  if (it == nf->deconv.end()) {
    if (nf->deconv.size() > 64) {  // volume moves walk through a few (nmax -> beta) values
      for (auto &kv : nf->deconv) cudaFree(kv.second);
      nf->deconv.clear();
    }
    std::vector<double> h = deconv_table(n, w, beta);
    double *d = nullptr;
    NCK(cudaMalloc(&d, sizeof(double) * h.size()));
    NCK(cudaMemcpy(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    it = nf->deconv.emplace(key, d).first;
  }
  *out = it->second;
  return 0;
}

GridGeom make_geom(const NufftGrid &g, const double L[3]) {
  GridGeom gg;
  for (int d = 0; d < 3; ++d) {
    gg.n[d] = g.n[d];
    gg.nb[d] = g.n[d] / kBrick;
    gg.invL[d] = 1.0 / L[d];
  }
  return gg;
}

template <int W>
void launch_tables(const GridGeom &gg, int nAtoms, double beta, const double4 *packed,
                   const int *sortedIdx, int4 *start, double *tab, double *dtab,
                   cudaStream_t st, const int *sortedKeys /* binStart */ = nullptr, int px0 = 0,
                   int px1 = 0) {
  const long long total = (long long)nAtoms * 3 * W;
  const long long blocks = std::min<long long>((total + 255) / 256, 148LL * 32);
  k_nufft_tables<W><<<(unsigned)blocks, 256, 0, st>>>(
      gg, nAtoms, beta, packed, sortedIdx, 1, start, tab, dtab, sortedKeys, px0, px1);
}

// bin the atoms and build their window tables (dtab too when wantD)
// bx0 < bx1: tables only for the atoms that can reach the bricks bx0 .. bx1-1 along x
int bin_atoms(Nufft *nf, cudaStream_t st, const NufftGrid &g, const GridGeom &gg,
              const double4 *packed, int nAtoms, bool wantD, long long *launches, int bx0 = 0,
              int bx1 = 0) {
  const int nBins = gg.nb[0] * gg.nb[1] * gg.nb[2];
  NCK(nf->keys.reserve(nAtoms + 1));
  NCK(nf->keysSorted.reserve(nAtoms + 1));
  NCK(nf->vals.reserve(nAtoms + 1));
  NCK(nf->sortedIdx.reserve(nAtoms + 1));
  NCK(nf->binStart.reserve(nBins + 2));
  NCK(nf->start.reserve(nAtoms + 1));
  NCK(nf->tab.reserve((size_t)nAtoms * 3 * g.w + 16));
  if (wantD) NCK(nf->dtab.reserve((size_t)nAtoms * 3 * g.w + 16));
  const int blocks = (nAtoms + 255) / 256;
  k_nufft_keys<<<blocks, 256, 0, st>>>(gg, nAtoms, packed, nf->keys.p, nf->vals.p);
  int bits = 1;
  while ((1 << bits) < nBins + 1) ++bits;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, nf->keys.p, nf->keysSorted.p, nf->vals.p,
                                  nf->sortedIdx.p, nAtoms, 0, bits, st);
  NCK(nf->cubTemp.reserve(tmp + 16));
  NCK(cub::DeviceRadixSort::SortPairs(nf->cubTemp.p, tmp, nf->keys.p, nf->keysSorted.p,
                                      nf->vals.p, nf->sortedIdx.p, nAtoms, 0, bits, st));
  k_nufft_bounds<<<(nAtoms + 1 + 255) / 256, 256, 0, st>>>(nBins, nAtoms, nf->keysSorted.p,
                                                          nf->binStart.p);
  double *dt = wantD ? nf->dtab.p : nullptr;
  const int *sk = nullptr;
  int px0 = 0, px1 = 0;
  if (bx1 > bx0 && bx1 - bx0 + 2 < gg.nb[0]) {  // one bin of halo on either side, cyclic
    sk = nf->binStart.p;
    px0 = (bx0 - 1 + gg.nb[0]) % gg.nb[0];
    px1 = bx1 % gg.nb[0];
  }
  if (g.w == 12)
    launch_tables<12>(gg, nAtoms, g.beta, packed, nf->sortedIdx.p, nf->start.p, nf->tab.p, dt, st,
                      sk, px0, px1);
  else if (g.w == 14)
    launch_tables<14>(gg, nAtoms, g.beta, packed, nf->sortedIdx.p, nf->start.p, nf->tab.p, dt, st,
                      sk, px0, px1);
  else
    launch_tables<16>(gg, nAtoms, g.beta, packed, nf->sortedIdx.p, nf->start.p, nf->tab.p, dt, st,
                      sk, px0, px1);
  NCK(cudaGetLastError());
  *launches += 5;
  return 0;
}

size_t fft_smem(int n, int lines) { return sizeof(double2) * ((size_t)n / 2 + (size_t)lines * n); }

// lines per CTA of a strided pass: enough work per CTA, <= 64 KB of shared memory
int pick_cb(int n, int C) {
  int cb = std::max(1, std::min(8, (int)((64 * 1024 / sizeof(double2) - n / 2) / n)));
  return std::min(cb, C);
}

template <typename K>
int set_smem(Nufft *nf, K kernel, size_t bytes) {
  if (bytes > 32 * 1024)  // static + dynamic above 48 KB needs the opt-in
    NCK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

int nufft_type1(Nufft *nf, cudaStream_t st, const NufftGrid &gIn, const double L[3],
                const double4 *packed, int nAtoms, const int4 *rows, int nRows, double *outR,
                double *outI, long long *launches, const NufftShard *shard) {
  if (!nf) return -1;
  NufftGrid g = gIn;  // the type-1 window
  g.w = gIn.w1;
  g.beta = gIn.beta1;
  if (g.w != 12 && g.w != 14 && g.w != 16) {
    nf->err = "unsupported window width";
    return -1;
  }
  const GridGeom gg = make_geom(g, L);
  const int n1 = g.n[0], n2 = g.n[1], n3 = g.n[2];
  const size_t nGrid = (size_t)n1 * n2 * n3;
  const int C1 = g.nmax[2] + 1, NB = 2 * g.nmax[1] + 1, NA = 2 * g.nmax[0] + 1;
  NCK(nf->grid.reserve(nGrid));
  NCK(nf->h1.reserve((size_t)n1 * n2 * C1));
  NCK(nf->h2.reserve((size_t)n1 * NB * C1));
  NCK(nf->h3.reserve((size_t)NA * NB * C1));
  const double2 *tw1, *tw2, *tw3;
  const double *dc0, *dc1, *dc2;
  int rc;
  if ((rc = get_twiddle(nf, n1, &tw1)) || (rc = get_twiddle(nf, n2, &tw2)) ||
      (rc = get_twiddle(nf, n3, &tw3)) || (rc = get_deconv(nf, n1, g.w, g.beta, &dc0)) ||
      (rc = get_deconv(nf, n2, g.w, g.beta, &dc1)) || (rc = get_deconv(nf, n3, g.w, g.beta, &dc2)))
    return rc;
  // GOMCB200_NUFFT_TRACE=1: per-phase device times of this call on stderr (debugging aid)
  static const bool trace = getenv("GOMCB200_NUFFT_TRACE") != nullptr;
  cudaEvent_t tev[8];
  int nTev = 0;
  auto mark = [&]() {
    if (!trace) return;
    cudaEventCreate(&tev[nTev]);
    cudaEventRecord(tev[nTev++], st);
  };
  mark();
  // this rank's slab of bricks / planes along x (everything when not sharded)
  int bx0 = 0, bx1 = gg.nb[0];
  const bool sharded = shard && shard->world > 1 && shard->allgather &&
                       gg.nb[0] % shard->world == 0;
  if (sharded) {
    const int per = gg.nb[0] / shard->world;
    bx0 = shard->rank * per;
    bx1 = bx0 + per;
  }
  const int x0 = bx0 * kBrick, nx = (bx1 - bx0) * kBrick;  // planes of the slab
  rc = bin_atoms(nf, st, g, gg, packed, nAtoms, false, launches, sharded ? bx0 : 0,
                 sharded ? bx1 : 0);
  if (rc) return rc;
  nf->binnedAtoms = sharded ? -1 : nAtoms;  // sharded: the tables are partial
  nf->binnedW = g.w;
  mark();
  // spread
  {
    const int perX = gg.nb[1] * gg.nb[2];
    const int nBricks = (bx1 - bx0) * perX, brick0 = bx0 * perX;
    if (g.w == 12) {
      if ((rc = set_smem(nf, k_nufft_spread<12>, kSpreadSmem))) return rc;
      k_nufft_spread<12><<<nBricks, 256, kSpreadSmem, st>>>(gg, nf->binStart.p, nf->start.p, nf->tab.p, nf->grid.p, brick0);
    } else if (g.w == 14) {
      if ((rc = set_smem(nf, k_nufft_spread<14>, kSpreadSmem))) return rc;
      k_nufft_spread<14><<<nBricks, 256, kSpreadSmem, st>>>(gg, nf->binStart.p, nf->start.p, nf->tab.p, nf->grid.p, brick0);
    } else {
      if ((rc = set_smem(nf, k_nufft_spread<16>, kSpreadSmem))) return rc;
      k_nufft_spread<16><<<nBricks, 256, kSpreadSmem, st>>>(gg, nf->binStart.p, nf->start.p, nf->tab.p, nf->grid.p, brick0);
    }
  }
  mark();
  // pruned FFT: z and y on the slab's planes, then x over everything
  {
    const int LP = std::max(1, std::min(16, (int)((96 * 1024 / sizeof(double2) - n3 / 2) / n3)));
    const long long nPairs = (long long)nx * n2 / 2;
    const size_t smem = fft_smem(n3, LP);
    if ((rc = set_smem(nf, k_fft_z_fwd, smem))) return rc;
    k_fft_z_fwd<<<(unsigned)((nPairs + LP - 1) / LP), 256, smem, st>>>(
        nx, n2, n3, ilog2(n3), C1, LP, nf->grid.p + (size_t)x0 * n2 * n3, tw3,
        nf->h1.p + (size_t)x0 * n2 * C1);
  }
  {
    const int CB = pick_cb(n2, C1);
    const size_t smem = fft_smem(n2, CB);
    if ((rc = set_smem(nf, k_fft_mid, smem))) return rc;
    k_fft_mid<<<nx * ((C1 + CB - 1) / CB), 256, smem, st>>>(
        n2, ilog2(n2), C1, CB, NB, g.nmax[1], nf->h1.p + (size_t)x0 * n2 * C1, tw2,
        nf->h2.p + (size_t)x0 * NB * C1);
  }
  mark();
  if (sharded) {
    NCK(cudaGetLastError());
    if (shard->allgather(shard->ctx, nf->h2.p, sizeof(double2) * (size_t)nx * NB * C1, st)) {
      nf->err = "all-gather of the pruned slabs failed";
      return -1;
    }
  }
  mark();
  {
    const int CB = pick_cb(n1, C1);
    const size_t smem = fft_smem(n1, CB);
    if ((rc = set_smem(nf, k_fft_outer, smem))) return rc;
    k_fft_outer<<<NB * ((C1 + CB - 1) / CB), 256, smem, st>>>(n1, ilog2(n1), NB, C1, CB, NA,
                                                             g.nmax[0], nf->h2.p, tw1, nf->h3.p);
  }
  k_nufft_finish<<<(nRows + 7) / 8, 256, 0, st>>>(nRows, rows, g.nmax[0], g.nmax[1], NB, C1,
                                                 nf->h3.p, dc0, dc1, dc2, outR, outI);
  NCK(cudaGetLastError());
  *launches += 5;
  if (trace) {
    mark();
    cudaEventSynchronize(tev[nTev - 1]);
    const char *names[] = {"bin+tables", "spread", "fft z,y", "all-gather", "fft x + finish"};
    fprintf(stderr, "[nufft rank %d/%d]", shard ? shard->rank : 0, shard ? shard->world : 1);
    for (int i = 0; i + 1 < nTev; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
      fprintf(stderr, " %s %.3f", names[i], ms);
    }
    fprintf(stderr, "\n");
    for (int i = 0; i < nTev; ++i) cudaEventDestroy(tev[i]);
  }
  return 0;
}

int nufft_type2_force(Nufft *nf, cudaStream_t st, const NufftGrid &g, const double L[3],
                      const double4 *packed, const int *atomIndex, int nAtoms,
                      const int4 *rows, int nRows, const double *prefact, const double *sumR,
                      const double *sumI, double *fx, double *fy, double *fz, int reuseBins,
                      long long *launches) {
  if (!nf) return -1;
  if (g.w != 12 && g.w != 14 && g.w != 16) {
    nf->err = "unsupported window width";
    return -1;
  }
  (void)reuseBins;
  const GridGeom gg = make_geom(g, L);
  const int n1 = g.n[0], n2 = g.n[1], n3 = g.n[2];
  const size_t nGrid = (size_t)n1 * n2 * n3;
  const int C1 = g.nmax[2] + 1, NB = 2 * g.nmax[1] + 1, NA = 2 * g.nmax[0] + 1;
  NCK(nf->grid.reserve(nGrid));
  NCK(nf->h1.reserve((size_t)n1 * n2 * C1));
  NCK(nf->h2.reserve((size_t)n1 * NB * C1));
  NCK(nf->h3.reserve((size_t)NA * NB * C1));
  const double2 *tw1, *tw2, *tw3;
  const double *dc0, *dc1, *dc2;
  int rc;
  if ((rc = get_twiddle(nf, n1, &tw1)) || (rc = get_twiddle(nf, n2, &tw2)) ||
      (rc = get_twiddle(nf, n3, &tw3)) || (rc = get_deconv(nf, n1, g.w, g.beta, &dc0)) ||
      (rc = get_deconv(nf, n2, g.w, g.beta, &dc1)) || (rc = get_deconv(nf, n3, g.w, g.beta, &dc2)))
    return rc;
  rc = bin_atoms(nf, st, g, gg, packed, nAtoms, true, launches);
  if (rc) return rc;
  NCK(cudaMemsetAsync(nf->h3.p, 0, sizeof(double2) * (size_t)NA * NB * C1, st));
  k_nufft_fill<<<(nRows + 7) / 8, 256, 0, st>>>(nRows, rows, g.nmax[0], g.nmax[1], NB, C1, prefact,
                                               sumR, sumI, dc0, dc1, dc2, nf->h3.p);
  {
    const int CB = pick_cb(n1, C1);
    const size_t smem = fft_smem(n1, CB);
    if ((rc = set_smem(nf, k_ifft_outer, smem))) return rc;
    k_ifft_outer<<<NB * ((C1 + CB - 1) / CB), 256, smem, st>>>(n1, ilog2(n1), NB, C1, CB, NA,
                                                              g.nmax[0], nf->h3.p, tw1, nf->h2.p);
  }
  {
    const int CB = pick_cb(n2, C1);
    const size_t smem = fft_smem(n2, CB);
    if ((rc = set_smem(nf, k_ifft_mid, smem))) return rc;
    k_ifft_mid<<<n1 * ((C1 + CB - 1) / CB), 256, smem, st>>>(n2, ilog2(n2), C1, CB, NB, g.nmax[1],
                                                            nf->h2.p, tw2, nf->h1.p);
  }
  {
    const int LP = std::max(1, std::min(16, (int)((96 * 1024 / sizeof(double2) - n3 / 2) / n3)));
    const long long nPairs = (long long)n1 * n2 / 2;
    const size_t smem = fft_smem(n3, LP);
    if ((rc = set_smem(nf, k_ifft_z, smem))) return rc;
    k_ifft_z<<<(unsigned)((nPairs + LP - 1) / LP), 256, smem, st>>>(n1, n2, n3, ilog2(n3), C1, LP,
                                                                   nf->h1.p, tw3, nf->grid.p);
  }
  const int blocks = (nAtoms + 7) / 8;
  if (g.w == 12)
    k_nufft_interp_force<12><<<blocks, 256, 0, st>>>(gg, nAtoms, nf->start.p, nf->tab.p, nf->dtab.p,
                                                     nf->grid.p, atomIndex, fx, fy, fz);
  else if (g.w == 14)
    k_nufft_interp_force<14><<<blocks, 256, 0, st>>>(gg, nAtoms, nf->start.p, nf->tab.p, nf->dtab.p,
                                                     nf->grid.p, atomIndex, fx, fy, fz);
  else
    k_nufft_interp_force<16><<<blocks, 256, 0, st>>>(gg, nAtoms, nf->start.p, nf->tab.p, nf->dtab.p,
                                                     nf->grid.p, atomIndex, fx, fy, fz);
  NCK(cudaGetLastError());
  *launches += 6;
  return 0;
}

}  // namespace gbn
