// nufft.h -- host interface of the non-uniform FFT path (nufft.cu) used by engine.cu.
//
// Structure factor of an orthogonal box (Ewald::BoxReciprocalSetup / BoxReciprocalSums,
// src/Ewald.cpp:193-361) and the reciprocal force (Ewald::BoxForceReciprocal, :1496-1596)
// as type-1 / type-2 non-uniform FFTs: O(N w^3 + n^3 log n) instead of O(N nk).
#pragma once
#include <cuda_runtime.h>

namespace gbn {

// Fine grid and spreading window for one k set.
struct NufftGrid {
  int n[3];     // fine grid points per axis (powers of two, >= 64)
  int nmax[3];  // largest |mode| per axis (Ewald kmax per axis)
  int w;        // window width in grid points (even, <= 16): type 2 (forces), ~1e-13
  double beta;  // exponential-of-semicircle shape parameter
  int w1;       // width of the type-1 transform (structure factor / energy), ~1e-11 of
  double beta1; // max |S|: two points narrower where the oversampling allows it
};

struct Nufft;  // scratch buffers + cached tables; one per engine

Nufft *nufft_create();
void nufft_destroy(Nufft *);
const char *nufft_last_error(const Nufft *);

// Counter bumped by every (re)allocation inside this module (all engines of the process).
long long nufft_alloc_generation();

// Grid for modes |a| <= nmax[0], |b| <= nmax[1], |c| <= nmax[2] at the accuracy of the
// widest window.  Returns 0, or -1 when the path does not apply (a mode range of zero).
int nufft_choose(const int nmax[3], NufftGrid *g);

// Multi-GPU type 1 (one engine per GPU, every rank holds all atoms): rank r spreads onto its
// slab of x planes only (window tables only for the atoms that reach it), runs the z and y
// passes of the pruned FFT on that slab, the ranks all-gather the pruned slabs
// (n1 x (2 nmax_y + 1) x (nmax_z + 1) complex: 5.8 MB for the 100k-atom box instead of a
// 16.8 MB grid), and the x pass + deconvolution run replicated, so every rank ends with the
// complete sums.  allgather(ctx, buf, bytesPerRank, stream): in place, rank r's share at
// buf + r * bytesPerRank.  Applies when world divides the bricks along x.
struct NufftShard {
  int rank, world;
  int (*allgather)(void *ctx, void *buf, size_t bytesPerRank, cudaStream_t stream);
  void *ctx;
};

// Type 1.  packed[i] = {x, y, z, q} of the nAtoms charged atoms, L = box axes.
// rows[nRows] = {a, b, cmax, first}: entries first.. of the reference's k list hold
// c = -cmax..cmax (1..cmax for a = b = 0); rows with cmax < 0 are padding.
// outR/outI[k] = sum_i q_i (cos, sin)(k.r_i) in that order.
// Returns 0 on success (kernels queued on `stream`); *launches is incremented.
int nufft_type1(Nufft *, cudaStream_t stream, const NufftGrid &g, const double L[3],
                const double4 *packed, int nAtoms, const int4 *rows, int nRows, double *outR,
                double *outI, long long *launches, const NufftShard *shard = nullptr);

// Type 2.  Reciprocal force of the same atoms from the sums:
//   F_i = sum_k 2 q_i prefact_k (sin(k.r_i) R_k - cos(k.r_i) I_k) k   (src/Ewald.cpp:1540-1552)
// added in place to fx/fy/fz[atomIndex[i]] (global atom indexing).  Reuses the atom
// binning of the last nufft_type1 call when `reuseBins` (same coordinates, same grid).
int nufft_type2_force(Nufft *, cudaStream_t stream, const NufftGrid &g, const double L[3],
                      const double4 *packed, const int *atomIndex, int nAtoms,
                      const int4 *rows, int nRows, const double *prefact, const double *sumR,
                      const double *sumI, double *fx, double *fy, double *fz, int reuseBins,
                      long long *launches);

}  // namespace gbn
