// MultiParticle move on the device: trial transform (CallTranslateParticlesGPU /
// CallRotateParticlesGPU, src/GPU/TransformParticlesCUDAKernel.cuh:20-37; CPU form
// MultiParticle::CalculateTrialDistRot, src/moves/MultiParticle.h:566-715) and the
// acceptance weight MultiParticle::GetCoeff (:443-513).
//
// The random numbers are those of Random123Wrapper (src/Random123Wrapper.cpp:16-22):
// Philox4x64-10 with counter {molecule index, key value, 0, 0} and key {step, seed},
// so that a trajectory draws the same variates as the reference for the same seed.
#pragma once
#include "common.cuh"

namespace gb {

// Philox4x64-10 (Salmon et al., SC'11; constants of lib/Random123/philox.h:229-251)
struct Philox4 {
  unsigned long long v[4];
};
__device__ __forceinline__ Philox4 philox4x64_10(unsigned long long c0, unsigned long long c1,
                                                 unsigned long long k0, unsigned long long k1) {
  unsigned long long c[4] = {c0, c1, 0ull, 0ull};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) {
      k0 += 0x9E3779B97F4A7C15ull;
      k1 += 0xBB67AE8584CAA73Bull;
    }
    const unsigned long long m0 = 0xD2E7470EE14C6C93ull, m1 = 0xCA5A826395121157ull;
    unsigned long long hi0 = __umul64hi(m0, c[0]), lo0 = m0 * c[0];
    unsigned long long hi1 = __umul64hi(m1, c[2]), lo1 = m1 * c[2];
    unsigned long long n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
  }
  Philox4 o = {{c[0], c[1], c[2], c[3]}};
  return o;
}
// r123::u01<double> / r123::uneg11<double>, lib/Random123/uniform.hpp:175-215
__device__ __forceinline__ double r123_u01(unsigned long long in) {
  const double factor = 5.421010862427522170037264004349708557128906250e-20;  // 2^-64
  return __dadd_rn(__dmul_rn(__ull2double_rn(in), factor), 0.5 * factor);
}
__device__ __forceinline__ double r123_uneg11(unsigned long long in) {
  const double factor = 1.084202172485504434007452800869941711425781250e-19;  // 2^-63
  return __dadd_rn(__dmul_rn(__ll2double_rn((long long)in), factor), 0.5 * factor);
}

// BoxDimensions::WrapPBC, src/BoxDimensions.cpp:261-295 / BoxDimensionsNonOrth.cpp:268-281
__device__ __forceinline__ double wrap_scalar(double v, double ax) {
  if (v >= ax)
    v -= ax;
  else if (v < 0)
    v += ax;
  return v;
}
__device__ __forceinline__ void wrap_vec(const BoxParams &p, double &x, double &y, double &z) {
  if (p.nonOrth) {
    double ux = x * p.Bi[0] + y * p.Bi[3] + z * p.Bi[6];
    double uy = x * p.Bi[1] + y * p.Bi[4] + z * p.Bi[7];
    double uz = x * p.Bi[2] + y * p.Bi[5] + z * p.Bi[8];
    ux = wrap_scalar(ux, p.ax[0]);
    uy = wrap_scalar(uy, p.ax[1]);
    uz = wrap_scalar(uz, p.ax[2]);
    x = ux * p.B[0] + uy * p.B[3] + uz * p.B[6];
    y = ux * p.B[1] + uy * p.B[4] + uz * p.B[7];
    z = ux * p.B[2] + uy * p.B[5] + uz * p.B[8];
  } else {
    x = wrap_scalar(x, p.ax[0]);
    y = wrap_scalar(y, p.ax[1]);
    z = wrap_scalar(z, p.ax[2]);
  }
}

struct MpArgs {
  int brownian;  // 0: MultiParticle (force-biased, range test); 1: MultiParticleBrownian
  int moveType;  // 0 displace (mp::MPDISPLACE), 1 rotate (mp::MPROTATE)
  int nMolsBox;
  double max, lambdaBeta;
  unsigned long long step, seed, key;
  const int *molList, *molStart;
  const signed char *involved;  // per molecule, or null: every molecule of the box
  // reference state: coordinates, COMs, molecule force (+ reciprocal) or torque
  const double *x, *y, *z, *cx, *cy, *cz, *fx, *fy, *fz, *rfx, *rfy, *rfz;
  // outputs: trial coordinates / COMs (pre-filled with the reference), t_k or r_k, flags
  double *nx, *ny, *nz, *ncx, *ncy, *ncz, *kx, *ky, *kz;
  int *inForceRange;
};

// One thread per molecule of the box.
__global__ void __launch_bounds__(128) k_mp_transform(BoxParams p, MpArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nMolsBox) return;
  const int m = a.molList[t];
  if (a.involved && !a.involved[m]) return;
  double f[3] = {a.fx[m], a.fy[m], a.fz[m]};
  if (a.rfx) {
    f[0] += a.rfx[m];
    f[1] += a.rfy[m];
    f[2] += a.rfz[m];
  }
  double lb[3], lbmax[3], val[3] = {0.0, 0.0, 0.0};
  bool inRange = true;
  for (int d = 0; d < 3; ++d) {  // CalcRandomTransform, :545-564
    lb[d] = f[d] * a.lambdaBeta;
    lbmax[d] = lb[d] * a.max;
    inRange = inRange && fabs(lbmax[d]) > 1E-12 && fabs(lbmax[d]) < 30;
  }
  const Philox4 r = philox4x64_10((unsigned long long)m, a.key, a.step, a.seed);
  if (a.brownian) {
    // MultiParticleBrownian::CalcRandomTransform: lb * max + N(0, sqrt(2 max)), variates of
    // Random123Wrapper::GetGaussianCoords (Box-Muller, lib/Random123/boxmuller.hpp:126-137);
    // lambdaBeta carries BETA here, every molecule takes the force-biased branch
    inRange = true;
    const double PI = 3.1415926535897932;
    const double stdDev = sqrt(2.0 * a.max);
    double s0, c0, s1, c1;
    sincos(PI * r123_uneg11(r.v[0]), &s0, &c0);
    sincos(PI * r123_uneg11(r.v[2]), &s1, &c1);
    const double rad0 = sqrt(-2. * log(r123_u01(r.v[1])));
    const double rad1 = sqrt(-2. * log(r123_u01(r.v[3])));
    (void)c1;
    val[0] = lbmax[0] + (s0 * rad0) * stdDev;
    val[1] = lbmax[1] + (c0 * rad0) * stdDev;
    val[2] = lbmax[2] + (s1 * rad1) * stdDev;
  } else if (inRange) {
    for (int d = 0; d < 3; ++d)
      val[d] = log(exp(-1.0 * lbmax[d]) + 2.0 * r123_u01(r.v[d]) * sinh(lbmax[d])) / lb[d];
  }
  a.kx[m] = val[0];
  a.ky[m] = val[1];
  a.kz[m] = val[2];
  a.inForceRange[m] = inRange ? 1 : 0;
  const double cx = a.cx[m], cy = a.cy[m], cz = a.cz[m];
  const int s = a.molStart[m], e = a.molStart[m + 1];
  if (a.moveType == 1) {
    // RotateForceBiased / RotateRandom: Rodrigues matrix about the COM
    double theta, ax[3];
    if (inRange) {
      theta = sqrt(val[0] * val[0] + val[1] * val[1] + val[2] * val[2]);
      const double inv = 1.0 / theta;
      ax[0] = val[0] * inv;
      ax[1] = val[1] * inv;
      ax[2] = val[2] * inv;
    } else {
      theta = a.max * r123_uneg11(r.v[0]);
      const double u = r123_uneg11(r.v[1]);
      const double phi = 2.0 * 3.14159265358979323846 * r123_u01(r.v[2]);
      const double root = sqrt(1.0 - u * u);
      ax[0] = root * cos(phi);
      ax[1] = root * sin(phi);
      ax[2] = u;
    }
    const double c = cos(theta), sn = sin(theta), omc = 1 - c;
    double mt[3][3];
    const double cr[3][3] = {{0.0, -ax[2], ax[1]}, {ax[2], 0.0, -ax[0]}, {-ax[1], ax[0], 0.0}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        mt[i][j] = (i == j ? c : 0.0) + (sn * cr[i][j] + omc * (ax[i] * ax[j]));
    for (int i = s; i < e; ++i) {
      double x = a.x[i], y = a.y[i], z = a.z[i];
      unwrap_vec(p, x, y, z, cx, cy, cz);
      x -= cx;
      y -= cy;
      z -= cz;
      double rx = mt[0][0] * x + mt[0][1] * y + mt[0][2] * z + cx;
      double ry = mt[1][0] * x + mt[1][1] * y + mt[1][2] * z + cy;
      double rz = mt[2][0] * x + mt[2][1] * y + mt[2][2] * z + cz;
      wrap_vec(p, rx, ry, rz);
      a.nx[i] = rx;
      a.ny[i] = ry;
      a.nz[i] = rz;
    }
  } else {
    // TranslateForceBiased / TranslateRandom
    double sh[3] = {val[0], val[1], val[2]};
    if (!inRange)
      for (int d = 0; d < 3; ++d) sh[d] = a.max * r123_uneg11(r.v[d]);
    for (int i = s; i < e; ++i) {
      double x = a.x[i] + sh[0], y = a.y[i] + sh[1], z = a.z[i] + sh[2];
      wrap_vec(p, x, y, z);
      a.nx[i] = x;
      a.ny[i] = y;
      a.nz[i] = z;
    }
    double ncx = cx + sh[0], ncy = cy + sh[1], ncz = cz + sh[2];
    wrap_vec(p, ncx, ncy, ncz);
    a.ncx[m] = ncx;
    a.ncy[m] = ncy;
    a.ncz[m] = ncz;
  }
}

// GetCoeff: product over the in-range molecules of CalculateWRatio(new, old, k, max).
// Per-block products in a fixed tree order; part[block].
__global__ void __launch_bounds__(256)
    k_mp_coeff(int nMolsBox, const int *__restrict__ molList,
               const int *__restrict__ inForceRange, double max, double lBeta,
               const double *ofx, const double *ofy, const double *ofz, const double *orx,
               const double *ory, const double *orz, const double *nfx, const double *nfy,
               const double *nfz, const double *nrx, const double *nry, const double *nrz,
               const double *__restrict__ kx, const double *__restrict__ ky,
               const double *__restrict__ kz, double *__restrict__ part) {
  __shared__ double sm[256];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double w = 1.0;
  if (t < nMolsBox) {
    const int m = molList[t];
    if (inForceRange[m]) {
      double o[3] = {ofx[m], ofy[m], ofz[m]}, n[3] = {nfx[m], nfy[m], nfz[m]};
      if (orx) {
        o[0] += orx[m];
        o[1] += ory[m];
        o[2] += orz[m];
        n[0] += nrx[m];
        n[1] += nry[m];
        n[2] += nrz[m];
      }
      const double k[3] = {kx[m], ky[m], kz[m]};
      for (int d = 0; d < 3; ++d) {
        const double lbn = n[d] * lBeta, lbo = o[d] * lBeta;
        w *= lbn * exp(-lbn * k[d]) / (2.0 * sinh(lbn * max));
        w /= lbo * exp(lbo * k[d]) / (2.0 * sinh(lbo * max));
      }
    }
  }
  sm[threadIdx.x] = w;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] *= sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
// MultiParticleBrownian::GetCoeff (log of the weight ratio): sum over the molecules of
// (|old*BETA*max - k|^2 - |new*BETA*max + k|^2) / (4 max); part[block].
__global__ void __launch_bounds__(256)
    k_bm_coeff(int nMolsBox, const int *__restrict__ molList, double max, double beta,
               const double *ofx, const double *ofy, const double *ofz, const double *orx,
               const double *ory, const double *orz, const double *nfx, const double *nfy,
               const double *nfz, const double *nrx, const double *nry, const double *nrz,
               const double *__restrict__ kx, const double *__restrict__ ky,
               const double *__restrict__ kz, double *__restrict__ part) {
  __shared__ double scratch[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double w = 0.0;
  if (t < nMolsBox) {
    const int m = molList[t];
    double o[3] = {ofx[m], ofy[m], ofz[m]}, n[3] = {nfx[m], nfy[m], nfz[m]};
    if (orx) {
      o[0] += orx[m];
      o[1] += ory[m];
      o[2] += orz[m];
      n[0] += nrx[m];
      n[1] += nry[m];
      n[2] += nrz[m];
    }
    const double k[3] = {kx[m], ky[m], kz[m]};
    double so = 0.0, sn = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double ov = o[d] * beta * max - k[d], nv = n[d] * beta * max + k[d];
      so += ov * ov;
      sn += nv * nv;
    }
    const double max4 = 4.0 * max;
    w = so / max4 - sn / max4;
  }
  double s = block_sum(w, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void k_mp_coeff_final(int n, const double *__restrict__ part, double *out) {
  __shared__ double sm[256];
  double w = 1.0;
  for (int t = threadIdx.x; t < n; t += 256) w *= part[t];
  sm[threadIdx.x] = w;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] *= sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = isfinite(sm[0]) ? sm[0] : 0.0;
}

}  // namespace gb
