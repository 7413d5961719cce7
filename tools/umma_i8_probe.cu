// umma_i8_probe.cu -- feasibility probe for the integer-tensor-core (Ozaki-sliced)
// structure factor of DESIGN.md section 7: one CTA issues tcgen05.mma kind::i8
// (M=128, N=64, K=32, u8 x u8 / s8 mixes -> s32 in TMEM) on operands that CUDA-core
// threads wrote to shared memory in the no-swizzle K-major canonical layout, reads the
// accumulators back with tcgen05.ld and checks them against integer arithmetic on the
// host; then times a long back-to-back MMA stream.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_i8_probe tools/umma_i8_probe.cu
//   tools/umma_i8_probe            (prints JSON)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

#ifndef PROBE_N
#define PROBE_N 64
#endif
constexpr int M = 128, N = PROBE_N, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// K-major, no swizzle: 16-byte unit offset of (row, kchunk) = (row%8) + (row/8)*SBO + kchunk*LBO
__host__ __device__ constexpr uint32_t op_off(int row, int k, int sboUnits, int lboUnits) {
  return (uint32_t)(((row & 7) + (row >> 3) * sboUnits + (k >> 4) * lboUnits) * 16 + (k & 15));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lboBytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sboBytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor (upper word of the CUTLASS 64-bit form)
__host__ __device__ constexpr uint32_t make_idesc(int aSigned, int bSigned) {
  return (2u << 4)                      // c_format = S32
         | ((uint32_t)aSigned << 7)     // a_format: 0 u8, 1 s8
         | ((uint32_t)bSigned << 10)    // b_format
         | (0u << 15) | (0u << 16)      // K-major A and B
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}

__global__ void __launch_bounds__(128) k_probe(int aSigned, int bSigned, int reps, int mmasPerRep,
                                               int32_t *out, long long *cycles) {
  __shared__ __align__(128) uint8_t sA[M * K];
  __shared__ __align__(128) uint8_t sB[N * K];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmemBase;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands: A[m][k] = (3m + 5k + 1) mod 256, B[n][k] = (7n + 11k + 2) mod 256
  constexpr int SBO_A = 2, LBO_A = 1;  // group stride 32 B... (units of 16 B): (row/8)*2 + kchunk
  constexpr int SBO_B = 2, LBO_B = 1;
  for (int i = tid; i < M * K; i += 128) {
    int m = i / K, k = i % K;
    sA[op_off(m, k, SBO_A * 8, LBO_A * 8)] = (uint8_t)(3 * m + 5 * k + 1);
  }
  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i % K;
    sB[op_off(n, k, SBO_B * 8, LBO_B * 8)] = (uint8_t)(7 * n + 11 * k + 2);
  }
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmemBase)),
                 "n"(PROBE_N < 32 ? 32 : PROBE_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmemBase;
  // core matrices of one 8-row group: the two K chunks are adjacent (LBO = 128 B), groups
  // follow at SBO = 256 B
  const uint64_t da = make_desc(smem_u32(sA), LBO_A * 128, SBO_A * 128);
  const uint64_t db = make_desc(smem_u32(sB), LBO_B * 128, SBO_B * 128);
  const uint32_t idesc = make_idesc(aSigned, bSigned);
  uint32_t parity = 0;
  if (tid == 0) {
    mma_i8(tmem, da, db, idesc, 0);
    mma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), parity);
  parity ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // accumulators: lane = row m, column = n; warp w reads lanes 32w..32w+31
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int cb = 0; cb < N; cb += 64) {
  uint32_t v[64];
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[c]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]),
          "=r"(v[c + 5]), "=r"(v[c + 6]), "=r"(v[c + 7])
        : "r"(taddr + cb + c));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int c = 0; c < 64; ++c) out[tid * N + cb + c] = (int32_t)v[c];
  }

  // throughput: reps x mmasPerRep back-to-back MMAs on the same operands
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (tid == 0) {
      for (int i = 0; i < mmasPerRep; ++i) mma_i8(tmem, da, db, idesc, 1);
      mma_commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1;
  }
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(PROBE_N < 32 ? 32 : PROBE_N)
                 : "memory");
}

int main() {
  int32_t *dOut;
  long long *dCyc;
  CK(cudaMalloc(&dOut, M * N * sizeof(int32_t)));
  CK(cudaMalloc(&dCyc, sizeof(long long)));
  std::vector<int32_t> h(M * N);
  printf("{");
  for (int combo = 0; combo < 4; ++combo) {
    const int aS = combo & 1, bS = combo >> 1;
    const int reps = 200, per = 26;
    k_probe<<<1, 128>>>(aS, bS, reps, per, dOut, dCyc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), dOut, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dCyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        long long ref = 0;
        for (int k = 0; k < K; ++k) {
          int a = (uint8_t)(3 * m + 5 * k + 1), b = (uint8_t)(7 * n + 11 * k + 2);
          if (aS) a = (int8_t)a;
          if (bS) b = (int8_t)b;
          ref += (long long)a * b;
        }
        if (ref != h[m * N + n]) ++bad;
      }
    printf("%s\"n%d_a_%s_b_%s\": {\"mismatches\": %lld, \"cycles_per_mma\": %.1f}", combo ? ", " : "", N,
           aS ? "s8" : "u8", bS ? "s8" : "u8", bad, (double)cyc / (reps * per));
  }
  printf("}\n");
  return 0;
}
