// Micro-test of the building blocks of the warp-specialised forward-mode decode (wsq kernel):
//   * tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st 32x32b: thread = row = TMEM
//     lane, one 32-bit column per K element), B K-major in shared memory, 3xTF32 split;
//   * two independent 4-warp groups of one CTA, each issuing its own MMA chains into its own TMEM columns, with their
//     own mbarriers and named barriers (no __syncthreads after the prologue);
//   * a two-layer chain  h = relu(A W0^T) ; D = h W1^T  with the intermediate written back to TMEM as the next A;
//   * clock64 timings of the chain pieces.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_ts_test umma_ts_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr uint32_t TF32_MASK = 0xffffe000u;
constexpr int W_LBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ int w_off(int n, int k, int K) { return (n / 8) * ((K / 4) * W_LBO) + (k / 4) * W_LBO + (n % 8) * 16 + (k % 4) * 4; }

__device__ __forceinline__ int mbar_wait(uint32_t bar, uint32_t phase) {
  uint32_t done = 0;
  int it = 0;
  for (; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(phase)
        : "memory");
  }
  if (!done) __trap();
  return it;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Each of the two groups (4 warps) runs `reps` chains on its own inputs X_g [128][K0]:  h = relu(X W0^T), D = h W1^T
__global__ void __launch_bounds__(256) chain_kernel(const float* X, const float* W0, const float* W1, float* D, int K0, int reps,
                                                    long long* clocks, int* polls) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = warp >> 2, qd = warp & 3, gt = tid & 127;
  unsigned char* w0_hi = smem;
  unsigned char* w0_lo = w0_hi + 64 * 64 * 4;
  unsigned char* w1_hi = w0_lo + 64 * 64 * 4;
  unsigned char* w1_lo = w1_hi + 64 * 64 * 4;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int e = tid; e < 64 * K0; e += 256) {
    const int n = e / K0, k = e % K0;
    const float w = W0[e], h = __uint_as_float(__float_as_uint(w) & TF32_MASK);
    *reinterpret_cast<float*>(w0_hi + w_off(n, k, K0)) = h;
    *reinterpret_cast<float*>(w0_lo + w_off(n, k, K0)) = w - h;
  }
  for (int e = tid; e < 64 * 64; e += 256) {
    const int n = e / 64, k = e % 64;
    const float w = W1[e], h = __uint_as_float(__float_as_uint(w) & TF32_MASK);
    *reinterpret_cast<float*>(w1_hi + w_off(n, k, 64)) = h;
    *reinterpret_cast<float*>(w1_lo + w_off(n, k, 64)) = w - h;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- from here on the two groups never meet
  const uint32_t tb = tmem_base + grp * 256;  // group's columns: [0,64) D, [64,128) A hi, [128,192) A lo
  const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
  const uint32_t bar = smem_u32(&bars[grp]);
  uint32_t phase = 0;
  const int row = gt;  // row of the group's tile
  const float* Xg = X + (size_t)grp * 128 * K0;
  long long t_st = 0, t_mma0 = 0, t_epi = 0, t_mma1 = 0;
  int npoll = 0;
  for (int r = 0; r < reps; ++r) {
    long long c0 = clock64();
    // layer-0 A operand: this thread's row -> TMEM (hi / lo), 16 columns per store
    for (int c = 0; c < K0; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = c + j < K0 ? Xg[row * K0 + c + j] : 0.f;
        const float h = __uint_as_float(__float_as_uint(x) & TF32_MASK);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(x - h);
      }
      tmem_st16(tb + lane_base + 64 + c, hi);
      tmem_st16(tb + lane_base + 128 + c, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar(1 + grp, 128);
    long long c1 = clock64();
    if (gt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(128, 64);
      const uint32_t sbo = (K0 / 4) * W_LBO;
      for (int s = 0; s < K0 / 8; ++s) {
        const uint64_t bh = make_desc(smem_u32(w0_hi) + s * 2 * W_LBO, W_LBO, sbo);
        const uint64_t bl = make_desc(smem_u32(w0_lo) + s * 2 * W_LBO, W_LBO, sbo);
        umma_ts(tb, tb + 128 + 8 * s, bh, idesc, s > 0);
        umma_ts(tb, tb + 64 + 8 * s, bl, idesc, 1);
        umma_ts(tb, tb + 64 + 8 * s, bh, idesc, 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (lane == 0) npoll += mbar_wait(bar, phase);
    __syncwarp();
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long c2 = clock64();
    // epilogue: relu, split, next A operand
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16], hi[16], lo[16];
      tmem_ld16(tb + lane_base + c, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float z = fmaxf(__uint_as_float(v[j]), 0.f);
        const float h = __uint_as_float(__float_as_uint(z) & TF32_MASK);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(z - h);
      }
      tmem_st16(tb + lane_base + 64 + c, hi);
      tmem_st16(tb + lane_base + 128 + c, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar(1 + grp, 128);
    long long c3 = clock64();
    if (gt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(128, 64);
      const uint32_t sbo = (64 / 4) * W_LBO;
      for (int s = 0; s < 8; ++s) {
        const uint64_t bh = make_desc(smem_u32(w1_hi) + s * 2 * W_LBO, W_LBO, sbo);
        const uint64_t bl = make_desc(smem_u32(w1_lo) + s * 2 * W_LBO, W_LBO, sbo);
        umma_ts(tb, tb + 128 + 8 * s, bh, idesc, s > 0);
        umma_ts(tb, tb + 64 + 8 * s, bl, idesc, 1);
        umma_ts(tb, tb + 64 + 8 * s, bh, idesc, 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (lane == 0) npoll += mbar_wait(bar, phase);
    __syncwarp();
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long c4 = clock64();
    t_st += c1 - c0;
    t_mma0 += c2 - c1;
    t_epi += c3 - c2;
    t_mma1 += c4 - c3;
    if (r == reps - 1) {
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(tb + lane_base + c, v);
        for (int j = 0; j < 16; ++j) D[((size_t)grp * 128 + row) * 64 + c + j] = __uint_as_float(v[j]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar(1 + grp, 128);  // D is rewritten by the next repetition's first MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (gt == 0) {
    clocks[grp * 4 + 0] = t_st / reps;
    clocks[grp * 4 + 1] = t_mma0 / reps;
    clocks[grp * 4 + 2] = t_epi / reps;
    clocks[grp * 4 + 3] = t_mma1 / reps;
    polls[grp] = npoll / reps;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

static int run(int K0, int reps) {
  std::vector<float> X(2 * 128 * K0), W0(64 * K0), W1(64 * 64), D(2 * 128 * 64);
  srand(7 + K0);
  for (auto& x : X) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : W0) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : W1) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dX, *dW0, *dW1, *dD;
  long long* dC;
  int* dP;
  cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dW0, W0.size() * 4);
  cudaMalloc(&dW1, W1.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMalloc(&dC, 8 * 8);
  cudaMalloc(&dP, 8);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW0, W0.data(), W0.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = 4 * 64 * 64 * 4 + 1024;
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  chain_kernel<<<1, 256, smem>>>(dX, dW0, dW1, dD, K0, reps, dC, dP);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("K0 %d: CUDA error %s\n", K0, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long long C[8];
  int P[2];
  cudaMemcpy(C, dC, sizeof(C), cudaMemcpyDeviceToHost);
  cudaMemcpy(P, dP, sizeof(P), cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int g = 0; g < 2; ++g)
    for (int m = 0; m < 128; ++m) {
      double h[64];
      for (int n = 0; n < 64; ++n) {
        double s = 0;
        for (int k = 0; k < K0; ++k) s += (double)X[(g * 128 + m) * K0 + k] * W0[n * K0 + k];
        h[n] = s > 0 ? s : 0;
      }
      for (int n = 0; n < 64; ++n) {
        double s = 0;
        for (int k = 0; k < 64; ++k) s += h[k] * W1[n * 64 + k];
        maxerr = fmax(maxerr, fabs(s - D[(g * 128 + m) * 64 + n]));
        maxref = fmax(maxref, fabs(s));
      }
    }
  const bool ok = maxerr < 2e-5 * maxref;
  printf("K0 %d reps %d: max |err| %.3e (max |ref| %.2f) %s\n", K0, reps, maxerr, maxref, ok ? "OK" : "MISMATCH");
  for (int g = 0; g < 2; ++g)
    printf("  group %d clocks: st+bar %lld  mma0 %lld  epi+bar %lld  mma1 %lld   polls/rep %d\n", g, C[g * 4], C[g * 4 + 1], C[g * 4 + 2],
           C[g * 4 + 3], P[g]);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(40, 1);
  bad += run(40, 64);
  bad += run(16, 64);
  bad += run(24, 64);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
