// Micro-test of the tcgen05 building blocks K1b uses (run on a B200 before integrating):
//   * TF32 tcgen05.mma (cta_group::1, M=128) with operands in the no-swizzle canonical shared-memory layouts,
//     3xTF32 split (a_hi b_hi + a_lo b_hi + a_hi b_lo), accumulator in TMEM, tcgen05.ld 32x32b epilogue;
//   * forward form  D[128 x N] = A[128 x K] W^T   with W [N][K] (nn.Linear layout)  -> B operand K-major;
//   * backward form D[128 x N'] = G[128 x 64] W   with the SAME shared-memory copy of W -> B operand MN-major
//     (LBO / SBO swapped, b_major = 1).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_test umma_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr uint32_t TF32_MASK = 0xffffe000u;
constexpr int A_LBO = 144;  // bytes between the 16-byte K chunks of a row group (128 + 16: conflict-free row gathers)
constexpr int W_LBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void split(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & TF32_MASK);
  lo = v - hi;
}

// byte offsets inside the canonical K-major tiles
__device__ __host__ inline int a_off(int m, int k, int K) { return (m / 8) * ((K / 4) * A_LBO) + (k / 4) * A_LBO + (m % 8) * 16 + (k % 4) * 4; }
__device__ __host__ inline int w_off(int n, int k, int K) { return (n / 8) * ((K / 4) * W_LBO) + (k / 4) * W_LBO + (n % 8) * 16 + (k % 4) * 4; }

// mode 0: D = A W^T (A [128][K], W [N][K]);  mode 1: D = A W (A [128][N], W [N][K]) -> D [128][K]
__global__ void __launch_bounds__(128) umma_kernel(const float* A, const float* W, float* D, int N, int K, int mode, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int KA = mode == 0 ? K : N;  // contraction length
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + 16 * (64 / 4) * A_LBO;
  unsigned char* w_hi = a_lo + 16 * (64 / 4) * A_LBO;
  unsigned char* w_lo = w_hi + 64 * 64 * 4;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // operands -> shared memory (hi / lo split)
  for (int e = tid; e < 128 * KA; e += 128) {
    const int m = e / KA, k = e % KA;
    float hi, lo;
    split(A[m * KA + k], hi, lo);
    *reinterpret_cast<float*>(a_hi + a_off(m, k, KA)) = hi;
    *reinterpret_cast<float*>(a_lo + a_off(m, k, KA)) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int n = e / K, k = e % K;
    float hi, lo;
    split(W[n * K + k], hi, lo);
    *reinterpret_cast<float*>(w_hi + w_off(n, k, K)) = hi;
    *reinterpret_cast<float*>(w_lo + w_off(n, k, K)) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base;
  if (tid == 0) {
    const uint32_t a_sbo = (KA / 4) * A_LBO, w_sbo = (K / 4) * W_LBO;
    const int ND = mode == 0 ? N : K;  // accumulator columns
    const uint32_t idesc = make_idesc(128, ND, mode);
    for (int s = 0; s < KA / 8; ++s) {
      const uint64_t ah = make_desc(smem_u32(a_hi) + s * 2 * A_LBO, A_LBO, a_sbo);
      const uint64_t al = make_desc(smem_u32(a_lo) + s * 2 * A_LBO, A_LBO, a_sbo);
      uint64_t bh, bl;
      if (mode == 0) {  // B[n][k] K-major: 8 k = 2 chunks per step
        bh = make_desc(smem_u32(w_hi) + s * 2 * W_LBO, W_LBO, w_sbo);
        bl = make_desc(smem_u32(w_lo) + s * 2 * W_LBO, W_LBO, w_sbo);
      } else {  // B'[n' = k][k' = n] MN-major: one step = 8 rows n = one row group; MN groups (4 k) are W_LBO apart
        if (variant == 0) {
          bh = make_desc(smem_u32(w_hi) + s * w_sbo, /*lbo = K'-group stride*/ w_sbo, /*sbo = MN-group stride*/ W_LBO);
          bl = make_desc(smem_u32(w_lo) + s * w_sbo, w_sbo, W_LBO);
        } else {
          bh = make_desc(smem_u32(w_hi) + s * w_sbo, W_LBO, w_sbo);
          bl = make_desc(smem_u32(w_lo) + s * w_sbo, W_LBO, w_sbo);
        }
      }
      umma_tf32(tmem_d, al, bh, idesc, s > 0);
      umma_tf32(tmem_d, ah, bl, idesc, 1);
      umma_tf32(tmem_d, ah, bh, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs (bounded spin: a wrong descriptor must not hang the box)
  {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, P1;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0));
    }
    if (!done) {
      if (tid == 0) printf("umma_test: mbarrier timeout\n");
      __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[64];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int ND = mode == 0 ? N : K;
  for (int j = 0; j < ND; ++j) D[tid * ND + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64));
}

static int run(int N, int K, int mode, int variant = 0) {
  const int KA = mode == 0 ? K : N, ND = mode == 0 ? N : K;
  std::vector<float> A(128 * KA), W(N * K), D(128 * ND);
  srand(1 + N + 7 * K + mode);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : W) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = 2 * 16 * 16 * A_LBO + 2 * 64 * 64 * 4 + 1024;
  cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_kernel<<<1, 128, smem>>>(dA, dW, dD, N, K, mode, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d N %d K %d: CUDA error %s\n", mode, N, K, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int j = 0; j < ND; ++j) {
      double ref = 0;
      for (int k = 0; k < KA; ++k) ref += (double)A[m * KA + k] * (mode == 0 ? W[j * K + k] : W[k * K + j]);
      maxerr = fmax(maxerr, fabs(ref - D[m * ND + j]));
      maxref = fmax(maxref, fabs(ref));
    }
  if (mode == 1) {
    // diagnostics: which W element pattern did the hardware read?  Compare D[0][j] against candidate contractions.
    printf("D[0][0..7]   =");
    for (int j = 0; j < 8; ++j) printf(" %8.4f", D[j]);
    printf("\nref[0][0..7] =");
    for (int j = 0; j < 8; ++j) {
      double r = 0;
      for (int k = 0; k < KA; ++k) r += (double)A[k] * W[k * K + j];
      printf(" %8.4f", r);
    }
    printf("\nA W^T (if K-major were used) =");
    for (int j = 0; j < 8; ++j) {
      double r = 0;
      for (int k = 0; k < KA && k < K; ++k) r += (double)A[k] * W[j * K + k];
      printf(" %8.4f", r);
    }
    printf("\n");
  }
  printf("variant %d ", variant);
  printf("mode %d (%s) N %d K %d: max |err| %.3e (max |ref| %.2f) %s\n", mode, mode ? "D = A W, B MN-major" : "D = A W^T, B K-major",
         N, K, maxerr, maxref, maxerr < 2e-5 * maxref ? "OK" : "MISMATCH");
  cudaFree(dA);
  cudaFree(dW);
  cudaFree(dD);
  return maxerr < 2e-5 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(64, 40, 0);  // layer 0 forward
  bad += run(64, 64, 0);  // layer 1 forward
  bad += run(64, 64, 1, 0);  // layer 1 backward
  bad += run(64, 48, 1, 0);  // layer 0 backward (input width padded to 48)
  bad += run(64, 64, 1, 1);
  bad += run(64, 48, 1, 1);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
