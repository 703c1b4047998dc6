// Layout probe: tcgen05.mma.kind::f16 with the A operand in TMEM (written by the threads with tcgen05.st, lane = row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_a tmem_a.cu && ./tmem_a
// A[t][k] = t + 128 k (exact in fp16), B = identity (K-major, no swizzle), D[t][n] must come back as A[t][n].
// Prints the number of mismatches for the packing "column c holds K elements (2c, 2c+1), low half first".
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

constexpr int KT = 32, N = 16;   // two K = 16 steps

__global__ void probe(float *out) {
  __shared__ __align__(128) __half Bs[N * KT];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x;
  // B operand: element (n, kk) = (n == kk % 16 && kk / 16 == 0) -> D[t][n] = A[t][n] for the first K step;
  // second K step adds A[t][16 + n] * 0 (zeros) -- checks that step 2 reads the NEXT 8 columns without disturbing.
  // K-major chunks: c = kk / 8 at c * (N * 8) + n * 8 + kk % 8
  for (int i = t; i < N * KT; i += blockDim.x) {
    const int c = i / (N * 8), rem = i - c * (N * 8), n = rem >> 3, kk = c * 8 + (rem & 7);
    Bs[i] = __float2half((kk == n) ? 1.f : (kk == 16 + n ? 0.5f : 0.f));
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(t & ~31) << 16);
  // A: 32 fp16 per row = 16 columns at tmem + 32; D: 16 fp32 columns at tmem + 0
  uint32_t w[16];
  for (int c = 0; c < 16; ++c) {
    const __half lo = __float2half((float)(t + 128 * ((2 * c) % 16)) * ((2 * c) < 16 ? 1.f : 0.001f));
    const __half hi = __float2half((float)(t + 128 * ((2 * c + 1) % 16)) * ((2 * c + 1) < 16 ? 1.f : 0.001f));
    w[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          lane_base + 32),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
      "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lboB = N * 16;
    for (int k = 0; k < KT / 16; ++k) {
      const uint64_t bdesc = umma_desc(smem_u32(Bs) + k * 2 * lboB, lboB, 128);
      const uint32_t a_addr = tmem + 32 + k * 8;     // 8 columns (16 fp16) per K step
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
          "r"(a_addr), "l"(bdesc), "r"(idesc), "r"((uint32_t)(k > 0))
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 16; ++n) out[t * 16 + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  float *d, h[128 * 16];
  cudaMalloc(&d, sizeof(h));
  cudaMemset(d, 0, sizeof(h));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int t = 0; t < 128; ++t)
    for (int n = 0; n < 16; ++n) {
      // step 1: A[t][n] * 1; step 2: A[t][16 + n] * 0.5 with A[t][16 + n] = (t + 128 n) * 0.001
      const float a1 = (float)(t + 128 * n);
      const float a2 = __half2float(__float2half((float)(t + 128 * n) * 0.001f));
      const float want = a1 + 0.5f * a2;
      if (fabsf(h[t * 16 + n] - want) > 1e-3f * fmaxf(1.f, fabsf(want))) {
        if (bad < 12) printf("t=%d n=%d got %g want %g\n", t, n, h[t * 16 + n], want);
        ++bad;
      }
    }
  printf("mismatches: %d of 2048\n", bad);
  printf("row 5: ");
  for (int n = 0; n < 16; ++n) printf("%g ", h[5 * 16 + n]);
  printf("\nrow 37: ");
  for (int n = 0; n < 16; ++n) printf("%g ", h[37 * 16 + n]);
  printf("\n");
  return 0;
}
