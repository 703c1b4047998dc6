// Parameter-adjoint accumulation of the tensor path as a tcgen05 / TMEM split-K reduction.
//
// The tensor-path reverse kernels leave, per GP, the outer-product operands over the
// L = (live steps x particles) columns as bfloat16 hi/lo pairs in MMA-ready tiles (common.cuh TcMats):
// a_bar, k', a^2, w (M rows each), g_mean, g_var (Dout), [x~, 1] (Din+1).  The parameter adjoints are the
// four thin GEMMs over that huge K dimension
//     P_bar' = Ab K^T [M x M],  alpha_bar' = K Gm^T,  S_bar = A2 Gv^T,  [U | r] = W X1^T .
// They are memory-bound (~14 FLOP/B), so the kernel is a pure stream: each CTA owns a contiguous range of
// tiles; one producer thread brings a tile (~55 KB at M = 100) into a shared-memory ring with 14 bulk
// copies (cp.async.bulk, completion on an mbarrier) -- no thread touches the data; one thread issues
// 4 blocks x 3 passes (hi.hi + lo.hi + hi.lo: ~2^-16 relative per product, unbiased) x 2 k-steps of
// tcgen05.mma.kind::f16 (bf16 inputs, both operands MN-major, no swizzle) into fp32 accumulators in
// TMEM.  Every kODrain tiles the accumulators are drained with tcgen05.ld and added (round-to-nearest)
// into per-thread float32 registers of 8 drain warps, because the tensor core's own fp32 accumulation
// degrades over long chains (measured with the earlier 3xTF32 version: 2-3e-4 relative error in P_bar
// with one chain per CTA, 5e-6 with a drain every 512 columns).  Per-CTA partials are summed in float64
// by a last kernel.
#pragma once
#include "kernels_tc.cuh"

namespace cbf {

constexpr int kODrainers = 256;    // warps 1..8 hold the drained accumulators
constexpr int kOThreads = 32 + kODrainers + 32;   // warp 0: MMA issue, warps 1..8: drains, warp 9: bulk copies
constexpr int kODrain = 16;        // tiles (512 columns) between accumulator drains
constexpr int kOStages = 3;        // shared-memory ring depth (3 x ~72 KB)
constexpr int kOCols = 192;        // TMEM columns used: P_bar 0..127 | alpha 128..143 | S 144..159 | U 160..191

struct OuterArgs {
  TcMats m;
  int M, dout, din;
};

// D = fp32, A = B = bf16, both MN-major (cute::UMMA::InstrDescriptor bits 15/16).
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Shared-memory stage (row-blocks of kOBlk bytes): the four left parts as [hi 16 blocks | lo 16 blocks]
// (a_bar, k', a^2, w; blocks >= MB stay zero), then g_mean [hi 2 | lo 2], g_var [hi 2 | lo 2],
// [x~,1] [hi NXB | lo NXB].
__host__ __device__ inline int outer_stage_blocks(int din) { return 128 + 8 + 2 * (round_up(din + 1, 16) / 8); }

// Stage hand-over through mbarriers only: full[s] (bulk-copy bytes) -> issuer; empty[s] (tcgen05.commit) ->
// producer; acc (commit) -> drain warps; drained (256 arrivals) -> issuer.
// Rpart: [gridDim.x][128][kOCols] float64, fully written by every CTA.
__global__ void __launch_bounds__(kOThreads) tc_outer_kernel(OuterArgs a, double *__restrict__ Rpart) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[kOStages], empty_bar[kOStages], acc_bar, drained_bar;
  __shared__ uint32_t tmem_storage;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MP = round_up(a.M, 16);
  const int NX = round_up(a.din + 1, 16), NXB = NX / 8;
  const uint32_t stage_bytes = (uint32_t)outer_stage_blocks(a.din) * kOBlk;
  {
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    for (uint32_t i = tid; i < kOStages * stage_bytes / 16; i += kOThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < kOStages; ++q) { mbar_init(smem_u32(&full_bar[q]), 1); mbar_init(smem_u32(&empty_bar[q]), 1); }
    mbar_init(smem_u32(&acc_bar), 1);
    mbar_init(smem_u32(&drained_bar), kODrainers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_storage)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  async_proxy_fence();     // the zero fill (generic proxy) is ordered before the bulk copies / MMA reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_storage;

  const size_t L = a.m.L;
  const size_t ntile = (L + kOT - 1) / kOT;
  const size_t per = (ntile + gridDim.x - 1) / gridDim.x;
  const size_t t0 = (size_t)blockIdx.x * per, t1 = (t0 + per < ntile) ? t0 + per : ntile;
  const uint32_t sbase = smem_u32(smem_raw);

  if (warp == 0) {
    // ================= MMA issuer =================
    const uint32_t idP = umma_idesc_bf16_mn(128, MP), id16 = umma_idesc_bf16_mn(128, 16), idX = umma_idesc_bf16_mn(128, NX);
    // The 24 MMAs of one tile (4 blocks x 3 passes x 2 k-steps).  Their operand descriptors for stage 0 stay
    // in registers of the single issuing thread; stage s only adds s * stage_bytes / 16 to the start-address
    // field.  MN-major, no swizzle: leading-dim offset = 128 B between the two 8-column core matrices of a
    // k-step, stride-dim offset = kOBlk between row-blocks.
    uint64_t da[24], db[24];
    {
      const uint32_t right_hi[4] = {sbase + 2 * 16 * kOBlk, sbase + 128 * kOBlk, sbase + 132 * kOBlk, sbase + 136 * kOBlk};
      const uint32_t right_lo[4] = {16u * kOBlk, 2u * kOBlk, 2u * kOBlk, (uint32_t)NXB * kOBlk};
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        const int b = i / 6, pass = (i % 6) / 2, k = i % 2;
        const uint32_t la = sbase + (uint32_t)(2 * b + (pass == 1 ? 1 : 0)) * 16 * kOBlk;   // hi.hi + lo.hi + hi.lo
        const uint32_t ra = right_hi[b] + (pass == 2 ? right_lo[b] : 0u);
        da[i] = umma_desc(la + k * 256, 128, kOBlk);
        db[i] = umma_desc(ra + k * 256, 128, kOBlk);
      }
    }
    const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
    uint32_t fph[kOStages], dph = 0;
#pragma unroll
    for (int q = 0; q < kOStages; ++q) fph[q] = 0;
    int since = 0, st = 0;
    for (size_t tile = t0; tile < t1; ++tile) {
#pragma unroll
      for (int q = 0; q < kOStages; ++q)
        if (q == st) { mbar_wait(smem_u32(&full_bar[q]), fph[q]); fph[q] ^= 1; }
      tc_fence_after();
      if (lane == 0) {
        const uint64_t off = stage_step * (uint64_t)st;
        const uint32_t acc0 = since > 0 ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          const int b = i / 6;
          const uint32_t dcol = (b == 0) ? 0u : (b == 1) ? 128u : (b == 2) ? 144u : 160u;
          const uint32_t idd = (b == 0) ? idP : (b == 3) ? idX : id16;
          umma_f16(tmem + dcol, da[i] + off, db[i] + off, idd, (i % 6 == 0) ? acc0 : 1u);
        }
        umma_commit(smem_u32(&empty_bar[st]));
      }
      __syncwarp();
      if (++since == kODrain || tile + 1 == t1) {
        if (lane == 0) umma_commit(smem_u32(&acc_bar));       // all MMAs so far are complete when this fires
        __syncwarp();
        mbar_wait(smem_u32(&drained_bar), dph);               // the drain warps have read the accumulators
        dph ^= 1;
        tc_fence_after();
        since = 0;
      }
      st = (st + 1 == kOStages) ? 0 : st + 1;
    }
  } else if (warp == 9) {
    // ================= bulk-copy producer =================
    if (lane == 0) {
      const TcMats &m = a.m;
      const uint32_t tile_bytes = (uint32_t)m.tile_bytes();
      const int part[4] = {m.bAb, m.bK, m.bA2, m.bW};
      uint32_t eph[kOStages];
#pragma unroll
      for (int q = 0; q < kOStages; ++q) eph[q] = 0;
      int st = 0;
      size_t done = 0;
      for (size_t tile = t0; tile < t1; ++tile, ++done) {
        if (done >= (size_t)kOStages) {     // the MMAs that last read this stage must be complete
#pragma unroll
          for (int q = 0; q < kOStages; ++q)
            if (q == st) { mbar_wait(smem_u32(&empty_bar[q]), eph[q]); eph[q] ^= 1; }
        }
        const uint32_t bar = smem_u32(&full_bar[st]);
        const uint32_t dst = sbase + (uint32_t)st * stage_bytes;
        const unsigned char *src = m.blk + tile * (size_t)tile_bytes;
        const size_t lo = (size_t)m.RB * kOBlk;
        mbar_expect_tx(bar, tile_bytes);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          bulk_g2s(dst + (2 * p) * 16 * kOBlk, src + (size_t)part[p] * kOBlk, m.MB * kOBlk, bar);
          bulk_g2s(dst + (2 * p + 1) * 16 * kOBlk, src + (size_t)part[p] * kOBlk + lo, m.MB * kOBlk, bar);
        }
        bulk_g2s(dst + 128 * kOBlk, src + (size_t)m.bGm * kOBlk, m.DB * kOBlk, bar);
        bulk_g2s(dst + 130 * kOBlk, src + (size_t)m.bGm * kOBlk + lo, m.DB * kOBlk, bar);
        bulk_g2s(dst + 132 * kOBlk, src + (size_t)m.bGv * kOBlk, m.DB * kOBlk, bar);
        bulk_g2s(dst + 134 * kOBlk, src + (size_t)m.bGv * kOBlk + lo, m.DB * kOBlk, bar);
        bulk_g2s(dst + 136 * kOBlk, src + (size_t)m.bX1 * kOBlk, m.XB * kOBlk, bar);
        bulk_g2s(dst + (136 + NXB) * kOBlk, src + (size_t)m.bX1 * kOBlk + lo, m.XB * kOBlk, bar);
        st = (st + 1 == kOStages) ? 0 : st + 1;
      }
    }
  } else {
    // ================= accumulator drains =================
    // this thread's share: TMEM lane quarter warp % 4 (warps 1..8 give every quarter two warps), columns
    // [c_begin, c_begin + 96)
    constexpr int kHalf = kOCols / 2;
    float racc[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) racc[e] = 0.f;
    const int c_begin = (warp <= 4) ? 0 : kHalf;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c_begin;
    uint32_t aph = 0;
    const size_t mine = t1 > t0 ? t1 - t0 : 0;
    const size_t ndrain = (mine + kODrain - 1) / kODrain;
    for (size_t d = 0; d < ndrain; ++d) {
      mbar_wait(smem_u32(&acc_bar), aph);
      aph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < kHalf; cc += 16) {
        float v[16];
        tmem_ld16(trow + cc, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) racc[cc + e] += v[e];
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&drained_bar));
    }
    double *out = Rpart + ((size_t)blockIdx.x * 128 + (warp & 3) * 32 + lane) * kOCols + c_begin;
#pragma unroll
    for (int e = 0; e < kHalf; ++e) out[e] = (double)racc[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// R[m][c] (float64, [M][Ctot]) (+)= sum over CTAs of the partials, with the TMEM column blocks mapped to
// [P_bar' (M) | alpha_bar' (dout) | S_bar (dout) | U, r (din+1)].
__global__ void outer_reduce_kernel(const double *__restrict__ Rpart, int nparts, int M, int dout, int din,
                                    double *__restrict__ R, int accumulate) {
  const int Ctot = M + 2 * dout + din + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * Ctot) return;
  const int m = i / Ctot, c = i - m * Ctot;
  int src;
  if (c < M) src = c;
  else if (c < M + dout) src = 128 + (c - M);
  else if (c < M + 2 * dout) src = 144 + (c - M - dout);
  else src = 160 + (c - M - 2 * dout);
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += Rpart[((size_t)p * 128 + m) * kOCols + src];
  R[i] = accumulate ? R[i] + s : s;   // later time windows add to the first one's result
}

inline size_t outer_smem_bytes(int din) { return (size_t)kOStages * outer_stage_blocks(din) * kOBlk; }

}  // namespace cbf
