// Parameter-adjoint accumulation of the tensor path as a tcgen05 / TMEM split-K reduction.
//
// The tensor-path reverse kernels leave, per GP, a tile-major float32 operand matrix (TcMats) over the
// L = (live steps x particles) columns: a_bar, k', a^2, w (M rows each), g_mean, g_var (Dout),
// [x~, 1] (Din+1).  The parameter adjoints are the four thin GEMMs over that huge K dimension
//     P_bar' = Ab K^T [M x M],  alpha_bar' = K Gm^T,  S_bar = A2 Gv^T,  [U | r] = W X1^T .
// They are memory-bound (13.5 FLOP/B is above the FP32-SIMT ridge but far below the tensor one),
// so each CTA streams a contiguous range of 16-column blocks once: 256 threads load the block (one
// contiguous ~26 KB region) with coalesced 16-byte loads, split every value x into hi = x truncated to TF32 and lo = x - hi
// (exact), and store both in the canonical K-major no-swizzle UMMA layout; one thread issues
// 4 blocks x 3 passes (hi.hi + lo.hi + hi.lo, the 3xTF32 scheme: ~2^-21 relative) x 2 k-steps of
// tcgen05.mma.kind::tf32 into fp32 accumulators in TMEM; the next tile's global loads overlap the
// MMAs.  Every kODrain tiles the accumulators are drained with tcgen05.ld and added (round-to-nearest)
// into per-thread float32 registers, because the tensor core's own fp32 accumulation degrades over long
// chains: measured against the SIMT path at L = 5e6, one chain per CTA gave 2-3e-4 relative error in
// P_bar, a drain every 32 tiles (192 accumulations) 5e-6 (tools/tc_accuracy.py).  Per-CTA partials are
// summed in float64 by a last kernel.
#pragma once
#include "kernels_tc.cuh"

namespace cbf {

constexpr int kOT = 16;            // columns per tile = 2 MMA k-steps of 8
constexpr int kOProducers = 256;   // producer threads (warps 1..8)
constexpr int kOThreads = kOProducers + 32;   // + the MMA-issuing warp 0
constexpr int kODrain = 32;        // tiles between accumulator drains (see below: longer chains lose accuracy)
constexpr int kOStages = 3;        // shared-memory ring depth (3 x ~72 KB)
constexpr int kOCols = 192;        // TMEM columns used: P_bar 0..127 | alpha 128..143 | S 144..159 | U 160..191

struct OuterArgs {
  TcMats m;
  int M, dout, din;
};

__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp-specialised: warp 0 issues the MMAs, warps 1..8 (256 producer threads) load, split and stage the
// blocks and drain the accumulators.  Stage hand-over through mbarriers only (no CTA-wide barrier in
// the loop): full[s] (256 producer arrivals) -> issuer; empty[s] (tcgen05.commit) -> producers.
// Rpart: [gridDim.x][128][kOCols] float64, fully written by every CTA.
__global__ void __launch_bounds__(kOThreads) tc_outer_kernel(OuterArgs a, double *__restrict__ Rpart) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[kOStages], empty_bar[kOStages], acc_bar, drained_bar;
  __shared__ uint32_t tmem_storage;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ptid = tid - 32;                                  // producer thread index (warps 1..8)
  const int M = a.M, MP = round_up(M, 16);
  const int nx = a.din + 1, NX = round_up(nx, 16);            // rows of the X1 tile (16 or 32)
  // One stage (floats): 8 left tiles [128 x 16] (hi, lo of Ab, K, A2, W), then Gm, Gv (16 rows), X1 (NX rows),
  // hi+lo each.
  const int nfl = 8 * 128 * kOT + 4 * 16 * kOT + 2 * NX * kOT;
  float *stage0 = reinterpret_cast<float *>(smem_raw);
  for (int i = tid; i < kOStages * nfl; i += kOThreads) stage0[i] = 0.f;   // rows >= M / unused rows stay zero
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < kOStages; ++q) { mbar_init(smem_u32(&full_bar[q]), kOProducers); mbar_init(smem_u32(&empty_bar[q]), 1); }
    mbar_init(smem_u32(&acc_bar), 1);
    mbar_init(smem_u32(&drained_bar), kOProducers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_storage)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  async_proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_storage;

  const size_t L = a.m.L;
  const int R = a.m.R;
  const size_t ntile = (L + kOT - 1) / kOT;
  const size_t per = (ntile + gridDim.x - 1) / gridDim.x;
  const size_t t0 = (size_t)blockIdx.x * per, t1 = (t0 + per < ntile) ? t0 + per : ntile;

  if (warp == 0) {
    // ================= MMA issuer =================
    const uint32_t idP = umma_idesc_tf32(128, MP), id16 = umma_idesc_tf32(128, 16), idX = umma_idesc_tf32(128, NX);
    const uint32_t lsz = 128 * kOT * 4;                      // bytes of one left tile
    // The 24 MMAs of one tile (4 blocks x 3 passes x 2 k-steps).  Their operand descriptors for stage 0 stay
    // in registers of the single issuing thread; stage s only adds s * stage_bytes / 16 to the start-address
    // field (shared-memory addresses fit its 14 bits), so issuing one tile is ~6 instructions per MMA.
    uint64_t da[24], db[24];
    {
      const uint32_t lbase = smem_u32(stage0);
      const uint32_t sgm_a = lbase + 8 * lsz, sgv_a = sgm_a + 2 * 16 * kOT * 4, sx1_a = sgv_a + 2 * 16 * kOT * 4;
      // block b: left matrix b (a_bar, k', a^2, w) x right tile (k' | g_mean | g_var | [x~,1]) -> TMEM column block
      const uint32_t rb[4] = {lbase + 2 * lsz, sgm_a, sgv_a, sx1_a};
      const uint32_t rrows[4] = {128u, 16u, 16u, (uint32_t)NX};
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        const int b = i / 6, pass = (i % 6) / 2, k = i % 2;
        const uint32_t lh = lbase + (2 * b) * lsz, ll = lh + lsz;
        const uint32_t rh = rb[b], rl = rh + rrows[b] * kOT * 4;
        const uint32_t lboL = 128 * 16, lboR = rrows[b] * 16;
        const uint32_t la = (pass == 1) ? ll : lh, ra = (pass == 2) ? rl : rh;   // hi.hi + lo.hi + hi.lo
        da[i] = umma_desc(la + k * 2 * lboL, lboL, 128);
        db[i] = umma_desc(ra + k * 2 * lboR, lboR, 128);
      }
    }
    const uint64_t stage_step = (uint64_t)((nfl * 4) >> 4);
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
          umma_tf32(tmem + dcol, da[i] + off, db[i] + off, idd, (i % 6 == 0) ? acc0 : 1u);
        }
        umma_commit(smem_u32(&empty_bar[st]));
      }
      __syncwarp();
      if (++since == kODrain || tile + 1 == t1) {
        if (lane == 0) umma_commit(smem_u32(&acc_bar));       // all MMAs so far are complete when this fires
        __syncwarp();
        mbar_wait(smem_u32(&drained_bar), dph);               // producers have read the accumulators
        dph ^= 1;
        tc_fence_after();
        since = 0;
      }
      st = (st + 1 == kOStages) ? 0 : st + 1;
    }
  } else {
    // ================= producers =================
    // one block = 4 chunks x R rows of 16 bytes, contiguous in memory: float4 index = chunk * R + row.
    // Slot q of a thread: chunk = q & 3, row = ptid + (q >> 2) * 256  (R <= 3 * 256).
    constexpr int kSlots = 12;
    float4 regsA[kSlots];
    auto fetch = [&](size_t tile, float4 (&regs)[kSlots]) {
      const float4 *src = reinterpret_cast<const float4 *>(a.m.blk + tile * ((size_t)R * 16));
      const size_t col0 = tile * kOT;
      const bool fullblk = col0 + kOT <= L;
#pragma unroll
      for (int q = 0; q < kSlots; ++q) {
        const int ch = q & 3, row = ptid + (q >> 2) * kOProducers;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < R) {
          v = src[ch * R + row];
          if (!fullblk) {                 // columns >= L of the last block were never written
            const size_t c = col0 + 4 * ch;
            if (c >= L) v.x = 0.f;
            if (c + 1 >= L) v.y = 0.f;
            if (c + 2 >= L) v.z = 0.f;
            if (c + 3 >= L) v.w = 0.f;
          }
        }
        regs[q] = v;
      }
    };
    auto stash = [&](float *left, const float4 (&regs)[kSlots]) {
      float *sgm = left + 8 * 128 * kOT, *sgv = sgm + 2 * 16 * kOT, *sx1 = sgv + 2 * 16 * kOT;
#pragma unroll
      for (int rr = 0; rr < kSlots / 4; ++rr) {
        const int grow = ptid + rr * kOProducers;
        if (grow < R) {
          // destination tile of this row: a_bar | k' | a^2 | w -> left tiles 0..3, then the small right tiles
          float *th;
          int row, rows;
          if (grow < M) { th = left; row = grow; rows = 128; }
          else if (grow < 2 * M) { th = left + 2 * 128 * kOT; row = grow - M; rows = 128; }
          else if (grow < 3 * M) { th = left + 4 * 128 * kOT; row = grow - 2 * M; rows = 128; }
          else if (grow < 4 * M) { th = left + 6 * 128 * kOT; row = grow - 3 * M; rows = 128; }
          else if (grow < 4 * M + a.dout) { th = sgm; row = grow - 4 * M; rows = 16; }
          else if (grow < 4 * M + 2 * a.dout) { th = sgv; row = grow - 4 * M - a.dout; rows = 16; }
          else { th = sx1; row = grow - 4 * M - 2 * a.dout; rows = NX; }
          float *tl = th + rows * kOT;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 v = regs[rr * 4 + ch];
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            // canonical K-major layout: 16-byte chunk ch of row `row` at ch*(rows*16 B) + row*16 B;
            // consecutive lanes = consecutive rows -> conflict-free 128-bit stores
            *reinterpret_cast<float4 *>(th + (size_t)ch * rows * 4 + row * 4) = hi;
            *reinterpret_cast<float4 *>(tl + (size_t)ch * rows * 4 + row * 4) = lo;
          }
        }
      }
    };
    // this thread's share of the drained accumulators: TMEM lane quarter warp % 4 (warps 1..8 give every
    // quarter two warps), columns [c_begin, c_begin + 96)
    constexpr int kHalf = kOCols / 2;
    float racc[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) racc[e] = 0.f;
    const int c_begin = (warp <= 4) ? 0 : kHalf;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c_begin;

    uint32_t eph[kOStages], aph = 0;
#pragma unroll
    for (int q = 0; q < kOStages; ++q) eph[q] = 0;
    int since = 0, st = 0;
    size_t done = 0;
    auto produce = [&](size_t tile, float4 (&regs)[kSlots]) {
      if (done >= (size_t)kOStages) {     // the MMAs that last read this stage must be complete
#pragma unroll
        for (int q = 0; q < kOStages; ++q)
          if (q == st) { mbar_wait(smem_u32(&empty_bar[q]), eph[q]); eph[q] ^= 1; }
      }
      stash(stage0 + (size_t)st * nfl, regs);
      async_proxy_fence();
#pragma unroll
      for (int q = 0; q < kOStages; ++q)
        if (q == st) mbar_arrive(smem_u32(&full_bar[q]));
      if (tile + 1 < t1) fetch(tile + 1, regs);
      if (++since == kODrain || tile + 1 == t1) {
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
        since = 0;
      }
      st = (st + 1 == kOStages) ? 0 : st + 1;
      ++done;
    };
    if (t0 < t1) fetch(t0, regsA);
    for (size_t tile = t0; tile < t1; ++tile) produce(tile, regsA);
    double *out = Rpart + ((size_t)blockIdx.x * 128 + (warp & 3) * 32 + lane) * kOCols + c_begin;
#pragma unroll
    for (int e = 0; e < kHalf; ++e) out[e] = (double)racc[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// R[m][c] (float64, [M][Ctot]) = sum over CTAs of the partials, with the TMEM column blocks mapped to
// [P_bar' (M) | alpha_bar' (dout) | S_bar (dout) | U, r (din+1)].
__global__ void outer_reduce_kernel(const double *__restrict__ Rpart, int nparts, int M, int dout, int din,
                                    double *__restrict__ R) {
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
  R[i] = s;
}

inline size_t outer_smem_bytes(int din) {
  const int NX = round_up(din + 1, 16);
  return sizeof(float) * kOStages * (8 * 128 * kOT + 4 * 16 * kOT + 2 * NX * kOT);
}

}  // namespace cbf
