// C ABI of cbfssm_b200 (see include/cbfssm_b200.h): shape checks, workspace layout,
// chain schedule, kernel dispatch and the small float64 reduction kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "dims_list.h"
#include "f64_path.h"
#include "kernels_outer.cuh"

namespace cbf {

static thread_local long long g_launches = 0;
void cbf_note_launch() { ++g_launches; }

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define CBF_DECLARE_OPS(DX, DU, DY) const DimOps *ops_##DX##_##DU##_##DY();
CBF_DIMS_LIST(CBF_DECLARE_OPS)
#undef CBF_DECLARE_OPS

#define CBF_DECLARE_FAST(DX, DU, DY, M) const DimOps *fast_ops_##DX##_##DU##_##DY##_##M();
CBF_FAST_LIST(CBF_DECLARE_FAST)
#undef CBF_DECLARE_FAST

const DimOps *find_ops(int dx, int du, int dy, int M, bool allow_fast) {
  if (allow_fast) {
#define CBF_MATCH_FAST(DX, DU, DY, MM) \
  if (dx == DX && du == DU && dy == DY && M == MM) return fast_ops_##DX##_##DU##_##DY##_##MM();
    CBF_FAST_LIST(CBF_MATCH_FAST)
#undef CBF_MATCH_FAST
  }
#define CBF_MATCH_OPS(DX, DU, DY) \
  if (dx == DX && du == DU && dy == DY) return ops_##DX##_##DU##_##DY();
  CBF_DIMS_LIST(CBF_MATCH_OPS)
#undef CBF_MATCH_OPS
  return nullptr;
}

// ---- optional per-kernel device timing (bench.py roofline): thread-local, off by default ----
struct TimingPool {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;      // pairs (start, stop)
  std::vector<int> kind;            // 0 bm_forward 1 fw_forward 2 fw_reverse 3 bm_reverse 4/5 outer-product accumulation f/b
  size_t used = 0;
};
static thread_local TimingPool g_timing;

struct ScopedTiming {
  cudaStream_t st;
  bool on;
  size_t idx;
  ScopedTiming(int kind, cudaStream_t s) : st(s), on(g_timing.enabled), idx(0) {
    if (!on) return;
    TimingPool &t = g_timing;
    if (t.used * 2 >= t.ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
      t.ev.push_back(a); t.ev.push_back(b); t.kind.push_back(kind);
    }
    idx = t.used++;
    t.kind[idx] = kind;
    cudaEventRecord(t.ev[2 * idx], st);
  }
  ~ScopedTiming() { if (on) cudaEventRecord(g_timing.ev[2 * idx + 1], st); }
};

constexpr int kMaxGridRev = 148 * 8;   // upper bound on persistent reverse CTAs (workspace sizing)
constexpr size_t kMaxSmem = 227 * 1024;
// Particle counts from which the specialised paths are used.  With the final kernels both beat the cooperative
// path at every size measured (M=20: 3.0 vs 3.2-3.9 ms/step at 50-3 200 particles; M=100: 5.2 vs 7.7 ms at
// 50-3 200 particles, tools/bench_cross.sh), so the cooperative kernels serve only M without a specialised path.
constexpr int kMinParticlesRegisterPath = 1;
constexpr int kMinParticlesTensorPath = 1;
#ifndef CBF_MIN_TENSOR_M
#define CBF_MIN_TENSOR_M 16
#endif
constexpr int kMinTensorM = CBF_MIN_TENSOR_M;   // smallest M served by the tensor path

// Live chain segments of both backward-message runs (cbfssm.py:123-136, SURVEY 8a note 5).
static std::vector<Chain> build_chains(int T, int R) {
  std::vector<Chain> out;
  for (int run = 0; run < 2; ++run) {
    const int off = run == 0 ? 1 : R + 1;
    std::vector<int> starts;
    starts.push_back(T - 1);
    for (int t = T - 2; t >= 0; --t)
      if ((t + off) % (2 * R) == 0) starts.push_back(t);
    for (size_t i = 0; i < starts.size(); ++i) {
      const int t_hi = starts[i];
      const int t_next = (i + 1 < starts.size()) ? starts[i + 1] : -1;
      const bool resample = (t_hi + off) % (2 * R) == 0;
      int t_lo = -1;
      for (int t = t_hi; t > t_next; --t)
        if (writer_run(t, R) == run) t_lo = t;
      if (t_lo < 0) continue;   // segment writes nothing: dead work
      out.push_back(Chain{run, t_hi, t_lo, resample ? 1 : 0, 0, t_hi, 0, 0});
    }
  }
  int col = 0;
  for (Chain &c : out) { c.col0 = col; col += c.t_hi - c.t_lo + 1; }
  for (size_t i = 0; i < out.size(); ++i) out[i].id = (int)i;
  return out;
}

struct Plan {
  Dims D;
  int dx, du, dy, dh, din;
  int ptiles, slots_per_cta;
  std::vector<Chain> chains;
  AccLayout Lf, Lb;
  size_t off_X, off_H, off_Yb, off_fbm, off_ffw, off_gf, off_gb, off_accf, off_accb, off_stats, off_cpack, total;
  const DimOps *ops;
  bool half;                     // CBFSSMHALF: no backward-message GP, x_0 supplied by the caller
  size_t off_x0b;
  size_t off_kaf, off_kab;       // register path: saved evaluations (0 = not saved)
  bool save_eval;
  size_t off_fvf, off_fvb;       // tensor path: saved (fmean, fvar, amax) per evaluation
  bool save_fv;
  bool f64;                      // float64 batched path (f64_path.cu): M > 128, dims without an instantiation, CBF_FLAG_FP64
  size_t off_f64;
  // tensor-core path (16 <= M <= 128, enough particles)
  bool tc_fwd, tc_rev;
  size_t colsf, colsb;           // columns (live steps x particles) of the operand matrices
  size_t off_rpart;              // per-CTA float64 partials of the tcgen05 outer-product kernel
  int ctot_f, ctot_b, nsc_f, nsc_b, nspart_f, nspart_b;
  size_t off_mats, off_spf, off_spb, off_rdf, off_rdb, off_carry_f, off_carry_b;
  // time windows of the tensor-path reverse pass (the operand tiles of one window fit the window budget)
  std::vector<TimeWin> win_f;                 // forward-rollout reverse: descending time
  std::vector<std::vector<Chain>> rounds_b;   // message reverse: r-th piece of every chain, ascending time
};

// Budget for the outer-product operand tiles of one time window of the tensor-path reverse pass (one buffer,
// used by the forward-rollout GP's windows and then by the message GP's).  Calls whose tiles fit run as one
// window; CBFSSM_B200_TC_WINDOW_BYTES overrides it (the tests use it to force many windows).
constexpr size_t kTcWindowBytes = (size_t)40 << 30;
static size_t tc_window_budget() {
  const char *e = getenv("CBFSSM_B200_TC_WINDOW_BYTES");
  if (e != nullptr && *e) {
    const long long v = atoll(e);
    if (v > 0) return (size_t)v;
  }
  return kTcWindowBytes;
}
// Budget for the register path's saved evaluations (176-192 B per GP evaluation at M = 20: 37 GB at the bench
// shape of 227 200 particles x 300 steps); larger problems recompute in the reverse pass as before.
// CBFSSM_B200_SAVE_EVAL_BYTES overrides it (0 disables saving).
constexpr size_t kSavedEvalBytes = (size_t)64 << 30;
static size_t saved_eval_budget() {
  const char *e = getenv("CBFSSM_B200_SAVE_EVAL_BYTES");
  if (e != nullptr && *e) {
    const long long v = atoll(e);
    if (v >= 0) return (size_t)v;
  }
  return kSavedEvalBytes;
}
constexpr int kOuterMaxGrid = 148 * 2;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bytes of one GP's outer-product operand tiles (common.cuh TcMats) for `cols` columns.
static size_t mats_bytes(size_t cols, int M, int dout, int din) {
  const size_t rb = (size_t)4 * ceil_div(M, 8) + 2 * ceil_div(dout, 8) + ceil_div(din + 1, 8);
  return ((cols + kOT - 1) / kOT) * 2 * rb * kOBlk;
}

static int make_plan(const cbf_shape *s, Plan &p, bool need_ops) {
  if (!s) { set_error("shape is NULL"); return CBF_ERR_NULL; }
  if (s->B < 1 || s->S < 1 || s->T < 1 || s->M < 1 || s->dx < 2 || s->du < 0 || s->dy < 1 || s->dy >= s->dx ||
      s->R < 1 || s->n_local < 1 || s->n_offset < 0 || (int64_t)s->n_offset + s->n_local > (int64_t)s->B * s->S) {
    set_error("invalid shape B=%d S=%d T=%d M=%d dx=%d du=%d dy=%d R=%d n_offset=%d n_local=%d", s->B, s->S, s->T,
              s->M, s->dx, s->du, s->dy, s->R, s->n_offset, s->n_local);
    return CBF_ERR_INVALID_SHAPE;
  }
  p.dx = s->dx; p.du = s->du; p.dy = s->dy; p.dh = s->dx - s->dy; p.din = s->dx + s->du;
  // Kernel selection (measured, tools/bench_small.sh): the register-resident kernels need enough
  // particles to fill the SMs with one particle per thread; below that the cooperative kernels,
  // which split one particle's M-loop over several warps, have the shorter serial chain per step.
  const bool want_fast = !(s->flags & CBF_FLAG_FORCE_COOPERATIVE) &&
                         ((s->flags & CBF_FLAG_FORCE_REGISTER) || s->n_local >= kMinParticlesRegisterPath);
  p.ops = find_ops(s->dx, s->du, s->dy, s->M, want_fast);
  // float64 batched path: asked for, or the only one that can take the shape (M beyond the tensor path's 128, or
  // dims without a compiled instantiation); it needs no per-dims kernels
  p.f64 = (s->flags & CBF_FLAG_FP64) != 0 || s->M > 128 || !p.ops;
  if (p.f64 && (s->dx + s->du + 1 > 32 || s->dx > 16)) {
    set_error("float64 path: dims (dx=%d,du=%d,dy=%d) exceed dx <= 16, dx + du <= 31", s->dx, s->du, s->dy);
    return CBF_ERR_UNSUPPORTED_DIMS;
  }
  if (p.f64) p.ops = nullptr;
  if (need_ops && !p.ops && !p.f64) {
    set_error("dims (dx=%d,du=%d,dy=%d) are not compiled in (see csrc/dims_list.h)", s->dx, s->du, s->dy);
    return CBF_ERR_UNSUPPORTED_DIMS;
  }
  // The cooperative kernels keep P, the [m][n] vectors and the accumulators in shared memory (M <= ~110); above
  // that only the tensor path can run the shape (checked once it has been selected below).
  bool coop_fits = true;
  size_t coop_need = 0;
  if (p.ops) {
    for (int w = 0; w < 4; ++w)
      if (p.ops->smem_bytes(s->M, w) > kMaxSmem) { coop_fits = false; coop_need = p.ops->smem_bytes(s->M, w); }
  }
  Dims &D = p.D;
  D.B = s->B; D.S = s->S; D.T = s->T; D.M = s->M; D.R = s->R; D.condition = s->condition ? 1 : 0;
  D.n_offset = s->n_offset; D.n_local = s->n_local; D.kap = s->k_factor;
  D.npad = round_up(s->n_local, 32);
  p.half = (s->flags & CBF_FLAG_HALF_MODEL) != 0;
  D.half = p.half ? 1 : 0;
  D.ncond = p.half ? s->dy : s->dx;
  p.ptiles = ceil_div(s->n_local, kNP);            // upper bound for every path (smallest CTA tile)
  p.chains = build_chains(s->T, s->R);
  if (p.half) p.chains.clear();
  if ((s->flags & CBF_FLAG_PREDICT_ONLY) && !s->condition) {
    // Free-running prediction (cbfssm.py:227-228): the rollout is conditioned only while t < R - 1, so of the
    // backward message only y2[0 .. R-1] is ever read -- keep the chains that write one of those steps (one
    // chain of <= 2R steps instead of ~2T)
    std::vector<Chain> need;
    for (const Chain &c : p.chains) {
      bool used = false;
      for (int t = c.t_lo; t <= c.t_hi && t <= s->R - 1; ++t) used = used || writer_run(t, s->R) == c.run;
      if (used) need.push_back(c);
    }
    p.chains.swap(need);
  }
  p.slots_per_cta = 1;
  if (p.ops) {
    p.ops->layouts(s->M, &p.Lf, &p.Lb);
    p.slots_per_cta = p.ops->slots_per_cta;
    p.ptiles = ceil_div(s->n_local, p.ops->particles_per_cta);
  } else {
    p.Lf = AccLayout(s->M, p.din, p.dx, p.dx);
    p.Lb = AccLayout(s->M, p.din, p.dh, p.dx);
  }
  size_t o = 0;
  const size_t np = D.npad;
  p.off_X = o; o = align_up(o + sizeof(float) * s->T * p.dx * np, 256);
  p.off_H = o; o = align_up(o + sizeof(float) * 2 * s->T * p.dh * np, 256);
  p.off_Yb = o; o = align_up(o + sizeof(float) * s->T * p.dh * np, 256);
  p.off_fbm = o; o = align_up(o + sizeof(float) * p.chains.size() * p.ptiles, 256);
  p.off_ffw = o; o = align_up(o + sizeof(float) * p.ptiles * (p.dy + 1), 256);
  p.off_gf = o; o = align_up(o + sizeof(float) * (size_t)kMaxGridRev * p.slots_per_cta * p.Lf.slot(), 256);
  p.off_gb = o; o = align_up(o + sizeof(float) * (size_t)kMaxGridRev * p.slots_per_cta * p.Lb.slot(), 256);
  p.off_accf = o; o = align_up(o + sizeof(double) * p.Lf.slot(), 256);
  p.off_accb = o; o = align_up(o + sizeof(double) * p.Lb.slot(), 256);
  p.off_stats = o; o = align_up(o + sizeof(double) * (p.dy + 2), 256);
  p.off_cpack = o; o = align_up(o + sizeof(float) * 2 * 2048, 256);
  p.off_x0b = o; o = align_up(o + sizeof(float) * p.dx * np, 256);
  // ---- register path: keep every GP evaluation (k, a, fmean, fvar) for the reverse pass when it fits ----
  p.save_eval = false; p.off_kaf = p.off_kab = 0;
  if (p.ops && p.ops->fixed_M && p.ops->saved_planes && !(s->flags & CBF_FLAG_PREDICT_ONLY) && s->T > 1) {
    const size_t bf = sizeof(float4) * (size_t)p.ops->saved_planes(0) * np * (size_t)(s->T - 1);
    const size_t bb = p.half ? 0 : sizeof(float4) * (size_t)p.ops->saved_planes(1) * np * (size_t)(2 * s->T);
    if (bf + bb <= saved_eval_budget()) {
      p.save_eval = true;
      p.off_kaf = o; o = align_up(o + bf, 256);
      p.off_kab = o; o = align_up(o + bb + 16, 256);
    }
  }
  // ---- tensor-core path ----
  p.tc_fwd = p.tc_rev = false;
  p.save_fv = false; p.off_fvf = p.off_fvb = 0;
  if (p.ops && p.ops->fw_forward_tc != nullptr && s->M >= kMinTensorM && s->M <= 128 &&
      !(s->flags & (CBF_FLAG_FORCE_COOPERATIVE | CBF_FLAG_NO_TENSOR_CORES)) &&
      ((s->flags & CBF_FLAG_FORCE_TENSOR_CORES) || s->n_local >= kMinParticlesTensorPath)) {
    p.tc_fwd = p.ops->smem_tc(s->M, 0) <= kMaxSmem && p.ops->smem_tc(s->M, 1) <= kMaxSmem;
    int live = 0;
    for (const Chain &c : p.chains) live += c.t_hi - c.t_lo + 1;
    p.colsf = (size_t)(s->T > 1 ? s->T - 1 : 0) * s->n_local;
    p.colsb = (size_t)live * s->n_local;
    p.ctot_f = s->M + 2 * p.dx + p.din + 1;
    p.ctot_b = s->M + 2 * p.dh + p.din + 1;
    p.tc_rev = p.tc_fwd && p.ops->fw_reverse_tc != nullptr && p.ops->smem_tc(s->M, 2) <= kMaxSmem &&
               p.ops->smem_tc(s->M, 3) <= kMaxSmem && p.colsf > 0 && (p.colsb > 0 || p.half);
    if (p.tc_rev) {
      // ---- time windows ----
      const size_t budget = tc_window_budget();
      const size_t col_f = mats_bytes(kOT, s->M, p.dx, p.din) / kOT, col_b = mats_bytes(kOT, s->M, p.dh, p.din) / kOT;
      const int steps_f = s->T - 1;
      size_t wmax = budget / (col_f * (size_t)s->n_local);
      if (wmax < 1) wmax = 1;
      if (wmax * (size_t)s->n_local >= ((size_t)1 << 31)) wmax = (((size_t)1 << 31) - 1) / s->n_local;
      const int nwin_f = (int)((steps_f + wmax - 1) / wmax), wf = ceil_div(steps_f, nwin_f);
      p.win_f.clear();
      for (int hi = steps_f - 1; hi >= 0; hi -= wf) {
        const int lo = hi - wf + 1 > 0 ? hi - wf + 1 : 0;
        p.win_f.push_back(TimeWin{hi, lo, hi == steps_f - 1 ? 1 : 0, lo == 0 ? 1 : 0});
      }
      size_t bytes_win = mats_bytes((size_t)wf * s->n_local, s->M, p.dx, p.din);
      p.rounds_b.clear();
      if (!p.half) {
        // Pack whole chains into launches of at most `cap` columns (chain-steps x particles within the budget);
        // a chain is cut only where it does not fit, and its pieces go to successive launches in ascending
        // time so that the message adjoint can cross them through the carry buffer.
        size_t cap = budget / (col_b * (size_t)s->n_local);
        if (cap < 1) cap = 1;
        if (cap * (size_t)s->n_local >= ((size_t)1 << 31)) cap = (((size_t)1 << 31) - 1) / s->n_local;
        if (cap < 1) { set_error("too many particles for one time step of operand tiles"); return CBF_ERR_INVALID_SHAPE; }
        std::vector<Chain> cur;
        size_t cur_cols = 0;
        for (const Chain &c : p.chains) {
          int lo = c.t_lo;
          bool first = true, in_cur = false;
          while (lo <= c.t_hi) {
            if (cur_cols >= cap || in_cur || (int)cur.size() >= kMaxChains) {   // launch full, or it holds this chain's previous piece
              p.rounds_b.push_back(cur);
              cur.clear(); cur_cols = 0; in_cur = false;
            }
            const size_t space = cap - cur_cols, left = (size_t)(c.t_hi - lo + 1);
            const int take = (int)(left < space ? left : space);
            Chain piece = c;
            piece.t_lo = lo;
            piece.t_hi = lo + take - 1;
            piece.carry = (first ? 0 : 1) | (piece.t_hi < c.t_hi ? 2 : 0);
            cur.push_back(piece);
            cur_cols += (size_t)take;
            in_cur = piece.t_hi < c.t_hi;   // the next piece of this chain must wait for this launch
            lo += take;
            first = false;
          }
        }
        if (!cur.empty()) p.rounds_b.push_back(cur);
        for (auto &round : p.rounds_b) {
          int col = 0;
          for (Chain &c : round) { c.col0 = col; col += c.t_hi - c.t_lo + 1; }
          const size_t b = mats_bytes((size_t)col * s->n_local, s->M, p.dh, p.din);
          if (b > bytes_win) bytes_win = b;
        }
      }
      p.nsc_f = p.Lf.slot() - p.Lf.scal_off();
      p.nsc_b = p.Lb.slot() - p.Lb.scal_off();
      const int pt = ceil_div(s->n_local, 128);
      p.nspart_f = pt;
      p.nspart_b = pt * (int)p.chains.size();
      p.off_mats = o; o = align_up(o + bytes_win, 256);
      p.off_rpart = o; o = align_up(o + sizeof(double) * (size_t)kOuterMaxGrid * 128 * kOCols, 256);
      p.off_spf = o; o = align_up(o + sizeof(float) * (size_t)p.nspart_f * p.nsc_f, 256);
      p.off_spb = o; o = align_up(o + sizeof(float) * (size_t)p.nspart_b * p.nsc_b, 256);
      p.off_rdf = o; o = align_up(o + sizeof(double) * (size_t)s->M * p.ctot_f, 256);
      p.off_rdb = o; o = align_up(o + sizeof(double) * (size_t)s->M * p.ctot_b, 256);
      p.off_carry_f = o; o = align_up(o + sizeof(float) * p.dx * np, 256);
      p.off_carry_b = o; o = align_up(o + sizeof(float) * (p.chains.size() + 1) * p.dh * np, 256);
      // moments of every evaluation, left by the forward kernels so that the reverse kernels recompute only the
      // kernel vector and a = P k (36 / 20 B per evaluation at D = 4)
      const size_t fvf = sizeof(float) * (size_t)(2 * p.dx + 1) * np * (size_t)(s->T - 1);
      const size_t fvb = p.half ? 0 : sizeof(float) * (size_t)(2 * p.dh + 1) * np * (size_t)(2 * s->T);
      if (!(s->flags & CBF_FLAG_PREDICT_ONLY) && fvf + fvb <= saved_eval_budget()) {
        p.save_fv = true;
        p.off_fvf = o; o = align_up(o + fvf, 256);
        p.off_fvb = o; o = align_up(o + fvb + 16, 256);
      }
    }
  }
  if (p.f64) {
    p.off_f64 = o;
    o = align_up(o + f64_scratch_bytes(s->n_local, s->T, s->M, p.dx, p.dy, p.din), 256);
  }
  if (!p.f64 && !coop_fits && !(p.tc_fwd && p.tc_rev)) {
    set_error("M=%d: resident parameter set needs %zu B of shared memory (> %zu) and the tensor path (16 <= M <= 128) "
              "is not available for this call", s->M, coop_need, kMaxSmem);
    return CBF_ERR_UNSUPPORTED_M;
  }
  p.total = o;
  return 0;
}

static int device_sms();

static TcMats bind_mats(void *base, size_t off, size_t L, int M, int dout, int din) {
  TcMats m;
  m.blk = reinterpret_cast<unsigned char *>(base) + off;
  m.L = L;
  m.MB = ceil_div(M, 8); m.DB = ceil_div(dout, 8); m.XB = ceil_div(din + 1, 8);
  m.RB = 4 * m.MB + 2 * m.DB + m.XB;
  m.bAb = 0; m.bK = m.MB; m.bA2 = 2 * m.MB; m.bW = 3 * m.MB;
  m.bGm = 4 * m.MB; m.bGv = m.bGm + m.DB; m.bX1 = m.bGv + m.DB;
  return m;
}

// The reverse kernels write only the L valid columns; the bulk copies of the accumulation kernel read whole
// tiles, so the unwritten tail of the last tile is cleared first.
static cudaError_t clear_mats_tail(const TcMats &m, cudaStream_t st) {
  if (m.L == 0 || m.L % kOT == 0) return cudaSuccess;
  return cudaMemsetAsync(m.blk + (m.L / kOT) * m.tile_bytes(), 0, m.tile_bytes(), st);
}

// P_bar', alpha_bar', S_bar, [U|r] of one GP on the tensor cores (kernels_outer.cuh) -> Rd [M x Ctot] float64.
static cudaError_t tc_outer(const TcMats &m, int M, int dout, int din, double *rpart, double *Rd, bool accumulate,
                            cudaStream_t st) {
  OuterArgs a{m, M, dout, din};
  const size_t ntile = (m.L + kOT - 1) / kOT;
  int grid = device_sms();          // one CTA per SM: the 3-stage ring takes ~216 KB of shared memory
  if (grid > kOuterMaxGrid) grid = kOuterMaxGrid;
  if ((size_t)grid > ntile) grid = (int)ntile;
  const size_t smem = outer_smem_bytes(din);
  cudaError_t e = cudaFuncSetAttribute(tc_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_outer_kernel<<<grid, kOThreads, smem, st>>>(a, rpart); cbf_note_launch();
  const int Ctot = M + 2 * dout + din + 1;
  outer_reduce_kernel<<<ceil_div(M * Ctot, 256), 256, 0, st>>>(rpart, grid, M, dout, din, Rd, accumulate ? 1 : 0); cbf_note_launch();
  return cudaGetLastError();
}

static Workspace bind_workspace(const Plan &p, void *base) {
  char *b = static_cast<char *>(base);
  Workspace w;
  w.X = reinterpret_cast<float *>(b + p.off_X);
  w.H = reinterpret_cast<float *>(b + p.off_H);
  w.Yb = reinterpret_cast<float *>(b + p.off_Yb);
  w.fpart_bm = reinterpret_cast<float *>(b + p.off_fbm);
  w.fpart_fw = reinterpret_cast<float *>(b + p.off_ffw);
  w.gpart_f = reinterpret_cast<float *>(b + p.off_gf);
  w.gpart_b = reinterpret_cast<float *>(b + p.off_gb);
  w.acc_f = reinterpret_cast<double *>(b + p.off_accf);
  w.acc_b = reinterpret_cast<double *>(b + p.off_accb);
  w.stats = reinterpret_cast<double *>(b + p.off_stats);
  w.cpack = reinterpret_cast<float *>(b + p.off_cpack);
  w.carry_f = p.tc_rev ? reinterpret_cast<float *>(b + p.off_carry_f) : nullptr;
  w.carry_b = p.tc_rev ? reinterpret_cast<float *>(b + p.off_carry_b) : nullptr;
  w.FVf = p.save_fv ? reinterpret_cast<float *>(b + p.off_fvf) : nullptr;
  w.FVb = (p.save_fv && !p.half) ? reinterpret_cast<float *>(b + p.off_fvb) : nullptr;
  w.KAf = p.save_eval ? reinterpret_cast<float4 *>(b + p.off_kaf) : nullptr;
  w.KAb = (p.save_eval && !p.half) ? reinterpret_cast<float4 *>(b + p.off_kab) : nullptr;
  w.x0 = nullptr;
  w.x0b = reinterpret_cast<float *>(b + p.off_x0b);
  w.npad = p.D.npad;
  return w;
}

static cbf_grad_layout grad_layout(const Plan &p) {
  cbf_grad_layout g;
  const int64_t M = p.D.M;
  int64_t o = 0;
  g.f_P = o; o += M * M;
  g.f_alpha = o; o += M * p.dx;
  g.f_S = o; o += M * p.dx;
  g.f_Z = o; o += M * p.din;
  g.f_ell = o; o += p.din;
  g.f_sig2 = o; o += 1;
  g.b_P = o; o += M * M;
  g.b_alpha = o; o += M * p.dh;
  g.b_S = o; o += M * p.dh;
  g.b_Z = o; o += M * p.din;
  g.b_ell = o; o += p.din;
  g.b_sig2 = o; o += 1;
  g.var_x = o; o += p.dx;
  g.var_y = o; o += p.dx;
  g.total = o;
  return g;
}

static bool misaligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

// ------------------------------------------------------------------------------------
// small float64 kernels
// ------------------------------------------------------------------------------------
// terms[0] = loglik (cbfssm.py:245-251), terms[1] = kl_x (:183), terms[2] = entropy (:99)
__global__ void finalize_terms_kernel(const float *__restrict__ fbm, int nbm, const float *__restrict__ ffw,
                                      int ptiles, int dy, const float *__restrict__ vy, double n_times_t,
                                      double *__restrict__ stats, double *__restrict__ terms) {
  __shared__ double sh[256];
  const int tid = threadIdx.x;
  for (int q = 0; q < dy + 2; ++q) {
    double s = 0.0;
    if (q <= dy) {
      for (int i = tid; i < ptiles; i += blockDim.x) s += (double)ffw[(size_t)i * (dy + 1) + q];
    } else {
      for (int i = tid; i < nbm; i += blockDim.x) s += (double)fbm[i];
    }
    sh[tid] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (tid < o) sh[tid] += sh[tid + o];
      __syncthreads();
    }
    if (tid == 0) stats[q] = sh[0];
    __syncthreads();
  }
  if (tid == 0) {
    double ll = 0.0;
    for (int j = 0; j < dy; ++j) {
      const double v = (double)vy[j];
      ll += -0.5 * stats[j] / v - 0.5 * n_times_t * (log(v) + 1.8378770664093454836);
    }
    terms[0] = ll;
    terms[1] = stats[dy];
    terms[2] = stats[dy + 1];
  }
}

__global__ void reduce_slots_kernel(const float *__restrict__ part, int nslots, int slot, double *__restrict__ out,
                                    int accumulate = 0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= slot) return;
  double s = 0.0;
  for (int i = 0; i < nslots; ++i) s += (double)part[(size_t)i * slot + e];
  out[e] = accumulate ? out[e] + s : s;
}

// Decode one GP's reduced accumulators into the flat kernel-level gradient.
__global__ void finalize_gp_grad_kernel(AccLayout L, const double *__restrict__ acc, GpDev gp,
                                        double *__restrict__ gP, double *__restrict__ galpha,
                                        double *__restrict__ gS, double *__restrict__ gZ,
                                        double *__restrict__ gell, double *__restrict__ gsig2) {
  const int M = L.M, Din = L.Din, Dout = L.Dout;
  auto at = [&](int m, int col) -> double { return acc[L.index(m, col)]; };
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < M * M; i += nt) gP[i] = at(i / M, i % M);
  for (int i = tid; i < M * Dout; i += nt) {
    const int m = i / Dout, d = i % Dout;
    galpha[i] = at(m, L.colGm + d);
    gS[i] = at(m, L.colGv + d);
  }
  const int xc = L.colX;
  for (int i = tid; i < M * Din; i += nt) {
    const int m = i / Din, j = i % Din;
    const double ell = (double)gp.ell[j];
    const double zt = (double)gp.Z[i] / ell;
    gZ[i] = (at(m, xc + j) - zt * at(m, xc + Din)) / ell;
  }
  const double *sc = acc + L.scal_off();
  for (int j = tid; j < Din; j += nt) gell[j] = sc[j] / (double)gp.ell[j];
  if (tid == 0) gsig2[0] = sc[Din] / (double)gp.sig2[0] + sc[Din + 1];
}

// Tensor path: R[m][c] (float64) = sum over (step, particle) of left[m] * right[c] with column blocks
// [a_bar k'^T (M) | k' g_mean^T (Dout) | a^2 g_var^T (Dout) | w [x~,1]^T (Din+1)]; sc = scalar sums.
__global__ void finalize_tc_grad_kernel(int M, int Din, int Dout, int Ctot, const double *__restrict__ R,
                                        const double *__restrict__ sc, GpDev gp, double *__restrict__ gP,
                                        double *__restrict__ galpha, double *__restrict__ gS,
                                        double *__restrict__ gZ, double *__restrict__ gell,
                                        double *__restrict__ gsig2) {
  const double sig2 = (double)gp.sig2[0];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < M * M; i += nt) gP[i] = sig2 * R[(size_t)(i / M) * Ctot + (i % M)];
  for (int i = tid; i < M * Dout; i += nt) {
    const int m = i / Dout, d = i % Dout;
    galpha[i] = sig2 * R[(size_t)m * Ctot + M + d];
    gS[i] = R[(size_t)m * Ctot + M + Dout + d];
  }
  const int xc = M + 2 * Dout;
  for (int i = tid; i < M * Din; i += nt) {
    const int m = i / Din, j = i % Din;
    const double ell = (double)gp.ell[j];
    const double zt = (double)gp.Z[i] / ell;
    gZ[i] = (R[(size_t)m * Ctot + xc + j] - zt * R[(size_t)m * Ctot + xc + Din]) / ell;
  }
  for (int j = tid; j < Din; j += nt) gell[j] = sc[j] / (double)gp.ell[j];
  if (tid == 0) gsig2[0] = sc[Din] / sig2 + sc[Din + 1];
}

// CBFSSMHALF: d loss / d x_0[b][j] = sum over the S particles of sequence b (x_0 is tiled, cbfssmhalf.py:80,92)
__global__ void reduce_x0b_kernel(const float *__restrict__ x0b, int npad, int S, int nb, int dx,
                                  double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * dx) return;
  const int b = i / dx, j = i % dx;
  double s = 0.0;
  for (int k = 0; k < S; ++k) s += (double)x0b[(size_t)j * npad + (size_t)b * S + k];
  out[i] = s;
}

__global__ void finalize_noise_grad_kernel(int dx, int dy, int Din, const double *__restrict__ sc_f,
                                           const double *__restrict__ sc_b, const double *__restrict__ stats,
                                           const float *__restrict__ vy, double w_ll, double n_times_t,
                                           double *__restrict__ gvx, double *__restrict__ gvy) {
  const int j = threadIdx.x;
  if (j >= dx) return;
  gvx[j] = sc_f[Din + 2 + j] + sc_b[Din + 2 + j];
  double g = sc_f[Din + 2 + dx + j];
  if (j < dy) {
    const double v = (double)vy[j];
    g += w_ll * (0.5 * stats[j] / (v * v) - 0.5 * n_times_t / v);
  }
  gvy[j] = g;
}

// x_final / y_tilde in the reference layout [nb, T, S, dx]  (cbfssm.py:97,181)
__global__ void export_states_kernel(Dims D, int dx, int dy, const float *__restrict__ y, Workspace ws,
                                     float *__restrict__ x_final, float *__restrict__ y_tilde) {
  const int dh = dx - dy;
  const size_t total = (size_t)D.n_local * D.T * dx;
  const size_t np = ws.npad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % dx);
    size_t r = i / dx;
    const int s = (int)(r % D.S); r /= D.S;
    const int t = (int)(r % D.T);
    const int bl = (int)(r / D.T);
    const int nl = bl * D.S + s;
    if (x_final) x_final[i] = ws.X[((size_t)t * dx + j) * np + nl];
    if (y_tilde) {
      float v;
      if (j < dy) {
        const int b = (D.n_offset + nl) / D.S;
        v = y[((size_t)b * D.T + t) * dy + j];
      } else {
        v = ws.H[(((size_t)writer_run(t, D.R) * D.T + t) * dh + (j - dy)) * np + nl];
      }
      y_tilde[i] = v;
    }
  }
}

// Per-sequence partial sums [sum_s x, sum_s x^2] over THIS shard's particles (any particle range, not only
// sequence-aligned ones): one warp per (local sequence row, t).  With the particles of a sequence spread over
// ranks, tf.nn.moments over the particle axis (cbfssm.py:267-269) is an all-reduce of these sums.
__global__ void state_sums_kernel(Dims D, int dx, Workspace ws, double *__restrict__ sums) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= D.B * D.T) return;
  const int b = warp / D.T, t = warp - b * D.T;
  const long long lo = (long long)b * D.S - D.n_offset, hi = lo + D.S;
  const int n0 = (int)(lo < 0 ? 0 : lo), n1 = (int)(hi > D.n_local ? D.n_local : hi);
  const size_t np = ws.npad;
  for (int j = 0; j < dx; ++j) {
    double s1 = 0.0, s2 = 0.0;
    const float *xp = ws.X + ((size_t)t * dx + j) * np;
    for (int n = n0 + lane; n < n1; n += 32) {
      const double v = (double)xp[n];
      s1 += v;
      s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      sums[((size_t)warp * dx + j) * 2] = s1;
      sums[((size_t)warp * dx + j) * 2 + 1] = s2;
    }
  }
}

// tf.nn.moments(axes=[2]) over the particle axis (cbfssm.py:267-269): one warp per (b,t).
__global__ void moments_kernel(const float *__restrict__ x, int rows, int S, int d, int d_keep,
                               const float *__restrict__ add_var, float *__restrict__ mean,
                               float *__restrict__ var) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  for (int j = 0; j < d_keep; ++j) {
    double s1 = 0.0;
    for (int s = lane; s < S; s += 32) s1 += (double)x[((size_t)warp * S + s) * d + j];
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    const double mu = s1 / S;
    double s2 = 0.0;
    for (int s = lane; s < S; s += 32) {
      const double e = (double)x[((size_t)warp * S + s) * d + j] - mu;
      s2 += e * e;
    }
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) {
      mean[(size_t)warp * d_keep + j] = (float)mu;
      var[(size_t)warp * d_keep + j] = (float)(s2 / S + (add_var ? (double)add_var[j] : 0.0));
    }
  }
}

__global__ void adam_kernel(int64_t n, double *__restrict__ theta, const double *__restrict__ grad,
                            double *__restrict__ m, double *__restrict__ v, double lr_t, double b1, double b2,
                            double eps) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double g = grad[i];
  const double mi = b1 * m[i] + (1.0 - b1) * g;
  const double vi = b2 * v[i] + (1.0 - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  theta[i] -= lr_t * mi / (sqrt(vi) + eps);
}

// Philox4x32-10 (Salmon et al. 2011) -> Box-Muller; counter = element index / 4.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void fill_normal_kernel(float *__restrict__ out, int64_t n, uint64_t seed, uint64_t stream_id) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  float z[4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;   // (0,1]
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
    const float rr = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincospif(2.f * u2, &sn, &cs);
    z[2 * h] = rr * cs;
    z[2 * h + 1] = rr * sn;
  }
  if (q * 4 + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {   // one 16-byte store per thread: full sectors
    *reinterpret_cast<float4 *>(out + q * 4) = make_float4(z[0], z[1], z[2], z[3]);
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (q * 4 + h < n) out[q * 4 + h] = z[h];
  }
}

static int device_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

static int rev_grid(const Plan &p, int which, int items) {
  int occ = p.ops->occupancy(p.D.M, which);
  if (occ < 1) occ = 1;
  int g = device_sms() * occ;
  if (g > kMaxGridRev) g = kMaxGridRev;
  if (g > items) g = items;
  return g < 1 ? 1 : g;
}

static GpDev to_dev(const cbf_gp *g) { return GpDev{g->Z, g->ell, g->sig2, g->P, g->alpha, g->S}; }

static int check_gp(const cbf_gp *g, const char *name) {
  if (!g || !g->Z || !g->ell || !g->sig2 || !g->P || !g->alpha || !g->S) {
    set_error("%s has a NULL operand", name);
    return CBF_ERR_NULL;
  }
  if (misaligned(g->Z) || misaligned(g->P) || misaligned(g->alpha) || misaligned(g->S)) {
    set_error("%s operands must be 16-byte aligned", name);
    return CBF_ERR_ALIGNMENT;
  }
  return 0;
}

#define CBF_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));               \
      return (int)_e;                                                          \
    }                                                                          \
  } while (0)

}  // namespace cbf

using namespace cbf;

extern "C" {

CBF_API int cbf_abi_version(void) { return CBF_ABI_VERSION; }

CBF_API const char *cbf_last_error_string(void) { return g_err; }

CBF_API int cbf_supported(int32_t M, int32_t dx, int32_t du, int32_t dy) {
  const DimOps *ops = find_ops(dx, du, dy, M, true);
  if (M < 1) return 0;
  const bool f64_ok = dx >= 2 && dy >= 1 && dy < dx && du >= 0 && dx <= 16 && dx + du + 1 <= 32;
  if (!ops || M > 128) return f64_ok ? 3 : 0;     // only the float64 batched path takes these
  bool coop_fits = true;
  for (int w = 0; w < 4; ++w)
    if (ops->smem_bytes(M, w) > kMaxSmem) coop_fits = false;
  if (!coop_fits) {   // only the tensor path can take it
    if (ops->fw_forward_tc == nullptr || ops->fw_reverse_tc == nullptr || M < kMinTensorM || M > 128) return 0;
    for (int w = 0; w < 4; ++w)
      if (ops->smem_tc(M, w) > kMaxSmem) return 0;
    return 1;
  }
  return ops->fixed_M ? 2 : 1;
}

CBF_API int cbf_workspace_bytes(const cbf_shape *shape, size_t *bytes_out) {
  if (!bytes_out) { set_error("bytes_out is NULL"); return CBF_ERR_NULL; }
  Plan p;
  int rc = make_plan(shape, p, false);
  if (rc) return rc;
  *bytes_out = p.total;
  return 0;
}

CBF_API int cbf_grad_layout_get(const cbf_shape *shape, cbf_grad_layout *out) {
  if (!out) { set_error("out is NULL"); return CBF_ERR_NULL; }
  Plan p;
  int rc = make_plan(shape, p, false);
  if (rc) return rc;
  *out = grad_layout(p);
  return 0;
}

}  // extern "C"

static int elbo_forward_impl(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b, const float *var_x,
                             const float *var_y, const float *u, const float *y, const float *x0,
                             const float *eps_b, const float *z_b, const float *eps_f, double *terms,
                             void *workspace, void *stream) {
  Plan p;
  int rc = make_plan(shape, p, true);
  if (rc) return rc;
  if ((rc = check_gp(gp_f, "gp_f")) || (!p.half && (rc = check_gp(gp_b, "gp_b")))) return rc;
  if (!var_x || !var_y || !u || !y || (!p.half && (!eps_b || !z_b)) || (p.half && !x0) ||
      (!eps_f && shape->T > 1) || !terms || !workspace) {
    set_error("cbf_elbo_forward: NULL argument");
    return CBF_ERR_NULL;
  }
  if (misaligned(workspace)) { set_error("workspace must be 16-byte aligned"); return CBF_ERR_ALIGNMENT; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = bind_workspace(p, workspace);
  ws.x0 = x0;
  if (p.f64) {
    if (!gp_f->state || (!p.half && !gp_b->state)) { set_error("float64 path: cbf_gp.state (the cbf_gp_prologue state) is NULL"); return CBF_ERR_NULL; }
    F64Args a{p.D, p.dx, p.du, p.dy, &p.chains, gp_f->state, p.half ? nullptr : gp_b->state, var_x, var_y, u, y, eps_b, z_b,
              eps_f, ws, static_cast<char *>(workspace) + p.off_f64, st};
    ScopedTiming tm(1, st);
    return f64_forward(a, terms);
  }
  const int nch = (int)p.chains.size();
  // tensor-core forward kernels: a 128-particle tile makes the M x M contraction a real GEMM
  const bool tc = p.tc_fwd;
  const int pt = tc ? ceil_div(p.D.n_local, 128) : p.ptiles;
  for (int c0 = 0; c0 < nch; c0 += kMaxChains) {
    ChainTable ct;
    ct.count = nch - c0 < kMaxChains ? nch - c0 : kMaxChains;
    memcpy(ct.c, p.chains.data() + c0, sizeof(Chain) * ct.count);
    ScopedTiming tm(0, st);
    CBF_CUDA((tc ? p.ops->bm_forward_tc : p.ops->bm_forward)(p.D, ct, to_dev(gp_b), var_x, u, y, eps_b, z_b, ws,
                                                             ws.fpart_bm + (size_t)c0 * pt, st));
  }
  {
    ScopedTiming tm(1, st);
    CBF_CUDA((tc ? p.ops->fw_forward_tc : p.ops->fw_forward)(p.D, to_dev(gp_f), var_x, var_y, u, y, eps_f, ws,
                                                             ws.fpart_fw, st));
  }
  finalize_terms_kernel<<<1, 256, 0, st>>>(ws.fpart_bm, nch * pt, ws.fpart_fw, pt, p.dy, var_y,
                                           (double)p.D.n_local * p.D.T, ws.stats, terms); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" {

CBF_API int cbf_elbo_forward(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b, const float *var_x,
                     const float *var_y, const float *u, const float *y, const float *eps_b, const float *z_b,
                     const float *eps_f, double *terms, void *workspace, void *stream) {
  if (shape && (shape->flags & CBF_FLAG_HALF_MODEL)) { set_error("use cbf_elbo_forward_half with CBF_FLAG_HALF_MODEL"); return CBF_ERR_INVALID_SHAPE; }
  return elbo_forward_impl(shape, gp_f, gp_b, var_x, var_y, u, y, nullptr, eps_b, z_b, eps_f, terms, workspace, stream);
}

CBF_API int cbf_elbo_forward_half(const cbf_shape *shape, const cbf_gp *gp_f, const float *var_x, const float *var_y,
                          const float *u, const float *y, const float *x0, const float *eps_f, double *terms,
                          void *workspace, void *stream) {
  if (!shape || !(shape->flags & CBF_FLAG_HALF_MODEL)) { set_error("cbf_elbo_forward_half needs CBF_FLAG_HALF_MODEL in shape.flags"); return CBF_ERR_INVALID_SHAPE; }
  return elbo_forward_impl(shape, gp_f, nullptr, var_x, var_y, u, y, x0, nullptr, nullptr, eps_f, terms, workspace, stream);
}

}  // extern "C"

static int elbo_backward_impl(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b, const float *var_x,
                              const float *var_y, const float *u, const float *y, const float *x0,
                              const float *eps_b, const float *z_b, const float *eps_f,
                              const double *term_weights_host, double *grad_flat, double *x0_bar, void *workspace,
                              void *stream) {
  Plan p;
  int rc = make_plan(shape, p, true);
  if (rc) return rc;
  if ((rc = check_gp(gp_f, "gp_f")) || (!p.half && (rc = check_gp(gp_b, "gp_b")))) return rc;
  if (shape->flags & CBF_FLAG_PREDICT_ONLY) {
    set_error("cbf_elbo_backward: the forward pass ran with CBF_FLAG_PREDICT_ONLY (no entropy, partial message)");
    return CBF_ERR_INVALID_SHAPE;
  }
  if (!var_x || !var_y || !u || !y || (!p.half && (!eps_b || !z_b)) || (p.half && (!x0 || !x0_bar)) ||
      (!eps_f && shape->T > 1) || !term_weights_host || !grad_flat || !workspace) {
    set_error("cbf_elbo_backward: NULL argument");
    return CBF_ERR_NULL;
  }
  if (p.half && (shape->n_local % shape->S != 0 || shape->n_offset % shape->S != 0)) {
    set_error("CBFSSMHALF needs a sequence-aligned shard (n_offset, n_local multiples of S)");
    return CBF_ERR_INVALID_SHAPE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = bind_workspace(p, workspace);
  ws.x0 = x0;
  const cbf_grad_layout gl = grad_layout(p);
  const float w_ll = (float)term_weights_host[0], w_kl = (float)term_weights_host[1],
              w_en = (float)term_weights_host[2];

  if (p.f64) {
    if (!gp_f->state || (!p.half && !gp_b->state)) { set_error("float64 path: cbf_gp.state (the cbf_gp_prologue state) is NULL"); return CBF_ERR_NULL; }
    F64Args a{p.D, p.dx, p.du, p.dy, &p.chains, gp_f->state, p.half ? nullptr : gp_b->state, var_x, var_y, u, y, eps_b, z_b,
              eps_f, ws, static_cast<char *>(workspace) + p.off_f64, st};
    F64Grad g{grad_flat, (long long)gl.total,
              grad_flat + gl.f_P, grad_flat + gl.f_alpha, grad_flat + gl.f_S, grad_flat + gl.f_Z, grad_flat + gl.f_ell, grad_flat + gl.f_sig2,
              grad_flat + gl.b_P, grad_flat + gl.b_alpha, grad_flat + gl.b_S, grad_flat + gl.b_Z, grad_flat + gl.b_ell, grad_flat + gl.b_sig2,
              grad_flat + gl.var_x, grad_flat + gl.var_y};
    {
      ScopedTiming tm(2, st);
      rc = f64_backward(a, term_weights_host[0], term_weights_host[1], term_weights_host[2], g);
    }
    if (rc) return rc;
    if (p.half) {
      const int nb = p.D.n_local / p.D.S;
      reduce_x0b_kernel<<<ceil_div(nb * p.dx, 256), 256, 0, st>>>(ws.x0b, ws.npad, p.D.S, nb, p.dx, x0_bar); cbf_note_launch();
      CBF_CUDA(cudaGetLastError());
    }
    return 0;
  }
  const int nch = (int)p.chains.size();
  if (p.tc_rev) {
    // ---- tensor-core path: rollout adjoints on tcgen05, parameter outer products as a tcgen05 split-K stream ----
    char *wb = static_cast<char *>(workspace);
    float *spf = reinterpret_cast<float *>(wb + p.off_spf), *spb = reinterpret_cast<float *>(wb + p.off_spb);
    double *rdf = reinterpret_cast<double *>(wb + p.off_rdf), *rdb = reinterpret_cast<double *>(wb + p.off_rdb);
    double *rpart = reinterpret_cast<double *>(wb + p.off_rpart);
    const int pt = ceil_div(p.D.n_local, 128);
    // forward-rollout GP: time windows from T-2 down to 0; each window's operand tiles are accumulated before
    // the next window overwrites them (one window when everything fits the budget)
    for (size_t wi = 0; wi < p.win_f.size(); ++wi) {
      const TimeWin win = p.win_f[wi];
      const TcMats mf = bind_mats(workspace, p.off_mats, (size_t)(win.t_hi - win.t_lo + 1) * p.D.n_local, p.D.M, p.dx, p.din);
      CBF_CUDA(clear_mats_tail(mf, st));
      {
        ScopedTiming tm(2, st);
        CBF_CUDA(p.ops->fw_reverse_tc(p.D, to_dev(gp_f), var_x, var_y, u, y, eps_f, w_ll, w_kl, ws, mf, win, spf, p.nsc_f, st));
      }
      reduce_slots_kernel<<<ceil_div(p.nsc_f, 256), 256, 0, st>>>(spf, p.nspart_f, p.nsc_f, ws.acc_f + p.Lf.scal_off(),
                                                                  wi > 0 ? 1 : 0); cbf_note_launch();
      CBF_CUDA(cudaGetLastError());
      ScopedTiming tm(4, st);
      CBF_CUDA(tc_outer(mf, p.D.M, p.dx, p.din, rpart, rdf, wi > 0, st));
    }
    // message GP: round r runs the r-th piece of every live chain (ascending time; the message adjoint crosses
    // pieces through the carry buffer).  Without a message GP (CBFSSMHALF) its scalar sums are zero.
    if (p.rounds_b.empty())
      CBF_CUDA(cudaMemsetAsync(ws.acc_b + p.Lb.scal_off(), 0, sizeof(double) * (size_t)p.nsc_b, st));
    for (size_t ri = 0; ri < p.rounds_b.size(); ++ri) {
      const std::vector<Chain> &round = p.rounds_b[ri];
      int cols = 0;
      for (const Chain &c : round) cols += c.t_hi - c.t_lo + 1;
      const TcMats mb = bind_mats(workspace, p.off_mats, (size_t)cols * p.D.n_local, p.D.M, p.dh, p.din);
      CBF_CUDA(clear_mats_tail(mb, st));
      const int nr = (int)round.size();
      for (int c0 = 0; c0 < nr; c0 += kMaxChains) {
        ChainTable ct;
        ct.count = nr - c0 < kMaxChains ? nr - c0 : kMaxChains;
        memcpy(ct.c, round.data() + c0, sizeof(Chain) * ct.count);
        ScopedTiming tm(3, st);
        CBF_CUDA(p.ops->bm_reverse_tc(p.D, ct, to_dev(gp_b), var_x, u, y, eps_b, z_b, w_en, ws, mb,
                                      spb + (size_t)c0 * pt * p.nsc_b, p.nsc_b, st));
      }
      reduce_slots_kernel<<<ceil_div(p.nsc_b, 256), 256, 0, st>>>(spb, pt * nr, p.nsc_b, ws.acc_b + p.Lb.scal_off(),
                                                                  ri > 0 ? 1 : 0); cbf_note_launch();
      CBF_CUDA(cudaGetLastError());
      ScopedTiming tm(5, st);
      CBF_CUDA(tc_outer(mb, p.D.M, p.dh, p.din, rpart, rdb, ri > 0, st));
    }
    finalize_tc_grad_kernel<<<8, 256, 0, st>>>(p.D.M, p.din, p.dx, p.ctot_f, rdf, ws.acc_f + p.Lf.scal_off(), to_dev(gp_f),
                                               grad_flat + gl.f_P, grad_flat + gl.f_alpha, grad_flat + gl.f_S,
                                               grad_flat + gl.f_Z, grad_flat + gl.f_ell, grad_flat + gl.f_sig2); cbf_note_launch();
    if (!p.half)
      finalize_tc_grad_kernel<<<8, 256, 0, st>>>(p.D.M, p.din, p.dh, p.ctot_b, rdb, ws.acc_b + p.Lb.scal_off(), to_dev(gp_b),
                                                 grad_flat + gl.b_P, grad_flat + gl.b_alpha, grad_flat + gl.b_S,
                                                 grad_flat + gl.b_Z, grad_flat + gl.b_ell, grad_flat + gl.b_sig2); cbf_note_launch();
    CBF_CUDA(cudaGetLastError());
  } else {
  // reverse of the forward rollout (writes the y2 adjoints), then of the message chains
  const int grid_f = rev_grid(p, 2, p.ptiles);
  {
    ScopedTiming tm(2, st);
    CBF_CUDA(p.ops->fw_reverse(p.D, to_dev(gp_f), var_x, var_y, u, y, eps_f, w_ll, w_kl, ws, ws.gpart_f, grid_f, st));
  }
  reduce_slots_kernel<<<ceil_div(p.Lf.slot(), 256), 256, 0, st>>>(ws.gpart_f, grid_f * p.slots_per_cta, p.Lf.slot(),
                                                                  ws.acc_f); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());

  int nslots_b = 0;
  for (int c0 = 0; c0 < nch; c0 += kMaxChains) {
    ChainTable ct;
    ct.count = nch - c0 < kMaxChains ? nch - c0 : kMaxChains;
    memcpy(ct.c, p.chains.data() + c0, sizeof(Chain) * ct.count);
    int grid_b = rev_grid(p, 3, p.ptiles * ct.count);
    if (nslots_b + grid_b > kMaxGridRev) grid_b = kMaxGridRev - nslots_b;
    if (grid_b < 1) { set_error("too many chain batches"); return CBF_ERR_INVALID_SHAPE; }
    {
      ScopedTiming tm(3, st);
      CBF_CUDA(p.ops->bm_reverse(p.D, ct, to_dev(gp_b), var_x, u, y, eps_b, z_b, w_en, ws,
                                 ws.gpart_b + (size_t)nslots_b * p.slots_per_cta * p.Lb.slot(), grid_b, st));
    }
    nslots_b += grid_b;
  }
  reduce_slots_kernel<<<ceil_div(p.Lb.slot(), 256), 256, 0, st>>>(ws.gpart_b, nslots_b * p.slots_per_cta, p.Lb.slot(),
                                                                  ws.acc_b); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());

  finalize_gp_grad_kernel<<<8, 256, 0, st>>>(p.Lf, ws.acc_f, to_dev(gp_f), grad_flat + gl.f_P, grad_flat + gl.f_alpha,
                                             grad_flat + gl.f_S, grad_flat + gl.f_Z, grad_flat + gl.f_ell,
                                             grad_flat + gl.f_sig2); cbf_note_launch();
  if (!p.half)
    finalize_gp_grad_kernel<<<8, 256, 0, st>>>(p.Lb, ws.acc_b, to_dev(gp_b), grad_flat + gl.b_P, grad_flat + gl.b_alpha,
                                               grad_flat + gl.b_S, grad_flat + gl.b_Z, grad_flat + gl.b_ell,
                                               grad_flat + gl.b_sig2); cbf_note_launch();
  }
  if (p.half) {   // no backward-message GP: its block of the flat gradient is zero; x_0 adjoint per sequence
    CBF_CUDA(cudaMemsetAsync(grad_flat + gl.b_P, 0, sizeof(double) * (size_t)(gl.var_x - gl.b_P), st));
    const int nb = p.D.n_local / p.D.S;
    reduce_x0b_kernel<<<ceil_div(nb * p.dx, 256), 256, 0, st>>>(ws.x0b, ws.npad, p.D.S, nb, p.dx, x0_bar); cbf_note_launch();
  }
  finalize_noise_grad_kernel<<<1, 32 * ceil_div(p.dx, 32), 0, st>>>(
      p.dx, p.dy, p.din, ws.acc_f + p.Lf.scal_off(), ws.acc_b + p.Lb.scal_off(), ws.stats, var_y, (double)term_weights_host[0],
      (double)p.D.n_local * p.D.T, grad_flat + gl.var_x, grad_flat + gl.var_y); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" {

CBF_API int cbf_elbo_backward(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b, const float *var_x,
                      const float *var_y, const float *u, const float *y, const float *eps_b, const float *z_b,
                      const float *eps_f, const double *term_weights_host, double *grad_flat, void *workspace,
                      void *stream) {
  if (shape && (shape->flags & CBF_FLAG_HALF_MODEL)) { set_error("use cbf_elbo_backward_half with CBF_FLAG_HALF_MODEL"); return CBF_ERR_INVALID_SHAPE; }
  return elbo_backward_impl(shape, gp_f, gp_b, var_x, var_y, u, y, nullptr, eps_b, z_b, eps_f, term_weights_host,
                            grad_flat, nullptr, workspace, stream);
}

CBF_API int cbf_elbo_backward_half(const cbf_shape *shape, const cbf_gp *gp_f, const float *var_x, const float *var_y,
                           const float *u, const float *y, const float *x0, const float *eps_f,
                           const double *term_weights_host, double *grad_flat, double *x0_bar, void *workspace,
                           void *stream) {
  if (!shape || !(shape->flags & CBF_FLAG_HALF_MODEL)) { set_error("cbf_elbo_backward_half needs CBF_FLAG_HALF_MODEL in shape.flags"); return CBF_ERR_INVALID_SHAPE; }
  return elbo_backward_impl(shape, gp_f, nullptr, var_x, var_y, u, y, x0, nullptr, nullptr, eps_f, term_weights_host,
                            grad_flat, x0_bar, workspace, stream);
}

CBF_API int cbf_export_states(const cbf_shape *shape, const float *y, float *x_final, float *y_tilde,
                      const void *workspace, void *stream) {
  Plan p;
  int rc = make_plan(shape, p, false);
  if (rc) return rc;
  if (!workspace || !y) { set_error("cbf_export_states: NULL argument"); return CBF_ERR_NULL; }
  if (shape->n_local % shape->S != 0 || shape->n_offset % shape->S != 0) {
    set_error("cbf_export_states needs a sequence-aligned shard (n_offset, n_local multiples of S)");
    return CBF_ERR_INVALID_SHAPE;
  }
  Workspace ws = bind_workspace(p, const_cast<void *>(workspace));
  const size_t total = (size_t)p.D.n_local * p.D.T * p.dx;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  export_states_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p.D, p.dx, p.dy, y, ws, x_final, y_tilde); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_state_sums(const cbf_shape *shape, double *sums, const void *workspace, void *stream) {
  Plan p;
  int rc = make_plan(shape, p, false);
  if (rc) return rc;
  if (!workspace || !sums) { set_error("cbf_state_sums: NULL argument"); return CBF_ERR_NULL; }
  Workspace ws = bind_workspace(p, const_cast<void *>(workspace));
  const int rows = p.D.B * p.D.T;
  state_sums_kernel<<<ceil_div(rows * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p.D, p.dx, ws, sums); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_moments(const float *x, int32_t nb, int32_t T, int32_t S, int32_t d, int32_t d_keep, const float *add_var,
                float *mean, float *var, void *stream) {
  if (!x || !mean || !var) { set_error("cbf_moments: NULL argument"); return CBF_ERR_NULL; }
  if (nb < 1 || T < 1 || S < 1 || d < 1 || d_keep < 1 || d_keep > d) {
    set_error("cbf_moments: invalid shape");
    return CBF_ERR_INVALID_SHAPE;
  }
  const int rows = nb * T;
  moments_kernel<<<ceil_div(rows * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, S, d, d_keep,
                                                                                         add_var, mean, var); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_adam_step(int64_t n, double *theta, const double *grad, double *m, double *v, int64_t step, double lr,
                  double beta1, double beta2, double eps, void *stream) {
  if (!theta || !grad || !m || !v) { set_error("cbf_adam_step: NULL argument"); return CBF_ERR_NULL; }
  if (n < 1 || step < 1) { set_error("cbf_adam_step: invalid n/step"); return CBF_ERR_INVALID_SHAPE; }
  const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)step)) / (1.0 - pow(beta1, (double)step));
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, theta, grad, m, v, lr_t,
                                                                                         beta1, beta2, eps); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_timing_enable(int enable) {
  g_timing.enabled = enable != 0;
  g_timing.used = 0;
  return 0;
}

CBF_API int cbf_timing_read(double *ms_sum_host, int64_t *count_host) {
  if (!ms_sum_host || !count_host) { set_error("cbf_timing_read: NULL argument"); return CBF_ERR_NULL; }
  for (int k = 0; k < 8; ++k) { ms_sum_host[k] = 0.0; count_host[k] = 0; }
  TimingPool &t = g_timing;
  for (size_t i = 0; i < t.used; ++i) {
    CBF_CUDA(cudaEventSynchronize(t.ev[2 * i + 1]));
    float ms = 0.f;
    CBF_CUDA(cudaEventElapsedTime(&ms, t.ev[2 * i], t.ev[2 * i + 1]));
    ms_sum_host[t.kind[i]] += ms;
    count_host[t.kind[i]] += 1;
  }
  t.used = 0;
  return 0;
}

CBF_API int cbf_launches_read(int64_t *count_host, int reset) {
  if (!count_host) { set_error("cbf_launches_read: NULL argument"); return CBF_ERR_NULL; }
  *count_host = (int64_t)g_launches;
  if (reset) g_launches = 0;
  return 0;
}

// Measured FP32 roof of this GPU for bench.py's roofline: packed FMA (fma.rn.f32x2, the instruction the rollout
// kernels' inner loops are made of), 8 independent accumulator pairs per thread, 8 CTAs of 256 threads per SM.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float s) {
  auto pk = [](float x, float y) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; };
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const unsigned long long ss = pk(s, s), h = pk(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ss), "l"(h));
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i]));
    t += x + y;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

CBF_API int cbf_measure_fp32_peak(float *scratch, int iters, double *tflops_host, void *stream) {
  if (!scratch || !tflops_host) { set_error("cbf_measure_fp32_peak: NULL argument"); return CBF_ERR_NULL; }
  if (iters < 1) { set_error("cbf_measure_fp32_peak: iters < 1"); return CBF_ERR_INVALID_SHAPE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  CBF_CUDA(cudaGetDevice(&dev));
  CBF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  CBF_CUDA(cudaEventCreate(&e0));
  CBF_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {       // first repetition warms up; best of the rest
    cudaEventRecord(e0, st);
    fp32_peak_kernel<<<sms * 8, 256, 0, st>>>(scratch, iters, 0.999f); cbf_note_launch();
    cudaEventRecord(e1, st);
    CBF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CBF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 2.0 * 32.0 * (double)iters * 256.0 * 8.0 * sms;   // 2 lanes x 2 flop x 32 fma2 per iteration
    if (rep > 0 && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_host = best;
  return 0;
}

CBF_API int cbf_fill_normal(float *out, int64_t n, uint64_t seed, uint64_t stream_id, void *stream) {
  if (!out) { set_error("cbf_fill_normal: NULL argument"); return CBF_ERR_NULL; }
  if (n < 1) return 0;
  const int64_t quads = (n + 3) / 4;
  fill_normal_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed,
                                                                                                    stream_id); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
