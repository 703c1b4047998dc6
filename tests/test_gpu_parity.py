"""CUDA path (through the C ABI) against the float64 oracle on identical inputs and
identical injected normal draws.  Tolerance: 1e-4 relative (north_star), measured per
tensor as max|a-b| / max|b|."""
import numpy as np
import pytest
import torch

from oracle import cbfssm_oracle as O
from oracle import kernel_math as KM
from tests.helpers import make_problem, rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-4


def run_engine(cfg, params, u, y, eps_b, z_b, eps_f, condition=True, flags=0):
    from cbf_ssm_b200.engine import ElboEngine, ModelDims
    dims = ModelDims(cfg.dim_x, cfg.dim_u, cfg.dim_y, cfg.ind_pnt_num, cfg.samples, cfg.recog_len,
                     cfg.k_factor, tuple(cfg.loss_factors))
    eng = ElboEngine(dims)
    eng.flags = flags
    eng.set_params({k: v.numpy() for k, v in params.items()})
    dev = eng.device
    B, T, _ = u.shape
    N = B * cfg.samples
    f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    ud, yd = f32(u), f32(y)
    eb, zb, ef = f32(eps_b.reshape(2, T, N)), f32(z_b.reshape(2, T, N)), f32(eps_f.reshape(T - 1, N))
    out = eng.forward(ud, yd, eb, zb, ef, condition)
    eng.backward()
    torch.cuda.synchronize()
    return eng, out, yd


CASES = [
    # dx du dy  M  S  B   T  R  kap   lf           cond  strong
    (4, 2, 2, 7, 3, 2, 11, 3, 1.0, (10.0, 0.5), True, True),
    (3, 1, 1, 5, 4, 3, 9, 2, 5.0, (6.0, 1.0), False, True),
    (4, 1, 1, 6, 2, 2, 6, 50, 1.0, (10.0, 0.0), True, True),
    (4, 2, 2, 20, 50, 3, 40, 8, 1.0, (20.0, 0.0), True, False),
    (4, 1, 1, 100, 10, 2, 30, 10, 1.0, (10.0, 0.0), True, False),
    (14, 7, 7, 33, 5, 2, 12, 3, 50.0, (6.0, 0.0), True, True),
    (13, 6, 7, 20, 9, 2, 10, 4, 1.0, (20.0, 0.2), True, True),
    (2, 1, 1, 12, 40, 2, 20, 4, 1.0, (10.0, 1.0), True, True),
    (8, 1, 4, 16, 7, 2, 14, 16, 2.0, (10.0, 0.0), True, True),
    (16, 1, 8, 9, 6, 1, 8, 2, 1.0, (10.0, 0.3), True, True),
    # shapes with a register-resident instantiation (csrc/dims_list.h CBF_FAST_LIST)
    (3, 1, 1, 12, 37, 5, 13, 3, 3.0, (6.0, 1.0), True, True),
    (4, 1, 1, 20, 50, 3, 30, 16, 10.0, (10.0, 0.0), True, False),
    (2, 1, 1, 20, 33, 4, 21, 4, 1.0, (10.0, 1.0), False, True),
    (8, 1, 4, 20, 7, 5, 14, 16, 2.0, (10.0, 0.0), True, True),
    (16, 1, 8, 20, 6, 3, 8, 2, 1.0, (10.0, 0.3), True, True),
    (13, 6, 7, 20, 40, 4, 12, 4, 1.0, (20.0, 0.2), True, False),
    # D >= 8 on the tensor path: two 128-particle tiles per CTA; 3 tiles = one full CTA + one CTA whose second
    # tile is empty and whose first is partly filled
    (8, 1, 4, 64, 30, 11, 10, 4, 1.0, (10.0, 0.5), True, True),
    (14, 7, 7, 50, 26, 10, 8, 3, 50.0, (6.0, 0.0), True, True),
    # M = 128: the resident set of the cooperative kernels no longer fits an SM, only the tensor path runs it
    (4, 2, 2, 128, 20, 7, 12, 4, 1.0, (10.0, 0.3), True, True),
    # sweep corner of BASELINE.json configs[4] at M = 100, D = 16.  (D = 2 at M = 100 -- 3 input dims, the densest
    # inducing set, cond K_zz ~ 1e5 -- is beyond float32: see the float64 cases and the ill-conditioned test below.)
    (16, 1, 8, 100, 12, 3, 10, 4, 1.0, (10.0, 0.3), True, True),
]

# 12 = CBF_FLAG_FORCE_REGISTER | CBF_FLAG_FORCE_TENSOR_CORES: the register-resident kernels
#      (small M) / tcgen05 forward kernels (M >= 48) regardless of the particle count;
# 1  = CBF_FLAG_FORCE_COOPERATIVE: shared-memory cooperative kernels only
PATHS = [12, 1]


@pytest.mark.parametrize("flags", PATHS, ids=["register_or_tensor", "cooperative"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "dx%d_du%d_dy%d_M%d_S%d_B%d_T%d_R%d" % c[:8])
def test_elbo_and_gradients_match_oracle(case, flags):
    dx, du, dy, M, S, B, T, R, kap, lf, cond, strong = case
    if M > 110 and flags == 1:
        pytest.skip("cooperative kernels need M <= ~110")
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, kap, lf, seed=7, strong=strong)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, cond)
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, flags)

    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        ref = float(getattr(res, k).detach())
        got = float(out[k])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref)

    xf, yt = eng.export_states(yd)
    assert rel_inf(xf.cpu().numpy(), res.x_final.detach().numpy()) < TOL
    assert rel_inf(yt.cpu().numpy(), res.y_tilde.detach().numpy()) < TOL

    pm, pv = eng.moments(xf, dy, eng.var_y)
    im, iv = eng.moments(xf, dx, None)
    torch.cuda.synchronize()
    assert rel_inf(pm.cpu().numpy(), res.pred_mean.detach().numpy()) < TOL
    assert rel_inf(pv.cpu().numpy(), res.pred_var.detach().numpy()) < TOL
    assert rel_inf(im.cpu().numpy(), res.internal_mean.detach().numpy()) < TOL
    assert rel_inf(iv.cpu().numpy(), res.internal_var.detach().numpy()) < TOL

    grads = eng.get_grads()
    worst = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in worst.items() if not v < TOL}
    assert not bad, bad


def _random_cases(count, seed):
    """Seeded random shapes over the compiled (dx, du, dy) triples: ragged tiles, R above and below T, one-particle
    and one-sequence batches, both branches of `condition`."""
    rng = np.random.RandomState(seed)
    dims = [(4, 2, 2), (4, 1, 1), (14, 7, 7), (13, 6, 7), (2, 1, 1), (8, 1, 4), (16, 1, 8), (3, 1, 1)]
    out = []
    for _ in range(count):
        dx, du, dy = dims[rng.randint(len(dims))]
        m_hi = 14 if dx + du <= 4 else 48          # few input dims: keep the inducing set well conditioned for float32
        M = int(rng.randint(2, m_hi + 1))
        S, B, T = int(rng.randint(1, 41)), int(rng.randint(1, 7)), int(rng.randint(2, 25))
        R = int(rng.randint(1, 31))
        kap = float(rng.choice([1.0, 2.0, 10.0, 50.0]))
        lf = (float(rng.choice([6.0, 10.0, 20.0])), float(rng.choice([0.0, 0.3, 1.0])))
        out.append((dx, du, dy, M, S, B, T, R, kap, lf, bool(rng.rand() < 0.75), True))
    return out


@pytest.mark.parametrize("flags", [0, 12], ids=["default_dispatch", "register_or_tensor"])
@pytest.mark.parametrize("case", _random_cases(14, seed=20261018), ids=lambda c: "dx%d_du%d_dy%d_M%d_S%d_B%d_T%d_R%d" % c[:8])
def test_random_shapes_match_oracle(case, flags):
    dx, du, dy, M, S, B, T, R, kap, lf, cond, strong = case
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, kap, lf, seed=101, strong=strong)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, cond)
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, flags)
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        ref, got = float(getattr(res, k).detach()), float(out[k])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref)
    xf, yt = eng.export_states(yd)
    assert rel_inf(xf.cpu().numpy(), res.x_final.detach().numpy()) < TOL
    assert rel_inf(yt.cpu().numpy(), res.y_tilde.detach().numpy()) < TOL
    grads = eng.get_grads()
    bad = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in bad.items() if not v < TOL}
    assert not bad, bad


def _f32_exact(*arrays):
    """The C ABI takes u, y and the draws as float32.  For ill-conditioned inducing sets the *problem* is sensitive to
    that rounding (the float64 oracle's own gradients move by up to 1e-2 at cond 1e9 when its inputs are rounded), so
    the float64 comparisons give the oracle the very numbers the GPU gets."""
    return tuple(np.asarray(a, np.float32).astype(np.float64) for a in arrays)


# The float64 batched path (CBF_FLAG_FP64 = 128; automatic for M > 128 and for dims without an instantiation)
F64_CASES = [
    (4, 2, 2, 7, 3, 2, 11, 3, 1.0, (10.0, 0.5), True, True),
    (3, 1, 1, 5, 4, 3, 9, 2, 5.0, (6.0, 1.0), False, True),
    (4, 1, 1, 6, 2, 2, 6, 50, 1.0, (10.0, 0.0), True, True),
    (14, 7, 7, 33, 5, 2, 12, 3, 50.0, (6.0, 0.0), True, True),
    (16, 1, 8, 20, 6, 3, 8, 2, 1.0, (10.0, 0.3), True, True),
    # BASELINE.json configs[4] corners: M = 500 at D = 8; D = 2 and D = 16 at M = 100
    (8, 1, 4, 500, 6, 2, 10, 4, 1.0, (10.0, 0.3), True, True),
    (2, 1, 1, 100, 30, 4, 20, 5, 1.0, (10.0, 0.5), True, True),
    (2, 1, 1, 100, 30, 4, 20, 5, 1.0, (10.0, 0.0), False, False),
    (16, 1, 8, 100, 12, 3, 10, 4, 1.0, (10.0, 0.3), True, True),
    # dims nobody compiled kernels for
    (5, 2, 3, 40, 6, 2, 10, 3, 2.0, (10.0, 0.2), True, True),
]


@pytest.mark.parametrize("case", F64_CASES, ids=lambda c: "dx%d_du%d_dy%d_M%d_S%d_B%d_T%d_R%d" % c[:8])
def test_float64_path_matches_oracle(case):
    dx, du, dy, M, S, B, T, R, kap, lf, cond, strong = case
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, kap, lf, seed=7, strong=strong)
    u, y, eps_b, z_b, eps_f = _f32_exact(u, y, eps_b, z_b, eps_f)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, cond)
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, 128)
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        ref, got = float(getattr(res, k).detach()), float(out[k])
        assert abs(got - ref) <= 1e-5 * max(abs(ref), 1e-3), (k, got, ref)
    xf, yt = eng.export_states(yd)
    pm, pv = eng.moments(xf, dy, eng.var_y)
    torch.cuda.synchronize()
    assert rel_inf(xf.cpu().numpy(), res.x_final.detach().numpy()) < 1e-5
    assert rel_inf(yt.cpu().numpy(), res.y_tilde.detach().numpy()) < 1e-5
    assert rel_inf(pv.cpu().numpy(), res.pred_var.detach().numpy()) < 1e-5
    grads = eng.get_grads()
    bad = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in bad.items() if not v < 2e-5}       # float64 arithmetic; float32 only where states are stored
    assert not bad, bad


@pytest.mark.parametrize("flags", PATHS, ids=["register_or_tensor", "cooperative"])
def test_kernel_level_gradients_match_kernel_math(flags):
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(4, 2, 2, 7, 3, 2, 11, 3, 1.0, (10.0, 0.5), seed=3, strong=True)
    pn = {k: v.numpy() for k, v in params.items()}
    out, _ = KM.elbo_value_and_grad(cfg, pn, u, y, eps_b, z_b, eps_f, True)
    eng, _, _ = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, True, flags)
    kg = eng.kernel_level_grads()
    ref = out["kernel_level"]
    for tag in ("f", "b"):
        for nm in ("P", "alpha", "S", "Z", "ell"):
            assert rel_inf(kg[f"{tag}.{nm}"], ref[tag][nm]) < TOL, (tag, nm)
        assert rel_inf(kg[f"{tag}.sig2"], np.asarray([ref[tag]["sig2"]])) < TOL
    assert rel_inf(kg["var_x"], ref["var_x"]) < TOL
    assert rel_inf(kg["var_y"], ref["var_y"]) < TOL


@pytest.mark.parametrize("budget", [40_000, 150_000, 600_000], ids=["1-step windows", "few-step windows", "two windows"])
@pytest.mark.parametrize("shape", [(4, 1, 1, 100, 10, 2, 30, 10), (8, 1, 4, 64, 30, 11, 10, 4)],
                         ids=["dx4_M100_T30_R10", "dx8_M64_T10_R4"])
def test_tensor_path_time_windows_match_oracle(shape, budget, monkeypatch):
    """The tensor-path reverse pass cut into time windows (operand tiles of one window at a time, state and
    message adjoints carried between launches) gives the same gradients as the oracle."""
    monkeypatch.setenv("CBFSSM_B200_TC_WINDOW_BYTES", str(budget * (shape[4] * shape[5]) // 20))
    dx, du, dy, M, S, B, T, R = shape
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, 1.0, (10.0, 0.7), seed=13, strong=True)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, True)
    eng, out, _ = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, True, 12)
    assert abs(float(out["loss"]) - float(res.loss.detach())) <= TOL * abs(float(res.loss.detach()))
    grads = eng.get_grads()
    bad = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES if not rel_inf(grads[k], gd[k].numpy()) < TOL}
    assert not bad, bad


@pytest.mark.parametrize("split", ["0", "1", "tiles3"],
                         ids=["one_thread_per_particle", "two_threads_per_particle", "three_tiles_per_cta"])
@pytest.mark.parametrize("case", [
    (4, 1, 1, 100, 10, 2, 30, 10, 1.0, (10.0, 0.0), True, False),        # M = 100 compile-time instantiation
    (4, 2, 2, 100, 16, 8, 40, 10, 1.0, (10.0, 0.3), True, True),
    (4, 2, 2, 128, 20, 7, 12, 4, 1.0, (10.0, 0.3), True, True),          # runtime M, 8 chunks: 2 per group
    (4, 1, 1, 24, 40, 2, 20, 4, 1.0, (10.0, 1.0), False, True),          # 2 chunks: groups 2, 3 own no rows
    (3, 1, 1, 40, 37, 5, 13, 3, 3.0, (6.0, 1.0), True, True),            # 3 chunks, 2 ragged tiles
    (4, 2, 2, 100, 50, 11, 12, 4, 1.0, (10.0, 0.5), True, True),         # 5 tiles: a full and a partly filled 3-tile CTA
], ids=lambda c: "dx%d_du%d_dy%d_M%d_S%d_B%d_T%d_R%d" % c[:8])
def test_tensor_path_latency_variant_matches_oracle(case, split, monkeypatch):
    """The tensor path has a latency variant (2 threads per particle, each half of the M rows; chosen when a launch
    has fewer CTAs than the GPU has SMs) besides the throughput variants (1 thread per particle; one particle tile
    per CTA, or three tiles sharing P where that fills whole waves better).  Every parity case in this file is small
    enough to get the first by default; this test forces each variant in turn, also through the time-window
    machinery."""
    monkeypatch.setenv("CBFSSM_B200_TC_SPLIT", "0" if split == "tiles3" else split)
    monkeypatch.setenv("CBFSSM_B200_TC_TILES", "3" if split == "tiles3" else "1")
    dx, du, dy, M, S, B, T, R, kap, lf, cond, strong = case
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, kap, lf, seed=17, strong=strong)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, cond)
    for budget in (None, "200000"):
        if budget is None:
            monkeypatch.delenv("CBFSSM_B200_TC_WINDOW_BYTES", raising=False)
        else:
            monkeypatch.setenv("CBFSSM_B200_TC_WINDOW_BYTES", budget)
        eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, 8)
        for k in ("loss", "loglik", "kl_x", "entropy"):
            ref, got = float(getattr(res, k).detach()), float(out[k])
            assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref, budget)
        xf, yt = eng.export_states(yd)
        assert rel_inf(xf.cpu().numpy(), res.x_final.detach().numpy()) < TOL
        assert rel_inf(yt.cpu().numpy(), res.y_tilde.detach().numpy()) < TOL
        grads = eng.get_grads()
        bad = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES}
        bad = {k: v for k, v in bad.items() if not v < TOL}
        assert not bad, (bad, budget)


def _cond_kzz(params, tag):
    """2-norm condition number of K_zz + 1e-8 I as the reference forms it (gp_tf.py:48-54)."""
    import torch
    kern = O.RBF(params[f"{tag}.variance_unc"], params[f"{tag}.lengthscales_unc"])
    K = kern.K(params[f"{tag}.zeta_pos"]).numpy() + 1e-8 * np.eye(params[f"{tag}.zeta_pos"].shape[0])
    ev = np.linalg.eigvalsh(K)
    return float(ev[-1] / ev[0])


@pytest.mark.parametrize("zeta_pos", [2.0, 1.0, 0.3], ids=["spread_cond_1e5", "crowded_cond_1e9", "clustered_cond_5e9"])
def test_ill_conditioned_inducing_sets_stay_finite_and_report_their_error(zeta_pos, capsys):
    """The kernels use an explicit float32 inverse P = (K_zz + 1e-8 I)^-1 where the reference does float64 Cholesky
    solves (gp_tf.py:137,145).  As the inducing points crowd together cond(K_zz) grows towards sigma^2 M / 1e-8 and
    float32 runs out of digits (a = P k amplifies the 1e-7 rounding of k by cond(K_zz)): for M = 100 inducing points
    in 3 input dims even the reference's own initialisation (zeta_pos = 2) has cond 1.4e5, so this corner is
    accuracy-limited on the float32 paths BY DESIGN.  What must hold on them: the GP variance never
    goes negative (gp_var_clamp), so loss, states and all gradients stay finite; their measured errors are
    printed.  The float64 path (CBF_FLAG_FP64) is the answer for such inducing sets and must match at 1e-4."""
    dx, du, dy, M, S, B, T, R = 2, 1, 1, 100, 20, 3, 16, 4
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(dx, du, dy, M, S, B, T, R, 1.0, (10.0, 0.5), seed=21, strong=True,
                                                        zeta_pos=zeta_pos)
    u, y, eps_b, z_b, eps_f = _f32_exact(u, y, eps_b, z_b, eps_f)
    res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, True)
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, True, 12)
    grads = eng.get_grads()
    xf, _ = eng.export_states(yd)
    torch.cuda.synchronize()
    assert np.isfinite(float(out["loss"])) and bool(torch.isfinite(xf).all())
    assert all(np.all(np.isfinite(grads[k])) for k in O.PARAM_NAMES)
    loss_err = abs(float(out["loss"]) - float(res.loss.detach())) / abs(float(res.loss.detach()))
    worst = max(rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES)
    with capsys.disabled():
        print(f"\n[ill-conditioned] zeta_pos={zeta_pos}: cond(K_zz) f={_cond_kzz(params, 'f'):.2e} b={_cond_kzz(params, 'b'):.2e}  "
              f"loss rel err {loss_err:.2e}  worst gradient rel err {worst:.2e}")
    # The float64 path is as accurate as float64 allows.  At cond ~1e9 that is itself limited: the two CPU restatements
    # (Cholesky solves + autograd vs explicit inverse + hand-derived adjoint) differ from each other by up to 1e-1 in
    # the kernel-variance gradient.  Bound per tensor: 1e-4 + 3 x that measured float64 floor.
    _, g2 = KM.elbo_value_and_grad(cfg, {k: v.numpy() for k, v in params.items()}, u, y, eps_b, z_b, eps_f, True)
    floor = {k: rel_inf(g2[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    eng64, out64, _ = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, True, 128)
    g64 = eng64.get_grads()
    loss64 = abs(float(out64["loss"]) - float(res.loss.detach())) / abs(float(res.loss.detach()))
    err64 = {k: rel_inf(g64[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    with capsys.disabled():
        print(f"[ill-conditioned] zeta_pos={zeta_pos}: float64 path loss rel err {loss64:.2e}  worst gradient rel err "
              f"{max(err64.values()):.2e}  (float64 floor between the CPU restatements {max(floor.values()):.2e})")
    assert loss64 < 1e-6
    # The kernel-variance gradient is one scalar left over from the cancellation of M^2 terms; at cond >= 1e9 the two
    # CPU float64 evaluations themselves disagree on it by 1e-2 .. 1e-1 depending on the BLAS build, so there it is
    # only required to be finite and of the right order.
    loose = ("f.variance_unc", "b.variance_unc") if zeta_pos < 2.0 else ()
    bad = {k: (v, floor[k]) for k, v in err64.items() if not v < (0.5 if k in loose else TOL + 3 * floor[k])}
    assert not bad, bad
    if zeta_pos == 2.0:
        assert max(err64.values()) < 1e-5        # cond 1.4e5: far inside float64
