"""CPU checks of the oracle.  The pin: fixtures written by executing the reference's own,
unmodified model source under an eager TensorFlow stand-in (tests/golden/ref_*.npz); beside it
identities, finite differences and a second restatement in restructured algebra."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import cbfssm_oracle as O
from oracle import kernel_math as KM
from tests.helpers import (HALF_REF_CASES, NAMED_CASES, STRONG_REF_CASES, half_ref_case, make_problem, named_case,
                           rel_inf, strong_ref_case)


def test_positive_transform_round_trip_and_guards():
    y = np.array([1e-9, 1e-4, 0.3, 1.0, 20.0, 36.0, 50.0])
    x = O.positive_backward(y)
    back = O.positive_forward(torch.tensor(x)).numpy()
    assert np.allclose(back, y, rtol=1e-9, atol=1e-12)
    assert x[-1] == pytest.approx(50.0 - 1e-10)          # y > 35 branch (tf_transform.py:16)
    with pytest.raises(AssertionError):
        O.positive_backward(np.array([1e-10]))


def test_rbf_matches_brute_force():
    g = np.random.default_rng(0)
    X, Z = g.standard_normal((7, 3)), g.standard_normal((5, 3))
    ell, var = np.array([0.5, 1.3, 2.0]), 0.7
    kern = O.RBF(torch.tensor(O.positive_backward(var)), torch.tensor(O.positive_backward(ell)))
    K = kern.K(torch.tensor(Z), torch.tensor(X)).numpy()
    ref = np.array([[var * math.exp(-0.5 * np.sum(((z - x) / ell) ** 2)) for x in X] for z in Z])
    assert np.allclose(K, ref, rtol=1e-12)
    assert np.allclose(kern.Kdiag(torch.tensor(X)).numpy(), var)


def _small_gp(seed=0, M=6, din=3, dout=2):
    g = np.random.default_rng(seed)
    Z = torch.tensor(g.uniform(-2, 2, (M, din)))
    m = torch.tensor(g.standard_normal((M, dout)))
    Su = torch.tensor(O.positive_backward(g.uniform(0.01, 0.3, (M, dout))))
    gp = O.GPModel(Z, m, Su, torch.tensor(O.positive_backward(0.8)), torch.tensor(O.positive_backward(np.full(din, 1.1))))
    return gp, g


def test_predict_matches_dense_gp_formulas():
    gp, g = _small_gp()
    X = torch.tensor(g.standard_normal((9, 3)))
    fm, fv = gp.predict(X)
    Kzz = gp.kern.K(gp.zeta_pos).numpy() + 1e-8 * np.eye(6)
    Kzx = gp.kern.K(gp.zeta_pos, X).numpy()
    A = np.linalg.solve(Kzz, Kzx)                                    # [M, N]
    assert np.allclose(fm.numpy(), A.T @ gp.zeta_mean.numpy(), rtol=1e-9)
    base = float(gp.kern.variance) - np.sum(Kzx * A, axis=0)
    ref = base[:, None] + (A ** 2).T @ gp.zeta_var.numpy()
    assert np.allclose(fv.numpy(), ref, rtol=1e-8)


def test_prior_kl_matches_torch_distributions():
    gp, _ = _small_gp(1)
    total = 0.0
    for d in range(gp.out_dim):
        q = torch.distributions.MultivariateNormal(gp.zeta_mean[:, d], covariance_matrix=torch.diag(gp.zeta_var[:, d]))
        p = torch.distributions.MultivariateNormal(torch.zeros(6, dtype=O.DT), scale_tril=gp.cholesky)
        total += float(torch.distributions.kl_divergence(q, p))
    assert float(gp.prior_kl()) == pytest.approx(total, rel=1e-10)


@pytest.mark.parametrize("T,R,live", [(100, 50, 150), (300, 50, 550), (250, 16, 484), (50, 16, 84), (64, 16, 112),
                                      (500, 16, 984), (500, 50, 950)])
def test_schedule_and_live_steps(T, R, live):
    """Literal flags of cbfssm.py:123-128; every t written exactly once; the chain table
    drops exactly the dead tail (SURVEY 8a note 5 table)."""
    writes = np.zeros(T, int)
    for run in (0, 1):
        for t in range(T):
            rs, wr = O.backward_schedule(run, t, R)
            assert rs == (((t + 1) if run == 0 else (t + R + 1)) % (2 * R) == 0)
            writes[t] += wr
    assert np.all(writes == 1)
    chains = KM.build_chains(T, R)
    assert sum(hi - lo + 1 for (_, hi, lo, _) in chains) == live
    # brute-force liveness: a step is live iff some later (lower t) step of its segment is written
    for run in (0, 1):
        seg_live = set()
        t = T - 1
        while t >= 0:
            seg = [t]
            t -= 1
            while t >= 0 and not O.backward_schedule(run, t, R)[0]:
                seg.append(t)
                t -= 1
            written = [s for s in seg if O.backward_schedule(run, s, R)[1]]
            if written:
                seg_live.update(s for s in seg if s >= min(written))
        table = set()
        for (r, hi, lo, _) in chains:
            if r == run:
                table.update(range(lo, hi + 1))
        assert table == seg_live


def test_oracle_gradients_match_finite_differences():
    cfg, params, u, y, eb, zb, ef = make_problem(3, 1, 1, 4, 2, 2, 7, 2, 3.0, (6.0, 1.0), seed=5, strong=True)
    _, gd = O.loss_and_grads(cfg, params, u, y, eb, zb, ef, True)
    rng = np.random.default_rng(0)
    for name in O.PARAM_NAMES:
        base = params[name]
        flat = base.reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                p2 = {k: v.clone() for k, v in params.items()}
                p2[name].reshape(-1)[idx] += sgn * h
                vals.append(float(O.elbo(cfg, p2, u, y, eb, zb, ef, True).loss))
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(gd[name].reshape(-1)[idx])
            assert fd == pytest.approx(an, rel=2e-5, abs=1e-6), (name, idx)


@pytest.mark.parametrize("cond", [True, False])
def test_kernel_algebra_restatement_agrees(cond):
    cfg, params, u, y, eb, zb, ef = make_problem(4, 2, 2, 7, 3, 2, 11, 3, 2.0, (10.0, 0.5), seed=2, strong=True)
    res, gd = O.loss_and_grads(cfg, params, u, y, eb, zb, ef, cond)
    out, g2 = KM.elbo_value_and_grad(cfg, {k: v.numpy() for k, v in params.items()}, u, y, eb, zb, ef, cond)
    assert out["loss"] == pytest.approx(float(res.loss.detach()), rel=1e-12)
    for k in O.PARAM_NAMES:
        assert rel_inf(g2[k], gd[k].numpy()) < 1e-10, k


def test_tf_adam_differs_from_torch_adam_only_in_epsilon_placement():
    th = torch.tensor([0.3, -1.2], dtype=O.DT)
    g = torch.tensor([0.5, -2.0], dtype=O.DT)
    t1, m, v = O.adam_step_tf(th, g, torch.zeros(2, dtype=O.DT), torch.zeros(2, dtype=O.DT), 1, 0.01)
    # first step: m_hat/sqrt(v_hat) = sign(g); TF: lr*sqrt(1-b2)/(1-b1) * m/(sqrt(v)+eps)
    lr_t = 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9)
    ref = th - lr_t * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-8)
    assert torch.allclose(t1, ref, rtol=1e-14)


# --------------------------------------------------------------------------------------
# The pin: fixtures produced by executing the reference's unmodified source
# (tests/golden/make_golden.py -> oracle/run_reference.py under oracle/tf_shim)
# --------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(__file__), "golden")
PIN_TOL = 1e-10          # float64 both sides; differences are summation order only


def _check_against_reference_fixture(name, case):
    gold = np.load(os.path.join(GOLD, "ref_" + name + ".npz"))
    cfg, params, u, y, eb, zb, ef, cond = case
    res, gd = O.loss_and_grads(cfg, params, u, y, eb, zb, ef, cond)
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        assert float(getattr(res, k).detach()) == pytest.approx(float(gold[k]), rel=PIN_TOL, abs=1e-9), k
    for k in O.PARAM_NAMES:
        assert rel_inf(gd[k].numpy(), gold["grad." + k]) < PIN_TOL, k
        # element-wise, too: the reference gradient is reproduced entry by entry
        assert np.allclose(gd[k].numpy(), gold["grad." + k], rtol=1e-7, atol=1e-9 * np.max(np.abs(gold["grad." + k])) + 1e-300), k
    for k in ("pred_mean", "pred_var", "internal_mean", "internal_var"):
        assert rel_inf(getattr(res, k).detach().numpy(), gold[k]) < PIN_TOL, k
    assert rel_inf(res.x_final.detach().numpy()[:, ::15, ::10, :], gold["x_final_sample"]) < PIN_TOL
    assert rel_inf(res.y_tilde.detach().numpy()[:, ::15, ::10, :], gold["y_tilde_sample"]) < PIN_TOL
    # the resample steps the reference's tf.cond predicates took == the restated schedule
    T, R = u.shape[1], cfg.recog_len
    mine = [(run, t) for run in (0, 1) for t in range(T - 1, -1, -1) if O.backward_schedule(run, t, R)[0]]
    assert [tuple(r) for r in gold["resampled_at"].tolist()] == mine
    # TF-Adam first step on the reference's gradients (cbfssm.py:274-275)
    for k in O.PARAM_NAMES:
        new, _, _ = O.adam_step_tf(params[k], gd[k], torch.zeros_like(params[k]), torch.zeros_like(params[k]), 1, 0.01)
        assert np.allclose(new.numpy(), gold["adam." + k], rtol=1e-9, atol=1e-12), k


@pytest.mark.parametrize("name", list(NAMED_CASES))
def test_oracle_reproduces_reference_source_on_named_configurations(name):
    _check_against_reference_fixture(name, named_case(name))


@pytest.mark.parametrize("name", list(STRONG_REF_CASES))
def test_oracle_reproduces_reference_source_on_strong_cases(name):
    _check_against_reference_fixture(name, strong_ref_case(name))


@pytest.mark.parametrize("name", list(HALF_REF_CASES))
def test_half_oracle_reproduces_reference_source(name):
    from oracle import cbfssmhalf_oracle as H
    gold = np.load(os.path.join(GOLD, "ref_" + name + ".npz"))
    cfg, params, w, u, y, ef, cond, recog = half_ref_case(name)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    wl = {k: torch.tensor(v, requires_grad=True) for k, v in w.items()}
    x0 = H.recog_rnn(wl, u, y, cfg.recog_len) if recog == "rnn" else H.recog_output(y, cfg.dim_x)
    res = H.elbo_half(cfg, leaf, u, y, x0, ef, cond)
    names = list(H.HALF_PARAM_NAMES) + (list(wl) if recog == "rnn" else [])
    leaves = [leaf[k] for k in H.HALF_PARAM_NAMES] + (list(wl.values()) if recog == "rnn" else [])
    grads = torch.autograd.grad(res["loss"], leaves, allow_unused=True)
    for k in ("loss", "kl_x", "kl_z_f"):
        assert float(res[k].detach()) == pytest.approx(float(gold[k]), rel=PIN_TOL), k
    for k, g, v in zip(names, grads, leaves):
        g = torch.zeros_like(v) if g is None else g
        assert rel_inf(g.numpy(), gold["grad." + k]) < PIN_TOL, k
    for k in ("x_final", "pred_mean", "pred_var"):
        assert rel_inf(res[k].detach().numpy(), gold[k]) < PIN_TOL, k


def test_reference_source_runs_here_and_matches_its_fixture():
    """Live run of the unmodified reference (only where the checkout exists, i.e. the build
    container): the committed fixture is what the reference source produces today."""
    from oracle import run_reference as RR
    if not RR.available():
        pytest.skip("reference checkout not present on this machine")
    name = "strong_m20"
    cfg, params, u, y, eb, zb, ef, cond = strong_ref_case(name)
    ref = RR.run_cbfssm(cfg, [params[k].numpy() for k in O.PARAM_NAMES], u, y, eb, zb, ef, cond)
    gold = np.load(os.path.join(GOLD, "ref_" + name + ".npz"))
    # (not bit-for-bit: the BLAS summation order may differ with the thread count of the machine running this)
    assert float(ref["loss"]) == pytest.approx(float(gold["loss"]), rel=1e-12)
    for k, g in zip(O.PARAM_NAMES, ref["grads"]):
        assert rel_inf(g.reshape(gold["grad." + k].shape), gold["grad." + k]) < 1e-11, k
    # every draw the graph asked for came from the reference's own loop bodies
    bodies = {b for b, _, _, _ in ref["draw_log"]}
    assert bodies == {"_backward_body", "_forward_body"}


# --------------------------------------------------------------------------------------
# CBFSSMHALF oracle (oracle/cbfssmhalf_oracle.py)
# --------------------------------------------------------------------------------------
def test_half_oracle_gradients_match_finite_differences():
    from oracle import cbfssmhalf_oracle as H
    cfg, _, u, y, _, _, ef = make_problem(4, 2, 2, 5, 2, 2, 7, 2, 3.0, (6.0, 0.0), seed=8, strong=True)
    params = H.init_params_half(cfg, 8)
    x0 = torch.tensor(np.random.default_rng(1).standard_normal((2, 4)))
    _, gd = H.loss_and_grads_half(cfg, params, u, y, x0, ef, True)
    rng = np.random.default_rng(0)
    h = 1e-6
    for name in H.HALF_PARAM_NAMES:
        flat = params[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(2, flat.numel()), replace=False):
            vals = []
            for sgn in (+1, -1):
                p2 = {k: v.clone() for k, v in params.items()}
                p2[name].reshape(-1)[idx] += sgn * h
                vals.append(float(H.elbo_half(cfg, p2, u, y, x0, ef, True)["loss"]))
            assert (vals[0] - vals[1]) / (2 * h) == pytest.approx(float(gd[name].reshape(-1)[idx]), rel=2e-5, abs=1e-6)
    for idx in (0, 5):
        vals = []
        for sgn in (+1, -1):
            x2 = x0.clone()
            x2.reshape(-1)[idx] += sgn * h
            vals.append(float(H.elbo_half(cfg, params, u, y, x2, ef, True)["loss"]))
        assert (vals[0] - vals[1]) / (2 * h) == pytest.approx(float(gd["x0"].reshape(-1)[idx]), rel=2e-5, abs=1e-6)


def test_half_unconditioned_dims_have_zero_kl_and_follow_the_prior():
    """cbfssmhalf.py:144-149: dims >= dim_y get k = 0, so with dim_y conditioned dims masked out the step is
    the GP prior step; with condition=False after recog_len-1 steps KL_x stops growing."""
    from oracle import cbfssmhalf_oracle as H
    cfg, _, u, y, _, _, ef = make_problem(3, 1, 1, 4, 2, 1, 6, 2, 1.0, (10.0, 0.0), seed=3, strong=True)
    params = H.init_params_half(cfg, 3)
    x0 = H.recog_output(y, 3)
    assert x0.shape == (1, 3) and float(x0[0, 1]) == 0.0 and float(x0[0, 0]) == pytest.approx(float(y[0, 0, 0]))
    a = H.elbo_half(cfg, params, u, y, x0, ef, True)
    b = H.elbo_half(cfg, params, u, y, x0, ef, False)
    assert float(b["kl_x"]) < float(a["kl_x"])          # only t < recog_len - 1 contribute when not conditioning


def test_tf_gru_cell_restatement_one_step_by_hand():
    from oracle import cbfssmhalf_oracle as H
    g = torch.Generator().manual_seed(0)
    d, Hn = 2, 16
    w = {"gates_kernel": torch.randn(d + Hn, 2 * Hn, generator=g, dtype=O.DT), "gates_bias": torch.ones(2 * Hn, dtype=O.DT),
         "candidate_kernel": torch.randn(d + Hn, Hn, generator=g, dtype=O.DT), "candidate_bias": torch.zeros(Hn, dtype=O.DT),
         "dense_kernel": torch.eye(Hn, 3, dtype=O.DT), "dense_bias": torch.zeros(3, dtype=O.DT)}
    u = torch.randn(1, 4, 1, generator=g, dtype=O.DT)
    y = torch.randn(1, 4, 1, generator=g, dtype=O.DT)
    out = H.recog_rnn(w, u, y, 1)                       # one step from h = 0 on [u_0, y_0]
    x = torch.cat((u[:, 0], y[:, 0]), dim=1)
    gates = torch.sigmoid(x @ w["gates_kernel"][:d] + 1.0)
    z = gates[:, Hn:]
    c = torch.tanh(x @ w["candidate_kernel"][:d])       # r * h = 0
    h = (1 - z) * c
    assert torch.allclose(out, h[:, :3], rtol=1e-12)
