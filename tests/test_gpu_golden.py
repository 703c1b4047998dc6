"""CUDA path against fixtures produced by executing the reference's own, unmodified model source
(tests/golden/ref_*.npz, written by tests/golden/make_golden.py through oracle/run_reference.py)
at the named configurations of SURVEY.md 8(d) and at harder "strong" cases.  Tolerance 1e-4
relative (north_star): per tensor max|a-b| / max|b|, plus an element-wise check with an absolute
floor so that entries far below a tensor's maximum are not waved through."""
import os

import numpy as np
import pytest
import torch

from oracle import cbfssm_oracle as O      # only for PARAM_NAMES / shapes (checker side)
from oracle import cbfssmhalf_oracle as H
from tests.helpers import (HALF_REF_CASES, NAMED_CASES, STRONG_REF_CASES, elementwise_err, elementwise_ok, half_ref_case, named_case,
                           rel_inf, strong_ref_case)
from tests.test_gpu_parity import run_engine

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = ("f.variance_unc", "f.lengthscales_unc", "b.variance_unc", "b.lengthscales_unc", "var_x_unc", "var_y_unc")


def _check(name, case, flags):
    gold = np.load(os.path.join(GOLD, "ref_" + name + ".npz"))
    cfg, params, u, y, eps_b, z_b, eps_f, cond = case
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, flags)
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        ref, got = float(gold[k]), float(out[k])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref)
    xf, yt = eng.export_states(yd)
    pm, pv = eng.moments(xf, cfg.dim_y, eng.var_y)
    im, iv = eng.moments(xf, cfg.dim_x, None)
    torch.cuda.synchronize()
    assert rel_inf(xf.cpu().numpy()[:, ::15, ::10, :], gold["x_final_sample"]) < TOL
    assert rel_inf(yt.cpu().numpy()[:, ::15, ::10, :], gold["y_tilde_sample"]) < TOL
    assert rel_inf(pm.cpu().numpy(), gold["pred_mean"]) < TOL
    assert rel_inf(pv.cpu().numpy(), gold["pred_var"]) < TOL
    assert rel_inf(im.cpu().numpy(), gold["internal_mean"]) < TOL
    assert rel_inf(iv.cpu().numpy(), gold["internal_var"]) < TOL
    assert elementwise_ok(pv.cpu().numpy(), gold["pred_var"], TOL)          # variances: every entry
    grads = eng.get_grads()
    worst = {k: rel_inf(grads[k], gold["grad." + k]) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in worst.items() if not v < TOL}
    assert not bad, bad
    # the small tensors (kernel variance / lengthscales, noise terms): entry by entry
    bad = {k: elementwise_err(grads[k], gold["grad." + k]) for k in SMALL}
    print("elementwise", name, flags, {k: "%.1e" % v for k, v in bad.items()})
    bad = {k: v for k, v in bad.items() if not v < TOL}
    assert not bad, bad
    # one TF-Adam step on these gradients lands on the reference's updated variables (cbfssm.py:274-275).
    # Adam's first step is lr_t * g / (sqrt(.001) |g| + 1e-8): it forgets |g| unless |g| ~ 1e-8, so only entries
    # whose gradient is significant within its tensor are compared (the others amplify float32 noise by 1e8).
    eng.adam_step(0.01)
    torch.cuda.synchronize()
    new = eng.get_params()
    for k in O.PARAM_NAMES:
        ref_new, g = gold["adam." + k], np.abs(gold["grad." + k])
        sel = (g > 1e-3 * g.max()) & (g > 1e-4)
        if not sel.any():
            continue
        step = np.max(np.abs(ref_new - params[k].numpy())) + 1e-300
        diff = np.abs(np.asarray(new[k]).reshape(ref_new.shape) - ref_new)
        assert np.max(diff[sel]) <= 2e-3 * step, k


@pytest.mark.parametrize("flags", [12, 1], ids=["register_or_tensor", "cooperative"])
@pytest.mark.parametrize("name", list(NAMED_CASES))
def test_named_configuration_matches_reference_source(name, flags):
    _check(name, named_case(name), flags)


@pytest.mark.parametrize("flags", [12, 1], ids=["register_or_tensor", "cooperative"])
@pytest.mark.parametrize("name", list(STRONG_REF_CASES))
def test_strong_case_matches_reference_source(name, flags):
    _check(name, strong_ref_case(name), flags)


@pytest.mark.parametrize("flags", [12, 1], ids=["register_or_tensor", "cooperative"])
@pytest.mark.parametrize("name", list(HALF_REF_CASES))
def test_half_model_matches_reference_source(name, flags):
    """CBFSSMHALF: rollout through the C ABI, recognition network on the host side (torch float64), against the
    reference's cbfssmhalf.py executed unmodified: loss, states, moments, the 7 model gradients and -- through
    the x_0 adjoint -- the 6 recognition-network gradients."""
    from cbf_ssm_b200.engine import ElboEngine, ModelDims
    gold = np.load(os.path.join(GOLD, "ref_" + name + ".npz"))
    cfg, params, w, u, y, eps_f, cond, recog = half_ref_case(name)
    B, T, _ = u.shape
    eng = ElboEngine(ModelDims(cfg.dim_x, cfg.dim_u, cfg.dim_y, cfg.ind_pnt_num, cfg.samples, cfg.recog_len,
                               cfg.k_factor, tuple(cfg.loss_factors), half=True))
    eng.flags = flags
    eng.set_params({k: v.numpy() for k, v in params.items()})
    dev = eng.device
    f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    wl = {k: torch.tensor(v, requires_grad=True) for k, v in w.items()}
    x0 = H.recog_rnn(wl, u, y, cfg.recog_len) if recog == "rnn" else H.recog_output(y, cfg.dim_x)
    ud, yd = f32(u), f32(y)
    out = eng.forward(ud, yd, None, None, f32(eps_f.reshape(T - 1, B * cfg.samples)), cond, x0=f32(x0.detach().numpy()))
    eng.backward()
    torch.cuda.synchronize()
    for k in ("loss", "kl_x", "kl_z_f"):
        ref, got = float(gold[k]), float(out[k])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref)
    xf, _ = eng.export_states(yd)
    pm, pv = eng.moments(xf, cfg.dim_y, eng.var_y)
    torch.cuda.synchronize()
    assert rel_inf(xf.cpu().numpy(), gold["x_final"]) < TOL
    assert rel_inf(pm.cpu().numpy(), gold["pred_mean"]) < TOL
    assert rel_inf(pv.cpu().numpy(), gold["pred_var"]) < TOL
    grads = eng.get_grads()
    bad = {k: rel_inf(grads[k], gold["grad." + k]) for k in H.HALF_PARAM_NAMES}
    bad = {k: v for k, v in bad.items() if not v < TOL}
    assert not bad, bad
    if recog == "rnn":
        gw = torch.autograd.grad(x0, list(wl.values()), grad_outputs=eng.x0_bar.cpu().double())
        for k, g in zip(wl, gw):
            assert rel_inf(g.numpy(), gold["grad." + k]) < TOL, k
