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
