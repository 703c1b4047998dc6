"""CPU oracle for the CBF-SSM sampled-ELBO hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, not the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``cbf_ssm_b200/`` does.

PARITY PINNED BY EXECUTING THE REFERENCE'S SOURCE.  The reference (silvanmelchior/CBF-SSM)
builds a TensorFlow 1.8 graph; TensorFlow cannot be imported in this image (Python 3.12,
no network) and the reference ships no tests, golden vectors or seeds.  This oracle is a
float64 op-for-op *restatement* of the reference files cited on each function, with every
``tf.random_normal`` draw turned into an explicit input.  It is pinned by
``tests/golden/ref_*.npz``: fixtures written by ``oracle/run_reference.py``, which imports the
reference's UNMODIFIED ``cbfssm/model/{tf_transform,gp_tf,base_model,cbfssm,cbfssmhalf}.py``
from the checkout and executes them under ``oracle/tf_shim`` (an eager float64 stand-in for
the ~60 TensorFlow symbols they use).  ``tests/test_oracle.py`` asserts this restatement
reproduces those fixtures to 1e-10 (loss terms, all 12 gradients, states, moments, first Adam
step, and the resample schedule the reference's own ``tf.cond`` predicates took) on the six
named configurations and four harder cases.  What remains restated rather than executed is
TensorFlow's *library* behaviour (softplus, cholesky, triangular solve,
MVN log-prob / KL formulas, moments, Adam), written out in the shim; it has not been compared
with a TensorFlow binary.  Further checks: finite differences of the autograd gradients, an
independent NumPy restatement in restructured algebra (``oracle/kernel_math.py``) and closed
forms.

All arithmetic is float64 (the reference default, cbfssm/model/base_model.py:8)
on PyTorch-CPU so that autograd supplies the reference gradients
(``tf.gradients`` through the three ``tf.while_loop`` s, cbfssm/model/cbfssm.py:274-275).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

DT = torch.float64

# Creation order of the 12 trainable tensors in the reference graph
# (cbfssm/model/gp_tf.py:112-127 inside cbfssm/model/cbfssm.py:30-54).
PARAM_NAMES = (
    "f.zeta_pos", "f.zeta_mean", "f.zeta_var_unc", "f.variance_unc", "f.lengthscales_unc",
    "b.zeta_pos", "b.zeta_mean", "b.zeta_var_unc", "b.variance_unc", "b.lengthscales_unc",
    "var_x_unc", "var_y_unc",
)


# --------------------------------------------------------------------------
# cbfssm/model/tf_transform.py
# --------------------------------------------------------------------------
def positive_backward(y):
    """Inverse of ``positive_forward`` used for initialisation (tf_transform.py:13-16)."""
    y = np.asarray(y, dtype=np.float64)
    if np.any(y <= 1e-10):
        raise AssertionError("Input to backward transformation should be greater 1e-10")
    with np.errstate(over="ignore"):
        res = np.log(np.exp(y - 1e-10) - 1.0)
    return np.where(y > 35, y - 1e-10, res)


def positive_forward(x: torch.Tensor) -> torch.Tensor:
    """softplus(x) + 1e-10 (tf_transform.py:19-21)."""
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9) + 1e-10


# --------------------------------------------------------------------------
# Configuration: the dict of run/template.py:19-40 plus the dataset dims
# --------------------------------------------------------------------------
@dataclass
class OracleConfig:
    dim_x: int
    dim_u: int
    dim_y: int
    ind_pnt_num: int
    samples: int
    recog_len: int
    k_factor: float = 1.0
    loss_factors: tuple = (10.0, 0.0)
    zeta_pos: float = 2.0
    zeta_mean: float = 0.01
    zeta_var: float = 1e-4
    var_x: object = None  # array [dim_x]
    var_y: object = None  # array [dim_x]  (sic, cbfssm.py:53, run/template.py:37)
    gp_var: float = 0.01
    gp_len: float = 1.0

    def __post_init__(self):
        if self.var_x is None:
            self.var_x = np.full(self.dim_x, 0.01)
        if self.var_y is None:
            self.var_y = np.full(self.dim_x, 1.0)
        self.var_x = np.asarray(self.var_x, dtype=np.float64)
        self.var_y = np.asarray(self.var_y, dtype=np.float64)

    @property
    def dim_h(self):
        return self.dim_x - self.dim_y

    @property
    def dim_in(self):
        return self.dim_x + self.dim_u


def init_params(cfg: OracleConfig, seed: int) -> Dict[str, torch.Tensor]:
    """The 12 raw (unconstrained) tensors, drawn in the reference's creation order.

    gp_tf.py:112-127 (zeta_pos ~ U(-zp, zp); zeta_mean = zm * U(0,1); zeta_var_unc,
    kern variance_unc, kern lengthscales_unc via ``backward``), for gp_f then gp_b
    (cbfssm.py:30-48), then var_x_unc, var_y_unc (cbfssm.py:51-54).  The reference
    uses the unseeded global NumPy RNG; here a seeded RandomState issues the same
    calls in the same order.
    """
    rs = np.random.RandomState(seed)
    M, din = cfg.ind_pnt_num, cfg.dim_in
    out: Dict[str, torch.Tensor] = {}
    for tag, dout in (("f", cfg.dim_x), ("b", cfg.dim_h)):
        out[f"{tag}.zeta_pos"] = rs.uniform(low=-cfg.zeta_pos, high=cfg.zeta_pos, size=(M, din))
        out[f"{tag}.zeta_mean"] = cfg.zeta_mean * rs.rand(M, dout)
        out[f"{tag}.zeta_var_unc"] = positive_backward(cfg.zeta_var * np.ones((M, dout)))
        out[f"{tag}.variance_unc"] = positive_backward(cfg.gp_var).reshape(())
        out[f"{tag}.lengthscales_unc"] = positive_backward(np.asarray([cfg.gp_len] * din))
    out["var_x_unc"] = positive_backward(cfg.var_x)
    out["var_y_unc"] = positive_backward(cfg.var_y)
    return {k: torch.tensor(np.asarray(v), dtype=DT) for k, v in out.items()}


# --------------------------------------------------------------------------
# cbfssm/model/gp_tf.py
# --------------------------------------------------------------------------
class RBF:
    """ARD squared-exponential kernel (gp_tf.py:20-49)."""

    def __init__(self, variance_unc, lengthscales_unc):
        self.variance = positive_forward(variance_unc)
        self.lengthscales = positive_forward(lengthscales_unc)

    def square_dist(self, X, X2=None):
        # gp_tf.py:33-43 -- expanded form, no clamp at zero
        X = X / self.lengthscales
        Xs = torch.sum(X * X, dim=1)
        if X2 is None:
            return -2.0 * (X @ X.T) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
        X2 = X2 / self.lengthscales
        X2s = torch.sum(X2 * X2, dim=1)
        return -2.0 * (X @ X2.T) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)

    def Kdiag(self, X):
        # gp_tf.py:45-46
        return self.variance.reshape(()).expand(X.shape[0])

    def K(self, X, X2=None):
        # gp_tf.py:48-49
        return self.variance * torch.exp(-0.5 * self.square_dist(X, X2))


def jitter_cholesky(mat, jitter=1e-8):
    """gp_tf.py:52-65 (always float64 here, so the cast is the identity)."""
    mat = mat + jitter * torch.eye(mat.shape[0], dtype=mat.dtype)
    return torch.linalg.cholesky(mat)


class GPModel:
    """Sparse GP with un-whitened, diagonal q(u) (gp_tf.py:103-172)."""

    def __init__(self, zeta_pos, zeta_mean, zeta_var_unc, variance_unc, lengthscales_unc):
        self.zeta_pos = zeta_pos
        self.zeta_mean = zeta_mean
        self.zeta_var = positive_forward(zeta_var_unc)       # gp_tf.py:122
        self.zeta_std = torch.sqrt(self.zeta_var)            # gp_tf.py:123
        self.kern = RBF(variance_unc, lengthscales_unc)      # gp_tf.py:125-127
        self.num_points, self.out_dim = zeta_mean.shape
        self.cholesky = jitter_cholesky(self.kern.K(zeta_pos), 1e-8)   # gp_tf.py:129-130

    def predict(self, Xnew):
        """gp_tf.py:132-161; returns (fmean [N, Dout], fvar [N, Dout])."""
        Kmn = self.kern.K(self.zeta_pos, Xnew)                                       # :134
        A = torch.linalg.solve_triangular(self.cholesky, Kmn, upper=False)           # :137
        fvar = self.kern.Kdiag(Xnew) - torch.sum(A * A, dim=0)                       # :140
        fvar = fvar.unsqueeze(0).expand(self.out_dim, -1)                            # :141-142
        A = torch.linalg.solve_triangular(self.cholesky.T, A, upper=True)            # :145
        fmean = A.T @ self.zeta_mean                                                 # :148
        LTA = A.unsqueeze(0) * self.zeta_std.T.unsqueeze(2)                          # :152
        fvar = fvar + torch.sum(LTA * LTA, dim=1)                                    # :159
        return fmean, fvar.T                                                         # :161

    def prior_kl(self):
        """gp_tf.py:163-172 with TF-1.8's kl(MVNLinearOperator a || b) written out:
        log|b.scale| - log|a.scale| + 0.5 (-n + ||b^-1 a.scale||_F^2 + ||b^-1 (mu_b - mu_a)||^2),
        batched over the Dout output functions, then reduce_sum.
        """
        L = self.cholesky
        n = self.num_points
        total = torch.zeros((), dtype=DT)
        logdet_b = torch.sum(torch.log(torch.abs(torch.diagonal(L))))
        for d in range(self.out_dim):
            std = self.zeta_std[:, d]
            b_inv_a = torch.linalg.solve_triangular(L, torch.diag(std), upper=False)
            b_inv_m = torch.linalg.solve_triangular(L, (-self.zeta_mean[:, d]).reshape(-1, 1), upper=False)
            logdet_a = torch.sum(torch.log(torch.abs(std)))
            total = total + (logdet_b - logdet_a
                             + 0.5 * (-n + torch.sum(b_inv_a * b_inv_a) + torch.sum(b_inv_m * b_inv_m)))
        return total


# --------------------------------------------------------------------------
# cbfssm/model/cbfssm.py
# --------------------------------------------------------------------------
def backward_schedule(run: int, t: int, recog_len: int):
    """(resample, write) flags of cbfssm.py:123-128."""
    R = recog_len
    if run == 0:
        return ((t + 1) % (2 * R) == 0), (t % (2 * R) < R)
    return ((t + R + 1) % (2 * R) == 0), (t % (2 * R) >= R)


@dataclass
class OracleResult:
    loss: torch.Tensor
    elbo: torch.Tensor
    loglik: torch.Tensor
    kl_x: torch.Tensor
    entropy: torch.Tensor
    kl_z_f: torch.Tensor
    kl_z_b: torch.Tensor
    x_final: torch.Tensor      # [B, T, S, dx]
    y_tilde: torch.Tensor      # [B, T, S, dx]
    pred_mean: torch.Tensor    # [B, T, dy]
    pred_var: torch.Tensor
    internal_mean: torch.Tensor
    internal_var: torch.Tensor
    extras: dict = field(default_factory=dict)


def build_gps(params):
    gp_f = GPModel(params["f.zeta_pos"], params["f.zeta_mean"], params["f.zeta_var_unc"],
                   params["f.variance_unc"], params["f.lengthscales_unc"])
    gp_b = GPModel(params["b.zeta_pos"], params["b.zeta_mean"], params["b.zeta_var_unc"],
                   params["b.variance_unc"], params["b.lengthscales_unc"])
    return gp_f, gp_b


def elbo(cfg: OracleConfig, params: Dict[str, torch.Tensor], u, y, eps_b, z_b, eps_f,
         condition: bool = True) -> OracleResult:
    """One execution of the reference graph on one minibatch.

    u [B,T,du], y [B,T,dy]; draws (Appendix B of SURVEY.md):
    eps_b [2,T,B,S] (cbfssm.py:149), z_b [2,T,B,S] read only at resample steps
    (cbfssm.py:134), eps_f [T-1,B,S] (cbfssm.py:209).  Each scalar draw is tiled
    over the output dimension exactly like ``tf.tile(..., [1,1,dim])``.
    """
    u = torch.as_tensor(u, dtype=DT)
    y = torch.as_tensor(y, dtype=DT)
    eps_b = torch.as_tensor(eps_b, dtype=DT)
    z_b = torch.as_tensor(z_b, dtype=DT)
    eps_f = torch.as_tensor(eps_f, dtype=DT)
    B, T, du = u.shape
    S, dx, dy, dh, R = cfg.samples, cfg.dim_x, cfg.dim_y, cfg.dim_h, cfg.recog_len
    assert du == cfg.dim_u and y.shape == (B, T, dy)
    assert eps_b.shape == (2, T, B, S) and z_b.shape == (2, T, B, S) and eps_f.shape == (T - 1, B, S)

    gp_f, gp_b = build_gps(params)
    var_x = positive_forward(params["var_x_unc"])      # cbfssm.py:52
    var_y = positive_forward(params["var_y_unc"])      # cbfssm.py:54

    # cbfssm.py:69-82 -- time-major, tiled over the particles
    u_arr = u.permute(1, 0, 2).unsqueeze(2).expand(T, B, S, du)
    y_arr = y.permute(1, 0, 2).unsqueeze(2).expand(T, B, S, dy)

    # ---- backward message, two runs (cbfssm.py:84-158) ----
    y2 = [None] * T
    prob = [None] * T
    log_2pie = math.log(2.0 * math.pi * math.e)
    for run in (0, 1):
        h = torch.zeros((B, S, dh), dtype=DT)                                  # :106
        for t in range(T - 1, -1, -1):
            resample, write = backward_schedule(run, t, R)
            hidden = z_b[run, t].unsqueeze(-1).expand(B, S, dh) if resample else h   # :133-136
            in_t = torch.cat((hidden, u_arr[t], y_arr[t]), dim=2)                    # :137
            fmean, fvar = gp_b.predict(in_t.reshape(B * S, dx + du))                 # :140-141
            fmean = fmean.reshape(B, S, dh) + in_t[:, :, :dh]                        # :143,145
            fvar = fvar.reshape(B, S, dh) + var_x[:dh]                               # :144,146
            eps = eps_b[run, t].unsqueeze(-1).expand(B, S, dh)                       # :149
            out = fmean + eps * torch.sqrt(fvar)                                     # :150
            if write:
                assert y2[t] is None
                y2[t] = out                                                          # :151
                prob[t] = 0.5 * torch.sum(log_2pie + torch.log(fvar))                # :154-156
            h = out                                                                  # :158
    y2_arr = torch.stack(y2, dim=0).permute(1, 0, 2, 3)                              # :95  [B,T,S,dh]
    out_dub = y.unsqueeze(2).expand(B, T, S, dy)                                     # :96
    y_tilde = torch.cat((out_dub, y2_arr), dim=3)                                    # :97
    entropy = torch.sum(torch.stack(prob))                                           # :99

    # ---- forward conditional rollout (cbfssm.py:160-237) ----
    xs = [y_tilde[:, 0]]                                                             # :168-169
    kls = []
    yt_arr = y_tilde.permute(1, 0, 2, 3)                                             # :173
    for t in range(T - 1):
        x_t = xs[t]
        in_t = torch.cat((x_t, u_arr[t]), dim=2)                                     # :197
        fmean, fvar = gp_f.predict(in_t.reshape(B * S, du + dx))                     # :200-201
        fmean = fmean.reshape(B, S, dx) + in_t[:, :, :dx]                            # :203,205
        fvar = fvar.reshape(B, S, dx) + var_x                                        # :204,206
        eps = eps_f[t].unsqueeze(-1).expand(B, S, dx)                                # :209
        var_y_t = var_y.reshape(1, 1, dx) + (cfg.k_factor - 1.0) * fvar              # :212-214
        y_diff = yt_arr[t + 1] - fmean                                               # :215
        s = var_y_t + fvar                                                           # :216
        k = fvar * torch.reciprocal(s)                                               # :217
        mu = fmean + k * y_diff                                                      # :218
        sig = (1.0 - k) ** 2 * fvar + k ** 2 * var_y_t                               # :219-220
        x_c = mu + eps * torch.sqrt(sig)                                             # :221
        x_nc = fmean + eps * torch.sqrt(fvar)                                        # :224
        do_cond = bool(condition) or (t < R - 1)                                     # :227
        xs.append(x_c if do_cond else x_nc)                                          # :228-229
        kl_reg = torch.log(fvar) - torch.log(sig) + (sig + (mu - fmean) ** 2) / fvar - 1.0   # :232
        kls.append(torch.sum(kl_reg) * (0.5 if do_cond else 0.0))                    # :233-235
    x_final = torch.stack(xs, dim=0).permute(1, 0, 2, 3)                             # :181 [B,T,S,dx]
    y_final = x_final[..., :dy]                                                      # :182
    kl_x = torch.sum(torch.stack(kls)) if kls else torch.zeros((), dtype=DT)         # :183

    # ---- loss (cbfssm.py:239-262) ----
    sd = torch.sqrt(var_y[:dy]).reshape(1, 1, 1, dy)                                 # :245-248
    obs = y.unsqueeze(2).expand(B, T, S, dy)                                         # :249
    zs = (obs - y_final) / sd
    log_probs = -0.5 * torch.sum(zs * zs, dim=-1) - torch.sum(torch.log(sd)) \
        - 0.5 * dy * math.log(2.0 * math.pi)                                         # :250 MVNDiag.log_prob
    loglik = torch.sum(log_probs)                                                    # :251
    kl_z_f = gp_f.prior_kl()                                                         # :254
    kl_z_b = gp_b.prior_kl()                                                         # :255
    divisor = 1.0 / float(S)                                                         # :257
    lf = cfg.loss_factors
    elbo_v = (loglik * lf[0] * divisor - kl_x * lf[0] * divisor
              + entropy * lf[1] * divisor - kl_z_f - kl_z_b)                         # :258-261
    loss = -elbo_v                                                                   # :262

    # ---- prediction moments (cbfssm.py:264-271) ----
    pred_mean = y_final.mean(dim=2)
    pred_var = y_final.var(dim=2, unbiased=False) + var_y[:dy]                       # :267-268
    internal_mean = x_final.mean(dim=2)
    internal_var = x_final.var(dim=2, unbiased=False)                                # :269

    return OracleResult(loss=loss, elbo=elbo_v, loglik=loglik, kl_x=kl_x, entropy=entropy,
                        kl_z_f=kl_z_f, kl_z_b=kl_z_b, x_final=x_final, y_tilde=y_tilde,
                        pred_mean=pred_mean, pred_var=pred_var,
                        internal_mean=internal_mean, internal_var=internal_var)


def loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, condition=True):
    """loss plus d loss / d (each of the 12 raw tensors) -- what
    ``AdamOptimizer.minimize`` differentiates (cbfssm.py:274-275)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    res = elbo(cfg, leaf, u, y, eps_b, z_b, eps_f, condition)
    grads = torch.autograd.grad(res.loss, [leaf[k] for k in PARAM_NAMES], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(leaf[k])) for k, g in zip(PARAM_NAMES, grads)}
    return res, gd


def draw_noise(B, S, T, seed):
    """Standard-normal draws with the shapes of SURVEY Appendix B, seeded."""
    g = np.random.default_rng(seed)
    eps_b = g.standard_normal((2, T, B, S))
    z_b = g.standard_normal((2, T, B, S))
    eps_f = g.standard_normal((T - 1, B, S))
    return eps_b, z_b, eps_f


def adam_step_tf(theta, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One TF-1.8 ``AdamOptimizer`` update (cbfssm.py:274): epsilon sits outside the
    bias-corrected root, unlike torch.optim.Adam.  ``step`` is 1-based."""
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad * grad
    lr_t = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    theta = theta - lr_t * m / (torch.sqrt(v) + eps)
    return theta, m, v
