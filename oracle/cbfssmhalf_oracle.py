"""CPU oracle for ``CBFSSMHALF`` (cbfssm/model/cbfssmhalf.py).  TEST INFRASTRUCTURE ONLY.

Same status as ``cbfssm_oracle.py``: a float64 PyTorch-CPU restatement of the TensorFlow-1.8 graph with
injected normal draws, pinned by fixtures from the reference's unmodified ``cbfssmhalf.py`` executed under
``oracle/tf_shim`` (``tests/golden/ref_half_*.npz``; the GRU cell / dense layer are TF library code restated
in the shim).
CBFSSMHALF has no backward-message GP: x_0 comes from a recognition model (zero-padded first output,
or a GRU(16)+dense over the reversed first ``recog_len`` steps, cbfssmhalf.py:64-95) and the forward
step conditions only the first ``dim_y`` state dimensions (cbfssmhalf.py:144-156); ``var_y`` has
length ``dim_y``; the ELBO has no entropy term and one inducing KL (cbfssmhalf.py:176-193).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .cbfssm_oracle import DT, GPModel, OracleConfig, positive_backward, positive_forward

HALF_PARAM_NAMES = ("f.zeta_pos", "f.zeta_mean", "f.zeta_var_unc", "f.variance_unc", "f.lengthscales_unc",
                    "var_x_unc", "var_y_unc")


def init_params_half(cfg: OracleConfig, seed: int) -> Dict[str, torch.Tensor]:
    """gp_f (gp_tf.py:112-127) then var_x_unc [dx], var_y_unc [dy] (cbfssmhalf.py:34-37)."""
    rs = np.random.RandomState(seed)
    M, din = cfg.ind_pnt_num, cfg.dim_in
    out = {
        "f.zeta_pos": rs.uniform(low=-cfg.zeta_pos, high=cfg.zeta_pos, size=(M, din)),
        "f.zeta_mean": cfg.zeta_mean * rs.rand(M, cfg.dim_x),
        "f.zeta_var_unc": positive_backward(cfg.zeta_var * np.ones((M, cfg.dim_x))),
        "f.variance_unc": positive_backward(cfg.gp_var).reshape(()),
        "f.lengthscales_unc": positive_backward(np.asarray([cfg.gp_len] * din)),
        "var_x_unc": positive_backward(cfg.var_x),
        "var_y_unc": positive_backward(np.asarray(cfg.var_y)[:cfg.dim_y]),
    }
    return {k: torch.tensor(np.asarray(v), dtype=DT) for k, v in out.items()}


def recog_output(y, dim_x):
    """x_0 = [y_0, 0]  (cbfssmhalf.py:76-80), per sequence [B, dx]."""
    y = torch.as_tensor(y, dtype=DT)
    B, _, dy = y.shape
    return torch.cat((y[:, 0, :], torch.zeros((B, dim_x - dy), dtype=DT)), dim=1)


def recog_rnn(w: Dict[str, torch.Tensor], u, y, recog_len):
    """TF-1.8 ``GRUCell(16)`` run over the reversed first ``recog_len`` steps of [u, y], then a dense
    layer (cbfssmhalf.py:82-92).  TF's cell: [r, z] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc);
    h' = z*h + (1-z)*c  (reset gate applied *before* the candidate matmul, unlike torch.nn.GRU)."""
    u = torch.as_tensor(u, dtype=DT)
    y = torch.as_tensor(y, dtype=DT)
    xy = torch.cat((u, y), dim=2)[:, :recog_len, :]
    xy = torch.flip(xy, dims=[1])
    B = xy.shape[0]
    H = w["gates_kernel"].shape[1] // 2
    h = torch.zeros((B, H), dtype=DT)
    for t in range(xy.shape[1]):
        g = torch.sigmoid(torch.cat((xy[:, t], h), dim=1) @ w["gates_kernel"] + w["gates_bias"])
        r, z = g[:, :H], g[:, H:]
        c = torch.tanh(torch.cat((xy[:, t], r * h), dim=1) @ w["candidate_kernel"] + w["candidate_bias"])
        h = z * h + (1.0 - z) * c
    return h @ w["dense_kernel"] + w["dense_bias"]


def elbo_half(cfg: OracleConfig, params, u, y, x0, eps_f, condition=True):
    """One graph execution given x_0 [B, dx] (output of the recognition model)."""
    u = torch.as_tensor(u, dtype=DT)
    y = torch.as_tensor(y, dtype=DT)
    eps_f = torch.as_tensor(eps_f, dtype=DT)
    B, T, du = u.shape
    S, dx, dy, R = cfg.samples, cfg.dim_x, cfg.dim_y, cfg.recog_len
    gp_f = GPModel(params["f.zeta_pos"], params["f.zeta_mean"], params["f.zeta_var_unc"],
                   params["f.variance_unc"], params["f.lengthscales_unc"])
    var_x = positive_forward(params["var_x_unc"])
    var_y = positive_forward(params["var_y_unc"])                       # length dy
    u_arr = u.permute(1, 0, 2).unsqueeze(2).expand(T, B, S, du)
    y_arr = y.permute(1, 0, 2).unsqueeze(2).expand(T, B, S, dy)
    xs = [x0.unsqueeze(1).expand(B, S, dx)]                             # :80 / :92 tile over samples
    kls = []
    pad = torch.zeros((B, S, dx - dy), dtype=DT)
    for t in range(T - 1):
        x_t = xs[t]
        in_t = torch.cat((x_t, u_arr[t]), dim=2)                                     # :124
        fmean, fvar = gp_f.predict(in_t.reshape(B * S, du + dx))                     # :127-128
        fmean = fmean.reshape(B, S, dx) + in_t[:, :, :dx]                            # :130,132
        fvar = fvar.reshape(B, S, dx) + var_x                                        # :131,133
        eps = eps_f[t].unsqueeze(-1).expand(B, S, dx)                                # :136
        var_y_t = var_y.reshape(1, 1, dy) + (cfg.k_factor - 1.0) * fvar[:, :, :dy]   # :139-141
        y_diff = y_arr[t + 1] - fmean[:, :, :dy]                                     # :142
        s = var_y_t + fvar[:, :, :dy]                                                # :143
        k = fvar[:, :, :dy] * torch.reciprocal(s)                                    # :144
        mu = fmean + torch.cat((k * y_diff, pad), dim=2)                             # :146
        sig = (1.0 - torch.cat((k, pad), dim=2)) ** 2 * fvar                         # :147-148
        sig = sig + torch.cat((k ** 2 * var_y_t, pad), dim=2)                        # :149
        x_c = mu + eps * torch.sqrt(sig)                                             # :150
        x_nc = fmean + eps * torch.sqrt(fvar)                                        # :153
        do_cond = bool(condition) or (t < R - 1)                                     # :156
        xs.append(x_c if do_cond else x_nc)
        kl_reg = torch.log(fvar) - torch.log(sig) + (sig + (mu - fmean) ** 2) / fvar - 1.0   # :161
        kls.append(torch.sum(kl_reg) * (0.5 if do_cond else 0.0))
    x_final = torch.stack(xs, dim=0).permute(1, 0, 2, 3)
    y_final = x_final[..., :dy]
    kl_x = torch.sum(torch.stack(kls)) if kls else torch.zeros((), dtype=DT)
    sd = torch.sqrt(var_y).reshape(1, 1, 1, dy)
    zs = (y.unsqueeze(2).expand(B, T, S, dy) - y_final) / sd
    loglik = torch.sum(-0.5 * torch.sum(zs * zs, dim=-1) - torch.sum(torch.log(sd)) - 0.5 * dy * math.log(2 * math.pi))
    kl_z_f = gp_f.prior_kl()
    lf = cfg.loss_factors
    elbo = loglik * lf[0] / S - kl_x * lf[0] / S - kl_z_f                           # :188-191
    pred_mean = y_final.mean(dim=2)
    pred_var = y_final.var(dim=2, unbiased=False) + var_y
    return dict(loss=-elbo, loglik=loglik, kl_x=kl_x, kl_z_f=kl_z_f, x_final=x_final,
                pred_mean=pred_mean, pred_var=pred_var)


def loss_and_grads_half(cfg, params, u, y, x0, eps_f, condition=True):
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    x0l = torch.as_tensor(x0, dtype=DT).detach().clone().requires_grad_(True)
    res = elbo_half(cfg, leaf, u, y, x0l, eps_f, condition)
    grads = torch.autograd.grad(res["loss"], [leaf[k] for k in HALF_PARAM_NAMES] + [x0l], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(leaf[k])) for k, g in zip(HALF_PARAM_NAMES, grads[:-1])}
    gd["x0"] = grads[-1] if grads[-1] is not None else torch.zeros_like(x0l)
    return res, gd
