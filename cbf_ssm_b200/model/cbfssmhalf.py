"""``CBFSSMHALF`` with the reference's constructor and handles (cbfssm/model/cbfssmhalf.py).

No backward-message GP: x_0 comes from a recognition model -- ``config['recog_model']`` = ``'output'``
(first output zero-padded to dim_x, cbfssmhalf.py:76-80) or ``'rnn'`` (default: TF-1.8 ``GRUCell(16)``
over the reversed first ``recog_len`` steps of [u, y], then a dense layer, cbfssmhalf.py:82-92) --
and the forward step conditions only the first dim_y state dims.  The rollout, its reverse, the
likelihood and KL run through ``cbf_elbo_forward_half`` / ``cbf_elbo_backward_half``; the recognition
network is a few small PyTorch ops on the device (it is not on the hot path) whose parameters are
updated by the same TF-style Adam kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import ElboEngine, ModelDims, init_param_arrays
from .base_model import BaseModel
from .cbfssm import CBFSSM, Saver, precision_flags

GRU_UNITS = 16


def gru_tf(w, xy):
    """TF-1.8 GRUCell over xy [B, L, d] (already reversed); returns the final state [B, 16].
    [r, z] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = z*h + (1-z)*c."""
    B = xy.shape[0]
    h = torch.zeros(B, GRU_UNITS, dtype=xy.dtype, device=xy.device)
    for t in range(xy.shape[1]):
        g = torch.sigmoid(torch.cat((xy[:, t], h), dim=1) @ w["gates_kernel"] + w["gates_bias"])
        r, z = g[:, :GRU_UNITS], g[:, GRU_UNITS:]
        c = torch.tanh(torch.cat((xy[:, t], r * h), dim=1) @ w["candidate_kernel"] + w["candidate_bias"])
        h = z * h + (1.0 - z) * c
    return h


class CBFSSMHALF(CBFSSM):

    def _build_graph(self):
        cfg = self.config
        self.dims = ModelDims(dim_x=int(cfg["dim_x"]), dim_u=int(cfg["ds"].dim_u), dim_y=int(cfg["ds"].dim_y),
                              ind_pnt_num=int(cfg["ind_pnt_num"]), samples=int(cfg["samples"]),
                              recog_len=int(cfg["recog_len"]), k_factor=float(cfg["k_factor"]),
                              loss_factors=tuple(float(v) for v in cfg["loss_factors"]), half=True)
        self.world = torch.distributed.get_world_size(self._group) if self._group is not None else 1
        self.rank = torch.distributed.get_rank(self._group) if self._group is not None else 0
        if self.world > 1:
            raise NotImplementedError("CBFSSMHALF: single GPU (replicas only) in this round")
        self.engine = ElboEngine(self.dims, device=self._device, group=None)
        self.engine.flags = precision_flags(cfg)
        for name in ("loss", "train", "init", "entropy", "kl_x", "pred_mean", "pred_var", "internal_mean",
                     "internal_var", "mse", "sde", "x_final", "y_final", "y_tilde"):
            setattr(self, name, self._handle(name))
        self.var_dict = {k: self._handle("var:" + k) for k in (                   # cbfssmhalf.py:39-45
            'process noise', 'observation noise', 'kernel lengthscales f', 'kernel variance f', 'IP pos f',
            'IP mean f', 'IP var f')}
        self.saver = Saver(self)
        self.recog = cfg.get("recog_model", "rnn")
        if self.recog not in ("rnn", "output"):
            raise AssertionError('invalid config for recognition model')
        d = self.dims
        din = d.dim_u + d.dim_y
        # recognition-network weights: one flat float64 vector (creation order of TF's variables)
        self._phi_shapes = {} if self.recog == "output" else {
            "gates_kernel": (din + GRU_UNITS, 2 * GRU_UNITS), "gates_bias": (2 * GRU_UNITS,),
            "candidate_kernel": (din + GRU_UNITS, GRU_UNITS), "candidate_bias": (GRU_UNITS,),
            "dense_kernel": (GRU_UNITS, d.dim_x), "dense_bias": (d.dim_x,)}
        n = sum(int(np.prod(s)) for s in self._phi_shapes.values())
        dev = self.engine.device
        self.phi = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)
        self.phi_grad = torch.zeros_like(self.phi)
        self.phi_m = torch.zeros_like(self.phi)
        self.phi_v = torch.zeros_like(self.phi)
        self._draw_seed = 0x5EED if self._seed is None else int(self._seed)
        self._draw_counter = 0
        self._injected = None
        self._bufs = {}
        self.initialize()

    def phi_views(self, base=None):
        base = self.phi if base is None else base
        out, o = {}, 0
        for k, shp in self._phi_shapes.items():
            sz = int(np.prod(shp))
            out[k] = base[o:o + sz].view(shp)
            o += sz
        return out

    def initialize(self):
        eng = self.engine
        eng.set_params(init_param_arrays(self.dims, self.config, self._seed))
        eng.adam_m.zero_(); eng.adam_v.zero_(); eng.adam_t = 0
        rs = np.random.RandomState(None if self._seed is None else self._seed + 1)
        views = self.phi_views()
        for k, v in views.items():          # TF defaults: glorot-uniform kernels, gate bias 1, other biases 0
            if k.endswith("kernel"):
                lim = np.sqrt(6.0 / (v.shape[0] + v.shape[1]))
                v.copy_(torch.as_tensor(rs.uniform(-lim, lim, size=tuple(v.shape))))
            else:
                v.fill_(1.0 if k == "gates_bias" else 0.0)
        self.phi_m.zero_(); self.phi_v.zero_()

    def state_dict(self):
        """The GP / noise variables plus the recognition network and its Adam slots (the reference's Saver keeps
        the GRU and dense-layer variables, cbfssmhalf.py:85-91, like any other global variable)."""
        ck = super().state_dict()
        ck.update(recog_model=self.recog, phi=self.phi.cpu(), phi_m=self.phi_m.cpu(), phi_v=self.phi_v.cpu())
        return ck

    def load_state_dict(self, ck):
        if ck.get("recog_model") != self.recog or ck["phi"].numel() != self.phi.numel():
            raise ValueError("checkpoint was written with a different recognition model")
        super().load_state_dict(ck)
        self.phi.copy_(ck["phi"])
        self.phi_m.copy_(ck["phi_m"])
        self.phi_v.copy_(ck["phi_v"])

    def inject_draws(self, eps_f):
        """Use these N(0,1) draws [T-1, B, S] for the next minibatch."""
        self._injected = np.asarray(eps_f)

    def recognise(self, ud, yd, phi=None):
        """x_0 [B, dim_x] (float64, differentiable w.r.t. the recognition weights)."""
        d = self.dims
        if self.recog == "output":
            B = yd.shape[0]
            return torch.cat((yd[:, 0, :].double(), torch.zeros(B, d.dim_h, dtype=torch.float64, device=yd.device)), 1)
        w = self.phi_views(phi)
        xy = torch.cat((ud, yd), dim=2)[:, :d.recog_len, :].double()
        h = gru_tf(w, torch.flip(xy, dims=[1]))
        return h @ w["dense_kernel"] + w["dense_bias"]

    def evaluate_batch(self, u_host, y_host, names, condition=True):
        eng, d = self.engine, self.dims
        dev = eng.device
        as_f32 = lambda a: (a if (torch.is_tensor(a) and a.dtype == torch.float32 and a.is_contiguous())
                            else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)))
        u, y = as_f32(u_host), as_f32(y_host)
        B, T, _ = u.shape
        ud, yd = u.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        N = B * d.samples
        key = (T, N)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(max(T - 1, 1), N, dtype=torch.float32, device=dev)
        ef = self._bufs[key]
        if self._injected is not None:
            ef[:T - 1].copy_(torch.as_tensor(self._injected.reshape(T - 1, N), dtype=torch.float32))
            self._injected = None
        else:
            eng.fill_normal(ef, self._draw_seed, self._draw_counter * 4096)
            self._draw_counter += 1
        train = "train" in names
        phi = self.phi.detach().requires_grad_(train and self.recog == "rnn")
        x0 = self.recognise(ud, yd, phi)
        x0f = x0.detach().float().contiguous()
        out = eng.forward(ud, yd, None, None, ef, condition=condition, x0=x0f)
        if train:
            eng.backward()
            if self.recog == "rnn":
                (g,) = torch.autograd.grad(x0, phi, grad_outputs=eng.x0_bar)
                self.phi_grad.copy_(g)
            out = eng.loss_terms(eng.terms)
            lr = float(self.config["learning_rate"])
            eng.adam_step(lr)
            if self.recog == "rnn":        # same TF-Adam kernel, same step count
                from .._lib import check, ptr
                check(eng.lib.cbf_adam_step(self.phi.numel(), ptr(self.phi), ptr(self.phi_grad), ptr(self.phi_m),
                                            ptr(self.phi_v), eng.adam_t, lr, 0.9, 0.999, 1e-8, eng._stream()))
        res = {}
        if any(n in ("pred_mean", "pred_var", "internal_mean", "internal_var", "mse", "sde", "x_final", "y_final",
                     "y_tilde") for n in names):
            xf, _ = eng.export_states(yd)
            pm, pv = eng.moments(xf, d.dim_y, eng.var_y)
            im, iv = eng.moments(xf, d.dim_x, None)
            res.update(x_final=xf, y_final=xf[..., :d.dim_y], y_tilde=xf, pred_mean=pm, pred_var=pv,
                       internal_mean=im, internal_var=iv)
            if "mse" in names:
                res["mse"] = torch.mean((pm - yd) ** 2)
            if "sde" in names:
                res["sde"] = torch.abs(pm - yd) / torch.sqrt(pv)
        res.update(loss=out["loss"], entropy=out["entropy"], kl_x=out["kl_x"])
        vals = []
        for n in names:
            if n in ("train", "init"):
                vals.append(None)
            elif n.startswith("var:"):
                vals.append(self._var_value(n[4:]))
            else:
                vals.append(res[n].detach().cpu().numpy())
        return vals
