"""``CBFSSM`` with the reference's constructor, config dict and attribute vocabulary
(cbfssm/model/cbfssm.py:10-277), executed on a B200 through the C ABI.

Fetch handles: ``loss, train, pred_mean, pred_var, internal_mean, internal_var, mse,
sde, x_final, y_final, y_tilde, entropy, kl_x, init`` and ``var_dict`` (same keys as
cbfssm.py:56-67).  One ``sess.run`` evaluates one minibatch:

* any fetch            -> GP prologues + backward message + forward rollout (CUDA)
* ``train``            -> + reverse kernels, [all-reduce], prologue adjoints, TF-Adam
* prediction handles   -> + state export and particle moments (cbfssm.py:264-271)

Normal draws come from the in-library Philox generator (replacing ``tf.random_normal``,
cbfssm.py:134,149,209) unless ``model.inject_draws(eps_b, z_b, eps_f)`` supplied them.
With a ``torch.distributed`` group the particles n = b*S+s of each minibatch are split
contiguously over the ranks.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..engine import ElboEngine, ModelDims, init_param_arrays
from .base_model import BaseModel, Fetch, OutOfRangeError

_PRED = ("pred_mean", "pred_var", "internal_mean", "internal_var", "mse", "sde", "x_final", "y_final", "y_tilde")


def precision_flags(cfg):
    """``config['gpu_precision']`` (not a reference key): 'float32' (default) runs the float32 rollout kernels the
    north-star asks for (1e-4 against the float64 reference while cond(K_zz) <~ 1e3); 'float64' selects the float64
    batched path (CBF_FLAG_FP64) -- the reference's own default dtype (base_model.py:8), ~20x slower at M = 100, for
    inducing sets float32 cannot resolve.  M > 128 always runs in float64."""
    prec = str(cfg.get("gpu_precision", "float32"))
    if prec not in ("float32", "float64"):
        raise ValueError("config['gpu_precision'] must be 'float32' or 'float64'")
    return 128 if prec == "float64" else 0


class Saver:
    """tf.train.Saver look-alike (cbfssm.py:276; trainer.py:31,59,63): like the reference's, it persists
    *every* variable of the model -- whatever ``model.state_dict()`` returns (the 12 tensors and their Adam
    slots; CBFSSMHALF adds its recognition network and that network's Adam slots)."""

    def __init__(self, model):
        self.model = model

    def save(self, sess, path):
        model = self.model
        if model.rank == 0:
            tmp = path + ".tmp"
            torch.save(model.state_dict(), tmp)
            os.replace(tmp, path)
        if model.world > 1:      # nobody restores before rank 0 has finished writing
            torch.distributed.barrier(group=model._group)
        return path

    def restore(self, sess, path):
        self.model.load_state_dict(torch.load(path, map_location="cpu"))


class CBFSSM(BaseModel):

    def __init__(self, config, dtype=np.float64, device="cuda", group=None, seed=None):
        self._device = device
        self._group = group
        self._seed = seed
        super().__init__(config, dtype=dtype)

    # ---- "graph construction" (cbfssm.py:15-23) ----
    def _build_graph(self):
        cfg = self.config
        self.dims = ModelDims(dim_x=int(cfg["dim_x"]), dim_u=int(cfg["ds"].dim_u), dim_y=int(cfg["ds"].dim_y),
                              ind_pnt_num=int(cfg["ind_pnt_num"]), samples=int(cfg["samples"]),
                              recog_len=int(cfg["recog_len"]), k_factor=float(cfg["k_factor"]),
                              loss_factors=tuple(float(v) for v in cfg["loss_factors"]))
        self.world = torch.distributed.get_world_size(self._group) if self._group is not None else 1
        self.rank = torch.distributed.get_rank(self._group) if self._group is not None else 0
        self.engine = ElboEngine(self.dims, device=self._device, group=self._group if self.world > 1 else None)
        self.engine.flags = precision_flags(cfg)
        for name in ("loss", "train", "init", "entropy", "kl_x") + _PRED:
            setattr(self, name, self._handle(name))
        # cbfssm.py:56-67
        self.var_dict = {k: self._handle("var:" + k) for k in (
            'process noise', 'observation noise', 'kernel lengthscales f', 'kernel variance f', 'IP pos f',
            'IP mean f', 'IP var f', 'kernel lengthscales b', 'kernel variance b', 'IP pos b', 'IP mean b',
            'IP var b')}
        self.saver = Saver(self)
        self._draw_seed = 0x5EED if self._seed is None else int(self._seed)
        self._draw_counter = 0
        self._injected = None
        self._bufs = {}
        self.initialize()

    def initialize(self):
        """``sess.run(model.init)`` (trainer.py:33): draw the initial values."""
        eng = self.engine
        eng.set_params(init_param_arrays(self.dims, self.config, self._seed))
        if self.world > 1:      # one set of initial values for all ranks
            torch.distributed.broadcast(eng.theta, src=0, group=self._group)
        eng.adam_m.zero_()
        eng.adam_v.zero_()
        eng.adam_t = 0
        self._warn_if_ill_conditioned()

    def _warn_if_ill_conditioned(self):
        """The float32 kernels form a = K_zz^-1 k with an explicit float32 inverse where the reference does float64
        Cholesky solves; tell the user once if the initial inducing set is beyond their reach."""
        eng = self.engine
        if eng.flags & 128 or self.dims.ind_pnt_num > 128:
            return
        eng.prologue()
        worst = max(eng.cond_kzz().values())
        if worst > 1e4:
            import warnings
            warnings.warn("cond_1(K_zz) = %.1e at initialisation: the float32 rollout kernels are accuracy-limited for "
                          "this inducing set (gradient errors ~ cond * 1e-7); set config['gpu_precision'] = 'float64'"
                          % worst, RuntimeWarning, stacklevel=3)

    def state_dict(self):
        """Everything a checkpoint must hold: parameters and optimiser state (tf.train.Saver saves all global
        variables, Adam slots included)."""
        eng = self.engine
        return {"theta": eng.theta.cpu(), "adam_m": eng.adam_m.cpu(), "adam_v": eng.adam_v.cpu(),
                "adam_t": eng.adam_t, "names": eng.names}

    def load_state_dict(self, ck):
        eng = self.engine
        if tuple(ck["names"]) != tuple(eng.names) or ck["theta"].numel() != eng.theta.numel():
            raise ValueError("checkpoint was written by a model with different variables")
        eng.theta.copy_(ck["theta"])
        eng.adam_m.copy_(ck["adam_m"])
        eng.adam_v.copy_(ck["adam_v"])
        eng.adam_t = int(ck["adam_t"])

    def inject_draws(self, eps_b, z_b, eps_f):
        """Use these N(0,1) draws ([2,T,B,S], [2,T,B,S], [T-1,B,S]) for the next minibatch."""
        self._injected = (np.asarray(eps_b), np.asarray(z_b), np.asarray(eps_f))

    # ---- one minibatch ----
    def _shard(self, B):
        N = B * self.dims.samples
        per = -(-N // self.world)
        n0 = min(self.rank * per, N)
        n1 = min(n0 + per, N)
        return n0, n1 - n0

    def _draws(self, B, T, n0, nl):
        dev = self.engine.device
        key = (T, nl)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(2, T, nl, dtype=torch.float32, device=dev),
                               torch.empty(2, T, nl, dtype=torch.float32, device=dev),
                               torch.empty(max(T - 1, 1), nl, dtype=torch.float32, device=dev))
        eb, zb, ef = self._bufs[key]
        if self._injected is not None:
            ie, iz, if_ = self._injected
            self._injected = None
            N = B * self.dims.samples
            eb.copy_(torch.as_tensor(ie.reshape(2, T, N)[:, :, n0:n0 + nl], dtype=torch.float32))
            zb.copy_(torch.as_tensor(iz.reshape(2, T, N)[:, :, n0:n0 + nl], dtype=torch.float32))
            if T > 1:
                ef[:T - 1].copy_(torch.as_tensor(if_.reshape(T - 1, N)[:, n0:n0 + nl], dtype=torch.float32))
        else:
            c = self._draw_counter
            self._draw_counter += 1
            sid = lambda slot: (c << 32) | (slot << 8) | (self.rank & 0xFF)     # Philox stream: (step, tensor / row, rank)
            self.engine.fill_normal(eb, self._draw_seed, sid(0))
            self.engine.fill_normal(ef, self._draw_seed, sid(1))
            # z_b is read only where a run resamples (cbfssm.py:123-136: the random_normal sits inside the tf.cond
            # branch), i.e. at 2 * ceil(T / 2R) of its 2T rows: draw just those rows
            R = self.dims.recog_len
            for run in (0, 1):
                for t in range(T):
                    if (t + (1 if run == 0 else R + 1)) % (2 * R) == 0:
                        self.engine.fill_normal(zb[run, t], self._draw_seed, sid(2 + run * T + t))
        return eb, zb, ef

    def evaluate_batch(self, u_host, y_host, names, condition=True):
        """Evaluate handle names on one minibatch given as host arrays [B,T,du], [B,T,dy]."""
        eng, d = self.engine, self.dims
        dev = eng.device
        as_f32 = lambda a: (a if (torch.is_tensor(a) and a.dtype == torch.float32 and a.is_contiguous())
                            else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)))
        u, y = as_f32(u_host), as_f32(y_host)      # pinned tensors are copied asynchronously
        B, T, _ = u.shape
        n0, nl = self._shard(B)
        # decided from (N, world) alone, so every rank raises together instead of some entering the all-reduce
        per = -(-(B * d.samples) // self.world)
        if per * (self.world - 1) >= B * d.samples:
            raise ValueError("minibatch has too few particles for this many ranks")
        # this rank reads u, y only of the sequences its particles belong to: copy just those rows and make the
        # particle offset relative to the first of them
        b_lo, b_hi = n0 // d.samples, (n0 + nl - 1) // d.samples
        u_part, y_part = u[b_lo:b_hi + 1], y[b_lo:b_hi + 1]
        n0_rel = n0 - b_lo * d.samples
        # host -> device copies go on a side stream and overlap the generation of the step's normal draws
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            ud = u_part.to(dev, non_blocking=True)
            yd = y_part.to(dev, non_blocking=True)
        eb, zb, ef = self._draws(B, T, n0, nl)
        main.wait_stream(self._copy_stream)
        ud.record_stream(main)
        yd.record_stream(main)
        n0 = n0_rel
        # Outputs.create_all fetches only prediction handles with condition False (outputs.py:68-71,128-130): then
        # the message is needed for t < recog_len only
        pred_only = (not condition) and all(n in ("pred_mean", "pred_var", "internal_mean", "internal_var", "mse", "sde",
                                                  "x_final", "y_final") or n.startswith("var:") for n in names)
        out = eng.forward(ud, yd, eb, zb, ef, condition=condition, n_offset=n0, n_local=nl, predict_only=pred_only)
        if "train" in names:
            eng.backward()                   # all-reduces gradient + terms when sharded
            out = eng.loss_terms(eng.terms)
            eng.adam_step(float(self.config["learning_rate"]))
        elif self.world > 1:
            t = eng.terms.clone()
            torch.distributed.all_reduce(t, group=self._group)
            out = eng.loss_terms(t)
        res = {}
        if any(n in _PRED for n in names):
            if self.world > 1:
                # the particles of a sequence are spread over ranks: moments over the particle axis become one
                # all-reduce of per-sequence [sum x, sum x^2] (cbfssm.py:267-269)
                if any(n in ("x_final", "y_final", "y_tilde") for n in names):
                    raise NotImplementedError("per-particle states are not gathered across ranks; fetch the moments")
                sums = torch.zeros(B, T, d.dim_x, 2, dtype=torch.float64, device=dev)
                sums[b_lo:b_hi + 1] = eng.state_sums()
                torch.distributed.all_reduce(sums, group=self._group)
                mean = sums[..., 0] / d.samples
                var = sums[..., 1] / d.samples - mean * mean
                yfull = y.to(dev)
                pm = mean[..., :d.dim_y].float()
                pv = (var[..., :d.dim_y] + eng.var_y[:d.dim_y].double()).float()
                res.update(pred_mean=pm, pred_var=pv, internal_mean=mean.float(), internal_var=var.float())
                yd_all = yfull
            else:
                xf, yt = eng.export_states(yd)
                pm, pv = eng.moments(xf, d.dim_y, eng.var_y)
                im, iv = eng.moments(xf, d.dim_x, None)
                res.update(x_final=xf, y_tilde=yt, y_final=xf[..., :d.dim_y], pred_mean=pm, pred_var=pv,
                           internal_mean=im, internal_var=iv)
                yd_all = yd
            if "mse" in names:   # tf.losses.mean_squared_error casts to float32 (cbfssm.py:270)
                res["mse"] = torch.mean((pm - yd_all) ** 2)
            if "sde" in names:
                res["sde"] = torch.abs(pm - yd_all) / torch.sqrt(pv)
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

    def _var_value(self, key):
        import torch.nn.functional as F
        eng = self.engine
        sp = lambda t: (F.softplus(t) + 1e-10).cpu().numpy()
        tag = key[-1] if key[-2:] in (" f", " b") else None
        table = {'process noise': lambda: sp(eng.view("var_x_unc")),
                 'observation noise': lambda: sp(eng.view("var_y_unc"))}
        if tag:
            table.update({f'kernel lengthscales {tag}': lambda: sp(eng.view(f"{tag}.lengthscales_unc")),
                          f'kernel variance {tag}': lambda: sp(eng.view(f"{tag}.variance_unc")),
                          f'IP pos {tag}': lambda: eng.view(f"{tag}.zeta_pos").cpu().numpy(),
                          f'IP mean {tag}': lambda: eng.view(f"{tag}.zeta_mean").cpu().numpy(),
                          f'IP var {tag}': lambda: sp(eng.view(f"{tag}.zeta_var_unc"))})
        return table[key]()

    def _session_run(self, fetches, feed_dict):
        single = not isinstance(fetches, (tuple, list))
        flist = [fetches] if single else list(fetches)
        names = [f.name for f in flist]
        if all(n == "init" for n in names):
            self.initialize()
            return None if single else tuple(None for _ in names)
        if all(n.startswith("var:") for n in names):
            vals = [self._var_value(n[4:]) for n in names]
            return vals[0] if single else tuple(vals)
        cond = True
        for k, v in feed_dict.items():
            if isinstance(k, Fetch) and k.name == "condition":
                cond = bool(v)
        u, y = self._next_batch()          # raises OutOfRangeError when drained
        vals = self.evaluate_batch(u, y, names, cond)
        return vals[0] if single else tuple(vals)
