"""Device-side driver of the sampled-ELBO hot path through the C ABI.

``ElboEngine`` owns the float64 master copy of the 12 trainable tensors
(creation order of cbfssm/model/gp_tf.py:112-127 within cbfssm/model/cbfssm.py:30-54),
the float32 kernel operands, the workspace and the gradient buffers, and issues,
per minibatch, on the current CUDA stream:

    noise / GP prologue (float64)  ->  cbf_elbo_forward  ->  cbf_elbo_backward
    -> [one all-reduce of the flat kernel-level gradient + the three ELBO terms]
    -> prologue adjoints (float64) -> TF-style Adam.

PyTorch is used for device memory, streams and ``torch.distributed`` only.
"""
from __future__ import annotations

import contextlib
import os
import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import cbf_gp, cbf_grad_layout, cbf_shape, check, ptr

F64 = torch.float64
F32 = torch.float32

GP_FIELDS = ("zeta_pos", "zeta_mean", "zeta_var_unc", "variance_unc", "lengthscales_unc")


def _softplus_inverse(y):
    """tf_transform.backward (cbfssm/model/tf_transform.py:13-16)."""
    y = np.asarray(y, dtype=np.float64)
    if np.any(y <= 1e-10):
        raise AssertionError("Input to backward transformation should be greater 1e-10")
    with np.errstate(over="ignore"):
        r = np.log(np.exp(y - 1e-10) - 1.0)
    return np.where(y > 35, y - 1e-10, r)


@dataclass
class ModelDims:
    dim_x: int
    dim_u: int
    dim_y: int
    ind_pnt_num: int
    samples: int
    recog_len: int
    k_factor: float = 1.0
    loss_factors: tuple = (10.0, 0.0)
    half: bool = False        # CBFSSMHALF (cbfssm/model/cbfssmhalf.py): no backward GP, var_y of length dim_y

    @property
    def dim_h(self):
        return self.dim_x - self.dim_y

    @property
    def dim_in(self):
        return self.dim_x + self.dim_u


def param_shapes(d: ModelDims):
    """Name -> shape of the 12 raw tensors, in the reference's creation order."""
    M, din = d.ind_pnt_num, d.dim_in
    out = {}
    for tag, dout in ((("f", d.dim_x),) if d.half else (("f", d.dim_x), ("b", d.dim_h))):
        out[f"{tag}.zeta_pos"] = (M, din)
        out[f"{tag}.zeta_mean"] = (M, dout)
        out[f"{tag}.zeta_var_unc"] = (M, dout)
        out[f"{tag}.variance_unc"] = ()
        out[f"{tag}.lengthscales_unc"] = (din,)
    out["var_x_unc"] = (d.dim_x,)
    # CBFSSM: length dim_x, sic (cbfssm.py:53, run/template.py:37); CBFSSMHALF: dim_y (cbfssmhalf.py:36)
    out["var_y_unc"] = (d.dim_y,) if d.half else (d.dim_x,)
    return out


def init_param_arrays(d: ModelDims, config: dict, seed: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Initial values (gp_tf.py:112-127, cbfssm.py:51-54).  The reference draws from the
    unseeded global NumPy RNG; ``seed`` makes the same sequence of calls reproducible."""
    rs = np.random.RandomState(seed) if seed is not None else np.random
    M, din = d.ind_pnt_num, d.dim_in
    out = {}
    for tag, dout in ((("f", d.dim_x),) if d.half else (("f", d.dim_x), ("b", d.dim_h))):
        out[f"{tag}.zeta_pos"] = rs.uniform(low=-config["zeta_pos"], high=config["zeta_pos"], size=(M, din))
        out[f"{tag}.zeta_mean"] = config["zeta_mean"] * rs.rand(M, dout)
        out[f"{tag}.zeta_var_unc"] = _softplus_inverse(config["zeta_var"] * np.ones((M, dout)))
        out[f"{tag}.variance_unc"] = _softplus_inverse(config["gp_var"]).reshape(())
        out[f"{tag}.lengthscales_unc"] = _softplus_inverse(np.asarray([config["gp_len"]] * din, dtype=np.float64))
    out["var_x_unc"] = _softplus_inverse(config["var_x"])
    out["var_y_unc"] = _softplus_inverse(np.asarray(config["var_y"], dtype=np.float64)[:d.dim_y] if d.half
                                         else config["var_y"])
    return out


def count_chain_batches(T, R, per_launch=120):
    """Launches the backward-message kernels need: live chain segments of both runs
    (cbfssm.py:123-136), at most ``per_launch`` per launch (csrc/common.cuh kMaxChains)."""
    n = 0
    for run in (0, 1):
        off = 1 if run == 0 else R + 1
        starts = [T - 1] + [t for t in range(T - 2, -1, -1) if (t + off) % (2 * R) == 0]
        for i, t_hi in enumerate(starts):
            t_next = starts[i + 1] if i + 1 < len(starts) else -1
            if any(((t % (2 * R)) < R) == (run == 0) for t in range(t_hi, t_next, -1)):
                n += 1
    return max(1, -(-n // per_launch)) if n else 0


class _GpBuffers:
    """float32 operands + float64 prologue state of one GP."""

    def __init__(self, M, din, dout, device, lib):
        self.M, self.din, self.dout = M, din, dout
        z = lambda *s: torch.zeros(*s, dtype=F32, device=device)
        self.Z, self.ell, self.sig2 = z(M, din), z(max(din, 4)), z(4)
        self.P, self.alpha, self.S = z(M, M), z(M, dout), z(M, dout)
        self.kl = torch.zeros(1, dtype=F64, device=device)
        self.state = torch.zeros(int(lib.cbf_gp_prologue_state_doubles(M, din, dout)), dtype=F64, device=device)
        self.c = cbf_gp(*(C.c_void_p(t.data_ptr()) for t in (self.Z, self.ell, self.sig2, self.P, self.alpha, self.S,
                                                              self.state)))


class ElboEngine:
    def __init__(self, dims: ModelDims, device="cuda", group=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("cbf_ssm_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.dims = dims
        self.device = torch.device(device)
        self.group = group
        d = dims
        self.kernel_path = self.lib.cbf_supported(d.ind_pnt_num, d.dim_x, d.dim_u, d.dim_y)
        if not self.kernel_path:
            raise _lib.CbfError(-2, f"M={d.ind_pnt_num}, dims=({d.dim_x},{d.dim_u},{d.dim_y}) not supported by the "
                                    "compiled library (float64 path: dim_x <= 16, dim_x + dim_u <= 31)")
        # ---- flat float64 parameter vector with named views ----
        shapes = param_shapes(d)
        self.names = tuple(shapes)
        sizes = [int(np.prod(s)) if len(s) else 1 for s in shapes.values()]
        self.theta = torch.zeros(sum(sizes), dtype=F64, device=self.device)
        self.grad = torch.zeros_like(self.theta)
        self.adam_m = torch.zeros_like(self.theta)
        self.adam_v = torch.zeros_like(self.theta)
        self.adam_t = 0
        self.offsets, o = {}, 0
        for n, sz in zip(self.names, sizes):
            self.offsets[n] = (o, sz)
            o += sz
        self.gp_f = _GpBuffers(d.ind_pnt_num, d.dim_in, d.dim_x, self.device, self.lib)
        self.gp_b = None if d.half else _GpBuffers(d.ind_pnt_num, d.dim_in, d.dim_h, self.device, self.lib)
        self.var_x = torch.zeros(max(d.dim_x, 4), dtype=F32, device=self.device)
        self.var_y = torch.zeros(max(d.dim_x, 4), dtype=F32, device=self.device)
        self.terms = torch.zeros(4, dtype=F64, device=self.device)
        self._scratch32 = torch.zeros(max(d.dim_x, 4), dtype=F32, device=self.device)
        self._scratch64 = torch.zeros(max(d.dim_x, 4), dtype=F64, device=self.device)
        self.x0_bar = None        # CBFSSMHALF: d loss / d x0 [nb, dim_x] after backward()
        self._ws = None
        self._ws_key = None
        self._side = None
        self._launch_base = 0
        self._gl = None
        self._gflat = None
        self._shape = None
        self._saved = None
        self.flags = 0             # CBF_FLAG_* passed in cbf_shape.flags

    # ---------------- parameters ----------------
    def view(self, name, of=None):
        o, sz = self.offsets[name]
        base = self.theta if of is None else of
        return base[o:o + sz].view(param_shapes(self.dims)[name])

    def set_params(self, arrays: Dict[str, np.ndarray]):
        for n in self.names:
            self.view(n).copy_(torch.as_tensor(np.asarray(arrays[n], dtype=np.float64)).reshape(self.view(n).shape))

    def get_params(self) -> Dict[str, np.ndarray]:
        return {n: self.view(n).detach().cpu().numpy().copy() for n in self.names}

    def get_grads(self) -> Dict[str, np.ndarray]:
        return {n: self.view(n, self.grad).detach().cpu().numpy().copy() for n in self.names}

    # ---------------- plumbing ----------------
    @property
    def launches(self):
        """Kernels of this library launched on this host thread since the counter was last set (the library
        counts every launch itself: cbf_launches_read; bench.py's gpu_launches)."""
        n = C.c_int64(0)
        check(self.lib.cbf_launches_read(C.byref(n), 0))
        return int(n.value) + self._launch_base

    @launches.setter
    def launches(self, value):
        n = C.c_int64(0)
        check(self.lib.cbf_launches_read(C.byref(n), 1))
        self._launch_base = int(value)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @contextlib.contextmanager
    def _two_gps(self):
        """Streams (f, b) for the float64 single-CTA kernels of the two GPs.  From M = 48 the O(M^3) prologue
        and its adjoint take 0.2-0.4 ms each and the two GPs are independent, so the backward-message GP's
        kernel runs on a side stream next to the other one (fork / join by events); below that the kernels
        are tens of microseconds and both stay on the caller's stream."""
        main = torch.cuda.current_stream(self.device)
        if self.dims.ind_pnt_num < 48 or self.dims.half:
            yield self._stream(), self._stream()
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork.record(main)
        self._side.wait_event(self._ev_fork)
        yield self._stream(), C.c_void_p(self._side.cuda_stream)
        self._ev_join.record(self._side)
        main.wait_event(self._ev_join)

    def make_shape(self, B, T, condition=True, n_offset=0, n_local=None, predict_only=False):
        d = self.dims
        n_local = B * d.samples - n_offset if n_local is None else n_local
        return cbf_shape(B, d.samples, T, d.ind_pnt_num, d.dim_x, d.dim_u, d.dim_y, d.recog_len,
                         1 if condition else 0, n_offset, n_local, float(d.k_factor),
                         int(self.flags) | (32 if d.half else 0) | (256 if predict_only and not condition else 0))

    def _ensure_workspace(self, shape):
        # The plan the library binds onto the buffer depends on everything in ``shape`` (flags select the kernel
        # path and with it extra sections) and on the tensor path's window budget, so the size is asked for on
        # every call (a host-side computation) and the buffer grows whenever the answer exceeds it.
        nbytes = C.c_size_t(0)
        check(self.lib.cbf_workspace_bytes(C.byref(shape), C.byref(nbytes)))
        key = (shape.B, shape.T, shape.n_local, int(shape.flags))
        if self._ws is None or nbytes.value > self._ws.numel():
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            if os.environ.get("CBFSSM_B200_POISON_WS"):     # test aid: every byte 0xFF = NaN in float32 / float64,
                self._ws.fill_(0xFF)                        # so a read of never-written workspace cannot go unnoticed
        if self._ws_key != key:
            gl = cbf_grad_layout()
            check(self.lib.cbf_grad_layout_get(C.byref(shape), C.byref(gl)))
            self._gl = gl
            # flat kernel-level gradient, followed by the three ELBO terms so that one
            # all-reduce covers both
            if self._gflat is None or self._gflat.numel() != gl.total + 4:
                self._gflat = torch.zeros(gl.total + 4, dtype=F64, device=self.device)
            self._ws_key = key

    def prologue(self):
        """Raw tensors -> float32 kernel operands, KL(q(u)||p(u)) per GP (float64)."""
        lib, st, d = self.lib, self._stream(), self.dims
        if d.half:      # var_x has dx entries, var_y only dy: constrain them separately
            check(lib.cbf_noise_forward(d.dim_x, ptr(self.view("var_x_unc")), ptr(self.view("var_x_unc")),
                                        ptr(self.var_x), ptr(self._scratch32), st))
            check(lib.cbf_noise_forward(d.dim_y, ptr(self.view("var_y_unc")), ptr(self.view("var_y_unc")),
                                        ptr(self.var_y), ptr(self._scratch32), st))
        else:
            check(lib.cbf_noise_forward(d.dim_x, ptr(self.view("var_x_unc")), ptr(self.view("var_y_unc")),
                                        ptr(self.var_x), ptr(self.var_y), st))
        with self._two_gps() as streams:
            for tag, g, s2 in ((("f", self.gp_f, streams[0]),) if d.half else
                               (("f", self.gp_f, streams[0]), ("b", self.gp_b, streams[1]))):
                check(lib.cbf_gp_prologue(g.M, g.din, g.dout, *(ptr(self.view(f"{tag}.{f}")) for f in GP_FIELDS),
                                          ptr(g.Z), ptr(g.ell), ptr(g.sig2), ptr(g.P), ptr(g.alpha), ptr(g.S),
                                          ptr(g.kl), ptr(g.state), s2))

    def cond_kzz(self):
        """1-norm condition numbers of K_zz + 1e-8 I of the GPs as of the last prologue (synchronises).  The float32
        rollout kernels hold 1e-4 against the float64 reference up to ~1e3."""
        out = {}
        for tag, g in (("f", self.gp_f),) + ((("b", self.gp_b),) if self.gp_b is not None else ()):
            out[tag] = float(g.state[-1])      # ProState.cond is the last slot of the state buffer
        return out

    def forward(self, u, y, eps_b, z_b, eps_f, condition=True, n_offset=0, n_local=None, run_prologue=True, x0=None,
                predict_only=False):
        """u [B,T,du], y [B,T,dy] float32 device tensors; draws float32 device tensors
        eps_b/z_b [2,T,n_local], eps_f [T-1,n_local].  Returns a dict of 0-d device
        tensors (this shard's loglik/kl_x/entropy; loss is global when a group is set
        only after ``backward``)."""
        B, T, _ = u.shape
        # predict_only: only the prediction outputs will be read (no loss / entropy / y_tilde / backward): with
        # condition False the library then runs just the message chain the free-running rollout needs
        shape = self.make_shape(B, T, condition, n_offset, n_local, predict_only)
        self._ensure_workspace(shape)
        if run_prologue:
            self.prologue()
        if self.dims.half:
            if x0 is None:
                raise ValueError("CBFSSMHALF: forward needs x0 [B, dim_x] from the recognition model")
            check(self.lib.cbf_elbo_forward_half(C.byref(shape), C.byref(self.gp_f.c), ptr(self.var_x), ptr(self.var_y),
                                                 ptr(u), ptr(y), ptr(x0), ptr(eps_f), ptr(self.terms), ptr(self._ws),
                                                 self._stream()))
        else:
            check(self.lib.cbf_elbo_forward(C.byref(shape), C.byref(self.gp_f.c), C.byref(self.gp_b.c),
                                            ptr(self.var_x), ptr(self.var_y), ptr(u), ptr(y), ptr(eps_b), ptr(z_b),
                                            ptr(eps_f), ptr(self.terms), ptr(self._ws), self._stream()))
        self._shape = shape
        self._saved = (u, y, eps_b, z_b, eps_f, x0)
        return self.loss_terms(self.terms)

    def loss_terms(self, terms):
        d = self.dims
        l1, l2 = (float(v) for v in d.loss_factors)
        S = float(d.samples)
        kl_f = self.gp_f.kl[0]
        if d.half:      # cbfssmhalf.py:188-191: no entropy term, one inducing KL
            elbo = (l1 / S) * (terms[0] - terms[1]) - kl_f
            return dict(loss=-elbo, loglik=terms[0], kl_x=terms[1], entropy=terms[2], kl_z_f=kl_f)
        kl_b = self.gp_b.kl[0]
        elbo = (l1 / S) * (terms[0] - terms[1]) + (l2 / S) * terms[2] - kl_f - kl_b    # cbfssm.py:257-261
        return dict(loss=-elbo, loglik=terms[0], kl_x=terms[1], entropy=terms[2], kl_z_f=kl_f, kl_z_b=kl_b)

    def backward(self):
        """Gradient of the loss w.r.t. the 12 raw tensors into ``self.grad`` (flat float64).
        With a process group: all-reduces [kernel-level gradient | ELBO terms] once, so
        every rank ends with the global gradient and ``self.terms`` holds global terms."""
        lib, st, d, shape, gl = self.lib, self._stream(), self.dims, self._shape, self._gl
        u, y, eps_b, z_b, eps_f, x0 = self._saved
        l1, l2 = (float(v) for v in d.loss_factors)
        S = float(d.samples)
        w = (C.c_double * 3)(-l1 / S, l1 / S, -l2 / S)
        gflat = self._gflat
        if d.half:
            nb = shape.n_local // d.samples
            if self.x0_bar is None or self.x0_bar.shape[0] != nb:
                self.x0_bar = torch.zeros(nb, d.dim_x, dtype=F64, device=self.device)
            check(lib.cbf_elbo_backward_half(C.byref(shape), C.byref(self.gp_f.c), ptr(self.var_x), ptr(self.var_y),
                                             ptr(u), ptr(y), ptr(x0), ptr(eps_f), w, ptr(gflat), ptr(self.x0_bar),
                                             ptr(self._ws), st))
        else:
            check(lib.cbf_elbo_backward(C.byref(shape), C.byref(self.gp_f.c), C.byref(self.gp_b.c), ptr(self.var_x),
                                        ptr(self.var_y), ptr(u), ptr(y), ptr(eps_b), ptr(z_b), ptr(eps_f), w,
                                        ptr(gflat), ptr(self._ws), st))
        if self.group is not None:
            gflat[gl.total:gl.total + 3].copy_(self.terms[:3])
            torch.distributed.all_reduce(gflat, group=self.group)
            self.terms[:3].copy_(gflat[gl.total:gl.total + 3])
        at = lambda off: C.c_void_p(gflat.data_ptr() + 8 * off)
        gat = lambda name: ptr(self.view(name, self.grad))
        gps = [("f", self.gp_f, (gl.f_P, gl.f_alpha, gl.f_S, gl.f_Z, gl.f_ell, gl.f_sig2))]
        if not d.half:
            gps.append(("b", self.gp_b, (gl.b_P, gl.b_alpha, gl.b_S, gl.b_Z, gl.b_ell, gl.b_sig2)))
        with self._two_gps() as streams:
            for (tag, g, o), s2 in zip(gps, streams):
                check(lib.cbf_gp_prologue_backward(g.M, g.din, g.dout, *(at(x) for x in o), 1.0, ptr(g.state),
                                                   *(gat(f"{tag}.{f}") for f in GP_FIELDS), s2))
        if d.half:
            sc = ptr(self._scratch64)
            check(lib.cbf_noise_backward(d.dim_x, ptr(self.view("var_x_unc")), ptr(self.view("var_x_unc")),
                                         at(gl.var_x), at(gl.var_x), gat("var_x_unc"), sc, st))
            check(lib.cbf_noise_backward(d.dim_y, ptr(self.view("var_y_unc")), ptr(self.view("var_y_unc")),
                                         at(gl.var_y), at(gl.var_y), gat("var_y_unc"), sc, st))
        else:
            check(lib.cbf_noise_backward(d.dim_x, ptr(self.view("var_x_unc")), ptr(self.view("var_y_unc")),
                                         at(gl.var_x), at(gl.var_y), gat("var_x_unc"), gat("var_y_unc"), st))
        return self.grad

    def adam_step(self, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        """TF-1.8 AdamOptimizer update of all 12 tensors (cbfssm.py:274)."""
        self.adam_t += 1
        check(self.lib.cbf_adam_step(self.theta.numel(), ptr(self.theta), ptr(self.grad), ptr(self.adam_m),
                                     ptr(self.adam_v), self.adam_t, float(lr), beta1, beta2, eps, self._stream()))

    def export_states(self, y):
        """x_final, y_tilde as [nb, T, S, dx] float32 (cbfssm.py:97,181) of the last forward."""
        shape, d = self._shape, self.dims
        nb = shape.n_local // d.samples
        xf = torch.empty(nb, shape.T, d.samples, d.dim_x, dtype=F32, device=self.device)
        yt = None if d.half else torch.empty_like(xf)       # CBFSSMHALF has no y_tilde
        check(self.lib.cbf_export_states(C.byref(shape), ptr(y), ptr(xf), ptr(yt), ptr(self._ws), self._stream()))
        return xf, yt

    def state_sums(self):
        """[B, T, dx, 2] float64 (sum_s x, sum_s x^2) of the last forward over this shard's particles
        (sharded prediction moments: all-reduce, then mean = s1/S, var = s2/S - mean^2)."""
        shape, d = self._shape, self.dims
        sums = torch.empty(shape.B, shape.T, d.dim_x, 2, dtype=F64, device=self.device)
        check(self.lib.cbf_state_sums(C.byref(shape), ptr(sums), ptr(self._ws), self._stream()))
        return sums

    def moments(self, x, d_keep, add_var=None):
        """tf.nn.moments(axes=[2]) (+ add_var) over the particle axis of [nb,T,S,d]."""
        nb, T, S, dd = x.shape
        mean = torch.empty(nb, T, d_keep, dtype=F32, device=self.device)
        var = torch.empty_like(mean)
        check(self.lib.cbf_moments(ptr(x), nb, T, S, dd, d_keep, ptr(add_var), ptr(mean), ptr(var), self._stream()))
        return mean, var

    def fill_normal(self, out, seed, stream_id):
        check(self.lib.cbf_fill_normal(ptr(out), out.numel(), int(seed), int(stream_id), self._stream()))
        return out

    def kernel_level_grads(self):
        """The flat kernel-level gradient of the last backward, split by name (tests)."""
        gl, g, d = self._gl, self._gflat, self.dims
        M, din = d.ind_pnt_num, d.dim_in
        out = {}
        for tag, dout in ((("f", d.dim_x),) if d.half else (("f", d.dim_x), ("b", d.dim_h))):
            for nm, shp in (("P", (M, M)), ("alpha", (M, dout)), ("S", (M, dout)), ("Z", (M, din)),
                            ("ell", (din,)), ("sig2", (1,))):
                off = getattr(gl, f"{tag}_{nm}")
                out[f"{tag}.{nm}"] = g[off:off + int(np.prod(shp))].view(shp).cpu().numpy().copy()
        out["var_x"] = g[gl.var_x:gl.var_x + d.dim_x].cpu().numpy().copy()
        out["var_y"] = g[gl.var_y:gl.var_y + d.dim_x].cpu().numpy().copy()
        return out
