"""Second, independent CPU restatement (NumPy float64).  TEST INFRASTRUCTURE ONLY.

``cbfssm_oracle.py`` follows the reference op-for-op (two triangular solves per
step, autograd).  This file restates the same path in the *restructured* algebra
the sm_100a kernels use (SURVEY.md section 8a notes 3-5): a resident
``P = (K_zz + 1e-8 I)^-1``, ``alpha = P m``, hand-derived reverse mode, chain-wise
backward message with dead-step elimination.  Agreement of the two to ~1e-10 is
(a) an independent check on the oracle and (b) the proof that the kernel algebra
and its adjoint are the reference's mathematics.  Nothing in the product imports
this file.

Reference citations are to /root/reference (cbfssm/model/gp_tf.py, cbfssm.py).
"""
from __future__ import annotations

import math
import numpy as np

LOG_2PIE = math.log(2.0 * math.pi * math.e)


def softplus(x):
    return np.logaddexp(0.0, x) + 1e-10          # tf_transform.py:19-21


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------
# parameter-only prologue and its adjoint (gp_tf.py:104-130, 163-172)
# --------------------------------------------------------------------------
def gp_prologue(Z, m, Su, vu, lu):
    """raw tensors -> resident kernel operands + KL(q(u) || p(u))."""
    ell = softplus(lu)
    sig2 = float(softplus(vu))
    S = softplus(Su)
    Zt = Z / ell
    diff = Zt[:, None, :] - Zt[None, :, :]
    D2 = np.sum(diff * diff, axis=2)
    K0 = sig2 * np.exp(-0.5 * D2)                 # gp_tf.py:48-49
    K = K0 + 1e-8 * np.eye(Z.shape[0])            # gp_tf.py:52-53
    L = np.linalg.cholesky(K)
    P = np.linalg.inv(K)
    P = 0.5 * (P + P.T)
    alpha = P @ m
    M, Dout = m.shape
    kl = (Dout * np.sum(np.log(np.diag(L))) - 0.5 * np.sum(np.log(S))
          + 0.5 * (-M * Dout + np.sum(np.diag(P)[:, None] * S) + np.sum(m * alpha)))
    return dict(Z=Z, m=m, ell=ell, sig2=sig2, S=S, Zt=Zt, K0=K0, P=P, alpha=alpha, kl=kl,
                Su=Su, vu=vu, lu=lu)


def gp_prologue_adjoint(g, Pbar, alphabar, Sbar, Zbar, ellbar, sig2bar, kl_weight=1.0):
    """Chain kernel-level adjoints (w.r.t. P, alpha, S, Z, ell, sig2 as the rollout
    consumes them) plus ``kl_weight * dKL`` back to the five raw tensors."""
    P, m, S, alpha, Zt, ell, sig2, K0 = (g[k] for k in ("P", "m", "S", "alpha", "Zt", "ell", "sig2", "K0"))
    M, Dout = m.shape
    Pbar = Pbar + alphabar @ m.T                          # alpha = P m
    mbar = P @ alphabar
    # KL terms
    Sbar = Sbar + kl_weight * (0.5 * np.diag(P)[:, None] - 0.5 / S)
    mbar = mbar + kl_weight * alpha
    Pbar = Pbar + kl_weight * 0.5 * (np.diag(np.sum(S, axis=1)) + m @ m.T)
    Kbar = -P @ Pbar @ P + kl_weight * 0.5 * Dout * P     # P = K^-1 ; sum log L_ii = 0.5 logdet K
    E = Kbar * K0
    sig2bar = sig2bar + np.sum(E) / sig2
    D2bar = -0.5 * E
    W = D2bar + D2bar.T
    Ztbar = 2.0 * (np.sum(W, axis=1)[:, None] * Zt - W @ Zt)
    Zbar = Zbar + Ztbar / ell
    ellbar = ellbar - np.sum(Ztbar * Zt, axis=0) / ell
    return {
        "zeta_pos": Zbar,
        "zeta_mean": mbar,
        "zeta_var_unc": Sbar * sigmoid(g["Su"]),
        "variance_unc": np.asarray(sig2bar * sigmoid(g["vu"])),
        "lengthscales_unc": ellbar * sigmoid(g["lu"]),
    }


# --------------------------------------------------------------------------
# one sparse-GP evaluation and its reverse (gp_tf.py:132-161; SURVEY 8a notes 3-4)
# --------------------------------------------------------------------------
def gp_eval(g, X):
    """X [N, Din] -> fmean [N, Dout], fvar [N, Dout] (no process noise)."""
    Xt = X / g["ell"]
    d = Xt[:, None, :] - g["Zt"][None, :, :]
    k = g["sig2"] * np.exp(-0.5 * np.sum(d * d, axis=2))          # [N, M]
    a = k @ g["P"]                                                # [N, M]  (P symmetric)
    fmean = k @ g["alpha"]
    fvar = g["sig2"] - np.sum(k * a, axis=1)[:, None] + (a * a) @ g["S"]
    return fmean, fvar, (Xt, d, k, a)


def gp_eval_reverse(g, cache, gm, gv, acc):
    """Given d/dfmean (gm) and d/dfvar (gv) [N, Dout]: accumulate kernel-level
    parameter adjoints into ``acc`` and return d/dX [N, Din]."""
    Xt, d, k, a = cache
    G = np.sum(gv, axis=1)                                        # [N]
    c = gv @ g["S"].T                                             # [N, M]
    b = a * c
    abar = 2.0 * b - G[:, None] * k
    kbar = gm @ g["alpha"].T + 2.0 * (b @ g["P"]) - 2.0 * G[:, None] * a
    w = kbar * k
    acc["P"] += abar.T @ k
    acc["alpha"] += k.T @ gm
    acc["S"] += (a * a).T @ gv
    wd = w[:, :, None] * d                                        # [N, M, Din]
    acc["Z"] += np.sum(wd, axis=0) / g["ell"]
    acc["ell"] += np.sum(wd * d, axis=(0, 1)) / g["ell"]
    acc["sig2"] += np.sum(w) / g["sig2"] + np.sum(G)
    return -np.sum(wd, axis=1) / g["ell"]


def new_acc(M, Din, Dout):
    return dict(P=np.zeros((M, M)), alpha=np.zeros((M, Dout)), S=np.zeros((M, Dout)),
                Z=np.zeros((M, Din)), ell=np.zeros(Din), sig2=0.0)


# --------------------------------------------------------------------------
# backward-message chains (cbfssm.py:101-158; SURVEY 8a notes 1 and 5)
# --------------------------------------------------------------------------
def writer_run(t, R):
    """Which of the two runs writes y2[t] (cbfssm.py:125,128)."""
    return 0 if (t % (2 * R)) < R else 1


def build_chains(T, R):
    """Live chain segments: list of (run, t_hi, t_lo, init) with steps t = t_hi .. t_lo
    (descending); init in {"zero", "resample"}.  Steps after a segment's last
    written step are dead (feed nothing) and are dropped."""
    chains = []
    for run in (0, 1):
        off = 1 if run == 0 else R + 1
        starts = [T - 1] + [t for t in range(T - 2, -1, -1) if (t + off) % (2 * R) == 0]
        for i, t_hi in enumerate(starts):
            t_next = starts[i + 1] if i + 1 < len(starts) else -1
            resample = (t_hi + off) % (2 * R) == 0
            written = [t for t in range(t_hi, t_next, -1) if writer_run(t, R) == run]
            if not written:
                continue
            chains.append((run, t_hi, min(written), "resample" if resample else "zero"))
    return chains


# --------------------------------------------------------------------------
# full ELBO value + gradient in kernel algebra
# --------------------------------------------------------------------------
def elbo_value_and_grad(cfg, params, u, y, eps_b, z_b, eps_f, condition=True, want_grad=True):
    """Same contract as ``cbfssm_oracle.loss_and_grads`` (params: dict of NumPy arrays
    keyed like ``PARAM_NAMES``).  Particles are flattened n = b*S + s."""
    u = np.asarray(u, np.float64)
    y = np.asarray(y, np.float64)
    B, T, du = u.shape
    S_, dx, dy, R = cfg.samples, cfg.dim_x, cfg.dim_y, cfg.recog_len
    dh, Din, M = dx - dy, dx + du, cfg.ind_pnt_num
    N = B * S_
    kap = float(cfg.k_factor)
    lam1, lam2 = (float(v) for v in cfg.loss_factors)
    eb = np.asarray(eps_b, np.float64).reshape(2, T, N)
    zb = np.asarray(z_b, np.float64).reshape(2, T, N)
    ef = np.asarray(eps_f, np.float64).reshape(T - 1, N)
    un = np.repeat(u, S_, axis=0)          # [N, T, du]   (cbfssm.py:74-76 tiling)
    yn = np.repeat(y, S_, axis=0)

    gf = gp_prologue(params["f.zeta_pos"], params["f.zeta_mean"], params["f.zeta_var_unc"],
                     params["f.variance_unc"], params["f.lengthscales_unc"])
    gb = gp_prologue(params["b.zeta_pos"], params["b.zeta_mean"], params["b.zeta_var_unc"],
                     params["b.variance_unc"], params["b.lengthscales_unc"])
    var_x = softplus(params["var_x_unc"])
    var_y = softplus(params["var_y_unc"])

    # ---- backward-message chains: H[run, t] = `out` of that step ----
    chains = build_chains(T, R)
    H = np.full((2, T, N, dh), np.nan)
    entropy = 0.0

    def hidden_in(run, t, t_hi, init):
        if t == t_hi:
            return np.repeat(zb[run, t][:, None], dh, axis=1) if init == "resample" else np.zeros((N, dh))
        return H[run, t + 1]

    for (run, t_hi, t_lo, init) in chains:
        for t in range(t_hi, t_lo - 1, -1):
            hid = hidden_in(run, t, t_hi, init)
            fm, fv, _ = gp_eval(gb, np.concatenate((hid, un[:, t], yn[:, t]), axis=1))
            fv = fv + var_x[:dh]
            H[run, t] = fm + hid + eb[run, t][:, None] * np.sqrt(fv)
            if writer_run(t, R) == run:
                entropy += 0.5 * np.sum(LOG_2PIE + np.log(fv))
    y2 = np.stack([H[writer_run(t, R), t] for t in range(T)], axis=0)      # [T, N, dh]
    ytil = np.concatenate((np.transpose(yn, (1, 0, 2)), y2), axis=2)       # [T, N, dx]

    # ---- forward rollout ----
    X = np.zeros((T, N, dx))
    X[0] = ytil[0]
    kl_x = 0.0
    for t in range(T - 1):
        fm, fv, _ = gp_eval(gf, np.concatenate((X[t], un[:, t]), axis=1))
        fm = fm + X[t]
        fv = fv + var_x
        vy = var_y + (kap - 1.0) * fv
        s = vy + fv
        kg = fv / s
        mu = fm + kg * (ytil[t + 1] - fm)
        sig = (1.0 - kg) ** 2 * fv + kg ** 2 * vy
        do = bool(condition) or (t < R - 1)
        e = ef[t][:, None]
        X[t + 1] = (mu + e * np.sqrt(sig)) if do else (fm + e * np.sqrt(fv))
        if do:
            kl_x += 0.5 * np.sum(np.log(fv) - np.log(sig) + (sig + (mu - fm) ** 2) / fv - 1.0)

    sse = np.sum((np.transpose(yn, (1, 0, 2)) - X[:, :, :dy]) ** 2, axis=(0, 1))   # [dy]
    loglik = -0.5 * np.sum(sse / var_y[:dy]) - 0.5 * N * T * np.sum(np.log(var_y[:dy]) + math.log(2 * math.pi))
    elbo = (lam1 / S_) * (loglik - kl_x) + (lam2 / S_) * entropy - gf["kl"] - gb["kl"]
    out = dict(loss=-elbo, loglik=loglik, kl_x=kl_x, entropy=entropy, kl_z_f=gf["kl"], kl_z_b=gb["kl"],
               X=X, ytil=ytil, H=H, chains=chains)
    if not want_grad:
        return out, None

    # ---- reverse mode; weights are d loss / d term ----
    w_ll, w_kl, w_en = -lam1 / S_, lam1 / S_, -lam2 / S_
    accf, accb = new_acc(M, Din, dx), new_acc(M, Din, dh)
    vxbar = np.zeros(dx)
    vybar = np.zeros(dx)
    # likelihood: explicit var_y dependence
    vybar[:dy] += w_ll * (0.5 * sse / var_y[:dy] ** 2 - 0.5 * N * T / var_y[:dy])
    Ybar = np.zeros((T, N, dh))            # adjoint of y2[t]
    yT = np.transpose(yn, (1, 0, 2))

    def lik_grad(t):
        gx = np.zeros((N, dx))
        gx[:, :dy] = w_ll * (yT[t] - X[t, :, :dy]) / var_y[:dy]
        return gx

    xbar = lik_grad(T - 1)
    for t in range(T - 2, -1, -1):
        fm0, fv0, cache = gp_eval(gf, np.concatenate((X[t], un[:, t]), axis=1))
        fm = fm0 + X[t]
        fv = fv0 + var_x
        vy = var_y + (kap - 1.0) * fv
        s = vy + fv
        kg = fv / s
        yd = ytil[t + 1] - fm
        mu = fm + kg * yd
        sig = (1.0 - kg) ** 2 * fv + kg ** 2 * vy
        do = bool(condition) or (t < R - 1)
        e = ef[t][:, None]
        fmb = np.zeros((N, dx)); fvb = np.zeros((N, dx))
        if do:
            mub = xbar + w_kl * (mu - fm) / fv
            sigb = xbar * e * 0.5 / np.sqrt(sig) + w_kl * 0.5 * (1.0 / fv - 1.0 / sig)
            fvb += w_kl * 0.5 * (1.0 / fv - (sig + (mu - fm) ** 2) / fv ** 2)
            fmb += -w_kl * (mu - fm) / fv
            # sig = (1-kg)^2 fv + kg^2 vy
            kgb = sigb * (-2.0 * (1.0 - kg) * fv + 2.0 * kg * vy)
            fvb += sigb * (1.0 - kg) ** 2
            vyb = sigb * kg ** 2
            # mu = fm + kg * yd ; yd = ytil - fm
            fmb += mub * (1.0 - kg)
            kgb += mub * yd
            ytb = mub * kg
            # kg = fv / s ; s = vy + fv
            fvb += kgb / s
            sb = -kgb * fv / s ** 2
            vyb += sb
            fvb += sb
            # vy = var_y + (kap-1) fv
            vybar += np.sum(vyb, axis=0)
            fvb += (kap - 1.0) * vyb
            Ybar[t + 1] += ytb[:, dy:]
        else:
            fmb += xbar
            fvb += xbar * e * 0.5 / np.sqrt(fv)
        vxbar += np.sum(fvb, axis=0)
        xin_bar = gp_eval_reverse(gf, cache, fmb, fvb, accf)
        xbar = xin_bar[:, :dx] + fmb + lik_grad(t)
    Ybar[0] += xbar[:, dy:]                 # x_0 = y_tilde[:, 0]  (cbfssm.py:168)

    for (run, t_hi, t_lo, init) in chains:
        hbar = np.zeros((N, dh))
        for t in range(t_lo, t_hi + 1):
            hid = hidden_in(run, t, t_hi, init)
            fm0, fv0, cache = gp_eval(gb, np.concatenate((hid, un[:, t], yn[:, t]), axis=1))
            fv = fv0 + var_x[:dh]
            ob = hbar.copy()
            fvb = np.zeros((N, dh))
            if writer_run(t, R) == run:
                ob += Ybar[t]
                fvb += w_en * 0.5 / fv
            fvb += ob * eb[run, t][:, None] * 0.5 / np.sqrt(fv)
            vxbar[:dh] += np.sum(fvb, axis=0)
            xin_bar = gp_eval_reverse(gb, cache, ob, fvb, accb)
            hbar = xin_bar[:, :dh] + ob

    gF = gp_prologue_adjoint(gf, accf["P"], accf["alpha"], accf["S"], accf["Z"], accf["ell"], accf["sig2"], 1.0)
    gB = gp_prologue_adjoint(gb, accb["P"], accb["alpha"], accb["S"], accb["Z"], accb["ell"], accb["sig2"], 1.0)
    grads = {f"f.{k}": v for k, v in gF.items()}
    grads.update({f"b.{k}": v for k, v in gB.items()})
    grads["var_x_unc"] = vxbar * sigmoid(params["var_x_unc"])
    grads["var_y_unc"] = vybar * sigmoid(params["var_y_unc"])
    out["kernel_level"] = dict(f=accf, b=accb, var_x=vxbar, var_y=vybar)
    return out, grads
