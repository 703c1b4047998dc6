import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import cbfssm_oracle as O
from tests.helpers import make_problem, rel_inf
from tests.test_gpu_parity import run_engine, _cond_kzz
for zp in (2.0, 1.0, 0.3):
    cfg, params, u, y, eb, zb, ef = make_problem(2, 1, 1, 100, 20, 3, 16, 4, 1.0, (10.0, 0.5), seed=21, strong=True, zeta_pos=zp)
    f = lambda a: np.asarray(a, np.float32).astype(np.float64)
    u, y, eb, zb, ef = f(u), f(y), f(eb), f(zb), f(ef)
    res, gd = O.loss_and_grads(cfg, params, u, y, eb, zb, ef, True)
    for flags in (12, 1, 128):
        eng, out, yd = run_engine(cfg, params, u, y, eb, zb, ef, True, flags)
        g = eng.get_grads()
        errs = {k: rel_inf(g[k], gd[k].numpy()) for k in O.PARAM_NAMES}
        le = abs(float(out["loss"]) - float(res.loss.detach())) / abs(float(res.loss.detach()))
        print(f"zeta_pos={zp} cond={_cond_kzz(params,'f'):.1e}/{_cond_kzz(params,'b'):.1e} flags={flags} loss {le:.1e} " + " ".join(f"{k.split('.')[0][0]}.{k.split('.')[-1][:6]}={v:.1e}" for k, v in errs.items()))
