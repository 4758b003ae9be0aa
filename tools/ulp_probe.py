import sys, numpy as np
sys.path.insert(0, '.')
from microclimc_b200._lib import Context
from tests.test_gpu_api import _ulp_err
ctx = Context(20, 10, 1, [0]); rng = np.random.default_rng(5)
x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-5, 5, 200000), rng.uniform(-1e-3, 1e-3, 50000)])
p = np.concatenate([rng.uniform(1e-6, 1e6, 200000), np.exp(rng.uniform(-300, 300, 100000)), rng.uniform(0.5, 2.0, 100000)])
q = np.concatenate([rng.uniform(-1e3, 1e3, 300000), rng.uniform(-1, 1, 100000)])
print("exp", _ulp_err(ctx.eval_primitive("exp", x), np.exp(x.astype(np.longdouble))))
print("log", _ulp_err(ctx.eval_primitive("log", p), np.log(p.astype(np.longdouble))))
print("rcp", _ulp_err(ctx.eval_primitive("rcp", p), 1 / p.astype(np.longdouble)))
print("sqrt", _ulp_err(ctx.eval_primitive("sqrt", p), np.sqrt(p.astype(np.longdouble))))
print("div", _ulp_err(ctx.eval_primitive("div", q, p), q.astype(np.longdouble) / p.astype(np.longdouble)))
