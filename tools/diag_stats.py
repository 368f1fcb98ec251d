import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy import stats
from conftest import tiled_pinit
from mcpar_b200 import engine as eng
from oracle import mh

N = 1 << 14
e = eng.Engine(2, N, mode="normal", pl=1.0, thin=20, history_steps=400)
e.run(8000, 500, tiled_pinit(N, 2), "rosenbrock1")
h = e.history()
for a, b in [(0, 50), (50, 150), (150, 400)]:
    r = h[a:b].reshape(-1, 3)
    print("rosen local", a, b, r[:, 0].mean(), r[:, 1].mean(), r[:, 0].var(), r[:, 1].var(), np.cov(r[:, 0], r[:, 1])[0, 1])
print("factor", e.factor(), "acc", e.stats()["accepted"] / e.stats()["tried"])
mean, cov = e.moments(); hh = h.reshape(-1, 3)
print("moments dev", mean, cov, "host", hh[:, :2].mean(0), np.cov(hh[:, :2].T, bias=True))
e.close()

N = 1 << 13
e = eng.Engine(2, N, mode="normal", pl=1.0, thin=400, history_steps=3)
e.run(1200, 500, np.zeros((N, 2)), "dualgaussian", [5.0])
x = e.history()[2, :, 0]
print("dgauss local: mean var", x.mean(), x.var(), "frac>2.4", (np.abs(x) > 2.4).mean(), "factor", e.factor())
xx = x[np.abs(x) < 2.4]
print(stats.kstest(xx, stats.truncnorm(-2.4, 2.4).cdf))
e.close()

N, M = 2048, 16
pin = tiled_pinit(N, 2)
for seed in (12345, 777):
    o = mh.run_counter("dualgaussian", 2, N, 400, 300, pin, par=[5.0], pool_m=M, seed=seed, thin=100)
    c = o["rows"][3, :, :2]
    print("oracle seed", seed, "mean", c.mean(0), "var", c.var(0), "frac mode2", (c[:, 0] > 2.5).mean())
for seed in (8675309, 4242):
    e = eng.Engine(2, N, mode="normal", pool_m=M, thin=100, history_steps=4, seed=seed)
    e.run(400, 300, pin, "dualgaussian", [5.0])
    g = e.history()[3, :, :2]
    print("gpu seed", seed, "mean", g.mean(0), "var", g.var(0), "frac mode2", (g[:, 0] > 2.5).mean())
    e.close()
o = mh.run_counter("dualgaussian", 2, N, 400, 300, pin, par=[5.0], pool_m=M, seed=8675309, thin=100)
e = eng.Engine(2, N, mode="normal", pool_m=M, thin=100, history_steps=4, seed=8675309)
e.run(400, 300, pin, "dualgaussian", [5.0])
g = e.history()
print("same seed: max |diff| rows", np.abs(g - o["rows"]).max(), "frac rows equal(1e-6)", np.all(np.isclose(g, o["rows"], atol=1e-6), axis=-1).mean())
