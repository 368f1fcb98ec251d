"""One-off check quoted in DESIGN.md section 4: does the reference build (oracle/_ref) keep the DualGaussian target
invariant from a stationary start?  PLOCAL 1: yes; PLOCAL 0.9: the small mode loses mass (unnormalised Q_i in cfac)."""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle.ref import Ref
ref = Ref(64)
R, C = 8, 8                      # 64 chains, all-pairs remote proposals over 64 components (the reference's algorithm)
res = {}
for pl in (1.0, 0.9):
    xs = []
    t0 = time.time()
    for seed in range(1500):
        rng = np.random.default_rng(1000 + seed)
        comp = rng.random(R * C) < 1.0 / 6.0
        pin = rng.standard_normal((R * C, 2)) + 5.0 * comp[:, None]
        o = ref.run("dualgaussian", 2, C, R, 300, 100, pin, par=[5.0], seed=seed + 1, pl=pl, want_maxl=False)
        rows = np.asarray(o["rows"]).reshape(-1, 3)
        # last 100 steps of every chain
        xs.append(rows[-100 * R * C:, 0])
    x = np.concatenate(xs)
    res[pl] = (x.mean() - 5.0 / 6.0, (x > 2.5).mean(), time.time() - t0)
    print("reference build, pl=%.1f: mean - 5/6 = %+.4f   P(x0>2.5) = %.4f (exact 0.1708)   [%.0f s]" % (pl, *res[pl]))
