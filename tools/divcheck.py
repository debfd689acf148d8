import sys; sys.path.insert(0,'.')
import numpy as np
from spotify_recommender_b200.engine import Engine
eng=Engine(0)
rng = np.random.default_rng(11)
n = 4_000_000
a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
b = np.abs(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32))
b[~(b > 0)] = 1.0
a[:1_000_000] = (rng.random(1_000_000, dtype=np.float32) * 2 - 1) * b[:1_000_000]
a[1_000_000:1_200_000] = b[1_000_000:1_200_000] * np.float32(1e-38) * rng.random(200_000, dtype=np.float32)
edge = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-45, -1e-45, 3.4028235e38, 1.1754944e-38], np.float32)
eb = np.array([1e-8, 1.0, 3.0, np.inf, 3.4028235e38, 1.1754944e-38, 1e-45, 7.0, 0.1], np.float32)
a = np.concatenate([a, np.repeat(edge, eb.size)])
b = np.concatenate([b, np.tile(eb, edge.size)])
got = eng.selftest_div(a, b)
with np.errstate(all="ignore"):
    want = a / b
nan = np.isnan(want)
print("nan mismatch", np.sum(np.isnan(got)!=nan))
bad = np.where((np.isnan(got)!=nan) | ((~nan) & (got.view(np.uint32)!=want.view(np.uint32))))[0]
print(len(bad), "bad of", a.size)
for i in bad[:40]:
    print(i, a[i], b[i], got[i], want[i], hex(a.view(np.uint32)[i]), hex(b.view(np.uint32)[i]))
# categorize
bb=b[bad]; aa=a[bad]
print("b subnormal:", np.sum(bb<1.1754944e-38), "a subnormal:", np.sum(np.abs(aa)<1.1754944e-38), "want subnormal", np.sum(np.abs(want[bad])<1.1754944e-38), "want inf", np.sum(np.isinf(want[bad])))
