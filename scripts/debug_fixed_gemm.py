import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import largecrimsoncanine as L
np.set_printoptions(linewidth=220, precision=3, suppress=True)
sig = (7, 0, 0); nb = 128; n = 256
alg = L.Algebra(*sig)
x = (np.arange(n)[:, None] * 100 + np.arange(nb)[None, :]).astype(np.float32)  # x[r,k] = 100 r + k
m = np.zeros((1, nb), np.float32); m[0, 0] = 1.0
M, X = L.MultivectorBatch.from_numpy(alg, m), L.MultivectorBatch.from_numpy(alg, x)
got = M.geometric_product(X).to_numpy()
print("identity: equal?", np.array_equal(got, x), "zeros frac", (got == 0).mean())
print(got[:6, :12]); print(got[30:36, :12]); print(got[126:132, :12])
bad = np.argwhere(got != x)
print("num bad", len(bad), "first bad", bad[:10].tolist())
# which rows / cols are right?
print("rows fully right:", np.where((got == x).all(axis=1))[0][:40])
print("cols fully right:", np.where((got == x).all(axis=0))[0][:64])
