"""Small-call latency of the host API (the live-recognition use of the reference: one utterance at a time)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmm_training_b200 import _lib, engine, synthetic
_lib.init(0)
rng = np.random.default_rng(0)
C = synthetic.random_codebook(3, 256)
X = rng.normal(size=(100, 13))
pi, A, B = engine.default_init(4, 256)
W = 10
Bm = rng.dirichlet(np.ones(256) * 0.3, size=(W, 4)); pim, Am = np.tile(pi, (W, 1)), np.tile(A, (W, 1, 1))
obs = rng.integers(0, 256, size=100).astype(np.uint8); off = np.array([0, 100], dtype=np.int64)
def timeit(f, n=200):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
print("vq_encode 100 frames x 256 centroids: %.1f us" % timeit(lambda: engine.vq_encode(X, C)))
print("score 1 utterance (T=100) x 10 models: %.1f us" % timeit(lambda: engine.score(obs, off, 4, 256, pim, Am, Bm)))
o1, f1, w1 = synthetic.fixed_length_codewords(1, 10, 20, 100, 4, 256)
p0, a0, b0 = np.tile(pi, (10, 1)), np.tile(A, (10, 1, 1)), np.tile(B, (10, 1, 1))
print("bw_fit config 1 (10 words x 20 utterances x T=100), 10 iterations: %.1f us" % timeit(lambda: engine.bw_fit(o1, f1, w1, 10, 4, 256, p0, a0, b0, max_iterations=10, epsilon=-1.0), 50))
print("bw_fit same, 1 iteration: %.1f us" % timeit(lambda: engine.bw_fit(o1, f1, w1, 10, 4, 256, p0, a0, b0, max_iterations=1, epsilon=-1.0), 50))
