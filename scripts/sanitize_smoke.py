"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmm_training_b200 import _lib, engine, synthetic

_lib.init(0)
rng = np.random.default_rng(0)
# VQ + LBG
X = synthetic.mfcc_mixture(0, 3000, K=16)
C = synthetic.random_codebook(1, 256)
engine.vq_encode(X, C)
engine.lbg_fit(X, 16, 5, 1e-3)
# N = 4 (both A structures), ragged lengths, a few iterations
for dense in (False, True):
    corpus = synthetic.word_corpus(1, 3, 45, tmin=1, tmax=60)
    obs, off, wos = synthetic.pack_corpus(corpus, 256)
    pi0, A0, B0 = engine.default_init(4, 256)
    A0 = A0.copy()
    if dense:
        A0 = rng.dirichlet(np.ones(4), size=4)
    W = 3
    out = engine.bw_fit(obs, off, wos, W, 4, 256, np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)), max_iterations=3)
    engine.score(obs, off, 4, 256, out[0], out[1], out[2])
# left-to-right N = 16 / 8 and generic N = 6, 16 dense
for N, M, dense in ((16, 1024, False), (8, 64, False), (6, 32, True), (16, 300, True)):
    corpus = [synthetic.clustered_sequences(rng, 37, N=N, M=M, tmin=1, tmax=50, spread=max(2, M // (2 * N))) for _ in range(2)]
    obs, off, wos = synthetic.pack_corpus(corpus, M)
    pi0, A0, B0 = engine.default_init(N, M)
    if dense:
        A0 = rng.dirichlet(np.ones(N), size=N)
    W = 2
    with engine.BaumWelch(obs, off, wos, W, N, M) as bw:
        bw.set_params(np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)))
        fam = bw.kernel_family()
        bw.iterate(2, 1e-6, 2)
        pi, A, B = bw.params()
    engine.score(obs, off, N, M, pi, A, B)
    print(N, M, fam, "ok")
print("sanitize smoke done, launches", _lib.load().hmmb_launch_count())
