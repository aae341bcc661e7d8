import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from hmm_training_b200 import engine, synthetic
from oracle import hmm_oracle as O
from test_gpu_parity import _tiny_models
np.set_printoptions(linewidth=200, precision=4)
N, M = 4, 16
rng = np.random.default_rng(5 + N)
W = 3
pi0, A0, B0 = _tiny_models(rng, W, N, M)
corpus = [[rng.integers(0, M, size=int(rng.integers(2, 30))) for _ in range(30)] for _ in range(W)]
obs, offsets, wos = synthetic.pack_corpus(corpus, M)
with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:
    bw.set_params(pi0, A0, B0)
    bw.iterate(1, 1e-6, 5)
    ll = bw.seq_ll()
    pi, A, B = bw.params(finalize=False)
w = 0
Ao, Bo, pio, h, it = O.hmm_training(corpus[w], N=N, M=M, max_iterations=1, init=(pi0[w], A0[w], B0[w]), return_history=True)
obs_p, lens, valid = O._pad(corpus[w])
la, logP, _ = O.forward_log(obs_p, lens, O.safe_log(pi0[w]), O.safe_log(A0[w]), O.safe_log(B0[w]))
print("seq ll gpu", ll[:30]); print("seq ll ref", logP)
print("pi gpu", pi[w], "ref(norm)", pio)
print("A gpu\n", A[w], "\nA ref (normalised)\n", Ao)
Bn = B[w] / B[w].sum(axis=1, keepdims=True)
print("B gpu (rownorm)\n", Bn, "\nB ref\n", Bo)
