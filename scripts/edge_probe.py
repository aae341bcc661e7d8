"""Edge shapes through the training / scoring API next to the oracle: a word without sequences, R = 0, one state, one
codeword, 32 states with a 65536-codeword alphabet, every sequence impossible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmm_training_b200 import engine
from oracle import hmm_oracle as O

rng = np.random.default_rng(3)

def fit(seqs, wos, W, N, M, init, iters=2):
    obs = np.concatenate(seqs) if seqs else np.zeros(0, np.int64)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    return engine.bw_fit(obs, off, np.asarray(wos, np.int32), W, N, M, *init, max_iterations=iters, epsilon=-1.0)

def rand_init(W, N, M):
    return (rng.dirichlet(np.ones(N), size=W), rng.dirichlet(np.ones(N), size=(W, N)), rng.dirichlet(np.ones(M), size=(W, N)))

def check(name, seqs, wos, W, N, M, init, iters=2):
    pi, A, B, hist, it = fit(seqs, wos, W, N, M, init, iters)
    worst = 0.0
    for w in range(W):
        mine = [seqs[r] for r in range(len(seqs)) if wos[r] == w]
        if not mine:
            same = np.array_equal(pi[w], init[0][w]) and np.array_equal(A[w], init[1][w]) and np.array_equal(B[w], init[2][w])
            print(f"  {name}: word {w} has no sequences -> parameters unchanged: {same}, iterations {it[w]}, history {hist[w][:2]}")
            continue
        Ao, Bo, pio, ho, _ = O.hmm_training(mine, N=N, M=M, epsilon=-1.0, max_iterations=iters, init=(init[0][w], init[1][w], init[2][w]), return_history=True)
        for x, y in ((A[w], Ao), (B[w], Bo), (pi[w], pio), (hist[w, :iters], np.array(ho))):
            with np.errstate(invalid="ignore", divide="ignore"):
                d = np.abs(x - y) / np.maximum(np.abs(y), 1e-300)
            d = np.where((x == y) | (np.isnan(x) & np.isnan(y)), 0.0, d)
            worst = max(worst, float(np.nanmax(d)))
    print(f"{name}: worst relative difference to the oracle {worst:.2e}")

W, N, M = 3, 4, 16
seqs = [rng.integers(0, M, size=int(rng.integers(3, 20))) for _ in range(10)]
wos = [0] * 5 + [2] * 5
check("word 1 without sequences (N=4)", seqs, wos, W, N, M, rand_init(W, N, M))
check("word 1 without sequences (N=6)", [s % 12 for s in seqs], wos, W, 6, 12, rand_init(W, 6, 12))
check("one state, one codeword", [np.zeros(5, np.int64), np.zeros(1, np.int64)], [0, 0], 1, 1, 1, (np.ones((1, 1)), np.ones((1, 1, 1)), np.ones((1, 1, 1))))
check("two states, two codewords", [rng.integers(0, 2, size=9) for _ in range(4)], [0] * 4, 1, 2, 2, rand_init(1, 2, 2))
check("32 states, 65536 codewords", [rng.integers(0, 65536, size=12) for _ in range(3)], [0] * 3, 1, 32, 65536, rand_init(1, 32, 65536), iters=1)
# every sequence impossible: state 0 is the only entry and cannot emit codeword 1
pi0 = np.array([[1.0, 0, 0, 0]]); A0 = np.array([[[.5, .5, 0, 0], [0, .5, .5, 0], [0, 0, .5, .5], [0, 0, 0, 1.0]]])
B0 = np.full((1, 4, 3), 1 / 3); B0[0, 0] = [1.0, 0.0, 0.0]
try:
    check("every sequence impossible", [np.array([1, 0, 2]), np.array([1, 1])], [0, 0], 1, 4, 3, (pi0, A0, B0))
except Exception as e:
    print("every sequence impossible ->", type(e).__name__, e)
try:
    out = fit([], [], 2, 4, 16, rand_init(2, 4, 16))
    print("R = 0: iterations", out[4], "history", out[3][:, :1].ravel())
except Exception as e:
    print("R = 0 ->", type(e).__name__, e)
# what a word without sequences comes back as (the reference: A = 0, B = 0, pi = NaN, -inf statistic every iteration)
pi, A, B, hist, it = fit(seqs, wos, 3, 4, 16, rand_init(3, 4, 16), iters=3)
print("word without sequences: A", np.unique(A[1]), "B", np.unique(B[1]), "pi", pi[1], "history", hist[1], "iterations", it[1])
pi, A, B, hist, it = fit([], [], 1, 4, 16, rand_init(1, 4, 16), iters=3)
print("R = 0: A", np.unique(A[0]), "B", np.unique(B[0]), "pi", pi[0], "history", hist[0], "iterations", it[0])
# recognition against degenerate models (what training returns for a word without sequences, negative / NaN entries):
# the reference's safe_log maps everything that is not > 0 to -inf, NaN included
U = [rng.integers(0, 16, size=int(rng.integers(1, 12))) for _ in range(9)]
obs = np.concatenate(U); off = np.concatenate([[0], np.cumsum([len(u) for u in U])]).astype(np.int64)
for N in (4, 6, 8, 16):
    pi_, A_, B_ = rand_init(4, N, 16)
    if N in (8, 16):  # bidiagonal: the left-to-right scorer
        A_ = np.zeros((4, N, N))
        for i in range(N):
            A_[:, i, i] = 0.6 if i + 1 < N else 1.0
            if i + 1 < N:
                A_[:, i, i + 1] = 0.4
    pi_[1] = np.nan; A_[1] = 0.0; B_[1] = 0.0          # the model of a word without sequences
    B_[2] = -B_[2]                                       # negative entries
    pi_[3, 0] = np.nan                                   # one NaN entry
    ll, arg = engine.score(obs, off, N, 16, pi_, A_, B_)
    want = O.score_batch(U, [(A_[w], B_[w], pi_[w]) for w in range(4)])
    with np.errstate(invalid="ignore"):
        ok = np.array_equal(np.isneginf(ll), np.isneginf(want)) and not np.isnan(ll).any() and \
            np.allclose(np.where(np.isneginf(want), 0, ll), np.where(np.isneginf(want), 0, want), rtol=1e-9, atol=0)
    print(f"degenerate models, N = {N}: matches the oracle: {ok}; argmax equal: {np.array_equal(arg, O.argmax_first(want))}")
    if not ok:
        print(ll[:3]); print(want[:3])
# training FROM a degenerate model (warm start from the file of a word that had no sequences; one NaN; negative entries)
for N in (4, 6, 8):
    M = 12
    seqs2 = [rng.integers(0, M, size=int(rng.integers(2, 15))) for _ in range(6)]
    init = rand_init(3, N, M)
    init[0][0] = np.nan; init[1][0] = 0.0; init[2][0] = 0.0
    init[0][1, 0] = np.nan
    init[2][2, 0] = -init[2][2, 0]
    try:
        check(f"degenerate initial models, N = {N}", seqs2 * 3, [0] * 6 + [1] * 6 + [2] * 6, 3, N, M, init, iters=3)
    except Exception as e:
        print(f"degenerate initial models, N = {N} ->", type(e).__name__, e)
# codebook build on edge inputs next to the C oracle (which matches the reference on every one of them, checked on the CPU)
from oracle import vq_oracle as V
from hmm_training_b200 import synthetic as S
X = S.mfcc_mixture(3, 200, K=8)
Xn = X.copy(); Xn[3, 4] = np.nan
for name, Xc, K, mi, eps in (("max_iterations = 0", X, 8, 0, 1e-3), ("max_iterations = 1", X, 8, 1, 1e-3), ("epsilon = 0", X, 8, 20, 0.0),
                             ("K = 2", X, 2, 20, 1e-3), ("K = 3", X, 3, 20, 1e-3), ("K > frames", X[:5], 16, 20, 1e-3),
                             ("one frame", X[:1], 4, 20, 1e-3), ("identical frames", np.tile(X[:1], (30, 1)), 8, 20, 1e-3),
                             ("a NaN coordinate", Xn, 4, 5, 1e-3)):
    try:
        got = engine.lbg_fit(Xc, K, mi, eps)
        want = V.lbg(Xc, K, mi, eps)
        ok = got[0].shape == want[0].shape and np.allclose(got[0], want[0], rtol=1e-9, atol=0, equal_nan=True) and \
            np.array_equal(got[2][:len(Xc)], want[2]) and list(got[3]) == list(want[3])
        print(f"LBG {name}: matches the oracle: {ok}; iterations {list(got[3])} vs {list(want[3])}")
    except Exception as e:
        print(f"LBG {name} ->", type(e).__name__, e)
for K in (0, -1):
    try:
        engine.lbg_fit(X, K, 5, 1e-3); print(f"LBG K = {K}: no error")
    except Exception as e:
        print(f"LBG K = {K} ->", type(e).__name__, e)
# VQ encode on edge inputs next to the C oracle (which matches the reference on all of them but "no centroids")
X = S.mfcc_mixture(0, 4000, K=8); C = S.random_codebook(1, 256)
Xn = X.copy(); Xn[2, 5] = np.nan; Xn[3, 0] = np.nan; Xn[4, 1] = np.inf; Xn[5, 2] = -np.inf; Xn[100:200, 7] = np.nan
Cn = C.copy(); Cn[0, 3] = np.nan; Cn[7] = np.inf; Cn[2, 0] = np.nan
for name, Xc, Cc in (("NaN / inf in frames", Xn, C), ("NaN / inf in centroids", X, Cn), ("every centroid NaN", X, np.full((4, 13), np.nan)),
                     ("one centroid", X, C[:1]), ("huge values", X * 1e200, C * 1e200), ("large values (fp32 overflow only)", X * 1e30, C * 1e30),
                     ("tiny values", X * 1e-200, C * 1e-200), ("small values (fp32 underflow only)", X * 1e-30, C * 1e-30)):
    try:
        got = engine.vq_encode(Xc, Cc); want = V.encode(Xc, Cc)
        print(f"VQ {name}: equal to the oracle: {np.array_equal(got, want)} ({int((got != want).sum())} of {len(want)} differ)")
    except Exception as e:
        print(f"VQ {name} ->", type(e).__name__, e)
try:
    print("VQ no centroids ->", engine.vq_encode(X[:5], C[:0]))
except Exception as e:
    print("VQ no centroids ->", type(e).__name__, e)
