"""Property tests (hypothesis): invariants of Baum-Welch re-estimation that hold for ANY input,
checked on the CPU oracle (not gpu) and on the CUDA path (gpu) with shapes and data drawn by hypothesis:

  * rows of A, B and pi sum to 1, or are all zero where nothing was observed (HMM/hmm_training.py:524-539);
  * the log-likelihood never decreases (up to the 1e-20-floor artefact, :493-497).  The statistic the reference
    tracks is log_sum_exp_r log P_r (:503), which EM does not maximise — the sum over r is what it maximises —
    so the property is asserted where the two coincide: words trained on a single sequence;
  * zeros of A and pi stay zeros (:415-455);
  * permuting the sequences leaves the result unchanged to 1e-12;
  * G "virtual ranks" on one device — every rank's accumulators captured by the all-reduce hook, summed on the
    host and injected — give the single-rank result to 1e-12, which checks the whole
    shard / reduce / replicated-M-step logic on a box with a single GPU.
"""
import ctypes

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from helpers import assert_close, assert_same_support
from oracle import hmm_oracle as O

# The committed runs are derandomised (the same examples every time: a round-end run must not depend on luck);
# HMMB_HYP_RANDOM=1 explores fresh examples, HMMB_HYP_EXAMPLES=n sets how many.
import os

_RANDOM = bool(os.environ.get("HMMB_HYP_RANDOM"))
_SCALE = int(os.environ.get("HMMB_HYP_EXAMPLES", "0"))


def _settings(n):
    return settings(max_examples=_SCALE or n, deadline=None, derandomize=not _RANDOM, suppress_health_check=list(HealthCheck))


def _random_model(rng, W, N, M, left_to_right, zero_frac):
    pi = rng.random((W, N)) + 0.05
    A = rng.random((W, N, N)) + 0.05
    if left_to_right:
        A *= (np.triu(np.ones((N, N))) - np.triu(np.ones((N, N)), 2))[None]
        pi[:, 1:] = 0.0
    elif zero_frac > 0:
        mask = rng.random((W, N, N)) < zero_frac
        mask[:, np.arange(N), np.arange(N)] = False  # every state keeps its self loop
        A[mask] = 0.0
    B = rng.random((W, N, M)) ** 3 + 1e-6
    pi /= pi.sum(axis=1, keepdims=True)
    A /= A.sum(axis=2, keepdims=True)
    B /= B.sum(axis=2, keepdims=True)
    return pi, A, B


def _random_corpus(rng, W, S, Tmin, Tmax, M):
    seqs, wos = [], []
    for w in range(W):
        lens = np.sort(rng.integers(Tmin, Tmax + 1, size=S))[::-1]
        for T in lens:
            seqs.append(rng.integers(0, M, size=int(T)).astype(np.int64))
            wos.append(w)
    return seqs, np.array(wos, dtype=np.int32)


def _pack(seqs):
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    return np.concatenate(seqs).astype(np.uint16), off


def _check_invariants(pi, A, B, hist, iters, pi0, A0, what, monotone):
    # a row is normalised to 1 unless nothing was ever observed for it (state never occupied before a sequence's
    # last step, word without a possible sequence): then it is all zeros (:524-539)
    for x, name in ((A, "rows of A"), (B, "rows of B"), (pi, "pi")):
        rs = x.sum(axis=-1)
        assert ((rs == 0) | (np.abs(rs - 1.0) <= 1e-12)).all(), f"{what} {name}: sums {rs}"
    assert not (A[A0 == 0] != 0).any(), what + ": a zero of A came alive"
    assert not (pi[pi0 == 0] != 0).any(), what + ": a zero of pi came alive"
    for w in range(hist.shape[0] if monotone else 0):
        h = hist[w, :int(iters[w])]
        h = h[np.isfinite(h)]
        # EM is monotone; the 1e-20 floor of unseen codewords (:493-497) may cost a little likelihood once
        assert (np.diff(h) >= -1e-6 * np.maximum(1.0, np.abs(h[:-1]))).all(), f"{what}: statistic decreased for word {w}: {h}"


shape = st.tuples(st.sampled_from([2, 3, 4, 5, 8, 16]), st.integers(3, 40), st.integers(1, 3), st.sampled_from([1, 1, 2, 5, 9]),
                  st.booleans(), st.sampled_from([0.0, 0.3]), st.integers(0, 2 ** 31 - 1))


@_settings(25)
@given(shape)
def test_oracle_reestimation_invariants(p):
    N, M, W, S, ltr, zf, seed = p
    rng = np.random.default_rng(seed)
    pi0, A0, B0 = _random_model(rng, W, N, M, ltr, zf)
    seqs, wos = _random_corpus(rng, W, S, 1, 12, M)
    perm = rng.permutation(len(seqs))
    for w in range(W):
        mine = [seqs[r] for r in range(len(seqs)) if wos[r] == w]
        A, B, pi, hist, it = O.hmm_training(mine, N=N, M=M, epsilon=-1.0, max_iterations=4, init=(pi0[w], A0[w], B0[w]),
                                            return_history=True)
        _check_invariants(pi[None], A[None], B[None], np.array(hist)[None], [it], pi0[w][None], A0[w][None], "oracle", S == 1)
        shuffled = [seqs[r] for r in perm if wos[r] == w]
        A2, B2, pi2 = O.hmm_training(shuffled, N=N, M=M, epsilon=-1.0, max_iterations=4, init=(pi0[w], A0[w], B0[w]))
        assert_close(A2, A, "oracle A under permutation", rtol=1e-12)
        assert_close(B2, B, "oracle B under permutation", rtol=1e-12)
        assert_close(pi2, pi, "oracle pi under permutation", rtol=1e-12)


gpu_shape = st.tuples(st.sampled_from([2, 3, 4, 4, 6, 8, 16]), st.sampled_from([5, 16, 64, 256, 300]), st.integers(1, 4),
                      st.sampled_from([1, 1, 3, 20, 70]), st.booleans(), st.sampled_from([0.0, 0.3]), st.integers(0, 2 ** 31 - 1))


@pytest.mark.gpu
@_settings(20)
@given(gpu_shape)
def test_gpu_reestimation_invariants_and_oracle(p):
    from hmm_training_b200 import engine
    N, M, W, S, ltr, zf, seed = p
    rng = np.random.default_rng(seed)
    pi0, A0, B0 = _random_model(rng, W, N, M, ltr, zf)
    seqs, wos = _random_corpus(rng, W, S, 1, 60, M)
    obs, off = _pack(seqs)
    pi, A, B, hist, iters = engine.bw_fit(obs, off, wos, W, N, M, pi0, A0, B0, epsilon=-1.0, max_iterations=4)
    _check_invariants(pi, A, B, hist, iters, pi0, A0, "gpu", S == 1)
    for w in range(W):
        mine = [seqs[r] for r in range(len(seqs)) if wos[r] == w]
        Ao, Bo, pio, ho, _ = O.hmm_training(mine, N=N, M=M, epsilon=-1.0, max_iterations=4, init=(pi0[w], A0[w], B0[w]),
                                            return_history=True)
        assert_close(A[w], Ao, "A"); assert_close(B[w], Bo, "B"); assert_close(pi[w], pio, "pi")
        assert_same_support(A[w], Ao, "A"); assert_same_support(pi[w], pio, "pi")
        # (a log-likelihood that is 0 up to rounding — P(O) = 1 after re-estimation on one short sequence — has no
        # meaningful relative error: absolute 1e-12 beside the relative 1e-9)
        assert_close(hist[w, :4], np.array(ho), "statistic", atol=1e-12)
    # permutation of the sequences
    perm = rng.permutation(len(seqs))
    obs2, off2 = _pack([seqs[r] for r in perm])
    pi2, A2, B2, hist2, _ = engine.bw_fit(obs2, off2, wos[perm], W, N, M, pi0, A0, B0, epsilon=-1.0, max_iterations=4)
    assert_close(A2, A, "A under permutation", rtol=1e-12); assert_close(B2, B, "B under permutation", rtol=1e-12)
    assert_close(pi2, pi, "pi under permutation", rtol=1e-12); assert_close(hist2, hist, "statistic under permutation", rtol=1e-12, atol=1e-12)


def _one_iteration(engine, seqs, wos, W, N, M, init, rank=0, world=1, hook=None):
    obs, off = _pack(seqs)
    with engine.BaumWelch(obs, off, wos, W, N, M) as bw:
        bw.set_params(*init)
        if world > 1:
            bw.set_dist(rank, world, hook)
        bw.iterate(1, -1.0, 1)
        return bw.params() + bw.history(1)


@pytest.mark.gpu
@pytest.mark.parametrize("N,M,G", [(4, 256, 3), (4, 40, 2), (16, 1024, 4), (8, 64, 2), (5, 33, 3)])
def test_virtual_ranks_on_one_device_equal_single_rank(N, M, G):
    """Shard-and-sum on one GPU: rank r runs its shard with a hook that copies the accumulator buffer (device
    pointer, n doubles) to the host; rank 0 then runs again with a hook that overwrites its buffer with the sum of
    all captured ones — exactly what the sum-all-reduce leaves on every rank."""
    from hmm_training_b200 import dist, engine
    rng = np.random.default_rng(100 * N + G)
    W, S = 3, 37
    init = _random_model(rng, W, N, M, N in (8, 16), 0.0)
    seqs, wos = _random_corpus(rng, W, S, 5, 50, M)
    want = _one_iteration(engine, seqs, wos, W, N, M, init)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaDeviceSynchronize.argtypes = []
    captured = []

    def capture(ptr, n):
        buf = np.empty(n)
        assert rt.cudaDeviceSynchronize() == 0
        assert rt.cudaMemcpy(buf.ctypes.data, ptr, n * 8, 2) == 0  # device -> host
        captured.append(buf)

    shards = [dist.shard_sequences_round_robin(wos, r, G) for r in range(G)]
    for r in range(G):
        _one_iteration(engine, [seqs[i] for i in shards[r]], wos[shards[r]], W, N, M, init, r, G, capture)
    assert len(captured) == G and len({len(c) for c in captured}) == 1
    total = np.sum(captured, axis=0)

    def inject(ptr, n):
        assert n == len(total)
        assert rt.cudaDeviceSynchronize() == 0
        assert rt.cudaMemcpy(ptr, total.ctypes.data, n * 8, 1) == 0  # host -> device

    got = _one_iteration(engine, [seqs[i] for i in shards[0]], wos[shards[0]], W, N, M, init, 0, G, inject)
    for x, y, name in zip(got, want, ("pi", "A", "B", "statistic", "iterations")):
        assert_close(x, y, f"{G} virtual ranks: {name}", rtol=1e-12, atol=1e-13)


# ---------------------------------------------------------------- VQ encode and recognition
vq_shape = st.tuples(st.integers(1, 700), st.sampled_from([1, 2, 7, 31, 32, 33, 256, 300, 600]), st.sampled_from([0, 1, 2, 3]),
                     st.integers(0, 2 ** 31 - 1))


@pytest.mark.gpu
@_settings(25)
@given(vq_shape)
def test_gpu_vq_encode_properties(p):
    """get_observations (HMM/hmm_training.py:95-118) for drawn frame / codebook shapes: bit-exact against the C oracle
    (every frame, whether the fp32 prefilter decides it or the exact list kernel does), invariant under a permutation
    of the frames, lowest index among duplicated centroids, and encoding a centroid returns the first copy of itself."""
    from hmm_training_b200 import engine
    from oracle import vq_oracle
    F, K, kind, seed = p
    rng = np.random.default_rng(seed)
    C = rng.normal(size=(K, 13)) * 10.0
    if kind == 1 and K > 1:    # duplicated and all-zero centroids
        C[rng.integers(0, K, size=max(1, K // 4))] = C[0]
        C[K // 2] = 0.0
    if kind == 2:              # tightly clustered codebook: everything is a near-tie for the prefilter
        C = C[:1] + rng.normal(size=(K, 13)) * 1e-7
    X = C[rng.integers(0, K, size=F)] + rng.normal(size=(F, 13)) * (1e-9 if kind == 3 else 3.0)
    X[:, 0] = rng.normal(size=F) * 1e3  # dimension 0 never takes part
    idx = engine.vq_encode(X, C)
    assert np.array_equal(idx, vq_oracle.encode(X, C))
    perm = rng.permutation(F)
    assert np.array_equal(engine.vq_encode(X[perm], C), idx[perm])
    own = engine.vq_encode(C, C)
    first = np.array([np.flatnonzero((C[:, 1:] == c[1:]).all(axis=1))[0] for c in C])
    assert np.array_equal(own, first)


score_shape = st.tuples(st.sampled_from([2, 4, 4, 5, 8, 16]), st.sampled_from([4, 16, 256, 300]), st.integers(1, 4),
                        st.integers(1, 40), st.booleans(), st.sampled_from([0.0, 0.3]), st.integers(0, 2 ** 31 - 1))


@pytest.mark.gpu
@_settings(20)
@given(score_shape)
def test_gpu_scoring_properties(p):
    """calculate_log_likelihood / test_hmm (HMM/hmm_testing.py:49-104, 139-161): the [U, W] matrix equals the oracle's,
    is invariant under a permutation of the utterances, the argmax is the first maximum (-1 = "unknown" when no model
    can emit the utterance), and the scorer agrees with the trainer's own first-iteration statistic."""
    from hmm_training_b200 import engine
    N, M, W, U, ltr, zf, seed = p
    rng = np.random.default_rng(seed)
    pi, A, B = _random_model(rng, W, N, M, ltr, zf)
    if zf > 0:
        B[rng.random(B.shape) < 0.2] = 0.0  # structural zeros in the emissions: some utterances become impossible
        B[:, :, 0] += 1e-3
        B /= B.sum(axis=2, keepdims=True)
    seqs, _ = _random_corpus(rng, 1, U, 1, 80, M)
    obs, off = _pack(seqs)
    ll, best = engine.score(obs, off, N, M, pi, A, B)
    ref = O.score_batch(seqs, [(A[w], B[w], pi[w]) for w in range(W)])
    assert_close(ll, ref, "log-likelihood matrix", atol=1e-12)
    assert np.array_equal(best, O.argmax_first(ref))
    perm = rng.permutation(U)
    obs2, off2 = _pack([seqs[u] for u in perm])
    ll2, best2 = engine.score(obs2, off2, N, M, pi, A, B)
    assert np.array_equal(ll2, ll[perm]) and np.array_equal(best2, best[perm])
    # the trainer's statistic of its first iteration is log_sum_exp over the same per-utterance values
    w = int(rng.integers(0, W))
    _, _, _, hist, _ = engine.bw_fit(obs, off, np.zeros(U, dtype=np.int32), 1, N, M, pi[w:w + 1], A[w:w + 1], B[w:w + 1],
                                     epsilon=-1.0, max_iterations=1)
    assert_close(hist[0, 0], O.log_sum_exp(ref[:, w]), "E-step statistic vs scorer", atol=1e-12)
