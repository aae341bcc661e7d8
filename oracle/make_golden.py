"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # all cases (a few minutes: the reference is
    python -m oracle.make_golden bw_small   # pure-Python loops), or selected ones

The reference has no tests, fixtures or golden vectors of its own, so
these files are the parity pin: outputs of the reference's own functions
(hmm_training, calculate_log_likelihood, get_observations, createCodeVector) on seeded
synthetic inputs.  Inputs are stored alongside the outputs so the GPU box needs neither
the reference nor this script.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hmm_training_b200 import synthetic as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _train_ref(ref, obs_list, N, M, iters, eps=1e-6, init=None):
    """Run the reference trainer; with ``init`` use its own warm-start route
    (HMM/hmm_training.py:275-297) from a temp Data/ tree."""
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "HMM"))
        kwargs = dict(load_initial_params=False)
        if init is not None:
            pi0, A0, B0 = init
            d = os.path.join(tmp, "Data", "Eighty-five-percent_20")
            os.makedirs(d)
            with open(os.path.join(d, "w.json"), "w") as f:
                json.dump({"states": N, "symbols": M, "A": np.asarray(A0).tolist(),
                           "B": np.asarray(B0).tolist(), "Pi": np.asarray(pi0).tolist(), "word": "w"}, f)
            kwargs = dict(load_initial_params=True, word_name="w")
        os.chdir(os.path.join(tmp, "HMM"))
        try:
            with ref_shim.quiet(), ref_shim.LLCapture(ref.training) as cap:
                A, B, pi = ref.training.hmm_training(obs_list, N=N, M=M, epsilon=eps, max_iterations=iters,
                                                     show_progress=False, **kwargs)
        finally:
            os.chdir(cwd)
    return A, B, pi, np.array(cap.values)


def _bw_case(ref, name, corpus, N, M, iters, eps=1e-6, init=None):
    W = len(corpus)
    A = np.zeros((W, N, N)); B = np.zeros((W, N, M)); pi = np.zeros((W, N))
    ll = np.full((W, iters), np.nan); its = np.zeros(W, np.int32)
    for w, word in enumerate(corpus):
        a, b, p, hist = _train_ref(ref, word, N, M, iters, eps, None if init is None else
                                   (init[0][w], init[1][w], init[2][w]))
        A[w], B[w], pi[w] = a, b, p
        ll[w, :len(hist)] = hist
        its[w] = len(hist)
        print(f"  {name}: word {w} done, {len(hist)} iterations", flush=True)
    obs, offsets, word_of_seq = S.pack_corpus(corpus, M)
    extra = {}
    if init is not None:
        extra = dict(pi0=np.asarray(init[0]), A0=np.asarray(init[1]), B0=np.asarray(init[2]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), obs=obs, offsets=offsets, word_of_seq=word_of_seq,
                        N=N, M=M, max_iterations=iters, epsilon=eps, A=A, B=B, pi=pi, ll_hist=ll,
                        iters=its, **extra)
    return A, B, pi


def case_bw_c1(ref):
    """BASELINE config 1: 10 words x 20 utterances x U{90..110} frames, N=4, M=256."""
    corpus = S.word_corpus(0, 10, 20, kind="clustered")
    A, B, pi = _bw_case(ref, "bw_c1_clustered_s0_it10", corpus, 4, 256, 10)
    # recognition golden on held-out utterances, scored against those 10 models
    test = S.word_corpus(100, 10, 4, kind="clustered")
    HMM = ref.classes.HMMTrained
    models = [HMM(4, 256, A[w], B[w], pi[w], f"w{w}") for w in range(10)]
    seqs = [s for word in test for s in word]
    ll = np.array([[ref.testing.calculate_log_likelihood(s, m) for m in models] for s in seqs])
    obs, offsets, true_word = S.pack_corpus(test, 256)
    np.savez_compressed(os.path.join(OUT, "score_c1.npz"), obs=obs, offsets=offsets, true_word=true_word,
                        A=A, B=B, pi=pi, ll=ll)


def case_bw_small(ref):
    _bw_case(ref, "bw_uniform_s1_it3", S.word_corpus(1, 3, 20, kind="uniform"), 4, 256, 3)
    _bw_case(ref, "bw_clustered_s2_it1", S.word_corpus(2, 3, 20, kind="clustered"), 4, 256, 1)
    # converges before max_iterations (iteration-count parity)
    _bw_case(ref, "bw_converge_eps", S.word_corpus(3, 4, 6, M=16, tmin=15, tmax=25), 4, 16, 60, eps=1e-3)


def case_bw_empty(ref):
    """A word without a single sequence next to two ordinary ones: the reference does not refuse it — every log sum is
    -inf, A and B come back as zeros, pi as NaN (0 / 0 at hmm_training.py:529), the statistic is -inf in every iteration
    and |(-inf) - (-inf)| = NaN never ends the loop early."""
    corpus = S.word_corpus(7, 3, 5, M=16, tmin=5, tmax=15)
    corpus[1] = []
    _bw_case(ref, "bw_word_without_sequences", corpus, 4, 16, 3)


def case_bw_warm(ref):
    """N != 4 through the reference's own warm-start route, and structural zeros."""
    rng = np.random.default_rng(5)
    N, M, W = 6, 32, 2
    corpus = S.word_corpus(4, W, 8, N=N, M=M, tmin=25, tmax=35)
    pi0 = np.zeros((W, N)); A0 = np.zeros((W, N, N)); B0 = np.zeros((W, N, M))
    for w in range(W):
        pi0[w] = rng.dirichlet(np.ones(N))
        A0[w] = rng.dirichlet(np.ones(N), size=N)
        B0[w] = rng.dirichlet(np.ones(M), size=N)
    _bw_case(ref, "bw_warm_n6_m32", corpus, N, M, 5, init=(pi0, A0, B0))

    # structural zeros: pi = e0, strict left-to-right A, exact zeros in B so that some
    # sequences are impossible (-inf) and some states/symbols have no finite term;
    # includes T=1 and T=2 sequences.
    N, M, W = 4, 12, 2
    corpus = []
    for w in range(W):
        seqs = S.clustered_sequences(rng, 6, N=N, M=M, tmin=8, tmax=14, spread=3)
        seqs.append(np.array([0], dtype=np.int64))
        seqs.append(np.array([1, 4], dtype=np.int64))
        seqs.append(np.array([11, 0, 3], dtype=np.int64))  # starts with a symbol state 0 cannot emit
        corpus.append(seqs)
    pi0 = np.tile(np.array([1.0, 0, 0, 0]), (W, 1))
    A0 = np.tile(np.array([[0.5, 0.5, 0, 0], [0, 0.5, 0.5, 0], [0, 0, 0.5, 0.5], [0, 0, 0, 1.0]]), (W, 1, 1))
    B0 = np.zeros((W, N, M))
    for w in range(W):
        b = rng.random((N, M)) + 0.05
        b[0, 9:] = 0.0   # state 0 cannot emit symbols 9..11
        b[3, :3] = 0.0   # state 3 cannot emit symbols 0..2
        b[2, 5] = 0.0
        B0[w] = b / b.sum(axis=1, keepdims=True)
    _bw_case(ref, "bw_structural_zeros", corpus, N, M, 4, init=(pi0, A0, B0))


def case_bw_ltr(ref):
    """The left-to-right kernel family (N = 8 / 16, bidiagonal A: BASELINE config 4's model family) pinned to the reference
    itself, through its warm-start route: pi entered in state 0 only (N = 16) or everywhere (N = 8), ragged lengths
    including T = 1, more than 32 sequences per word (a full block and a partly filled one), 5 iterations."""
    rng = np.random.default_rng(16)
    for N, M, W, S_, tmin, tmax, iters, e0 in ((16, 64, 2, 40, 18, 45, 5, True), (8, 24, 2, 37, 6, 22, 5, False)):
        corpus = []
        for w in range(W):
            seqs = S.clustered_sequences(rng, S_, N=N, M=M, tmin=tmin, tmax=tmax, shift=3 * w, spread=max(2, M // (2 * N)))
            seqs[0] = seqs[0][:1]
            corpus.append(seqs)
        pi0 = np.zeros((W, N)); A0 = np.zeros((W, N, N)); B0 = np.zeros((W, N, M))
        for w in range(W):
            if e0:
                pi0[w, 0] = 1.0
            else:
                pi0[w] = rng.dirichlet(np.ones(N))
            for i in range(N):
                stay = 0.4 + 0.4 * rng.random()
                A0[w, i, i] = stay if i + 1 < N else 1.0
                if i + 1 < N:
                    A0[w, i, i + 1] = 1.0 - stay
            B0[w] = rng.dirichlet(np.ones(M) * 2.0, size=N)
        A, B, pi = _bw_case(ref, f"bw_ltr_n{N}_m{M}", corpus, N, M, iters, init=(pi0, A0, B0))
        # recognition golden for the same family: held-out utterances (one of a single frame, one the models of word 1
        # find far less likely than those of word 0) against the trained models and against the initial ones
        HMM = ref.classes.HMMTrained
        models = [HMM(N, M, A[w], B[w], pi[w], f"w{w}") for w in range(W)] + \
                 [HMM(N, M, A0[w], B0[w], pi0[w], f"init{w}") for w in range(W)]
        test = []
        for w in range(W):
            t = S.clustered_sequences(rng, 6, N=N, M=M, tmin=tmin, tmax=3 * tmax, shift=3 * w, spread=max(2, M // (2 * N)))
            t[0] = t[0][:1]
            test.append(t)
        seqs = [s_ for word in test for s_ in word]
        ll = np.array([[ref.testing.calculate_log_likelihood(s_, m) for m in models] for s_ in seqs])
        obs, offsets, true_word = S.pack_corpus(test, M)
        np.savez_compressed(os.path.join(OUT, f"score_ltr_n{N}_m{M}.npz"), obs=obs, offsets=offsets, true_word=true_word,
                            N=N, M=M, A=np.stack([m.A for m in models]), B=np.stack([m.B for m in models]),
                            pi=np.stack([m.Pi for m in models]), ll=ll)


def case_vq(ref):
    Raw, Cen = ref.cvc.RawDataMFCC, ref.cvc.CentroidDataMFCC
    X = S.mfcc_mixture(0, 2000, K=64)
    C = S.random_codebook(1, 256)
    C[17] = C[5]          # exact duplicate centroid: lowest index must win
    C[200] = 0.0; C[201] = 0.0  # twin all-zero centroids (empty-cluster artefact)
    X[10] = C[5]; X[11] = C[200]; X[12, 1:] = C[33, 1:]  # exact hits, energy dim ignored
    frames = [Raw(raw_samples=np.array([]), mfcc=x.copy()) for x in X]
    cents = [Cen(mfcc=c.copy(), id=i) for i, c in enumerate(C)]
    recs = [frames[:700], frames[700:701], frames[701:]]
    obs = ref.training.get_observations(recs, cents)
    np.savez_compressed(os.path.join(OUT, "vq_2000x256.npz"), X=X, C=C, idx=np.concatenate(obs).astype(np.int32),
                        rec_lens=np.array([len(r) for r in recs]))


def _lbg_case(ref, name, X, K, max_iter, eps=0.001):
    Raw = ref.cvc.RawDataMFCC
    frames = [Raw(raw_samples=np.array([]), mfcc=x.copy()) for x in X]
    with ref_shim.quiet() as buf:
        cents, gens = ref.cvf.createCodeVector(frames, centroids_quantity=K, max_iterations=max_iter,
                                               epsilon=eps, save_updates=False)
    iters = [int(line.split("Converged after")[1].split()[0]) for line in buf.getvalue().splitlines()
             if "Converged after" in line]
    gens_flat = np.concatenate([np.array([c.mfcc for c in g]) for g in gens])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, K=K, max_iterations=max_iter, epsilon=eps,
                        C=np.array([c.mfcc for c in cents]), ids=np.array([c.id for c in cents]),
                        gens=gens_flat, gen_sizes=np.array([len(g) for g in gens]),
                        assign=np.array([f.parent_centroid_id for f in frames], dtype=np.int32),
                        generation=np.array([f.generation for f in frames], dtype=np.int32),
                        iters=np.array(iters, dtype=np.int32),
                        stdout=np.array(buf.getvalue()))  # the reference's own prints (progress lines, :472, :512-516)
    print(f"  {name}: iters {iters}", flush=True)


def case_lbg(ref):
    _lbg_case(ref, "lbg_600_k32", S.mfcc_mixture(3, 600, K=16), 32, 100)
    _lbg_case(ref, "lbg_1200_k256_it3", S.mfcc_mixture(4, 1200, K=64), 256, 3)
    # few distinct points -> empty clusters -> all-zero twin centroids (:435)
    rng = np.random.default_rng(7)
    base = S.mfcc_mixture(5, 3, K=3)
    _lbg_case(ref, "lbg_empty_clusters", base[rng.integers(0, 3, size=60)], 16, 20)
    _lbg_case(ref, "lbg_k1", S.mfcc_mixture(6, 40, K=2), 1, 5)
    _lbg_case(ref, "lbg_k24_nonpow2", S.mfcc_mixture(8, 200, K=8), 24, 10)


def case_frames(ref):
    """Frame file written by the reference's own DataStorage.save_raw_data (codevector_classes.py:
    438-444) — the input of the native loader (hmmb_frames_json_scan).  raw_samples stay empty so
    that __post_init__ does not call librosa (:217-220); names and values are chosen to be awkward."""
    Raw = ref.cvc.RawDataMFCC
    rng = np.random.default_rng(21)
    X = S.mfcc_mixture(9, 40, K=4)
    X[3] = [0.0, -0.0, 5e-324, 1e-320, 2.2250738585072014e-308, 1.7976931348623157e308, -1e300, 1.0, 3.0, 1e22,
            123456789012345680.0, 0.1, 1 / 3]
    X[7] = np.round(X[7])  # integral floats print as "12.0"
    names = ['finish-04', 'say "mfcc_vector": [1, 2]', 'back\\slash\\', 'quote\\"end', "üñí", "mfcc_vector"]
    frames = [Raw(raw_samples=np.array([]), mfcc=X[i].copy(), frame_number=i, recording=names[i % len(names)],
                  parent_centroid_id=int(rng.integers(0, 256)), generation=int(rng.integers(0, 9)))
              for i in range(len(X))]
    path = os.path.join(OUT, "frames_ref.json")
    ref.cvc.DataStorage.save_raw_data(frames, path)
    np.savez_compressed(os.path.join(OUT, "frames_ref.npz"), X=X)


def case_pipeline(ref):
    """The callers either side of the hot path, run by the reference itself on a small Data/ tree:
    training_with_save per word (HMM/main.py:147-154) then test_hmm (HMM/hmm_testing.py:107-163)."""
    Raw, Cen = ref.cvc.RawDataMFCC, ref.cvc.CentroidDataMFCC
    rng = np.random.default_rng(33)
    W, S_tr, S_te, K = 3, 6, 4, 16
    C = S.random_codebook(12, K)
    words = ["alpha", "bravo", "charlie"]

    def recording(w):
        T = int(rng.integers(18, 34))
        seg = np.sort(rng.integers(0, 4, size=T))
        cent = C[(4 * w + seg + rng.integers(0, 2, size=T)) % K]
        return cent + rng.normal(size=(T, 13)) * np.concatenate([[30.0], np.linspace(4.0, 0.5, 12)])

    train = [[recording(w) for _ in range(S_tr)] for w in range(W)]
    test = [[recording(w) for _ in range(S_te)] for w in range(W)]
    centroids = [Cen(mfcc=C[k].copy(), id=k) for k in range(K)]
    to_frames = lambda rec: [Raw(raw_samples=np.array([]), mfcc=x.copy()) for x in rec]
    cwd = os.getcwd()
    models = []
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "HMM"))
        os.makedirs(os.path.join(tmp, "Data", "CodeVector"))
        os.chdir(os.path.join(tmp, "HMM"))
        try:
            with ref_shim.quiet():
                ref.cvc.DataStorage.save_centroids(centroids, os.path.join(tmp, "Data", "CodeVector", "codevector.json"))
                for w in range(W):
                    models.append(ref.training.training_with_save([to_frames(r) for r in train[w]], centroids, words[w],
                                                                  max_iterations=3, show_progress=False))
                true, pred = ref.testing.test_hmm(models, {words[w]: [to_frames(r) for r in test[w]] for w in range(W)},
                                                  base_dir=os.path.join(tmp, "Data"))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "pipeline_ref.npz"), C=C, words=np.array(words),
                        train=np.concatenate([r for word in train for r in word]),
                        train_len=np.array([[len(r) for r in word] for word in train]),
                        test=np.concatenate([r for word in test for r in word]),
                        test_len=np.array([[len(r) for r in word] for word in test]),
                        A=np.stack([m.A for m in models]), B=np.stack([m.B for m in models]),
                        pi=np.stack([m.Pi for m in models]), true=np.array(true), pred=np.array(pred))


CASES = {"bw_c1": case_bw_c1, "bw_small": case_bw_small, "bw_warm": case_bw_warm, "bw_ltr": case_bw_ltr, "bw_empty": case_bw_empty,
         "vq": case_vq, "lbg": case_lbg, "frames": case_frames, "pipeline": case_pipeline}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    for name in (sys.argv[1:] or list(CASES)):
        print(f"[golden] {name}", flush=True)
        CASES[name](ref)
