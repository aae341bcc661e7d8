"""Parity at BASELINE.json's full sizes (VERDICT r1 "Missing 4"): the CUDA path against the oracle on the same
inputs where the oracle finishes in seconds — VQ on 1 M frames, LBG on 200 k frames x K = 256, one EM iteration
of a whole config-3 word (100 000 sequences x 200 frames), 2 000 sampled utterances of config 5's million — plus
the near-tie report of the VQ contract and the reference's createCodeVector prints."""
import os

import numpy as np
import pytest

from hmm_training_b200 import codevector_functions as CF
from hmm_training_b200 import engine, synthetic
from hmm_training_b200.codevector_classes import RawDataMFCC
from oracle import hmm_oracle as O
from oracle import vq_oracle

from helpers import assert_close, assert_same_support, load_golden

pytestmark = pytest.mark.gpu


def test_vq_encode_one_million_frames_bit_exact():
    """BASELINE config 2's encode in full: 1 M frames x 256 centroids, indices equal to the C oracle's."""
    F, K = 1_000_000, 256
    X = synthetic.mfcc_mixture(0, F, K)
    C = synthetic.random_codebook(1, K)
    idx = engine.vq_encode(X, C)
    ref = vq_oracle.encode(X, C)
    assert np.array_equal(idx, ref), f"{int((idx != ref).sum())} of {F} indices differ"
    # the same frames against a codebook trained on them (close centroids, the regime after LBG)
    Ct = engine.lbg_fit(X[:50_000], 256, 3, 1e-3)[0]
    assert np.array_equal(engine.vq_encode(X, Ct), vq_oracle.encode(X, Ct))


def test_vq_prefilter_adversarial_codebooks():
    """Codebooks built to defeat the fp32 prefilter: exact twins, twins one ulp apart, centroids on a line through
    the frame at distances that differ in the 10th digit, huge and tiny magnitudes, non-finite rows.  Every index
    must still be the exact scan's."""
    rng = np.random.default_rng(11)
    X = synthetic.mfcc_mixture(5, 4096, K=32)
    C = synthetic.random_codebook(6, 64)
    C[10] = C[3]                                   # exact twin: lowest index wins
    C[11] = np.nextafter(C[4], np.inf)             # one ulp away in every coordinate
    C[12] = X[7]; C[13] = X[7]                     # zero distance twice
    d = rng.normal(size=13); d /= np.linalg.norm(d[1:])
    C[20] = X[100] + 3.0 * d; C[21] = X[100] - 3.0 * (1 + 1e-10) * d   # |x - c| equal to 1e-10 relative
    C[22] = X[200] + 5.0 * d; C[23] = X[200] - 5.0 * (1 + 1e-14) * d
    for scale in (1.0, 1e-30, 1e25, 1e160):
        idx = engine.vq_encode(X * scale, C * scale)
        assert np.array_equal(idx, vq_oracle.encode(X * scale, C * scale)), f"scale {scale}"
    Xn = X.copy(); Xn[5, 3] = np.nan; Xn[6, 2] = np.inf
    Cn = C.copy(); Cn[30, 4] = np.nan
    assert np.array_equal(engine.vq_encode(Xn, Cn), vq_oracle.encode(Xn, Cn))


def test_vq_near_tie_report():
    """North-star contract: indices bit-exact, near-ties ((d2 - d1) / d1 < 1e-12) LISTED."""
    X = synthetic.mfcc_mixture(2, 3000, K=16)
    C = synthetic.random_codebook(3, 32)
    idx, near, n_near = engine.vq_encode(X, C, near_ties=True)
    ref, _ = vq_oracle.encode(X, C, return_dist=True)
    assert np.array_equal(idx, ref) and n_near == 0 and len(near) == 0   # generic data: no near-ties
    C2 = C.copy()
    C2[9] = C2[2]                                    # every frame nearest to 2 is an exact tie with 9
    d = np.zeros(13); d[1] = 1.0
    C2[20] = X[17] + 2.0 * d; C2[21] = X[17] - 2.0 * d * (1 + 1e-13)   # frame 17: two centroids at 2 (1 +- 1e-13)
    idx, near, n_near = engine.vq_encode(X, C2, near_ties=True)
    ref = vq_oracle.encode(X, C2)
    assert np.array_equal(idx, ref)
    # what the report must contain, from fp64 distances on the host
    D = np.sqrt(((X[:, None, 1:] - C2[None, :, 1:]) ** 2).sum(-1))
    D.sort(axis=1)
    want = np.flatnonzero((D[:, 1] - D[:, 0]) <= 1e-12 * D[:, 0])
    assert 17 in want and n_near == len(near)
    # (frames within a few ulps of the 1e-12 boundary may fall on either side of it)
    loose = np.flatnonzero((D[:, 1] - D[:, 0]) <= 1.001e-12 * D[:, 0])
    tight = np.flatnonzero((D[:, 1] - D[:, 0]) <= 0.999e-12 * D[:, 0])
    assert set(tight) <= set(near.tolist()) <= set(loose)
    # capacity smaller than the list: count still complete
    _, near3, n3 = engine.vq_encode(X, C2, near_ties=True, near_cap=3)
    assert n3 == n_near and len(near3) == min(3, n_near)


def test_lbg_200k_frames_k256_matches_oracle():
    """The Lloyd passes the bench times (BASELINE config 2: K = 256) against vq_oracle.lbg on 200 000 frames:
    iteration counts and assignments equal, centroids of every generation to 1e-9
    (CodeVector/codevector_functions.py:485-510)."""
    F, K, it = 200_000, 256, 12
    X = synthetic.mfcc_mixture(0, F, K)
    C, gens, assign, iters, gd = engine.lbg_fit(X, K, it, 1e-3)
    Co, genso, assigno, iterso, gdo = vq_oracle.lbg(X, K, it, 1e-3)
    assert np.array_equal(iters, iterso)
    assert_close(gd, gdo, "summed distances", rtol=1e-12, atol=0)
    assert_close(C, Co, "centroids", rtol=1e-9, atol=1e-12)
    assert_close(np.concatenate(gens), np.concatenate(genso), "generations", rtol=1e-9, atol=1e-12)
    assert np.array_equal(assign, assigno), f"{int((assign != assigno).sum())} assignments differ"


@pytest.mark.parametrize("name", ["lbg_600_k32", "lbg_1200_k256_it3", "lbg_k1", "lbg_k24_nonpow2"])
def test_create_code_vector_prints_match_reference(name, capsys):
    """The reference's own stdout (progress line every tenth pass, final diff of every generation,
    codevector_functions.py:447-523), captured by oracle/make_golden.py, against the drop-in's."""
    g = load_golden(name)
    frames = [RawDataMFCC(raw_samples=np.array([]), mfcc=x.copy()) for x in g["X"]]
    capsys.readouterr()
    CF.createCodeVector(frames, centroids_quantity=int(g["K"]), max_iterations=int(g["max_iterations"]),
                        epsilon=float(g["epsilon"]), save_updates=False)
    got = capsys.readouterr().out
    assert got == str(g["stdout"])


def test_config3_whole_word_em_iteration_matches_oracle():
    """BASELINE config 3 in full on the GPU (10 words x 100 000 sequences x T = 200); the third EM iteration of one
    whole word — pi, A, B re-estimated from 20 M frames — against the oracle's log-space iteration over the same
    100 000 sequences (chunked, oracle.em_iteration_chunked), at the contract's 1e-9."""
    N, M, W, S, T = 4, 256, 10, 100_000, 200
    obs, offsets, wos = synthetic.fixed_length_codewords(1000, W, S, T, N, M)
    pi0, A0, B0 = engine.default_init(N, M)
    with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:
        bw.set_params(np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)))
        bw.iterate(2, -1.0, 100)
        pi_raw, A_raw, B_raw = bw.params(finalize=False)
        bw.iterate(1, -1.0, 100)
        pi, A, B = bw.params(finalize=True)
        hist, iters = bw.history(100)
        assert bw.diagnostics() == (0, 0)
    w = 3
    rows = obs.reshape(W * S, T)[w * S:(w + 1) * S]
    seqs = [r.astype(np.int64) for r in rows]
    (Ao, Bo, pio), ll = O.em_iteration_chunked(seqs, (pi_raw[w], A_raw[w], B_raw[w]), M, chunk=2500,
                                               procs=min(16, os.cpu_count() or 1))
    assert_close(hist[w, 2:3], np.array([ll]), "statistic of iteration 3")
    assert_close(A[w], Ao, "A"); assert_close(B[w], Bo, "B"); assert_close(pi[w], pio, "pi")
    assert_same_support(A[w], Ao); assert_same_support(pi[w], pio)


def test_config5_sampled_utterances_match_oracle():
    """BASELINE config 5 in full (1 M utterances x 10 trained-like models): 2 000 utterances sampled from the [U, W]
    matrix against the oracle's scorer, and the argmax (first maximum wins, hmm_testing.py:143-153) on the sample."""
    U, Wm, T = 1_000_000, 10, 100
    rng = np.random.default_rng(5)
    obs, offsets, _ = synthetic.fixed_length_codewords(77, Wm, U // Wm, T, 4, 256)
    pi, A, _ = engine.default_init(4, 256)
    Bm = rng.dirichlet(np.ones(256) * 0.3, size=(Wm, 4))
    Bm[3, 2, :40] = 1e-20   # floored entries (hmm_training.py:497) in one model
    pim, Am = np.tile(pi, (Wm, 1)), np.tile(A, (Wm, 1, 1))
    ll, arg = engine.score(obs, offsets, 4, 256, pim, Am, Bm)
    sample = np.unique(np.concatenate([np.arange(40), np.arange(U - 40, U), rng.choice(U, 1920, replace=False)]))
    seqs = [obs[offsets[u]:offsets[u + 1]].astype(np.int64) for u in sample]
    ref = O.score_batch(seqs, [(Am[w], Bm[w], pim[w]) for w in range(Wm)])
    assert_close(ll[sample], ref, "sampled [U, W] log-likelihoods")
    assert np.array_equal(arg[sample], O.argmax_first(ref))
    assert np.array_equal(arg, np.argmax(ll, axis=1))   # all finite here: first maximum
