"""The CPU oracle (oracle/) pinned against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU-only; runs everywhere."""
import numpy as np
import pytest

from helpers import assert_close, assert_same_support, floored_set, load_golden, split_corpus
from oracle import hmm_oracle as O
from oracle import vq_oracle

BW_CASES = ["bw_c1_clustered_s0_it10", "bw_uniform_s1_it3", "bw_clustered_s2_it1", "bw_converge_eps",
            "bw_warm_n6_m32", "bw_structural_zeros", "bw_ltr_n16_m64", "bw_ltr_n8_m24", "bw_word_without_sequences"]


@pytest.mark.parametrize("name", BW_CASES)
def test_oracle_baum_welch_matches_reference(name):
    g = load_golden(name)
    N, M = int(g["N"]), int(g["M"])
    corpus = split_corpus(g)
    for w, word in enumerate(corpus):
        init = (g["pi0"][w], g["A0"][w], g["B0"][w]) if "pi0" in g else None
        A, B, pi, hist, it = O.hmm_training(word, N=N, M=M, epsilon=float(g["epsilon"]),
                                            max_iterations=int(g["max_iterations"]), init=init,
                                            return_history=True)
        assert it == int(g["iters"][w]), f"{name} word {w}: iteration count"
        assert_close(hist, g["ll_hist"][w, :it], f"{name} w{w} ll")
        assert_close(A, g["A"][w], f"{name} w{w} A")
        assert_close(pi, g["pi"][w], f"{name} w{w} pi")
        assert_close(B, g["B"][w], f"{name} w{w} B")
        assert_same_support(A, g["A"][w], "A")
        assert_same_support(pi, g["pi"][w], "pi")
        assert np.array_equal(floored_set(B, M), floored_set(g["B"][w], M))


def test_oracle_scoring_matches_reference():
    g = load_golden("score_c1")
    off = g["offsets"]
    seqs = [g["obs"][off[r]:off[r + 1]].astype(np.int64) for r in range(len(off) - 1)]
    models = [(g["A"][w], g["B"][w], g["pi"][w]) for w in range(g["A"].shape[0])]
    ll = O.score_batch(seqs, models)
    assert_close(ll, g["ll"], "score ll")
    one = O.calculate_log_likelihood(seqs[3], *models[2])
    assert_close(one, g["ll"][3, 2], "single ll")
    assert np.array_equal(O.argmax_first(ll), O.argmax_first(g["ll"]))


@pytest.mark.parametrize("name", ["score_ltr_n16_m64", "score_ltr_n8_m24"])
def test_oracle_scoring_matches_reference_left_to_right(name):
    """Bidiagonal models with 16 / 8 states (trained and initial), utterances up to 3x the training length, one of a
    single frame; the 16-state initial models enter in state 0 only, so some utterances are impossible (-inf)."""
    g = load_golden(name)
    off = g["offsets"]
    seqs = [g["obs"][off[r]:off[r + 1]].astype(np.int64) for r in range(len(off) - 1)]
    models = [(g["A"][w], g["B"][w], g["pi"][w]) for w in range(g["A"].shape[0])]
    ll = O.score_batch(seqs, models)
    assert np.array_equal(np.isneginf(ll), np.isneginf(g["ll"]))
    assert_close(ll, g["ll"], "score ll")
    assert np.array_equal(O.argmax_first(ll), O.argmax_first(g["ll"]))


def test_oracle_vq_matches_reference_bit_exact():
    g = load_golden("vq_2000x256")
    idx = vq_oracle.encode(g["X"], g["C"])
    assert np.array_equal(idx, g["idx"])
    assert idx[10] == 5 and idx[11] == 200 and idx[12] == 33  # duplicates: lowest index wins


@pytest.mark.parametrize("name", ["lbg_600_k32", "lbg_1200_k256_it3", "lbg_empty_clusters", "lbg_k1",
                                  "lbg_k24_nonpow2"])
def test_oracle_lbg_matches_reference(name):
    g = load_golden(name)
    C, gens, assign, iters, _ = vq_oracle.lbg(g["X"], int(g["K"]), int(g["max_iterations"]), float(g["epsilon"]))
    assert C.shape == g["C"].shape
    assert np.array_equal(iters, g["iters"])
    assert_close(C, g["C"], "centroids", rtol=1e-12, atol=0)
    assert_close(np.concatenate(gens), g["gens"], "generations", rtol=1e-12, atol=0)
    assert [len(x) for x in gens] == list(g["gen_sizes"])
    if len(g["iters"]):
        assert np.array_equal(assign, g["assign"])


def test_oracle_errors():
    with pytest.raises(IndexError):
        O.hmm_training([np.array([], dtype=np.int64)], max_iterations=1)
    with pytest.raises(ValueError):
        vq_oracle.lbg(np.zeros((0, 13)), 4)


def test_chunked_em_iteration_equals_pinned_trainer():
    """oracle.em_iteration_chunked (used by the full-size GPU parity tests) is the loop body of the pinned
    hmm_training, reduced chunk by chunk with log_sum_exp: one and two iterations must agree to rounding,
    including the 1e-20 floor pattern and structural zeros."""
    import numpy as np
    from hmm_training_b200 import synthetic
    from oracle import hmm_oracle as O
    corpus = synthetic.word_corpus(3, 2, 24)
    for w, seqs in enumerate(corpus):
        init = O.default_init(4, 256)
        A1, B1, p1, h, it = O.hmm_training(seqs, max_iterations=1, return_history=True)
        (A2, B2, p2), ll = O.em_iteration_chunked(seqs, init, 256, chunk=5, procs=2 if w else 1)
        assert abs(ll - h[0]) <= 1e-12 * abs(h[0])
        for x, y in ((A1, A2), (B1, B2), (p1, p2)):
            assert np.array_equal(x == 0, y == 0)
            assert np.all(np.abs(x - y) <= 1e-12 * np.abs(x) + 1e-300)
        # a second iteration chained from the un-normalised log parameters
        lp, lA, lB = O.mstep_from_logsums(O.merge_logsums([O.estep_logsums(seqs, O.safe_log(init[0]), O.safe_log(init[1]),
                                                                           O.safe_log(init[2]), 256)]), 4, 256)
        (A3, B3, p3), _ = O.em_iteration_chunked(seqs, (O.safe_exp(lp), O.safe_exp(lA), O.safe_exp(lB)), 256, chunk=7)
        A4, B4, p4 = O.hmm_training(seqs, max_iterations=2, epsilon=-1.0)
        for x, y in ((A4, A3), (B4, B3), (p4, p3)):
            assert np.all(np.abs(x - y) <= 1e-11 * np.abs(x) + 1e-300)
