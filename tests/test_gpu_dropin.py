"""The reference-shaped Python surface (same names / arguments / side effects as the
reference's modules) on top of the CUDA path.  GPU required."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from helpers import assert_close, load_golden, split_corpus
from hmm_training_b200 import codevector_functions as cvf
from hmm_training_b200 import hmm_testing, hmm_training, synthetic
from hmm_training_b200.codevector_classes import CentroidDataMFCC, DataStorage, RawDataMFCC
from hmm_training_b200.hmm_classes import DataStorageHMM, HMMTrained
from oracle import hmm_oracle as O
from oracle import vq_oracle

pytestmark = pytest.mark.gpu


def _frames(X, name="rec"):
    return [RawDataMFCC(raw_samples=np.array([]), mfcc=x.copy(), frame_number=i, recording=name) for i, x in enumerate(X)]


def _centroids(C):
    return [CentroidDataMFCC(mfcc=c.copy(), id=i) for i, c in enumerate(C)]


def test_get_observations_matches_reference_golden():
    g = load_golden("vq_2000x256")
    frames = _frames(g["X"])
    lens = g["rec_lens"]
    recs, pos = [], 0
    for n in lens:
        recs.append(frames[pos:pos + n]); pos += n
    recs.insert(1, [])  # an empty recording yields an empty array, like the reference
    obs = hmm_training.get_observations(recs, _centroids(g["C"]))
    assert len(obs) == len(recs) and len(obs[1]) == 0
    assert obs[0].dtype == np.int64
    assert np.array_equal(np.concatenate([obs[0], obs[2], obs[3]]), g["idx"])
    none = hmm_training.get_observations(recs[:2], [])  # no centroids: every frame keeps index 0 (:103)
    assert len(none) == 2 and len(none[1]) == 0 and none[0].dtype == np.int64 and len(none[0]) == lens[0] and not none[0].any()


def test_hmm_training_signature_and_prints(tmp_path, monkeypatch):
    g = load_golden("bw_converge_eps")
    corpus = split_corpus(g)
    monkeypatch.chdir(tmp_path)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        A, B, pi = hmm_training.hmm_training(corpus[0], N=4, M=16, epsilon=float(g["epsilon"]), max_iterations=60,
                                             show_progress=True, word_name="nofile", load_initial_params=True)
    out = buf.getvalue()
    it = int(g["iters"][0])
    assert "Could not load saved model for word 'nofile'" in out and "Using default transition matrix" in out
    assert f"Converged after {it} iterations" in out and out.count("Iteration ") == it
    assert f"Log-likelihood: {g['ll_hist'][0, it - 1]:.6f}" in out
    assert_close(A, g["A"][0], "A"); assert_close(B, g["B"][0], "B"); assert_close(pi, g["pi"][0], "pi")
    with pytest.raises(IndexError):  # the reference's defaults are 4-state literals
        hmm_training.hmm_training(corpus[0], N=5, M=16, max_iterations=1, show_progress=False, load_initial_params=False)
    with pytest.raises(IndexError):  # empty recording: hmm_training.py:376
        hmm_training.hmm_training([np.array([], dtype=np.int64)], N=4, M=16, max_iterations=1, show_progress=False,
                                  load_initial_params=False)
    with pytest.raises(IndexError, match="only integers"):  # float codewords: numpy's message at :353
        hmm_training.hmm_training([np.array([1.0, 2.0, 3.0])], N=4, M=16, max_iterations=1, show_progress=False,
                                  load_initial_params=False)
    with pytest.raises(UnboundLocalError):  # max_iterations = 0: the closing print at :516 reads an unassigned local
        hmm_training.hmm_training(corpus[0], N=4, M=16, max_iterations=0, show_progress=False, load_initial_params=False)


def test_hmm_training_without_recordings_behaves_like_the_reference():
    """An empty list of recordings is not an error in the reference: zeros for A and B, NaN for pi, "-inf" statistics, all
    iterations run (the lines below are its output for this call; tests/golden/bw_word_without_sequences.npz holds the
    same case as one word of three)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        A, B, pi = hmm_training.hmm_training([], N=4, M=16, epsilon=1e-6, max_iterations=2, show_progress=True,
                                             load_initial_params=False)
    assert buf.getvalue().splitlines() == [
        "Using default initial state probabilities", "Using default transition matrix", "Using default emission matrix",
        "Iteration 1", "Log-likelihood: -inf, Diff: inf", "Iteration 2", "Log-likelihood: -inf, Diff: inf",
        "Log-likelihood: -inf, Diff: inf", "Reached maximum iterations (2)"]
    assert A.shape == (4, 4) and B.shape == (4, 16) and not A.any() and not B.any() and np.isnan(pi).all()


def test_warm_start_file_route(tmp_path, monkeypatch):
    """../Data/Eighty-five-percent_20/<word>.json relative to the CWD (hmm_training.py:278)."""
    g = load_golden("bw_warm_n6_m32")
    corpus = split_corpus(g)
    (tmp_path / "HMM").mkdir(); d = tmp_path / "Data" / "Eighty-five-percent_20"; d.mkdir(parents=True)
    DataStorageHMM.save_hmm(HMMTrained(6, 32, g["A0"][1], g["B0"][1], g["pi0"][1], "w1"), str(d), print_messages=False)
    monkeypatch.chdir(tmp_path / "HMM")
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        A, B, pi = hmm_training.hmm_training(corpus[1], N=6, M=32, max_iterations=5, show_progress=True, word_name="w1",
                                             load_initial_params=True)
    assert "Loaded initial parameters from saved model for word 'w1'" in buf.getvalue()
    assert "Reached maximum iterations (5)" in buf.getvalue()
    assert_close(A, g["A"][1], "A"); assert_close(B, g["B"][1], "B"); assert_close(pi, g["pi"][1], "pi")


def test_training_with_save_and_test_hmm_end_to_end(tmp_path, monkeypatch):
    """VQ -> Baum-Welch -> JSON -> recognition through the reference's entry points, checked
    against the same pipeline on the CPU oracle."""
    (tmp_path / "HMM").mkdir(); (tmp_path / "Data" / "CodeVector").mkdir(parents=True)
    monkeypatch.chdir(tmp_path / "HMM")
    rng = np.random.default_rng(3)
    K = 32
    C = synthetic.random_codebook(9, K)
    words = ["alpha", "beta", "gamma"]

    def utterances(w, n):
        out = []
        for _ in range(n):
            T = int(rng.integers(30, 45))
            seg = np.sort(rng.integers(0, 4, size=T))
            cid = (8 * seg + 3 * w + rng.integers(0, 5, size=T)) % K
            X = C[cid] + rng.normal(size=(T, 13)) * 0.5
            out.append(_frames(X, f"{words[w]}_{len(out)}"))
        return out

    train = {words[w]: utterances(w, 12) for w in range(3)}
    test = {words[w]: utterances(w, 4) for w in range(3)}
    cents = _centroids(C)
    with contextlib.redirect_stdout(io.StringIO()):
        DataStorage.save_centroids(cents, str(tmp_path / "Data" / "CodeVector" / "codevector.json"))
        models = [hmm_training.training_with_save(train[w], cents, w, max_iterations=4, show_progress=False) for w in words]
    for m in models:
        saved = json.load(open(tmp_path / "Data" / "ResultsHMM" / f"{m.word}.json"))
        assert saved["states"] == 4 and saved["symbols"] == K and saved["word"] == m.word
        assert np.array_equal(np.array(saved["B"]), m.B)
    # oracle pipeline
    for w, m in zip(words, models):
        obs = [vq_oracle.encode(np.array([f.mfcc for f in rec]), C).astype(np.int64) for rec in train[w]]
        Ao, Bo, pio = O.hmm_training(obs, N=4, M=K, max_iterations=4)
        assert_close(m.A, Ao, "A"); assert_close(m.B, Bo, "B"); assert_close(m.Pi, pio, "pi")
    with contextlib.redirect_stdout(io.StringIO()):
        loaded = DataStorageHMM.load_all_hmms(str(tmp_path / "Data" / "ResultsHMM"), print_messages=False)
        true, pred = hmm_testing.test_hmm(loaded, test, base_dir=str(tmp_path / "Data"))
    assert true == [w for w in words for _ in range(4)]
    seqs = [vq_oracle.encode(np.array([f.mfcc for f in rec]), C).astype(np.int64) for w in words for rec in test[w]]
    ref = O.score_batch(seqs, [(m.A, m.B, m.Pi) for m in loaded])
    want = [loaded[k].word if k >= 0 else "unknown" for k in O.argmax_first(ref)]
    assert pred == want
    assert sum(t == p for t, p in zip(true, pred)) >= 10  # the synthetic words are separable
    one = hmm_testing.calculate_log_likelihood(seqs[0], loaded[0])
    assert_close(one, ref[0, 0], "calculate_log_likelihood")
    # batched trainer == per-word trainer
    with contextlib.redirect_stdout(io.StringIO()):
        batched = hmm_training.train_hmm_batched(train, cents, max_iterations=4, save=False)
    for a, b in zip(batched, models):
        assert_close(a.A, b.A, "batched A", rtol=1e-12); assert_close(a.B, b.B, "batched B", rtol=1e-12)


def test_create_code_vector_side_effects(tmp_path):
    g = load_golden("lbg_600_k32")
    frames = _frames(g["X"])
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        cents, gens = cvf.createCodeVector(frames, centroids_quantity=32, max_iterations=100, epsilon=0.001,
                                           save_updates=True, output_dir=str(tmp_path / "cv"))
    assert "Creating codevector with 32 centroids..." in buf.getvalue()
    assert [c.id for c in cents] == list(range(32)) and [len(x) for x in gens] == [1, 2, 4, 8, 16, 32]
    assert_close(np.array([c.mfcc for c in cents]), g["C"], "centroids", atol=1e-12)
    assert np.array_equal([f.parent_centroid_id for f in frames], g["assign"])  # inputs mutated in place (:502)
    assert all(f.generation == 5 for f in frames)                                # (:479-480)
    summary = json.load(open(tmp_path / "cv" / "training_summary.json"))
    assert summary["total_frames"] == 600 and summary["max_generation"] == 5
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(OverflowError):  # int(np.log2(0)) at :465
            cvf.createCodeVector(frames[:10], centroids_quantity=0, save_updates=False)
        with pytest.raises(ValueError):     # int(np.log2(-1))
            cvf.createCodeVector(frames[:10], centroids_quantity=-1, save_updates=False)
        with pytest.raises(ValueError, match="No raw data provided"):
            cvf.createCodeVector([], centroids_quantity=4, save_updates=False)
    reloaded = DataStorage.load_raw_data_mfcc(str(tmp_path / "cv" / "codevector_frames_updated.json"), print_messages=False)
    assert reloaded[7].parent_centroid_id == frames[7].parent_centroid_id and np.array_equal(reloaded[7].mfcc, frames[7].mfcc)
    with pytest.raises(ValueError):
        cvf.createCodeVector([])


def _write_tree(tmp, g, save_raw):
    """Data/ tree of the reference's layout from the pipeline golden."""
    from hmm_training_b200.codevector_classes import CentroidDataMFCC, DataStorage, RawDataMFCC
    words = [str(w) for w in g["words"]]
    data = os.path.join(tmp, "Data")
    os.makedirs(os.path.join(data, "CodeVector"))
    os.makedirs(os.path.join(tmp, "HMM"))
    DataStorage.save_centroids([CentroidDataMFCC(mfcc=c, id=k) for k, c in enumerate(g["C"])],
                               os.path.join(data, "CodeVector", "codevector.json"))
    for purpose, key, fname in (("TrainHMM", "train", "hmm_frames.json"), ("Test", "test", "test_frames.json")):
        pos = 0
        for w, word in enumerate(words):
            for r, T in enumerate(g[key + "_len"][w]):
                d = os.path.join(data, purpose, word, f"{word}-{r:02d}")
                os.makedirs(d)
                frames = [RawDataMFCC(raw_samples=save_raw(T, i), mfcc=x, frame_number=i, recording=f"{word}-{r:02d}")
                          for i, x in enumerate(g[key][pos:pos + T])]
                DataStorage.save_raw_data(frames, os.path.join(d, fname))
                pos += T
    return words


@pytest.mark.parametrize("batched,fast", [(True, True), (False, False)])
def test_train_and_test_callers_match_reference_pipeline(batched, fast, tmp_path, monkeypatch, capsys):
    """HMM/main.py train_hmm + test on a Data/ tree, against models and predictions produced by
    the reference's own training_with_save / test_hmm on the same frames (pipeline_ref.npz)."""
    from helpers import assert_close, load_golden
    from hmm_training_b200 import main as hmain
    g = load_golden("pipeline_ref")
    rng = np.random.default_rng(1)
    words = _write_tree(str(tmp_path), g, lambda T, i: rng.normal(size=8) if fast else np.array([]))
    monkeypatch.chdir(tmp_path / "HMM")  # the reference's paths are relative to the CWD (../Data/...)
    models = hmain.train_hmm(show_progress=False, max_iterations=3, batched=batched, fast=fast)
    assert models is not None and sorted(m.word for m in models) == sorted(words)
    by_word = {m.word: m for m in models}
    for w, word in enumerate(words):
        m = by_word[word]
        assert_close(m.A, g["A"][w], f"{word} A"); assert_close(m.B, g["B"][w], f"{word} B")
        assert_close(m.Pi, g["pi"][w], f"{word} pi")
        assert os.path.exists(tmp_path / "Data" / "ResultsHMM" / f"{word}.json")
    true, pred = hmain.test(fast=fast)
    order = np.argsort(np.array(true), kind="stable")  # directory iteration order is arbitrary
    gorder = np.argsort(g["true"], kind="stable")
    assert [true[i] for i in order] == [str(x) for x in g["true"][gorder]]
    assert [pred[i] for i in order] == [str(x) for x in g["pred"][gorder]]
    out = capsys.readouterr().out
    assert "Overall Accuracy: 100.00%" in out and os.path.exists(tmp_path / "Data" / "Plots" / "confusion_matrix.csv")


def _pinned_copy(a):
    """numpy view of pinned host memory (hmmb_host_alloc) holding a copy of `a`."""
    import ctypes
    from hmm_training_b200 import _lib
    lib = _lib.load()
    p = lib.hmmb_host_alloc(a.nbytes)
    assert p
    buf = np.ctypeslib.as_array((ctypes.c_uint8 * a.nbytes).from_address(p)).view(a.dtype).reshape(a.shape)
    buf[...] = a
    return buf, p


@pytest.mark.parametrize("S,ragged", [(40000, False), (80000, False), (60000, True)])
def test_pipelined_upload_matches_plain_create(S, ragged):
    """engine.bw_fit on a large PINNED codeword buffer takes the pipelined path (chunked upload on the
    copy stream, first E-step in stages behind it, parameters uploaded by the create); results must be
    those of create + set_params + iterate on the same data — to rounding: the staged E-step launches from a finer
    work list than the resident iterations, so its partial sums are grouped differently — and a codeword >= M must still
    surface as IndexError.  36 MB of codewords: two upload chunks -> two stages on the compute stream;
    72 MB: four stages alternating between the two side streams; ragged: stage boundaries from the
    sequence offsets (lengths descending inside each word, as the blocked layout wants them)."""
    from hmm_training_b200 import _lib, engine
    N, M, W, T = 4, 256, 6, 150
    obs, offsets, wos = synthetic.fixed_length_codewords(3, W, S, T, N, M)
    if ragged:
        rng = np.random.default_rng(11)
        lens = np.concatenate([np.sort(rng.integers(120, 2 * T + 1, size=S))[::-1] for _ in range(W)]).astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        obs = rng.integers(0, M, size=int(offsets[-1])).astype(np.uint8)
    pi0, A0, B0 = engine.default_init(N, M)
    pi0, A0, B0 = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1))
    pinned, handle = _pinned_copy(obs)
    try:
        a = engine.bw_fit(pinned, offsets, wos, W, N, M, pi0, A0, B0, max_iterations=3)
        with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:  # pageable input: plain path
            bw.set_params(pi0, A0, B0)
            bw.iterate(3, 1e-6, 3)
            b = bw.params() + bw.history(3)
        for x, y, what in zip(a[:4], b[:4], ("pi", "A", "B", "statistic")):
            assert_close(x, y, f"pipelined vs plain create: {what}", rtol=1e-12)
        assert np.array_equal(a[4], b[4])
        with engine.BaumWelch(pinned, offsets, wos, W, N, M, pipeline_upload=True, init=(pi0, A0, B0)) as bw:
            assert bw.kernel_family() == "n4_left_to_right"
            bw.iterate(1, 1e-6, 3)
            bw.set_params(pi0, A0, B0)  # restart on the resident data: every iteration from the resident work list
            bw.iterate(3, 1e-6, 3)
            c = bw.params() + bw.history(3)
        for x, y in zip(b, c):
            assert np.array_equal(x, y, equal_nan=True)  # ... which is what the plain create runs, bit for bit
        a2 = engine.bw_fit(pinned, offsets, wos, W, N, M, pi0, A0, B0, max_iterations=3)
        for x, y in zip(a, a2):
            assert np.array_equal(x, y, equal_nan=True)  # and the pipelined call repeats itself bit for bit
        pinned16, h16 = _pinned_copy(obs.astype(np.uint16))
        try:
            pinned16[len(pinned16) // 2 + 7] = 300  # >= M, in the second half of the upload
            with pytest.raises(IndexError):
                engine.bw_fit(pinned16, offsets, wos, W, N, M, pi0, A0, B0, max_iterations=2)
        finally:
            _lib.load().hmmb_host_free(h16)
    finally:
        _lib.load().hmmb_host_free(handle)


@pytest.mark.parametrize("N,M,S,ragged", [(16, 1024, 600, False), (8, 300, 900, True)])
def test_pipelined_upload_left_to_right_kernels(N, M, S, ragged):
    """Config 4's shape through the pipelined create: pinned codewords + left-to-right initial parameters (pinned or
    pageable) build ONLY the blocked layout, whose repack and first E-step run in stages behind the upload.  The
    accumulators of these kernels take fp64 atomics, so the comparison with the plain path is 1e-12, not bit-exact.
    A later set_params with a dense A must still work (the generic layout is then built from the kept codewords),
    and a codeword >= M must still surface as IndexError."""
    from hmm_training_b200 import _lib, engine
    W, T = 40, 150
    obs, offsets, wos = synthetic.fixed_length_codewords(5, W, S, T, N, M)
    if ragged:
        rng = np.random.default_rng(12)
        lens = np.concatenate([np.sort(rng.integers(100, 2 * T + 1, size=S))[::-1] for _ in range(W)]).astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        obs = rng.integers(0, M, size=int(offsets[-1])).astype(np.uint16)
    assert obs.nbytes >= 4 << 20  # large enough for the chunked upload
    pi0, A0, B0 = engine.default_init(N, M)
    pi0, A0, B0 = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1))
    pinned, handle = _pinned_copy(obs)
    pB, hB = _pinned_copy(B0)
    pA, hA = _pinned_copy(A0)
    pP, hP = _pinned_copy(pi0)
    lib = _lib.load()
    try:
        with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:  # pageable input: plain path, both layouts
            bw.set_params(pi0, A0, B0)
            assert bw.kernel_family() == "left_to_right"
            bw.iterate(3, 1e-6, 3)
            ref = bw.params() + bw.history(3)
        for params in ((pi0, A0, B0), (pP, pA, pB)):
            got = engine.bw_fit(pinned, offsets, wos, W, N, M, *params, max_iterations=3)
            for x, y in zip(got[:4], ref[:4]):
                np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-300)
            assert np.array_equal(got[4], ref[4])
        # dense parameters after a pipelined left-to-right create
        rng = np.random.default_rng(4)
        Ad = rng.random((W, N, N)) + 0.1
        Ad /= Ad.sum(axis=2, keepdims=True)
        with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:
            bw.set_params(pi0, Ad, B0)
            assert bw.kernel_family() == "generic"
            bw.iterate(2, 1e-6, 2)
            refd = bw.params() + bw.history(2)
        for first_iterate in (False, True):
            with engine.BaumWelch(pinned, offsets, wos, W, N, M, pipeline_upload=True, init=(pi0, A0, B0)) as bw:
                assert bw.kernel_family() == "left_to_right"
                if first_iterate:
                    bw.iterate(1, 1e-6, 3)
                bw.set_params(pi0, Ad, B0)
                assert bw.kernel_family() == "generic"
                bw.iterate(2, 1e-6, 2)
                gotd = bw.params() + bw.history(2)
            for x, y in zip(gotd[:4], refd[:4]):
                np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-300)
        bad, hb = _pinned_copy(obs.astype(np.uint16))
        try:
            bad[len(bad) // 2 + 7] = M + 5  # >= M, in the second half of the upload
            with pytest.raises(IndexError):
                engine.bw_fit(bad, offsets, wos, W, N, M, pi0, A0, B0, max_iterations=2)
        finally:
            lib.hmmb_host_free(hb)
    finally:
        for hnd in (handle, hB, hA, hP):
            lib.hmmb_host_free(hnd)


def test_pipelined_upload_with_forward_handover():
    """in the pipelined first E-step the exact kernel runs after the staged backward passes, so
    the per-CTA convergence statistic taken inside k_bw_bwd4 missed every sequence the forward precision guard had
    handed over in that very pass.  Warm start from models with denormal / 1e-300 / zero emissions (what saved
    reference models look like after safe_exp underflow): the pipelined fit must reproduce the plain path's
    log-likelihood history bit for bit, and the case must actually exercise the exact kernel."""
    from hmm_training_b200 import _lib, engine
    from test_gpu_parity import _tiny_models
    N, M, W, S = 4, 16, 3, 70000
    rng = np.random.default_rng(9)
    pi0, A0, B0 = _tiny_models(rng, W, N, M)
    lens = np.concatenate([np.sort(rng.integers(2, 60, size=S))[::-1] for _ in range(W)]).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    obs = rng.integers(0, M, size=int(offsets[-1])).astype(np.uint8)
    wos = np.repeat(np.arange(W, dtype=np.int32), S)
    assert obs.nbytes >= 4 << 20  # large enough for the chunked upload
    pinned, handle = _pinned_copy(obs)
    try:
        with engine.BaumWelch(pinned, offsets, wos, W, N, M, pipeline_upload=True, init=(pi0, A0, B0)) as bw:
            bw.iterate(3, -1.0, 3)
            a = bw.params() + bw.history(3)
            exact_a, _ = bw.diagnostics()
        with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:  # pageable input: plain path
            bw.set_params(pi0, A0, B0)
            bw.iterate(3, -1.0, 3)
            b = bw.params() + bw.history(3)
            exact_b, _ = bw.diagnostics()
        assert exact_a > 0 and exact_b > 0  # (the counts differ: the staged pass runs the exact kernel after its backward passes)
        assert np.isfinite(b[3][:, 0]).all()
        # the statistic of the first iteration is reduced in a fixed order: bit-identical; everything downstream of the
        # exact kernel's fp64 atomics (parameters, later iterations) agrees to rounding
        assert np.array_equal(a[3][:, 0], b[3][:, 0])
        assert np.array_equal(a[4], b[4])
        for x, y, what in zip(a[:4], b[:4], ("pi", "A", "B", "history")):
            assert_close(x, y, what, rtol=1e-11)
    finally:
        _lib.load().hmmb_host_free(handle)


@pytest.mark.parametrize("tiny", [False, True])
def test_pipelined_scorer_matches_plain(tiny, monkeypatch):
    """Recognition on a large PINNED codeword buffer runs in stages behind the upload, with each stage's
    rows of the [U, W] matrix copied back while the other half is scored; results must equal the unpipelined
    path bit for bit — also when the precision guard marks pairs (denormal emissions) and everything is sent
    again after the exact log-space recomputation."""
    from hmm_training_b200 import _lib, engine
    rng = np.random.default_rng(17)
    N, M, W, U, T = 4, 16 if tiny else 256, 6, 400_000, 90  # 36 MB of codewords -> two upload chunks / stages
    if tiny:
        pi = np.zeros((W, N)); pi[:, 0] = 1.0
        A = np.zeros((W, N, N))
        for i in range(N):
            A[:, i, i] = 0.7
            A[:, i, min(i + 1, N - 1)] += 0.3
        B = rng.dirichlet(np.ones(M), size=(W, N))
        tiny_vals = np.array([1e-300, 3e-310, 1e-320, 4.9e-324, 0.0, 1e-250])
        for w in range(W):
            for k in range(M):
                if k % 3 == w % 3:
                    keep = rng.integers(0, N)
                    for j in range(N):
                        if j != keep:
                            B[w, j, k] = tiny_vals[rng.integers(0, len(tiny_vals))]
        obs = rng.integers(0, M, size=U * T).astype(np.uint8)
    else:
        pi0, A0, B0 = engine.default_init(N, M)
        pi, A = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1))
        B = rng.dirichlet(np.ones(M) * 0.3, size=(W, N))
        obs, _, _ = synthetic.fixed_length_codewords(5, 4, U // 4, T, N, M)
    offsets = np.arange(U + 1, dtype=np.int64) * T
    pinned, handle = _pinned_copy(obs)
    try:
        monkeypatch.delenv("HMMB_SCORE_NO_PIPELINE", raising=False)
        ll, arg = engine.score(pinned, offsets, N, M, pi, A, B)
        monkeypatch.setenv("HMMB_SCORE_NO_PIPELINE", "1")
        ll0, arg0 = engine.score(pinned, offsets, N, M, pi, A, B)
        assert np.array_equal(ll, ll0, equal_nan=True) and np.array_equal(arg, arg0)
        assert not np.isnan(ll).any()
        # a pinned result matrix leaves stage by stage on the D2H stream; the pageable one above went through
        # the bounce buffers after the last stage
        monkeypatch.delenv("HMMB_SCORE_NO_PIPELINE", raising=False)
        ll_pin, hll = _pinned_copy(np.zeros((U, W)))
        try:
            ll1, arg1 = engine.score(pinned, offsets, N, M, pi, A, B, out_ll=ll_pin)
            assert np.array_equal(ll1, ll, equal_nan=True) and np.array_equal(arg1, arg)
        finally:
            _lib.load().hmmb_host_free(hll)
        if tiny:
            fin = np.isfinite(ll)
            assert fin.any() and (ll[fin] < -700).any()  # the denormal / exact paths were exercised
        sub = slice(0, 300)
        ref = O.score_batch([obs[u * T:(u + 1) * T].astype(np.int64) for u in range(300)], [(A[w], B[w], pi[w]) for w in range(W)])
        assert_close(ll[sub], ref, "pipelined scorer vs oracle")
    finally:
        _lib.load().hmmb_host_free(handle)


def test_large_pageable_parameters_round_trip():
    """Parameter sets above 4 MB in PAGEABLE memory cross PCIe through the library's pinned bounce buffers
    (multi-threaded memcpy overlapping the DMA) in both directions; set -> get must return the same bits, and
    HMMB_NO_BOUNCE (plain cudaMemcpyAsync) must agree."""
    from hmm_training_b200 import engine
    rng = np.random.default_rng(23)
    N, M, W, T = 8, 1024, 150, 12  # B: 150 x 8 x 1024 doubles = 9.8 MB -> two bounce chunks
    obs = rng.integers(0, M, size=W * T).astype(np.uint16)
    offsets = np.arange(W + 1, dtype=np.int64) * T
    wos = np.arange(W, dtype=np.int32)
    pi = rng.dirichlet(np.ones(N), size=W)
    A = rng.dirichlet(np.ones(N), size=(W, N))
    B = rng.dirichlet(np.ones(M), size=(W, N))
    outs = []
    for env in (None, "1"):
        if env:
            os.environ["HMMB_NO_BOUNCE"] = env
        try:
            with engine.BaumWelch(obs, offsets, wos, W, N, M) as bw:
                bw.set_params(pi, A, B)
                got = bw.params(finalize=False)
                bw.iterate(2, 1e-6, 2)
                outs.append(got + bw.params())
        finally:
            os.environ.pop("HMMB_NO_BOUNCE", None)
    for x, y in zip(outs[0][:3], (pi, A, B)):
        assert np.array_equal(x, y)
    for x, y in zip(outs[0], outs[1]):
        assert np.array_equal(x, y, equal_nan=True)


def test_plain_c_caller_trains_and_recognises(tmp_path):
    """examples/c_caller.c: the C ABI used from plain C (hmmb_bw_fit, hmmb_score, hmmb_vq_encode) — exit code 0 means
    the statistic did not decrease, B rows sum to 1, every utterance was recognised as its own word and every
    centroid encoded to itself."""
    import subprocess
    from test_cabi_load import _build_c_caller
    exe, env = _build_c_caller(tmp_path)
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "16 of 16 utterances" in r.stdout
