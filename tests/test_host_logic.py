"""CPU-only tests of the host-side mirror of the reference interface: data formats, helper
semantics, sequence packing, multi-rank plumbing (gloo, world_size 2)."""
import json
import os

import numpy as np
import pytest

from hmm_training_b200 import _lib, codevector_functions as cvf, dist as hdist, hmm_training, synthetic
from hmm_training_b200.codevector_classes import CentroidDataMFCC, DataStorage, RawDataMFCC, frames_matrix
from hmm_training_b200.hmm_classes import DataStorageHMM, HMMTrained
from oracle import hmm_oracle as O


def test_log_helpers_match_reference_semantics():
    x = np.array([0.0, -1.0, 0.5, 1e-320])
    assert np.array_equal(hmm_training.safe_log(x), O.safe_log(x))
    assert hmm_training.safe_log(0.0) == float("-inf") and hmm_training.safe_exp(float("-inf")) == 0.0
    v = np.array([-3.0, float("-inf"), -1.0])
    assert hmm_training.log_sum_exp(v) == O.log_sum_exp(v)
    assert hmm_training.log_sum_exp(np.array([float("-inf")] * 3)) == float("-inf")


def test_model_json_round_trip_matches_reference_layout(tmp_path):
    m = HMMTrained(4, 8, np.eye(4), np.full((4, 8), 0.125), np.array([1.0, 0, 0, 0]), "begin")
    DataStorageHMM.save_hmm(m, str(tmp_path), print_messages=False)
    raw = json.load(open(tmp_path / "begin.json"))
    assert list(raw.keys()) == ["states", "symbols", "A", "B", "Pi", "word"]  # hmm_classes.py:25-34
    assert open(tmp_path / "begin.json").read().startswith('{\n  "states": 4')  # indent=2
    back = DataStorageHMM.load_hmm("begin", str(tmp_path), print_messages=False)
    assert np.array_equal(back.A, m.A) and back.word == "begin" and back.symbols == 8
    (tmp_path / "junk.json").write_text("{not json")
    assert [h.word for h in DataStorageHMM.load_all_hmms(str(tmp_path), print_messages=False)] == ["begin"]


def test_frame_and_centroid_json_layout(tmp_path):
    fr = RawDataMFCC(raw_samples=np.array([]), mfcc=np.arange(13.0), parent_centroid_id=3, generation=2,
                     frame_number=7, recording="r1")
    assert list(fr.to_dict().keys()) == ["raw_samples", "sample_rate", "n_channels", "frame_duration_ms", "mfcc_vector",
                                         "parent_centroid_id", "generation", "frame_number", "recording"]
    DataStorage.save_raw_data([fr], str(tmp_path / "hmm_frames.json"))
    back = DataStorage.load_raw_data_mfcc(str(tmp_path / "hmm_frames.json"), print_messages=False)[0]
    assert np.array_equal(back.mfcc, fr.mfcc) and back.recording == "r1" and back.frame_number == 7
    cents = [CentroidDataMFCC(mfcc=np.ones(13) * i, id=i) for i in range(3)]
    DataStorage.save_centroids(cents, str(tmp_path / "codevector.json"))
    DataStorage.save_generations([cents[:1], cents[:2]], str(tmp_path / "generations.json"))
    assert json.load(open(tmp_path / "codevector.json"))[2] == {"mfcc": [2.0] * 13, "id": 2}
    assert [c.id for c in DataStorage.load_centroids(str(tmp_path / "codevector.json"), print_messages=False)] == [0, 1, 2]
    assert [len(g) for g in DataStorage.load_generations(str(tmp_path / "generations.json"))] == [1, 2]
    DataStorage.save_data_binary(cents, str(tmp_path / "codevector.pkl"))
    assert DataStorage.load_data_binary(str(tmp_path / "codevector.pkl"))[1].id == 1
    with pytest.raises(ValueError):
        frames_matrix([RawDataMFCC(mfcc=np.zeros(12))])


def test_split_and_adjust_host_helpers():
    cents = [CentroidDataMFCC(mfcc=np.full(13, 2.0), id=0), CentroidDataMFCC(mfcc=np.zeros(13), id=1)]
    out = cvf.new_epsilon_centroids(cents)
    assert [c.id for c in out] == [0, 1, 2, 3]
    assert np.array_equal(out[0].mfcc, np.full(13, 2.0) * 1.001) and np.array_equal(out[1].mfcc, np.full(13, 2.0) * 0.999)
    assert np.array_equal(out[2].mfcc, out[3].mfcc)  # twins of an empty (all-zero) centroid stay identical
    assert cvf.new_epsilon_centroids(cents + cents[:1]) is not None  # non power of two: returned unchanged
    frames = [RawDataMFCC(mfcc=np.full(13, float(i)), parent_centroid_id=i % 2, generation=2) for i in range(4)]
    adj = cvf.new_adjust_centroids(frames)
    assert len(adj) == 4 and np.array_equal(adj[0].mfcc, np.full(13, 1.0)) and np.array_equal(adj[3].mfcc, np.zeros(13))
    assert cvf.euclidian_distance(np.zeros(3), np.array([3.0, 4.0, 0.0])) == 5.0
    with pytest.raises(ValueError):
        cvf.euclidian_distance(np.zeros(3), np.zeros(4))


def test_pack_sequences_and_errors():
    obs, off = _lib.pack_sequences([np.array([1, 2, 3]), np.array([], dtype=np.int64), np.array([255])], 256)
    assert obs.dtype == np.uint8 and list(off) == [0, 3, 3, 4]
    obs, _ = _lib.pack_sequences([np.array([1000, 2])], 1024)
    assert obs.dtype == np.uint16
    with pytest.raises(IndexError):
        _lib.pack_sequences([np.array([256])], 256)
    with pytest.raises(IndexError):
        _lib.pack_sequences([np.array([-1])], 256)


def test_synthetic_generators_are_seeded():
    a = synthetic.fixed_length_codewords(3, 4, 5, 20)
    b = synthetic.fixed_length_codewords(3, 4, 5, 20)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].dtype == np.uint8 and a[1][-1] == 400
    assert synthetic.fixed_length_codewords(3, 2, 3, 10, 16, 1024)[0].dtype == np.uint16
    X = synthetic.mfcc_mixture(0, 100)
    assert X.shape == (100, 13) and np.array_equal(X, synthetic.mfcc_mixture(0, 100))


def test_sharding_helpers():
    wos = np.array([0, 0, 0, 1, 1, 2, 2, 2, 2, 0], dtype=np.int32)
    parts = [hdist.shard_sequences_round_robin(wos, r, 3) for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(10))
    for w in range(3):
        counts = [int((wos[p] == w).sum()) for p in parts]
        assert max(counts) - min(counts) <= 1
    assert [hdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ll = np.array([-10.0, -12.0, float("-inf"), -11.0, -400.0])
    stats = []
    for part in (ll[:2], ll[2:3], ll[3:]):
        fin = part[np.isfinite(part)]
        m = fin.max() if len(fin) else float("-inf")
        stats.append((m, float(np.exp(fin - m).sum()) if len(fin) else 0.0))
    assert abs(hdist.combine_llstats(stats) - O.log_sum_exp(ll)) < 1e-12
    assert hdist.combine_llstats([(float("-inf"), 0.0)]) == float("-inf")


def _gloo_worker(rank, world, port, q):
    import ctypes
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the allreduce hook the C library calls, on a host buffer (gloo stands in for NCCL)
    fn = hdist.make_allreduce(device="cpu")
    W, stride = 3, 6
    buf = np.zeros(W * stride + world * W * 2)
    buf[: W * stride] = np.arange(W * stride) * (rank + 1)         # this rank's accumulators
    ll = np.array([[-5.0 - rank, 2.0], [float("-inf"), 0.0], [-7.0 + rank, 1.5]])
    buf[W * stride + rank * W * 2: W * stride + (rank + 1) * W * 2] = ll.reshape(-1)  # own llstats slot only
    fn(buf.ctypes.data_as(ctypes.c_void_p).value, buf.size)
    acc = buf[: W * stride]
    stats = buf[W * stride:].reshape(world, W, 2)
    combined = [hdist.combine_llstats([(stats[r, w, 0], stats[r, w, 1]) for r in range(world)]) for w in range(W)]
    q.put((rank, acc.copy(), combined))
    dist.destroy_process_group()


def test_allreduce_hook_and_llstats_gather_over_gloo():
    """world_size 2 on CPU: accumulators are summed, and zero-filled foreign llstats slots turn
    the sum-allreduce into the all-gather the M-step kernel relies on (-inf survives 0 + -inf)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, acc, comb in res:
        assert np.array_equal(acc, np.arange(18) * 3.0)
        want0 = O.log_sum_exp(np.array([-5.0 + np.log(2.0), -6.0 + np.log(2.0)]))
        assert abs(comb[0] - want0) < 1e-12 and comb[1] == float("-inf")
        assert abs(comb[2] - O.log_sum_exp(np.array([-7.0 + np.log(1.5), -6.0 + np.log(1.5)]))) < 1e-12
    assert np.array_equal(res[0][1], res[1][1])


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hmm_training_b200 import synthetic
    from oracle import vq_oracle
    rng = np.random.default_rng(4)
    seqs = [rng.integers(0, 16, size=int(rng.integers(1, 30))) for _ in range(37)]  # ragged, not divisible by 2
    obs = np.concatenate(seqs).astype(np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).astype(np.int64)
    W, N, M = 3, 4, 16
    pi = rng.dirichlet(np.ones(N), size=W); A = rng.dirichlet(np.ones(N), size=(W, N)); B = rng.dirichlet(np.ones(M), size=(W, N))

    def cpu_scorer(o, off, N_, M_, pi_, A_, B_):  # the oracle stands in for the GPU scorer: this test is about the shards
        ss = [o[off[u]:off[u + 1]].astype(np.int64) for u in range(len(off) - 1)]
        ll = O.score_batch(ss, [(A_[w], B_[w], pi_[w]) for w in range(len(pi_))]) if ss else np.zeros((0, len(pi_)))
        return ll, O.argmax_first(ll).astype(np.int32) if len(ss) else np.zeros(0, np.int32)

    ll, arg = hdist.score_sharded(obs, offsets, N, M, pi, A, B, rank, world, device="cpu", scorer=cpu_scorer)
    X = synthetic.mfcc_mixture(1, 101, K=8); C = synthetic.random_codebook(2, 8)
    idx = hdist.vq_encode_sharded(X, C, rank, world, device="cpu", encoder=vq_oracle.encode)
    full_ll, full_arg = cpu_scorer(obs, offsets, N, M, pi, A, B)
    q.put((rank, bool(np.array_equal(ll, full_ll)), bool(np.array_equal(arg, full_arg)),
           bool(np.array_equal(idx, vq_oracle.encode(X, C)))))
    dist.destroy_process_group()


def test_sharded_scoring_and_encoding_over_gloo():
    """world_size 2 on CPU: utterance / frame ranges per rank, blocks all-gathered in rank order (config 5's
    "1-8 B200" path; the per-shard compute is the GPU's in production and the oracle's here)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: exactly one JSON
    line on stdout with the contract's keys, everything else on stderr."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HMMB_CPU_SEQ_PER_WORD="8")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "baum_welch_frames_per_s_per_iter" and d["value"] > 0
    for key in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
