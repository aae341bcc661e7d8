"""Multi-GPU path: sequences sharded over ranks, one NCCL all-reduce of the accumulators per
EM iteration.  Needs >= 2 GPUs (skipped otherwise); spawns one process per GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["HMMB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["HMMB_ROOT"], "tests"))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from hmm_training_b200 import _lib, engine, synthetic, dist as hdist
from helpers import load_golden
_lib.init(int(os.environ["LOCAL_RANK"])); hdist.bind_torch_stream()
g = load_golden("bw_c1_clustered_s0_it10")
obs, off, wos = g["obs"], g["offsets"], g["word_of_seq"]
mine = hdist.shard_sequences_round_robin(wos, rank, world)
seqs = [obs[off[r]:off[r + 1]] for r in mine]
sobs = np.concatenate(seqs); soff = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
W, N, M = 10, 4, 256
pi0, A0, B0 = engine.default_init(N, M)
out = engine.bw_fit(sobs, soff, wos[mine], W, N, M, np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)),
                    max_iterations=10, allreduce=hdist.make_allreduce(), rank=rank, world=world)
# LBG with frames sharded over ranks
X = synthetic.mfcc_mixture(3, 600, K=16)
lo, hi = hdist.shard_range(600, rank, world)
C, gens, assign, iters, gd = engine.lbg_fit(X[lo:hi], 32, 100, 0.001, allreduce=hdist.make_allreduce())
# left-to-right N = 16 models: plain all-reduce vs word groups overlapped with the backward pass
rng = np.random.default_rng(5)
N16, M16, W16 = 16, 1024, 6
corpus = [synthetic.clustered_sequences(rng, 40, N=N16, M=M16, tmin=20, tmax=60, shift=11 * w, spread=32) for w in range(W16)]
o16, off16, wos16 = synthetic.pack_corpus(corpus, M16)
mine16 = hdist.shard_sequences_round_robin(wos16, rank, world)
seqs16 = [o16[off16[r]:off16[r + 1]] for r in mine16]
so16 = np.concatenate(seqs16); soff16 = np.concatenate([[0], np.cumsum([len(q) for q in seqs16])]).astype(np.int64)
p16, a16, b16 = engine.default_init(N16, M16)
init16 = (np.tile(p16, (W16, 1)), np.tile(a16, (W16, 1, 1)), np.tile(b16, (W16, 1, 1)))
ltr = {}
for name, groups in (("plain", 1), ("overlap", 3)):
    with engine.BaumWelch(so16, soff16, wos16[mine16], W16, N16, M16) as bw:
        bw.set_params(*init16)
        assert bw.kernel_family() == "left_to_right"
        bw.set_dist(rank, world, hdist.make_allreduce(overlap=groups > 1))
        bw.set_overlap(groups)
        bw.iterate(3, 1e-6, 3)
        ltr[name] = bw.params() + bw.history(3)
# the library's own NCCL communicator (hmmb_comm_init; libnccl opened with dlopen) instead of the torch hook:
# a two-rank sum has one order, so the results must be identical bit for bit
hdist.native_comm_init(rank, world)
out_native = engine.bw_fit(sobs, soff, wos[mine], W, N, M, np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)),
                           max_iterations=10, allreduce="native", rank=rank, world=world)
for x, y in zip(out, out_native):
    assert np.array_equal(x, y, equal_nan=True), "native communicator disagrees with the torch.distributed hook"
C2 = engine.lbg_fit(X[lo:hi], 32, 100, 0.001, allreduce="native")[0]
assert np.allclose(C, C2, rtol=1e-12, atol=1e-12)  # LBG sums use fp64 atomics: order-dependent at 1e-16
hdist.native_comm_destroy()
np.savez(os.environ["HMMB_OUT"] + f".{rank}.npz", pi=out[0], A=out[1], B=out[2], hist=out[3], iters=out[4], C=C,
         lbg_iters=iters, assign=assign, **{f"ltr_{k}_{i}": v for k, r in ltr.items() for i, v in enumerate(r)})
torch.cuda.synchronize()
_lib.load().hmmb_set_stream(None); _lib.load().hmmb_shutdown()
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_training_matches_reference_golden(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import assert_close, load_golden
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HMMB_ROOT=ROOT, HMMB_OUT=str(tmp_path / "out"))
    port = 29500 + os.getpid() % 1000
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], check=True, env=env, timeout=600)
    g = load_golden("bw_c1_clustered_s0_it10")
    r0, r1 = np.load(str(tmp_path / "out") + ".0.npz"), np.load(str(tmp_path / "out") + ".1.npz")
    for k in ("pi", "A", "B", "hist", "iters", "C", "lbg_iters"):
        assert np.array_equal(r0[k], r1[k], equal_nan=True), f"ranks disagree on {k}"  # replicated M-step
    assert np.array_equal(r0["iters"], g["iters"])
    assert_close(r0["hist"][:, :10], g["ll_hist"], "ll")
    assert_close(r0["A"], g["A"], "A"); assert_close(r0["B"], g["B"], "B"); assert_close(r0["pi"], g["pi"], "pi")
    # left-to-right path: ranks agree, overlapped == plain bit for bit, and both match a single-process fit
    from hmm_training_b200 import engine, synthetic
    rng = np.random.default_rng(5)
    N16, M16, W16 = 16, 1024, 6
    corpus = [synthetic.clustered_sequences(rng, 40, N=N16, M=M16, tmin=20, tmax=60, shift=11 * w, spread=32) for w in range(W16)]
    o16, off16, wos16 = synthetic.pack_corpus(corpus, M16)
    p16, a16, b16 = engine.default_init(N16, M16)
    single = engine.bw_fit(o16, off16, wos16, W16, N16, M16, np.tile(p16, (W16, 1)), np.tile(a16, (W16, 1, 1)),
                           np.tile(b16, (W16, 1, 1)), max_iterations=3)
    for i in range(5):
        assert np.array_equal(r0[f"ltr_plain_{i}"], r1[f"ltr_plain_{i}"], equal_nan=True)
        assert np.array_equal(r0[f"ltr_plain_{i}"], r0[f"ltr_overlap_{i}"], equal_nan=True)
        assert np.array_equal(r0[f"ltr_overlap_{i}"], r1[f"ltr_overlap_{i}"], equal_nan=True)
        assert_close(r0[f"ltr_overlap_{i}"], single[i], f"2-rank left-to-right vs single process [{i}]", rtol=1e-10)
    gl = load_golden("lbg_600_k32")
    assert np.array_equal(r0["lbg_iters"], gl["iters"])
    assert_close(r0["C"], gl["C"], "centroids", atol=1e-12)
    assert np.array_equal(np.concatenate([r0["assign"], r1["assign"]]), gl["assign"])
