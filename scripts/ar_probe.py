"""8-rank probe: latency of the per-iteration accumulator all-reduce (84 KB fp64) through the same hook the
library uses, and the EM iteration time with and without the collective."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from hmm_training_b200 import _lib, engine, synthetic, dist as hdist
_lib.init(lr); hdist.bind_torch_stream()
buf = torch.zeros(10455 + 2 * world * 10, dtype=torch.float64, device="cuda")
for _ in range(20): dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize()
ar_us = e0.elapsed_time(e1) / 200 * 1e3
W, S, T, N, M = 10, 100000, 200, 4, 256
obs, off, wos = synthetic.fixed_length_codewords(1000 + rank, W, S, T, N, M)
pi0, A0, B0 = engine.default_init(N, M)
init = (np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1)))
res = {}
for mode in ("allreduce", "none"):
    bw = engine.BaumWelch(obs, off, wos, W, N, M)
    bw.set_params(*init)
    if mode == "allreduce":
        bw.set_dist(rank, world, hdist.make_allreduce())
    bw.iterate(3, -1.0, 100, sync_each=False)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record(); bw.iterate(10, -1.0, 100, sync_each=False); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    res[mode] = [round(float(x), 3) for x in allt]
    bw.close()
if rank == 0:
    print(f"all_reduce of {buf.numel() * 8} B back to back: {ar_us:.1f} us each")
    print("ms per iteration with all-reduce, per rank:", res["allreduce"])
    print("ms per iteration without (independent ranks):", res["none"])
torch.cuda.synchronize(); _lib.load().hmmb_set_stream(None); _lib.load().hmmb_shutdown()
dist.barrier(); dist.destroy_process_group()
