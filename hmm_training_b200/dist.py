"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for
the single exchange step of each path — a sum-allreduce of the small per-iteration
accumulators (pi / xi / emission-count sums, or centroid sums and counts).  Sequences, frames
and utterances are sharded; nothing else crosses ranks.

The C library calls back into ``make_allreduce()``'s closure with a DEVICE pointer; the
closure wraps it zero-copy as a torch tensor and all-reduces it on torch's current stream,
which ``bind_torch_stream()`` has made the library's stream as well, so ordering needs no
host synchronisation.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import numpy as np

from . import _lib


class _DeviceBuffer:
    """Zero-copy view of n fp64 values at a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


_bound_stream = None


def bind_torch_stream():
    """Make libhmmb200 launch on torch's current CUDA stream, so torch events, NCCL
    collectives and our kernels are ordered without host synchronisation.  The legacy default
    stream (handle 0) cannot be handed over, so a dedicated torch stream is made current
    first in that case.  Returns the torch stream."""
    global _bound_stream
    import torch
    s = torch.cuda.current_stream()
    if s.cuda_stream == 0:
        s = torch.cuda.Stream()
        torch.cuda.set_stream(s)
    _bound_stream = s  # keep it alive
    _lib.check(_lib.load().hmmb_set_stream(ctypes.c_void_p(s.cuda_stream)))
    return s


def make_allreduce(group=None, device: str = "cuda", overlap: bool = False):
    """Returns fn(ptr, n_doubles) that sums the buffer in place over the process group.
    device='cpu' treats ptr as host memory (gloo; used by the CPU tests of the plumbing).
    overlap=True (with BaumWelch.set_overlap): every collective runs on a side stream, ordered behind what is
    queued on the library's stream at the time of the call, so that it overlaps the kernels queued afterwards;
    the library's closing fn(0, 0) call makes its stream wait for the side stream."""
    import torch
    import torch.distributed as dist
    comm_stream = torch.cuda.Stream() if (overlap and device != "cpu") else None

    def fn(ptr: int, n: int) -> None:
        if n == 0:  # join
            if comm_stream is not None:
                torch.cuda.current_stream().wait_stream(comm_stream)
            return
        if comm_stream is not None:
            lib_stream = _lib.load().hmmb_get_stream()
            if lib_stream and torch.cuda.current_stream().cuda_stream != lib_stream:
                raise RuntimeError("overlapping all-reduce: torch's current stream is not the library's (call "
                                   "dist.bind_torch_stream() and keep that stream current)")
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                t = torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return
        if device == "cpu":
            buf = (ctypes.c_double * n).from_address(ptr)
            t = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return
        # The collective must be queued on the stream the library launches on: on any other stream it would race
        # with k_bw_reduce / k_bw_mstep.  Run it there explicitly instead of trusting torch's current stream.
        lib_stream = _lib.load().hmmb_get_stream()
        t = torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda")
        if lib_stream and torch.cuda.current_stream().cuda_stream != lib_stream:
            with torch.cuda.stream(torch.cuda.ExternalStream(lib_stream)):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return fn


def native_comm_init(rank: int, world: int, group=None) -> None:
    """Create the library's OWN NCCL communicator (hmmb_comm_init) — the path a C caller without torch takes.  The
    128-byte NCCL id is made on rank 0 and travels through torch.distributed here; afterwards pass
    ``allreduce="native"`` to engine.bw_fit / BaumWelch.set_dist / engine.lbg_fit."""
    import torch.distributed as dist
    lib = _lib.load()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.hmmb_comm_unique_id(buf, 128))
    box = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    _lib.check(lib.hmmb_comm_init(int(rank), int(world), box[0]))


def native_comm_destroy() -> None:
    _lib.check(_lib.load().hmmb_comm_destroy())


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sequences_round_robin(word_of_seq: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Indices of this rank's sequences: the sequences of every word are dealt round-robin over
    the ranks, so each rank sees (almost) the same number of sequences of every word."""
    word_of_seq = np.asarray(word_of_seq)
    order = np.argsort(word_of_seq, kind="stable")
    ws = word_of_seq[order]
    starts = np.flatnonzero(np.r_[True, ws[1:] != ws[:-1]])
    pos_in_word = np.arange(len(ws)) - np.repeat(starts, np.diff(np.r_[starts, len(ws)]))
    return np.sort(order[pos_in_word % world == rank])


def combine_llstats(stats: Sequence[Tuple[float, float]]) -> float:
    """log_sum_exp over all ranks' sequences from per-rank (max, sum exp(ll - max)) pairs —
    the host mirror of what k_bw_mstep does on the device (HMM/hmm_training.py:503)."""
    mx = max((m for m, _ in stats), default=float("-inf"))
    if mx == float("-inf"):
        return float("-inf")
    s = sum(sv * np.exp(m - mx) for m, sv in stats if m != float("-inf"))
    return float(mx + np.log(s))


def cpu_affinity_info(device_index: int) -> dict:
    """What bind_to_gpu_cpus sees: CPUs NVML calls local to the GPU, CPUs this process may run on, their overlap."""
    import os
    info = {"cpu_count": os.cpu_count()}
    try:
        info["allowed"] = len(os.sched_getaffinity(0))
    except Exception as exc:
        info["allowed_error"] = repr(exc)
    try:
        import pynvml
        pynvml.nvmlInit()
        phys = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 1, 1024) + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        info["nvml_local"] = len(local)
        info["overlap"] = len(local & os.sched_getaffinity(0))
    except Exception as exc:
        info["nvml_error"] = repr(exc)
    return info


def bind_to_gpu_cpus(device_index: int) -> int:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (its NUMA node), so
    that the pinned host buffers it allocates afterwards (first touch) and the library's host threads sit
    next to that GPU's PCIe root.  With eight ranks uploading at once this is what decides the aggregate
    host-to-device bandwidth.  Returns the number of CPUs bound to (0 = left unchanged: NVML or the
    affinity call unavailable, or the container's CPU set does not meet the GPU's)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        phys = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(ncpu, 1024) + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target or target == allowed:
            return 0
        os.sched_setaffinity(0, target)
        return len(target)
    except Exception:
        return 0


# ----------------------------------------------------------------------------- pure shards
def _gather_rows(local: np.ndarray, counts: Sequence[int], group=None, device: str = "cuda") -> np.ndarray:
    """All-gather of per-rank row blocks (rank r holds counts[r] rows) into the full array on every rank.  The
    blocks are padded to the largest one so that one all_gather moves them (NCCL over NVLink for device='cuda',
    gloo for the CPU tests of this plumbing)."""
    import torch
    import torch.distributed as dist
    world = len(counts)
    if world == 1:
        return local
    mx = max(int(c) for c in counts)
    shape = (mx,) + tuple(local.shape[1:])
    pad = np.zeros(shape, dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device != "cpu":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o.cpu().numpy()[: int(c)] for o, c in zip(outs, counts)], axis=0)


def score_sharded(obs, offsets, N: int, M: int, pi, A, B, rank: int, world: int, group=None, device: str = "cuda",
                  gather: bool = True, scorer=None):
    """test_hmm's scoring loop (HMM/hmm_testing.py:139-153) sharded over `world` ranks: rank r scores the contiguous
    utterance range shard_range(U, r, world) against all models on its GPU; no collective on the data path, the
    [U_r, W] blocks and argmax vectors are all-gathered afterwards (gather=False returns the local block and its
    range instead).  Returns (ll [U, W], argmax [U]) on every rank."""
    from . import engine
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    U = len(offsets) - 1
    lo, hi = shard_range(U, rank, world)
    loc_off = offsets[lo:hi + 1] - offsets[lo]
    loc_obs = obs[int(offsets[lo]):int(offsets[hi])]
    ll, arg = (scorer or engine.score)(loc_obs, loc_off, N, M, pi, A, B)
    if not gather:
        return ll, arg, (lo, hi)
    counts = [shard_range(U, r, world)[1] - shard_range(U, r, world)[0] for r in range(world)]
    return (_gather_rows(np.ascontiguousarray(ll), counts, group, device),
            _gather_rows(np.ascontiguousarray(arg), counts, group, device))


def vq_encode_sharded(X: np.ndarray, C: np.ndarray, rank: int, world: int, group=None, device: str = "cuda",
                      gather: bool = True, encoder=None):
    """get_observations' frame loop (HMM/hmm_training.py:95-118) sharded over ranks: contiguous frame ranges, no
    collective on the data path, indices all-gathered afterwards."""
    from . import engine
    F = X.shape[0]
    lo, hi = shard_range(F, rank, world)
    idx = (encoder or engine.vq_encode)(X[lo:hi], C)
    if not gather:
        return idx, (lo, hi)
    counts = [shard_range(F, r, world)[1] - shard_range(F, r, world)[0] for r in range(world)]
    return _gather_rows(np.ascontiguousarray(idx), counts, group, device)
