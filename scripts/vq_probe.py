"""Kernel-level VQ encode rate (device-resident frames), 1 M frames x 256 centroids."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hmm_training_b200 import _lib, synthetic
lib = _lib.load(); _lib.init(0)
F, K = 1_000_000, 256
X = synthetic.mfcc_mixture(0, F, K); C = synthetic.random_codebook(1, K)
dX, dC = torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda()
dI = torch.empty(F, dtype=torch.int32, device="cuda")
call = lambda: _lib.check(lib.hmmb_vq_encode_dev(dX.data_ptr(), F, dC.data_ptr(), K, dI.data_ptr(), None))
for _ in range(3): call()
lib.hmmb_synchronize()
_lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
for _ in range(20): call()
for k in ("vq_encode", "vq_exact"):
    print(k, _lib.phase_ms(k))
ms, n = _lib.phase_ms("vq_encode")
print(f"vq_encode: {ms / n:.4f} ms per launch, {F / (ms / n * 1e-3) / 1e9:.3f} G frames/s, {F * K * 36 / (ms / n * 1e-3) / 1e12:.2f} TFLOP/s fp64, checksum {int(dI.sum())}")

ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.check(lib.hmmb_set_profiling(0))
from hmm_training_b200 import dist
dist.bind_torch_stream()
for _ in range(3): call()
ev0.record(torch.cuda.current_stream())
for _ in range(20): call()
ev1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
print(f"whole call (events, profiling off): {ev0.elapsed_time(ev1) / 20:.4f} ms")
