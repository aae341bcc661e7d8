"""Per-generation cost of the LBG codebook build on BASELINE config 2 (1 M frames, K up to 256)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hmm_training_b200 import _lib, engine, synthetic
lib = _lib.load(); _lib.init(0)
F = 1_000_000
X = synthetic.mfcc_mixture(0, F, 256)
dX = torch.from_numpy(X).cuda()
engine.lbg_fit(None, 4, 5, 1e-3, x_dev_ptr=dX.data_ptr(), F=F)
prev_t, prev_a, prev_n = 0.0, 0.0, 0
for K in (2, 4, 8, 16, 32, 64, 128, 256):
    _lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
    t0 = time.perf_counter()
    C, gens, assign, iters, gd = engine.lbg_fit(None, K, 100, 1e-3, x_dev_ptr=dX.data_ptr(), F=F)
    dt = (time.perf_counter() - t0) * 1e3
    a_ms, a_n = _lib.phase_ms("lbg_assign")
    _lib.check(lib.hmmb_set_profiling(0))
    passes = int(iters[-1])
    print(f"K={K:3d}: generation passes {passes:3d}  wall {dt - prev_t:7.2f} ms ({(dt - prev_t) / passes:.3f} ms/pass)  "
          f"assign kernel {(a_ms - prev_a) / max(a_n - prev_n, 1):.3f} ms/pass   total {dt:.1f} ms")
    prev_t, prev_a, prev_n = dt, a_ms, a_n
