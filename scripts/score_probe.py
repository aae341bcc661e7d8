"""Phase timing of engine.score on BASELINE config 5 (1M utterances x 10 models)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hmm_training_b200 import _lib, engine, synthetic
lib = _lib.load(); _lib.init(0)
U, Wm = 1_000_000, 10
rng = np.random.default_rng(5)
obs, offsets, _ = synthetic.fixed_length_codewords(77, Wm, U // Wm, 100, 4, 256)
pi, A, B = engine.default_init(4, 256)
Bm = rng.dirichlet(np.ones(256) * 0.3, size=(Wm, 4))
pim, Am = np.tile(pi, (Wm, 1)), np.tile(A, (Wm, 1, 1))
engine.score(obs[:100 * 1000], offsets[:1001], 4, 256, pim, Am, Bm)
for pinned in (False, True):
    o = obs
    if pinned:
        t = torch.empty(obs.shape, dtype=torch.uint8, pin_memory=True); o = t.numpy(); o[:] = obs
    for want_ll in (True, False):
        _lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
        t0 = time.perf_counter()
        ll, arg = engine.score(o, offsets, 4, 256, pim, Am, Bm, want_ll=want_ll)
        dt = time.perf_counter() - t0
        ph = {k: _lib.phase_ms(k)[0] for k in ("prepare", "score_load", "score", "score_exact", "score_argmax")}
        _lib.check(lib.hmmb_set_profiling(0))
        print(f"pinned={pinned} want_ll={want_ll}: {dt*1e3:.1f} ms total -> {U/dt/1e6:.2f} M utt/s; kernels {ph}")
