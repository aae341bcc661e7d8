"""engine.score end to end on BASELINE config 5 with pinned input / output (HMMB_SCORE_STAGES A/B)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hmm_training_b200 import _lib, engine, synthetic
_lib.init(0)
U, Wm = 1_000_000, 10
rng = np.random.default_rng(5)
obs, offsets, _ = synthetic.fixed_length_codewords(77, Wm, U // Wm, 100, 4, 256)
pi, A, B = engine.default_init(4, 256)
Bm = rng.dirichlet(np.ones(256) * 0.3, size=(Wm, 4))
pim, Am = np.tile(pi, (Wm, 1)), np.tile(A, (Wm, 1, 1))
obs_p = torch.empty(obs.shape, dtype=torch.uint8, pin_memory=True).numpy(); obs_p[:] = obs
ll_p = torch.empty((U, Wm), dtype=torch.float64, pin_memory=True).numpy()
for _ in range(2): engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, out_ll=ll_p)
for want in (True, False):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        if want: engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, out_ll=ll_p)
        else: engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, want_ll=False)
        ts.append(time.perf_counter() - t0)
    print(f"stages={os.environ.get('HMMB_SCORE_STAGES','default')} want_ll={want}: median {np.median(ts)*1e3:.2f} ms min {min(ts)*1e3:.2f} ms")
