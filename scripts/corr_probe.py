"""How the N = 4 E-step depends on how many sequences of a block see the same codeword at the same step: config 3's
shape (10 words x S sequences x T = 200, M = 256) with the per-segment alphabet narrowed from the benchmark's 40
codewords down to 2 (the count update of k_bw_bwd4 takes one round per rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmm_training_b200 import _lib, engine
lib = _lib.load(); _lib.init(0)
W, S, T, N, M = 10, int(os.environ.get("S", "40000")), 200, 4, 256
pi0, A0, B0 = engine.default_init(N, M)
pi0, A0, B0 = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1))
for spread in (256, 40, 16, 8, 4, 2, 1):
    rng = np.random.default_rng(1)
    R = W * S
    if spread == 256:
        obs = rng.integers(0, M, size=(R, T)).astype(np.uint8)
    else:
        cuts = np.sort(rng.integers(0, T + 1, size=(R, N - 1)), axis=1)
        seg = (np.arange(T)[None, :, None] >= cuts[:, None, :]).sum(axis=2)
        obs = ((seg * (M // N) + rng.integers(0, spread, size=(R, T)) + (np.arange(R) // S * 7)[:, None]) % M).astype(np.uint8)
    off = np.arange(R + 1, dtype=np.int64) * T
    wos = np.repeat(np.arange(W, dtype=np.int32), S)
    with engine.BaumWelch(obs.reshape(-1), off, wos, W, N, M) as bw:
        bw.set_params(pi0, A0, B0)
        bw.iterate(2, -1.0, 1 << 14, sync_each=False)
        _lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
        bw.iterate(5, -1.0, 1 << 14, sync_each=False)
        lib.hmmb_synchronize()
        f, b = _lib.phase_ms("bw_forward"), _lib.phase_ms("bw_backward")
        _lib.check(lib.hmmb_set_profiling(0))
        print(f"codewords per segment {spread:3d}: forward {f[0] / f[1]:.3f} ms, backward {b[0] / b[1]:.3f} ms per iteration "
              f"({R * T / 1e6:.0f} M frames), diagnostics {bw.diagnostics()}", flush=True)
