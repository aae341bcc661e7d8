"""What a thin state costs: config 3's shape at 20 % with utterances of only TWO acoustic segments on the 4-state model,
so that the tail states die during training; prints the time of every EM iteration and the number of rescued states."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmm_training_b200 import _lib, engine
lib = _lib.load(); _lib.init(0)
W, S, T, N, M = 10, 20000, 200, 4, 256
rng = np.random.default_rng(5)
R = W * S
cut = rng.integers(40, 160, size=(R, 1))
tg = np.arange(T)[None, :]
obs = np.where(tg < cut, rng.integers(0, 40, size=(R, T)), 128 + rng.integers(0, 40, size=(R, T)))
obs = ((obs + (np.arange(R) // S * 7)[:, None]) % M).astype(np.uint8)
off = np.arange(R + 1, dtype=np.int64) * T
wos = np.repeat(np.arange(W, dtype=np.int32), S)
pi0, A0, B0 = engine.default_init(N, M)
pi0, A0, B0 = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1))
with engine.BaumWelch(obs.reshape(-1), off, wos, W, N, M) as bw:
    bw.set_params(pi0, A0, B0)
    for it in range(1, 21):
        lib.hmmb_synchronize(); t0 = time.perf_counter()
        bw.iterate(1, -1.0, 100)
        lib.hmmb_synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print(f"iteration {it:2d}: {dt:8.2f} ms, thin states {bw.thin_states()}, diagnostics {bw.diagnostics()}", flush=True)
