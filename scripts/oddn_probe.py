"""EM iteration time by state count on config 3's shape at 20 % (10 words x 20 000 sequences x T = 200, M = 256):
N = 4 (one sequence per thread), N = 8 / 16 (left-to-right kernels) and every other N (generic lanes-per-state kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmm_training_b200 import _lib, engine, synthetic
lib = _lib.load(); _lib.init(0)
W, S, T, M = 10, 20000, 200, 256
for N in (3, 4, 5, 6, 8, 12, 16, 32):
    obs, off, wos = synthetic.fixed_length_codewords(3, W, S, T, N, M)
    pi0 = np.zeros((W, N)); pi0[:, 0] = 1.0
    A0 = np.zeros((W, N, N))
    for i in range(N):
        A0[:, i, i] = 0.6 if i + 1 < N else 1.0
        if i + 1 < N:
            A0[:, i, i + 1] = 0.4
    B0 = np.full((W, N, M), 1.0 / M)
    with engine.BaumWelch(obs, off, wos, W, N, M) as bw:
        bw.set_params(pi0, A0, B0)
        bw.iterate(2, -1.0, 1 << 14, sync_each=False)
        _lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
        bw.iterate(3, -1.0, 1 << 14, sync_each=False)
        lib.hmmb_synchronize()
        f, b = _lib.phase_ms("bw_forward"), _lib.phase_ms("bw_backward")
        _lib.check(lib.hmmb_set_profiling(0))
        print(f"N = {N:2d} ({bw.kernel_family():16s}): forward {f[0] / f[1]:7.3f} ms, backward {b[0] / b[1]:7.3f} ms per iteration over {W * S * T / 1e6:.0f} M frames", flush=True)
