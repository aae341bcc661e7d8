"""Dense transition matrices on the lanes-per-state kernels: E-step time at N = 8 / 16 (config 4's alphabet, 20 M frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmm_training_b200 import _lib, engine, synthetic
lib = _lib.load(); _lib.init(0)
rng = np.random.default_rng(0)
W, S, T, M = 200, 500, 200, 1024
for N in (16, 8):
    obs, off, wos = synthetic.fixed_length_codewords(3, W, S, T, N, M)
    pi0 = rng.dirichlet(np.ones(N), size=W); A0 = rng.dirichlet(np.ones(N), size=(W, N)); B0 = rng.dirichlet(np.ones(M) * 4, size=(W, N))
    with engine.BaumWelch(obs, off, wos, W, N, M) as bw:
        bw.set_params(pi0, A0, B0)
        bw.iterate(2, -1.0, 1 << 14, sync_each=False)
        _lib.check(lib.hmmb_set_profiling(1)); _lib.check(lib.hmmb_phase_reset())
        bw.iterate(3, -1.0, 1 << 14, sync_each=False)
        lib.hmmb_synchronize()
        f, b = _lib.phase_ms("bw_forward"), _lib.phase_ms("bw_backward")
        _lib.check(lib.hmmb_set_profiling(0))
        print(f"dense N = {N:2d} ({bw.kernel_family()}): forward {f[0] / f[1]:7.3f} ms, backward {b[0] / b[1]:7.3f} ms per iteration over {W * S * T / 1e6:.0f} M frames", flush=True)
