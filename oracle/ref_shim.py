"""TEST INFRASTRUCTURE — never imported by the product path.

Imports the UNMODIFIED reference (DemianMArin/HMM_Training) from /root/reference so
that (a) the numpy restatement in ``oracle/hmm_oracle.py`` can be validated against
it and (b) golden vectors can be generated (``oracle/make_golden.py``).  The reference
tree only exists in the build container, never on the GPU box, so nothing under
``tests/ -m gpu``, ``bench.py`` or ``__graft_entry__.smoke()`` may import this module.

The reference imports four packages that are absent here (librosa, spectrum,
matplotlib, seaborn) at module scope but never *calls* them on the hot path
(HMM/hmm_training.py:4, HMM/hmm_testing.py:10-11, CodeVector/codevector_classes.py:7-8,
CodeVector/codevector_functions.py:12); they are stubbed with empty modules.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("HMM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "HMM", "hmm_training.py"))


_loaded = None


def load():
    """Return a namespace with the reference's hot-path modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    for name in ["librosa", "spectrum", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.ticker", "seaborn"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["spectrum"].poly2lsf = None
    sys.modules["spectrum"].lsf2poly = None
    # Keep our own package's same-named modules out of the way: the reference uses
    # flat imports (``import hmm_training``) resolved through sys.path.
    saved = {k: sys.modules.pop(k) for k in
             ["hmm_training", "hmm_testing", "hmm_classes", "codevector_functions",
              "codevector_classes", "CodeVector", "CodeVector.codevector_classes"]
             if k in sys.modules}
    paths = [os.path.join(REF_ROOT, "HMM"), REF_ROOT, os.path.join(REF_ROOT, "CodeVector")]
    sys.path[:0] = paths
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import hmm_training as ref_training
            import hmm_testing as ref_testing
            import hmm_classes as ref_classes
            import codevector_functions as ref_cvf
            import CodeVector.codevector_classes as ref_cvc
    finally:
        for p in paths:
            sys.path.remove(p)
    ns = types.SimpleNamespace(training=ref_training, testing=ref_testing,
                               classes=ref_classes, cvf=ref_cvf, cvc=ref_cvc)
    # detach the flat names so they cannot shadow anything else later
    for k in ["hmm_training", "hmm_testing", "hmm_classes", "codevector_functions",
              "codevector_classes"]:
        sys.modules.pop(k, None)
    sys.modules.update(saved)
    _loaded = ns
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        yield buf


class LLCapture:
    """Capture the full-precision per-iteration convergence statistic.

    The reference only prints it at %.6f (HMM/hmm_training.py:511); the value is the
    result of the ``log_sum_exp`` call at HMM/hmm_training.py:503.  We wrap the module
    global and record results whose caller line is 503 — the reference is not modified.
    """

    def __init__(self, ref_training):
        self.mod = ref_training
        self.values = []

    def __enter__(self):
        self._orig = self.mod.log_sum_exp
        orig = self._orig
        values = self.values

        def wrapped(x):
            r = orig(x)
            f = sys._getframe(1)
            if f.f_lineno == 503 and f.f_code.co_name == "hmm_training":
                values.append(float(r))
            return r

        self.mod.log_sum_exp = wrapped
        return self

    def __exit__(self, *exc):
        self.mod.log_sum_exp = self._orig
        return False
