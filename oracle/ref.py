"""Loader for oracle/_ref: the real reference package built by oracle/make_ref.sh (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke(), bench.py's reference arm / cpu_baseline leg and the input generation of
tests and bench (``ref_inputs``) import this.  Nothing under gglasso_b200/ does; the product path never runs it.
/root/reference is never read here: the package must have been built into oracle/_ref beforehand.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "gglasso", "__init__.py"))


def load():
    """import the reference as ``gglasso`` (from oracle/_ref) and return the package."""
    if not available():
        raise RuntimeError("oracle/_ref/gglasso is missing: run oracle/make_ref.sh in the build container "
                           "(it is git-ignored and travels with the gpurun snapshot)")
    mod = sys.modules.get("gglasso")
    if mod is not None and os.path.dirname(os.path.dirname(os.path.abspath(mod.__file__))) != REF_DIR:
        for name in [n for n in sys.modules if n == "gglasso" or n.startswith("gglasso.")]:
            del sys.modules[name]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module("gglasso")


def fresh_solvers():
    """un-patched solver callables of the reference (``gglasso_b200.install()`` rebinds names inside the package;
    this returns the originals from the defining modules' source by reloading them)."""
    load()
    import gglasso.solver.admm_solver as a
    import gglasso.solver.single_admm_solver as s
    a = importlib.reload(a)
    s = importlib.reload(s)
    return a.ADMM_MGL, s.ADMM_SGL, s.block_SGL
