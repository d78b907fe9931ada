"""Import the UNMODIFIED reference (fabian-sp/GGLasso, /root/reference) in this container.

Only usable where /root/reference exists (the build container). Nothing under tests/ that
runs on the GPU box imports this module; it is used by make_golden.py to generate fixtures.

numba 0.65 rejects ``np.arange(start=..., stop=...)`` keywords inside @njit
(ggl_helper.py:58,169,199,242), so the package is copied to a temp dir and those four
calls are rewritten positionally -- semantics unchanged (route 1).
"""
import os, re, shutil, sys, tempfile

REF_SRC = "/root/reference/src/gglasso"


def load_reference():
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not present (only available in the build container)")
    tmp = os.path.join(tempfile.gettempdir(), "gglasso_ref_shim")
    dst = os.path.join(tmp, "gglasso")
    if not os.path.isdir(dst):
        os.makedirs(tmp, exist_ok=True)
        shutil.copytree(REF_SRC, dst)
        f = os.path.join(dst, "solver", "ggl_helper.py")
        src = open(f).read()
        src2 = re.sub(r"np\.arange\(start\s*=\s*([^,]+?)\s*,\s*stop\s*=\s*([^)]+?)\)", r"np.arange(\1, \2)", src)
        assert src2.count("np.arange(start") == 0
        open(f, "w").write(src2)
    if tmp not in sys.path:
        sys.path.insert(0, tmp)
    import gglasso  # noqa
    return gglasso
