"""install() must rebind every holder of the solver names inside the reference package
(src/gglasso/problem.py:10-11, helper/model_selection.py:13, solver/ppdna_solver.py:12).
Needs the reference tree, which only exists in the build container -> skipped elsewhere."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gglasso"), reason="reference tree not available")
def test_install_rebinds_reference_names():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from _refshim import load_reference
    load_reference()
    import gglasso.problem as prob
    import gglasso.helper.model_selection as ms
    import gglasso.solver.admm_solver as am
    import gglasso.solver.single_admm_solver as sm
    import gglasso_b200
    orig = (prob.ADMM_MGL, prob.ADMM_SGL, prob.block_SGL)
    try:
        patched = gglasso_b200.install()
        assert "gglasso.problem.ADMM_MGL" in patched and "gglasso.helper.model_selection.block_SGL" in patched
        assert prob.ADMM_MGL is gglasso_b200.ADMM_MGL and prob.ADMM_SGL is gglasso_b200.ADMM_SGL
        assert prob.block_SGL is gglasso_b200.block_SGL
        assert ms.ADMM_SGL is gglasso_b200.ADMM_SGL and ms.block_SGL is gglasso_b200.block_SGL
        assert am.ADMM_MGL is gglasso_b200.ADMM_MGL and sm.ADMM_SGL is gglasso_b200.ADMM_SGL
    finally:
        prob.ADMM_MGL, prob.ADMM_SGL, prob.block_SGL = orig
