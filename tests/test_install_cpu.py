"""install() must rebind every holder of the solver names inside the reference package
(src/gglasso/problem.py:10-11, helper/model_selection.py:13, solver/ppdna_solver.py:12), and uninstall() must
restore them.  Needs oracle/_ref (built by oracle/make_ref.sh from the reference tree) -> skipped where it is absent."""
import pytest

from oracle import ref


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_install_rebinds_reference_names_and_uninstall_restores_them():
    ref.load()
    import gglasso.problem as prob
    import gglasso.helper.model_selection as ms
    import gglasso.solver.admm_solver as am
    import gglasso.solver.single_admm_solver as sm
    import gglasso_b200
    gglasso_b200.uninstall()
    orig = (prob.ADMM_MGL, prob.ADMM_SGL, prob.block_SGL, ms.ADMM_SGL, am.ADMM_MGL)
    assert orig[0] is not gglasso_b200.ADMM_MGL
    try:
        patched = gglasso_b200.install()
        assert "gglasso.problem.ADMM_MGL" in patched and "gglasso.helper.model_selection.block_SGL" in patched
        assert prob.ADMM_MGL is gglasso_b200.ADMM_MGL and prob.ADMM_SGL is gglasso_b200.ADMM_SGL
        assert prob.block_SGL is gglasso_b200.block_SGL
        assert ms.ADMM_SGL is gglasso_b200.ADMM_SGL and ms.block_SGL is gglasso_b200.block_SGL
        assert am.ADMM_MGL is gglasso_b200.ADMM_MGL and sm.ADMM_SGL is gglasso_b200.ADMM_SGL
        assert gglasso_b200.install() == patched          # idempotent: a second install does not stack originals
    finally:
        gglasso_b200.uninstall()
    assert (prob.ADMM_MGL, prob.ADMM_SGL, prob.block_SGL, ms.ADMM_SGL, am.ADMM_MGL) == orig


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_reference_inputs_match_recorded_fingerprints():
    """the reference generators reproduce the inputs the large fixtures were made from (small and medium configs
    here; cfg3 / cfg5 are checked by the GPU tests that use them)."""
    import os
    from oracle import ref_inputs
    fp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large_inputs.json")
    for name in ("cfg1", "cfg2_ggl", "cfg2_fgl", "cfg3_small", "cfg4_small", "cfg4"):
        S = ref_inputs.load(name, cache=False)
        assert ref_inputs.check_fingerprint(name, S, fp) < 1e-9, name
