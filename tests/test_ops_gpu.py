"""Op-level GPU checks (each gpinv:: custom op and its backward against torch references)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["check_gemm", "check_kcore", "check_potrf", "check_trsm",
                                  "check_sample", "check_kron", "check_pack", "check_liks", "check_shard"])
def test_op_group(name):
    from tests import gpu_selfcheck as sc
    sc.FAIL.clear()
    getattr(sc, name)()
    assert not sc.FAIL, sc.FAIL
