"""Drop-in at the CShare seam: the UNMODIFIED reference Decision.cpp / Planning.cpp, compiled against the
product's host/Share.h so that every SearchObstacle / CreateNewPath / BezierPlanning / MeanPoints call they
make is a CUDA launch in libdmpp_b200.so, must publish exactly what they publish with the CPU specification."""
import numpy as np
import pytest

from conftest import assert_records_equal, same
from test_oracle_vs_ref import REC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,cycles", [("highway", 24, 25), ("junction", 8, 70)])
def test_reference_with_gpu_share(the_map, oracle, kind, n, cycles):
    from dmpp_b200 import scenes
    from oracle import binding
    if not binding.Reference.available(gpu_share=True):
        pytest.skip("oracle/_ref/libref_gpu.so not built")
    ref = binding.Reference(gpu_share=True)
    ref.set_map(the_map)
    ep = scenes.Episodes(the_map, np.arange(40, 40 + n), cycles=cycles, kind=kind)
    H, OX, OY = ep.all_cycles()
    got = ref.run(H, OX, OY, calls=False)
    want = oracle.run(H, OX, OY, exhaustive=False)
    clean = want["trace"]["ub_hits"] == 0
    assert_records_equal(got["rec"], want["rec"], REC, mask=clean, what="reference + GPU CShare")
    assert (same(got["path_xy"], want["path_xy"]) | ~clean[..., None, None]).all()
    assert got["traj"] == want["traj"] and got["traj"] > 0
