"""A few cases of tools/fuzz_parity.py in the suite (the tool itself runs hundreds): random non-cubic meshes, source
positions, sub-box sizes, SED mixes, clumping / LLS hooks and partially ionized states through one source pass and one
global pass, GPU against oracle."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1001, 1005, 1021, 1032, 1055, 2053, 2082, 2202])
def test_random_case(seed):
    import fuzz_parity
    tag, upd = fuzz_parity.one_case(seed)
    assert upd > 0, tag
