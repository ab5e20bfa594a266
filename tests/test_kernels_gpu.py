"""GPU parity tests of every kernel behind the C ABI (see tests/kernel_checks.py)."""
import pytest

import kernel_checks


@pytest.mark.gpu
@pytest.mark.parametrize("group", list(kernel_checks.GROUPS))
def test_kernel_group(group):
    kernel_checks.GROUPS[group]()
