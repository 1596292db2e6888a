"""Helpers for the `-m gpu` parity tests: everything goes through the C ABI (cornell_moe_b200.capi)."""
import numpy as np

import oracle as orc


def checker():
    """The CPU checker: the compiled reference when its .so travelled with the snapshot, else the C oracle."""
    return orc.load_reference() if orc.have_reference() else orc.load_oracle()


def tril_close(A, B, rtol, atol):
    np.testing.assert_allclose(np.tril(A), np.tril(B), rtol=rtol, atol=atol)
