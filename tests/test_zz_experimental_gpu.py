"""Collected LAST on purpose (file name): kernels that have not run on a B200 yet.  A device-side fault here cannot take
the CUDA context away from the tests that gate the round."""
import pytest

import train_cases as tc
from oracle.cases import SYN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 1e-4


@pytest.fixture(scope="module")
def lib():
    from sound_bubble_b200 import _lib
    return _lib.load()


def _ok(errs, tol=GRAD_TOL):
    assert all(v <= tol for v in errs.values()), {k: v for k, v in errs.items() if v > tol}


# The attention backward kernels (first version) have been checked on the host-emulated build only; this is their first run
# on a B200, hence non-strict xfail: a pass shows up as XPASS, a failure does not hide the rest of the suite.
@pytest.mark.xfail(strict=False, reason="attention backward kernels: first B200 run (validated on the host-emulated build so far)")
def test_experimental_attention_gradients(lib, monkeypatch):
    from sound_bubble_b200 import training
    monkeypatch.setattr(training, "EXPERIMENTAL_ATTENTION", True)
    _ok(tc.check_golden_grads(lib, DEV, "grad_syn_attn"))
    _ok(tc.check_net(lib, DEV, "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=10), B=2, T=25))
    _ok(tc.check_next_state(lib, DEV, "dis_embed", dict(SYN, B=1, use_attn=True, local_atten_len=4), B=2, T=9))
