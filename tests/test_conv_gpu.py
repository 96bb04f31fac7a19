"""K1/K2/K3 parity: tcgen05 implicit-GEMM convolution, transposed convolution and weight gradient
through the C ABI against a plain PyTorch fp32 evaluation of the same op on the same bf16-rounded
operands (tolerances in conv_cases.py)."""
import pytest

import conv_cases as cc
from cycle_depth_estimation_b200 import _lib

pytestmark = pytest.mark.gpu


def _assert_ok(res):
    assert _lib.lib().cdb_device_abort_flag() == 0, "kernel aborted on a bounded mbarrier wait"
    assert res["ok"], res


@pytest.mark.parametrize("name", sorted(cc.FWD_CASES))
def test_conv_fwd(name):
    _assert_ok(cc.conv_fwd_case(**cc.FWD_CASES[name]))


@pytest.mark.parametrize("name", sorted(cc.ROWPACK_CASES))
def test_conv_rowpack(name):
    _assert_ok(cc.conv_rowpack_case(**cc.ROWPACK_CASES[name]))


@pytest.mark.parametrize("name", sorted(cc.WGRAD_CASES))
def test_conv_wgrad(name):
    _assert_ok(cc.conv_wgrad_case(**cc.WGRAD_CASES[name]))


@pytest.mark.parametrize("name", sorted(cc.WGRAD_FEWCOUT_CASES))
def test_conv_wgrad_few_output_channels(name):
    _assert_ok(cc.conv_wgrad_fewcout_case(**cc.WGRAD_FEWCOUT_CASES[name]))


@pytest.mark.parametrize("name", sorted(cc.TF32_FWD_CASES))
def test_conv_fwd_tf32(name):
    """TF32 variant (kind::tf32) of conv / transposed conv / flipped data gradient: relative L2 <= 1e-3 vs float64."""
    _assert_ok(cc.conv_fwd_tf32_case(**cc.TF32_FWD_CASES[name]))


@pytest.mark.parametrize("name", sorted(cc.TF32_WGRAD_CASES))
def test_conv_wgrad_tf32(name):
    _assert_ok(cc.conv_wgrad_tf32_case(**cc.TF32_WGRAD_CASES[name]))
