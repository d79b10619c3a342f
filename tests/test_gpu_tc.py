"""GPU tests of the tcgen05 path: parity against the oracle through the C-ABI with backend=2."""
import pytest
import torch

from tests.cases import CASES, TOL, run_case

pytestmark = pytest.mark.gpu

TC_CASES = {
    "ns_f32_tc_256": dict(in_keys=("x", "y"), out_keys=("u", "v", "p"), hidden=[256, 256, 256], act="tanh",
                          exprs=CASES["ns_f32"]["exprs"], dtype=torch.float32),
    "ns_f32_tc_128_sin": dict(in_keys=("x", "y"), out_keys=("u", "v", "p"), hidden=[128, 128, 128], act="sin",
                              exprs=CASES["ns_f32"]["exprs"], dtype=torch.float32),
    "ac_f32_tc_128": dict(in_keys=("t", "x"), out_keys=("u",), hidden=[128] * 4, act="tanh",
                          exprs=CASES["allen_cahn_period_f32"]["exprs"], dtype=torch.float32,
                          periods={"x": (2.0, False)}, oracle_exprs=CASES["allen_cahn_period_f32"]["oracle_exprs"],
                          ranges={"x": (-1, 1)}),
}
CASES.update(TC_CASES)


@pytest.mark.parametrize("name", sorted(TC_CASES))
def test_tc_case_matches_oracle(name):
    assert torch.cuda.is_available()
    r = run_case(name, 3000, device="cuda:0", backend=2)
    assert r["tc"], "tcgen05 backend was not selected"
    tl, tr, tg = TOL[torch.float32]
    assert r["loss"] <= tl, r
    assert r["res"] <= tr, r
    assert r["grad"] <= tg, r
