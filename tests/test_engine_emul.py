"""CPU tests of the SIMT kernel SOURCES through the emulation build (tests/emul/cuda_emul.h):
every CUDA thread is an OS thread, blocks run one at a time.  Same C-ABI, same host code."""
import pytest
import torch

from paddlescience_b200.engine import binding as B
from tests.cases import CASES, TOL, run_case
from tests.emul.build_emul import build


@pytest.fixture(scope="module")
def emul_lib():
    return B.Library(build())


@pytest.mark.parametrize("name", sorted(CASES))
def test_case_matches_oracle(emul_lib, name):
    n = 45 if CASES[name]["hidden"][0] > 64 else 70
    r = run_case(name, n, library=emul_lib, device="cpu")
    tl, tr, tg = TOL[CASES[name]["dtype"]]
    assert r["loss"] <= tl, r
    assert r["res"] <= tr, r
    assert r["grad"] <= tg, r
    assert r["fwd_vs_fused"] == 0.0, r


@pytest.mark.parametrize("name", ["ns_f32", "allen_cahn_period_f64", "biharmonic_f64"])
def test_tile_gemm_path_for_thin_layers(emul_lib, name, monkeypatch):
    """The dedicated first/last-layer kernels can be switched off: the generic tile GEMMs must agree."""
    monkeypatch.setenv("PPSCI_B200_NO_THIN", "1")
    r = run_case(name, 60, library=emul_lib, device="cpu")
    tl, tr, tg = TOL[CASES[name]["dtype"]]
    assert r["loss"] <= tl and r["res"] <= tr and r["grad"] <= tg, r


def test_tcgen05_pair_kernels_through_the_emulated_primitives(emul_lib):
    """The CTA-pair tensor-core kernels (k_tc2_fwd / k_tc2_dx / k_tc2_dw + thin first / last layers) compiled for the
    CPU: UMMA descriptors decoded, 128-byte swizzle, mbarrier phases, TMEM lane quadrants, cta_group::2 operand split,
    round-toward-zero accumulation (tests/emul/cuda_emul.h).  60 points = 3 point tiles: one full tile pair + a pair
    whose second CTA has an empty tile."""
    r = run_case("ns_f32_tc_256", 60, library=emul_lib, device="cpu", backend=2)
    assert r["tc"]
    assert r["loss"] <= 1e-5 and r["res"] <= 1e-5 and r["grad"] <= 2e-5, r
    assert r["fwd_vs_fused"] == 0.0, r
