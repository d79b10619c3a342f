"""GPU bring-up: a case with several tile pairs per CTA (multi-pass persistent loops), small enough for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.cases import run_case
n = int(os.environ.get("N_POINTS", 9000))
r = run_case("ns_f32_tc_256", n, device="cuda:0", backend=2)
print({k: r[k] for k in ("loss", "res", "grad", "tc", "launches")})
