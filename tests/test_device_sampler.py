"""Device-side uniform sampling (Philox4x32-10, ``ppsci_b200_sample_uniform``) — checked through the CPU emulation build
of the same kernel: known-answer vector of the generator, range, determinism, disjoint counter ranges, and use as the
``input`` callable of ContinuousNamedArrayDataset."""
import numpy as np
import pytest
import torch

import ppsci
from paddlescience_b200.engine import binding as B


def _philox_ref(ctr, key):
    """Philox4x32-10 restated with numpy integers (Salmon et al. 2011); known answer from the Random123 KAT file:
    counter 0 key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    c = [int(v) for v in ctr]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c


def test_philox_known_answer():
    assert _philox_ref([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


@pytest.fixture()
def emul(monkeypatch):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))


def test_sampler_matches_the_restated_generator_and_is_deterministic(emul):
    smp = ppsci.data.dataset.DeviceUniformSampler(("t", "x", "y"), (0.0, -1.0, 2.0), (1.0, 1.0, 5.0), 257, seed=0x1234567890, device="cpu")
    a = smp()
    b = smp()
    assert set(a) == {"t", "x", "y"} and a["x"].shape == (257, 1)
    for k, (lo, hi) in zip(("t", "x", "y"), ((0, 1), (-1, 1), (2, 5))):
        assert float(a[k].min()) >= lo and float(a[k].max()) < hi
    # counter ranges of consecutive calls are disjoint: different points; a fresh sampler with the same seed repeats them
    assert not torch.equal(a["x"], b["x"])
    smp2 = ppsci.data.dataset.DeviceUniformSampler(("t", "x", "y"), (0.0, -1.0, 2.0), (1.0, 1.0, 5.0), 257, seed=0x1234567890, device="cpu")
    assert torch.equal(smp2()["y"], a["y"])
    # word-for-word against the restated generator (fp32: top 24 bits of word 2 j of block d // 2)
    key = [0x1234567890 & 0xFFFFFFFF, 0x1234567890 >> 32]
    for i in (0, 1, 200, 256):
        w = _philox_ref([i, 0, 0, 0], key)
        assert float(a["t"][i]) == pytest.approx(np.float32(0.0 + 1.0 * ((w[0] >> 8) / 16777216.0)), rel=0, abs=1e-7)
        assert float(a["x"][i]) == pytest.approx(np.float32(-1.0 + 2.0 * ((w[2] >> 8) / 16777216.0)), rel=0, abs=1e-7)
        w2 = _philox_ref([i, 0, 1, 0], key)
        assert float(a["y"][i]) == pytest.approx(np.float32(2.0 + 3.0 * ((w2[0] >> 8) / 16777216.0)), rel=0, abs=3e-7)


def test_sampler_feeds_the_continuous_dataset(emul):
    smp = ppsci.data.dataset.DeviceUniformSampler(("x", "y"), (0, 0), (1, 1), 64, seed=7, device="cpu", dtype=torch.float64)
    ds = ppsci.data.dataset.ContinuousNamedArrayDataset(smp, lambda inp: {"u": inp["x"] * 0}, None)
    it = iter(ds)
    i1, l1, _ = next(it)
    i2, _, _ = next(it)
    assert i1["x"].dtype == torch.float64 and l1["u"].shape == (64, 1) and not torch.equal(i1["x"], i2["x"])
    u = torch.cat([i1["x"], i2["x"], i1["y"], i2["y"]])
    assert 0.35 < float(u.mean()) < 0.65
