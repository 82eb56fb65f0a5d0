"""Pin the oracle restatements of the path's callers (prior simulator, posterior summary, optimiser tail)
against outputs of the real reference (tests/golden/{em_ou,em_lv,summary_lv,ema}.pt from make_golden_aux.py)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import oracle_torch as O

GOLD = Path(__file__).parent / "golden"


def _sde(kind):
    return O.OrnsteinUhlenbeck() if kind == "ou" else O.LotkaVolterra()


@pytest.mark.parametrize("name", ["em_ou", "em_lv"])
def test_oracle_euler_maruyama_matches_reference(name):
    g = torch.load(GOLD / f"{name}.pt")
    theta = g["theta"].clone().requires_grad_(True)
    x0 = g["x0"].clone().requires_grad_(True)
    n_steps = round(g["horizon"] / g["dt"])
    paths = O.euler_maruyama(_sde(g["kind"]), x0, theta, n_steps, g["dt"], g["positive_dims"], g["noise"])
    torch.testing.assert_close(paths, g["paths"], rtol=1e-5, atol=1e-6)
    if g["kind"] == "lv":
        assert (g["paths"] == 1e-6).any(), "the golden case must exercise the clamp(min=1e-6) branch"
    obs_idx = (g["obs_times"] / g["dt"]).round().long()
    mse = ((paths[:, obs_idx] - g["obs_values"]) ** 2).mean()
    torch.testing.assert_close(mse, g["mse"], rtol=1e-5, atol=0)
    mse.backward()
    for a, b in ((theta.grad, g["g_theta"]), (x0.grad, g["g_x0"])):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * (b.abs().max().item() + 1e-12))
    # the fused pre-training objective restated from trainer.py:253-259 (x0 = first observation)
    if g["kind"] == "ou":
        th2 = g["theta"].clone()
        m2 = O.pretrain_mse(_sde("ou"), th2, g["obs_times"], g["obs_values"], n_steps, g["dt"], [], g["noise"])
        torch.testing.assert_close(m2, g["mse"], rtol=1e-5, atol=0)


def test_oracle_path_summary_matches_reference():
    g = torch.load(GOLD / "summary_lv.pt")
    x, mean, std = O.path_summary(g["z"], g["positive_dims"])
    torch.testing.assert_close(x, g["x"], rtol=0, atol=0)
    torch.testing.assert_close(mean, g["mean"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(std, g["std"], rtol=1e-6, atol=1e-6)


def test_oracle_adamw_ema_matches_reference():
    g = torch.load(GOLD / "ema.pt")
    params, shadow, norms = O.adamw_ema_steps(g["init"], g["grads"], g["lrs"], g["max_norm"], g["decay"])
    torch.testing.assert_close(norms, g["norms"], rtol=1e-6, atol=0)
    for a, b in zip(params, g["params"]):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
    for a, b in zip(shadow, g["shadow"]):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors published with Random123 (kat_vectors: zero counter / key,
    all-ones, and the pi digits case) pin the integer core of the restatement."""
    import oracle.oracle_torch as M

    def rounds(ctr, key):
        m32 = 0xFFFFFFFF
        c, k0, k1 = list(ctr), key[0], key[1]
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & m32, (p0 >> 32) ^ c[3] ^ k1, p0 & m32]
            k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
        return c

    assert rounds([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert rounds([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert rounds([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    # the vectorised numpy restatement agrees with the scalar rounds and is standard normal
    n = M.philox_normal(seed=0x299F31D0A4093822, B=3, T=5, S=4).numpy()
    c = rounds([4, 0, 2, 0], [0xA4093822, 0x299F31D0])
    u = [((x >> 8) + 0.5) * 2.0**-24 for x in c]
    ref = [np.sqrt(-2 * np.log(u[0])) * np.cos(2 * np.pi * u[1]), np.sqrt(-2 * np.log(u[0])) * np.sin(2 * np.pi * u[1]),
           np.sqrt(-2 * np.log(u[2])) * np.cos(2 * np.pi * u[3]), np.sqrt(-2 * np.log(u[2])) * np.sin(2 * np.pi * u[3])]
    np.testing.assert_allclose(n[2, 4], ref, rtol=1e-12)
    big = M.philox_normal(seed=7, B=64, T=256, S=4).numpy()
    assert abs(big.mean()) < 0.02 and abs(big.std() - 1.0) < 0.02
