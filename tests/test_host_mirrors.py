"""Host-side mirrors of the reference interface that run in plain PyTorch (no GPU): the single-step
``DiffusionTransitionHead.forward`` (models/head.py:68-97) and ``StateSpace`` (inference/state_space.py) against the
oracle's restatement of the same reference lines, values and gradients."""
import pytest
import torch

from oracle import oracle_torch as O
from tests._util import build_head


@pytest.mark.parametrize("kind,S,NL,H", [("ou", None, 2, 16), ("lv", None, 1, 8), ("l96", 5, 3, 12)])
def test_head_forward_matches_oracle_step(kind, S, NL, H):
    p = O.make_problem(kind, 4, 3, context_dim=6, hidden_dim=H, num_layers=NL, state_dim=S, dtype=torch.float64)
    w = p.weights
    # push some diagonal entries below DIAG_MIN so the floored branch and its gradient rule are exercised
    w.out_b[w.state_dim] = -0.5
    head = build_head(p, device="cpu").double()
    with torch.no_grad():  # build_head copied through fp32 parameters: restore the exact float64 weights
        for k in range(NL):
            for nm, src in (("weight_ih", w.w_ih), ("weight_hh", w.w_hh), ("bias_ih", w.b_ih), ("bias_hh", w.b_hh)):
                getattr(head.gru, f"{nm}_l{k}").copy_(src[k])
        head.out_proj.weight.copy_(w.out_w)
        head.out_proj.bias.copy_(w.out_b)
    hidden0 = [0.1 * torch.randn(4, H, dtype=torch.float64) for _ in range(NL)]
    z = p.x0.clone().requires_grad_(True)
    mu_r, L_r, hid_r = O.head_step(w.map(lambda t: t.clone().requires_grad_(True)), z, p.context[:, 0], p.theta, hidden0)
    z2 = p.x0.clone().requires_grad_(True)
    mu, L, hid = head(z2, p.context[:, 0], p.theta, torch.stack(hidden0))
    assert torch.allclose(mu, mu_r, atol=1e-12) and torch.allclose(L, L_r, atol=1e-12)
    assert torch.allclose(hid, torch.stack(hid_r), atol=1e-12)
    assert (torch.triu(L, diagonal=1) == 0).all() and (torch.diagonal(L, dim1=-2, dim2=-1) >= 1e-2).all()
    g = torch.randn_like(L)
    (L_r * g).sum().backward()
    (L * g).sum().backward()
    assert torch.allclose(z2.grad, z.grad, atol=1e-12)
    # hidden=None starts from zeros like nn.GRU
    mu0, _, _ = head(z2.detach(), p.context[:, 0], p.theta)
    mu0_r, _, _ = O.head_step(w, p.x0, p.context[:, 0], p.theta, [torch.zeros(4, H, dtype=torch.float64)] * NL)
    assert torch.allclose(mu0, mu0_r, atol=1e-12)


def test_floored_diagonal_gradient_rule():
    from viforsdes_b200.head import _FlooredDiagonal

    x = torch.tensor([0.5, 0.001, 0.001, 0.01], requires_grad=True)
    y = _FlooredDiagonal.apply(x, 1e-2)
    assert torch.equal(y.detach(), torch.tensor([0.5, 0.01, 0.01, 0.01]))
    y.backward(torch.tensor([1.0, 1.0, -1.0, 1.0]))
    assert torch.equal(x.grad, torch.tensor([1.0, 0.0, -1.0, 1.0]))  # primitives/bounds.py:20


def test_state_space_matches_oracle():
    from viforsdes_b200.state_space import StateSpace

    torch.manual_seed(5)
    z = torch.randn(3, 7, 4, dtype=torch.float64) * 15
    z[0, 0, 1] = 30.0  # beyond softplus' threshold
    for pos in ([], [1, 3], [0, 1, 2, 3]):
        sp = StateSpace(4, pos)
        zz = z.clone().requires_grad_(True)
        zr = z.clone().requires_grad_(True)
        x, xr = sp.to_state(zz), O.to_state(zr, pos)
        # the vectorised and the gathered-column softplus of torch's CPU backend differ in the last bit for some arguments
        assert torch.allclose(x, xr, rtol=4e-16, atol=0)
        lj = sp.log_jacobian(zz[:, 1:]).sum(-1)
        assert torch.allclose(lj, O.log_jacobian(zr, pos), atol=1e-12)
        (x.sum() + lj.sum()).backward()
        (xr.sum() + O.log_jacobian(zr, pos).sum()).backward()
        assert torch.allclose(zz.grad, zr.grad, atol=1e-12)
        xs = torch.rand(5, 4, dtype=torch.float64) * 3
        assert torch.allclose(sp.to_latent(xs), O.to_latent(xs, pos), atol=1e-12)
        assert torch.allclose(sp.to_state(sp.to_latent(xs)), xs.clamp(min=1e-6) if len(pos) == 4 else sp.to_state(sp.to_latent(xs)))
        assert sp.positive_mask == sum(1 << d for d in pos)
    for bad in (dict(dim=0), dict(dim=2, positive_dims=[2]), dict(dim=2, positive_dims=[0, 0])):
        with pytest.raises(ValueError):
            StateSpace(**bad)
