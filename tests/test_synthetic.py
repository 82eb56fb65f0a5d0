"""The benchmark's input factory (product side) and the oracle's factory produce the same tensors."""
import torch

from oracle import oracle_torch as O
from viforsdes_b200.synthetic import make_inputs


def test_synthetic_inputs_match_oracle_factory():
    for kind, kw in (("ou", {}), ("lv", {}), ("l96", dict(state_dim=5))):
        a = make_inputs(kind, 3, 40, context_dim=12, hidden_dim=16, num_layers=2, seed=3, **kw)
        p = O.make_problem(kind, 3, 40, context_dim=12, hidden_dim=16, num_layers=2, seed=3, **kw)
        assert torch.equal(a.x0, p.x0) and torch.equal(a.theta, p.theta) and torch.equal(a.eps, p.eps)
        assert torch.equal(a.context_full[:, :-1], p.context)
        for x, y in zip([*a.w_ih, *a.w_hh, *a.b_ih, *a.b_hh, a.out_w, a.out_b], p.weights.tensors()):
            assert torch.equal(x, y)
        assert torch.equal(a.obs_times, p.obs_times) and torch.equal(a.obs_values, p.obs_values)
        assert a.obs_variance == p.obs_variance and tuple(a.positive_dims) == tuple(p.positive_dims)
        assert torch.equal(a.obs_idx, O.obs_indices(p.obs_times, p.dt, 40))
