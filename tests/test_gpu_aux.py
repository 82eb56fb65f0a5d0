"""GPU parity tests for the callers either side of the path: the fused Euler-Maruyama prior simulator
(forward + reverse-mode), the posterior-sample summary, the fused clip + AdamW + EMA step and the context-producer
fold, each against the CPU oracle and the committed reference outputs (tests/golden/*.pt)."""
from __future__ import annotations

from pathlib import Path

import pytest
import torch
from torch import nn

from oracle import oracle_torch as O
from tests._util import assert_close, assert_parity, build_head

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _sdes(kind):
    from viforsdes_b200 import sde as vs

    return (vs.OrnsteinUhlenbeck(), O.OrnsteinUhlenbeck()) if kind == "ou" else (vs.LotkaVolterra(), O.LotkaVolterra())


def _em_inputs(kind, B, T, seed, near_clamp=False):
    g = torch.Generator().manual_seed(seed)
    if kind == "ou":
        theta = torch.stack([torch.exp(0.3 * torch.randn(B, generator=g)), 1 + 0.3 * torch.randn(B, generator=g),
                             torch.exp(-1 + 0.3 * torch.randn(B, generator=g))], 1)
        x0 = 0.5 + 0.2 * torch.randn(B, 1, generator=g)
        pos = []
    else:
        theta = torch.exp(torch.log(torch.tensor([0.5, 0.0025, 0.3])) + 0.1 * torch.randn(B, 3, generator=g))
        x0 = torch.tensor([71.0, 79.0]).expand(B, 2) * torch.exp(0.1 * torch.randn(B, 2, generator=g))
        pos = [0, 1]
    noise = torch.randn(B, T, x0.shape[1], generator=g)
    if near_clamp and kind == "lv" and B >= 4:
        x0[:4] = torch.tensor([[2e-6, 3.0], [1.5, 1e-6], [1e-5, 1e-5], [4e-6, 40.0]])
        noise[:4, : min(T, 4)] = -2.5
    return x0.contiguous(), theta, noise, pos


# ---------------------------------------------------------------------------------------------------------------
# Euler-Maruyama simulator
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["em_ou", "em_lv"])
def test_em_matches_reference_golden(name):
    from viforsdes_b200.euler_maruyama import euler_maruyama

    g = torch.load(GOLD / f"{name}.pt")
    sde, _ = _sdes(g["kind"])
    x0 = g["x0"].cuda().requires_grad_(True)
    theta = g["theta"].cuda().requires_grad_(True)
    paths = euler_maruyama(sde, x0, theta, g["horizon"], g["dt"], g["positive_dims"], noise=g["noise"].cuda())
    assert_close(paths, g["paths"], name="paths")
    if g["kind"] == "lv":
        assert (paths == 1e-6).any()
    obs_idx = (g["obs_times"] / g["dt"]).round().long().cuda()
    mse = ((paths[:, obs_idx] - g["obs_values"].cuda()) ** 2).mean()
    assert_close(mse, g["mse"], name="mse")
    mse.backward()
    assert_close(theta.grad, g["g_theta"], name="g_theta")
    assert_close(x0.grad, g["g_x0"], name="g_x0")


@pytest.mark.parametrize("kind,B,T", [("ou", 70, 37), ("lv", 70, 45), ("lv", 1, 1), ("ou", 33, 32), ("lv", 5, 16),
                                      ("lv", 64, 200)])
def test_em_fwd_bwd_vs_oracle(kind, B, T):
    """Ragged warps (B % 32 != 0), step counts that do not divide the 32/S-step chunk, clamp hits; random cotangents on the
    whole trajectory (not only the observation points)."""
    x0, theta, noise, pos = _em_inputs(kind, B, T, seed=B * 1000 + T, near_clamp=True)
    gsde, osde = _sdes(kind)
    gp = torch.randn(B, T + 1, x0.shape[1], generator=torch.Generator().manual_seed(1))
    refs = []
    for dt_ in (torch.float32, torch.float64):
        a, th = x0.detach().clone().to(dt_).requires_grad_(True), theta.detach().clone().to(dt_).requires_grad_(True)
        p = O.euler_maruyama(osde, a, th, T, 0.05, pos, noise.to(dt_))
        p.backward(gp.to(dt_))
        refs.append((p.detach(), th.grad, a.grad))
    a, th = x0.cuda().requires_grad_(True), theta.cuda().requires_grad_(True)
    from viforsdes_b200 import _lib
    from viforsdes_b200.euler_maruyama import _mask

    p = torch.ops.visde.em_fwd(a, th, noise.cuda(), 0, T, 0.05, gsde.device_kind, _mask(pos))
    p.backward(gp.cuda())
    assert _lib.load() is not None
    for got, i, nm in ((p, 0, "paths"), (th.grad, 1, "g_theta"), (a.grad, 2, "g_x0")):
        assert_parity(got, refs[0][i], refs[1][i], name=f"{kind} B={B} T={T} {nm}")


def test_em_philox_noise():
    from viforsdes_b200.euler_maruyama import euler_maruyama, philox_normal
    from viforsdes_b200.sde import LotkaVolterra

    seed, B, T = 0x1234_5678_9ABC_DEF0, 77, 53
    n = philox_normal(seed, B, T, 2)
    ref = O.philox_normal(seed, B, T, 2)
    assert (n.cpu().double() - ref).abs().max().item() < 2e-5
    x0, theta, _, pos = _em_inputs("lv", B, T, seed=3)
    a, th = x0.cuda(), theta.cuda().requires_grad_(True)
    p1 = euler_maruyama(LotkaVolterra(), a, th, T * 0.05, 0.05, pos, noise=None, seed=seed)
    th2 = theta.cuda().requires_grad_(True)
    p2 = euler_maruyama(LotkaVolterra(), a, th2, T * 0.05, 0.05, pos, noise=n)
    assert torch.equal(p1, p2), "in-kernel Philox draws differ from visde_philox_normal"
    gp = torch.randn_like(p1)
    p1.backward(gp)
    p2.backward(gp)
    assert torch.equal(th.grad, th2.grad), "backward must regenerate the forward's draws"
    p3 = euler_maruyama(LotkaVolterra(), a, th, T * 0.05, 0.05, pos, noise=None, seed=seed + 1)
    assert not torch.equal(p1, p3)
    # S = 1 uses the first normal of the same block
    assert torch.equal(philox_normal(seed, B, T, 1)[..., 0], n[..., 0])


def test_philox_noise_for_the_path_op_wide_state():
    """Counter-based noise for sample_diffusion_paths (opt-in replacement of torch.randn, diffusion_path_sampler.py:57): S = 10
    draws against the numpy restatement, shard-layout independence via batch_offset, and the sampler wiring."""
    from viforsdes_b200.euler_maruyama import philox_normal
    from viforsdes_b200.observations import Observations
    from viforsdes_b200.sampler import sample_diffusion_paths
    from viforsdes_b200.state_space import StateSpace

    seed, B, T, S = 0x1234ABCD5678, 37, 19, 10
    n = philox_normal(seed, B, T, S)
    ref = O.philox_normal(seed, B, T, S)
    assert (n.cpu().double() - ref).abs().max().item() < 2e-5
    assert torch.equal(n[..., :4], philox_normal(seed, B, T, 4)), "dims 0..3 must be the original four-normal stream"
    assert abs(n.mean().item()) < 0.05 and abs(n.std().item() - 1.0) < 0.05
    p = O.make_problem("l96", 8, 12, context_dim=16, hidden_dim=32, num_layers=1, state_dim=S)
    head = build_head(p).eval()
    ctx_full = torch.zeros(8, 13, 16, device="cuda")
    ctx_full[:, :12] = p.context.cuda()
    enc = lambda *a: ctx_full  # noqa: E731
    obs = Observations(times=p.obs_times, values=p.obs_values)
    full = sample_diffusion_paths(enc, head, obs, p.theta.cuda(), p.x0.cuda(), 0.6, p.dt, StateSpace(S), seed=seed)
    again = sample_diffusion_paths(enc, head, obs, p.theta.cuda(), p.x0.cuda(), 0.6, p.dt, StateSpace(S), seed=seed)
    assert torch.equal(full.z, again.z), "seeded sampling must be reproducible"
    enc_hi = lambda *a: ctx_full[4:]  # noqa: E731
    shard = sample_diffusion_paths(enc_hi, head, obs, p.theta.cuda()[4:], p.x0.cuda()[4:], 0.6, p.dt, StateSpace(S), seed=seed,
                                   batch_offset=4)
    assert torch.equal(shard.z, full.z[4:]), "a shard with batch_offset must reproduce its rows of the full batch"


def test_em_pretraining_size_properties():
    """inference/trainer.py:208-259 sizes (4096 simulations, LV horizon 40 at dt 0.05 = 800 steps): bit-determinism, finite
    values, linearity of the reverse mode in the cotangent, fp64 oracle on a slice, and the fused objective."""
    from viforsdes_b200.euler_maruyama import _mask, pretrain_mse
    from viforsdes_b200.sde import LotkaVolterra

    B, T = 4096, 800
    x0, theta, _, pos = _em_inputs("lv", B, T, seed=9)
    x0[:] = torch.tensor([71.0, 79.0])
    a, seed = x0.cuda(), 2024
    kind, mask = LotkaVolterra.device_kind, _mask(pos)
    th = theta.cuda().requires_grad_(True)
    p = torch.ops.visde.em_fwd(a, th, None, seed, T, 0.05, kind, mask)
    assert torch.isfinite(p).all() and (p >= 1e-6).all()
    assert torch.equal(p, torch.ops.visde.em_fwd(a, th, None, seed, T, 0.05, kind, mask))
    g1, g2 = torch.randn_like(p) * 1e-3, torch.randn_like(p) * 1e-3
    outs = [torch.autograd.grad(p, th, g, retain_graph=True)[0] for g in (g1, g2, 2.0 * g1 - 0.5 * g2)]
    assert torch.equal(outs[0], torch.autograd.grad(p, th, g1, retain_graph=True)[0])
    lin = 2.0 * outs[0] - 0.5 * outs[1]
    assert (outs[2] - lin).abs().max().item() <= 1e-4 * lin.abs().max().item()
    # slice against the fp64 oracle with the same draws
    from viforsdes_b200.euler_maruyama import philox_normal

    sl = slice(100, 108)
    noise = philox_normal(seed, B, T, 2)[sl].cpu()
    refs = []
    for dt_ in (torch.float32, torch.float64):
        t_ = theta[sl].detach().clone().to(dt_).requires_grad_(True)
        pr = O.euler_maruyama(O.LotkaVolterra(), x0[sl].to(dt_), t_, T, 0.05, pos, noise.to(dt_))
        pr.backward(g1[sl].cpu().to(dt_))
        refs.append((pr.detach(), t_.grad))
    assert_parity(p[sl], refs[0][0], refs[1][0], name="paths slice")
    assert_parity(outs[0][sl], refs[0][1], refs[1][1], name="g_theta slice")
    obs_t = torch.tensor([0.0, 10.0, 20.0, 30.0, 40.0]).cuda()
    obs_v = torch.tensor([[71.0, 79.0], [120.0, 60.0], [160.0, 140.0], [60.0, 200.0], [50.0, 90.0]]).cuda()
    mse = pretrain_mse(LotkaVolterra(), th, obs_t, obs_v, 40.0, 0.05, pos, seed=seed)
    idx = (obs_t / 0.05).round().long()
    assert torch.allclose(mse, ((p[:, idx] - obs_v) ** 2).mean())
    mse.backward()
    assert torch.isfinite(th.grad).all()


def test_em_user_sde_steps_in_pytorch_and_errors():
    from viforsdes_b200 import _lib
    from viforsdes_b200.euler_maruyama import euler_maruyama

    sde = O.Lorenz96(6)
    g = torch.Generator().manual_seed(0)
    x0, theta = torch.randn(5, 6, generator=g), torch.stack([8 + torch.randn(5, generator=g), 0.3 * torch.ones(5)], 1)
    noise = torch.randn(5, 12, 6, generator=g)
    ref = O.euler_maruyama(sde, x0, theta, 12, 0.05, [], noise)
    got = euler_maruyama(sde, x0.cuda(), theta.cuda(), 0.6, 0.05, [], noise=noise.cuda())
    assert_close(got, ref, name="l96 paths")
    with pytest.raises(ValueError):
        euler_maruyama(sde, x0.cuda(), theta.cuda(), 0.6, -0.05)
    with pytest.raises(ValueError):
        euler_maruyama(sde, x0.cuda(), theta.cuda(), 0.0, 0.05)
    lib = _lib.load()
    assert lib.visde_em_fwd(4, 4, _lib.SDE_GENERIC, 0, 0.05, None, None, None, 0, None, None) == _lib.EINVAL
    assert lib.visde_em_fwd(0, 4, _lib.SDE_OU, 0, 0.05, None, None, None, 0, None, None) == _lib.OK  # empty batch
    assert lib.visde_em_fwd(4, 4, _lib.SDE_OU, 0, 0.0, None, None, None, 0, None, None) == _lib.EINVAL


# ---------------------------------------------------------------------------------------------------------------
# posterior summary
# ---------------------------------------------------------------------------------------------------------------
def test_summary_matches_reference_golden():
    from viforsdes_b200.posterior import summarise_paths
    from viforsdes_b200.state_space import StateSpace

    g = torch.load(GOLD / "summary_lv.pt")
    x, mean, std = summarise_paths(g["z"].cuda(), StateSpace(2, g["positive_dims"]))
    assert_close(x, g["x"], rtol=1e-6, name="x")
    assert_close(mean, g["mean"], rtol=1e-5, name="mean")
    assert_close(std, g["std"], rtol=1e-5, name="std")


@pytest.mark.parametrize("n,T1,S,pos", [(1000, 801, 2, [0, 1]), (1, 5, 1, []), (33, 7, 3, [1]), (5000, 101, 10, [])])
def test_summary_vs_oracle(n, T1, S, pos):
    from viforsdes_b200.posterior import summarise_paths
    from viforsdes_b200.state_space import StateSpace

    z = 4.0 * torch.randn(n, T1, S, generator=torch.Generator().manual_seed(n)) + 2.0
    x64, m64, s64 = O.path_summary(z.double(), pos)
    x, mean, std = summarise_paths(z.cuda(), StateSpace(S, pos))
    assert_close(x, x64, rtol=1e-6, name="x")
    assert_close(mean, m64, rtol=1e-5, name="mean")
    if n == 1:
        assert torch.isnan(std).all()  # torch.std of one sample
    else:
        assert_close(std, s64, rtol=1e-5, name="std")
    _, mean2, std2 = summarise_paths(z.cuda(), StateSpace(S, pos), want_x=False)
    assert torch.equal(mean, mean2) and torch.equal(std.nan_to_num(), std2.nan_to_num())


def test_posterior_sample_and_summary_through_the_head():
    """VariationalPosterior.sample / summary flow (variational_posterior.py:93-135) with a stub encoder: eval-mode,
    stash-less forward, x = from_latent(z), mean / std over the samples."""
    from viforsdes_b200.observations import Observations
    from viforsdes_b200.posterior import sample_posterior, summarise_posterior
    from viforsdes_b200.state_space import StateSpace

    p = O.make_problem("lv", 48, 30, context_dim=16, hidden_dim=32, num_layers=2)
    head = build_head(p)
    ctx_full = torch.zeros(48, 31, 16)
    ctx_full[:, :-1] = p.context

    class Post:
        def rsample(self, n):
            return p.theta.cuda()[:n]

    enc = lambda *a: ctx_full.cuda()  # noqa: E731
    obs = Observations(times=p.obs_times.cuda(), values=p.obs_values.cuda())
    ss = StateSpace(2, [0, 1])
    s = sample_posterior(enc, head, Post(), obs, 48, 30 * p.dt, p.dt, ss, noise=p.eps.cuda())
    assert head.training, "sampling must restore the training flag"
    x0 = p.obs_values[0].unsqueeze(0).expand(48, -1)
    zref, _, _ = O.sample_paths(p.weights, O.to_latent(x0, [0, 1]), p.context, p.theta, p.eps, p.dt)
    xref = O.to_state(zref, [0, 1])
    assert_close(s.diffusion_paths, xref.detach(), name="posterior x")
    summ = summarise_posterior(enc, head, Post(), obs, 30 * p.dt, p.dt, ss, n_samples=48, noise=p.eps.cuda())
    assert_close(summ.diffusion_path_mean, xref.mean(0).detach(), name="path mean")
    assert_close(summ.diffusion_path_std, xref.std(0).detach(), rtol=1e-3, name="path std")
    assert_close(summ.sde_parameter_mean, p.theta.mean(0), name="theta mean")


# ---------------------------------------------------------------------------------------------------------------
# fused clip + AdamW + EMA
# ---------------------------------------------------------------------------------------------------------------
def _golden_module(g):
    model = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 3))
    with torch.no_grad():
        for p, v in zip(model.parameters(), g["init"]):
            p.copy_(v)
    return model.cuda()


def test_fused_adamw_ema_matches_reference_golden():
    from viforsdes_b200.optim import FlatParameters, FusedAdamWEma

    g = torch.load(GOLD / "ema.pt")
    model = _golden_module(g)
    flat = FlatParameters([list(model[0].parameters()), list(model[2].parameters())])
    opt = FusedAdamWEma(flat, lrs=[1e-2, 3e-3], max_norm=g["max_norm"], ema_decay=g["decay"])
    for step, gs in enumerate(g["grads"]):
        flat.zero_grad()
        for p, gr in zip(model.parameters(), gs):
            p.grad.copy_(gr)  # gradients land in the flat bucket through the views
        opt.step()
        assert abs(opt.grad_norm.item() - g["norms"][step].item()) <= 1e-5 * g["norms"][step].item()
    for p, ref in zip(model.parameters(), g["params"]):
        assert_close(p, ref, rtol=1e-5, atol_scale=1e-6, name="param")
    for s, ref in zip(opt.ema_views(), g["shadow"]):
        assert_close(s, ref, rtol=1e-5, atol_scale=1e-6, name="ema shadow")


@pytest.mark.parametrize("n,max_norm,scale,ema", [(351_000, 1.0, 1.0, 0.999), (1_000_003, 0.0, 1.0, None),
                                                  (4099, 0.5, 1024.0, 0.9), (3, 10.0, 1.0, 0.5)])
def test_fused_adamw_ema_vs_oracle(n, max_norm, scale, ema):
    """Head-sized (351 KB bucket) and encoder-sized buffers, unaligned tails, GradScaler inv_scale, no-clip and no-EMA."""
    from viforsdes_b200.optim import FlatParameters, FusedAdamWEma

    gen = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=gen)
    grads = [[scale * torch.randn(n, generator=gen) * (10.0 if k == 1 else 0.1)] for k in range(4)]
    rp, rs, rn = O.adamw_ema_steps([p0.double()], [[g[0].double()] for g in grads], [2e-3], max_norm,
                                   ema if ema is not None else 0.0, inv_scale=1.0 / scale)
    param = nn.Parameter(p0.clone().cuda())
    flat = FlatParameters([[param]])
    opt = FusedAdamWEma(flat, lrs=[2e-3], max_norm=max_norm, ema_decay=ema)
    inv = torch.tensor([1.0 / scale], device="cuda") if scale != 1.0 else None
    for k, g in enumerate(grads):
        param.grad.copy_(g[0])
        opt.step(inv_scale=inv)
        if max_norm > 0:
            assert abs(opt.grad_norm.item() - rn[k].item()) <= 2e-5 * rn[k].item()
    assert_close(param, rp[0], rtol=2e-5, atol_scale=2e-6, name="param")
    if ema is not None:
        assert_close(opt.ema_views()[0], rs[0], rtol=2e-5, atol_scale=2e-6, name="ema")
    else:
        assert opt.ema_views() == []


# ---------------------------------------------------------------------------------------------------------------
# context-producer fold
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E,Cd,B,T", [(24, 16, 5, 13), (128, 128, 3, 40)])
def test_context_fold_matches_unfused(E, Cd, B, T):
    """head(tokens . W_op^T + b_op) == head_folded(tokens): outputs and every gradient, including both factors of the fold."""
    p = O.make_problem("lv", B, T, context_dim=Cd, hidden_dim=64, num_layers=2)
    g = torch.Generator().manual_seed(E)
    tokens = 0.3 * torch.randn(B, T, E, generator=g)
    proj = nn.Linear(E, Cd)
    outs = []
    for folded in (False, True):
        head = build_head(p)
        pr = nn.Linear(E, Cd).cuda()
        pr.load_state_dict(proj.state_dict())
        tk = tokens.cuda().requires_grad_(True)
        x0, th, eps = p.x0.cuda(), p.theta.cuda().requires_grad_(True), p.eps.cuda()
        if folded:
            paths, means, chol = head.sample_diffusion_paths_from_tokens(x0, tk, pr, th, eps, p.dt)
        else:
            paths, means, chol = head.sample_diffusion_paths(x0, pr(tk), th, eps, p.dt)
        gen = torch.Generator().manual_seed(7)
        loss = sum((o * torch.randn(o.shape, generator=gen).cuda()).sum() for o in (paths, means, chol))
        loss.backward()
        outs.append((paths.detach(), means.detach(), chol.detach(), tk.grad, th.grad, pr.weight.grad, pr.bias.grad,
                     head.gru.weight_ih_l0.grad, head.gru.bias_ih_l0.grad, head.gru.weight_hh_l1.grad, head.out_proj.weight.grad))
    names = ["paths", "means", "chol", "g_tokens", "g_theta", "g_proj_w", "g_proj_b", "g_w_ih_l0", "g_b_ih_l0", "g_w_hh_l1", "g_out_w"]
    for a, b, nm in zip(outs[1], outs[0], names):
        assert_close(a, b, rtol=2e-4, atol_scale=5e-5, name=nm)


def test_pretrain_loop_matches_reference_control_flow():
    """pretrain_sde_parameters (sync-free: device Adam, device best-so-far) against the reference's loop structure
    (inference/trainer.py:208-250: torch.optim.Adam, clip_grad_norm_, three .item() per step) on the same draws."""
    from viforsdes_b200.euler_maruyama import PretrainConfig, pretrain_mse, pretrain_sde_parameters
    from viforsdes_b200.sde import OrnsteinUhlenbeck

    sde, dt, horizon = OrnsteinUhlenbeck(), 0.05, 5.0
    true = torch.tensor([1.2, 0.8, 0.25])
    obs_t = torch.arange(0.0, 5.5, 1.0).cuda()
    obs_v = (true[1] + (0.2 - true[1]) * torch.exp(-true[0] * obs_t.cpu())).unsqueeze(1).cuda()
    cfg = PretrainConfig(n_iterations=120, batch_size=512, learning_rate=0.05, init_scale=0.5)
    pos = [0, 2]

    torch.manual_seed(11)
    best = pretrain_sde_parameters(sde, obs_t, obs_v, horizon, dt, pos, [], cfg, seed=100)

    torch.manual_seed(11)
    d = 3
    mu = torch.zeros(d, device="cuda")
    mu[[1]] = cfg.init_scale * torch.randn(1, device="cuda")
    mu = torch.nn.Parameter(mu)
    log_sigma = torch.nn.Parameter(torch.zeros(d, device="cuda"))
    opt = torch.optim.Adam([mu, log_sigma], lr=cfg.learning_rate)
    best_ref, best_mse, first_mse = mu.detach().clone(), float("inf"), None
    for step in range(cfg.n_iterations):
        opt.zero_grad()
        eps = torch.randn(cfg.batch_size, d, device="cuda")
        log_theta = mu + log_sigma.exp() * eps
        theta = log_theta.clone()
        theta[:, pos] = log_theta[:, pos].exp()
        mse = pretrain_mse(sde, theta, obs_t, obs_v, horizon, dt, [], seed=100 + step)
        first_mse = mse.item() if first_mse is None else first_mse
        if torch.isfinite(mse) and mse.item() < best_mse:
            best_ref, best_mse = mu.detach().clone(), mse.item()
        if torch.isfinite(mse):
            mse.backward()
            torch.nn.utils.clip_grad_norm_([mu, log_sigma], 1.0)
            opt.step()
    assert_close(best, best_ref, rtol=1e-4, atol_scale=1e-4, name="best_mu")
    assert best_mse < 0.5 * first_mse, (best_mse, first_mse)
    # the fitted mean reverts towards the generating parameters (theta_1 is identified by the plateau of the observations)
    assert abs(best[1].item() - true[1].item()) < 0.25


def _module_iteration(inp, head, theta_req=True):
    """One ELBO iteration through the public modules on the tensors of a synthetic `Inputs` (what PathIteration fuses)."""
    from viforsdes_b200 import sde as vs
    from viforsdes_b200.elbo import path_elbo_terms
    from viforsdes_b200.observations import GaussianObservationLikelihood, Observations
    from viforsdes_b200.state_space import StateSpace
    from viforsdes_b200.types import DiffusionPathSample

    x0 = inp.x0.cuda().requires_grad_(True)
    full = inp.context_full.cuda().requires_grad_(True)
    theta = inp.theta.cuda().requires_grad_(theta_req)
    paths, means, chol = head.sample_diffusion_paths(x0, full[:, :-1], theta, inp.eps.cuda(), inp.dt)
    S = inp.x0.shape[1]
    sample = DiffusionPathSample(paths, means, chol, StateSpace(S, list(inp.positive_dims)))
    sde = vs.OrnsteinUhlenbeck() if inp.kind == "ou" else vs.LotkaVolterra()
    terms = path_elbo_terms(sde, Observations(times=inp.obs_times, values=inp.obs_values),
                            GaussianObservationLikelihood(variance=inp.obs_variance), theta, sample, inp.dt)
    loss = -(terms[:, 0] + terms[:, 1] - terms[:, 2] + terms[:, 3]).mean()
    loss.backward()
    return paths, terms, x0.grad, full.grad, theta.grad, loss


def _head_from_inputs(inp):
    from viforsdes_b200.head import DiffusionTransitionHead, HeadConfig

    S, H, NL = inp.x0.shape[1], inp.w_hh[0].shape[1], len(inp.w_hh)
    head = DiffusionTransitionHead(S, inp.context_full.shape[2], inp.theta.shape[1], HeadConfig(hidden_dim=H, num_layers=NL))
    with torch.no_grad():
        for k in range(NL):
            getattr(head.gru, f"weight_ih_l{k}").copy_(inp.w_ih[k])
            getattr(head.gru, f"weight_hh_l{k}").copy_(inp.w_hh[k])
            getattr(head.gru, f"bias_ih_l{k}").copy_(inp.b_ih[k])
            getattr(head.gru, f"bias_hh_l{k}").copy_(inp.b_hh[k])
        head.out_proj.weight.copy_(inp.out_w)
        head.out_proj.bias.copy_(inp.out_b)
    return head.cuda().train()


@pytest.mark.parametrize("kind,B,T", [("lv", 16, 40), ("ou", 300, 12)])
def test_path_iteration_matches_module_path_and_trains(kind, B, T):
    """bench.py's device-resident iteration (runner.PathIteration: C ABI, static buffers, CUDA graph) against the
    torch.library / nn.Module path on the same synthetic inputs; then three complete head training steps (graph replay +
    fused clip + AdamW + EMA on the flat buffers) against torch.optim.AdamW + clip_grad_norm_ + EMA lerp on the module."""
    from viforsdes_b200.runner import PathIteration
    from viforsdes_b200.synthetic import make_inputs
    from tests._util import head_grads

    inp = make_inputs(kind, B, T, context_dim=128, hidden_dim=64, num_layers=2, seed=5)
    it = PathIteration(inp, "cuda")
    it.step()
    head = _head_from_inputs(inp)
    paths, terms, gx0, gctx, gth, _ = _module_iteration(inp, head)
    r = it.results()
    assert_close(r["paths"], paths, rtol=1e-6, name="paths")
    assert_close(r["terms"], terms, rtol=1e-5, name="terms")
    assert_close(r["grads"]["x0"], gx0, rtol=1e-5, name="g_x0")
    assert_close(r["grads"]["context"], gctx[:, :T], rtol=1e-5, name="g_ctx")
    assert_close(r["grads"]["theta"], gth, rtol=1e-5, name="g_theta")
    for nm, g in head_grads(head).items():
        assert_close(r["grads"][nm], g, rtol=1e-5, name=f"g_{nm}")
    # --- three training steps of the head
    lr, max_norm, decay = 2e-3, 1.0, 0.9
    opt = it.make_optimizer(lr=lr, max_norm=max_norm, ema_decay=decay)
    it.capture()
    ref_opt = torch.optim.AdamW(head.parameters(), lr=lr)
    shadow = [p.detach().clone() for p in head.parameters()]
    losses = []
    for _ in range(3):
        it.replay()
        opt.step()
        ref_opt.zero_grad()
        *_, loss = _module_iteration(inp, head)
        losses.append(loss.item())
        torch.nn.utils.clip_grad_norm_(head.parameters(), max_norm)
        ref_opt.step()
        with torch.no_grad():
            for s_, p_ in zip(shadow, head.parameters()):
                s_.lerp_(p_.detach(), 1 - decay)
    torch.cuda.synchronize()
    names = {f"w_ih_l{k}": it.w[0][k] for k in range(2)} | {f"w_hh_l{k}": it.w[1][k] for k in range(2)} | \
            {f"b_ih_l{k}": it.w[2][k] for k in range(2)} | {f"b_hh_l{k}": it.w[3][k] for k in range(2)} | \
            {"out_w": it.out_w, "out_b": it.out_b}
    mod = {f"w_ih_l{k}": getattr(head.gru, f"weight_ih_l{k}") for k in range(2)} | \
          {f"w_hh_l{k}": getattr(head.gru, f"weight_hh_l{k}") for k in range(2)} | \
          {f"b_ih_l{k}": getattr(head.gru, f"bias_ih_l{k}") for k in range(2)} | \
          {f"b_hh_l{k}": getattr(head.gru, f"bias_hh_l{k}") for k in range(2)} | \
          {"out_w": head.out_proj.weight, "out_b": head.out_proj.bias}
    for nm in names:
        assert_close(names[nm], mod[nm], rtol=2e-4, atol_scale=2e-5, name=f"trained {nm}")
    assert len(opt.ema_views()) == 10


def test_fused_optimizer_skips_non_finite_steps_like_gradscaler():
    """trainer.py:199-203: scaler.unscale_ -> clip_grad_norm_ -> scaler.step(optimizer) skips optimizer.step() when any
    gradient is inf / NaN; ema.update() (trainer.py:126) still runs.  Reference = torch.optim.AdamW + torch.amp.GradScaler
    semantics restated with plain PyTorch on the same gradients."""
    from viforsdes_b200.optim import FlatParameters, FusedAdamWEma

    gen = torch.Generator().manual_seed(5)
    n, lr, decay, scale = 1037, 3e-3, 0.9, 256.0
    p0 = torch.randn(n, generator=gen)
    grads = [scale * torch.randn(n, generator=gen) for _ in range(5)]
    grads[1][17] = float("inf")
    grads[3][n - 1] = float("nan")
    # reference
    ref = nn.Parameter(p0.clone().double())
    ropt = torch.optim.AdamW([ref], lr=lr)
    shadow = ref.detach().clone()
    for g in grads:
        ref.grad = g.double() / scale
        if torch.isfinite(ref.grad).all():
            torch.nn.utils.clip_grad_norm_([ref], 1.0)
            ropt.step()
        shadow.lerp_(ref.detach(), 1 - decay)
    # fused
    param = nn.Parameter(p0.clone().cuda())
    flat = FlatParameters([[param]])
    opt = FusedAdamWEma(flat, lrs=[lr], max_norm=1.0, ema_decay=decay)
    inv = torch.tensor([1.0 / scale], device="cuda")
    seen = []
    for g in grads:
        param.grad = None  # what optimizer.zero_grad(set_to_none=True) does: the step re-attaches the flat view
        opt._check_views()
        param.grad.copy_(g)
        before = param.detach().clone()
        opt.step(inv_scale=inv)
        seen.append(bool(opt.found_inf.item()))
        if seen[-1]:
            assert torch.equal(param.detach(), before), "a skipped step must leave the parameters untouched"
    assert seen == [False, True, False, True, False]
    assert int(opt.skipped_steps.item()) == 2
    assert torch.isfinite(param).all() and torch.isfinite(opt.exp_avg).all() and torch.isfinite(opt.ema).all()
    assert_close(param, ref.detach(), rtol=2e-5, atol_scale=2e-6, name="param after skipped steps")
    assert_close(opt.ema_views()[0], shadow, rtol=2e-5, atol_scale=2e-6, name="ema after skipped steps")
    assert opt.state_dict()["step"] == 3
    with pytest.raises(ValueError):
        opt.load_state_dict({"exp_avg": opt.exp_avg, "exp_avg_sq": opt.exp_avg_sq, "ema": None, "step": 1})


def test_fused_optimizer_ema_apply_and_resume():
    """EMA swap for sampling (exponential_moving_average.py:30-42) and checkpoint / resume of the fused optimiser."""
    from viforsdes_b200.optim import FlatParameters, FusedAdamWEma

    gen = torch.Generator().manual_seed(0)
    lin = nn.Linear(6, 5).cuda()
    flat = FlatParameters([list(lin.parameters())])
    opt = FusedAdamWEma(flat, lrs=[1e-2], max_norm=1.0, ema_decay=0.9)
    grads = [torch.randn(flat.grads.numel(), generator=gen).cuda() for _ in range(4)]
    for g in grads[:2]:
        flat.grads.copy_(g)
        opt.step()
    w_now = lin.weight.detach().clone()
    with opt.ema_applied():
        assert torch.equal(lin.weight, opt.ema_views()[0]) and not torch.equal(lin.weight, w_now)
    assert torch.equal(lin.weight, w_now)
    # resume: a second optimiser on a copy of the parameters continues bit-identically
    lin2 = nn.Linear(6, 5).cuda()
    lin2.load_state_dict(lin.state_dict())
    flat2 = FlatParameters([list(lin2.parameters())])
    opt2 = FusedAdamWEma(flat2, lrs=[1e-2], max_norm=1.0, ema_decay=0.9)
    opt2.load_state_dict(opt.state_dict())
    for g in grads[2:]:
        for f, o in ((flat, opt), (flat2, opt2)):
            f.grads.copy_(g)
            o.step()
    # (compare the tensors, not the raw buffers: the alignment padding between tensors carries whatever the random test
    # gradients put there and is not part of the checkpoint)
    for a, b2 in zip(lin.parameters(), lin2.parameters()):
        assert torch.equal(a, b2)
    for a, b2 in zip(opt.ema_views(), opt2.ema_views()):
        assert torch.equal(a, b2)
    assert opt2.step_count == 4
