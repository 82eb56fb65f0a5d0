"""Time one ELBO iteration of a user-defined SDE (BASELINE config 5 shape: stochastic Lorenz-96, S = 10) through
the public modules: head.sample_diffusion_paths -> PyTorch drift/diffusion -> path_elbo_terms -> backward.
    python tools/time_generic_sde.py [B] [T] [S]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from oracle import oracle_torch as O
from tests._util import build_head, cuda_inputs, cuda_sde
from viforsdes_b200 import _lib
from viforsdes_b200.elbo import path_elbo_terms
from viforsdes_b200.observations import GaussianObservationLikelihood, Observations
from viforsdes_b200.state_space import StateSpace
from viforsdes_b200.types import DiffusionPathSample

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
p = O.make_problem("l96", B, T, context_dim=256, hidden_dim=64, num_layers=2, state_dim=S)
head = build_head(p)
x0, full, view, theta, eps = cuda_inputs(p)
obs = Observations(times=p.obs_times, values=p.obs_values)
lik = GaussianObservationLikelihood(variance=p.obs_variance)
sde = cuda_sde(p)
space = StateSpace(p.weights.state_dim, list(p.positive_dims))


def step():
    for q in (x0, full, theta, *head.parameters()):
        q.grad = None
    paths, means, chol = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    terms = path_elbo_terms(sde, obs, lik, theta, DiffusionPathSample(paths, means, chol, space), p.dt)
    (-(terms[:, 0] + terms[:, 1] - terms[:, 2] + terms[:, 3]).mean()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
lib = _lib.load()
n = 10
_lib.check(lib.visde_profile_begin(n * 16))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = (C.c_double * len(_lib.STAGES))()
cnt = (C.c_int * len(_lib.STAGES))()
_lib.check(lib.visde_profile_end(ms, cnt))
dt = e0.elapsed_time(e1) / n
print(f"l96 S={S} B={B} T={T}: {dt:.2f} ms per iteration (device time incl. PyTorch drift/diffusion + autograd), "
      f"{B * T / dt * 1e3:.3e} traj-steps/s")
print({s: round(ms[i] / n, 3) for i, s in enumerate(_lib.STAGES)})
