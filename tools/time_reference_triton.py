"""Time the UNMODIFIED reference's own GPU path (its Triton kernels, kernels/forward.py + backward.py through
models/head.py:156-209) on this GPU, next to this library, on the same seeded inputs -- the "GPU bar" of SURVEY.md §8d.
Needs a copy of the reference sources under baseline/_ref/src (git-ignored; `cp -r /root/reference/src baseline/_ref/`
in the build container: it travels to the GPU box with the snapshot).  Not part of the product, tests or bench.py.

    python tools/time_reference_triton.py [kind] [B] [T]          (default lv 128 800)
"""
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
ref_src = ROOT / "baseline" / "_ref" / "src"
if not ref_src.exists():
    print(f"{ref_src} is missing: reference not available here")
    sys.exit(0)
sys.path.insert(0, str(ref_src))
for name, attrs in {"matplotlib": {}, "matplotlib.pyplot": {}, "matplotlib.axes": {"Axes": object},
                    "matplotlib.figure": {"Figure": object}}.items():
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)

import torch  # noqa: E402

from oracle import oracle_torch as O  # noqa: E402
from tests._util import build_head, cuda_inputs, normwise  # noqa: E402
from variational_sde.config import HeadConfig  # noqa: E402
from variational_sde.models.head import DiffusionTransitionHead as RefHead  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "lv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T = int(sys.argv[3]) if len(sys.argv) > 3 else 800
p = O.make_problem(kind, B, T, context_dim=256, hidden_dim=64, num_layers=2)
w = p.weights
ref = RefHead(w.state_dim, w.context_dim, w.param_dim, HeadConfig(hidden_dim=w.hidden_dim, num_layers=w.num_layers))
with torch.no_grad():
    for k in range(w.num_layers):
        getattr(ref.gru, f"weight_ih_l{k}").copy_(w.w_ih[k])
        getattr(ref.gru, f"weight_hh_l{k}").copy_(w.w_hh[k])
        getattr(ref.gru, f"bias_ih_l{k}").copy_(w.b_ih[k])
        getattr(ref.gru, f"bias_hh_l{k}").copy_(w.b_hh[k])
    ref.out_proj.weight.copy_(w.out_w)
    ref.out_proj.bias.copy_(w.out_b)
ref = ref.cuda().train()
ours = build_head(p)
g = torch.Generator(device="cuda").manual_seed(3)


def run(head, n, cts=None):
    x0, full, view, theta, eps = cuda_inputs(p)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    out = None
    for _ in range(n):
        for q in (x0, full, theta, *head.parameters()):
            q.grad = None
        ev[0].record()
        out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
        ev[1].record()
        c = cts or [torch.ones_like(o) for o in out]
        torch.autograd.backward(list(out), c)
        ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[1].elapsed_time(ev[2])
    return tf / n, tb / n, [o.detach() for o in out], full.grad.detach().clone()


x = run(ref, 1)  # Triton JIT
cts = [torch.randn(o.shape, device="cuda", generator=g) for o in x[2]]
run(ref, 2, cts)
rf, rb, ro, rg = run(ref, 5, cts)
run(ours, 3, cts)
of, ob, oo, og = run(ours, 5, cts)
print(f"{kind} B={B} T={T} C=256 H=64x2, fwd + bwd of sample_diffusion_paths (CUDA events, incl. op dispatch):")
print(f"  reference Triton kernels : fwd {rf:8.3f} ms  bwd {rb:8.3f} ms  total {rf + rb:8.3f} ms")
print(f"  this library (AUTO)      : fwd {of:8.3f} ms  bwd {ob:8.3f} ms  total {of + ob:8.3f} ms   speed-up {(rf + rb) / (of + ob):.1f}x")
for a, b, nm in zip(oo, ro, ("paths", "means", "chol")):
    print(f"  max normwise |ours - reference| {nm}: {normwise(a, b):.2e}")
print(f"  max normwise |ours - reference| grad_context: {normwise(og, rg):.2e}")
