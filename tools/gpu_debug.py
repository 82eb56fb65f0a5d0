"""Ad-hoc GPU diagnostics (not part of the product): compare TC vs SIMT GEMM stages on one case."""
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import oracle_torch as O
from tests._util import build_head, cuda_inputs, head_grads
from viforsdes_b200 import ops, _lib

def run(p, variant, seed=7):
    ops.set_variant(variant)
    B, T, S = p.eps.shape
    g = torch.Generator().manual_seed(seed)
    gP, gM, gL = torch.randn(B, T + 1, S, generator=g), torch.randn(B, T, S, generator=g), torch.randn(B, T, S, S, generator=g)
    head = build_head(p)
    x0, full, view, th, eps = cuda_inputs(p)
    out = head.sample_diffusion_paths(x0, view, th, eps, p.dt)
    torch.autograd.backward(list(out), [gP.cuda(), gM.cuda(), gL.cuda()])
    torch.cuda.synchronize()
    return [o.detach().cpu() for o in out], {"x0": x0.grad.cpu(), "context": full.grad[:, :T].cpu(), "theta": th.grad.cpu(),
                                              **{k: v.cpu() for k, v in head_grads(head).items()}}

kind, B, T, kw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), eval(sys.argv[4])
p = O.make_problem(kind, B, T, **kw)
o_s, g_s = run(p, _lib.VARIANT_GENERIC | 0x100)
o_t, g_t = run(p, _lib.VARIANT_GENERIC)
for nm, a, b in zip(("paths", "means", "chol"), o_t, o_s):
    print(f"{nm:10s} max|tc-simt| {(a-b).abs().max():.3e}  scale {b.abs().max():.3e}")
for k in g_s:
    a, b = g_t[k], g_s[k]
    print(f"grad {k:10s} max|tc-simt| {(a-b).abs().max():.3e}  scale {b.abs().max():.3e}  tc_absmax {a.abs().max():.3e} nan {int(torch.isnan(a).sum())}")
a, b = g_t["context"], g_s["context"]
print("ctx grad tc[0,0,:8]  ", a[0, 0, :8])
print("ctx grad simt[0,0,:8]", b[0, 0, :8])
print("ctx grad tc[0,5,32:40]  ", a[0, 5, 32:40])
print("ctx grad simt[0,5,32:40]", b[0, 5, 32:40])
r = (a / b)
print("ratio quantiles", torch.quantile(r.flatten()[:100000], torch.tensor([0.01, 0.25, 0.5, 0.75, 0.99])))
a, b = g_t["w_ih_l0"], g_s["w_ih_l0"]
print("w_ih_l0 tc[0, S:S+6]", a[0, 2:8], " simt", b[0, 2:8])
print("w_ih_l0 tc[5, S+40:S+44]", a[5, 42:46], " simt", b[5, 42:46])
a, b = g_t["w_hh_l0"], g_s["w_hh_l0"]
print("w_hh_l0 tc[0,:4]", a[0, :4], " simt", b[0, :4]); print("w_hh_l0 tc[130,:4]", a[130, :4], " simt", b[130, :4])
