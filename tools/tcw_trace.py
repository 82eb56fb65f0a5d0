"""Phase timing of one tile-step of the wide tensor-core recurrence kernels (clock64 stamps of CTA 0, steps 40..55).
Build first with the trace hooks:  make -C viforsdes_b200/csrc clean && make -C viforsdes_b200/csrc -j8 NVCCFLAGS_EXTRA=-DVISDE_TCW_TRACE
    python tools/tcw_trace.py"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from viforsdes_b200 import _lib
from viforsdes_b200.runner import PathIteration
from viforsdes_b200.synthetic import make_inputs

inp = make_inputs("l96", 8192, 100, context_dim=256, hidden_dim=64, num_layers=2)
it = PathIteration(inp, "cuda")
for _ in range(3):
    it.step()
torch.cuda.synchronize()
lib = C.CDLL(str(_lib.LIB_PATH))
GHZ = 1.965
for name, fn, labels in (
    ("forward", "visde_debug_tcw_trace_fwd", ["step start", "d0 ready", "L0 epilogue done", "a0 arrived+issue (warp0) / prefetch", "d1 ready",
                                                "L1 epilogue done", "a1 arrived+issue", "out ready", "out epilogue done"]),
    ("backward", "visde_debug_tcw_trace_bwd", ["step start (loads requested)", "previous step's MMAs done, d z_t read from TMEM", "d_out . W_out landed (k=1 pass-1 start)",
                                                 "k=1 pass 1 done (row scale agreed)", "k=1 pass 2 done (4 chunks issued)", "before in0 wait",
                                                 "in0 ready", "k=0 pass-1 start", "k=0 pass 1 done", "k=0 pass 2 done",
                                                 "   (d_out entries computed, row scale agreed)", "   (d_out operand written to the ring)"])):
    buf = (C.c_longlong * (2 * 16 * 16))()
    assert getattr(lib, fn)(buf) == 0
    for th, who in ((0, "thread 0 (warp 0, issuer)"), (1, "thread 224 (warp 7)")):
        print(f"== {name}, {who}: mean over steps 41..54, microseconds since step start")
        n_slot = len(labels)
        acc = [0.0] * n_slot
        cnt = 0
        total = 0.0
        for t in range(1, 15):
            base = buf[(th * 16 + t) * 16 + 0]
            nxt = buf[(th * 16 + t + (1 if name == "forward" else -1)) * 16 + 0]
            if base == 0 or nxt == 0:
                continue
            cnt += 1
            total += abs(nxt - base)
            for sl in range(n_slot):
                acc[sl] += buf[(th * 16 + t) * 16 + sl] - base
        if not cnt:
            print("   (no samples)")
            continue
        for sl, lab in enumerate(labels):
            if lab is None:
                continue
            print(f"   {acc[sl] / cnt / GHZ / 1e3:7.2f} us  {lab}")
        print(f"   {total / cnt / GHZ / 1e3:7.2f} us  = one full step")
