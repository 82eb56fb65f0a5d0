"""One e2e session iteration of the bench default, for `ncu --metrics gpu__time_duration.sum` (kernel list of the host-buffer path)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from viforsdes_b200.session import HostSession
from viforsdes_b200.synthetic import make_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
inp = make_inputs("l96", B, 100, context_dim=256, hidden_dim=64, num_layers=2)
sess = HostSession.from_inputs(inp, context_dtype=torch.bfloat16, device_noise_seed=7)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 7):
    sess.step()
sess.close()
