from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus) -> None:
    """Achieved parity errors of this session -> gpurun_out/parity_log.jsonl (GPU box runs; scratch, see tools/parity_table.py)."""
    import json
    import os

    from tests import _util

    if not _util.PARITY_LOG:
        return
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / os.environ.get("VISDE_PARITY_LOG", "parity_log.jsonl"), "a") as f:
            for rec in _util.PARITY_LOG:
                f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
