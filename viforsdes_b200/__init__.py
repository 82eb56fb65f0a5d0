"""B200-native variational path sampling for SDEs (drop-in for the hot path of Tom-Ryder/VIforSDEs)."""
from viforsdes_b200 import _lib  # noqa: F401

__all__ = ["_lib"]
