"""crossscore_b200: B200-native (sm_100a) implementation of the CrossScore inference hot path."""
from .config import default_cfg  # noqa: F401
from .model import CrossScoreNet, load_checkpoint  # noqa: F401

__all__ = ["CrossScoreNet", "load_checkpoint", "default_cfg"]
