"""Second-order variant of the checkpoint entry point (seistorch/checkpoint_new.py:220);
same implementation as seistorch_b200.checkpoint."""
from .checkpoint import checkpoint  # noqa: F401
