"""dirtorch.utils.pytorch_loader.get_loader (reference: dirtorch/utils/pytorch_loader.py:11-73)."""
from dirb200.loader import get_loader  # noqa: F401
