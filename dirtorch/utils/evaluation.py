"""dirtorch.utils.evaluation.compute_average_precision (reference: dirtorch/utils/evaluation.py:46-82)."""
from dirb200.datasets import compute_average_precision  # noqa: F401
