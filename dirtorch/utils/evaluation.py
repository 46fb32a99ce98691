"""dirtorch.utils.evaluation (reference: dirtorch/utils/evaluation.py:41-82): the two AP definitions of the path."""
from dirb200.datasets import compute_average_precision  # noqa: F401


def compute_AP(label, score):
    """Label-based AP (evaluation.py:41-43): scikit-learn's average_precision_score."""
    from sklearn.metrics import average_precision_score
    return average_precision_score(label, score)
