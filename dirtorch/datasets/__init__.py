"""dirtorch.datasets (reference: dirtorch/datasets/{create,generic,oxford,paris}.py)."""
from dirb200.datasets import (Dataset, ImageList, ImageListLabels, ImageListLabelsQ, ImageListRelevants,  # noqa: F401
                              ImageListROIs, Oxford5K, Paris6K, ROxford5K, RParis6K, create)
