"""dirtorch.datasets (reference: dirtorch/datasets/{create,generic,oxford,paris}.py)."""
from dirb200.datasets import (Dataset, ImageList, ImageListRelevants, ImageListROIs, Oxford5K, Paris6K,  # noqa: F401
                              ROxford5K, RParis6K, create)
