"""dirtorch.nets: model registry and factory (reference: dirtorch/nets/__init__.py:18-95)."""
from dirb200.nets import (ResNetRMAC, create_model, load_pretrained_weights, model_names,  # noqa: F401
                          resnet50_rmac, resnet101_rmac, resnet152_rmac)
