"""dirtorch.utils.common (reference: dirtorch/utils/common.py) on the B200 path."""
from dirb200.common import (load_checkpoint, matmul, model_size, pool, save_checkpoint, switch_model_to_cuda,  # noqa: F401
                            tonumpy, torch_set_gpu, torch_set_seed, transform, typename, variables, whiten_features)
