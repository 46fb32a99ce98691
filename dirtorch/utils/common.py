"""dirtorch.utils.common (reference: dirtorch/utils/common.py) on the B200 path."""
from dirb200.common import (load_checkpoint, matmul, pool, switch_model_to_cuda, tonumpy, torch_set_gpu,  # noqa: F401
                            torch_set_seed, transform, typename, variables, whiten_features)
