"""dirtorch.test_dir (reference: dirtorch/test_dir.py): evaluation driver on the B200 path."""
from dirb200.pipeline import (eval_model, expand_descriptors, extract_image_features, load_model,  # noqa: F401
                              test_dir_main)

if __name__ == "__main__":
    test_dir_main()
