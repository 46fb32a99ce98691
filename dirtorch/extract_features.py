"""dirtorch.extract_features (reference: dirtorch/extract_features.py): feature extraction driver on the B200 path."""
from dirb200.pipeline import extract_features, extract_features_main, load_model  # noqa: F401

if __name__ == "__main__":
    extract_features_main()
