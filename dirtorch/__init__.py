"""Drop-in ``dirtorch`` surface: the module paths of naver/deep-image-retrieval that sit on the descriptor
extraction + retrieval hot path, re-exported from the B200 implementation (package ``deep-image-retrieval_b200``,
importable as ``dirb200``).  ``python -m dirtorch.test_dir`` / ``python -m dirtorch.extract_features`` keep the
reference's command-line flags."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import dirb200 as _dirb200  # noqa: E402,F401
