"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/femocs_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from femocs_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "femocs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(built)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/femocs_b200.h is not exported" % n


def test_python_binding_covers_header(built):
    from femocs_b200 import lib
    assert sorted(lib.SIGNATURES) == _declared_symbols()
    lib.load()


def test_no_cpu_fallback(built):
    """Without a CUDA device fb_create must fail with a reason (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import femocs_b200 as fb
    with pytest.raises(fb.FemocsB200Error, match="no CUDA device|CPU fallback|driver"):
        fb.Context(0)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under femocs_b200/ or include/ may reference it."""
    bad = []
    for base in ("femocs_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h")):
                    src = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"\boracle\b|femocs_oracle|libfemocs_ref|/root/reference", src):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
