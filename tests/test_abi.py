"""CPU-side checks of the C ABI library: it loads, exports every symbol include/twilight_b200.h declares, and refuses to
run without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "twilight_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(twl_[a-z_0-9]+)\s*\(", text))
    assert "twl_align_profiles" in names
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from twilight_b200 import _lib, build
    build.build()
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in twilight_b200.h but not exported"
    # and the Python binding table covers the header one to one
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_device_means_hard_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import twilight_b200
    with pytest.raises(twilight_b200.TwilightError) as err:
        twilight_b200.Context()
    assert "no CPU fallback" in str(err.value) or "NO_DEVICE" in str(err.value)


def test_product_package_does_not_touch_the_oracle():
    """Nothing under twilight_b200/ may import, include, link or execute oracle/ (it is checker-only infrastructure;
    mentioning it in a comment is fine)."""
    pkg = os.path.join(ROOT, "twilight_b200")
    bad = re.compile(r'(^\s*(import|from)\s+[\w.]*oracle)|(#\s*include\s*[<"][^>"]*oracle)|(CDLL\([^)]*oracle)|(-l\s*twl_oracle)|(libtalco_ref)|(oracle_lib)', re.M)
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                hit = bad.search(text)
                assert hit is None, f"{os.path.join(base, f)}: {hit.group(0)}"


def test_numpy_struct_mirrors_match_ctypes():
    """The Python mirror assembles twl_level_pair / twl_level_result blocks as numpy structured arrays: their layout must
    equal the ctypes (= C header) structs field by field."""
    import ctypes as C
    from twilight_b200 import _lib, api
    for dt, st in ((api.Context._SIDE_DT, _lib.NodeSide), (api.Context._PAIR_DT, _lib.LevelPair), (api.Context._RES_DT, _lib.LevelResult)):
        assert dt.itemsize == C.sizeof(st)
        for name, _ in st._fields_:
            assert dt.fields[name][1] == getattr(st, name).offset, name
