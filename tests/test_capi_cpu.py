"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/svo_b200.h
declares (no compute calls here: there is no GPU in this container); the product refuses to run
without a CUDA device instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "svo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svo_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from rpg_svo_b200 import build, capi

    path = build.build()
    assert os.path.exists(path)
    lib = capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svo_b200.h but not exported"


def test_sass_is_sm100a_with_tma_and_no_cpu_fallback():
    from rpg_svo_b200 import build

    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", build.OUT], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", build.OUT], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "TMA bulk copy (cp.async.bulk) missing from the kernels"
    assert "SYNCS" in sass, "mbarrier instructions missing"


def test_context_creation_fails_loudly_without_gpu():
    import torch

    from rpg_svo_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SvoB200Error):
        capi.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rpg_svo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", "").replace("oracle/", "") \
                    or "import oracle" not in txt and "from oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt and "libsvo_oracle" not in txt, f


def test_public_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/svo_b200.h must compile as C99 with no C++ / torch types."""
    import subprocess

    src = tmp_path / "c.c"
    src.write_text('#include "svo_b200.h"\nint main(void){ svo_b200_map_view v; svo_b200_detect_options o; '
                   'svo_b200_sia_options s; (void)v; (void)o; (void)s; return 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(root, "include"), str(src)])
