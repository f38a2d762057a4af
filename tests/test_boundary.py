"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol of include/ssm.h,
the ctypes struct mirrors the C struct, and the product package never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from semantic_slam_mapping_b200 import build, lib as L
    build.build()
    return L.load()


def test_header_symbols_are_exported(lib):
    from semantic_slam_mapping_b200.lib import SYMBOLS
    hdr = open(os.path.join(ROOT, "include", "ssm.h")).read()
    declared = sorted(set(re.findall(r"\b(ssm_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_params_struct_layout_matches_c(lib, tmp_path):
    from semantic_slam_mapping_b200.params import CParams, Params
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ssm.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(ssm_params),'
                   ' offsetof(ssm_params, cx), offsetof(ssm_params, palette_bgr), offsetof(ssm_params, map_capacity));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, off_cx, off_pal, off_cap = map(int, subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(CParams) == size
    assert CParams.cx.offset == off_cx and CParams.palette_bgr.offset == off_pal and CParams.map_capacity.offset == off_cap
    d = CParams()
    lib.ssm_default_params(ctypes.byref(d))
    p = Params().c()
    for name, _ in CParams._fields_:
        if name == "palette_bgr":
            assert bytes(d.palette_bgr) == bytes(p.palette_bgr)
        else:
            assert getattr(d, name) == getattr(p, name), name   # python defaults == C defaults == reference values


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semantic_slam_mapping_b200 import Context, SsmError
    with pytest.raises(SsmError) as e:
        Context()
    assert e.value.code == -3   # SSM_ERR_NO_DEVICE: there is no CPU fallback


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semantic_slam_mapping_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "ssm_oracle" not in text, f


def test_voxel_owner_is_a_partition(lib):
    from semantic_slam_mapping_b200 import voxel_owner
    for n in (1, 2, 4, 8):
        owners = {voxel_owner(i, j, k, n) for i in range(-40, 40, 3) for j in range(-9, 9, 2) for k in range(0, 300, 7)}
        assert owners <= set(range(n)) and (n == 1 or len(owners) == n)
    # all voxels of one 8^3 brick share an owner
    assert len({voxel_owner(8 + a, -16 + b, 24 + c, 8) for a in range(8) for b in range(8) for c in range(8)}) == 1
