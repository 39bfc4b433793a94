"""C-ABI checks that need no GPU: libcsd_b200.so loads, exports every symbol include/csd_b200.h declares,
the ctypes prototype table (_lib._PROTOTYPES) matches the header in both directions, host-only entry points
work, and the product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from conditional_score_diffusion_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "csd_b200.h")


def _header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_declares_entry_points():
    syms = _header_symbols()
    assert "csd_upfirdn2d_f32" in syms and "csd_conv_gemm" in syms and "csd_langevin_update_f32" in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol(lib):
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/csd_b200.h but not exported by libcsd_b200.so"


def test_ctypes_table_matches_header():
    assert sorted(_lib.exported_names()) == _header_symbols()


def test_struct_layout_matches_header():
    """Field order of csd_conv_gemm_desc in the header == ctypes mirror (a silent mismatch corrupts launches)."""
    text = open(HEADER).read()
    body = re.search(r"typedef struct csd_conv_gemm_desc \{(.*?)\} csd_conv_gemm_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?[A-Za-z_0-9]+\s*\*?\s*", "", decl, count=1)
        for part in decl.split(","):
            part = part.strip().lstrip("*").strip()
            part = re.sub(r"\[.*\]", "", part)
            if part:
                names.append(part)
    assert names == [f[0] for f in _lib.ConvGemmDesc._fields_]


def test_host_only_entry_points(lib):
    assert lib.csd_abi_version() == 1
    oh, ow = ctypes.c_int(0), ctypes.c_int(0)
    # up=2, 4x4 filter, pad (2,1): 80 -> 160 (op/upfirdn2d.py:104-105)
    assert lib.csd_upfirdn2d_out_size(80, 80, 4, 4, 2, 2, 1, 1, 2, 1, 2, 1, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (160, 160)
    # down=2, 4x4 filter, pad (1,1): 160 -> 80
    assert lib.csd_upfirdn2d_out_size(160, 160, 4, 4, 1, 1, 2, 2, 1, 1, 1, 1, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (80, 80)
    # invalid arguments give a status and a message, not a crash
    assert lib.csd_upfirdn2d_out_size(4, 4, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, ctypes.byref(oh), ctypes.byref(ow)) != 0
    assert lib.csd_last_error()


def test_gn_fused_planning_is_host_side(lib):
    """The one-launch GroupNorm takes the <= 20 px shapes of NCSN++ (nf 96) and declines large images."""
    for c0, c1, hw in ((288, 0, 25), (288, 288, 25), (288, 0, 100), (288, 192, 100), (192, 0, 400), (288, 192, 400),
                       (192, 192, 400), (128, 0, 64)):
        groups = min((c0 + c1) // 4, 32)
        assert lib.csd_gn_fused_supported(c0, c1, hw, groups, 64) == 1, (c0, c1, hw)
    assert lib.csd_gn_fused_supported(96, 0, 160 * 160, 24, 64) == 0      # large image: statistics ride in the conv epilogue
    assert lib.csd_gn_fused_supported(100, 0, 25, 25, 64) == 0            # channels not a multiple of 8
    assert lib.csd_gn_fused_supported(96, 0, 25, 7, 64) == 0              # groups do not divide the channels


def test_descriptor_validation_is_host_side(lib):
    """A malformed descriptor is rejected before any CUDA call (so this runs without a GPU)."""
    d = _lib.ConvGemmDesc()
    assert lib.csd_conv_gemm(ctypes.byref(d), None) != 0
    assert b"nseg" in lib.csd_last_error()


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors / without the extension (no oracle, no eager torch)."""
    from conditional_score_diffusion_b200 import kernels as K
    from conditional_score_diffusion_b200 import op
    x = torch.zeros(1, 1, 4, 4)
    k = torch.ones(4, 4)
    with pytest.raises(Exception):
        op.upfirdn2d(x, k, up=2, pad=(2, 1))
    with pytest.raises(_lib.CsdError):
        K.upfirdn2d_planes(x.view(1, 4, 4), k, 1, 1, 1, 1, 0, 0, 0, 0)


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "conditional_score_diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
