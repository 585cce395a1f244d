"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol that
include/ngp_b200.h declares, and validates arguments without touching a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from taichi_nerfs_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 22
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ngp_b200.h but not exported"


def test_python_signatures_cover_header(lib):
    from taichi_nerfs_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared_symbols()


def test_version_and_error_string(lib):
    assert lib.ngp_version() >= 100
    assert isinstance(lib.ngp_last_error(), bytes)
    assert lib.ngp_launch_count() >= 0


def test_argument_validation_needs_no_gpu(lib):
    """Negative sizes / null pointers / bad dtypes are rejected (rc < 0) before any launch."""
    from taichi_nerfs_b200.layout import make_hash_layout
    lay = make_hash_layout(2 ** 19, 16, 16, 1024, 2).as_ctypes()
    assert lib.ngp_ray_aabb_intersect(None, None, 0.5, None, -1, None) < 0
    assert lib.ngp_ray_aabb_intersect(None, None, 0.5, None, 8, None) < 0
    assert b"null" in lib.ngp_last_error()
    assert lib.ngp_hash_encode_fwd(None, None, C.byref(lay), None, 7, 8, None) < 0
    assert b"dtype" in lib.ngp_last_error()
    bad = make_hash_layout(2 ** 19, 16, 16, 1024, 2).as_ctypes()
    bad.feat_dim = 9
    assert lib.ngp_hash_encode_fwd(None, None, C.byref(bad), None, 0, 8, None) < 0
    assert lib.ngp_adam_step(None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-15, 1.0, 0, 0, 8, None) < 0
    # empty inputs are a no-op success
    assert lib.ngp_ray_aabb_intersect(None, None, 0.5, None, 0, None) == 0
    assert lib.ngp_hash_encode_fwd(None, None, C.byref(lay), None, 1, 0, None) == 0
    assert lib.ngp_dir_encode(None, None, 0, None) == 0


def test_product_path_has_no_cpu_fallback():
    """ops refuse CPU tensors instead of silently computing on the host."""
    import torch
    from taichi_nerfs_b200 import ops, _lib
    with pytest.raises(_lib.NgpError):
        ops.ray_aabb_intersect(torch.zeros(4, 3), torch.ones(4, 3), 0.5)


def test_product_code_never_imports_oracle():
    for sub in ("taichi_nerfs_b200", "modules", "datasets"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), (dirpath, f)
