"""CPU-side checks of the C-ABI library: it loads, exports every symbol
include/mcgpu.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mcgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcgpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(mcgpu_lib):
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(mcgpu_lib, n)]
    assert not missing, "declared in mcgpu.h but not exported: %s" % missing


def test_python_binding_lists_the_same_symbols():
    from mcpar_b200 import engine
    assert sorted(engine.EXPORTS) == _declared()


def test_config_struct_layout_matches_header(mcgpu_lib):
    from mcpar_b200 import engine
    # 4*int32, 3*int64, 2*int32, 5*double, uint64, 4*int32, int64, 2*int32
    assert C.sizeof(engine.Config) == 16 + 24 + 8 + 40 + 8 + 16 + 8 + 8
    assert C.sizeof(engine.Stats) == 8 * 8 + 8 + 24
    hdr = open(os.path.join(ROOT, "include", "mcgpu.h")).read()
    assert "#define MCGPU_ABI_VERSION 2" in hdr and engine.Config().abi_version == 0


def test_no_cpu_fallback_without_gpu(mcgpu_lib):
    from mcpar_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.McgpuError, match="ENODEVICE"):
        engine.Engine(2, 64)
    with pytest.raises(engine.McgpuError, match="ENODEVICE"):
        engine.loglik("rosenbrock1", 2, [[1.0, 1.0]])


def test_product_does_not_reference_the_oracle():
    """Nothing under mcpar_b200/ or include/ may import, link or name oracle/."""
    bad = []
    for base in ("mcpar_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if "build" in dp or "__pycache__" in dp:
                continue
            for f in fn:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle/|oracle\.(mh|ref)\b|import oracle|from oracle|mh_oracle|libmcpar_ref", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
