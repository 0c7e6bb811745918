"""The C-ABI library loads on a machine without a GPU, exports every symbol include/misa_b200.h declares, and
refuses to compute without a CUDA device (no CPU fallback). No compute call is made here."""
import ctypes as C
import os
import re

import pytest

import misa_md_b200 as mb
from misa_md_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "misa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(misa_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = mb.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "misa_b200.h")).read()
    for needle in ("arch_imp.h:17-31", "atom.cpp:160,294,321", "simulation.cpp:164-194", "neighbour_index.inl:13-76",
                   "lat_particle_packer.cpp", "atom_element.h:18-41"):
        assert needle in text, needle


def test_compute_fails_loudly_without_a_device():
    lib = mb.load()
    if lib.misa_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert lib.misa_b200_env_init(-1) == -1  # MISA_B200_ENODEV
    assert b"no CUDA device" in lib.misa_b200_last_error()
    dom = capi.make_domain((8, 8, 8))
    h = C.c_void_p()
    assert lib.misa_b200_create(C.byref(dom), C.byref(h)) == -1
    with pytest.raises(mb.MisaError):
        mb.Context((8, 8, 8))


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi._build, "LIB", str(tmp_path / "libmisa_b200.so"))
    monkeypatch.setattr(capi._build, "build", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("nvcc absent")))
    with pytest.raises((mb.MisaError, RuntimeError)):
        capi.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under misa_md_b200/ may import, link or dlopen it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "misa_md_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in text.splitlines() if not l.strip().startswith(("//", "#", "*", "/*")))
                assert "liboracle" not in code and "from oracle" not in code and "import oracle" not in code, f
