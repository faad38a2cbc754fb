"""CPU checks of the drop-in boundary: the library builds for sm_100a, loads, and exports every
symbol include/facet_b200.h declares (no compute calls — there is no GPU here)."""
import os
import re

from facet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "facet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert syms, "no symbols parsed from the header"
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/facet_b200.h but not exported"
    assert sorted(_lib.declared_symbols()) == syms, "ctypes signatures out of sync with the header"
    assert lib.fb_abi_version() == 1


def test_contract_violation_reports_error():
    lib = _lib.load()
    # null pointers are rejected before any CUDA call is made
    rc = lib.fb_tech_stats(None, 1, 8, 8, 192, 0, None, None, None, 0, None)
    assert rc < 0
    assert b"null" in lib.fb_last_error()


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
