"""CPU: the C-ABI library builds, loads, and exports every function include/*.h declares; the ctypes
binding table covers the same set (no compute calls -- no GPU here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b((?:a3d|cd)_\w+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    from act3d_chained_diffuser_b200 import build, lib
    path = build.build()
    raw = ctypes.CDLL(path)
    declared = declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/ but not exported by {os.path.basename(path)}"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    loaded = lib.load()
    assert loaded.a3d_abi_version() == 1
    assert lib.xattn_layer_floats(60, 60) == 8 * 64 and lib.xattn_layer_words(60, 60) == 2 * 4 * 4 * 8 * 32 * 4
    assert lib.kv_bytes(2, 3, 65, 4) == 2 * 3 * 2 * 2 * 4 * 2048


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from act3d_chained_diffuser_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(lib.A3DError, match="no CPU/eager fallback"):
        lib.load()
