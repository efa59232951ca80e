"""CPU-side checks of the boundary: the shared library is present, loads, and exports every
symbol include/surreal_b200.h declares.  No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'surreal_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(sb200_[a-z0-9_]+)\s*\(', txt)))


def test_library_builds_loads_and_exports_header_symbols():
    from surreal_b200 import build, _lib
    build.build()
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 5
    for s in syms:
        assert hasattr(L, s), 'header declares %s but the library does not export it' % s
    declared = set(_lib.exported_symbols())
    assert set(syms) == declared, 'ctypes signatures out of sync with the header: %s' % (set(syms) ^ declared)
    assert L.sb200_version() >= 100
    assert L.sb200_status_string(-1) == b'invalid argument'


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from surreal_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.SB200Error):
        _lib.lib()
