"""CPU: the C-ABI library builds, loads and exports every symbol include/ramnet_b200.h declares.
No compute call is made (no GPU here)."""
import ctypes
import os
import re

from rpg_ramnet_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ramnet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ramnet_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/ramnet_b200.h but not exported'
    assert set(syms) == set(_lib.SIGNATURES), 'ctypes table and header disagree'


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, 'include', 'ramnet_b200.h')).read()
    assert lib.ramnet_version() == int(re.search(r'#define RAMNET_ABI_VERSION (\d+)', hdr).group(1))
    assert isinstance(lib.ramnet_last_error(), bytes)


def test_handle_raises_instead_of_hanging_without_a_gpu():
    """handle() used to call check() under a non-reentrant lock and deadlock when ramnet_create failed."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('needs a box without a GPU')
    code = ('from rpg_ramnet_b200 import _lib\n'
            'try:\n    _lib.handle(0)\nexcept _lib.RamnetError as e:\n    print("RAISED", e)\n')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert 'RAISED' in r.stdout, r.stdout + r.stderr


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-lelf', build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs
