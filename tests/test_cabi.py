"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol the header declares."""
import ctypes
import subprocess


def test_header_symbols_exported(pkg):
    L = pkg._lib.lib()
    decl = pkg._lib.parse_header()
    names = [n for n, _, _ in decl]
    assert len(names) >= 35 and len(set(names)) == len(names)
    out = subprocess.run(['nm', '-D', '--defined-only', str(pkg._lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l}
    missing = [n for n in names if n not in exported]
    assert not missing, f'declared in include/sessrec_b200.h but not exported: {missing}'
    assert set(L.functions) == set(names)


def test_sm100a_sass_present(pkg):
    out = subprocess.run(['cuobjdump', '-lelf', str(pkg._lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out[:400]


def test_error_channel_without_gpu(pkg):
    L = pkg._lib.lib()
    # split-K without accumulate is rejected before any CUDA call
    r = L.functions['srk_gemm'](4, 4, 4, None, 4, 1, None, 4, 1, None, 4, None, None, None, None, ctypes.c_float(1.0), 0, 2,
                                None)
    assert r == -1 and 'split' in L.last_error()
    r = L.functions['srk_embed_gather_fwd'](None, None, 4, 6, 0, None, None, None, None, None)
    assert r == -3 and 'unsupported' in L.last_error()
    assert L.functions['srk_version']() >= 100


def test_missing_library_fails_loudly(pkg, tmp_path, monkeypatch):
    import pytest
    monkeypatch.setattr(pkg._lib, 'LIB_PATH', tmp_path / 'nope.so')
    with pytest.raises(pkg._lib.SessRecError, match='no CPU'):
        pkg._lib._Lib()


def test_cpu_tensors_are_rejected(pkg):
    import pytest
    import torch
    from sessionrec_pytorch_b200.srgnn import SRGNN
    m = SRGNN(16, 8, 1)
    b = pkg.SessionBatch.build([[1, 2, 3]], [4], 'session')
    with pytest.raises(pkg._lib.SessRecError, match='no CPU fallback'):
        m(b)
