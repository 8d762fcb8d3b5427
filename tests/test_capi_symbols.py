"""The C-ABI library loads and exports every symbol include/oflib_b200.h declares (no compute calls: no GPU here),
and the product package never touches the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'oflib_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ofb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ('ofb_warp_t_fwd', 'ofb_warp_t_bwd', 'ofb_splat_s_fwd', 'ofb_splat_s_bwd', 'ofb_flow_stats',
                 'ofb_abi_version', 'ofb_error_string'):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from oflibpytorch_b200 import _native
    assert os.path.exists(_native.LIB_PATH), 'build the extension first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), 'symbol %s declared in include/oflib_b200.h is not exported' % name
    assert sorted(_native.SIGNATURES) == declared_functions(), 'ctypes signature table out of sync with the header'
    lib.ofb_abi_version.restype = ctypes.c_int
    assert lib.ofb_abi_version() == 1
    lib.ofb_error_string.restype = ctypes.c_char_p
    assert b'invalid' in lib.ofb_error_string(-1)


def test_missing_library_fails_loudly(monkeypatch):
    from oflibpytorch_b200 import _native
    import pytest
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', '/nonexistent/liboflib_b200.so')
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        _native.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'oflibpytorch_b200')
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                text = open(os.path.join(dirpath, fn)).read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(('import ', 'from ', '#include')):
                        assert 'oracle' not in s, '%s references the oracle: %s' % (fn, s)
                    assert 'liboracle' not in s and '/root/reference' not in s, '%s: %s' % (fn, s)


def test_header_is_valid_c99_and_links_from_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C (not only C++), and a plain C program must be able to
    link the library and call an entry point that needs no GPU."""
    import shutil
    import subprocess
    from oflibpytorch_b200 import _native
    gcc = shutil.which('gcc')
    if gcc is None:
        import pytest
        pytest.skip('no gcc')
    src = tmp_path / 'probe.c'
    src.write_text('#include <stdio.h>\n#include "oflib_b200.h"\n'
                   'int main(void) { printf("%d %s\\n", ofb_abi_version(), ofb_error_string(OFB_EINVAL)); return 0; }\n')
    exe = tmp_path / 'probe'
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-l:liboflib_b200.so', '-Wl,-rpath,' + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith('1 ') and 'invalid' in out


def test_torch_library_shim_registers_the_hot_ops():
    """csrc/torch_binding.cpp: the two hot W ops registered from C++ (TORCH_LIBRARY oflib_b200n), CUDA key only."""
    import pytest
    import torch
    from oflibpytorch_b200 import _ops
    assert _ops._NATIVE is not None, 'liboflib_b200_torch.so missing: run __graft_entry__.build()'
    for name in ('warp_t_fwd', 'warp_t_bwd'):
        assert hasattr(torch.ops.oflib_b200n, name)
    with pytest.raises(RuntimeError):        # no CPU kernel, no fallback: the package's own message comes first
        _ops.warp_t(torch.zeros(1, 2, 4, 4), torch.zeros(1, 1, 4, 4))
