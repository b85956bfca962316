"""The C-ABI library: loads without a GPU, exports every symbol include/b200ir.h declares, ctypes mirrors of the
structs have the C layout, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200ir.h')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from image_restoration_b200 import _lib
    return _lib.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200ir_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(lib):
    from image_restoration_b200 import _lib
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f'{n} declared in b200ir.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature'
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.b200ir_abi_version() == 3


def test_struct_layout_matches_c(tmp_path):
    from image_restoration_b200 import _lib
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200ir.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(b200ir_view), sizeof(b200ir_conv_desc),'
                   'offsetof(b200ir_conv_desc, weight), offsetof(b200ir_conv_desc, out),'
                   'offsetof(b200ir_conv_desc, res_scale));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    sv, sd, ow, oo, ors = map(int, subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(_lib.View) == sv and ctypes.sizeof(_lib.ConvDesc) == sd
    assert _lib.ConvDesc.weight.offset == ow and _lib.ConvDesc.out.offset == oo
    assert _lib.ConvDesc.res_scale.offset == ors


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.b200ir_device_check() != 0
    assert b'CUDA' in lib.b200ir_last_error() or b'sm_' in lib.b200ir_last_error()
    from image_restoration_b200 import _lib
    d = _lib.ConvDesc()
    assert lib.b200ir_conv_igemm(ctypes.byref(d), None) != 0
    assert lib.b200ir_launch_count() == 0


def test_sass_is_blackwell_native():
    """The built library contains tcgen05 MMA, TMEM loads and TMA loads (SASS mnemonics of B200_PROFILING.md)."""
    so = os.path.join(ROOT, 'image_restoration_b200', 'libb200ir.so')
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip('cuobjdump unavailable')
    # UTCHMMA.2CTA / UTCBAR.2CTA.MULTICAST: the CTA-pair conv kernel (tcgen05.mma.cta_group::2 + multicast commit)
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTCHMMA.2CTA', 'UTCBAR.2CTA.MULTICAST', 'UTMALDG.4D.2CTA'):
        assert mnemonic in sass, mnemonic
    assert 'HMMA.16816' not in sass          # no legacy mma.sync path
