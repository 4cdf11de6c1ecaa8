"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every
function include/ifcb_b200.h declares; argument errors come back as status codes."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'ifcb_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ifcb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(built_lib):
    names = _declared()
    assert 'ifcb_preprocess' in names and 'ifcb_plan_add_conv' in names and len(names) >= 15
    for n in names:
        assert hasattr(built_lib, n), n
    from ifcb_classifier_b200 import _lib
    assert sorted(_lib._SIGNATURES) == names            # the ctypes table mirrors the header one to one


def test_library_is_sm100a_with_tcgen05_and_tma():
    from ifcb_classifier_b200 import _lib
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert 'UTCHMMA' in sass          # tcgen05.mma
    assert 'UTMALDG.4D.IM2COL' in sass and 'UTMALDG.2D' in sass   # TMA im2col + tiled
    assert 'LDTM' in sass             # tcgen05.ld
    assert 'HMMA.16816' not in sass   # no legacy mma.sync path


def test_argument_errors_are_status_codes(built_lib):
    from ifcb_classifier_b200 import _lib
    rc = built_lib.ifcb_preprocess(None, 0, None, None, None, 4, 10, 10, 299, None, None, 0, None, 0, None, None)
    assert rc < 0 and b'null' in built_lib.ifcb_last_error()
    rc = built_lib.ifcb_conv_geometry(64, 64, 3, 3, 24, None, None, None, None)
    assert rc < 0 and b'tile_n' in built_lib.ifcb_last_error()
    g = _lib.conv_geometry(80, 192, 3, 3)
    assert g == dict(Cin_pad=128, K_pad=1152, tile_n=192, Cout_pad=192)
    assert _lib.conv_geometry(32, 64, 3, 3) == dict(Cin_pad=32, K_pad=288, tile_n=64, Cout_pad=64)
    assert _lib.conv_geometry(48, 64, 5, 5)['Cin_pad'] == 64
    g = _lib.conv_geometry(2048, 1344, 1, 1)
    assert g['tile_n'] == 224 and g['Cout_pad'] == 1344
    assert built_lib.ifcb_plan_num_layers(None) == -1
