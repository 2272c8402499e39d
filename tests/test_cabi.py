"""CPU: the C-ABI shared library loads and exports exactly the symbols include/vlg.h declares
(no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from vlg import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vlg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    lib = built.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vlg.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", built.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vlg_[a-z0-9_]+)", out)))
    assert exported == names, f"export list differs from the header: {set(exported) ^ set(names)}"


def test_binding_table_matches_header(built):
    assert sorted(built.SIGNATURES) == declared_symbols()
    assert built.lib().vlg_abi_version() == 1
    assert built.lib().vlg_launch_count() == 0


def test_struct_sizes_match_the_c_side(built, tmp_path):
    """Compile a tiny C program against vlg.h and compare sizeof() with the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "vlg.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(vlg_sample_params),sizeof(vlg_geometry),sizeof(vlg_layer_weights),sizeof(vlg_weights),'
                   'sizeof(vlg_generate_params),sizeof(vlg_diff_geometry),sizeof(vlg_diff_block),sizeof(vlg_diff_weights));'
                   'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    import ctypes as C
    assert sizes == [C.sizeof(built.SampleParams), C.sizeof(built.Geometry), C.sizeof(built.LayerWeights),
                     C.sizeof(built.Weights), C.sizeof(built.GenerateParams), C.sizeof(built.DiffGeometry),
                     C.sizeof(built.DiffBlock), C.sizeof(built.DiffWeights)]


def test_sass_is_blackwell_native():
    """The GEMM must be tcgen05 (UTCHMMA/LDTM) fed by TMA (UTMALDG), not a recompiled mma.sync kernel."""
    obj = os.path.join(ROOT, "video-llamagen_b200", "lib", "gemm.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    attn = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "video-llamagen_b200", "lib", "attn_tma.o")],
                          capture_output=True, text=True).stdout
    assert "UTMALDG" in attn
    tail = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "video-llamagen_b200", "lib", "layer_tail.o")],
                          capture_output=True, text=True).stdout
    assert "UTCHMMA" in tail and "UTMALDG" in tail and "LDTM" in tail and "HMMA." not in tail.replace("UTCHMMA", "")
