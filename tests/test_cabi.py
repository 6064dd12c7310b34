"""CPU suite: the C-ABI library loads and exports every symbol include/vitb200.h declares (no compute calls)."""

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "vitb200.h"
LIB = ROOT / "vit_plasticity_b200" / "libvitb200.so"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    if not LIB.exists():
        import __graft_entry__ as g

        g.build()
    return LIB


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("vb_gemm_bf16", "vb_layernorm_fwd", "vb_layernorm_bwd", "vb_attention_fwd", "vb_attention_bwd", "vb_attention_pair_delta",
                 "vb_im2col_patches", "vb_rowsumsq_diff_f32", "vb_layernorm_pair_sqdiff", "vb_last_error", "vb_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in vitb200.h but not exported: {missing}"


def test_ctypes_binding_matches_header(built):
    from vit_plasticity_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    handle = _lib.lib()
    assert handle.vb_version() >= 100
    assert isinstance(handle.vb_last_error(), bytes)
    # struct layout: the ctypes mirror must agree with what the C compiler makes of include/vitb200.h
    import subprocess
    import tempfile

    src = '#include <stddef.h>\n#include <stdio.h>\n#include "vitb200.h"\nint main(void){printf("%zu %zu %zu %zu", sizeof(vb_gemm_args), offsetof(vb_gemm_args, bias), offsetof(vb_gemm_args, split_k), offsetof(vb_gemm_args, out_colsum));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = Path(d) / "layout.c", Path(d) / "layout"
        cfile.write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(cfile), "-o", str(exe)], check=True)
        size, off_bias, off_split, off_cs = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_lib.GemmArgs) == size
    assert (_lib.GemmArgs.bias.offset, _lib.GemmArgs.split_k.offset, _lib.GemmArgs.out_colsum.offset) == (off_bias, off_split, off_cs)


def test_argument_validation_needs_no_gpu(built):
    from vit_plasticity_b200 import _lib

    handle = _lib.lib()
    args = _lib.GemmArgs()
    assert handle.vb_gemm_bf16(ctypes.byref(args), None) == 1  # VB_ERR_INVALID: bad shape, before any CUDA call
    assert b"bad shape" in handle.vb_last_error()
    assert handle.vb_layernorm_fwd(None, None, None, None, None, None, 4, 768, 1e-12, None) == 1
    assert handle.vb_attention_fwd(None, None, None, 1, 197, 12, 64, None) == 1


def test_sass_is_blackwell_native(built):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG/UTMAREDG (B200_PROFILING.md)."""
    obj = ROOT / "vit_plasticity_b200" / "csrc" / "gemm_tcgen05.o"
    if not obj.exists():
        pytest.skip("object file not kept")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
