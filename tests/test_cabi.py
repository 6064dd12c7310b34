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


def test_gemm_plan_tile_width_and_split_k_need_no_gpu(built):
    """vb_gemm_plan: the mapping vb_gemm_bf16 chooses, as host arithmetic (148 SMs assumed without a device). ViT-B token rows
    at 64 / 128 / 256 / 512 images per GPU: the N = 768 GEMMs with a 192-column variant (proj / fc2 forward with the residual
    epilogue, fc1 / qkv dgrad) take 192-column tiles exactly where that saves a wave of the 74 CTA pairs; everything else and
    ViT-L (N = 1024) stays at 256; weight gradients get the smallest split-K that fills >= 90 % of the last wave."""
    from vit_plasticity_b200 import _lib as L

    gp = ctypes.sizeof(L.GemmPlan)
    assert gp == 9 * 4
    handle = L.lib()
    before = handle.vb_get_gemm_tile_n(), handle.vb_get_gemm_cta_pair()
    handle.vb_set_gemm_tile_n(0)
    handle.vb_set_gemm_cta_pair(1)
    try:
        for images, want in ((64, 192), (128, 192), (256, 256), (512, 256)):
            m = images * 197
            for kw in (dict(k=768, epilogue=L.EPI_BF16_RESID), dict(k=3072, epilogue=L.EPI_BF16_RESID),
                       dict(k=3072, b_layout=1, epilogue=L.EPI_BF16), dict(k=2304, b_layout=1, epilogue=L.EPI_BF16)):
                p = L.gemm_plan(m, 768, **kw)
                assert (p["cta_pair"], p["tile_m"], p["tile_n"], p["units"]) == (1, 256, want, 74), (images, kw, p)
                assert p["n_tiles"] == 768 // want and p["m_tiles"] == -(-m // 256)
                assert p["waves"] == -(-p["m_tiles"] * p["n_tiles"] // 74)
            # no 192-column variant: the row-dot epilogue (proj dgrad), the GELU epilogue, a dgrad with fused column sums
            assert L.gemm_plan(m, 768, 768, b_layout=1, epilogue=L.EPI_BF16_ROWDOT)["tile_n"] == 256
            assert L.gemm_plan(m, 3072, 768, epilogue=L.EPI_BF16_GELU_GRAD)["tile_n"] == 256
            assert L.gemm_plan(m, 768, 3072, b_layout=1, epilogue=L.EPI_BF16, out_colsum=True)["tile_n"] == 256
            assert L.gemm_plan(m, 1024, 1024, epilogue=L.EPI_BF16_RESID)["tile_n"] == 256  # ViT-L: 1024 is no multiple of 192
        # 64 images: 50 row blocks x 3 tiles = 2.03 waves (3 paid) against 50 x 4 tiles of 3/4 the size = 2.7 waves (3 paid)
        assert L.gemm_plan(64 * 197, 768, 768, epilogue=L.EPI_BF16)["waves"] == 3  # K-major B without residual: 256 only
        assert L.gemm_plan(64 * 197, 768, 768, epilogue=L.EPI_BF16_RESID)["waves"] == 3
        # forced widths
        handle.vb_set_gemm_tile_n(256)
        assert L.gemm_plan(64 * 197, 768, 768, epilogue=L.EPI_BF16_RESID)["tile_n"] == 256
        handle.vb_set_gemm_tile_n(192)
        assert L.gemm_plan(512 * 197, 768, 768, epilogue=L.EPI_BF16_RESID)["tile_n"] == 192
        handle.vb_set_gemm_tile_n(0)
        # weight gradients (both operands MN-major, fp32 reduce-add): auto split-K; fc1 at 512 images: 3 x 12 tiles
        w = L.gemm_plan(3072, 768, 512 * 197, a_layout=1, b_layout=1, epilogue=L.EPI_F32_ADD, split_k=0)
        assert w["m_tiles"] * w["n_tiles"] == 36 and w["split_k"] >= 2
        fill = w["m_tiles"] * w["n_tiles"] * w["split_k"] / (w["waves"] * 74)
        assert fill >= 0.9, w
        assert w["k_blocks_per_split"] >= 8
        # a single 128-row block is not paired
        s1 = L.gemm_plan(100, 768, 768)
        assert (s1["cta_pair"], s1["tile_m"], s1["units"]) == (0, 128, 148)
        assert handle.vb_gemm_plan(None, None) == 1 and b"null" in handle.vb_last_error()
    finally:
        handle.vb_set_gemm_tile_n(before[0])
        handle.vb_set_gemm_cta_pair(before[1])


def test_sass_is_blackwell_native(built):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG/UTMAREDG (B200_PROFILING.md)."""
    obj = ROOT / "vit_plasticity_b200" / "csrc" / "gemm_tcgen05.o"
    if not obj.exists():
        pytest.skip("object file not kept")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
