"""ctypes binding of ``libvitb200.so`` (C ABI declared in ``include/vitb200.h``).

This is the only place where Python touches the native library. Every wrapper takes torch CUDA tensors, passes raw
device pointers / sizes / the current CUDA stream, and raises ``RuntimeError`` on a non-zero return code. There is no
CPU or library fallback: if the shared object is missing the import of :func:`lib` fails loudly.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

import torch

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libvitb200.so"

# enum vb_epilogue
EPI_BF16 = 0
EPI_BF16_RESID = 1
EPI_BF16_GELU = 2
EPI_BF16_DGELU = 3
EPI_F32 = 4
EPI_F32_ADD = 5
EPI_SUMSQ = 6
EPI_BF16_GELU_GRAD = 7
EPI_BF16_MULAUX = 8
EPI_BF16_ROWDOT = 9


class GemmArgs(Structure):
    """Mirror of ``struct vb_gemm_args``."""

    _fields_ = [
        ("a", c_void_p),
        ("b", c_void_p),
        ("lda", c_int64),
        ("ldb", c_int64),
        ("a_layout", c_int32),
        ("b_layout", c_int32),
        ("m", c_int32),
        ("n", c_int32),
        ("k", c_int32),
        ("epilogue", c_int32),
        ("bias", c_void_p),
        ("aux", c_void_p),
        ("ld_aux", c_int64),
        ("out", c_void_p),
        ("ld_out", c_int64),
        ("out2", c_void_p),
        ("ld_out2", c_int64),
        ("sumsq", c_void_p),
        ("rows_per_sample", c_int32),
        ("cols_per_group", c_int32),
        ("n_groups", c_int32),
        ("split_k", c_int32),
        ("out_colsum", c_void_p),
    ]


class GemmPlan(Structure):
    """Mirror of ``struct vb_gemm_plan_t``."""

    _fields_ = [(n, c_int32) for n in ("cta_pair", "tile_m", "tile_n", "m_tiles", "n_tiles", "split_k", "k_blocks_per_split", "units", "waves")]


# name -> (restype, argtypes); checked against `nm -D` by tests/test_cabi.py
SIGNATURES = {
    "vb_version": (c_int32, []),
    "vb_last_error": (c_char_p, []),
    "vb_launch_count": (c_int64, []),
    "vb_reset_launch_count": (None, []),
    "vb_gemm_bf16": (c_int32, [POINTER(GemmArgs), c_void_p]),
    "vb_preprocess_u8": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p],
    ),
    "vb_gemm_plan": (c_int32, [POINTER(GemmArgs), POINTER(GemmPlan)]),
    "vb_set_gemm_cta_pair": (None, [c_int32]),
    "vb_get_gemm_cta_pair": (c_int32, []),
    "vb_set_gemm_scheduler": (None, [c_int32]),
    "vb_get_gemm_scheduler": (c_int32, []),
    "vb_set_gemm_tile_n": (None, [c_int32]),
    "vb_get_gemm_tile_n": (c_int32, []),
    "vb_layernorm_fwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p],
    ),
    "vb_layernorm_bwd_workspace_bytes": (c_int64, [c_int32]),
    "vb_layernorm_bwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_attention_bwd_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "vb_attention_bwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_bwd_bias": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_bwd_with_delta": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_bwd_with_delta_qbias": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_pair_delta": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_pair_delta_layers": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_attention_perturb_delta_layers": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_layernorm_delta_sqdiff": (c_int32, [c_void_p, c_void_p, c_float, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "vb_scale_bf16": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "vb_philox_normal_f32": (c_int32, [c_void_p, c_int64, c_int64, c_uint64, c_uint64, c_void_p]),
    "vb_cast_f32_to_bf16": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_cast_bf16_to_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_im2col_patches": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_assemble_tokens": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p],
    ),
    "vb_assemble_tokens_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "vb_colsum_bf16": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p]),
    "vb_add_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_pool_tokens": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_rowsumsq_diff_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "vb_layernorm_pair_sqdiff": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "vb_sumsq_partials_f32": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p]),
    "vb_sgd_momentum_clip_step": (
        c_int32,
        [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_float, c_float, c_float, c_float, c_int32, c_void_p, c_void_p],
    ),
    "vb_adamw_clip_step": (
        c_int32,
        [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p] + [c_float] * 8 + [c_void_p, c_void_p],
    ),
}

_lib = None
# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream:
# entries are (start_event, end_event, algorithmic_flops)
GEMM_EVENTS = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {(_PKG_DIR / 'csrc')}`. There is no fallback path."
            )
        handle = ctypes.CDLL(os.fspath(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vb_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: vit_plasticity_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def launch_count() -> int:
    return int(lib().vb_launch_count())


def reset_launch_count() -> None:
    lib().vb_reset_launch_count()


# --------------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------------
def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    m: int,
    n: int,
    k: int,
    a_layout: int = 0,
    b_layout: int = 0,
    epilogue: int = EPI_BF16,
    bias: torch.Tensor | None = None,
    aux: torch.Tensor | None = None,
    out: torch.Tensor | None = None,
    out2: torch.Tensor | None = None,
    sumsq: torch.Tensor | None = None,
    rows_per_sample: int = 0,
    cols_per_group: int = 0,
    n_groups: int = 0,
    split_k: int = 1,
    out_colsum: torch.Tensor | None = None,
) -> None:
    """C[m,n] = epilogue(sum_k A[m,k] B[n,k]); see ``vb_gemm_bf16`` in include/vitb200.h.

    ``a``/``b`` are 2-D bf16 tensors whose last dim is contiguous (row stride = ``stride(0)``); with layout 0 they are
    stored [m,k] / [n,k], with layout 1 they are stored [k,m] / [k,n].
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    args = GemmArgs()
    args.a, args.b = a.data_ptr(), b.data_ptr()
    args.lda, args.ldb = a.stride(0), b.stride(0)
    args.a_layout, args.b_layout = a_layout, b_layout
    args.m, args.n, args.k = m, n, k
    args.epilogue = epilogue
    args.bias = _ptr(bias)
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if aux is not None:
        _req(aux, torch.bfloat16, "aux")
        args.aux, args.ld_aux = aux.data_ptr(), aux.stride(0)
    if out is not None:
        args.out, args.ld_out = out.data_ptr(), out.stride(0)
    if out2 is not None:
        args.out2, args.ld_out2 = out2.data_ptr(), out2.stride(0)
    if sumsq is not None:
        _req(sumsq, torch.float32, "sumsq")
        args.sumsq = sumsq.data_ptr()
    args.rows_per_sample, args.cols_per_group, args.n_groups = rows_per_sample, cols_per_group, n_groups
    args.split_k = split_k
    if out_colsum is not None:
        _req(out_colsum, torch.float32, "out_colsum")
        args.out_colsum = out_colsum.data_ptr()
    events = GEMM_EVENTS
    if events is not None:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
    _check(lib().vb_gemm_bf16(ctypes.byref(args), _stream()), "vb_gemm_bf16")
    if events is not None:
        end.record()
        events.append((start, end, 2.0 * m * n * k))


def gemm_plan(m: int, n: int, k: int, *, a_layout: int = 0, b_layout: int = 0, epilogue: int = EPI_BF16, split_k: int = 1, out_colsum: bool = False) -> dict:
    """How ``gemm`` would map this shape onto the persistent grid (tile mapping / width, split-K, waves): host arithmetic only,
    works without a GPU (the grid is then sized for 148 SMs)."""
    args = GemmArgs()
    args.m, args.n, args.k = m, n, k
    args.a_layout, args.b_layout, args.epilogue, args.split_k = a_layout, b_layout, epilogue, split_k
    args.out_colsum = 1 if out_colsum else None  # only its null-ness matters to the plan; never dereferenced there
    plan = GemmPlan()
    _check(lib().vb_gemm_plan(ctypes.byref(args), ctypes.byref(plan)), "vb_gemm_plan")
    return {name: getattr(plan, name) for name, _ in GemmPlan._fields_}


# --------------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, *, want_stats=True):
    _req(x, torch.bfloat16, "x")
    rows, cols = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    _check(
        lib().vb_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(mean), _ptr(rstd), rows, cols, float(eps), _stream()),
        "vb_layernorm_fwd",
    )
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, dgamma=None, dbeta=None, dres_colsum=None):
    """dx = dres + LN'(dy); dgamma / dbeta / dres_colsum (f32 [cols]) are accumulated into (+=) when given."""
    rows, cols = x.shape
    dx = torch.empty_like(x)
    if dres_colsum is not None:
        _req(dres_colsum, torch.float32, "dres_colsum")
    _check(
        lib().vb_layernorm_bwd(
            dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(dres), dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(dres_colsum), rows, cols, _stream()
        ),
        "vb_layernorm_bwd",
    )
    return dx


# --------------------------------------------------------------------------------------------------
# Attention core
# --------------------------------------------------------------------------------------------------
def attention_fwd(qkv, batch, seq, heads, head_dim, *, want_lse=True):
    _req(qkv, torch.bfloat16, "qkv")
    e = heads * head_dim
    out = torch.empty(batch * seq, e, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(batch, heads, seq, device=qkv.device, dtype=torch.float32) if want_lse else None
    _check(lib().vb_attention_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse), batch, seq, heads, head_dim, _stream()), "vb_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, batch, seq, heads, head_dim, *, dbias=None, delta=None, q_bias_only=False):
    """dqkv; with ``dbias`` (f32 [3E], accumulated into) the same kernel also reduces the column sums of dqkv. ``delta``
    (f32 [batch, heads, seq] = rowsum(dout * out) per head, e.g. from the ROWDOT epilogue of the GEMM that produced
    ``dout``): ``out`` is then not read and no delta pass is launched. ``q_bias_only`` (with delta and dbias): only the
    query third of dbias is reduced here; the caller gets the value third as the column sums of ``dout`` from the GEMM that
    produced it, and the key third is zero."""
    dqkv = torch.empty_like(qkv)
    if delta is not None:
        _req(delta, torch.float32, "delta")
        assert delta.numel() == batch * heads * seq and delta.is_contiguous()
        if dbias is not None:
            _req(dbias, torch.float32, "dbias")
        if q_bias_only and dbias is not None:
            _check(
                lib().vb_attention_bwd_with_delta_qbias(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), delta.data_ptr(), batch, seq, heads, head_dim, _stream()),
                "vb_attention_bwd_with_delta_qbias",
            )
            return dqkv
        _check(
            lib().vb_attention_bwd_with_delta(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), _ptr(dbias), delta.data_ptr(), batch, seq, heads, head_dim, _stream()),
            "vb_attention_bwd_with_delta",
        )
        return dqkv
    ws = torch.empty(int(lib().vb_attention_bwd_workspace_bytes(batch, seq, heads)), device=qkv.device, dtype=torch.uint8)
    if dbias is not None:
        _req(dbias, torch.float32, "dbias")
        _check(
            lib().vb_attention_bwd_bias(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), ws.data_ptr(), batch, seq, heads, head_dim, _stream()),
            "vb_attention_bwd_bias",
        )
        return dqkv
    _check(
        lib().vb_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), batch, seq, heads, head_dim, _stream()),
        "vb_attention_bwd",
    )
    return dqkv


def attention_pair_delta(qkv_a, qkv_b, delta, batch, seq, heads, head_dim):
    """delta[:, :] = attn(qkv_a) - attn(qkv_b); qkv_* are [batch*seq, 3E] views (row stride may exceed 3E)."""
    assert qkv_a.stride(0) == qkv_b.stride(0) and qkv_a.stride(1) == 1 and qkv_b.stride(1) == 1 and delta.stride(1) == 1
    _check(
        lib().vb_attention_pair_delta(qkv_a.data_ptr(), qkv_b.data_ptr(), qkv_a.stride(0), delta.data_ptr(), delta.stride(0), batch, seq, heads, head_dim, _stream()),
        "vb_attention_pair_delta",
    )


def attention_pair_delta_layers(qkv_a, qkv_b, layers, batch, seq, heads, head_dim):
    """[layers, batch*seq, E] bf16: attn(qkv_a) - attn(qkv_b) for every layer's slice of the concatenated projections."""
    assert qkv_a.stride(0) == qkv_b.stride(0) and qkv_a.stride(1) == 1 and qkv_b.stride(1) == 1
    delta = torch.empty(layers, batch * seq, heads * head_dim, device=qkv_a.device, dtype=torch.bfloat16)
    _check(
        lib().vb_attention_pair_delta_layers(qkv_a.data_ptr(), qkv_b.data_ptr(), qkv_a.stride(0), delta.data_ptr(), layers, batch, seq, heads, head_dim, _stream()),
        "vb_attention_pair_delta_layers",
    )
    return delta


def attention_perturb_delta_layers(qkv_a, dqkv, layers, batch, seq, heads, head_dim, out=None):
    """[layers, batch*seq, E] bf16: attn(a + d) - attn(a) for every layer, in perturbation form: ``qkv_a`` holds the
    projections of the base tokens (with bias), ``dqkv`` those of the token difference (no bias)."""
    _req(qkv_a, torch.bfloat16, "qkv_a")
    _req(dqkv, torch.bfloat16, "dqkv")
    assert qkv_a.stride(0) == dqkv.stride(0) and qkv_a.stride(1) == 1 and dqkv.stride(1) == 1 and qkv_a.shape == dqkv.shape
    if out is None:
        out = torch.empty(layers, batch * seq, heads * head_dim, device=qkv_a.device, dtype=torch.bfloat16)
    _check(
        lib().vb_attention_perturb_delta_layers(qkv_a.data_ptr(), dqkv.data_ptr(), qkv_a.stride(0), out.data_ptr(), layers, batch, seq, heads, head_dim, _stream()),
        "vb_attention_perturb_delta_layers",
    )
    return out


# --------------------------------------------------------------------------------------------------
# Element-wise helpers
# --------------------------------------------------------------------------------------------------
def scale_bf16(src: torch.Tensor, scale: float, dst: torch.Tensor | None = None) -> torch.Tensor:
    _req(src, torch.bfloat16, "src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty_like(src)
    _check(lib().vb_scale_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), float(scale), _stream()), "vb_scale_bf16")
    return dst


def philox_normal(n_images: int, shape, seed: int, first_image: int, device) -> torch.Tensor:
    """f32 [n_images, *shape] standard normals; row i is the noise of image ``first_image + i`` whatever the batching."""
    elems = 1
    for d in shape:
        elems *= int(d)
    out = torch.empty(n_images, *shape, device=device, dtype=torch.float32)
    _check(lib().vb_philox_normal_f32(out.data_ptr(), n_images, elems, int(seed), int(first_image), _stream()), "vb_philox_normal_f32")
    return out


def cast_f32_to_bf16(src: torch.Tensor, dst: torch.Tensor | None = None) -> torch.Tensor:
    _req(src, torch.float32, "src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    _check(lib().vb_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vb_cast_f32_to_bf16")
    return dst


def cast_bf16_to_f32(src: torch.Tensor) -> torch.Tensor:
    _req(src, torch.bfloat16, "src")
    assert src.is_contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    _check(lib().vb_cast_bf16_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vb_cast_bf16_to_f32")
    return dst


def im2col_patches(img: torch.Tensor, p: int, img2: torch.Tensor | None = None) -> torch.Tensor:
    _req(img, torch.float32, "img")
    assert img.is_contiguous() and (img2 is None or (img2.is_contiguous() and img2.shape == img.shape))
    n, c, h, w = img.shape
    patches = torch.empty(n * (h // p) * (w // p), c * p * p, device=img.device, dtype=torch.bfloat16)
    _check(lib().vb_im2col_patches(img.data_ptr(), _ptr(img2), patches.data_ptr(), n, c, h, w, p, _stream()), "vb_im2col_patches")
    return patches


def preprocess_u8(src, params, tab_bounds, tab_coef, max_src_rows, lut, out, *, want_f32=True, patch=0):
    """uint8 [n, h, w, 3] -> (f32 [n, 3, out, out] | None, bf16 patch rows | None); see ``vb_preprocess_u8``."""
    _req(src, torch.uint8, "src")
    assert src.dim() == 4 and src.shape[3] == 3 and src.is_contiguous()
    for t, name in ((tab_bounds, "tab_bounds"), (tab_coef, "tab_coef")):
        _req(t, torch.int32, name)
    _req(lut, torch.float32, "lut")
    if params is not None:
        _req(params, torch.int32, "params")
        assert params.shape == (src.shape[0], 8) and params.is_contiguous()
    n, h, w, _ = src.shape
    n_tables, out_t, ksize = tab_coef.shape
    assert out_t == out and tab_bounds.shape == (n_tables, out, 2)
    img = torch.empty(n, 3, out, out, device=src.device, dtype=torch.float32) if want_f32 else None
    patches = torch.empty(n * (out // patch) ** 2, 3 * patch * patch, device=src.device, dtype=torch.bfloat16) if patch else None
    _check(
        lib().vb_preprocess_u8(src.data_ptr(), n, h, w, _ptr(params), tab_bounds.data_ptr(), tab_coef.data_ptr(), n_tables, ksize, max_src_rows, lut.data_ptr(), out, _ptr(img), _ptr(patches), patch, _stream()),
        "vb_preprocess_u8",
    )
    return img, patches


def assemble_tokens(patch_out, patch_out_f32, cls, pos, batch, np_, e, *, want_bf16=True, want_f32=False):
    dev = cls.device
    tokens = torch.empty(batch * (np_ + 1), e, device=dev, dtype=torch.bfloat16) if want_bf16 else None
    tokens_f32 = torch.empty(batch * (np_ + 1), e, device=dev, dtype=torch.float32) if want_f32 else None
    _check(
        lib().vb_assemble_tokens(_ptr(patch_out), _ptr(patch_out_f32), cls.data_ptr(), pos.data_ptr(), _ptr(tokens), _ptr(tokens_f32), batch, np_, e, _stream()),
        "vb_assemble_tokens",
    )
    return tokens, tokens_f32


def assemble_tokens_bwd(dtokens, dcls, dpos, batch, np_, e):
    dpatch = torch.empty(batch * np_, e, device=dtokens.device, dtype=torch.bfloat16)
    _check(lib().vb_assemble_tokens_bwd(dtokens.data_ptr(), dpatch.data_ptr(), _ptr(dcls), _ptr(dpos), batch, np_, e, _stream()), "vb_assemble_tokens_bwd")
    return dpatch


def colsum_bf16(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[c] += sum_r x[r, c]"""
    _req(x, torch.bfloat16, "x")
    _req(out, torch.float32, "out")
    assert x.dim() == 2 and x.stride(1) == 1
    _check(lib().vb_colsum_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), x.shape[0], x.shape[1], _stream()), "vb_colsum_bf16")


def add_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(a)
    _check(lib().vb_add_bf16(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), "vb_add_bf16")
    return y


def pool_tokens(x: torch.Tensor, n: int, seq: int, cls_pooling: bool, normalize: bool = True) -> torch.Tensor:
    """f32 [n, dim]: cls row or token mean of a bf16 [n*seq, dim] / [n, seq, dim] activation, L2-normalised per row."""
    _req(x, torch.bfloat16, "x")
    assert x.is_contiguous()
    dim = x.shape[-1]
    out = torch.empty(n, dim, device=x.device, dtype=torch.float32)
    _check(lib().vb_pool_tokens(x.data_ptr(), out.data_ptr(), n, seq, dim, int(cls_pooling), int(normalize), _stream()), "vb_pool_tokens")
    return out


def rowsumsq_diff_f32(a, b, out, n_samples, rows_per_sample, cols):
    _check(lib().vb_rowsumsq_diff_f32(a.data_ptr(), _ptr(b), out.data_ptr(), n_samples, rows_per_sample, cols, _stream()), "vb_rowsumsq_diff_f32")


def layernorm_pair_sqdiff(a, b, u, n_samples, rows_per_sample, cols, eps):
    _check(lib().vb_layernorm_pair_sqdiff(a.data_ptr(), b.data_ptr(), u.data_ptr(), n_samples, rows_per_sample, cols, float(eps), _stream()), "vb_layernorm_pair_sqdiff")


def layernorm_delta_sqdiff(a, d, scale, u, n_samples, rows_per_sample, cols, eps):
    """u[s, c] += sum_rows (zhat(a + scale d) - zhat(a))^2, evaluated in perturbation form (no cancellation)"""
    _req(a, torch.float32, "a")
    _req(d, torch.float32, "d")
    _check(lib().vb_layernorm_delta_sqdiff(a.data_ptr(), d.data_ptr(), float(scale), u.data_ptr(), n_samples, rows_per_sample, cols, float(eps), _stream()), "vb_layernorm_delta_sqdiff")


# --------------------------------------------------------------------------------------------------
# Fused optimizer step
# --------------------------------------------------------------------------------------------------
def sumsq_partials_f32(x: torch.Tensor, partials: torch.Tensor) -> None:
    """partials[b] = block b's share of sum(x^2); deterministic (no atomics)"""
    _req(x, torch.float32, "x")
    _req(partials, torch.float32, "partials")
    assert x.is_contiguous() and partials.is_contiguous()
    _check(lib().vb_sumsq_partials_f32(x.data_ptr(), x.numel(), partials.data_ptr(), partials.numel(), _stream()), "vb_sumsq_partials_f32")


def adamw_clip_step(table, n_chunks, grad_arena, exp_avg, exp_avg_sq, partials, norm_out, max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2, hyper=None):
    _check(
        lib().vb_adamw_clip_step(table.data_ptr(), n_chunks, grad_arena.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), partials.data_ptr(), partials.numel(),
                                 _ptr(norm_out), float(max_norm), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), float(bc1), float(bc2),
                                 _ptr(hyper), _stream()),
        "vb_adamw_clip_step",
    )


def sgd_momentum_clip_step(table, n_chunks, grad_arena, momentum_arena, partials, norm_out, max_norm, lr, momentum, weight_decay, first_step, hyper=None):
    _check(
        lib().vb_sgd_momentum_clip_step(table.data_ptr(), n_chunks, grad_arena.data_ptr(), _ptr(momentum_arena), partials.data_ptr(), partials.numel(),
                                        _ptr(norm_out), float(max_norm), float(lr), float(momentum), float(weight_decay), int(first_step), _ptr(hyper), _stream()),
        "vb_sgd_momentum_clip_step",
    )
