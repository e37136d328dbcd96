"""CPU-side checks of the C ABI and the host shims (no kernel is launched):
libofb200.so loads and exports every symbol include/ofb200.h declares, the ctypes table binds all of
them, argument validation returns the documented codes, and the product path refuses to run without
a CUDA device instead of falling back."""
import ctypes
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
import ofb200  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(ofb200.LIB_PATH):
        ofb200.build()
    return ofb200.load()


def test_every_header_symbol_is_exported_and_bound(lib):
    syms = entry.header_symbols()
    assert len(syms) >= 16
    raw = ctypes.CDLL(ofb200.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/ofb200.h but not exported"
        assert s in ofb200.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.ofb_version() == 120
    assert lib.ofb_strerror(0) == b"ok"
    assert b"invalid" in lib.ofb_strerror(-1)


def test_pyramid_layout_is_host_only(lib):
    pyr = ofb200.Pyramid()
    elems = (ctypes.c_int64 * ofb200.MAX_LEVELS)()
    assert lib.ofb_pyramid_layout(47, 156, 4, 1, ctypes.byref(pyr), ctypes.byref(elems)) == 0
    assert [pyr.lvl_h[i] for i in range(4)] == [47, 23, 11, 5]
    assert [pyr.lvl_w[i] for i in range(4)] == [156, 78, 39, 19]       # floor: trailing odd rows / cols dropped
    for i in range(4):
        assert pyr.row_pitch[i] % 16 == 0 and pyr.row_pitch[i] >= pyr.lvl_w[i]
        assert pyr.q_stride[i] % 8 == 0 and pyr.q_stride[i] >= pyr.row_pitch[i] * pyr.lvl_h[i]
    assert pyr.layout == ofb200.LAYOUT_ROWS
    assert lib.ofb_pyramid_layout(47, 156, 4, 2, ctypes.byref(pyr), ctypes.byref(elems)) == 0
    assert pyr.layout == ofb200.LAYOUT_BLOCK8X4
    for i in range(4):      # 8x4 blocks: rows padded to a multiple of 4, columns to a multiple of 8
        assert pyr.row_pitch[i] == (pyr.lvl_w[i] + 7) // 8 * 8
        assert pyr.q_stride[i] == pyr.row_pitch[i] * ((pyr.lvl_h[i] + 3) // 4 * 4)
    assert lib.ofb_pyramid_layout(47, 156, 4, 3, ctypes.byref(pyr), ctypes.byref(elems)) == 0
    assert pyr.layout == ofb200.LAYOUT_QMINOR8X4 and [pyr.q_stride[i] for i in range(4)] == [32] * 4
    assert [elems[i] for i in range(4)] == [pyr.row_pitch[i] * ((pyr.lvl_h[i] + 3) // 4 * 4) for i in range(4)]
    assert lib.ofb_pyramid_layout(47, 156, 4, 0, ctypes.byref(pyr), ctypes.byref(elems)) == 0
    assert [pyr.row_pitch[i] for i in range(4)] == [156, 78, 39, 19]
    assert lib.ofb_pyramid_layout(4, 4, 4, 1, ctypes.byref(pyr), ctypes.byref(elems)) != 0   # level 3 would be empty
    assert lib.ofb_pyramid_layout(8, 8, 5, 1, ctypes.byref(pyr), ctypes.byref(elems)) != 0


def test_argument_validation_without_launch(lib):
    null = None
    assert lib.ofb_warp_f32(null, null, null, null, 1, 3, 8, 8, 0, 1, 0, 0, 0, 1.0, 1.0, null) == -1
    assert lib.ofb_convex_upsample_f32(null, null, null, 1, 4, 4, null) == -1
    assert lib.ofb_epe_reduce_f32(null, null, null, null, 1, 4, 4, null) == -1
    assert lib.ofb_corr_lookup(null, null, null, null, null, 1, 4, 4, 4, null) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from optical_flow import normalize, warp

    with pytest.raises(ofb200.OfbError):
        warp(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2))
    with pytest.raises(ofb200.OfbError):
        normalize(torch.zeros(1, 2, 2, 2))
    from model.corr import CorrBlock

    with pytest.raises(ofb200.OfbError):
        CorrBlock(torch.zeros(1, 64, 8, 8), torch.zeros(1, 64, 8, 8))


def test_reference_assertion_behaviour():
    """Shape preconditions are bare asserts, as in the reference (operator.py:74,77,105-106,128,144,160,163)."""
    from optical_flow import denormalize, integrate, normalize, resize, scale

    bad = torch.zeros(1, 3, 4, 4)
    for fn in (scale, normalize, denormalize, resize):
        with pytest.raises(AssertionError):
            fn(bad)
    with pytest.raises(AssertionError):
        scale(torch.zeros(1, 2, 4, 4), (1.0, 2.0, 3.0))
    with pytest.raises(AssertionError):
        integrate(torch.zeros(1, 2, 4, 4))


def test_pair_arena_layout_cpu():
    """ofb200.runner.PairArena (host logic, no GPU): every field is a view of one (pairs, bytes_per_pair) buffer at a
    256-byte aligned offset, with the field's usual shape and its own dtype; a pair's fields are contiguous in memory so
    a micro-batch is one slice of rows."""
    import torch

    from ofb200.runner import FIELDS, PairArena

    pairs, c, h, w, iters = 3, 4, 5, 6, 2
    g = torch.Generator().manual_seed(0)
    batch = {"fmap1": torch.randn(pairs, c, h, w, generator=g), "fmap2": torch.randn(pairs, c, h, w, generator=g),
             "coords": torch.randn(iters, pairs, 2, h, w, generator=g), "flow_lo": torch.randn(pairs, 2, h, w, generator=g),
             "up_mask": torch.randn(pairs, 576, h, w, generator=g), "frame": torch.rand(pairs, 3, 8 * h, 8 * w, generator=g),
             "target": torch.randn(pairs, 2, 8 * h, 8 * w, generator=g), "valid": torch.rand(pairs, 8 * h, 8 * w, generator=g)}
    arena = PairArena(pairs, PairArena.shapes_of(batch)).fill(batch)
    assert arena.buf.shape == (pairs, arena.pair_bytes) and arena.pair_bytes % PairArena.ALIGN == 0
    for k in FIELDS:
        assert arena.offsets[k] % PairArena.ALIGN == 0
        assert arena[k].shape == batch[k].shape and torch.equal(arena[k], batch[k])
        assert arena[k].data_ptr() == arena.buf.data_ptr() + arena.offsets[k]              # a view, not a copy
    # a micro-batch is a row range: its views see the same values
    for k in FIELDS:
        sub = arena.view(k, 1, 3)
        want = batch[k][:, 1:3] if k == "coords" else batch[k][1:3]
        assert torch.equal(sub, want)
    assert arena.payload_bytes_per_pair == 4 * sum(batch[k].numel() for k in FIELDS) // pairs
    # half-precision feature maps travel as they are: same views, half the bytes for those fields
    half = PairArena(pairs, PairArena.shapes_of(batch), dtypes={"fmap1": torch.bfloat16, "fmap2": torch.float16}).fill(batch)
    assert half["fmap1"].dtype == torch.bfloat16 and half["fmap2"].dtype == torch.float16 and half["up_mask"].dtype == torch.float32
    assert torch.equal(half["fmap1"], batch["fmap1"].bfloat16()) and torch.equal(half["fmap2"], batch["fmap2"].half())
    assert torch.equal(half["coords"], batch["coords"]) and torch.equal(half.view("frame", 1, 2), batch["frame"][1:2])
    assert half.payload_bytes_per_pair == arena.payload_bytes_per_pair - 2 * 2 * batch["fmap1"][0].numel()
    assert not half.same_layout(arena) and half.same_layout(half)


def test_input_padder_known_answers_cpu():
    """model.InputPadder (caller-side torch glue; behaviour of reference methods/raft/model/utils.py:38-61): the
    paddings SURVEY.md section 8 quotes -- Sintel 436 -> 440 centred, KITTI 375 x 1242 -> 376 x 1248 with the extra
    row at the bottom -- replicate values, and unpad inverts pad."""
    import torch

    from model.utils import InputPadder, coords_grid

    x = torch.arange(436 * 1024, dtype=torch.float32).view(1, 1, 436, 1024)
    p = InputPadder(x.shape)
    (y,) = p.pad(x)
    assert y.shape == (1, 1, 440, 1024) and torch.equal(y[0, 0, 0], x[0, 0, 0]) and torch.equal(y[0, 0, 2], x[0, 0, 0])
    assert torch.equal(y[0, 0, -1], x[0, 0, -1]) and torch.equal(p.unpad(y), x)
    k = torch.rand(2, 3, 375, 1242)
    pk = InputPadder(k.shape, mode="kitti")
    a, b = pk.pad(k, 2 * k)
    assert a.shape == (2, 3, 376, 1248) and torch.equal(a[:, :, 0, 3:-3], k[:, :, 0]) and torch.equal(a[:, :, -1, 3:-3], k[:, :, -1])
    assert torch.equal(a[:, :, :, 0], a[:, :, :, 3]) and torch.equal(pk.unpad(b), 2 * k)
    odd = torch.rand(1, 1, 7, 9)
    po = InputPadder(odd.shape)                      # 1 extra row -> bottom, 7 extra columns -> 3 left, 4 right
    (z,) = po.pad(odd)
    assert z.shape == (1, 1, 8, 16) and torch.equal(z[..., :7, 3:12], odd) and torch.equal(po.unpad(z), odd)
    g = coords_grid(2, 3, 4)
    assert g.shape == (2, 2, 3, 4) and g.dtype == torch.float32
    assert torch.equal(g[1, 0, 2], torch.arange(4.0)) and torch.equal(g[0, 1, :, 1], torch.arange(3.0))


def test_host_side_argument_errors_cpu():
    """Argument checks of the Python surface that fire before any device work (so they are testable without a GPU),
    with the reference's exception types: shape preconditions are bare asserts (operator.py:74,77,128,144,160)."""
    import pytest
    import torch

    from model import sequence_loss
    from optical_flow import denormalize, integrate, normalize, resize, scale, warp

    gt, valid = torch.zeros(1, 2, 4, 4), torch.ones(1, 4, 4)
    with pytest.raises(ValueError):
        sequence_loss([], gt, valid)
    with pytest.raises(NotImplementedError):
        sequence_loss([gt] * 25, gt, valid)
    with pytest.raises(RuntimeError):
        sequence_loss([torch.zeros(1, 2, 4, 5)], gt, valid)
    for fn in (normalize, denormalize, lambda f: scale(f, 2.0), lambda f: resize(f, size=(2, 2))):
        with pytest.raises(AssertionError):
            fn(torch.zeros(1, 3, 4, 4))
    with pytest.raises(AssertionError):
        scale(gt, (1.0, 2.0, 3.0))
    with pytest.raises(AssertionError):
        integrate(gt)
    with pytest.raises(AssertionError):
        integrate(gt, torch.zeros(1, 2, 4, 5))
    with pytest.raises(NotImplementedError):
        warp(torch.zeros(1, 3, 4, 4), gt, mode="bicubic")
    with pytest.raises(ValueError):
        warp(torch.zeros(1, 3, 4, 4), gt, padding_mode="wrap")
    with pytest.raises(NotImplementedError):
        resize(gt, size=(2, 2), mode="nearest")
    with pytest.raises(NotImplementedError):
        warp(torch.zeros(1, 3, 4, 4, dtype=torch.float64), gt.double())
