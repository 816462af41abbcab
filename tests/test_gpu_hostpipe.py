"""Host-buffer front end (supres.HostPipeline / the DSen2_20 facade): pinned and pageable host arrays, float32 and
uint16 digital numbers (s2_tiles_supres.py:311-315 hands GDAL uint16), sharded patch ranges.  Every route must give the
pixels of the device-resident pipeline bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def scene():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dsen2_b200.DSen2Net import s2model
    rng = np.random.RandomState(3)
    H, W = 1140, 692                               # 11 x 7 patches, clamped last row and column
    d10 = rng.randint(0, 12000, size=(H, W, 4)).astype(np.uint16)
    d20 = rng.randint(0, 12000, size=(H // 2, W // 2, 6)).astype(np.uint16)
    model = s2model(((4, None, None), (6, None, None)), num_layers=2, feature_size=128, seed=5)
    return torch, model, d10, d20


def _device_reference(torch, model, d10, d20):
    from dsen2_b200 import supres
    return supres.super_resolve_device(model, torch.from_numpy(d10.astype(np.float32)).cuda(),
                                       torch.from_numpy(d20.astype(np.float32)).cuda()).cpu().numpy()


def test_facade_uint16_float32_and_other_dtypes_are_bit_identical(scene):
    torch, model, d10, d20 = scene
    from dsen2_b200 import supres
    ref = _device_reference(torch, model, d10, d20)
    got_u16 = supres.DSen2_20(d10, d20, model=model)
    assert got_u16.dtype == np.float32 and got_u16.shape == ref.shape
    assert np.array_equal(got_u16, ref)
    assert np.array_equal(supres.DSen2_20(d10.astype(np.float32), d20.astype(np.float32), model=model), ref)
    assert np.array_equal(supres.DSen2_20(d10.astype(np.int32), d20.astype(np.float64), model=model), ref)   # cast while staged
    out = np.full(ref.shape, -1, np.float32)
    assert supres.DSen2_20(d10, d20, model=model, out=out) is out and np.array_equal(out, ref)
    # inputs are never modified (supres.py:17-19 copies)
    assert d10.dtype == np.uint16 and d10.max() < 12000


def test_pinned_and_pageable_routes_agree_on_sharded_ranges(scene):
    """Patch ranges that start and end inside a patch row (what a rank of N owns): partial-row output rectangles."""
    torch, model, d10, d20 = scene
    from dsen2_b200 import sharding, supres
    ref = _device_reference(torch, model, d10, d20)
    H, W = d10.shape[:2]
    for dtype, cast in ((torch.uint16, np.uint16), (torch.float32, np.float32)):
        pipe = supres.HostPipeline(model, H, W, dtype=dtype, chunk_patch_rows=2)
        h10, h20 = torch.from_numpy(d10.astype(cast)).pin_memory(), torch.from_numpy(d20.astype(cast)).pin_memory()
        hout = torch.full((H, W, 6), -1.0).pin_memory()
        got_np = np.full((H, W, 6), -1, np.float32)
        for r in (1, 2, 0):
            first, count = sharding.shard_range(77, r, 3)            # 25/26 patches: mid-row seams
            pipe.run(h10, h20, hout=hout, first_patch=first, num_patches=count)
            pipe.run_numpy(d10.astype(cast), d20.astype(cast), out=got_np, first_patch=first, num_patches=count)
        torch.cuda.synchronize()
        assert np.array_equal(hout.numpy(), ref)
        assert np.array_equal(got_np, ref)
        assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0


def test_facade_60m_uint16(scene):
    torch, _, _, _ = scene
    from dsen2_b200 import supres
    from dsen2_b200.DSen2Net import s2model
    rng = np.random.RandomState(9)
    H, W = 600, 348
    d10 = rng.randint(0, 9000, size=(H, W, 4)).astype(np.uint16)
    d20 = rng.randint(0, 9000, size=(H // 2, W // 2, 6)).astype(np.uint16)
    d60 = rng.randint(0, 9000, size=(H // 6, W // 6, 2)).astype(np.uint16)
    model = s2model(((4, None, None), (6, None, None), (2, None, None)), num_layers=1, feature_size=128, seed=6)
    a = supres.DSen2_60(d10, d20, d60, model=model)
    b = supres.DSen2_60(d10.astype(np.float32), d20.astype(np.float32), d60.astype(np.float32), model=model)
    assert a.shape == (H, W, 2) and np.array_equal(a, b)
    c = supres.super_resolve_device(model, *[torch.from_numpy(x.astype(np.float32)).cuda() for x in (d10, d20, d60)])
    assert np.array_equal(a, c.cpu().numpy())


def test_facade_rejects_mismatched_shapes(scene):
    torch, model, d10, d20 = scene
    from dsen2_b200 import supres
    with pytest.raises(ValueError):
        supres.DSen2_20(d10, d20[:-1], model=model)
    with pytest.raises(ValueError):
        supres.DSen2_20(d10[:, :, :3], d20, model=model)
    with pytest.raises(ValueError):
        supres.DSen2_20(d10[:-1], d20, model=model)          # odd 10 m height
