"""Host-side multi-rank logic on CPU: contiguous patch sharding, ownership rectangles, gloo world_size-2 assembly."""
import os
import sys

import numpy as np
import pytest

from dsen2_b200 import sharding
from oracle import patches_oracle as po

GEOMS = [(300, 412, 128, 8), (560, 560, 128, 8), (600, 348, 192, 12), (10980, 10980, 128, 8), (10980, 10980, 192, 12)]


def test_shard_range_is_a_partition():
    for n in (1, 7, 36, 9801, 4356):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


@pytest.mark.parametrize('geom', GEOMS)
@pytest.mark.parametrize('world', [1, 2, 8])
def test_owned_rects_tile_the_image_exactly_once(geom, world):
    H, W, P, B = geom
    ny, nx, S = sharding.tile_grid(H, W, P, B)
    n = ny * nx
    cover = np.zeros((H, W), np.uint8)
    for r in range(world):
        f, c = sharding.shard_range(n, r, world)
        for (y0, y1, x0, x1) in sharding.owned_rects(f, c, H, W, P, B):
            cover[y0:y1, x0:x1] += 1
        r0, r1 = sharding.input_rows(f, c, H, W, P, B)
        y0, y1 = sharding.output_rows(f, c, H, W, P, B)
        assert 0 <= r0 <= y0 < y1 <= r1 <= H
    assert cover.min() == 1 and cover.max() == 1


@pytest.mark.parametrize('geom', GEOMS)
@pytest.mark.parametrize('world', [1, 3])
@pytest.mark.parametrize('chunk_rows', [1, 6])
def test_plan_chunks_covers_range_once_with_monotonic_uploads(geom, world, chunk_rows):
    """Host-buffer pipeline plan: chunks partition the rank's patch range, every owned pixel is downloaded exactly
    once, and uploading only the rows beyond the previous chunk's still gives every chunk all the rows it reads."""
    H, W, P, B = geom
    ny, nx, S = sharding.tile_grid(H, W, P, B)
    cover = np.zeros((H, W), np.uint8)
    for r in range(world):
        first, count = sharding.shard_range(ny * nx, r, world)
        plan = sharding.plan_chunks(first, count, H, W, P, B, chunk_rows, 1 if chunk_rows > 1 else None)
        assert [p for p, _, _, _ in plan] == sorted(p for p, _, _, _ in plan)
        assert sum(c for _, c, _, _ in plan) == count and plan[0][0] == first
        resident = None                                    # [lo, hi) rows on the device, as HostPipeline tracks them
        for p0, cnt, (r0, r1), rects in plan:
            assert cnt > 0 and 0 <= r0 < r1 <= H
            lo, hi = (r0, r1) if resident is None else (resident[0], max(resident[1], r1))
            assert resident is None or r0 >= resident[0], "row ranges must advance monotonically"
            resident = (lo, hi)
            assert resident[0] <= r0 and r1 <= resident[1]
            for (y0, y1, x0, x1) in rects:
                cover[y0:y1, x0:x1] += 1
    assert cover.min() == 1 and cover.max() == 1


def test_ownership_equals_sequential_overwrite():
    for (H, W, P, B) in GEOMS[:3]:
        ny, nx, S = sharding.tile_grid(H, W, P, B)
        n = ny * nx
        pred = np.empty((n, 1, P, P), np.float32)
        for i in range(n):
            pred[i] = i
        ref = po.recompose_images(pred, B, (H, W))[:, :, 0]
        for world in (2, 3):
            for r in range(world):
                f, c = sharding.shard_range(n, r, world)
                for (y0, y1, x0, x1) in sharding.owned_rects(f, c, H, W, P, B):
                    v = ref[y0:y1, x0:x1]
                    assert v.min() >= f and v.max() < f + c


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    H, W, P, B = 300, 412, 128, 8
    ny, nx, S = sharding.tile_grid(H, W, P, B)
    n = ny * nx
    rng = np.random.RandomState(5)
    pred = rng.rand(n, 3, P, P).astype(np.float32)                 # every rank regenerates the same "prediction"
    f, c = sharding.shard_range(n, rank, world)
    # what the rank's GPU would hold after super_resolve_device on its patch range: owned pixels only
    canvas = np.zeros((H, W, 3), np.float32)
    full = po.recompose_images(pred, B, (H, W))
    for (y0, y1, x0, x1) in sharding.owned_rects(f, c, H, W, P, B):
        canvas[y0:y1, x0:x1] = full[y0:y1, x0:x1]
    y0, y1 = sharding.output_rows(f, c, H, W, P, B)
    band = torch.from_numpy(np.ascontiguousarray(canvas[y0:y1]))
    meta = [None] * world
    dist.all_gather_object(meta, (f, c, y0, y1))
    if rank == 0:
        parts = [(f, c, y0, band.numpy(), P, B)]
        for src in range(1, world):
            fs, cs, ys0, ys1 = meta[src]
            buf = torch.empty((ys1 - ys0, W, 3))
            dist.recv(buf, src=src)
            parts.append((fs, cs, ys0, buf.numpy(), P, B))
        out = sharding.assemble(np.zeros((H, W, 3), np.float32), parts)
        np.save(os.path.join(tmp, 'ok.npy'), np.array([np.array_equal(out, full)]))
    else:
        dist.send(band, dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_assembly(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(str(tmp_path / 'ok.npy'))[0]


def test_compat_shims_importable():
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dsen2_b200', 'compat')
    sys.path.insert(0, compat)
    try:
        for m in ('supres', 'utils', 'utils.DSen2Net', 'utils.patches', 'utils.imresize'):
            sys.modules.pop(m, None)
        import supres as s
        from utils.DSen2Net import s2model
        from utils.imresize import imresize
        from utils.patches import get_test_patches, get_test_patches60, recompose_images
        assert s.SCALE == 2000 and callable(s.DSen2_20) and callable(s.DSen2_60)
        assert callable(s2model) and callable(imresize) and callable(get_test_patches)
        assert callable(get_test_patches60) and callable(recompose_images)
    finally:
        sys.path.remove(compat)
        for m in ('supres', 'utils', 'utils.DSen2Net', 'utils.patches', 'utils.imresize'):
            sys.modules.pop(m, None)


def test_batch_and_chunk_sizing_rules():
    """Device batches are whole patch rows (so host-pipeline chunks split evenly); chunks follow the call's share."""
    from dsen2_b200 import sharding
    from dsen2_b200.supres import default_device_batch
    assert default_device_batch(10980, 128, 8) == 3 * 99          # full tile, 20 m path: three patch rows
    assert default_device_batch(10980, 192, 12) == 2 * 66         # 60 m path: two rows of 192-pixel patches
    assert default_device_batch(600, 128, 8) >= 36                # a 600 x 600 scene is one launch
    assert default_device_batch(2352, 128, 8) % 21 == 0
    assert sharding.auto_chunk_rows(9801, 99) == (3, 3)           # one GPU: 99 patch rows, 33 full launches
    assert sharding.auto_chunk_rows(1226, 99) == (3, 1)           # a rank of 8: small first / last chunk, big ones between
    assert sharding.auto_chunk_rows(36, 6) == (1, 1)              # small range, no batch size given
    assert sharding.auto_chunk_rows(36, 6, 300) == (6, 6)         # a 600 x 600 scene fits one device batch: one chunk
    assert sharding.auto_chunk_rows(1226, 99, 297) == (3, 1)
    plan = sharding.plan_chunks(0, 9801, 10980, 10980, 128, 8, *sharding.auto_chunk_rows(9801, 99))
    assert len(plan) == 33 and sum(c[1] for c in plan) == 9801 and all(c[1] == 297 for c in plan)
    # rank 3 of 8: patches [3676, 4901) = rows 37.1 .. 49.5: the rest of row 37, three-row chunks, one row at the end
    first, count = sharding.shard_range(9801, 3, 8)
    plan = sharding.plan_chunks(first, count, 10980, 10980, 128, 8, *sharding.auto_chunk_rows(count, 99))
    sizes = [c[1] for c in plan]
    assert sum(sizes) == count and [c[0] for c in plan] == [first + sum(sizes[:i]) for i in range(len(sizes))]
    assert sizes[0] <= 99 and sizes[-1] <= 99 and max(sizes) == 297 and len(plan) <= 7
