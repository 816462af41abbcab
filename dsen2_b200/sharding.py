"""Patch-range sharding of one tile across ranks (one process per GPU; no collective on the inference path).

The filled patch list of a tile (row-major (ty, tx), patches.py:58-72) is split into contiguous ranges.
Every patch depends only on its own crop of the inputs, and stitching is ownership-based
(``dsen2_recompose``), so ranks never exchange data; the host (or a peer copy) assembles the disjoint
rectangles each rank owns.
"""


def shard_range(n_items, rank, world):
    """Contiguous split of [0, n_items) -- sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def tile_grid(H, W, P, border):
    S = P - 2 * border
    return -(-H // S), -(-W // S), S


def _own_lo(t, n, size, S):
    """First output coordinate owned by tile t along one axis (last writer wins, patches.py:394-403)."""
    if t >= n:
        return size
    if t == n - 1 and size % S != 0:
        return size - S
    return t * S


def owned_rects(first, count, H, W, P, border):
    """Disjoint rectangles (y0, y1, x0, x1) of output pixels whose last writer is in [first, first+count)."""
    ny, nx, S = tile_grid(H, W, P, border)
    rects = []
    p, end = first, first + count
    while p < end:
        ty, tx0 = divmod(p, nx)
        tx1 = min(nx, tx0 + (end - p))          # exclusive
        y0, y1 = _own_lo(ty, ny, H, S), _own_lo(ty + 1, ny, H, S)
        x0, x1 = _own_lo(tx0, nx, W, S), _own_lo(tx1, nx, W, S)
        if rects and rects[-1][2] == 0 and rects[-1][3] == W and x0 == 0 and x1 == W and rects[-1][1] == y0:
            rects[-1] = (rects[-1][0], y1, 0, W)   # merge full tile rows
        elif y1 > y0 and x1 > x0:
            rects.append((y0, y1, x0, x1))
        p += tx1 - tx0
    return rects


def output_rows(first, count, H, W, P, border):
    """Bounding row range [y0, y1) of the pixels owned by the patch range."""
    rects = owned_rects(first, count, H, W, P, border)
    if not rects:
        return 0, 0
    return min(r[0] for r in rects), max(r[1] for r in rects)


def input_rows(first, count, H, W, P, border):
    """10 m input rows [r0, r1) the patch range reads (symmetric padding folds back inside the image)."""
    if count <= 0:
        return 0, 0
    ny, nx, S = tile_grid(H, W, P, border)
    ty0, ty1 = first // nx, (first + count - 1) // nx

    def origin(t):
        return min(t * S, H - S)
    r0 = max(0, origin(ty0) - border)
    r1 = min(H, origin(ty1) + S + border)
    return r0, r1


def plan_chunks(first, count, H, W, P, border, chunk_patch_rows, edge_patch_rows=None):
    """Cut the patch range [first, first+count) into chunks of whole patch rows for the host-buffer pipeline.

    Returns [(p0, cnt, (r0, r1), rects)]: the chunk's patch range, the 10 m input rows it reads and the output
    rectangles it owns.  Consecutive chunks read monotonically advancing row ranges, so each input row needs to be
    uploaded once (``supres.HostPipeline`` uploads only the rows beyond the previous chunk's).
    ``edge_patch_rows``: size of the FIRST and the LAST chunk (default: ``chunk_patch_rows``).  The first chunk's upload
    and the last chunk's download are the only transfers that do not overlap with compute, so they are kept small while
    the chunks in between are large enough for full-size launches."""
    ny, nx, _ = tile_grid(H, W, P, border)
    edge = chunk_patch_rows if edge_patch_rows is None else edge_patch_rows
    end = first + count
    last_row = (end - 1) // nx if count > 0 else 0
    plan, p0 = [], first
    while p0 < end:
        row = p0 // nx
        rows_left = last_row - row + 1
        if p0 == first:
            step = edge
        elif rows_left <= edge:
            step = rows_left                           # the last chunk
        else:
            step = min(chunk_patch_rows, rows_left - edge)   # keep `edge` rows for the last chunk
        p1 = min(end, (row + max(1, step)) * nx)
        cnt = p1 - p0
        plan.append((p0, cnt, input_rows(p0, cnt, H, W, P, border), owned_rects(p0, cnt, H, W, P, border)))
        p0 = p1
    return plan


def auto_chunk_rows(num_patches, nx, device_batch=None):
    """Patch rows per host-pipeline chunk -> (middle, edge): three patch rows (a full 297-patch launch on a Sentinel-2
    tile) in the middle of the range, one at both ends when the range is short (a rank of 8 holds ~12 patch rows of a
    full tile: its un-overlapped first upload / last download would otherwise be a quarter of its work).  A range that
    fits ONE device batch (a 600 x 600 scene: 36 patches) is a single chunk: there is nothing to overlap, and every
    extra chunk costs a set of launches."""
    rows = -(-int(num_patches) // int(nx))
    if device_batch is not None and num_patches <= device_batch:
        return rows, rows
    middle = max(1, min(3, rows // 4))
    edge = 1 if rows < 48 else middle
    return middle, edge


def assemble(canvas, parts):
    """Host-side assembly: parts = [(first, count, rows_y0, band ndarray (y1-y0, W, C))]; writes owned rects."""
    H, W = canvas.shape[:2]
    for first, count, y_off, band, P, border in parts:
        for (y0, y1, x0, x1) in owned_rects(first, count, H, W, P, border):
            canvas[y0:y1, x0:x1] = band[y0 - y_off:y1 - y_off, x0:x1]
    return canvas
