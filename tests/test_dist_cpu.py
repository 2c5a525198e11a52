"""Multi-rank sharding logic on CPU: world_size 2 and 3 over gloo.

The CUDA entry points are replaced by stand-ins that paint every pixel / voxel with a value
derived from its own coordinates, so a wrong tile mapping, slab offset, frame owner or gather
order shows up as a wrong value at rank 0.  (The real kernels' partition invariance is checked
on the GPU in tests/test_gpu_parity.py::test_tile_partition_is_bit_identical.)
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lyapunov3d_b200 import dist as ld


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fake_render(cam, prm, seq, lights, num_lights, width, height, mode="exact", tile=None, rank=0, world=1, compact=False, **kw):
    """Pixel (x,y) -> rgba (x&255, y&255, (x+y)&255, 7); points bytes = index pattern."""
    if tile is None:
        idx = torch.arange(width * height)
    else:
        idx = ld.tile_pixel_index(width, height, tile, rank, world)
    x, y = idx % width, idx // width
    rgba = torch.stack([x & 255, y & 255, (x + y) & 255, torch.full_like(x, 7)], -1).to(torch.uint8)
    pts = ((idx[:, None] * 3 + torch.arange(36)[None, :]) & 255).to(torch.uint8)
    rgba[idx < 0] = 0
    pts[idx < 0] = 0
    if tile is None:
        return rgba.view(height, width, 4), pts.view(height, width, 36), torch.tensor([width * height])
    return rgba, pts, torch.tensor([int((idx >= 0).sum())])


def fake_bake(prm, seq, nx, ny=None, nz=None, z0=0, z1=None, mode="fast", dtype="f32", out=None):
    vol = torch.zeros((nz, ny, nx))
    z = torch.arange(z0, z1, dtype=torch.float32)[:, None, None]
    y = torch.arange(ny, dtype=torch.float32)[None, :, None]
    x = torch.arange(nx, dtype=torch.float32)[None, None, :]
    vol[z0:z1] = z * 10000 + y * 100 + x
    return vol


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, tile = 37, 21, 8
        rgba, pts, ev = ld.render_frame_sharded(None, None, None, None, 0, w, h, tile=tile, render_fn=fake_render)
        total = torch.tensor([int(ev.item())])
        dist.all_reduce(total)
        vol = ld.bake_sharded(None, None, 7, 5, 9, bake_fn=fake_bake)
        if rank == 0:
            want_rgba, want_pts, _ = fake_render(None, None, None, None, 0, w, h)
            assert torch.equal(rgba, want_rgba) and torch.equal(pts, want_pts)
            assert int(total.item()) == w * h                  # every pixel rendered exactly once
            assert torch.equal(vol, fake_bake(None, None, 7, 5, 9, 0, 9))
        else:
            assert rgba is None and vol is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frame_and_bake_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_partitions_cover_everything_exactly_once():
    for (w, h, tile, world) in [(1920, 1080, 8, 8), (100, 52, 8, 3), (37, 21, 16, 2), (3840, 2160, 8, 8), (5, 3, 8, 4)]:
        seen = np.zeros(w * h, np.int32)
        for r in range(world):
            idx = ld.tile_pixel_index(w, h, tile, r, world).numpy()
            np.add.at(seen, idx[idx >= 0], 1)
        assert (seen == 1).all(), (w, h, tile, world)
    for nz, world in [(512, 8), (9, 4), (3, 8), (100, 7)]:
        planes = [z for r in range(world) for z in range(*ld.slab_range(nz, r, world))]
        assert planes == list(range(nz))
        sizes = [b - a for a, b in (ld.slab_range(nz, r, world) for r in range(world))]
        assert max(sizes) - min(sizes) <= 1
    frames = sorted(f for r in range(8) for f in ld.frames_of_rank(120, r, 8))
    assert frames == list(range(120)) and len(ld.frames_of_rank(120, 3, 8)) == 15


def test_tile_index_matches_library_count():
    from lyapunov3d_b200 import api
    for (w, h, tile, world) in [(1920, 1080, 8, 8), (100, 52, 8, 3), (37, 21, 16, 2)]:
        for r in range(world):
            assert ld.tile_pixel_index(w, h, tile, r, world).numel() == api.tile_count(w, h, tile, r, world)
