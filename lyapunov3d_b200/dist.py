"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

Every pixel and every voxel is independent, so there is no
communication inside a render or a bake; ranks only hand their finished piece to
rank 0 with point-to-point sends (NCCL over NVLink on GPUs, gloo in the CPU tests):

* frame      -> tile x tile pixel tiles dealt round-robin (tile j to rank j % world): costs
                vary 100x between pixels and are spatially coherent, so bands would not balance
* volume     -> contiguous z-slabs, received straight into place in rank 0's volume
* animation  -> whole frames, frame f to rank f % world

The partition is pure re-indexing: the assembled result is bit-identical to the
single-GPU one (tests/test_gpu_parity.py, tests/test_dist_cpu.py).

`render_fn` / `bake_fn` default to the CUDA entry points; the CPU tests inject stand-ins
to exercise the index arithmetic and the gather without a GPU.
"""
import torch
import torch.distributed as dist

from . import api

DEFAULT_TILE = 8


def _rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def slab_range(nz, rank, world):
    """Planes [z0, z1) of rank `rank`: contiguous, sizes differ by at most one."""
    return nz * rank // world, nz * (rank + 1) // world


def frames_of_rank(n_frames, rank, world):
    return list(range(rank, n_frames, world))


def tile_pixel_index(width, height, tile, rank, world):
    """Image index (x + y*width) of every work item of a rank, in work order; -1 marks the
    padding of ragged edge tiles.  Mirrors item_to_pixel() in csrc/kernels/kernels.cuh."""
    tiles_x = (width + tile - 1) // tile
    tiles_y = (height + tile - 1) // tile
    n_tiles = tiles_x * tiles_y
    mine = torch.arange(rank, max(n_tiles, rank), world, dtype=torch.int64)
    p = torch.arange(tile * tile, dtype=torch.int64)
    x = (mine % tiles_x)[:, None] * tile + (p % tile)[None, :]
    y = (mine // tiles_x)[:, None] * tile + (p // tile)[None, :]
    idx = x + y * width
    idx[(x >= width) | (y >= height)] = -1
    return idx.reshape(-1)


def _place(image_flat, compact, index):
    ok = index >= 0
    image_flat[index[ok].to(image_flat.device)] = compact[ok.to(compact.device)]


def _gather_to_root(mine, sizes, group=None):
    """Point-to-point gather of differently sized tensors; returns the list on rank 0."""
    rank, world = _rank_world(group)
    if world == 1:
        return [mine]
    if rank != 0:
        dist.send(mine.contiguous(), dst=0, group=group)
        return None
    out = [mine]
    for r in range(1, world):
        buf = torch.empty((sizes[r],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        dist.recv(buf, src=r, group=group)
        out.append(buf)
    return out


def render_frame_sharded(cam, prm, seq, lights, num_lights, width, height, mode="exact", tile=DEFAULT_TILE,
                         group=None, want_points=True, render_fn=None):
    """One frame across all ranks (BASELINE config 3).  Returns (rgba[h,w,4], points[h,w,36], evals)
    on rank 0 and (None, None, evals_of_this_rank) elsewhere."""
    rank, world = _rank_world(group)
    render_fn = render_fn or api.render
    c_rgba, c_pts, evals = render_fn(cam, prm, seq, lights, num_lights, width, height, mode=mode,
                                     tile=tile, rank=rank, world=world, compact=True)
    sizes = [api.tile_count(width, height, tile, r, world) if render_fn is api.render
             else int(tile_pixel_index(width, height, tile, r, world).numel()) for r in range(world)]
    g_rgba = _gather_to_root(c_rgba, sizes, group)
    g_pts = _gather_to_root(c_pts, sizes, group) if want_points else None
    if rank != 0:
        return None, None, evals
    dev = c_rgba.device
    rgba = torch.zeros((height * width, 4), dtype=torch.uint8, device=dev)
    pts = torch.zeros((height * width, 36), dtype=torch.uint8, device=dev) if want_points else None
    for r in range(world):
        if dev.type == "cuda" and render_fn is api.render:
            api.scatter_tiles(rgba, g_rgba[r], width, height, tile, r, world)
            if want_points:
                api.scatter_tiles(pts, g_pts[r], width, height, tile, r, world)
        else:
            index = tile_pixel_index(width, height, tile, r, world)
            _place(rgba, g_rgba[r], index)
            if want_points:
                _place(pts, g_pts[r], index)
    return rgba.view(height, width, 4), (pts.view(height, width, 36) if want_points else None), evals


def bake_sharded(prm, seq, nx, ny=None, nz=None, mode="fast", dtype="f32", group=None, bake_fn=None, device=None):
    """Volume bake across all ranks by z-slabs (BASELINE config 4).  Rank 0 returns the full
    [nz,ny,nx] volume, the others None."""
    rank, world = _rank_world(group)
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    bake_fn = bake_fn or api.bake
    z0, z1 = slab_range(nz, rank, world)
    if rank == 0:
        vol = bake_fn(prm, seq, nx, ny, nz, z0=z0, z1=z1, mode=mode, dtype=dtype)
        for r in range(1, world):
            a, b = slab_range(nz, r, world)
            if b > a:
                dist.recv(vol[a:b], src=r, group=group)      # contiguous: lands in place
        return vol
    # other ranks only ever hold their slab
    if bake_fn is api.bake:
        slab = _bake_slab_only(prm, seq, nx, ny, nz, z0, z1, mode, dtype)
    else:
        slab = bake_fn(prm, seq, nx, ny, nz, z0=z0, z1=z1, mode=mode, dtype=dtype)[z0:z1]
    if z1 > z0:
        dist.send(slab.contiguous(), dst=0, group=group)
    return None


def _bake_slab_only(prm, seq, nx, ny, nz, z0, z1, mode, dtype):
    """Bake planes z0..z1 into a buffer that holds just those planes."""
    tdt = torch.float16 if dtype in ("f16", api.F16) else torch.float32
    slab = torch.zeros((max(z1 - z0, 0), ny, nx), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
    if z1 > z0:
        # the kernel indexes the FULL volume: bias the base pointer so plane z0 is slab[0]
        api.bake_into(slab, z0, prm, seq, nx, ny, nz, z0, z1, mode, dtype)
    return slab


def render_animation_sharded(n_frames, width, height, prm, cam, seq, lights, num_lights, mode="exact",
                             group=None, render_fn=None, on_frame=None):
    """The scale.pl camera orbit, frame f on rank f % world (BASELINE config 5).  Rank 0 gets
    every frame (list of rgba tensors, or `on_frame(f, rgba)` calls); others return None."""
    from .structs import clone
    rank, world = _rank_world(group)
    render_fn = render_fn or api.render
    mine = {}
    for f in frames_of_rank(n_frames, rank, world):
        c = clone(cam)
        api.campath_frame(f, n_frames, c)
        api.scene_cam_recalculate(c, width, height, 1)
        mine[f] = render_fn(c, prm, seq, lights, num_lights, width, height, mode=mode)[0]
    frames = [None] * n_frames if rank == 0 else None
    for f in range(n_frames):
        owner = f % world
        if owner == 0:
            if rank == 0:
                frames[f] = mine[f]
        elif rank == owner:
            dist.send(mine[f].contiguous(), dst=0, group=group)
        elif rank == 0:
            buf = torch.empty((height, width, 4), dtype=torch.uint8, device=next(iter(mine.values())).device if mine else "cpu")
            dist.recv(buf, src=owner, group=group)
            frames[f] = buf
        if rank == 0 and on_frame is not None:
            on_frame(f, frames[f])
            frames[f] = None
    return frames


# ----------------------------------------------------------------------------------
# Peer-memory variants: no gather at all.  Rank 0 owns the result buffer; every other
# rank maps it through CUDA IPC and its kernels store their shard straight into rank 0's
# HBM over NVLink/NVSwitch while they compute (40 B per pixel per ~5e5 iterations, 4 B per
# voxel per 1026: the link is idle by comparison, so the transfer is free).
# ----------------------------------------------------------------------------------
class PeerBuffer:
    """`nbytes` of device memory on rank 0, mapped into every rank of the group."""

    def __init__(self, nbytes, group=None):
        import ctypes as C
        self.rank, self.world = _rank_world(group)
        self.nbytes, self.group = int(nbytes), group
        L = api.lib()
        handle = torch.zeros(64, dtype=torch.uint8)
        ptr = C.c_void_p()
        if self.rank == 0:
            api._check(L.lyap_peer_alloc(C.byref(ptr), self.nbytes), "lyap_peer_alloc")
            if self.world > 1:
                buf = (C.c_ubyte * 64)()
                api._check(L.lyap_peer_export(ptr, buf), "lyap_peer_export")
                handle = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device())
            h = handle.to(dev)
            dist.broadcast(h, src=0, group=group)
            if self.rank != 0:
                raw = (C.c_ubyte * 64)(*h.cpu().tolist())
                api._check(L.lyap_peer_open(raw, C.byref(ptr)), "lyap_peer_open")
        self.ptr = ptr.value

    def view(self, shape, typestr):
        return api.DevicePointer(self.ptr, shape, typestr)

    def zero_(self):
        """Owner only: clear the buffer (stream-ordered on the current stream)."""
        if self.rank == 0:
            self.view((self.nbytes,), "|u1").tensor().zero_()

    def close(self):
        """Importers unmap first, then a barrier, then the owner frees: cudaFree of an exported
        allocation while another process still has it open is undefined behaviour."""
        L = api.lib()
        if not self.ptr:
            return
        ptr, self.ptr = self.ptr, 0
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)      # every rank's kernels have finished storing
            rc = L.lyap_peer_close(ptr) if self.rank != 0 else 0
            dist.barrier(group=self.group)      # every importer has closed its mapping
            api._check(rc, "lyap_peer_close")
        if self.rank == 0:
            api._check(L.lyap_peer_free(ptr), "lyap_peer_free")


def bake_sharded_peer(volume, prm, seq, nx, ny, nz, mode="fast", f16=False, group=None):
    """Every rank bakes its z-slab directly into `volume` (a PeerBuffer of the full volume on
    rank 0).  Returns after all ranks' kernels have completed."""
    rank, world = _rank_world(group)
    z0, z1 = slab_range(nz, rank, world)
    api.bake_ptr(volume.ptr, f16, prm, seq, nx, ny, nz, z0, z1, mode)
    torch.cuda.current_stream().synchronize()
    if world > 1:
        dist.barrier(group=group)


def render_frame_sharded_peer(rgba, points, cam, prm, seq, d_lights, num_lights, width, height, mode="exact",
                              tile=DEFAULT_TILE, group=None, evals=None):
    """Every rank renders its interleaved tiles directly into rank 0's `rgba` / `points`
    PeerBuffers at their image positions.  `points` must have been zeroed by rank 0 (and a
    barrier passed) if defined miss pixels are wanted."""
    rank, world = _rank_world(group)
    api.render_into(rgba.ptr, points.ptr, cam, prm, seq, d_lights, num_lights, width, height, mode, tile, rank, world, evals)
    torch.cuda.current_stream().synchronize()
    if world > 1:
        dist.barrier(group=group)
