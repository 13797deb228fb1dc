"""Multi-GPU plumbing for the "batches of independent sequences, one scene per GPU" mode.

The fusion path shards by scene: every rank owns whole scenes (hash table + voxel blocks + tracking state stay on one
GPU), so there is no collective on the data path.  torch.distributed is used only to agree on the partition, to
barrier around the timed region and to combine the per-rank timings (max over ranks).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition_scenes(n_scenes: int, world_size: int, rank: int):
    """Contiguous, balanced assignment of scene ids to ranks (first ranks get the remainder)."""
    base, rem = divmod(n_scenes, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def sequence_start_for_scene(scene_id: int) -> int:
    """BASELINE configs[3]: 64 copies of the sequence with phase-shifted trajectories."""
    return (7 * scene_id) % 100


def combine_timing(local_ms: float, local_frames: int, device=None):
    """Returns (max over ranks of the elapsed ms, total frames over all ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_ms), int(local_frames)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([int(local_frames)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def aggregate_frames_per_second(local_ms: float, local_frames: int, device=None) -> float:
    ms, frames = combine_timing(local_ms, local_frames, device)
    return frames / (ms * 1e-3)


# ------------------------------------------------------------------------------------------------------------------
# Spatial sharding of ONE scene across the GPUs of an NVLink domain (BASELINE configs[2]; include/itm_b200.h)

def _floor_div(a: int, b: int) -> int:
    return a // b  # Python's // already floors


def owner_of_block(x: int, y: int, z: int, world: int, axis: int, origin_block: int, thickness_blocks: int) -> int:
    """Rank that owns the voxel block at block coordinate (x, y, z): slabs of `thickness_blocks` along `axis`, the outer
    ranks open-ended - the same function the kernels use (shard_owner_of_block in infinitam_b200/csrc/kernels.h)."""
    c = (x, y, z)[axis]
    return min(max(_floor_div(c - origin_block, thickness_blocks), 0), world - 1)


def block_resident(x: int, y: int, z: int, rank: int, world: int, axis: int, origin_block: int, thickness_blocks: int, halo: int = 1) -> bool:
    """Does `rank` keep the block's voxels?  Owned, or within `halo` blocks of its slab (the ray cast's trilinear taps and
    normals reach one block across the boundary; a wider halo lets a rank march more rays completely)."""
    if world <= 1:
        return True
    c = (x, y, z)[axis]
    lo = origin_block + rank * thickness_blocks
    hi = lo + thickness_blocks
    return (rank == 0 or c >= lo - halo) and (rank == world - 1 or c < hi + halo)


def slab_layout(world: int, voxel_size: float, extent_m=(-2.0, 2.0), axis: int = 0):
    """(axis, origin_block, thickness_blocks) cutting [extent_m) along `axis` into `world` slabs of whole blocks"""
    import math
    block_m = 8.0 * voxel_size
    lo = math.floor(extent_m[0] / block_m)
    hi = math.ceil(extent_m[1] / block_m)
    return axis, lo, max(2, math.ceil((hi - lo) / world))


def exchange_handles(local_handles: bytes, group=None):
    """all_gather of this rank's packed IPC handles -> list (one bytes object per rank).  Works on any backend."""
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(local_handles), dtype=torch.uint8)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = t.to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def compare_scene(eng, single, rank, world, layout, voxel_size, halo=1):
    """One rank of a sharded scene (its ITMMainEngine view) against a single-GPU engine that fused the same frames with the
    same poses: index bit-identical, ptr >= 0 exactly on resident blocks, resident voxel blocks bit-identical, composed
    raycast image within tolerance.  Test / bench instrumentation (host reads of whole buffers); returns a dict of findings."""
    import numpy as np

    from . import capi
    r = {}
    hs, h1 = eng.read(capi.BUF_HASH), single.read(capi.BUF_HASH)
    r["hash_pos_offset_equal"] = bool(np.array_equal(hs["pos"], h1["pos"]) and np.array_equal(hs["offset"], h1["offset"]))
    _, _, st_s = eng.get_state()
    _, _, st_1 = single.get_state()
    r["visible_count_equal"] = int(st_s[0]) == int(st_1[0])
    r["excess_counter_equal"] = int(st_s[2]) == int(st_1[2])
    r["visible_list_equal"] = bool(np.array_equal(eng.read(capi.BUF_VISIBLE_IDS)[: st_s[0]], single.read(capi.BUF_VISIBLE_IDS)[: st_1[0]]))
    alloc = np.nonzero(h1["ptr"] >= 0)[0]
    pos = h1["pos"][alloc].astype(np.int64)
    axis, origin, thick = layout
    c = pos[:, axis]
    lo = origin + rank * thick
    hi = lo + thick
    resident = ((rank == 0) | (c >= lo - halo)) & ((rank == world - 1) | (c < hi + halo))
    owner = np.clip((c - origin) // thick, 0, world - 1)
    r["allocated_blocks"] = int(len(alloc))
    r["owned_blocks"] = int((owner == rank).sum())
    r["resident_blocks"] = int(resident.sum())
    ptr_s = hs["ptr"][alloc]
    r["residency_matches_ptr"] = bool(np.array_equal(ptr_s >= 0, resident) and np.all(ptr_s[~resident] == -1))
    r["free_blocks_used"] = int(eng.params.sdf_local_block_num - 1 - st_s[1])
    # voxels of every resident block against the single GPU's
    vs = eng.read(capi.BUF_VOXELS).reshape(-1, 512)
    v1 = single.read(capi.BUF_VOXELS).reshape(-1, 512)
    ids = alloc[resident & (ptr_s >= 0)]
    a = vs[hs["ptr"][ids]] & 0x00FFFFFF
    b = v1[h1["ptr"][ids]] & 0x00FFFFFF
    r["resident_voxel_blocks_equal"] = bool(np.array_equal(a, b))
    r["resident_voxel_blocks_differing"] = int(np.any(a != b, axis=1).sum())
    # composed raycast
    W, H = eng.W, eng.H
    rs = eng.read(capi.BUF_RAYCAST_RESULT).reshape(H, W, 4)
    r1 = single.read(capi.BUF_RAYCAST_RESULT).reshape(H, W, 4)
    hit_s, hit_1 = rs[..., 3] > 0, r1[..., 3] > 0
    both = hit_s & hit_1
    d = np.abs(rs[..., :3] - r1[..., :3])[both].max(axis=1) * voxel_size if both.any() else np.zeros(1)
    r["raycast_hits_single"] = int(hit_1.sum())
    r["raycast_hit_mismatch"] = int((hit_s != hit_1).sum())
    r["raycast_max_diff_m"] = float(d.max())
    r["raycast_over_1e-4_m"] = int((d > 1e-4).sum())
    r["raycast_bit_equal_px"] = float(np.mean(np.all(rs[both] == r1[both], axis=1))) if both.any() else 1.0
    # pixels no rank could march completely are reported as misses and counted by the engine; every other pixel - hit or miss -
    # must be the single GPU's bits
    r["raycast_unresolved_px"] = int(eng.shard_unresolved())
    r["raycast_px_differing"] = int(np.any(rs.view(np.uint32) != r1.view(np.uint32), axis=2).sum())
    ps, p1 = eng.read(capi.BUF_POINTS).reshape(H, W, 4), single.read(capi.BUF_POINTS).reshape(H, W, 4)
    r["icp_point_validity_mismatch"] = int(((ps[..., 3] > 0) != (p1[..., 3] > 0)).sum())
    return r


class ShardedEngine:
    """ITMMainEngine for one scene spread over all ranks of the default process group (one process per GPU).

    Rank 0 feeds the frames; EnqueueFrame broadcasts the raw depth image with NCCL and every rank enqueues the frame on its
    own GPU.  The voxel payload is partitioned (slabs of block coordinates, see slab_layout), the per-rank partial raycast
    images are composed inside the frame by peer reads over NVLink; there is no other collective."""

    def __init__(self, params, stream=None, layout=None, halo=None, peers=None):
        """stream: raw cudaStream_t handle shared with torch (torch.cuda.Stream().cuda_stream, made current), so that the
        NCCL broadcast and the engine's kernels are ordered on one stream.  Must not be the legacy default stream (0).
        layout: (axis, origin_block, thickness_blocks); default: the synthetic room's x extent cut into world slabs.
        params.sdf_local_block_num is the voxel pool PER RANK."""
        import ctypes as C

        from . import capi
        from .engines import ITMMainEngine

        self.lib = capi.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > capi.MAX_SHARDS:
            raise ValueError("at most %d ranks" % capi.MAX_SHARDS)
        self.params = params
        if not stream:
            raise ValueError("ShardedEngine needs an explicit torch stream handle (see docstring)")
        if params.device != torch.cuda.current_device():
            raise ValueError("params.device (%d) must be torch's current device (%d)" % (params.device, torch.cuda.current_device()))
        self.layout = layout if layout is not None else slab_layout(self.world, params.voxel_size)
        # blocks beyond its slab a rank keeps resident.  One is what the trilinear taps need; three were measured at 1280x720 /
        # 2 mm on 4 GPUs: 8.3 k unresolved pixels instead of 9.5 k - they are silhouette rays that pass through the band of a
        # near surface in one slab and end on a far one in another, not rays near a slab boundary - so the default stays 1
        import os
        self.halo = int(halo if halo is not None else os.environ.get("ITM_B200_SHARD_HALO", 1))
        W, H = params.width, params.height
        tiles = ((W + 15) // 16) * ((H + 7) // 8)
        # one peer-visible allocation per rank: [partial image 0 | partial image 1 | tile flags 0 | tile flags 1 | barrier words]
        img = W * H * 16
        tile_bytes = (tiles + 255) // 256 * 256
        self._offsets = (0, img, 2 * img, 2 * img + tile_bytes, 2 * img + 2 * tile_bytes)
        total = self._offsets[4] + 256
        p, h = C.c_void_p(), C.create_string_buffer(capi.IPC_HANDLE_BYTES)
        capi.check(self.lib.itm_b200_ipc_alloc(total, C.byref(p), h))
        self._local = p
        all_handles = exchange_handles(h.raw)
        sh = capi.Shard()
        sh.rank, sh.world = self.rank, self.world
        sh.axis, sh.origin_block, sh.thickness_blocks = self.layout
        sh.stream = stream
        sh.halo_blocks = self.halo
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                base = self._local.value
            else:
                q = C.c_void_p()
                capi.check(self.lib.itm_b200_ipc_open(all_handles[r][:capi.IPC_HANDLE_BYTES], C.byref(q)))
                self._opened.append(q)
                base = q.value
            for parity in range(2):
                sh.partial_raycast_dev[parity][r] = base + self._offsets[parity]
                sh.tile_hit_dev[parity][r] = base + self._offsets[2 + parity]
            sh.barrier_flags_dev[r] = base + self._offsets[4]
        h = C.c_void_p()
        capi.check(self.lib.itm_b200_engine_create_sharded(C.byref(params), C.byref(sh), C.byref(h)))
        # a plain ITMMainEngine view of the handle gives Sync / read / stage access
        self.engine = ITMMainEngine.__new__(ITMMainEngine)
        self.engine.lib, self.engine.params, self.engine.W, self.engine.H, self.engine.h = self.lib, params, W, H, h
        self._raw = torch.empty((H, W), dtype=torch.int16, device=torch.device("cuda", torch.cuda.current_device()))
        # every rank's voxel pool and hash table, peer-visible: the few rays no rank can march on its own voxels read the
        # blocks held elsewhere from their owners (ITM_B200_SHARD_PEERS=0: they stay misses)
        import os
        self.peers = (bool(peers) if peers is not None else os.environ.get("ITM_B200_SHARD_PEERS", "1") != "0") and self.world > 1
        if self.peers:
            hv, hh = C.create_string_buffer(capi.IPC_HANDLE_BYTES), C.create_string_buffer(capi.IPC_HANDLE_BYTES)
            capi.check(self.lib.itm_b200_engine_shard_export(h, hv, hh))
            peers = exchange_handles(hv.raw + hh.raw)
            pv, ph = (C.c_void_p * capi.MAX_SHARDS)(), (C.c_void_p * capi.MAX_SHARDS)()
            for r in range(self.world):
                if r == self.rank:
                    continue
                for arr, off in ((pv, 0), (ph, capi.IPC_HANDLE_BYTES)):
                    q = C.c_void_p()
                    capi.check(self.lib.itm_b200_ipc_open(peers[r][off:off + capi.IPC_HANDLE_BYTES], C.byref(q)))
                    self._opened.append(q)
                    arr[r] = q.value
            capi.check(self.lib.itm_b200_engine_shard_attach(h, pv, ph))
        dist.barrier()  # nobody reads a peer's buffers before that peer exists

    def EnqueueFrame(self, raw_depth_dev=None):
        """raw_depth_dev: int16 CUDA tensor on rank 0 (ignored elsewhere).  Broadcast + enqueue, no host sync."""
        if self.rank == 0:
            self._raw.copy_(raw_depth_dev, non_blocking=True)
        dist.broadcast(self._raw.view(torch.uint8), src=0)  # NCCL has no int16: the frame travels as bytes
        self.engine.EnqueueFrameDevice(self._raw.data_ptr())

    def Sync(self):
        return self.engine.Sync()

    def close(self):
        dist.barrier()
        self.engine.close()
        for p in self._opened:
            self.lib.itm_b200_ipc_close(p)
        dist.barrier()
        self.lib.itm_b200_ipc_free(self._local)
        self._opened, self._local = [], None
