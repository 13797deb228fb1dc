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
# Work sharding of ONE scene across the GPUs of an NVLink domain (BASELINE configs[2]; include/itm_b200.h, "work sharding")

def owner_of_block(x: int, y: int, z: int, world: int) -> int:
    """Rank that integrates the voxel block at block coordinate (x, y, z) - same function as the kernels use
    (shard_owner_of_block in infinitam_b200/csrc/kernels.h)."""
    m = 0xFFFFFFFF
    h = ((x & m) * 73856093 & m) ^ ((y & m) * 19349669 & m) ^ ((z & m) * 83492791 & m)
    return (h >> 7) % world


def owner_of_raycast_tile(tile_x: int, tile_y: int, tiles_per_row: int, world: int) -> int:
    """16x8-pixel raycast tiles are dealt out round-robin in raster order (k_raycast)."""
    return (tile_y * tiles_per_row + tile_x) % world


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


class ShardedEngine:
    """ITMMainEngine for one scene shared by all ranks of the default process group (one process per GPU).

    Rank 0 feeds the frames; ProcessFrame broadcasts the raw depth image with NCCL and every rank enqueues the frame on
    its own GPU.  There is no other collective: voxel and raycast results travel as peer stores inside the kernels."""

    def __init__(self, params, stream=None):
        """stream: raw cudaStream_t handle shared with torch (torch.cuda.Stream().cuda_stream, made current), so that the
        NCCL broadcast and the engine's kernels are ordered on one stream.  Must not be the legacy default stream (0)."""
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
        W, H = params.width, params.height
        sizes = [params.sdf_local_block_num * 512 * 4, W * H * 16, capi.MAX_SHARDS * 4]
        self._local, handles = [], b""
        for nbytes in sizes:
            p, h = C.c_void_p(), C.create_string_buffer(capi.IPC_HANDLE_BYTES)
            capi.check(self.lib.itm_b200_ipc_alloc(nbytes, C.byref(p), h))
            self._local.append(p)
            handles += h.raw
        all_handles = exchange_handles(handles)
        sh = capi.Shard()
        sh.rank, sh.world = self.rank, self.world
        sh.stream = stream
        self._opened = []
        for r in range(self.world):
            for k, arr in enumerate((sh.voxel_blocks_dev, sh.raycast_result_dev, sh.barrier_flags_dev)):
                if r == self.rank:
                    arr[r] = self._local[k].value
                else:
                    p = C.c_void_p()
                    hb = all_handles[r][k * capi.IPC_HANDLE_BYTES:(k + 1) * capi.IPC_HANDLE_BYTES]
                    capi.check(self.lib.itm_b200_ipc_open(hb, C.byref(p)))
                    self._opened.append(p)
                    arr[r] = p.value
        h = C.c_void_p()
        capi.check(self.lib.itm_b200_engine_create_sharded(C.byref(params), C.byref(sh), C.byref(h)))
        # a plain ITMMainEngine view of the handle gives ProcessFrame / Sync / read / stage access
        self.engine = ITMMainEngine.__new__(ITMMainEngine)
        self.engine.lib, self.engine.params, self.engine.W, self.engine.H, self.engine.h = self.lib, params, W, H, h
        self._raw = torch.empty((H, W), dtype=torch.int16, device=torch.device("cuda", torch.cuda.current_device()))
        dist.barrier()  # nobody stores into a peer before that peer has reset its copy

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
        for p in self._local:
            self.lib.itm_b200_ipc_free(p)
        self._opened, self._local = [], []
