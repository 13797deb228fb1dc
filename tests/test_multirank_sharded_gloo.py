"""CPU (gloo, world_size 2): host-side logic of the sharded mode - ownership functions agree with the library, every
block / raycast tile has exactly one owner, IPC handle exchange returns every rank's bytes in rank order."""
import os
import socket
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from infinitam_b200 import capi, multi

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank * 16 + i for i in range(3)]) * 64  # three fake 64-byte handles
    got = multi.exchange_handles(mine)
    ok = len(got) == world and all(got[r] == bytes([r * 16 + i for i in range(3)]) * 64 for r in range(world))
    lib = capi.load()
    owned = 0
    for x in range(-6, 6):
        for y in range(-6, 6):
            for z in range(-6, 6):
                o = multi.owner_of_block(x, y, z, world)
                ok = ok and o == lib.itm_b200_shard_owner_of_block(x, y, z, world) and 0 <= o < world
                owned += o == rank
    tiles = [multi.owner_of_raycast_tile(tx, ty, 80, world) for ty in range(90) for tx in range(80)]
    ok = ok and abs(tiles.count(rank) - len(tiles) / world) <= 1
    import torch
    t = torch.tensor([owned])
    dist.all_reduce(t)
    ok = ok and int(t[0]) == 12 ** 3  # every block has exactly one owner
    ok = ok and abs(owned - 12 ** 3 / world) < 0.15 * 12 ** 3  # ... and the split is roughly even
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharding_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
