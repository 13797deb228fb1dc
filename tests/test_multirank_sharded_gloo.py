"""CPU (gloo, world_size 2): host-side logic of the sharded mode - ownership / residency functions agree with the library,
every block has exactly one owner and is resident on its owner plus at most one neighbour, IPC handle exchange returns
every rank's bytes in rank order."""
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
    axis, origin, thick = multi.slab_layout(world, 0.005, extent_m=(-0.24, 0.24))  # 12 blocks of 4 cm -> 6 per rank
    ok = ok and (axis, origin, thick) == (0, -6, 6)
    for x in range(-8, 8):
        for y in range(-2, 2):
            for z in range(-2, 2):
                o = multi.owner_of_block(x, y, z, world, axis, origin, thick)
                ok = ok and o == lib.itm_b200_shard_owner_of_block(x, y, z, world, axis, origin, thick) and 0 <= o < world
                res = [multi.block_resident(x, y, z, r, world, axis, origin, thick) for r in range(world)]
                ok = ok and res == [bool(lib.itm_b200_shard_block_resident(x, y, z, r, world, axis, origin, thick)) for r in range(world)]
                ok = ok and res[o] and 1 <= sum(res) <= 2  # on the owner, plus the neighbour for the boundary layer
                ok = ok and (sum(res) == 2) == (x in (-1, 0))  # ... which is exactly the two block layers at the cut
                owned += o == rank
    n_blocks = 16 * 4 * 4
    import torch
    t = torch.tensor([owned])
    dist.all_reduce(t)
    ok = ok and int(t[0]) == n_blocks  # every block has exactly one owner
    ok = ok and owned == n_blocks // world  # ... and this symmetric range splits evenly
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
