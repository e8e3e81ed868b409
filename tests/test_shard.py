"""GOP sharding logic on CPU: world_size 2 over gloo."""
import os
import random
import socket

import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import pcamv_loader
    pcamv = pcamv_loader.load()
    from pcamv_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gops = shard.gop_ranges(600, 75)
    mine = shard.assign_gops(len(gops), world)[rank]
    local = []
    for g in mine:
        rnd = random.Random(1000 + g)
        n = 50 + 7 * g
        import hashlib
        local.append({"gop": g, "n_bits": n, "payload": bytes(rnd.getrandbits(1) for _ in range(n)), "n_mv": 5 * n,
                      "n_flipped": n // 3, "bytes": 1000 + g, "md5": hashlib.md5(b"gop%d" % g).hexdigest(),
                      "payload_md5": hashlib.md5(b"pay%d" % g).hexdigest()})
    res = shard.gather_gop_results(local)
    t = shard.max_over_ranks([0.5 + rank, 2.0 - rank])
    s = shard.sum_over_ranks([float(len(mine))])
    q.put((rank, res, t, s))
    dist.barrier()
    dist.destroy_process_group()


def test_gop_sharding_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    got = dict()
    for _ in range(world):
        rank, res, t, s = q.get(timeout=120)
        got[rank] = (res, t, s)
    for p in procs: p.join(timeout=60)
    res0, t0, s0 = got[0]
    assert got[1][0] is None
    assert [r["gop"] for r in res0] == list(range(8))
    assert [r["rank"] for r in res0] == [0, 1] * 4            # GOP g ran on rank g mod N
    for r in res0:
        rnd = random.Random(1000 + r["gop"])
        assert r["payload"] == bytes(rnd.getrandbits(1) for _ in range(r["n_bits"]))     # payload arrives intact, in GOP order
        assert r["n_mv"] == 5 * r["n_bits"] and r["bytes"] == 1000 + r["gop"]
        import hashlib
        assert r["md5"] == hashlib.md5(b"gop%d" % r["gop"]).hexdigest()      # 128-bit digests survive the int64 transport
        assert r["payload_md5"] == hashlib.md5(b"pay%d" % r["gop"]).hexdigest()
    assert t0 == [1.5, 2.0] and got[1][1] == [1.5, 2.0]          # max over ranks
    assert s0 == [8.0]


def test_gop_ranges_and_assignment(pcamv):
    from pcamv_b200 import shard
    assert shard.gop_ranges(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert shard.gop_ranges(600, 75)[-1] == (525, 75) and len(shard.gop_ranges(600, 75)) == 8
    assert shard.assign_gops(8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert shard.assign_gops(3, 8)[3:] == [[]] * 5
