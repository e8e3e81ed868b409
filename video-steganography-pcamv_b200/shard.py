"""GOP sharding across the GPUs of one box.

An IDR resets the reference list, frame_num and POC (reference encoder/encoder.c:2247-2254, encoder/slicetype.c:601-620),
so with constant QP every IDR-bounded GOP is an independent encoder run: GOP g = frames [g*K, (g+1)*K) goes to rank
g mod N (the reference CLI's own --seek / --frames), one encoder context per shard, no data-path collective.  The one
exchange step is the end-of-run gather of per-GOP payload bits and statistics to rank 0, which concatenates the NAL
streams in GOP order; torch.distributed does it (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def gop_ranges(n_frames, keyint):
    """[(first_frame, n_frames_of_gop), ...] for IDR every `keyint` frames."""
    return [(s, min(keyint, n_frames - s)) for s in range(0, n_frames, keyint)]


def assign_gops(n_gops, world):
    """rank -> list of GOP indices (round robin: equal GOP lengths make it the longest-first schedule too)."""
    return [list(range(r, n_gops, world)) for r in range(world)]


def _md5_to_ints(hexdigest):
    v = int(hexdigest or "0", 16)
    hi, lo = v >> 64, v & ((1 << 64) - 1)
    return [hi - (1 << 64) if hi >= (1 << 63) else hi, lo - (1 << 64) if lo >= (1 << 63) else lo]


def _ints_to_md5(hi, lo):
    return "%032x" % (((hi & ((1 << 64) - 1)) << 64) | (lo & ((1 << 64) - 1)))


def gather_gop_results(local, group=None):
    """local: list of dicts {gop, n_bits, payload (bytes of 0/1), n_mv, n_flipped, bytes[, md5, payload_md5]} of this rank's
    GOPs (md5 = digest of the GOP's NAL stream, payload_md5 = digest of its message + stego bits; hex strings).
    Returns on rank 0 the records of all ranks sorted by GOP index (None elsewhere)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sorted(local, key=lambda r: r["gop"])
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    # fixed-size header per GOP + one flat payload tensor: two collectives, no pickling on the data path
    hdr = torch.tensor([[r["gop"], r["n_bits"], r["n_mv"], r["n_flipped"], r["bytes"]] + _md5_to_ints(r.get("md5")) +
                        _md5_to_ints(r.get("payload_md5")) for r in local] or [[-1, 0, 0, 0, 0, 0, 0, 0, 0]],
                       dtype=torch.int64, device=dev)
    pay = torch.tensor(list(b"".join(bytes(r["payload"]) for r in local)) or [0], dtype=torch.uint8, device=dev)
    sizes = torch.tensor([hdr.shape[0], pay.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_h = int(max(s[0] for s in all_sizes)); max_p = int(max(s[1] for s in all_sizes))
    hdr_pad = torch.full((max_h, 9), -1, dtype=torch.int64, device=dev); hdr_pad[:hdr.shape[0]] = hdr
    pay_pad = torch.zeros(max_p, dtype=torch.uint8, device=dev); pay_pad[:pay.numel()] = pay
    hdrs = [torch.zeros_like(hdr_pad) for _ in range(world)]
    pays = [torch.zeros_like(pay_pad) for _ in range(world)]
    dist.all_gather(hdrs, hdr_pad, group=group)
    dist.all_gather(pays, pay_pad, group=group)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        pos = 0
        raw = bytes(pays[r].cpu().tolist())
        for row in hdrs[r].cpu().tolist():
            if row[0] < 0:
                continue
            out.append({"gop": row[0], "n_bits": row[1], "n_mv": row[2], "n_flipped": row[3], "bytes": row[4],
                        "md5": _ints_to_md5(row[5], row[6]), "payload_md5": _ints_to_md5(row[7], row[8]),
                        "payload": raw[pos:pos + row[1]], "rank": r})
            pos += row[1]
    return sorted(out, key=lambda r: r["gop"])


def max_over_ranks(seconds, group=None):
    """Timing rule of bench.py: the slowest rank's time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(s) for s in seconds]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(seconds), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(values, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(v) for v in values]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.tolist()]
