"""Multi-GPU sharding of the burst batch (one process per GPU, torch.distributed).

Bursts (ARFCN x timeslot x frame) are independent and every rank holds its own copy of the tables
(SURVEY.md §8(e)), so the data path needs no collective: each rank processes a contiguous slice of the
burst index space.  The only exchange is result collection - per-burst records gathered to one rank and
counters summed - which maps to NCCL all_gather / all_reduce over NVLink on GPUs and to gloo in the CPU
tests.  This module is backend-agnostic plumbing; it never touches sample data.
"""
import torch
import torch.distributed as dist

COUNTER_NAMES = ("bursts", "detected", "clipped", "thresh_edge", "bisect_tie", "errors")


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n bursts for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_by_arfcn(n_arfcn, rank, world):
    """ARFCN-major partition: whole carriers per rank (all 8 timeslots of a carrier stay together)."""
    return shard_range(n_arfcn, rank, world)


def shard_time_blocks(n_blocks, rank, world, quantum=1):
    """Time-block partition of one wideband stream (cfg 5): contiguous block range [b0, b1) for
    `rank`, every boundary a multiple of `quantum` blocks (125 blocks of 192 wideband rows are exactly 52 slots of 625
    samples per channel at 65/48, so whole bursts stay on one rank).  Returns (b0, b1, halo_rows): the number of wideband
    rows in front of b0 the rank re-reads from the source to rebuild the channelizer and resampler histories (0 for the
    rank that starts the stream)."""
    q_total = n_blocks // quantum
    lo, hi = shard_range(q_total, rank, world)
    b0, b1 = lo * quantum, hi * quantum
    if rank == world - 1:
        b1 = n_blocks
    return b0, b1, (32 if b0 > 0 else 0)


def counters(res):
    """Per-shard statistics tensor (int64[6]) from a result dict (rc int32[n], flags uint8[n])."""
    rc, fl = res["rc"], res["flags"]
    vals = [rc.numel(), int((rc > 0).sum()), int((rc == -2).sum()), int(((fl & 1) != 0).sum()),
            int(((fl & 2) != 0).sum()), int(((rc < 0) & (rc != -2)).sum())]
    return torch.tensor(vals, dtype=torch.int64, device=rc.device)


def counters_device(res):
    """Same as counters() but without host synchronisation (stays on the device stream)."""
    rc, fl = res["rc"], res["flags"]
    return torch.stack([torch.tensor(rc.numel(), device=rc.device), (rc > 0).sum(), (rc == -2).sum(),
                        ((fl & 1) != 0).sum(), ((fl & 2) != 0).sum(), ((rc < 0) & (rc != -2)).sum()]).to(torch.int64)


def reduce_counters(c, group=None):
    """Sum the statistics over all ranks (ncclAllReduce on GPUs)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return c


def gather_results(res, n_total, group=None, keys=("rc", "amp", "toa", "tsc", "ci", "flags", "soft")):
    """All-gather per-burst result records into global burst order.

    `res[k]` is this rank's slice ([n_local, ...]); shards follow shard_range(n_total, rank, world).
    Returns a dict of tensors of length n_total on every rank.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {k: res[k] for k in keys}
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    maxn = max(sizes)
    out = {}
    for k in keys:
        t = res[k]
        pad_shape = (maxn,) + tuple(t.shape[1:])
        buf = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        out[k] = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return out


RECORD_BYTES = 22  # rc i32 | amp f32 x 2 | toa f32 | ci f32 | tsc u8 | flags u8


def alloc_packed_results(n, soft_stride=148, device="cuda"):
    """Result arrays for Trx.detect_demod(out=...) carved out of ONE contiguous buffer (SoA blocks back to back:
    rc[n] | amp[n][2] | toa[n] | ci[n] | tsc[n] | flags[n], 22 bytes per burst), so that a rank's per-burst records
    travel in a single collective (`records`, uint8[22 n]).  n must be a multiple of 4 (alignment of the float blocks).
    The soft-bit rows are a separate tensor: they stay with the rank that produced them."""
    assert n % 4 == 0
    rec = torch.zeros(RECORD_BYTES * n, dtype=torch.uint8, device=device)
    out = dict(rc=rec[0:4 * n].view(torch.int32), amp=rec[4 * n:12 * n].view(torch.float32).view(n, 2),
               toa=rec[12 * n:16 * n].view(torch.float32), ci=rec[16 * n:20 * n].view(torch.float32),
               tsc=rec[20 * n:21 * n], flags=rec[21 * n:22 * n],
               soft=torch.zeros((n, soft_stride), dtype=torch.float32, device=device))
    return out, rec


def unpack_records(rec_all, n, world):
    """[world, 22 n] gathered record blocks -> dict of [world * n, ...] arrays in global burst order"""
    r = rec_all.view(world, RECORD_BYTES * n)
    return dict(rc=r[:, 0:4 * n].contiguous().view(torch.int32).view(-1),
                amp=r[:, 4 * n:12 * n].contiguous().view(torch.float32).view(-1, 2),
                toa=r[:, 12 * n:16 * n].contiguous().view(torch.float32).view(-1),
                ci=r[:, 16 * n:20 * n].contiguous().view(torch.float32).view(-1),
                tsc=r[:, 20 * n:21 * n].contiguous().view(-1), flags=r[:, 21 * n:22 * n].contiguous().view(-1))


def gather_records(rec, group=None):
    """One all_gather of a rank's packed record block (ncclAllGather on GPUs): uint8[22 n] -> uint8[world, 22 n]"""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    out = torch.empty((world, rec.numel()), dtype=torch.uint8, device=rec.device)
    if world == 1:
        out[0] = rec
    else:
        dist.all_gather_into_tensor(out.view(-1), rec, group=group)
    return out

