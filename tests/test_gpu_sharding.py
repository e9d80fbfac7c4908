"""Multi-GPU rows of SURVEY.md 8(e) on hardware:

* time-block sharding of the wideband stream (cfg 5): every rank's block range, started from a halo re-read from the
  source, must be bit-identical to the same blocks of the unsharded run - checked on one GPU by running the ranks'
  ranges one after the other, and against the CPU checker's block-by-block chain;
* result collection over NCCL: two ranks (torch.distributed.run, backend nccl) detect+demodulate their shards,
  all-gather the per-burst records and all-reduce the counters; the gathered arrays must equal a single-rank run
  (skipped with fewer than two GPUs).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def beq(a, b):
    return bool(torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)))


@pytest.mark.parametrize("m,world", [(64, 2), (64, 8), (4, 3)])
def test_wideband_time_block_sharding_bit_exact(trx, m, world):
    import osmo_trx_b200
    from osmo_trx_b200 import sharding
    BL, Q = 192, 125                       # 125 blocks = 52 slots of 625 samples per channel
    nblk = Q * 2 * world + Q               # uneven split on purpose
    g = torch.Generator(device=trx.device)
    g.manual_seed(5)
    wide = torch.randn((nblk * BL * m, 2), generator=g, device=trx.device)
    full = osmo_trx_b200.WidebandRx(trx, m, BL).rotate(wide).clone()
    covered = 0
    for rank in range(world):
        b0, b1, halo = sharding.shard_time_blocks(nblk, rank, world, Q)
        assert (b1 - b0) % Q == 0 or rank == world - 1
        rx = osmo_trx_b200.WidebandRx(trx, m, BL)
        if halo:
            rx.prime(wide[(b0 * BL - halo) * m: b0 * BL * m])
        part = rx.rotate(wide[b0 * BL * m: b1 * BL * m])
        torch.cuda.synchronize()
        assert beq(part, full[:, b0 * 260: b1 * 260]), f"rank {rank} of {world}: blocks {b0}..{b1} differ from the unsharded run"
        covered += b1 - b0
    assert covered == nblk


def test_wideband_rx_matches_reference_chain(trx, checker):
    """WidebandRx (many blocks per call, channel rows behind carried history, one long resampler stream per channel)
    against the CPU checker's block-by-block Channelizer::rotate + Resampler::rotate; two calls, so the carried
    state (channelizer history and resampler tail) is exercised as well."""
    import osmo_trx_b200
    if not hasattr(checker, "wideband_rx") or checker.pfx != "ref_":
        pytest.skip("the block-by-block wideband chain is wrapped for the compiled reference only")
    m, BL, nblk = 64, 192, 24
    rng = np.random.default_rng(3)
    wide = rng.standard_normal((nblk * BL * m, 2)).astype(np.float32)
    want = checker.wideband_rx(wide, nblk, m, BL, 65, 48)
    rx = osmo_trx_b200.WidebandRx(trx, m, BL)
    w = torch.from_numpy(wide).cuda()
    half = (nblk // 2) * BL * m
    got = torch.cat([rx.rotate(w[:half]).clone(), rx.rotate(w[half:]).clone()], dim=1).cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err <= 1e-4, err   # the DFT's parity bar (FFTW unpinned in the reference); the resampler itself is exact


NCCL_WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import bench, osmo_trx_b200
from osmo_trx_b200 import sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
trx = osmo_trx_b200.Trx(local)
n = 6000 + 1
# every rank generates the same global batch (seeded), then keeps its shard
rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=77, device=dev)
trx.detect_config(16, 1)
lo, hi = sharding.shard_range(n, rank, world)
res = trx.detect_demod(rx[lo:hi].contiguous(), typ[lo:hi].contiguous(), tsc[lo:hi].contiguous(), mt[lo:hi].contiguous(), bound, n_gmsk_soft=148)
torch.cuda.synchronize()
allr = sharding.gather_results(res, n)
cnt = sharding.reduce_counters(sharding.counters_device(res))
whole = trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148)
torch.cuda.synchronize()
ok = True
for k in ("rc", "amp", "toa", "tsc", "ci", "flags", "soft"):
    a, b = allr[k], whole[k]
    if a.dtype == torch.float32:
        a, b = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    ok = ok and bool(torch.equal(a, b))
want = sharding.counters(whole)
ok = ok and bool(torch.equal(cnt.cpu(), want.cpu()))
# the packed form: results written straight into one record block per rank, gathered by a single ncclAllGather
m = (n // world) // 4 * 4
outp, rec = sharding.alloc_packed_results(m, 148, dev)
s0 = rank * m
trx.detect_demod(rx[s0:s0 + m].contiguous(), typ[s0:s0 + m].contiguous(), tsc[s0:s0 + m].contiguous(), mt[s0:s0 + m].contiguous(), bound,
                 n_gmsk_soft=148, out=outp)
allp = sharding.unpack_records(sharding.gather_records(rec), m, world)
torch.cuda.synchronize()
for k in ("rc", "amp", "toa", "tsc", "ci", "flags"):
    a, b = allp[k], whole[k][:world * m]
    if a.dtype == torch.float32:
        a, b = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    ok = ok and bool(torch.equal(a, b))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({{"ok": bool(flag.item()), "world": world, "backend": dist.get_backend(), "counters": cnt.tolist()}}))
dist.destroy_process_group()
"""


def test_nccl_gather_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER.format(root=ROOT))
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["world"] == 2 and rep["backend"] == "nccl", rep
