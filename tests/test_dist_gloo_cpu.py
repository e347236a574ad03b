"""CPU, world_size 2 over gloo: the N > 1 host logic of the path -- contig sharding covers every
record exactly once and the single collective (SUM all-reduce of the counter block) gives the
global pass/fail counts."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from variantcalling_b200 import dist as vdist
from variantcalling_b200.synth import CONTIG_LENGTHS


def test_lpt_partition_balances_contigs():
    total = 50_000_000
    ranges = vdist.contig_record_ranges(total, CONTIG_LENGTHS)
    assert ranges["chr1"][0] == 0 and ranges["chrY"][1] == total
    spans = sorted(ranges.values())
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    loads = {c: r1 - r0 for c, (r0, r1) in ranges.items()}
    for world in (1, 2, 4, 8):
        bins = vdist.lpt_partition(loads, world)
        assert sorted(c for b in bins for c in b) == sorted(loads)
        per = [sum(loads[c] for c in b) for b in bins]
        assert max(per) <= 1.10 * total / world, (world, per)


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 100_000
    ranges = vdist.contig_record_ranges(total, CONTIG_LENGTHS)
    loads = {c: r1 - r0 for c, (r0, r1) in ranges.items()}
    mine = vdist.lpt_partition(loads, world)[rank]
    # stand-in for the GPU pass: a deterministic per-record decision, counted per rank
    n = low = 0
    for c in mine:
        r0, r1 = ranges[c]
        idx = np.arange(r0, r1)
        n += idx.size
        low += int(np.count_nonzero((idx * 2654435761) % 7 < 2))
    counts = torch.tensor([n, low, n - low, 0], dtype=torch.int64)
    vdist.allreduce_counts(counts)
    out_q.put((rank, counts.tolist(), sorted(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_counts_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(100_000)
    low = int(np.count_nonzero((idx * 2654435761) % 7 < 2))
    assert res[0][1] == res[1][1] == [100_000, low, 100_000 - low, 0]
    assert sorted(res[0][2] + res[1][2]) == sorted(CONTIG_LENGTHS)
    assert not set(res[0][2]) & set(res[1][2])
