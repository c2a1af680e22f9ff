"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, "scan" their shard (the CPU oracle
stands in for the GPU scan here -- this test is about the sharding / reduce / gather plumbing), and must
reproduce the single-process result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from regex_fpga_b200 import shard, workloads as WL, MATCH_DTYPE
        from oracle import oracle_py as O
        z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
        E, n_states, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
        first, count = shard.shard_range(n_streams, rank, world)
        # each rank generates only its shard; the generator is counter based, so shards tile the full batch
        data = WL.make_batch_numpy("whi", lo, hi, count, 300, 320, first_stream=first)
        r = O.b_scan_many(E, n_states, data, count, 320, 300, n_threads=1)
        recs = r["recs"].astype(MATCH_DTYPE)
        recs["stream"] += first                                  # what stream_id_base does on the GPU
        counts = torch.from_numpy(r["counts"].astype(np.int64))
        shard.reduce_counts(counts, dist)
        allrecs = shard.gather_records(recs, dist)
        if rank == 0:
            q.put((counts.numpy(), allrecs))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    from regex_fpga_b200 import shard, workloads as WL
    from oracle import oracle_py as O
    n_streams = 101                                               # odd on purpose: uneven shards
    assert shard.shard_range(n_streams, 0, 2) == (0, 50) and shard.shard_range(n_streams, 1, 2) == (50, 51)
    assert sum(shard.shard_range(7, r, 4)[1] for r in range(4)) == 7
    z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
    E, n_states, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
    full = WL.make_batch_numpy("whi", lo, hi, n_streams, 300, 320)
    want = O.b_scan_many(E, n_states, full, n_streams, 320, 300, n_threads=2)
    assert want["n_recs"] > 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, recs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(counts.astype(np.uint64), want["counts"])
    assert recs.tolist() == want["recs"].tolist()
