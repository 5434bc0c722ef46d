"""N > 1 host logic on CPU: world_size-2 gloo, the oracle standing in for the GPU (tests may use it).
Checks that sharding + gather reproduces the single-process result bit for bit, and the timing reduce."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Oracle, SeedOpt
    sh, fm, sy = pkg("sharding"), pkg("fmindex"), pkg("synth")
    ref = sy.make_reference(150_000, 9)
    ix = fm.build_index(ref)
    seq, offs = sy.to_batch(sy.simulate_reads(ref, 1001, 101, 0.02, seed=3, n_frac=0.05))   # odd count: ragged shards
    lseq, loffs, (lo, hi) = sh.shard_batch(seq, offs, rank, world)
    assert (hi - lo) in (501, 500)
    local = Oracle(ix).collect(lseq, loffs, SeedOpt())
    merged = sh.gather_results(dict(intv=local["intv"], read_off=local["read_off"], step=local["step"]))
    t = sh.max_over_ranks(1.0 + rank)
    n = sh.sum_over_ranks(hi - lo)
    assert t == float(world) and n == 1001
    assert sh.gather_rows([rank, 2.5]) == [[0.0, 2.5], [1.0, 2.5]]
    if rank == 0:
        full = Oracle(ix).collect(seq, offs, SeedOpt())
        assert all(np.array_equal(merged[k], full[k]) for k in ("intv", "read_off", "step"))
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    sh = pkg("sharding")
    assert sh.shard_range(10, 0, 3) == (0, 4) and sh.shard_range(10, 2, 3) == (8, 10) and sh.shard_range(2, 3, 4) == (2, 2)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
