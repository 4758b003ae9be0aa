"""N > 1 path on CPU: two `gloo` ranks each own a contiguous column range (microclimc_b200.shard), advance their
columns independently (no data-path collective) and rank 0 gathers; the result must equal the single-process run
bit for bit.  The per-rank compute here is the host (g++) build of the device column code (tests/hostdbg), standing in
for the GPU that each rank drives in bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ncol, T, outfile):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from oracle import pyoracle
    from tests.common import make_cfg, transient_inputs
    from tests.test_host_logic import dbg_run_transient
    from microclimc_b200 import shard
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostdbg", "libdbg_host.so"))
    inp = transient_inputs(pyoracle, ncol=ncol, T=T)
    c0, c1 = shard.column_range(ncol, world, rank)
    mine = dict(inp, **{k: shard.slice_columns(inp[k], c0, c1) for k in ("veg", "soil", "lat", "lon", "fs", "fo")})
    r = dbg_run_transient(lib, mine, make_cfg(spin_steps=20))
    state = shard.gather_columns(r["state"], ncol)
    stats = shard.gather_columns(r["stats"], ncol)
    if rank == 0:
        np.savez(outfile, state=state, stats=stats)
    dist.barrier()
    dist.destroy_process_group()


def test_column_range_is_a_partition():
    from microclimc_b200 import shard
    for ncol in (1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.column_range(ncol, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == ncol
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1


def test_two_rank_gloo_run_equals_single_process(oracle, hostdbg, tmp_path):
    from tests.common import make_cfg, transient_inputs
    from tests.test_host_logic import dbg_run_transient
    ncol, T = 11, 24                      # odd: the tail rank owns fewer columns
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), ncol, T, out), nprocs=2, join=True)
    got = np.load(out)
    inp = transient_inputs(oracle, ncol=ncol, T=T)
    want = dbg_run_transient(hostdbg, inp, make_cfg(spin_steps=20))
    np.testing.assert_array_equal(got["state"], want["state"])
    np.testing.assert_array_equal(got["stats"], want["stats"])
