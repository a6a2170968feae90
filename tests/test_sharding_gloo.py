"""CPU, world_size 2 over gloo: the N>1 path of the benchmark / deployment — frame
sharding, the barrier + MAX-reduce timing protocol, and gathering shard results —
shards are independent, so the concatenation of per-rank results must equal the
unsharded result bit for bit (computed here with the CPU oracle standing in for
the per-rank sweep)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simplerecon_b200 import sharding
from simplerecon_b200.synthetic import make_tuple


def test_shard_range_partitions_exactly():
    for total in (1, 4, 7, 64):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 4)
    # BASELINE configs: 16 frames over 4 GPUs, 64 over 8
    assert sharding.shard_range(16, 3, 4) == (12, 16) and sharding.shard_range(64, 5, 8) == (40, 48)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import costvolume_oracle as O
    r, w, _ = sharding.init_distributed("gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    full = make_tuple(4, 2, 12, 16, seed=42)
    mine = sharding.shard_tuple(full, rank, world)
    assert mine["src_feats"].shape[0] == 2 and mine["min_depth"].shape == (1, 1, 1, 1)
    sharding.barrier()
    cost, lowest, _, _ = O.forward_dot(**mine, num_depth_bins=5)
    sharding.barrier()
    # timing protocol: every rank contributes its own elapsed time, all see the max
    t = sharding.max_over_ranks(10.0 + rank)
    assert t == 10.0 + (world - 1)
    assert sharding.sum_over_ranks(1.0) == float(world)
    gathered = sharding.gather_frames(cost)
    if rank == 0:
        ref, _, _, _ = O.forward_dot(**full, num_depth_bins=5)
        torch.save({"equal": bool(torch.equal(gathered, ref)), "shape": tuple(gathered.shape)},
                   os.path.join(out_dir, "result.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reassemble(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = torch.load(tmp_path / "result.pt")
    assert res["equal"] and res["shape"] == (4, 5, 12, 16)
