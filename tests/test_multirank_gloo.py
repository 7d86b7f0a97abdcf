"""N>1 host logic on CPU (world_size 2, gloo): validators sharded by rank, per-rank direct vote weights summed with
ONE all_reduce of int64[n_blocks], head computed on every rank from the reduced vector -- must equal the unsharded
result for any partition.  The GPU kernels are replaced by the numpy oracle behind the same
EpochProcessor plumbing (engine interface duck-typed), so what is tested is the sharding + collective, not the math."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenarios
from oracle import fast

N_VAL, N_BLK = 4096, 300


class _OracleEngine:
    """Duck-typed stand-in for engine.Engine: votes/head by the numpy oracle, BLS stages accept everything."""
    device = 0

    def __init__(self, tree, shard):
        self.parent, self.roots, self.leaf_viable = tree
        self.shard = shard

    def aggregate_dev(self, *a):
        pass

    def fast_aggregate_verify_dev(self, d_members, d_off, d_bits, d_msgs, d_sigs, d_ok):
        d_ok.fill_(1)

    def latest_messages_update_dev(self, *a):
        pass

    def vote_weights_dev(self, d_votes):
        msg_block, has_msg, equiv, active, eff = self.shard
        w = np.zeros(N_BLK, dtype=np.uint64)
        m = (has_msg != 0) & (active != 0) & (equiv == 0)
        np.add.at(w, msg_block[m].astype(np.int64), eff[m])
        d_votes.copy_(torch.from_numpy(w.astype(np.int64)))

    def head_from_votes_dev(self, d_votes, d_head, justified_idx=0, boost_idx=-1, boost_score=0, d_weight=None):
        w = d_votes.numpy().astype(np.uint64).copy()
        if boost_idx >= 0:
            w[boost_idx] += np.uint64(boost_score)
        for b in range(N_BLK - 1, 0, -1):
            w[self.parent[b]] += w[b]
        d_head[0] = fast.ghost_head(self.parent, self.roots, fast.ghost_viable(self.parent, self.leaf_viable), w, justified_idx)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pos_evolution_b200.epoch import EpochProcessor
    parent, slot, roots, leaf_viable = scenarios.fork_tree(N_BLK, 8)
    votes = scenarios.votes(N_VAL, N_BLK, 8)
    lo, hi = rank * N_VAL // world, (rank + 1) * N_VAL // world
    shard = tuple(v[lo:hi] for v in votes)
    ep = EpochProcessor(_OracleEngine((parent, roots, leaf_viable), shard), n_agg=4, n_sig=8, bits_stride=1, n_blocks=N_BLK,
                        process_group=dist.group.WORLD, device=torch.device("cpu"))
    ep.set_committees(np.arange(8, dtype=np.uint32), np.array([0, 2, 4, 6, 8], dtype=np.uint32))
    z = torch.zeros(1)
    ok, head = ep.process_epoch_dev(z, z, z, z, z, 0, N_BLK - 1, 12345)
    q.put((rank, int(head[0]), int(ok.sum())))
    dist.destroy_process_group()


def test_two_rank_sharding_equals_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parent, slot, roots, leaf_viable = scenarios.fork_tree(N_BLK, 8)
    msg_block, has_msg, equiv, active, eff = scenarios.votes(N_VAL, N_BLK, 8)
    w = fast.ghost_weights(parent, msg_block, has_msg, eff, active, equiv, N_BLK - 1, 12345)
    expect = fast.ghost_head(parent, roots, fast.ghost_viable(parent, leaf_viable), w, 0)
    assert [r[1] for r in res] == [expect, expect]
    assert [r[2] for r in res] == [4, 4]
