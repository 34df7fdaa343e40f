"""CPU, world_size 2 over gloo: the hot path shards by utterance with no exchange -- the
concatenation of per-rank results equals the single-process result -- and the only collective
(gradient all-reduce) and the max-over-ranks timing helper behave."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simulst_b200 import sharding


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import mma as omma
        torch.set_num_threads(1)
        g = torch.Generator().manual_seed(42)        # every rank builds the same full batch
        bsz, heads, t, s = 5, 2, 6, 40               # 5 utterances over 2 ranks: 3 + 2
        p = torch.sigmoid(torch.randn(bsz * heads, t, s, generator=g) - 2)
        e = torch.randn(bsz * heads, t, s, generator=g)
        r0, r1 = sharding.row_shard(bsz, heads, world, rank)
        a_loc, b_loc = omma.mma_process_train(p[r0:r1], e[r0:r1], None, 1e-6, True, None)
        gathered = [None] * world
        dist.all_gather_object(gathered, (r0, r1, a_loc, b_loc))
        grad = torch.full((7,), float(rank + 1))
        h = sharding.allreduce_gradients(grad, async_op=True)
        h.wait()
        slowest = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            a_full, b_full = omma.mma_process_train(p, e, None, 1e-6, True, None)
            a_cat = torch.cat([x[2] for x in sorted(gathered, key=lambda x: x[0])])
            b_cat = torch.cat([x[3] for x in sorted(gathered, key=lambda x: x[0])])
            ok = (torch.equal(a_cat, a_full) and torch.equal(b_cat, b_full)
                  and torch.equal(grad, torch.full((7,), 3.0)) and slowest == 11.0
                  and [(x[0], x[1]) for x in sorted(gathered, key=lambda x: x[0])] == [(0, 6), (6, 10)])
            with open(os.path.join(out_dir, "ok"), "w") as f:
                f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_utterance_shard_partition():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [sharding.utterance_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.row_shard(64, 8, 8, 3) == (192, 256)


def test_two_rank_sharding_matches_single_process(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1"
