"""CPU, world_size 2 (gloo): the host side of the multi-GPU path — candidate sharding with global
ids, the one all_gather of per-candidate scores, the shared seed broadcast — behaves the same on
every rank and for ragged shards.  The denoising loop itself has no inter-rank traffic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from its_b200.search import search_algorithm as S
        lo, hi = S._shard(n, rank, world)
        # this rank "scores" its own block: score of global candidate i is a fixed function of i
        all_scores = torch.tensor([((i * 37) % 11) / 10.0 for i in range(n)], dtype=torch.float32)
        local = all_scores[lo:hi].clone()
        gathered = S._gather_scores(local, n)
        seed = S._shared_seed(1000 + rank if rank == 0 else None, torch.device("cpu"))
        # strict '>' / first index selection on the gathered list (search_algorithm.py:79-81)
        best, best_i = float("-inf"), -1
        for i, v in enumerate(gathered.tolist()):
            if v > best:
                best, best_i = v, i
        torch.save({"gathered": gathered, "ref": all_scores, "seed": seed, "best": best_i, "shard": (lo, hi)},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7, 1])
def test_score_gather_and_selection_agree_on_every_rank(tmp_path, n):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    # shards tile [0, n) with global ids, ragged tail on the last rank
    assert res[0]["shard"][0] == 0 and res[-1]["shard"][1] == n
    assert res[0]["shard"][1] == res[1]["shard"][0]
    for r in res:
        assert torch.equal(r["gathered"], r["ref"])          # every rank sees every score, in candidate order
        assert r["seed"] == res[0]["seed"] == 1000           # rank 0's seed wins
        assert r["best"] == res[0]["best"]
    ref = res[0]["ref"]
    assert res[0]["best"] == int(torch.nonzero(ref == ref.max())[0])   # first index of the maximum


def test_shard_is_rank_count_invariant():
    from its_b200.search import search_algorithm as S
    for n in (1, 5, 64, 1000):
        for world in (1, 2, 3, 8):
            ids = []
            for r in range(world):
                lo, hi = S._shard(n, r, world)
                ids += list(range(lo, hi))
            assert ids == list(range(n))
