"""N>1 host logic on CPU: two gloo ranks run the sharded-search protocol of
SURVEY.md 8e — local top-k on a row-range shard, all-gather of the (sim,id)
candidates, merge with the stated tie-break — with the ORACLE standing in for
the device scan, and must reproduce the single-shard oracle result exactly
(shard-count invariance).  Also covers shard_rows and the unique-id broadcast."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _merge(sims, ids, k):
    """[G,k] candidate lists (padded with -inf / INT64_MAX) -> best k by (sim desc, id asc):
    what merge_candidates_kernel computes (perceive_b200/csrc/pcv_load.cuh)."""
    s, i = sims.reshape(-1), ids.reshape(-1)
    live = i != np.iinfo(np.int64).max
    s, i = s[live], i[live]
    order = np.lexsort((i, -s))[:k]
    return i[order], s[order]


def _worker(rank, world, port, n, dim, k, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from perceive_b200.distributed import shard_rows
        r0, r1 = shard_rows(n, rank, world)
        rows = orc.synth_rows(1, 0, r0, r1 - r0, dim)  # this rank's shard, global ids
        rows[0] = orc.synth_rows(1, 0, 0, 1, dim)[0]   # a duplicate of global row 0 on every shard: ties across shards
        ids = np.arange(r0 + 1, r1 + 1, dtype=np.int64)
        q = orc.synth_rows(2, 0, 0, 1, dim)[0]
        l_ids, _, l_sims = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1)
        pad = k - len(l_ids)
        send_s = torch.from_numpy(np.concatenate([l_sims.astype(np.float32), np.full(pad, -np.inf, np.float32)]))
        send_i = torch.from_numpy(np.concatenate([l_ids, np.full(pad, np.iinfo(np.int64).max, np.int64)]))
        gs = [torch.empty_like(send_s) for _ in range(world)]
        gi = [torch.empty_like(send_i) for _ in range(world)]
        dist.all_gather(gs, send_s)
        dist.all_gather(gi, send_i)
        m_ids, m_sims = _merge(torch.stack(gs).numpy(), torch.stack(gi).numpy(), k)
        # a broadcast of 128 opaque bytes from rank 0: the unique-id hand-off (no NCCL needed to test it)
        token = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(token, 0)
        assert token.tolist() == list(range(128))
        np.save(os.path.join(out, f"ids{rank}.npy"), m_ids)
        np.save(os.path.join(out, f"sims{rank}.npy"), m_sims)
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    from perceive_b200.distributed import shard_rows
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_rows(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_sharded_search_matches_single_shard(orc, tmp_path, world):
    import torch.multiprocessing as mp
    n, dim, k = 5000, 96, 10
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, dim, k, str(tmp_path)), nprocs=world, join=True)
    # single-shard truth over the same logical corpus
    rows = orc.synth_rows(1, 0, 0, n, dim)
    from perceive_b200.distributed import shard_rows
    for r in range(world):
        rows[shard_rows(n, r, world)[0]] = rows[0]
    ids = np.arange(1, n + 1, dtype=np.int64)
    q = orc.synth_rows(2, 0, 0, 1, dim)[0]
    w_ids, _, w_sims = orc.search(rows, ids, q, k, mode=orc.MODE_F32_V1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"ids{r}.npy"), w_ids)
        assert np.array_equal(np.load(tmp_path / f"sims{r}.npy"), w_sims.astype(np.float32))
