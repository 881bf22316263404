"""Host-side plumbing for row-range shards, one process per GPU.

The reference has no multi-process path (single process, rayon
over sources, search.rs:163-177); its per-source fan-out + concat + sort
(search.rs:163-181) becomes per-SHARD local top-k + all-gather + merge here.
`torch.distributed` is used only to hand the 128-byte NCCL unique id from rank 0
to the other ranks; the data path (ncclAllGather of the candidates and the merge
kernel) runs inside libperceive_cuda on the index's own stream.
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [r0, r1) of the (source,id)-ordered matrix owned by `rank`: contiguous,
    balanced to within one row, covering [0, n_rows) exactly once over all ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    return n_rows * rank // world, n_rows * (rank + 1) // world


def exchange_unique_id(dist, rank: int, device=None) -> bytes:
    """Rank 0 creates the NCCL unique id (pcv_comm_unique_id); every rank returns it.
    Works on any torch.distributed backend (gloo on CPU in the tests, nccl on GPUs)."""
    import torch

    from .searcher import comm_unique_id
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def exchange_ipc_handles(dist, mine: bytes, world: int, device=None) -> bytes:
    """All-gather of the 64-byte CUDA IPC handles; returns them in rank order."""
    import torch
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    allh = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(allh, t)
    return b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh)


def attach_shard(index, dist, rank: int, world: int, device=None, exchange: str = "p2p",
                 max_records: int = 1 << 16) -> str:
    """Make `index` shard `rank` of `world`: after this every search on it is collective.
    exchange = "nccl": ncclAllGather + merge kernel; "p2p": candidates stored straight into peer
    memory over NVLink, merged in the same launch (NCCL stays attached as the fallback for
    searches whose n_queries * k exceeds max_records).  Returns the exchange actually in use."""
    if world == 1:
        return "none"
    if exchange not in ("nccl", "p2p"):
        raise ValueError(f"unknown exchange {exchange!r}")
    index.attach_comm(exchange_unique_id(dist, rank, device), rank, world)
    if exchange == "p2p":
        # mapping peer memory can fail on one rank only (no IPC between the processes, no P2P path):
        # the choice of exchange must be unanimous, so the ranks vote and fall back to NCCL together
        import torch
        ok = 1
        try:
            if os.environ.get("PCV_P2P_FAIL_RANK") == str(rank):  # test hook: pretend this rank cannot map peers
                raise RuntimeError("simulated peer-mapping failure")
            mine = index.p2p_export(world, max_records)
        except Exception:  # noqa: BLE001 - any failure means "not this way"
            mine, ok = bytes(64), 0
        handles = exchange_ipc_handles(dist, mine, world, device)  # every rank takes part, whatever happened
        if ok:
            try:
                index.p2p_attach(handles, rank, world)
            except Exception:  # noqa: BLE001
                ok = 0
        vote = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if int(vote.item()) == 0:
            if ok:
                index.p2p_detach()  # somebody else could not map: nobody uses the peer path
            return "nccl"
        return "p2p"
    return "nccl"
