"""Multi-GPU parity check, one process per GPU (run under torchrun on a B200 box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py

Every rank holds a row-range shard of one synthetic corpus (global ids), attaches
the NCCL communicator, and searches collectively (local top-k -> ncclAllGather ->
merge kernel).  Every rank then checks its merged result against the oracle's scan of
the WHOLE corpus: fp32 scan path and the split filter + rescoring path bit-exact, bf16
tensor-core path within the K2 tolerance.  Prints one OK line per rank; any mismatch raises."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import perceive_b200 as pb  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from perceive_b200.distributed import attach_shard, shard_rows  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    orc.build()

    # ---- fp32 scan path (K1 + K5), bit-exact --------------------------------
    n, dim, k = 300_000, 384, 10
    r0, r1 = shard_rows(n, rank, world)
    full = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = orc.synth_rows(2, 0, 0, 6, dim)
    for exchange in ("nccl", "p2p"):
        with pb.Index(dim, device=local) as ix:
            ix.generate_synthetic(r1 - r0, 1, first_row=r0)
            attach_shard(ix, dist, rank, world, device=dev, exchange=exchange, max_records=64)
            st = ix.stats()
            assert st.world == world and st.rank == rank
            for rep in range(3):  # several epochs through both buffer parities
                got = ix.search(qs, k)  # 6 x 10 = 60 records: fits the 64-record peer buffers
                for b in range(qs.shape[0]):
                    w_ids, w_scores, w_sims = orc.search(full, ids, qs[b], k, mode=orc.MODE_F32_V1)
                    assert np.array_equal(got[0][b], w_ids), (exchange, rank, b, got[0][b], w_ids)
                    assert np.array_equal(got[2][b], w_sims.astype(np.float32)) and np.array_equal(got[1][b], w_scores)
                    assert int(got[3][b]) == k
            one = ix.search(qs[0], k)
            assert np.array_equal(one[0][0], got[0][0]) and np.array_equal(one[2][0], got[2][0])
            if exchange == "p2p":  # one query: the scan's last CTA carries the exchange — ONE launch in all
                assert ix.stats().last_launches == 1, ix.stats().last_launches
            big = ix.search(np.concatenate([qs, qs]), k)  # 120 records > 64: falls back to NCCL on both
            assert np.array_equal(big[0][:6], got[0]) and np.array_equal(big[0][6:], got[0])
            dist.barrier()
        print(f"rank {rank}/{world}: fp32 sharded scan ({exchange} exchange) == oracle over the whole corpus (bit-exact)", flush=True)

    # ---- fp32 rows as two 16-bit planes (K3: tensor-core filter + exact rescoring, + K5), bit-exact ------
    n, dim, k, nq = 200_000, 384, 10, 64
    r0, r1 = shard_rows(n, rank, world)
    full = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qsp = orc.synth_rows(2, 0, 0, nq, dim)
    for exchange in ("p2p", "nccl"):
        with pb.Index(dim, device=local, store=pb.PCV_F32_SPLIT) as ix:
            ix.generate_synthetic(r1 - r0, 1, first_row=r0)
            attach_shard(ix, dist, rank, world, device=dev, exchange=exchange, max_records=nq * k)
            got = ix.search(qsp, k)
            assert ix.stats().last_kernel == 2
            for b in range(0, nq, 7):
                w_ids, w_scores, w_sims = orc.search(full, ids, qsp[b], k, mode=orc.MODE_F32_V1)
                assert np.array_equal(got[0][b], w_ids), (exchange, rank, b, got[0][b], w_ids)
                assert np.array_equal(got[2][b], w_sims.astype(np.float32)) and np.array_equal(got[1][b], w_scores)
            dist.barrier()
    print(f"rank {rank}/{world}: split sharded filter + exact rescoring == oracle fp32 scan over the whole corpus (bit-exact)", flush=True)

    # ---- bf16 tensor-core path (K2 + K5), tolerance ---------------------------
    from test_gpu_gemm import check_batch
    n, dim, k, nq = 160_000, 384, 20, 96
    r0, r1 = shard_rows(n, rank, world)
    stored = orc.round_bf16(orc.synth_rows(1, 0, 0, n, dim))
    ids = np.arange(1, n + 1, dtype=np.int64)
    qb = orc.round_bf16(orc.synth_rows(2, 0, 0, nq, dim))
    with pb.Index(dim, device=local, store=pb.PCV_BF16) as ix:
        ix.generate_synthetic(r1 - r0, 1, first_row=r0)
        attach_shard(ix, dist, rank, world, device=dev, exchange="p2p", max_records=nq * k)
        res = ix.search(qb, k)
        assert ix.stats().last_kernel == 2
        err = check_batch(res, stored, ids, qb, k, what=f"rank {rank} sharded K2")
    print(f"rank {rank}/{world}: bf16 sharded tcgen05 search within tolerance (max err {err:.2e})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    if os.environ.get("CHECK_WATCHDOG"):  # debugging aid: dump every thread's stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["CHECK_WATCHDOG"]), exit=True)
    main()
