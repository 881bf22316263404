"""Multi-GPU correctness where a 2+-GPU box can see it: spawns two ranks (one process per GPU, torchrun)
running tools/multigpu_check.py — every rank checks its merged, sharded result against the oracle's scan
of the WHOLE corpus for both exchanges (stores into peer memory / ncclAllGather), for the fp32 scan
(bit-exact), the split filter + rescoring path (bit-exact) and the bf16 tensor path (tolerance).
Skipped, with the reason, on a one-GPU box; bench.py's `parity` block covers every N of the scaling run."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_rank_sharded_search_matches_the_oracle(pcv_lib, tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs 2 GPUs for one process per GPU, this box shows {n} (two ranks spinning on one GPU "
                    "would deadlock the peer exchange); bench.py --gpus N checks parity at every N")
    env = dict(os.environ, CHECK_WATCHDOG="240")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "multigpu_check_pytest.log").write_text(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for what in ("fp32 sharded scan (nccl exchange)", "fp32 sharded scan (p2p exchange)", "split sharded", "bf16 sharded tcgen05"):
        assert r.stdout.count(what) == 2, (what, r.stdout[-3000:])


def test_one_handle_over_many_gpus_matches_one_gpu(pcv_lib, orc):
    """pcv_index_create_multi: one handle, one row-range shard per GPU of THIS process, exchange over peer
    access.  Every path (fp32 scan, split filter + rescoring, bf16 tensor) returns what one GPU returns —
    bit for bit, the merge being order-independent — and the row-level calls address the shards as one matrix."""
    import numpy as np
    import torch

    import perceive_b200 as pb
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip(f"needs 2 GPUs in one process, this box shows {n_gpu}")
    devs = list(range(min(n_gpu, 4)))
    n, dim, k = 120_000, 384, 10
    rows = orc.synth_rows(1, 0, 0, n, dim)
    ids = np.arange(1, n + 1, dtype=np.int64)
    src = (ids % 3).astype(np.int64)
    qs = orc.synth_rows(2, 0, 0, 40, dim)
    for store in (pb.PCV_F32, pb.PCV_F32_SPLIT, pb.PCV_BF16):
        with pb.Index(dim, device=0, store=store) as one, pb.Index(dim, store=store, devices=devs) as many:
            one.set_rows(rows, ids, src)
            many.set_rows(rows, ids, src)
            st = many.stats()
            assert st.world == len(devs) and st.n_rows == n
            for q, flt in ((qs[:1], None), (qs[:5], [0, 2]), (qs, None), (qs, [1])):
                a, b = one.search(q, k, sources=flt), many.search(q, k, sources=flt)
                for x, y, name in zip(a, b, ("ids", "scores", "sims", "counts")):
                    assert np.array_equal(x, y), (store, name, flt)
            assert many.find_id(77) is not None and np.array_equal(many.embedding_of(77), one.embedding_of(77))
            # rebuild one source: its rows leave every shard, the new ones are dealt out again
            new_rows = orc.synth_rows(9, 0, 0, 9_000, dim)
            new_ids = np.arange(500_001, 509_001, dtype=np.int64)
            one.replace_source(1, new_rows, new_ids)
            many.replace_source(1, new_rows, new_ids)
            a, b = one.search(qs, k), many.search(qs, k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
            many.set_hidden([int(b[0][0][0])])
            one.set_hidden([int(a[0][0][0])])
            a, b = one.search(qs[:2], k), many.search(qs[:2], k)
            assert np.array_equal(a[0], b[0]) and int(a[0][0][0]) != int(b[0][0][1])
    # rejected: the same device twice (two exchange kernels waiting on one another must not share a GPU)
    with pytest.raises(pb.PcvError) as e:
        pb.Index(dim, devices=[0, 0])
    assert e.value.code == 1
