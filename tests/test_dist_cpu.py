"""world_size-2 gloo test of the N>1 host logic: sample sharding + the counter all-reduce.
On CPU the per-shard evaluator is the oracle (test infrastructure); on the GPU box the same
logic runs with the CUDA plan (tests/test_gpu_parity.py::test_sample_offset_and_sharding_invariance)."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_samples, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from qo100net import dist as qd
    from oracle import refbind as R
    d = qd.init_process_group("gloo")
    lad = R.ladder_lpf(R.cheby_g(11, 0.1), 10e6, 50.0, True, (60, 30, 0.1, 50))
    f = R.grid_log(4e6, 62.5e6, 129)
    tols = [(i, 0, i, 0, 0.05 if i % 2 == 0 else 0.02) for i in range(11)]
    specs = [(1, 0, 9.5e6, -2.0), (2, 13e6, 1e99, -49.0)]
    lo, hi = qd.shard_range(n_samples, rank, world)
    r = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(42, hi - lo, tols, sample_offset=lo, hist_bins=8, hist_spec=0, hist_lo=-4, hist_hi=0))
    cnt = torch.tensor([r["n_pass"], r["n_total"]] + [int(x) for x in r["fail_per_spec"]] + [int(x) for x in r["hist"]], dtype=torch.int64)
    qd.allreduce_counters(cnt)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), cnt.numpy())
    d.destroy_process_group()


def test_shard_range_partitions():
    from qo100net import dist as qd
    for n in (0, 1, 7, 1000, 10 ** 8):
        for w in (1, 2, 3, 4, 8):
            r = [qd.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_allreduce_equals_single_run(tmp_path, R):
    n = 101
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    lad = R.ladder_lpf(R.cheby_g(11, 0.1), 10e6, 50.0, True, (60, 30, 0.1, 50))
    f = R.grid_log(4e6, 62.5e6, 129)
    tols = [(i, 0, i, 0, 0.05 if i % 2 == 0 else 0.02) for i in range(11)]
    specs = [(1, 0, 9.5e6, -2.0), (2, 13e6, 1e99, -49.0)]
    r = R.mc_run(lad, 50, 50, f, specs, R.mc_cfg(42, n, tols, hist_bins=8, hist_spec=0, hist_lo=-4, hist_hi=0))
    one = [r["n_pass"], r["n_total"]] + [int(x) for x in r["fail_per_spec"]] + [int(x) for x in r["hist"]]
    assert a.tolist() == one and a[1] == n
