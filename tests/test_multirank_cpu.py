"""The N>1 path on CPU: world_size-2 gloo processes shard the spectra by precursor-mass range, run the (oracle) search on
their shard and gather the PSM records to rank 0 through pepquery_b200.shard — the same plumbing bench.py uses with NCCL.
The union over shards must equal the unsharded search (mass sharding is exact: every pair involves one spectrum)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pq_oracle
    from pepquery_b200 import _abi, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prot = synth.proteins(300); tgt = synth.targets(25); sp = synth.spectra(700, tgt, prot)
    mass = shard.precursor_mass(sp["prec_mz"], sp["charge"])
    bounds = shard.shard_bounds(mass, world)
    sub, gidx = shard.take_shard(sp, bounds, rank)
    p = _abi.default_params(n_random=100)
    o = pq_oracle.Oracle(p, threads=2)
    o.load_spectra(sub); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    ps = o.psms()
    ps["spectrum"] = gidx[ps["spectrum"]].astype(np.int32)       # local -> global spectrum index
    allp = shard.gather_records(ps, dist, rank, world)
    n_local = np.array([len(sub["prec_mz"])])
    if rank == 0:
        q.put((allp.tobytes(), str(allp.dtype.descr), bounds.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_mass_sharding_equals_unsharded():
    import torch.multiprocessing as mp
    from oracle import pq_oracle
    from pepquery_b200 import _abi, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    raw, descr, bounds = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.frombuffer(raw, dtype=pq_oracle.psm_dtype())
    got = np.sort(got, order=["spectrum", "form"])
    prot = synth.proteins(300); tgt = synth.targets(25); sp = synth.spectra(700, tgt, prot)
    o = pq_oracle.Oracle(_abi.default_params(n_random=100), threads=4)
    o.load_spectra(sp); o.set_targets(*tgt); o.step2(); o.build_reference(*prot); o.step3(); o.step4(); o.rank()
    want = o.psms()
    assert len(got) == len(want) > 0
    for f in ("spectrum", "form", "count_b", "count_y", "n_db", "total_db", "n_random", "total_random", "valid", "rank", "confident"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["score"], want["score"]) and np.array_equal(got["p_value"], want["p_value"])
    assert bounds[0] == -np.inf and bounds[-1] == np.inf and bounds[0] < bounds[1] < bounds[2]


def test_shard_bounds_balance_weights():
    from pepquery_b200 import shard
    rng = np.random.default_rng(0)
    m = rng.uniform(500, 4000, 10000)
    w = rng.integers(1, 100, 10000).astype(float)
    for world in (1, 2, 4, 8):
        b = shard.shard_bounds(m, world, w)
        assert len(b) == world + 1
        tot = [w[(m >= b[r]) & (m < b[r + 1])].sum() for r in range(world)]
        assert abs(sum(tot) - w.sum()) < 1e-6                      # a partition: nothing lost, nothing double counted
        assert max(tot) <= 1.05 * w.sum() / world + 100


def test_take_shard_is_a_partition(small_inputs):
    from pepquery_b200 import shard
    prot, tgt, sp = small_inputs
    mass = shard.precursor_mass(sp["prec_mz"], sp["charge"])
    b = shard.shard_bounds(mass, 3)
    seen = []
    for r in range(3):
        sub, gidx = shard.take_shard(sp, b, r)
        seen.extend(gidx.tolist())
        for k in (0, len(gidx) - 1):
            if len(gidx):
                g = gidx[k]
                a0, a1 = int(sp["peak_off"][g]), int(sp["peak_off"][g + 1])
                b0, b1 = int(sub["peak_off"][k]), int(sub["peak_off"][k + 1])
                assert np.array_equal(sp["mz"][a0:a1], sub["mz"][b0:b1]) and sub["prec_mz"][k] == sp["prec_mz"][g]
    assert sorted(seen) == list(range(len(mass)))
