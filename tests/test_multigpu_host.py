"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo processes exercise sharding, the target-blob broadcast
protocol, result gathering and max-over-ranks timing (the GPU box runs the same functions over NCCL)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from simpleslam_b200 import multigpu


class FakeCtx:
    """stands in for capi.Context: a 'built index' that is just a byte string living in host memory"""

    def __init__(self, payload=None):
        self.payload = payload
        self.imported = None

    def target_blob_size(self):
        return len(self.payload)

    def target_export(self, ptr, n):
        import ctypes
        ctypes.memmove(ptr, self.payload, n)

    def target_import(self, ptr, n):
        import ctypes
        self.imported = ctypes.string_at(ptr, n)


def _worker(rank, world, port, n_scans, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    payload = bytes(np.random.RandomState(5).randint(0, 256, 100003, dtype=np.uint8))
    ctx = FakeCtx(payload if rank == 0 else None)
    n, dt = multigpu.broadcast_target(ctx, dist, rank, dev)
    ok_blob = (rank == 0) or (ctx.imported == payload)
    lo, hi = multigpu.shard(n_scans, rank, world)
    # "register" my shard: the result encodes the global scan id so the gather can be verified
    T = [np.eye(4) * (i + 1) for i in range(lo, hi)]
    conv = [(i % 3) != 0 for i in range(lo, hi)]
    allT, allc = multigpu.gather_poses(dist, world, T, conv, n_scans, dev)
    tmax, = multigpu.max_over_ranks(dist, [1.0 + rank], dev)
    tsum, = multigpu.sum_over_ranks(dist, [float(hi - lo)], dev)
    good = ok_blob and n == len(payload) and len(allT) == n_scans and all(allT[i][0, 0] == i + 1 for i in range(n_scans)) \
        and all(allc[i] == ((i % 3) != 0) for i in range(n_scans)) and tmax == float(world) and tsum == float(n_scans)
    open(os.path.join(out_dir, "rank%d.ok" % rank), "w").write("1" if good else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition_is_exact():
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 4, 8):
            spans = [multigpu.shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, 11, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d.ok" % r)).read() == "1"


def test_job_scans_do_not_depend_on_the_sharding():
    """bench.py's fixed C4 job: every rank raycasts only its contiguous shard, and scan k (points, truth, guess) must be the
    same whatever the number of ranks — otherwise the N-GPU runs would not measure the same job"""
    from simpleslam_b200 import synth, workloads
    synth.build()
    n_total = 12
    ident = lambda pts, leaf: pts  # noqa: E731  (no downsample: host-only check)
    whole, wt, wg = workloads.c4_scans("ndt", ident, 0, n_total, n_total, tiles=1, workers=4)
    for world in (2, 4):
        got, gt, gg = [], [], []
        for r in range(world):
            lo, hi = multigpu.shard(n_total, r, world)
            s, t, g = workloads.c4_scans("ndt", ident, lo, hi, n_total, tiles=1, workers=2)
            got += s; gt += t; gg += g
        assert len(got) == n_total
        for a, b in zip(whole, got):
            assert np.array_equal(a, b)
        assert all(np.array_equal(a, b) for a, b in zip(wt, gt)) and all(np.array_equal(a, b) for a, b in zip(wg, gg))
