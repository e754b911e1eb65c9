"""Multi-GPU plumbing for batched independent registrations: one process per GPU, the static-map index is
built once on rank 0 and broadcast to the other ranks with torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests); scans are sharded in contiguous blocks; results are gathered at the end. There is NO per-iteration
collective: every registration's reductions are device-local."""
import numpy as np
import torch


def shard(n_items, rank, world):
    """contiguous block partition of n_items over `world` ranks"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_target(ctx, dist, rank, device, src=0):
    """Rank `src` exports its built target index into one contiguous blob; everyone else imports it.
    `ctx` needs target_blob_size() / target_export(ptr, n) / target_import(ptr, n) (capi.Context on the GPU box).
    Returns (blob_bytes, seconds spent in the broadcast collective)."""
    import time
    nbytes = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        nbytes[0] = ctx.target_blob_size()
    dist.broadcast(nbytes, src)
    n = int(nbytes.item())
    blob = torch.empty(n, dtype=torch.uint8, device=device)
    if rank == src:
        ctx.target_export(blob.data_ptr(), n)
    if blob.is_cuda:
        torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    dist.broadcast(blob, src)
    if blob.is_cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank != src:
        ctx.target_import(blob.data_ptr(), n)
    return n, dt


def gather_poses(dist, world, local_T, local_conv, n_total, device):
    """all-gather of the per-scan results (16 doubles + converged flag) of contiguous shards -> arrays of length n_total"""
    counts = [shard(n_total, r, world)[1] - shard(n_total, r, world)[0] for r in range(world)]
    mx = max(counts) if counts else 0
    buf = torch.zeros((mx, 17), dtype=torch.float64, device=device)
    k = len(local_T)
    if k:
        buf[:k, :16] = torch.as_tensor(np.asarray(local_T, dtype=np.float64).reshape(k, 16), device=device)
        buf[:k, 16] = torch.as_tensor(np.asarray(local_conv, dtype=np.float64), device=device)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    T = np.concatenate([o[:c, :16].cpu().numpy().reshape(c, 4, 4) for o, c in zip(outs, counts)]) if n_total else np.zeros((0, 4, 4))
    conv = np.concatenate([o[:c, 16].cpu().numpy() for o, c in zip(outs, counts)]) > 0.5 if n_total else np.zeros(0, bool)
    return T, conv


def max_over_ranks(dist, values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(dist, values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
