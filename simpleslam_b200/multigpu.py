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


_GATHER_BUFFERS = {}


def gather_poses(dist, world, local_T, local_conv, n_total, device):
    """all-gather of the per-scan results (16 doubles + converged flag) of contiguous shards -> arrays of length n_total.
    One host->device copy, one all_gather_into_tensor, one device->host copy per call; the buffers are kept between calls
    (at 8 ranks the job step is a few milliseconds: a copy per rank and fresh allocations were a fifth of it)."""
    counts = [shard(n_total, r, world)[1] - shard(n_total, r, world)[0] for r in range(world)]
    mx = max(counts) if counts else 0
    if n_total == 0 or mx == 0:
        return np.zeros((0, 4, 4)), np.zeros(0, bool)
    key = (world, mx, str(device))
    bufs = _GATHER_BUFFERS.get(key)
    if bufs is None:
        pin = device != "cpu" and str(device) != "cpu" and torch.cuda.is_available()
        h_in = torch.zeros((mx, 17), dtype=torch.float64, pin_memory=pin)
        h_out = torch.zeros((world * mx, 17), dtype=torch.float64, pin_memory=pin)
        bufs = (h_in, h_out, torch.zeros((mx, 17), dtype=torch.float64, device=device), torch.zeros((world * mx, 17), dtype=torch.float64, device=device))
        _GATHER_BUFFERS[key] = bufs
    h_in, h_out, d_in, d_out = bufs
    k = len(local_T)
    hv = h_in.numpy()
    if k:
        hv[:k, :16] = np.asarray(local_T, dtype=np.float64).reshape(k, 16)
        hv[:k, 16] = np.asarray(local_conv, dtype=np.float64)
    hv[k:] = 0.0
    d_in.copy_(h_in, non_blocking=True)
    if hasattr(dist, "all_gather_into_tensor") and d_in.is_cuda:
        dist.all_gather_into_tensor(d_out, d_in)
    else:  # gloo (CPU tests): list form
        outs = list(d_out.view(world, mx, 17).unbind(0))
        dist.all_gather(outs, d_in)
    h_out.copy_(d_out, non_blocking=True)
    if d_out.is_cuda:
        torch.cuda.current_stream().synchronize()
    allv = h_out.numpy().reshape(world, mx, 17)
    T = np.concatenate([allv[r, :c, :16].reshape(c, 4, 4) for r, c in enumerate(counts)])
    conv = np.concatenate([allv[r, :c, 16] for r, c in enumerate(counts)]) > 0.5
    return T, conv


def max_over_ranks(dist, values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(dist, values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
