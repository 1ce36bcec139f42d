"""Multi-GPU sharding of the rasteriser path: one process per GPU, views sharded by batch element.

The op is independent per batch element (every input and gradient is indexed by b,
CudaRenderer.cpp:312-324 / CudaRendererGrad.cpp:264-283), so whole batch elements go to ranks and
NO collective is needed inside the op.  A collective exists only for parameters the caller shares
across the batch (one SH set / one texture / identity-shared vertices for all batch elements):
their gradients are summed over the local batch slice and then all-reduced once per step in a
single flat buffer (NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no
multi-GPU path at all (python/utils/CheckGPU.py:51-52 masks all but one GPU).
"""
import torch
import torch.distributed as dist


def partition_batch(batch, world_size):
    """Contiguous [start, stop) slices of batch elements per rank, sizes differing by at most 1."""
    base, extra = divmod(int(batch), int(world_size))
    out, s = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def local_slice(batch, rank=None, world_size=None):
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    return partition_batch(batch, world_size)[rank]


def shard_inputs(inputs, rank=None, world_size=None):
    """inputs: dict of op inputs whose dim 0 is the batch.  Returns this rank's slice of each."""
    B = next(iter(inputs.values())).shape[0]
    s, e = local_slice(B, rank, world_size)
    return {k: v[s:e] for k, v in inputs.items()}


def allreduce_shared_grads(grads, group=None, async_op=False):
    """Sum gradients of batch-shared parameters across ranks with ONE collective.

    grads: list of tensors (already summed over the local batch slice).  They are packed into one
    flat fp32 buffer (latency-bound messages: SH 27*C*4 B, vertices N*12 B, texture texH*texW*12 B),
    all-reduced, and unpacked in place.  Returns a handle with .wait() when async_op."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return _Done()
    flat = _common_flat_buffer(grads)
    if flat is not None:
        # the gradients already live back to back in one buffer (NativeRenderer.backward(out=...) with views of
        # shared_grad_buffer): one in-place collective, no pack / unpack kernels
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        h = _Pending(work.wait)
        if not async_op:
            h.wait()
        return h
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def finish():
        work.wait()
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    h = _Pending(finish)
    if not async_op:
        h.wait()
    return h


def shared_grad_buffer(shapes, device):
    """One flat fp32 buffer + a view per shape, laid out back to back: pass the views as `out=` of the backward
    and the list of views to allreduce_shared_grads, which then reduces the buffer in place."""
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=device)
    views, off = [], 0
    for s, n in zip(shapes, sizes):
        views.append(flat[off:off + n].view(s))
        off += n
    return flat, views


def _common_flat_buffer(grads):
    """The flat tensor the gradients are consecutive views of, or None."""
    if not grads or any(g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
        return None
    st = grads[0].untyped_storage()
    off = grads[0].storage_offset()
    start = off
    for g in grads:
        if g.untyped_storage().data_ptr() != st.data_ptr() or g.storage_offset() != off:
            return None
        off += g.numel()
    return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, start, (off - start,))


class _Done:
    def wait(self):
        return None


class _Pending:
    def __init__(self, fn):
        self._fn = fn

    def wait(self):
        if self._fn is not None:
            self._fn()
            self._fn = None


def gather_batch(local, batch, group=None):
    """All-gather per-batch-element results (e.g. sh_coeff_grad rows, which are disjoint per (b,c))."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    parts = partition_batch(batch, world)
    nmax = max(e - s for s, e in parts)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:e - s] for o, (s, e) in zip(outs, parts)], 0)
