"""Multi-GPU sharding of the rasteriser path: one process per GPU, views sharded by batch element and camera.

The op is independent per batch element (every input and gradient is indexed by b,
CudaRenderer.cpp:312-324 / CudaRendererGrad.cpp:264-283), so whole batch elements go to ranks and
NO collective is needed inside the op.  A collective exists only for

  * parameters the caller shares across the batch (one SH set / one texture / identity-shared vertices for all
    batch elements): their gradients are summed over the local batch slice and all-reduced once per step;
  * the CAMERA SPLIT, used when there are fewer batch elements than ranks (plan_views): the cameras of one batch
    element all accumulate into the same vertex_pos / vertex_color / texture gradients
    (CudaRendererGrad.cpp:264-283), so those are summed over the ranks that hold the element's cameras, and the
    per-(b, c) sh_coeff_grad rows are gathered (reduce_camera_split).

Two transports for that one sum: torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests), and -- for
the latency-bound messages of this path (hundreds of KB) -- SymmetricGradBuffer: the gradients are written straight
into peer-mapped symmetric memory and summed by a few CTAs inside the backward's last kernel (one-shot all-reduce
over NVLink peer loads or NVLS multimem.ld_reduce, csrc/gvv_collective.cuh), no extra launch, no NCCL call.
The reference has no multi-GPU path at all (python/utils/CheckGPU.py:51-52 masks all but one GPU).
"""
import ctypes

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


# ---------------------------------------------------------------------------------------------------------------
# camera split: fewer batch elements than ranks
# ---------------------------------------------------------------------------------------------------------------
def plan_views(batch, cameras, world_size):
    """Per rank: (b0, b1, c0, c1) = batch elements [b0, b1) x cameras [c0, c1) it renders; None for an idle rank.

    batch >= world_size: whole batch elements per rank (partition_batch), all cameras.  batch < world_size: rank r
    belongs to batch element r * batch // world_size, and the ranks of one element split its cameras contiguously
    (sizes differing by at most one); ranks beyond an element's camera count stay idle."""
    batch, cameras, world_size = int(batch), int(cameras), int(world_size)
    if batch >= world_size:
        return [(s, e, 0, cameras) if e > s else None for s, e in partition_batch(batch, world_size)]
    owner = [r * batch // world_size for r in range(world_size)]
    plan = []
    for r in range(world_size):
        b = owner[r]
        team = [q for q in range(world_size) if owner[q] == b]
        c0, c1 = partition_batch(cameras, len(team))[team.index(r)]
        plan.append((b, b + 1, c0, c1) if c1 > c0 else None)
    return plan


def camera_teams(plan):
    """Lists of ranks that share a batch element (and therefore must sum its gradients), only where a split exists."""
    teams = {}
    for r, p in enumerate(plan):
        if p is not None:
            teams.setdefault((p[0], p[1]), []).append(r)
    return [t for t in teams.values() if len(t) > 1]


def shard_views(inputs, cameras, plan_entry):
    """This rank's slice of the op inputs (dict with the op's names) for plan entry (b0, b1, c0, c1):
    vertex_pos / vertex_color / texture by batch element; sh_coeff, target_image, extrinsics, intrinsics also by camera."""
    b0, b1, c0, c1 = plan_entry
    C = int(cameras)
    out = {}
    for k, v in inputs.items():
        v = v[b0:b1]
        if k == "sh_coeff":
            v = v[:, c0:c1]
        elif k == "target_image":
            v = v[:, c0:c1]
        elif k == "extrinsics":
            v = v.reshape(v.shape[0], C, 12)[:, c0:c1].reshape(v.shape[0], -1)
        elif k == "intrinsics":
            v = v.reshape(v.shape[0], C, 9)[:, c0:c1].reshape(v.shape[0], -1)
        out[k] = v.contiguous() if hasattr(v, "contiguous") else v
    return out


def reduce_camera_split(grads, plan, cameras, rank=None, groups=None):
    """grads = (vertex_pos_grad, vertex_color_grad, texture_grad, sh_coeff_grad) of THIS rank's slice (entries may be
    None).  Sums the first three over the ranks of this rank's camera team (one all-reduce of a flat buffer) and
    gathers the sh_coeff_grad rows of the team into [b1 - b0, cameras, 27].  groups: {tuple(team): ProcessGroup} made
    once with make_team_groups (new_group is collective over ALL ranks).  Without a split the gradients are returned
    unchanged."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    mine = plan[rank]
    team = next((t for t in camera_teams(plan) if rank in t), None)
    gpos, gcol, gtex, gsh = grads
    if mine is None or team is None:
        return gpos, gcol, gtex, gsh
    group = groups[tuple(team)]
    allreduce_shared_grads([g for g in (gpos, gcol, gtex) if g is not None], group=group)
    C = int(cameras)
    full = torch.zeros((gsh.shape[0], C, 27), dtype=gsh.dtype, device=gsh.device)
    full[:, mine[2]:mine[3]] = gsh
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)        # rows are disjoint: the sum is the gather
    return gpos, gcol, gtex, full


def make_team_groups(plan):
    """One process group per camera team; every rank must call this (dist.new_group is collective)."""
    return {tuple(t): dist.new_group(ranks=t) for t in camera_teams(plan)}


# ---------------------------------------------------------------------------------------------------------------
# one-shot all-reduce over symmetric memory (NVLink peer loads / NVLS), fused into the backward
# ---------------------------------------------------------------------------------------------------------------
class SymmetricGradBuffer:
    """Gradient outputs in symmetric memory + the descriptor gvv_backward needs to sum them across ranks itself.

    shapes: shapes of the gradients to be summed (e.g. [(1, C, 27), (1, N, 3)]).  Two SLOTS of views alternate from
    step to step (a slot is rewritten two barriers after it was read, so no second barrier per step is needed):

        buf = SymmetricGradBuffer([(1, C, 27), (1, N, 3)], device)
        gsh, gcol = buf.views(step % 2)                    # pass as out= of NativeRenderer.backward
        buf.attach(renderer, step % 2)                     # gvv_set_allreduce
        renderer.backward(..., out=(None, gcol, None, gsh))
        gsh_sum, gcol_sum = buf.results(step % 2)          # valid in stream order after the backward

    mode: "p2p" (every rank loads all peers' copies over NVLink), "nvls" (multimem.ld_reduce on the multicast
    mapping: the switch adds) or "auto" (nvls when the allocation has a multicast mapping).  after_backward: the
    range includes vertex_pos_grad (camera split) -> reduced by a launch of its own after the backward.
    Allocation and rendezvous go through torch.distributed._symmetric_memory (plumbing); the exchange itself is
    this repo's kernel code (csrc/gvv_collective.cuh)."""

    FIRST_CHANNEL = 16

    def __init__(self, shapes, device, group=None, mode="auto", channels=None, after_backward=False):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native
        self._native = _native
        group = group if group is not None else dist.group.WORLD
        self.shapes = [tuple(int(x) for x in s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.count = sum(self.sizes)
        self.slot = (self.count + 3) // 4 * 4
        self.device = torch.device(device)
        self.buf = symm_mem.empty(2 * self.slot, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if self.world > 8:
            raise RuntimeError("SymmetricGradBuffer: at most 8 ranks (one NVSwitch domain)")
        mc = int(self.hdl.multicast_ptr or 0)
        if mode == "auto":
            mode = "nvls" if mc else "p2p"
        if mode == "nvls" and not mc:
            raise RuntimeError("SymmetricGradBuffer: NVLS requested but the allocation has no multicast mapping")
        self.mode = mode
        pad_words = int(symm_mem.get_signal_pad_size()) // 4
        if channels is None:        # one pass of 2 float4 per thread over the range: remote loads are latency-bound (~1 NVLink round trip)
            channels = max(4, -(-self.slot // 4 // 512))
        self.epoch_word = pad_words - 64                    # the last 64 words of the pad: one barrier counter per CTA
        self.channels = max(1, min(int(channels), self.epoch_word // self.world - self.FIRST_CHANNEL, 64))
        self.after_backward = bool(after_backward)
        self.result = torch.empty(2 * self.slot, dtype=torch.float32, device=self.device)   # one per slot: step i's sums stay readable during step i+1
        self._descs = []
        for s in (0, 1):
            d = _native.gvv_allreduce_desc(int(self.hdl.buffer_ptrs_dev), int(self.hdl.signal_pad_ptrs_dev), mc if mode == "nvls" else 0,
                                           self.rank, self.world, s * self.slot, self.slot if mode == "nvls" else self.count,
                                           self.result.data_ptr() + 4 * s * self.slot, 1 if mode == "nvls" else 0, self.channels, self.FIRST_CHANNEL,
                                           self.epoch_word, int(self.after_backward))
            self._descs.append(d)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)          # every rank's buffer is zeroed and mapped before anyone's first step

    def _split(self, flat):
        out, off = [], 0
        for shp, n in zip(self.shapes, self.sizes):
            out.append(flat[off:off + n].view(shp))
            off += n
        return out

    def views(self, slot):
        return self._split(self.buf[slot * self.slot:(slot + 1) * self.slot])

    def results(self, slot):
        return self._split(self.result[slot * self.slot:(slot + 1) * self.slot])

    def attach(self, renderer, slot):
        renderer.set_allreduce(self._descs[slot])


class SharedGrads:
    """Which gradients of one backward go through a SymmetricGradBuffer slot (CudaRendererGpu(sharedGrads_attr=...) or
    NativeRenderer.backward(out=shared.outputs()) + buffer.attach).

    names: the op inputs whose gradients the buffer's shapes describe, in the buffer's order, out of
    ("vertex_pos", "vertex_color", "texture", "sh_coeff").  sh_rows = (c0, c1): with a camera split the buffer holds
    sh_coeff_grad for ALL cameras of the batch element and this rank fills rows [c0, c1) (the other rows stay zero,
    so the sum over the team is the gather)."""
    ORDER = ("vertex_pos", "vertex_color", "texture", "sh_coeff")

    def __init__(self, buffer, slot, names, sh_rows=None):
        if len(names) != len(buffer.shapes) or any(n not in self.ORDER for n in names):
            raise ValueError("SharedGrads: names must match the buffer's shapes and be op input names")
        self.buffer, self.slot, self.names, self.sh_rows = buffer, int(slot), tuple(names), sh_rows

    def outputs(self):
        """The out= tuple for NativeRenderer.backward: views of the slot for the named gradients, None elsewhere."""
        out = [None, None, None, None]
        for name, v in zip(self.names, self.buffer.views(self.slot)):
            if name == "sh_coeff" and self.sh_rows is not None:
                v = v[:, self.sh_rows[0]:self.sh_rows[1]]
            out[self.ORDER.index(name)] = v
        return tuple(out)

    def reduced(self, grads):
        """grads with the named entries replaced by their cross-rank sums (views of the slot's result buffer; valid
        until the slot is used again two steps later)."""
        g = list(grads)
        for name, v in zip(self.names, self.buffer.results(self.slot)):
            if name == "sh_coeff" and self.sh_rows is not None:
                v = v[:, self.sh_rows[0]:self.sh_rows[1]]
            g[self.ORDER.index(name)] = v
        return tuple(g)
