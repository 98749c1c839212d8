"""Frame-sharded data parallelism: one process per GPU, frames (or camera streams) split across ranks, and ONE
collective per batch -- the variable-length gather of feature points.

The reference is single-process / single-GPU (``recognition_testing.py:64``); nothing in S1-S8 couples frames, so the
data path needs no exchange. Points are tiny (32 B each, a handful per level): the gather is latency-bound, so it is
issued as one padded ``all_gather`` per batch rather than per frame. Works on any ``torch.distributed`` backend
(``nccl`` on the GPU box over NVLink, ``gloo`` in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def bind_to_device_numa_node(device_index):
    """Pin this process (CPU affinity, hence first-touch placement of the page-locked frame / result buffers it allocates
    afterwards) to the NUMA node its GPU hangs off. With one process per GPU on a two-socket host, host<->device copies
    that cross the socket interconnect cap the end-to-end rate of all ranks together; local buffers do not. Returns the
    node id, or None when the topology cannot be read (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device_index]) if visible and visible.replace(",", "").isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml reports an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(total, rank, world_size):
    """Contiguous block of ``total`` frames owned by ``rank`` (block sizes differ by at most one)."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_points_padded(points, count, frame_offset, levels_per_frame, capacity, group=None):
    """Sync-free form for steady-state loops: every rank contributes exactly ``capacity`` rows.

    :param count: 1-element int64 DEVICE tensor (as written by ``silent_pipeline_run``).
    :return: ``(everyone [world, capacity, 4], counts [world])``; rows beyond ``counts[r]`` are padding.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    padded = points[:capacity].clone()
    padded[:, 0] += frame_offset * levels_per_frame
    counts = torch.empty(world, dtype=torch.int64, device=points.device)
    everyone = torch.empty((world, capacity, 4), dtype=torch.int64, device=points.device)
    if world == 1:
        counts.copy_(count.reshape(1))
        everyone[0].copy_(padded)
        return everyone, counts
    dist.all_gather_into_tensor(counts, count.reshape(1), group=group)
    dist.all_gather_into_tensor(everyone.view(world * capacity, 4), padded, group=group)
    return everyone, counts


def gather_points(points, count=None, frame_offset=0, levels_per_frame=1, capacity=None, group=None):
    """Gather every rank's feature points in global frame order.

    :param points: int64 ``[cap_or_K, 4]`` rows ``(local_level, y, x, 0)`` of this rank (row-major order).
    :param count: number of valid rows (int or 1-element tensor); default ``len(points)``.
    :param frame_offset: index of this rank's first frame in the global batch; level ids are rebased to
        ``global_frame * levels_per_frame + level`` so the concatenation is in the reference's row-major order
        (ranks hold contiguous frame blocks, see :func:`shard_range`).
    :param capacity: fixed padded row count per rank (same on every rank); default = max count over ranks
        (costs one extra tiny all-reduce).
    :return: ``(points [K_total, 4], counts [world_size])`` on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = points.device
    n = int(count) if count is not None else int(points.shape[0])
    n = min(n, int(points.shape[0]))
    local = points[:n].clone()
    local[:, 0] += frame_offset * levels_per_frame
    if world == 1:
        return local, torch.tensor([n], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    cap = int(capacity) if capacity is not None else int(counts.max().item())
    cap = max(cap, 1)
    padded = torch.zeros((cap, 4), dtype=torch.int64, device=dev)
    padded[:min(n, cap)] = local[:cap]
    everyone = torch.empty((world * cap, 4), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everyone, padded, group=group)
    host_counts = counts.tolist()
    parts = [everyone[r * cap: r * cap + min(int(host_counts[r]), cap)] for r in range(world)]
    return torch.cat(parts, dim=0), counts


class PointGather:
    """The steady-state form of the gather: ONE fixed-size all-gather per batch, issued on a side stream so that it
    overlaps the next batch's kernels; double-buffered, no host sync, no allocation per step.

    ``submit`` (called on the compute stream right after the pipeline call) packs the rank's points and count into the
    slot's send buffer with one small kernel (``silent_pack_points``) and queues the collective behind it on the side
    stream; ``result`` waits for a slot and returns ``(points [K_total, 4], counts [world])`` in global frame order.
    CUDA tensors + NCCL only (the gloo/CPU form is :func:`gather_points`).
    """

    def __init__(self, capacity, levels_per_frame, device, group=None, depth=2, native=False):
        """``native``: issue the collective through the library's own entry point (``silent_gather_points``: one
        ``ncclAllGather`` on the side stream, communicator created from an id that rank 0 broadcasts through
        ``torch.distributed``) instead of ``torch.distributed.all_gather_into_tensor``."""
        from . import _lib, _ops
        self._lib, self._ops = _lib, _ops
        self.comm = None
        self.capacity, self.levels, self.group = int(capacity), int(levels_per_frame), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        self.side = torch.cuda.Stream(self.device)
        rows = self.capacity + 1
        self.send = [torch.zeros((rows, 4), dtype=torch.int64, device=self.device) for _ in range(depth)]
        self.recv = [torch.zeros((self.world, rows, 4), dtype=torch.int64, device=self.device) for _ in range(depth)]
        self.packed = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.work = [None] * depth      # the collective's Work handle: result() asks it for asynchronous NCCL errors
        self.submitted = 0
        if native and self.world > 1:
            self.comm = self._create_native_comm(group)

    def _create_native_comm(self, group):
        import ctypes
        rank = dist.get_rank(group)
        on_gpu = dist.get_backend(group) == "nccl"
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            self._lib.check(self._lib.lib().silent_comm_unique_id(buf), "silent_comm_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(self.device) if on_gpu else ident
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(ident.cpu().tolist())
        comm = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib().silent_comm_create(raw, self.world, rank, ctypes.byref(comm)),
                            "silent_comm_create")
        return comm

    def close(self):
        if self.comm is not None:
            self._lib.lib().silent_comm_destroy(self.comm)
            self.comm = None

    def submit(self, points, count, frame_offset):
        """points: int64 CUDA ``[>= capacity, 4]``, count: 1-element int64 CUDA tensor. Returns the slot index."""
        k = self.submitted % len(self.send)
        self.submitted += 1
        cur = torch.cuda.current_stream(self.device)
        if points.shape[0] < self.capacity:
            raise ValueError("points buffer holds %d rows, the gather capacity is %d" % (points.shape[0], self.capacity))
        cur.wait_event(self.done[k])      # the slot's previous collective has read the send buffer (no-op at first use)
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib().silent_pack_points(
                self._ops.ptr(points), self._ops.ptr(count), self.capacity, int(frame_offset) * self.levels,
                self._ops.ptr(self.send[k]), self._ops.stream_ptr()), "silent_pack_points")
        self.packed[k].record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.packed[k])
            if self.world == 1:
                self.recv[k][0].copy_(self.send[k], non_blocking=True)
            elif self.comm is not None:
                with torch.cuda.device(self.device):
                    self._lib.check(self._lib.lib().silent_gather_points(
                        self.comm, self._ops.ptr(self.send[k]), self.capacity + 1, self._ops.ptr(self.recv[k]),
                        self.side.cuda_stream), "silent_gather_points")
            else:
                self.work[k] = dist.all_gather_into_tensor(self.recv[k].view(-1, 4), self.send[k], group=self.group,
                                                           async_op=True)
                self.work[k].wait()     # orders the side stream behind the collective; does not block the host
            self.done[k].record(self.side)
        return k

    def result(self, slot):
        self.done[slot].synchronize()
        if self.comm is not None:
            self._lib.check(self._lib.lib().silent_comm_check(self.comm), "silent_comm_check")
        work = self.work[slot]
        if work is not None:
            # NCCL reports failures of an already enqueued collective asynchronously (a peer died, the communicator was
            # aborted, the watchdog timed out): surface them here instead of handing out a half-filled buffer
            work.wait()                     # raises the NCCL error, if any (torch's async error handling)
            if not work.is_completed():
                raise RuntimeError("feature-point all-gather did not complete")
            counts_ok = self.recv[slot][:, self.capacity, 1:].abs().sum().item() == 0   # the count row is (count, 0, 0, 0)
            if not counts_ok:
                raise RuntimeError("feature-point all-gather delivered a malformed count row")
        recv = self.recv[slot]
        counts = recv[:, self.capacity, 0].clone()
        host = counts.tolist()
        parts = [recv[r, :min(int(host[r]), self.capacity)] for r in range(self.world)]
        return torch.cat(parts, dim=0), counts
