"""Frame sharding across the GPUs of one box.

`process()` keeps no state across frames (plane_extractor.cpp:281,428), so a depth sequence is split into
static contiguous ranges, one per rank (one process per GPU), and every rank runs the ordinary batched entry
points on its range.  There is NO collective on the data path; the only exchange is the optional final gather
of the label maps (4 B/pixel) to one rank, done with torch.distributed (NCCL over NVLink on GPUs, gloo on CPU
for the tests).
"""
import numpy as np


def frame_range(n_frames, rank, world):
    """Contiguous, balanced range [begin, end) of `rank` out of `world`: sizes differ by at most one frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_frames), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def batches(begin, end, max_batch):
    """Split a rank's range into device batches of at most `max_batch` frames."""
    out = []
    b = begin
    while b < end:
        e = min(end, b + max_batch)
        out.append((b, e))
        b = e
    return out


def process_range(extractor, load_frames, begin, end, layout, max_batch):
    """Run `extractor` over frames [begin, end); `load_frames(b, e)` returns a host float32 array holding frames
    b..e-1 back to back.  Returns (end - begin, n_points) int32 labels (host)."""
    n = extractor.n_points
    out = np.empty((end - begin, n), dtype=np.int32)
    for b, e in batches(begin, end, max_batch):
        out[b - begin:e - begin] = extractor.process_batch_host(load_frames(b, e), layout)
    return out


def gather_labels(local_labels, n_frames, dst=0, group=None):
    """Gather every rank's (frames_of_rank, n_points) int32 label block to rank `dst`, in frame order.
    Works on CPU tensors (gloo) and CUDA tensors (NCCL).  Returns the full (n_frames, n_points) tensor on `dst`,
    None elsewhere.  This is the 'final gather' of the north star: 1.2 MB per VGA frame, off the hot path."""
    import torch
    import torch.distributed as dist
    t = local_labels if isinstance(local_labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_labels))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_points = t.shape[1]
    sizes = [frame_range(n_frames, r, world) for r in range(world)]
    assert t.shape[0] == sizes[rank][1] - sizes[rank][0], "local block does not match this rank's frame range"
    # gather needs equal shapes: pad every block to the largest range
    longest = max(e - b for b, e in sizes)
    padded = torch.zeros((longest, n_points), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    if t.is_cuda:
        # NCCL: all_gather into one buffer (gather is not universally supported), rank dst keeps the result
        buf = torch.empty((world, longest, n_points), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        blocks = list(buf) if rank == dst else None
    else:
        blocks = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, blocks, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([blocks[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers it
    allocates next (first touch) and its copy-engine traffic stay on the GPU's side of the socket interconnect.
    With 8 ranks each moving ~50 GB/s over PCIe, unbound buffers put most of that on the inter-socket link.
    Returns the node number, or None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        finally:
            pynvml.nvmlShutdown()
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
