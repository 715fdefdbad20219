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


def process_range_device(pipeline, pool, pool_frames, begin, end, layout, max_batch, labels_out, stream):
    """Device-resident counterpart of process_range for synthetic sequences: frame i of the sequence is frame
    i % pool_frames of `pool`, a CUDA tensor holding the unique frames tiled to at least max_batch + pool_frames frames,
    so that any batch of consecutive sequence frames is one contiguous slice of it.  Batches go through `pipeline`
    (deplex_b200.PipelinedExtractor, i.e. the C-ABI dpx_pipeline) on `stream`; labels land in labels_out[i - begin].
    Asynchronous: the caller joins the pipeline."""
    n = pipeline.n_points
    frame_bytes_in, frame_bytes_out = n * 3 * 4, n * 4
    assert pool.shape[0] >= max_batch + pool_frames and labels_out.shape[0] >= end - begin
    for b, e in batches(begin, end, max_batch):
        pipeline.submit_ptr(pool.data_ptr() + (b % pool_frames) * frame_bytes_in, e - b, layout,
                            labels_out.data_ptr() + (b - begin) * frame_bytes_out, stream.cuda_stream)


def gather_labels(local_labels, n_frames, dst=0, group=None, out=None):
    """Gather every rank's (frames_of_rank, n_points) label block to rank `dst`, in frame order: the 'final gather' of
    the north star (1.2 MB per VGA frame, off the hot path).  Point-to-point, straight into the destination's slices --
    no padding and no staging copy, so a 100 000-frame sequence needs only the result itself on `dst`.  Works on CPU
    tensors (gloo) and CUDA tensors (NCCL over NVLink).  `out` (on dst): a preallocated (n_frames, n_points) tensor;
    when dst's own block is already a view of its slice, nothing is copied for it.  Returns the full tensor on `dst`,
    None elsewhere."""
    import torch
    import torch.distributed as dist
    t = local_labels if isinstance(local_labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_labels))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if out is not None and out.data_ptr() != t.data_ptr():
            out[: t.shape[0]].copy_(t)
            return out
        return t
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = [frame_range(n_frames, r, world) for r in range(world)]
    assert t.shape[0] == sizes[rank][1] - sizes[rank][0], "local block does not match this rank's frame range"
    t = t.contiguous()
    if rank != dst:
        if t.shape[0] > 0:
            dist.send(t, dst=dst, group=group)
        return None
    if out is None:
        out = torch.empty((n_frames, t.shape[1]), dtype=t.dtype, device=t.device)
    mine = out[sizes[rank][0]:sizes[rank][1]]
    if mine.data_ptr() != t.data_ptr():
        mine.copy_(t)
    # frame-major rows: every rank's block is one contiguous slice of `out`, so it is received in place
    ops = [dist.P2POp(dist.irecv, out[sizes[r][0]:sizes[r][1]], r, group=group)
           for r in range(world) if r != dst and sizes[r][1] > sizes[r][0]]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


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
