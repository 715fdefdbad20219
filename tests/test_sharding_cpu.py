"""Host-side logic of the multi-GPU path (frame sharding + final gather) on CPU: world size 2 and 3 over gloo.
The per-frame 'extractor' is a stand-in that derives labels from the frame contents, so the test checks the
partitioning, batching, ordering and gather, not the kernels (those are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deplex_b200 import sharding  # noqa: E402


def test_frame_range_partitions_exactly():
    for n in (0, 1, 7, 256, 100000):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.frame_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.batches(3, 10, 4) == [(3, 7), (7, 10)]
    with pytest.raises(ValueError):
        sharding.frame_range(10, 2, 2)


class _FakeExtractor:
    """Labels = low byte of the frame's first float plus the pixel index parity: depends on frame content only."""
    n_points = 12

    def process_batch_host(self, xyz, layout):
        f = xyz.reshape(-1, self.n_points, 3)
        return (f[:, :, 0].astype(np.int32) * 2 + (np.arange(self.n_points, dtype=np.int32) & 1)[None, :])


def _frames(b, e):
    idx = np.arange(b, e, dtype=np.float32)
    return np.broadcast_to(idx[:, None, None], (e - b, _FakeExtractor.n_points, 3)).copy()


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = _FakeExtractor()
        b, e = sharding.frame_range(n_frames, rank, world)
        local = sharding.process_range(ex, _frames, b, e, 1, max_batch=3)
        full = sharding.gather_labels(local, n_frames, dst=0)
        if rank == 0:
            q.put(full.numpy())
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_frames", [(2, 11), (3, 7), (2, 1)])
def test_sharded_run_equals_single_process(world, n_frames):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = sharding.process_range(_FakeExtractor(), _frames, 0, n_frames, 1, max_batch=4)
    assert got.shape == (n_frames, _FakeExtractor.n_points)
    assert np.array_equal(got, want)
