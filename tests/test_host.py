"""Host-side logic that needs no GPU: error conventions of the operator surface, the no-fallback rule, frame sharding
and the feature-point gather on a 2-process gloo group."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402


def test_get_dimensions_contract():
    from pysilent_b200.util.get_dimensions import get_dimensions
    assert get_dimensions(np.zeros((4, 8, 8, 3))) == 2
    assert get_dimensions(torch.zeros(4, 8, 8, 8, 3)) == 3
    with pytest.raises(TypeError, match="must either be tensor or numpy array"):
        get_dimensions([[1, 2], [3, 4]])


def test_from_image_asserts_like_the_reference():
    from pysilent_b200.util import zoom
    img = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(AssertionError, match="Scale must be greater than one"):
        zoom.from_image(img, 3, (8, 8), 1.0)
    with pytest.raises(AssertionError, match="Number of colors"):
        zoom.from_image(img, 0, (8, 8), 1.5)
    with pytest.raises(AssertionError, match="Each dimension"):
        zoom.from_image(img, 3, (8, -1), 1.5)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks behaviour on a machine without a GPU")
def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA -- it never routes through the oracle or any CPU code."""
    from pysilent_b200 import LineEndPipeline, filters
    from pysilent_b200.util import zoom
    x = np.zeros((1, 8, 8, 3), np.float32)
    for call in (lambda: filters.rgc_filter(x), lambda: LineEndPipeline(output_size=(8, 8)).run(x),
                 lambda: zoom.from_image(np.zeros((32, 32, 3), np.uint8), 3, (8, 8), 1.5),
                 lambda: LineEndPipeline().run_host(np.zeros((1, 480, 640, 3), np.uint8))):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysilent_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, name)).read()
                assert "from oracle" not in text and "import oracle" not in text and "silent_oracle" not in text.replace(
                    "oracle/silent_oracle", ""), os.path.join(base, name)
    assert "scipy" not in open(os.path.join(pkg, "util", "zoom", "from_image.py")).read().replace(
        "scipy.ndimage.zoom(prefilter=False)", "")


def test_missing_library_fails_loudly(monkeypatch):
    from pysilent_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsilent_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_shard_range_partitions_frames():
    from pysilent_b200.distributed import shard_range
    for total in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pysilent_b200.distributed import gather_points, gather_points_padded, shard_range
    levels, total_frames = 3, 5
    lo, hi = shard_range(total_frames, rank, world)
    rs = np.random.RandomState(100 + rank)
    rows = []
    for f in range(hi - lo):                   # local rows (local_level, y, x, 0) in row-major order
        for lvl in range(levels):
            for _ in range(rs.randint(0, 4)):
                rows.append((f * levels + lvl, rs.randint(0, 192), rs.randint(0, 288), 0))
    rows.sort()
    local = torch.tensor(rows, dtype=torch.int64).reshape(-1, 4)
    padded_in = torch.zeros((64, 4), dtype=torch.int64)
    padded_in[: len(local)] = local
    pts, counts = gather_points(padded_in, len(local), frame_offset=lo, levels_per_frame=levels)
    everyone, counts2 = gather_points_padded(padded_in, torch.tensor([len(local)]), lo, levels, 64)
    trimmed = torch.cat([everyone[r, : int(counts2[r])] for r in range(world)])
    assert torch.equal(trimmed, pts) and torch.equal(counts, counts2)
    np.save(os.path.join(out_dir, "local_%d.npy" % rank), (local + torch.tensor([lo * levels, 0, 0, 0])).numpy())
    np.save(os.path.join(out_dir, "gathered_%d.npy" % rank), pts.numpy())
    dist.destroy_process_group()


def test_point_gather_two_ranks_gloo(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = np.concatenate([np.load(tmp_path / ("local_%d.npy" % r)) for r in range(world)])
    for r in range(world):
        got = np.load(tmp_path / ("gathered_%d.npy" % r))
        assert np.array_equal(got, want)
    key = (want[:, 0] * 192 + want[:, 1]) * 288 + want[:, 2]
    assert (np.diff(key) >= 0).all()          # global order = frame-major, then the reference's row-major order


def test_bench_cpu_arm_runs_on_worker_processes():
    """bench.py's CPU arm (`--impl reference`, `cpu_baseline`): one spawned worker process per core runs the oracle port;
    a round of two workers on the small C1 frames returns two finished frames, and a worker's result equals the parent's."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    port = bench.CpuPort(config=1, threads=2)
    try:
        assert port.round() == 2
        mine = port.one(0)
        theirs = list(port.pool.map(bench._cpu_worker_one, [0]))[0]
        assert mine == theirs and mine >= 0
    finally:
        port.close()
