"""CPU tests of the host-side logic: workload generators, settings construction, tile-row partitioning and the
multi-rank exchange (world_size 2, gloo) used by the tile-sharded path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import scenes
from conftest import ROOT


def test_settings_follow_reference_conventions():
    c2w = scenes.orbit_c2w(12)[3]
    v = scenes.make_view(c2w, 1920, 1080, super_sample=2)
    assert (v.image_width, v.image_height) == (3840, 2160)
    w2c = np.linalg.inv(c2w)
    assert np.allclose(v.viewmatrix, w2c.T, atol=1e-6)                       # transposed = column-major for the kernels
    P = scenes.projection_matrix(0.01, 100.0, np.pi / 4, np.pi / 4)
    assert np.allclose(v.projmatrix, (P @ w2c).T, atol=1e-5)                 # full world->clip, transposed
    assert np.allclose(v.campos, c2w[:3, 3]) and v.tanfovx == pytest.approx(1.0)
    p = np.array([0.1, -0.2, 0.05, 1.0], np.float32)                         # kernels compute m[0]x+m[4]y+m[8]z+m[12]
    assert np.allclose(v.viewmatrix.reshape(16)[[2, 6, 10, 14]] @ p, (w2c @ p)[2], atol=1e-6)


def test_workload_generators_are_deterministic_and_shaped():
    a, b = scenes.human_cloud(5000, seed=0), scenes.human_cloud(5000, seed=0)
    assert all(torch.equal(a[k], b[k]) for k in ("means3D", "scales", "rotations", "shs"))
    assert a["shs"].shape == (5000, 13, 3) and a["sh_degree"] == 1 and float(a["shs"][:, 1:].abs().max()) == 0.0
    assert float(a["opacities"].min()) == 1.0 and float(a["scales"].min()) >= 0.0
    m = a["means3D"]
    assert float(m[:, 0].abs().max()) <= 0.54 and float(m[:, 1].abs().max()) <= 1.0 and float(m[:, 2].abs().max()) <= 0.24
    vox = scenes.human_cloud(5000, seed=0, voxelize=64)
    assert torch.equal(vox["means3D"] * 64, torch.round(vox["means3D"] * 64))
    c4 = scenes.random_cloud(1000)
    assert c4["shs"].shape == (1000, 16, 3) and torch.allclose(c4["rotations"].norm(dim=1), torch.ones(1000), atol=1e-5)


def test_balanced_rows_partition():
    import sharding as bench
    cost = np.zeros(68)
    cost[10:50] = np.random.default_rng(0).uniform(1, 100, 40)
    for world in (1, 2, 4, 8):
        parts = bench.balanced_rows(cost, world)
        assert parts[0][0] == 0 and parts[-1][1] == 68
        assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        loads = [cost[a:b].sum() for a, b in parts]
        assert max(loads) <= cost.sum() / world + cost.max() + 1e-9
    assert bench.balanced_rows(np.zeros(5), 4)[-1][1] == 5


def test_rebalance_rows_moves_rows_towards_the_faster_ranks():
    import sharding
    rows = [(0, 22), (22, 40), (40, 55), (55, 70), (70, 85), (85, 100), (100, 114), (114, 128)]
    times = [0.66, 0.70, 0.72, 0.73, 0.73, 0.72, 0.70, 0.65]
    new = sharding.rebalance_rows(rows, times, 128)
    assert new[0][0] == 0 and new[-1][1] == 128
    assert all(new[k][1] == new[k + 1][0] and new[k][1] >= new[k][0] for k in range(7))
    n_old, n_new = [b - a for a, b in rows], [b - a for a, b in new]
    assert n_new[0] >= n_old[0] and n_new[7] >= n_old[7]          # the fast edge ranks take rows ...
    assert n_new[3] <= n_old[3] and n_new[4] <= n_old[4]          # ... from the slow middle ones
    assert sharding.rebalance_rows(rows, [1.0] * 8, 128) == rows   # equal times per row count? equal speed x rows: unchanged
    # an empty range stays empty, nothing is lost
    r2 = sharding.rebalance_rows([(0, 10), (10, 10), (10, 30)], [1.0, 0.0, 1.0], 30)
    assert r2[0][0] == 0 and r2[-1][1] == 30 and r2[1][0] == r2[1][1]


def test_host_block_is_one_buffer_with_aligned_views():
    """renderer.host_block: the five attribute arrays become views into ONE float32 block (one host->device copy per
    step in FramePipeline.enqueue_host); values unchanged, every array starts at a multiple of 64 floats."""
    import scenes
    from renderer import host_block, FramePipeline
    cl = scenes.tiny_cloud(1000, seed=3, sh_degree=1)
    blk = host_block(cl, pin=False)
    flat = blk["_flat"]
    assert flat.dtype == torch.float32 and flat.dim() == 1
    end = 0
    for (n, off, shp), name in zip(blk["_layout"], FramePipeline._ATTRS):
        assert n == name and off % 64 == 0 and off >= end and tuple(blk[n].shape) == tuple(cl[n].shape) == shp
        assert torch.equal(blk[n], cl[n].float())
        assert blk[n].data_ptr() == flat.data_ptr() + 4 * off          # a view, not a copy
        end = off + blk[n].numel()
    assert end <= flat.numel()


def test_grouped_recurrence_equals_the_reference_recurrence():
    """The grouped hit loop of the blend kernel applies alpha values in list order with the recurrence
    tt = T * (done ? 1 : 1 - a);  done' = done or tt < 1e-4;  C += (c * (done' ? 0 : a)) * T;  T = done' ? T : tt
    where a = 0 for an instance the reference skips at this pixel (alpha < 1/255 or power > 0).  This is the reference's loop
    (forward.cu:336-366: skip, test_T = T (1 - alpha), stop test, C += c alpha T, T = test_T, last contributor) with the
    branches turned into selects; checked here operation for operation in float32 on random alpha streams, including
    pixels that stop and streams with skipped instances (fused multiply-adds are emulated the same way on both sides)."""
    rng = np.random.default_rng(11)
    f32 = np.float32
    for trial in range(200):
        n = int(rng.integers(1, 400))
        alpha = rng.uniform(0.0, 0.99, n).astype(np.float32) * (rng.random(n) < 0.8)
        alpha[rng.random(n) < 0.2] = f32(0.001)               # below 1/255: skipped
        power_pos = rng.random(n) < 0.05                      # "power > 0": skipped
        col = rng.random((n, 3)).astype(np.float32)
        # --- the reference's loop
        T, C, last, done = f32(1.0), np.zeros(3, np.float32), 0, False
        for i in range(n):
            if done:
                break
            a = alpha[i]
            if power_pos[i] or a < f32(1.0 / 255.0):
                continue
            test_T = f32(T * f32(f32(1.0) - a))
            if test_T < f32(0.0001):
                done = True
                continue
            for ch in range(3):
                C[ch] = f32(f32(f32(col[i, ch] * a) * T) + C[ch])
            T = test_T
            last = i + 1
        # --- the grouped formulation (phase A: a = 0 where skipped; phase B: selects)
        T2, C2, last2, done2 = f32(1.0), np.zeros(3, np.float32), 0, False
        for i in range(n):
            a = f32(0.0) if (power_pos[i] or alpha[i] < f32(1.0 / 255.0)) else alpha[i]
            om = f32(f32(1.0) - a)
            tt = f32(T2 * (f32(1.0) if done2 else om))
            d = done2 or bool(tt < f32(0.0001))
            eff = f32(0.0) if d else a
            for ch in range(3):
                C2[ch] = f32(f32(f32(col[i, ch] * eff) * T2) + C2[ch])
            T2 = T2 if d else tt
            if (not d) and a != f32(0.0):
                last2 = i + 1
            done2 = d
        assert T.tobytes() == T2.tobytes() and C.tobytes() == C2.tobytes() and last == last2 and done == done2, trial


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sharding
    H, W = 100, 48  # 7 tile rows, last one partial (100 = 6*16 + 4)
    gy = (H + 15) // 16
    cost = np.array([0, 5, 50, 20, 1, 0, 3], float)
    rows = sharding.balanced_rows(cost, world)        # every rank derives the same partition locally
    full = torch.arange(3 * H * W, dtype=torch.float32).view(3, H, W)   # what a single GPU would render
    color = torch.zeros(3, H, W)
    a, b = sharding.pixel_rows(rows, H)[rank]
    color[:, a:b, :] = full[:, a:b, :]                 # this rank's shard; everything else stays zero
    sharding.exchange_image(color, rows, rank)         # the exchange bench.py uses
    q.put((rank, bool(torch.equal(color, full)), rows, gy))
    dist.destroy_process_group()


def test_tile_row_shard_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == res[1][2]  # identical partitions on both ranks


def _views_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sharding
    mine = sharding.views_of_rank(rank, world, 13)
    flags = torch.zeros(13, dtype=torch.int32)
    flags[mine] = 1
    dist.all_reduce(flags)  # test-only collective: every view must be rendered by exactly one rank
    q.put((rank, mine, flags.tolist()))
    dist.destroy_process_group()


def test_view_parallel_sharding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_views_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 2, 4, 6, 8, 10, 12] and res[1][1] == [1, 3, 5, 7, 9, 11]
    assert res[0][2] == [1] * 13 and res[1][2] == [1] * 13


def test_projection_entries_match_the_camera_oracle():
    """The four scalars the binding hands to gs_make_views are the non-trivial entries of getProjectionMatrix
    (simple_raw_render.py:50-69) as restated in oracle/camera.py; the struct mirrors of the new entry points exist."""
    import ctypes as C
    import math

    import numpy as np
    from diff_gaussian_rasterization import _C
    from oracle import camera
    for fx, fy in ((45.0, 45.0), (40.0, 40.0), (60.0, 35.0)):
        P = camera.projection_matrix(0.01, 100, np.pi * fx / 180, np.pi * fy / 180)
        got = np.array(_C.projection_entries(fx, fy), np.float64).astype(np.float32)
        assert np.array_equal(got, np.array([P[0, 0], P[1, 1], P[2, 2], P[2, 3]], np.float32))
        assert P[3, 2] == 1.0 and P[0, 2] == 0.0 and P[1, 2] == 0.0
    assert _C.GS_VIEW_STRIDE == 48 and C.sizeof(_C.GsHeadLayout) == 13 * 4
    assert math.isclose(_C.projection_entries(90.0, 90.0)[0], 1.0, rel_tol=1e-12)
