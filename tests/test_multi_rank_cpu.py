"""CPU, world_size 2, gloo: the multi-GPU plumbing of the path — frames shard into contiguous blocks,
rank r accumulates into histogram rows [r*B, (r+1)*B), and DetectCountPipeline.all_reduce (the path's only collective)
merges them.  The per-rank histograms are produced by the CPU oracle here (no GPU), so this covers the host logic:
slot ownership, the all-reduce call and the sparse-dict conversion of the merged table."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import floor_oracle as fo


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, B: int, Z: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from office_person_detection_vit_b200.pipeline import DetectCountPipeline

        zones = fo.grid_zones(Z)
        hist = torch.zeros(B * world, Z + 1, dtype=torch.int32)
        # rank r owns frames [r*B, (r+1)*B): 20 + frame_index points per frame, seeded by the global frame index
        for f in range(rank * B, (rank + 1) * B):
            pts = fo.camera_points(20 + f, seed=1000 + f)
            idx, _ = fo.project_classify_count(fo.H_CONFIG, pts, zones)
            hist[f] += torch.from_numpy(fo.count(zone_idx=idx, Z=Z)[0].astype(np.int32))
        assert int(hist[:rank * B].sum()) == 0 and int(hist[(rank + 1) * B:].sum()) == 0
        DetectCountPipeline(None, None, None).all_reduce(hist)
        q.put((rank, hist.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_frame_sharded_histogram_all_reduce():
    world, B, Z = 2, 3, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, Z, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    zones = fo.grid_zones(Z)
    expect = np.zeros((B * world, Z + 1), dtype=np.int64)
    for f in range(B * world):
        idx, _ = fo.project_classify_count(fo.H_CONFIG, fo.camera_points(20 + f, seed=1000 + f), zones)
        expect[f] = fo.count(zone_idx=idx, Z=Z)[0]
    assert (got[0] == expect).all() and (got[1] == expect).all()      # every rank holds the global table
    assert expect.sum(axis=1).tolist() == [20 + f for f in range(B * world)]
