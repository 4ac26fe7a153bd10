"""CPU: the N>1 host logic -- sharding, record packing and the single gather -- with world_size 2 on gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keras_ocr_b200 import distributed as D


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    counts = [3, 0, 2]
    boxes = [rng.uniform(0, 500, (c, 4, 2)).astype(np.float32) for c in counts]
    labels = rng.integers(-1, 37, (5, 48)).astype(np.int8)
    rec = D.pack_records(counts, boxes, labels, per_rank=4, max_boxes=8)
    got = D.unpack_records(rec, 8)
    assert [g[0] for g in got] == counts
    assert np.array_equal(got[0][1], boxes[0]) and np.array_equal(got[2][1], boxes[2])      # bit-exact
    assert np.array_equal(got[0][2], labels[:3].astype(np.int32)) and np.array_equal(got[2][2], labels[3:].astype(np.int32))


class _FakeStage:
    device = None
    alphabet = "0123456789abcdefghijklmnopqrstuvwxyz"


class _FakePipeline:
    """Stands in for the GPU pipeline: deterministic function of the image content."""
    detector = _FakeStage()
    recognizer = _FakeStage()

    def recognize(self, images):
        out = []
        for im in images:
            k = int(im[0, 0, 0]) % 4
            out.append([("w%dx%d" % (int(im[0, 0, 0]), j), np.full((4, 2), float(im[0, 0, 0]) + j, np.float32)) for j in range(k)])
        return out


def _worker(rank, world, port, images, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = D.recognize_sharded(_FakePipeline(), images, max_boxes=8)
        if rank == 0:
            ret.put([[(t, b.tolist()) for t, b in g] for g in res])
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_recognize_sharded_world2_gloo():
    images = np.zeros((7, 4, 4, 3), np.uint8)
    images[:, 0, 0, 0] = np.arange(7) + 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, images, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[(t, b.tolist()) for t, b in g] for g in _FakePipeline().recognize(images)]
    assert got == expect                       # global order preserved, payload bit-exact
