"""CPU: the N>1 host logic -- sharding, record packing and the single gather -- with world_size 2 on gloo."""
import os
import socket

import numpy as np
import pytest
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


class _FakeRecordsPipeline(_FakePipeline):
    """Also offers ``recognize_records`` (the device-record path of the real Pipeline), here built on the host."""
    calls = 0

    def recognize_records(self, images, rows=None, rec_boxes=128):
        type(self).calls += 1
        return D._host_records(_FakePipeline(), _FakePipeline().recognize(images), rows, rec_boxes)


class _FakeTwoPhasePipeline(_FakePipeline):
    """Offers records_begin / records_end like the real Pipeline (the stream decodes between the two)."""

    def records_begin(self, images, rows=None, rec_boxes=128):
        return {"images": images, "rows": rows, "rec_boxes": rec_boxes}

    def records_counts(self, state):
        return [len(g) for g in _FakePipeline().recognize(state["images"])]

    def records_end(self, state, rec_boxes=None):
        return D._host_records(_FakePipeline(), _FakePipeline().recognize(state["images"]), state["rows"],
                               state["rec_boxes"] if rec_boxes is None else rec_boxes)

    def recognize_records(self, images, rows=None, rec_boxes=128):
        return self.records_end(self.records_begin(images, rows, rec_boxes))


def _stream_worker(rank, world, port, batches, ret, two_phase):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stream = D.ShardedStream(_FakeTwoPhasePipeline() if two_phase else _FakePipeline(), max_boxes=8)
        outs = [stream.submit(b[rank * 2:(rank + 1) * 2]) for b in batches]      # two images per rank and batch
        outs.append(stream.flush())
        assert stream.flush() is None                                             # nothing left in flight
        if rank == 0:
            ret.put([None if o is None else [[(t, b.tolist()) for t, b in g] for g in o] for o in outs])
        else:
            assert all(o is None for o in outs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("two_phase", [False, True])
def test_sharded_stream_is_one_batch_deep_and_ordered(two_phase):
    """ShardedStream.submit(batch k) returns batch k-1's results on rank 0 (None first), flush() the last; global image
    order and payload are those of recognize_sharded."""
    batches = []
    for k in range(3):
        im = np.zeros((4, 4, 4, 3), np.uint8)
        im[:, 0, 0, 0] = np.arange(4) + 1 + 4 * k
        batches.append(im)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, batches, ret, two_phase)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[[(t, b.tolist()) for t, b in g] for g in _FakePipeline().recognize(b)] for b in batches]
    assert got[0] is None and got[1:] == expect


def _worker(rank, world, port, images, ret, kind="host", presharded=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe = _FakeRecordsPipeline() if kind == "records" else _FakePipeline()
        if presharded:                                  # every rank brings its own, equally long, slice
            images = images[rank * 3:(rank + 1) * 3]
        res = D.recognize_sharded(pipe, images, max_boxes=8, presharded=presharded)
        if kind == "records":
            assert _FakeRecordsPipeline.calls == 1
        if rank == 0:
            ret.put([[(t, b.tolist()) for t, b in g] for g in res])
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,presharded", [("host", False), ("records", False), ("records", True)])
def test_recognize_sharded_world2_gloo(kind, presharded):
    images = np.zeros((7, 4, 4, 3), np.uint8)
    images[:, 0, 0, 0] = np.arange(7) + 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, images, ret, kind, presharded)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = images[:6] if presharded else images     # two ranks x three images when every rank brings its own
    expect = [[(t, b.tolist()) for t, b in g] for g in _FakePipeline().recognize(covered)]
    assert got == expect                       # global order preserved, payload bit-exact


def test_unpack_blocks_matches_per_block_unpack():
    rng = np.random.default_rng(1)
    blocks, expect = [], []
    for per_rank, counts in ((3, [2, 0, 8]), (3, [1, 5])):       # second block is a short shard (one padding row)
        boxes = [rng.uniform(0, 900, (c, 4, 2)).astype(np.float32) for c in counts]
        labels = rng.integers(-1, 37, (sum(counts), 48)).astype(np.int8)
        blocks.append(D.pack_records(counts, boxes, labels, per_rank, 8))
        expect += D.unpack_records(blocks[-1], 8)
    counts, boxes, labels = D.unpack_blocks(blocks, 8)
    assert counts.tolist() == [2, 0, 8, 1, 5] == [e[0] for e in expect]
    assert np.array_equal(boxes, np.concatenate([e[1] for e in expect]))
    assert np.array_equal(labels.astype(np.int32), np.concatenate([e[2] for e in expect]))


# ---------------------------------------------------------------------------- dense pages: more words than a record holds
class _DensePipeline(_FakeTwoPhasePipeline):
    """Image value v -> v words (so one image can exceed any fixed record size)."""

    def recognize(self, images):
        return [[("w%d" % j, np.full((4, 2), float(j), np.float32)) for j in range(int(im[0, 0, 0]))] for im in images]

    def records_counts(self, state):
        return [len(g) for g in self.recognize(state["images"])]

    def records_end(self, state, rec_boxes=None):
        return D._host_records(self, self.recognize(state["images"]), state["rows"],
                               state["rec_boxes"] if rec_boxes is None else rec_boxes)


def _dense_worker(rank, world, port, images, ret, max_boxes, two_phase):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe = _DensePipeline()
        if not two_phase:                                # duck-typed pipeline: only recognize()
            pipe = type("P", (), {"detector": _FakeStage(), "recognizer": _FakeStage(), "recognize": _DensePipeline.recognize})()
        try:
            res = D.recognize_sharded(pipe, images, max_boxes=max_boxes)
            out = None if res is None else [[(t, b.tolist()) for t, b in g] for g in res]
        except D.RecordOverflow as exc:
            out = "overflow: " + str(exc)
        if rank == 0:
            ret.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("max_boxes,two_phase", [(128, True), ("auto", True), ("auto", False)])
def test_more_words_than_a_record_holds_is_never_silent(max_boxes, two_phase):
    """An image with more words than ``max_boxes`` (the 128-word cap used to drop words without a signal):
    a fixed record size makes rank 0 raise RecordOverflow naming the image; ``max_boxes='auto'`` agrees on a record
    size across the ranks (one all-reduce) and returns every word, as single-GPU ``Pipeline.recognize`` does."""
    images = np.zeros((4, 4, 4, 3), np.uint8)
    images[:, 0, 0, 0] = [3, 150, 0, 7]                 # image 1 (rank 0's shard) has 150 words
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_dense_worker, args=(r, 2, port, images, ret, max_boxes, two_phase)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if max_boxes == "auto":
        expect = [[(t, b.tolist()) for t, b in g] for g in _DensePipeline().recognize(images)]
        assert got == expect and len(got[1]) == 150
    else:
        assert isinstance(got, str) and "150 words" in got and "max_boxes=128" in got


def test_unpack_blocks_overflow_strict_and_capped():
    boxes = [np.zeros((5, 4, 2), np.float32)]
    rec = D.pack_records([5], boxes, np.zeros((5, 48), np.int8), 1, 4)
    with pytest.raises(D.RecordOverflow):
        D.unpack_blocks([rec], 4)
    counts, b, l = D.unpack_blocks([rec], 4, strict=False)
    assert counts.tolist() == [4] and b.shape == (4, 4, 2) and l.shape == (4, 48)
