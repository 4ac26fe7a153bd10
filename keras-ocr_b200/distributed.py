"""Multi-GPU ``Pipeline.recognize``: one process per GPU, images sharded by batch, ONE gather.

The reference is single-process; images are fully independent in
``Pipeline.recognize`` (reference pipeline.py:28-75), so the path shards with no data-path
collective.  Each rank runs the whole pipeline on its contiguous slice and the per-image result
records -- ``count`` (int32), ``boxes`` (M,4,2) float32, ``labels`` (M,48) int8 -- are gathered to
rank 0 with a single ``torch.distributed.gather`` over NCCL/NVLink (``gloo`` in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

STEPS = 48


def shard_bounds(n_items, world_size, rank):
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank`` (first ``n_items % world`` ranks get one more)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records(counts, boxes, labels, per_rank, max_boxes):
    """Fixed-size record block for one rank: float32 tensor (per_rank, 1 + max_boxes*8 + max_boxes*12).

    labels (int8, 48 per word) are bit-packed 4 per float32 slot so that a single dtype travels.
    counts (n,), boxes (n,M,4,2) float32, labels (sum(counts),48) int -> one contiguous CPU tensor.
    """
    n = len(counts)
    rec = np.zeros((per_rank, 1 + max_boxes * 8 + max_boxes * (STEPS // 4)), dtype=np.float32)
    lab8 = np.full((per_rank, max_boxes, STEPS), -1, dtype=np.int8)
    start = 0
    for i in range(n):
        c = min(int(counts[i]), max_boxes)
        rec[i, 0] = c
        rec[i, 1:1 + c * 8] = np.asarray(boxes[i][:c], dtype=np.float32).reshape(-1)
        if c:
            lab8[i, :c] = np.asarray(labels[start:start + c], dtype=np.int8)
        start += int(counts[i])
    rec[:, 1 + max_boxes * 8:] = lab8.reshape(per_rank, -1).view(np.float32)
    rec[n:, 0] = -1                                    # padding rows of a short last shard
    return torch.from_numpy(rec)


def unpack_records(rec, max_boxes):
    """Inverse of pack_records for one rank's block -> list of (count, boxes (c,4,2), labels (c,48))."""
    rec = rec.cpu().numpy()
    out = []
    for row in rec:
        c = int(row[0])
        if c < 0:
            continue
        boxes = row[1:1 + c * 8].reshape(c, 4, 2).copy()
        lab = np.ascontiguousarray(row[1 + max_boxes * 8:]).view(np.int8).reshape(max_boxes, STEPS)[:c].astype(np.int32)
        out.append((c, boxes, lab))
    return out


def gather_records(local, world_size, rank, device=None):
    """The single collective: gather every rank's record block to rank 0.  Returns the list of blocks
    on rank 0, None elsewhere."""
    if world_size == 1:
        return [local]
    t = local.to(device) if device is not None else local
    blocks = [torch.empty_like(t) for _ in range(world_size)] if rank == 0 else None
    dist.gather(t, gather_list=blocks, dst=0)
    return blocks


def recognize_sharded(pipeline, images, max_boxes=128):
    """Run ``pipeline.recognize`` on this rank's shard of ``images`` and gather to rank 0.

    Returns, on rank 0, the same list-of-lists as ``Pipeline.recognize`` for ALL images (global
    order); ``None`` on the other ranks.  Boxes are in source-image pixels.
    """
    from . import recognition

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(images), world, rank)
    per_rank = shard_bounds(len(images), world, 0)[1]
    local = pipeline.recognize(images[lo:hi]) if hi > lo else []
    alphabet = pipeline.recognizer.alphabet
    counts = [len(g) for g in local]
    boxes = [np.array([b for _, b in g], dtype=np.float32).reshape(-1, 4, 2) for g in local]
    labels = np.full((sum(counts), STEPS), -1, dtype=np.int8)
    k = 0
    for g in local:
        for text, _ in g:
            labels[k, :len(text)] = [alphabet.index(ch) for ch in text]
            k += 1
    device = pipeline.detector.device if dist.is_initialized() and dist.get_backend() == "nccl" else None
    blocks = gather_records(pack_records(counts, boxes, labels, per_rank, max_boxes), world, rank, device)
    if rank != 0:
        return None
    out = []
    for block in blocks:
        for c, bx, lab in unpack_records(block, max_boxes):
            texts = recognition.labels_to_text(lab, alphabet)
            out.append(list(zip(texts, bx)))
    return out
