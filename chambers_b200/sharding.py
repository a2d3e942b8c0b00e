"""Batch sharding across the GPUs of one box.

Images are independent, so the batch splits into contiguous slices, one per rank, and there is no
collective on the data path.  What keeps the result independent of the GPU count is that the
schedule RNG is keyed by the GLOBAL image index: rank r passes ``image_index_base = start`` and
``batch_total = B`` (Contrast's batch-mode constant needs the whole-batch size).
"""


def shard_bounds(total, rank, world):
    """Contiguous [start, stop) of ``rank`` when ``total`` images are split over ``world`` ranks;
    the first ``total % world`` ranks take one extra image."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %r / world %r" % (rank, world))
    base, extra = divmod(int(total), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_kwargs(total, rank, world):
    """Keyword arguments a rank passes to a policy layer for its slice."""
    start, _stop = shard_bounds(total, rank, world)
    return {"image_index_base": start, "batch_total": int(total)}
