"""Host <-> device plumbing for streams of volume pairs: the copy of item i+1 runs on a side stream
while item i is being registered (two CUDA streams, event-ordered; torch owns the memory)."""
from __future__ import annotations

import torch


def prefetch_to_device(items, device):
    """Yield tuples of device tensors for an iterable of tuples of (pinned) host tensors.  The H2D
    copies of the NEXT item are enqueued on a dedicated copy stream before the current item is
    handed out, so they overlap with the kernels the caller launches on the current stream."""
    device = torch.device(device)
    copy_stream = torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)

    def stage(item):
        with torch.cuda.stream(copy_stream):
            out = tuple(t.to(device, non_blocking=True) for t in item)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return out, ev

    it = iter(items)
    try:
        nxt = stage(next(it))
    except StopIteration:
        return
    while nxt is not None:
        cur, ev = nxt
        try:
            nxt = stage(next(it))
        except StopIteration:
            nxt = None
        compute.wait_event(ev)
        for t in cur:
            t.record_stream(compute)
        yield cur
