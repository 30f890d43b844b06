"""Host -> device input prefetcher: the next batch's pinned-memory copy runs on a side stream while the current step
computes (what torch's DataLoader(pin_memory=True) + non_blocking copies give the reference's train loop,
train_larva.py:112-131 there, made explicit).  Plumbing only: no compute, no fallback."""
import torch


class DevicePrefetcher:
    """Iterate over `source` (an iterable of tuples of HOST tensors, ideally pinned) and yield tuples of device
    tensors.  `depth` batches are in flight; a slot is only overwritten after the consumer's stream has passed the
    point where the slot was handed out `depth` iterations ago."""

    def __init__(self, source, device, depth=2, defer=False):
        self.it = iter(source)
        self.defer, self._owed = bool(defer), False
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * self.depth      # device tensors per slot
        self.ready = [None] * self.depth      # copy-done events
        self.free = [None] * self.depth       # consumer-done events
        self.queue = []                       # slot indices with a copy in flight, oldest first
        self.n = 0
        for _ in range(self.depth):
            self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        s = self.n % self.depth
        self.n += 1
        with torch.cuda.stream(self.stream):
            if self.free[s] is not None:
                self.stream.wait_event(self.free[s])
            if self.slots[s] is None or any(d.shape != h.shape or d.dtype != h.dtype
                                            for d, h in zip(self.slots[s], host)):
                self.slots[s] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            for d, h in zip(self.slots[s], host):
                d.copy_(h, non_blocking=True)
            if self.ready[s] is None:
                self.ready[s] = torch.cuda.Event()      # events are re-recorded, not re-created (host time per batch)
            self.ready[s].record(self.stream)
        self.queue.append(s)

    def __iter__(self):
        return self

    def __next__(self):
        if self._owed:             # nobody ran the postponed refill (run_deferred): catch up now
            self.refill()
        if not self.queue:
            raise StopIteration
        s = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        out = self.slots[s]
        # everything the consumer enqueues from now until its NEXT call uses `out`; mark the slot free at that call
        if getattr(self, '_last', None) is not None:
            if self.free[self._last] is None:
                self.free[self._last] = torch.cuda.Event()
            self.free[self._last].record(cur)
            if self.defer:
                # the refill (~25 us of host time: stream switch, two copies, an event) waits until the consumer has
                # enqueued its step -- models/LarvaNet.py calls run_deferred() between optim.step() and loss.item(), so
                # it overlaps the GPU's work instead of delaying it; the next __next__ catches up if nobody did
                self._owed = True
                if self not in _deferred:
                    _deferred.append(self)
            else:
                self._issue()          # refill the slot handed out one call ago
        self._last = s
        return out

    def refill(self):
        if self._owed:
            self._owed = False
            self._issue()


_deferred = []


def run_deferred():
    """Issue the refills that DevicePrefetcher(defer=True) instances postponed.  Called by the train step after its GPU
    work is enqueued."""
    for p in _deferred:
        p.refill()
