"""Host -> device minibatch staging for the training loop (the role torch's DataLoader(pin_memory=True)
plus `.cuda(non_blocking=True)` plays around the reference's `fit()` loop, src/DrVAE.py:774-782).

A minibatch of an ensemble step is tens of MB (32 models x 150 rows x 978 genes x 2 inputs), i.e.
~0.7 ms of PCIe time next to a ~1.3 ms step: copying it on the compute stream would serialise the
two.  DeviceFeeder keeps `depth` preallocated device slots and copies batch i+1 on its own stream
while step i runs; slot reuse is ordered with CUDA events, never with host synchronisation.

    feeder = DeviceFeeder(plan.device)
    feeder.put(host_batch_0)
    for i in range(n):
        if i + 1 < n: feeder.put(host_batch[i + 1])     # async H2D from pinned memory
        batch, slot = feeder.get()                      # compute stream waits for the copy of batch i
        losses = plan.train_step(batch, hp)
        feeder.done(slot)                               # the slot may be overwritten after this step
"""
import torch


class DeviceFeeder:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * self.depth
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.free = [None] * self.depth
        self.n_put = 0
        self.n_get = 0
        self.bytes_put = 0

    def put(self, host_batch):
        """Enqueue the copy of one minibatch (dict of CPU tensors, ideally pinned) into the next slot."""
        if self.n_put - self.n_get >= self.depth:
            raise RuntimeError("DeviceFeeder: all %d slots are in flight; call get()/done() first" % self.depth)
        s = self.n_put % self.depth
        if self.slots[s] is None or any(self.slots[s][k].shape != v.shape or self.slots[s][k].dtype != v.dtype
                                        for k, v in host_batch.items()):
            self.slots[s] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()}
        with torch.cuda.stream(self.copy_stream):
            if self.free[s] is not None:
                self.copy_stream.wait_event(self.free[s])  # the step that read this slot has finished
            for k, v in host_batch.items():
                self.slots[s][k].copy_(v, non_blocking=True)
                self.bytes_put += v.numel() * v.element_size()
            self.ready[s].record(self.copy_stream)
        self.n_put += 1

    def get(self):
        """-> (device batch, slot).  The current stream is made to wait for the slot's copy."""
        if self.n_get >= self.n_put:
            raise RuntimeError("DeviceFeeder: nothing was put()")
        s = self.n_get % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[s])
        self.n_get += 1
        return self.slots[s], s

    def done(self, slot):
        """Call after enqueueing the step that consumes `slot`."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev
