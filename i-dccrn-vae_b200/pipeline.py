"""Throughput modes.

HostPipeline: batches that live in host memory - the host -> device copy of batch i+1 and the device -> host copy of
batch i-1 run on two copy streams while batch i computes on the caller's stream (the kernels keep the whole GPU).

StreamPipeline: consecutive batches alternate over several CUDA streams.

A single forward is a dependency chain  encoder GEMMs -> LSTM recurrence (latency-bound, few SMs) -> decoder
GEMMs, so on one stream the tensor pipe idles during the recurrence.  With two batches in flight the recurrence of
batch i overlaps the tensor-core GEMMs of batch i+1 / i-1: the GEMM kernels claim their tiles from a global counter
(any number of free SMs is used evenly) and the recurrence is configured to occupy only 48 SMs
(``lstm_ncols = 64``, one launch per layer instead of the 144-CTA wavefront kernel).
"""
import contextlib

import torch

from . import lib, ops


class StreamPipeline:
    """with pipe.next_stream(): enc(...); dec(...)   — round-robins the batches over ``n_streams`` streams."""

    def __init__(self, device, n_streams=2):
        if n_streams < 1:
            raise ValueError("n_streams must be >= 1")
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)] if n_streams > 1 else []
        self._i = 0
        self._saved = None
        self._gate = None          # event: the previous batch has reached its LSTM recurrence
        self.stagger = True

    def __enter__(self):
        # overlap-friendly kernel configuration (restored on exit)
        self._saved = (ops.LSTM_WAVE[0],)
        if self.streams:
            ops.LSTM_WAVE[0] = False
            lib.set_option("lstm_ncols", 64)
            lib.set_option("gemm_dynamic_tiles", 1)
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                s.wait_stream(cur)
        return self

    def __exit__(self, *exc):
        self.join()
        ops.LSTM_WAVE[0] = self._saved[0]
        if self.streams:
            lib.set_option("lstm_ncols", 0)
            lib.set_option("gemm_dynamic_tiles", 0)
        return False

    @contextlib.contextmanager
    def next_stream(self):
        if not self.streams:
            yield torch.cuda.current_stream(self.device)
            return
        s = self.streams[self._i % len(self.streams)]
        self._i += 1
        # stagger: a batch starts its encoder GEMMs only once the previous batch has entered its (latency-bound)
        # recurrence, so the tensor pipe always has GEMM work from one batch while another one is in its LSTM
        if self.stagger and self._gate is not None:
            s.wait_event(self._gate)
        gate = torch.cuda.Event()
        fired = [False]

        def hook():
            if not fired[0]:
                gate.record(torch.cuda.current_stream(self.device))
                fired[0] = True
        old = ops.GATE_HOOK[0]
        ops.GATE_HOOK[0] = hook
        try:
            with torch.cuda.stream(s):
                yield s
                if not fired[0]:
                    gate.record(s)
        finally:
            ops.GATE_HOOK[0] = old
        self._gate = gate

    def join(self):
        """Make the caller's current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)


class HostPipeline:
    """Enhance batches held in (pinned) host memory with the copies overlapped:

        hp = HostPipeline(lambda x: enhance(x), device)       # x: device tensor -> device tensor
        for x_host, y_host in batches:
            hp.submit(x_host, y_host)                         # returns at once; y_host is filled asynchronously
        hp.join()                                             # current stream waits for every copy; then synchronise

    ``depth`` input buffers rotate: the copy of batch i + 1 starts as soon as the compute of batch i + 1 - depth has
    finished with its buffer.  One compute stream (the caller's current stream): nothing competes with the kernels for
    SMs, only the copy engines run beside them."""

    def __init__(self, fn, device, depth=2):
        self.fn, self.device, self.depth = fn, torch.device(device), int(depth)
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self._x = [None] * self.depth
        self._free = [None] * self.depth          # event: the compute that read input buffer i is done
        self._i = 0

    def submit(self, x_host, out_host):
        i = self._i % self.depth
        self._i += 1
        cur = torch.cuda.current_stream(self.device)
        if self._x[i] is None or self._x[i].shape != x_host.shape or self._x[i].dtype != x_host.dtype:
            self._x[i] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
            self.s_in.wait_stream(cur)                         # (allocated on the compute stream)
        with torch.cuda.stream(self.s_in):
            if self._free[i] is not None:
                self.s_in.wait_event(self._free[i])
            self._x[i].copy_(x_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        cur.wait_event(ready)
        y = self.fn(self._x[i])
        done = torch.cuda.Event()
        done.record(cur)
        self._free[i] = done
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out_host.copy_(y, non_blocking=True)
        y.record_stream(self.s_out)
        return out_host

    def join(self):
        """The caller's current stream waits for every submitted copy (synchronise it to read the host buffers)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_out)
