"""Host <-> device streaming around the model call (pinned-memory prefetch of the CMAQ batches).

`HostPipeline` overlaps the host->device copy of batch i+1 and the device->host copy of the predictions of batch i-1 with
the kernels of batch i: copies run on a side stream, ordered against the compute stream with events, input buffers on the
device are double-buffered.  The evaluation loop of the reference (`evaluation_vit.py:236-250`) moves every batch
synchronously with `.cuda()` before calling the model; at B200 inference rates that copy (844 MB per 64-sample batch) would
otherwise cost half of the step.
"""
from __future__ import annotations

import torch

from . import _lib

PM25_CHANNELS = (4, 10, 16, 22)                         # metnet3.py:362


def pack_host(x: torch.Tensor, pm25_mean: float, pm25_std: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """Data-loader side of the streaming path: (B,T,C,H,W) fp32 host tensor -> pinned bf16 host tensor holding exactly the
    values the device `prepare` kernel derives from it in the default precision -- the PM2.5 channels standardised in
    fp32 ((x - mean) / std, metnet3.py:369-373), then everything rounded to bf16 (round-to-nearest-even).  Feeding this to
    ``HostPipeline`` halves the host->device bytes of a batch (422 instead of 844 MB at B=64) and leaves the predictions
    bit-identical (tests/test_metnet3_gpu.py::test_packed_host_batches_bit_identical)."""
    assert x.dtype == torch.float32 and not x.is_cuda and x.dim() == 5
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16).pin_memory()
    out.copy_(x)
    idx = [c for c in PM25_CHANNELS if c < x.shape[2]]
    pm = (x[:, :, idx] - torch.tensor(pm25_mean, dtype=torch.float32)) / torch.tensor(pm25_std, dtype=torch.float32)
    out[:, :, idx] = pm.to(torch.bfloat16)
    return out


class HostPipeline:
    """pipe = HostPipeline(model);  for y_host in pipe.run(batches): ...   where `batches` yields
    (x_host, timestamps_host[, out_host]) with pinned host tensors; `y_host` is a pinned (B, L, H, W) fp32 tensor that is
    valid when it is yielded (the copy has completed).  `x_host` is the reference's fp32 tensor or a bf16 batch from
    ``pack_host`` (half the host->device bytes, same predictions)."""

    def __init__(self, model, depth: int = 2):
        self.model = model
        self.depth = max(2, depth)
        self.copy_stream = None

    @torch.no_grad()
    def run(self, batches):
        dev = next(self.model.parameters()).device
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream(dev)
        copy = self.copy_stream
        slots = [dict(x=None, ts=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(self.depth)]
        for s in slots:
            s["free"].record(compute)
        pending = []                                  # (out_host, event) whose device->host copy is in flight

        def upload(k, batch):
            s = slots[k % self.depth]
            x_h, ts_h = batch[0], batch[1]
            with torch.cuda.stream(copy):
                copy.wait_event(s["free"])             # the kernels that read this slot's previous contents are done
                if s["x"] is None or s["x"].shape != x_h.shape:
                    s["x"] = torch.empty(x_h.shape, dtype=x_h.dtype, device=dev)
                    s["ts"] = torch.empty(ts_h.shape, dtype=ts_h.dtype, device=dev)
                s["x"].copy_(x_h, non_blocking=True)
                s["ts"].copy_(ts_h, non_blocking=True)
                s["ready"].record(copy)
            return s, (batch[2] if len(batch) > 2 else None)

        it = iter(batches)
        k = 0
        try:
            nxt = upload(k, next(it))
        except StopIteration:
            return
        while nxt is not None:
            s, out_h = nxt
            try:
                nxt = upload(k + 1, next(it))          # the next batch's copy overlaps this batch's kernels
            except StopIteration:
                nxt = None
            compute.wait_event(s["ready"])
            if s["x"].dtype == torch.bfloat16:
                y = self.model.forward_packed(s["x"], s["ts"])
            else:
                y = self.model(s["x"], timestamps=s["ts"])
            s["free"].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            if out_h is None:
                out_h = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                out_h.copy_(y, non_blocking=True)
                y.record_stream(copy)
                ev = torch.cuda.Event()
                ev.record(copy)
            pending.append((out_h, ev))
            while len(pending) > 1:                    # hand back results whose copy has certainly been issued one step ago
                o, e = pending.pop(0)
                e.synchronize()
                _lib.raise_device_errors()             # e.g. an out-of-range timestamp: the reference raises IndexError
                yield o
            k += 1
        for o, e in pending:
            e.synchronize()
            _lib.raise_device_errors()
            yield o
