"""Data-parallel driver of the per-image pass (the role of `BatchProcessor`,
processing/batch_processor.py:27-658, for device-side work).

One process per GPU.  `ScoringPipeline.run_host` takes batches of same-shaped frames that live in
(pinned) host memory, streams them to the device in chunks on a copy stream while the previous
chunk is being scored on the compute stream, and reads back only the small per-image results.
Images are independent, so ranks never communicate here; the similarity stage
(utils/duplicate.py, processing/bursts.py) is the only place with a collective.
"""
from __future__ import annotations

import numpy as np

from .. import _lib, ops


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers
    allocated afterwards are local to the GPU's PCIe root (one process per GPU: without this, half of the
    ranks of an 8-GPU box stream their frames across the socket interconnect).  Returns the node id or None
    when the topology cannot be read; never raises."""
    import os
    try:
        torch = _lib.require_cuda()
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class ScoringPipeline:
    def __init__(self, scorer, chunk: int = 8):
        torch = _lib.require_cuda()
        self.scorer = scorer
        self.chunk = int(chunk)
        self.device = scorer.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = None

    def _buffers(self, h, w):
        torch = _lib.require_cuda()
        shape = (self.chunk, h, w, 3)
        if self._bufs is None or tuple(self._bufs[0].shape) != shape:
            self._bufs = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
        return self._bufs

    def run_host(self, host_batch, rgb_order=False, orientation=1):
        """host_batch: CPU uint8 tensor [n,H,W,3] (pinned for full PCIe rate); `orientation` is the EXIF code of
        the (same-shaped) frames when they are uploaded as decoded, without `exif_transpose` (the device does it,
        utils/image_loading.py:101 of the reference).  Returns a dict of
        host numpy arrays: hist256 [n,256], sums [n,4], derived [n,4], embedding [n,768],
        aesthetic_raw [n], tag_sims [n,T] (or None).  H2D and D2H happen inside this call."""
        return self.run_host_stream([host_batch], rgb_order=rgb_order, orientation=orientation)

    def run_host_stream(self, host_batches, rgb_order=False, orientation=1):
        """Like `run_host` for a sequence (or generator) of same-shaped host batches, e.g. what a loader thread hands
        over: the chunks of consecutive batches follow each other through the two device buffers without draining
        the copy / compute overlap in between, and the small results are read back once at the end."""
        torch = _lib.require_cuda()
        compute = torch.cuda.current_stream(self.device)
        outs = []
        ci = 0
        bufs = None
        for host_batch in host_batches:
            n, h, w, _ = host_batch.shape
            if bufs is None:
                bufs = self._buffers(h, w)
            elif tuple(bufs[0].shape[1:3]) != (h, w):
                raise ValueError("run_host_stream takes batches of one frame shape")
            for start in range(0, n, self.chunk):
                k = min(self.chunk, n - start)
                slot = ci & 1
                with torch.cuda.stream(self.copy_stream):
                    if ci >= 2:
                        self.copy_stream.wait_event(self._free[slot])     # compute finished with this buffer
                    bufs[slot][:k].copy_(host_batch[start:start + k], non_blocking=True)
                    self._ready[slot].record(self.copy_stream)
                compute.wait_event(self._ready[slot])
                frames = bufs[slot][:k] if orientation == 1 else ops.orient(bufs[slot][:k], orientation)
                dev = self.scorer.score_images_device(frames, rgb_order=rgb_order)
                self._free[slot].record(compute)
                outs.append(dev)
                ci += 1
        if not outs:
            raise ValueError("run_host_stream needs at least one frame")
        cat = lambda key: torch.cat([o[key] for o in outs]).cpu().numpy() if outs[0][key] is not None else None
        res = {key: cat(key) for key in ("hist256", "sums", "derived", "embedding", "aesthetic_raw", "tag_sims", "phash")}
        res["hist256"] = res["hist256"].view(np.uint32)
        if res["phash"] is not None:
            res["phash"] = res["phash"].view(np.uint64)
        return res

    @staticmethod
    def h2d_bytes(n, h, w):
        return int(n) * h * w * 3

    @staticmethod
    def d2h_bytes(n, n_tags):
        return int(n) * (256 * 4 + 4 * 8 + 4 * 8 + 768 * 4 + 4 + n_tags * 4 + 8)
