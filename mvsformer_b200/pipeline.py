"""Host-side streaming executor for the cascade: the call a user makes to turn HOST feature maps of
a list of reference views into HOST depth maps.

The reference's driver (test.py:232-251) moves one sample to the GPU, runs the model,
synchronises and copies every output back, strictly in sequence.  Here the three legs overlap:

* a copy stream uploads the pinned inputs of reference view i+1 while view i computes,
* the compute stream runs the cascade (all kernels of libmvs_b200.so),
* the depth map and confidence of view i are downloaded into a ring of pinned buffers right
  behind the compute, and handed to the caller one step later, so the host never stalls the GPU.

Only depth and confidence cross PCIe on the way back (14 MB at 1152x1536) — not the per-stage
probability volumes the reference's ``tensor2numpy(outputs)`` drags along (test.py:251).
"""
import torch


class PackedSample:
    """One reference view's inputs packed into a single pinned host buffer, so the upload is ONE
    DMA transfer instead of one per tensor (per-copy gaps cost ~10 % of PCIe bandwidth at 531 MB)."""

    def __init__(self, features, proj_matrices, depth_values):
        parts = [("f", k, v) for k, v in features.items()] + [("c", k, v) for k, v in proj_matrices.items()] + \
                [("d", "", depth_values)]
        self.layout, total = [], 0
        for kind, key, t in parts:
            n = t.numel()
            self.layout.append((kind, key, tuple(t.shape), total, n))
            total += (n + 63) // 64 * 64                      # keep every tensor 256-byte aligned
        self.flat = torch.empty(total, dtype=torch.float32).pin_memory()
        for (kind, key, shape, off, n), (_, _, t) in zip(self.layout, parts):
            self.flat[off:off + n].copy_(t.reshape(-1).float())

    def unpack(self, flat):
        feats, cams, dv = {}, {}, None
        for kind, key, shape, off, n in self.layout:
            view = flat[off:off + n].view(shape)
            if kind == "f":
                feats[key] = view
            elif kind == "c":
                cams[key] = view
            else:
                dv = view
        return feats, cams, dv


class StreamedCascade:
    def __init__(self, net, device, tmp, ring=2):
        self.net = net
        self.device = torch.device(device)
        self.tmp = tmp
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ring = ring
        self._out = None
        self._dev_ring = None            # [(device flat buffer, "consumed" event)] for PackedSample uploads
        self._up = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload(self, sample):
        """sample = PackedSample, or (features dict, proj_matrices dict, depth_values) of pinned host tensors."""
        main = torch.cuda.current_stream(self.device)
        if isinstance(sample, PackedSample):
            # persistent device staging ring (no allocator traffic): slot i % 2 is overwritten only
            # after the compute that read it (two uploads ago) has been enqueued and finished
            n = sample.flat.numel()
            if self._dev_ring is None or self._dev_ring[0][0].numel() != n:
                self._dev_ring = [[torch.empty(n, dtype=torch.float32, device=self.device), None] for _ in range(2)]
            slot = self._dev_ring[self._up % 2]
            self._up += 1
            with torch.cuda.stream(self.copy_stream):
                if slot[1] is not None:
                    self.copy_stream.wait_event(slot[1])
                slot[0].copy_(sample.flat, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            f, c, d = sample.unpack(slot[0])
            self.h2d_bytes = 4 * n
            return f, c, d, ready, slot
        feats, cams, dv = sample
        with torch.cuda.stream(self.copy_stream):
            f = {k: v.to(self.device, non_blocking=True) for k, v in feats.items()}
            c = {k: v.to(self.device, non_blocking=True) for k, v in cams.items()}
            d = dv.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        for t in list(f.values()) + list(c.values()) + [d]:
            t.record_stream(main)                      # consumed on the compute stream
        self.h2d_bytes = sum(4 * v.numel() for v in feats.values()) + sum(4 * v.numel() for v in cams.values()) + 4 * dv.numel()
        return f, c, d, ready, None

    def _ring_buffers(self, depth, conf):
        if self._out is None or self._out[0][0].shape != depth.shape:
            self._out = [(torch.empty(depth.shape, dtype=torch.float32).pin_memory(),
                          torch.empty(conf.shape, dtype=torch.float32).pin_memory(), torch.cuda.Event())
                         for _ in range(self.ring)]
        return self._out

    def run(self, samples):
        """Iterates over host samples; yields (depth, confidence) pinned host tensors, valid until
        ``ring`` further results have been produced."""
        main = torch.cuda.current_stream(self.device)
        it = iter(samples)
        nxt = next(it, None)
        if nxt is None:
            return
        staged = self._upload(nxt)
        pending = []
        i = 0
        with torch.no_grad():
            while staged is not None:
                f, c, d, ready, slot = staged
                nxt = next(it, None)
                staged = self._upload(nxt) if nxt is not None else None      # overlaps with this view's compute
                main.wait_event(ready)
                out = self.net(f, c, d, tmp=self.tmp)
                if slot is not None:                                         # staging slot may be refilled after this
                    slot[1] = torch.cuda.Event()
                    slot[1].record(main)
                bufs = self._ring_buffers(out["refined_depth"], out["photometric_confidence"])
                hd, hc, done = bufs[i % self.ring]
                hd.copy_(out["refined_depth"], non_blocking=True)
                hc.copy_(out["photometric_confidence"], non_blocking=True)
                done.record(main)
                self.d2h_bytes = 4 * (hd.numel() + hc.numel())
                pending.append((hd, hc, done))
                if len(pending) >= self.ring:                                # hand out the oldest result
                    od, oc, oe = pending.pop(0)
                    oe.synchronize()
                    yield od, oc
                i += 1
        for od, oc, oe in pending:
            oe.synchronize()
            yield od, oc
