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

``FeatureCache`` + ``StreamedCascade.run_scan`` add the first "next" row of SURVEY.md §8f: inside a
scan every image is the reference view once and a source view of ~4 neighbours
(datasets/general_eval.py:71), so the reference re-extracts (and this engine would re-upload) each
view's feature maps ~5 times.  The cache keeps per-view feature maps resident in HBM (49 DTU views
x 106 MB = 5.2 GB of the 180 GB) in the CHANNELS-LAST layout the cost-volume kernels sample (a view is re-laid
out once, when it is uploaded), and a sample's views are addressed in place through their pool slots
(``view_slots`` of mvs_cost_volume_cl_*): nothing is gathered or copied on the device.
"""
from collections import OrderedDict

import time

import torch


class ScanSample:
    """One reference view of a scan: ``view_ids[0]`` is the reference view, the rest its source
    views; ``load(view_id)`` returns that view's PINNED host features ``{"stageK": [C,h,w]}`` and is
    only called for views that are not resident; cameras / depth range as in ``StageNet.forward``
    (``proj_matrices["stageK"]`` is ``[1,V,2,4,4]`` in ``view_ids`` order)."""

    def __init__(self, view_ids, load, proj_matrices, depth_values):
        self.view_ids = list(view_ids)
        self.load = load
        self.proj_matrices = proj_matrices
        self.depth_values = depth_values


class FeatureCache:
    """Slot allocator (LRU) over per-stage device pools ``[capacity, C, h, w]``.  Pure bookkeeping:
    the pools are created by ``StreamedCascade`` on first use, so this class is testable on CPU."""

    def __init__(self, capacity):
        if capacity < 2:
            raise ValueError("FeatureCache needs at least 2 slots")
        self.capacity = capacity
        self.slot_of = OrderedDict()          # view id -> slot, least recently used first
        self.free = list(range(capacity - 1, -1, -1))
        self.hits = 0
        self.misses = 0

    def lookup(self, view_id):
        """Slot of a resident view (and mark it most recently used), else None."""
        slot = self.slot_of.get(view_id)
        if slot is not None:
            self.slot_of.move_to_end(view_id)
            self.hits += 1
        return slot

    def reserve(self, view_id, pinned=()):
        """Slot for a view that is about to be uploaded; evicts the least recently used view that
        is not in ``pinned``.  Returns (slot, evicted_view_or_None)."""
        self.misses += 1
        evicted = None
        if self.free:
            slot = self.free.pop()
        else:
            for cand in self.slot_of:
                if cand not in pinned:
                    evicted = cand
                    break
            if evicted is None:
                raise RuntimeError("FeatureCache: capacity %d is too small for the views pinned by one sample" % self.capacity)
            slot = self.slot_of.pop(evicted)
        self.slot_of[view_id] = slot
        return slot, evicted


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


class CascadeLanes:
    """Round-robin of reference views over ``lanes`` CUDA streams of one GPU.  A reference view's cascade is a strict chain
    of 70 launches, a third of which (stage 1-2 layers, heads, projections) cannot fill 148 SMs; with two views in flight
    the narrow kernels of one run under the wide kernels of the other.  Views are independent (no data is shared but the
    read-only weights), so the results are those of sequential calls."""

    def __init__(self, net, device, lanes=2):
        self.net = net
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self._n = 0

    def next_stream(self):
        """The lane of the next view; it is made to wait for everything enqueued on the caller's stream so far."""
        lane = self.streams[self._n % len(self.streams)]
        lane.wait_stream(torch.cuda.current_stream(self.device))
        self._n += 1
        return lane

    def submit(self, *args, **kwargs):
        """``net(*args, **kwargs)`` on the next lane; returns (outputs, lane).  The outputs are ordered on ``lane``: use them
        there, wait for an event recorded on it, or call ``join()``."""
        lane = self.next_stream()
        with torch.cuda.stream(lane):
            out = self.net(*args, **kwargs)
        if self._n == 1:
            # first call: the BN-folded / operand-packed weights were just built on this lane; the other lanes read them
            torch.cuda.current_stream(self.device).wait_stream(lane)
            torch.cuda.synchronize(self.device)
        return out, lane

    def join(self):
        """The caller's stream waits for every lane."""
        main = torch.cuda.current_stream(self.device)
        for lane in self.streams:
            main.wait_stream(lane)


class StreamedCascade:
    def __init__(self, net, device, tmp, ring=None, lanes=1):
        """``lanes`` > 1: consecutive reference views alternate between that many compute streams (CascadeLanes);
        ``ring`` = results in flight before the oldest is handed out (default lanes + 1, so ``lanes`` cascades stay enqueued
        while the host prepares the next)."""
        self.net = net
        self.device = torch.device(device)
        self.tmp = tmp
        self.lanes = CascadeLanes(net, self.device, lanes) if lanes > 1 else None
        ring = ring if ring is not None else (lanes + 1 if lanes > 1 else 2)
        self.copy_stream = torch.cuda.Stream(device=self.device)      # host -> device (+ re-layout of a new view)
        self.d2h_stream = torch.cuda.Stream(device=self.device)       # device -> host: the other copy engine, off the compute stream
        self.ring = ring
        self._out = None
        self._dev_ring = None            # [(device flat buffer, "consumed" event)] for PackedSample uploads
        self._up = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.cache = None                # FeatureCache bookkeeping (run_scan)
        self._pools = None               # {"stageK": [capacity, h, w, C]} channels-last device pools
        self._staging = None             # {"stageK": [C, h, w]} landing buffers of an upload (re-laid out into the pool)
        self._compute_done = None        # event: the last enqueued cascade has read the pools
        self._slot_used = {}             # slot -> event of the last cascade that read it
        self._small = None               # ring of device sets (cameras, depth range) of run_scan
        self._small_i = 0
        self._lanes_warm = False
        self.host_wait_s = 0.0           # time the host spent blocked on results (≈ 0 means the host, not the GPU, paces the loop)
        self.host_steps = 0

    def _upload(self, sample):
        """sample = PackedSample, or (features dict, proj_matrices dict, depth_values) of pinned host tensors."""
        main = torch.cuda.current_stream(self.device)
        if isinstance(sample, PackedSample):
            # persistent device staging ring (no allocator traffic): a slot is overwritten only after the
            # compute that read it (ring + 1 uploads ago) has been enqueued and finished
            n = sample.flat.numel()
            if self._dev_ring is None or self._dev_ring[0][0].numel() != n:
                self._dev_ring = [[torch.empty(n, dtype=torch.float32, device=self.device), None] for _ in range(self.ring + 1)]
            slot = self._dev_ring[self._up % len(self._dev_ring)]
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

    # -- scan mode: per-view feature cache ----------------------------------------------------------
    def _stage_scan(self, sample):
        """Upload the views of ``sample`` that are not resident (copy stream: H2D into a staging buffer, then one
        re-layout launch into the view's pool slot) and return the slots.  Slots of views this sample uses are pinned
        against eviction."""
        from . import engine
        cache = self.cache
        slots, uploaded = [], 0
        with torch.cuda.stream(self.copy_stream):
            for vid in sample.view_ids:
                slot = cache.lookup(vid)
                if slot is None:
                    slot, evicted = cache.reserve(vid, pinned=sample.view_ids)
                    feats = sample.load(vid)
                    keys = sorted(feats)
                    if self._pools is None:
                        self._pools = {k: torch.empty((cache.capacity,) + tuple(feats[k].shape[1:]) + (feats[k].shape[0],),
                                                      dtype=torch.float32, device=self.device) for k in keys}
                        self._staging = {k: torch.empty(tuple(feats[k].shape), dtype=torch.float32, device=self.device) for k in keys}
                    if evicted is not None and self._slot_used.get(slot) is not None:
                        # only the cascades that read this slot matter (LRU: usually finished long ago), not the one in flight
                        self.copy_stream.wait_event(self._slot_used.pop(slot))
                    for k in keys:
                        self._staging[k].copy_(feats[k], non_blocking=True)
                        uploaded += 4 * feats[k].numel()
                    for k0 in range(0, len(keys), 4):                        # NCHW -> channels-last, straight into the slot
                        engine.features_to_cl([self._staging[k] for k in keys[k0:k0 + 4]],
                                              outs=[self._pools[k][slot] for k in keys[k0:k0 + 4]])
                slots.append(slot)
            cams, dv = self._small_set(sample)
            for k, v in sample.proj_matrices.items():
                cams[k].copy_(v, non_blocking=True)
            dv.copy_(sample.depth_values, non_blocking=True)
            uploaded += sum(4 * v.numel() for v in sample.proj_matrices.values()) + 4 * sample.depth_values.numel()
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return slots, cams, dv, ready, uploaded

    def _small_set(self, sample):
        """Device buffers for a sample's cameras and depth range: ``ring + 1`` persistent sets used in rotation (allocated
        once from the compute stream's pool — no cross-stream allocator traffic in the loop).  The set handed to sample
        i+1 was last read by cascade i-ring, whose result has been synchronised on before sample i's turn."""
        shapes = {k: tuple(v.shape) for k, v in sample.proj_matrices.items()}
        shapes["dv"] = tuple(sample.depth_values.shape)
        if self._small is None or self._small[0] != shapes:
            with torch.cuda.stream(torch.cuda.current_stream(self.device)):
                sets = [({k: torch.empty(shapes[k], dtype=torch.float32, device=self.device) for k in sample.proj_matrices},
                         torch.empty(shapes["dv"], dtype=torch.float32, device=self.device)) for _ in range(self.ring + 1)]
            self._small = (shapes, sets)
        self._small_i += 1
        return self._small[1][self._small_i % (self.ring + 1)]

    def run_scan(self, samples, capacity=64, keep_cache=False):
        """Like ``run`` for ``ScanSample``s that share views: each view's features cross PCIe once
        (while resident).  Yields (depth, confidence) pinned host tensors, one step behind the GPU; a yielded pair
        stays valid until the next-but-one result is requested (``ring + 1`` pinned buffers rotate).

        View ids are the caller's and are only meaningful inside one scan (the reference numbers the views of every
        scan 0..48), so the cache is EMPTIED at the start of every call; pass ``keep_cache=True`` to continue the same
        scan across calls (same ids = same features)."""
        with torch.cuda.device(self.device):
            yield from self._run_scan(samples, capacity, keep_cache)

    def _run_scan(self, samples, capacity, keep_cache):
        main = torch.cuda.current_stream(self.device)
        if self.cache is None or self.cache.capacity != capacity:
            self.cache, self._pools, self._staging, self._compute_done = FeatureCache(capacity), None, None, None
            self._slot_used = {}
        elif not keep_cache:
            if self._compute_done is not None:
                self.copy_stream.wait_event(self._compute_done)     # pools may still be read by the last cascade
            self.cache = FeatureCache(capacity)                      # pools are reused, their contents forgotten
            self._slot_used = {}
        it = iter(samples)
        nxt = next(it, None)
        if nxt is None:
            return
        staged = self._stage_scan(nxt)
        pending = []
        i = 0
        with torch.no_grad():
            while staged is not None:
                slots, c, d, ready, uploaded = staged
                lane = self.lanes.next_stream() if self.lanes is not None else main
                lane.wait_event(ready)
                self.h2d_bytes = uploaded
                with torch.cuda.stream(lane):
                    out = self.net(None, c, d, tmp=self.tmp, pools_cl=self._pools, view_slots=slots)
                if self.lanes is not None and not self._lanes_warm:
                    torch.cuda.synchronize(self.device)             # weights packed on the first lane are read by the others
                    self._lanes_warm = True
                self._compute_done = torch.cuda.Event()
                self._compute_done.record(lane)
                for sl in slots:
                    self._slot_used[sl] = self._compute_done
                # staged only now: an upload that evicts a slot waits for the cascade enqueued above, which may read it
                nxt = next(it, None)
                staged = self._stage_scan(nxt) if nxt is not None else None  # overlaps with this view's compute
                pending.append(self._download(out, i, lane))
                del out
                if len(pending) >= self.ring:
                    od, oc, oe = pending.pop(0)[:3]
                    self._wait(oe)
                    yield od, oc
                i += 1
        if self.lanes is not None:
            self.lanes.join()
        for od, oc, oe, _, _ in pending:
            self._wait(oe)
            yield od, oc

    def _wait(self, event):
        t0 = time.perf_counter()
        event.synchronize()
        self.host_wait_s += time.perf_counter() - t0
        self.host_steps += 1

    def _download(self, out, i, lane):
        """Enqueue the device -> host copy of a finished cascade's result on the D2H stream (so it overlaps the next
        cascade instead of sitting between two of them).  The device tensors ride along in the returned tuple: they must
        stay allocated until the copy has finished, i.e. until ``done`` has been synchronised on."""
        depth, conf = out["refined_depth"], out["photometric_confidence"]
        bufs = self._ring_buffers(depth, conf)
        hd, hc, done = bufs[i % len(bufs)]
        computed = torch.cuda.Event()
        computed.record(lane)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(computed)
            hd.copy_(depth, non_blocking=True)
            hc.copy_(conf, non_blocking=True)
            done.record(self.d2h_stream)
        self.d2h_bytes = 4 * (hd.numel() + hc.numel())
        return hd, hc, done, depth, conf

    def _ring_buffers(self, depth, conf):
        if self._out is None or self._out[0][0].shape != depth.shape:
            self._out = [(torch.empty(depth.shape, dtype=torch.float32).pin_memory(),
                          torch.empty(conf.shape, dtype=torch.float32).pin_memory(), torch.cuda.Event())
                         for _ in range(self.ring + 1)]     # +1: a yielded pair survives one more next()
        return self._out

    def run(self, samples):
        """Iterates over host samples; yields (depth, confidence) pinned host tensors.  ``ring + 1`` pinned buffers
        rotate, so a yielded pair stays valid until the next-but-one result is requested; copy it (or use
        ``AsyncResultWriter.submit(..., copy=True)``) to keep it longer."""
        with torch.cuda.device(self.device):
            yield from self._run(samples)

    def _run(self, samples):
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
                lane = self.lanes.next_stream() if self.lanes is not None else main
                lane.wait_event(ready)
                with torch.cuda.stream(lane):
                    out = self.net(f, c, d, tmp=self.tmp)
                if self.lanes is not None and not self._lanes_warm:
                    torch.cuda.synchronize(self.device)             # weights packed on the first lane are read by the others
                    self._lanes_warm = True
                if slot is not None:                                         # staging slot may be refilled after this
                    slot[1] = torch.cuda.Event()
                    slot[1].record(lane)
                elif lane is not main:                                       # plain tensors from the copy stream's pool
                    for t in list(f.values()) + list(c.values()) + [d]:
                        t.record_stream(lane)
                pending.append(self._download(out, i, lane))
                del out
                if len(pending) >= self.ring:                                # hand out the oldest result
                    od, oc, oe = pending.pop(0)[:3]
                    self._wait(oe)
                    yield od, oc
                i += 1
        if self.lanes is not None:
            self.lanes.join()
        for od, oc, oe, _, _ in pending:
            self._wait(oe)
            yield od, oc
