"""Corpus-scale driver: pre-processors + frame computer + post-processors over packed batches.

This is the batched counterpart of the reference's per-utterance loop in
``signals-to-torch-feat-dir`` (``command_line.py:102-136, 585-606``): utterances are packed into
length-bucketed chunks, each chunk is copied host->device from pinned memory on a copy stream
while the previous chunk is in the fused kernel and the one before that is being copied back
(three CUDA streams, double-buffered device memory), so the end-to-end rate approaches the slower
of PCIe and the kernels.

Fusion rules: ``Dither`` and ``Preemphasize`` (each at most once) in front of an STFT computer are
folded into the kernel's sample staging; ``Deltas`` along time and a trailing ``Standardize`` with
global statistics run as device passes on the resident features.  Anything else falls back to the
processors' own ``apply`` on the host arrays, after the device part.
"""

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .compute import (
    BatchLayout,
    FrameComputer,
    PackedSignals,
    ShortIntegrationFrameComputer,
    ShortTimeFourierTransformFrameComputer,
)
from .post import Deltas, PostProcessor, Standardize
from .pre import Dither, PreProcessor, Preemphasize

__all__ = ["FeaturePipeline", "shard_utterances"]


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Deal utterances to ranks so that every rank gets about the same number of samples

    Longest first, each to the currently lightest rank (greedy LPT).  Utterances are the unit of
    sharding: they are independent, so no rank ever needs another rank's samples.
    Returns, per rank, the sorted indices of its utterances.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for idx in order:
        rank = int(np.argmin(loads))
        shards[rank].append(int(idx))
        loads[rank] += lengths[idx]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


class FeaturePipeline:
    """``preprocessors -> computer -> postprocessors`` over whole batches on one GPU"""

    def __init__(
        self,
        computer: FrameComputer,
        preprocessors: Sequence[PreProcessor] = (),
        postprocessors: Sequence[PostProcessor] = (),
        seed: int = 0,
        chunk_samples: int = 1 << 26,
        post_along_time: bool = True,
    ):
        """
        Parameters
        ----------
        post_along_time
            If True (API use), ``Deltas`` filter along the time axis and run on the device.  If
            False (the reference CLI's behaviour), every post-processor is applied
            per utterance through its own ``apply(features)`` with the default ``axis=-1``.
        """
        self.computer = computer
        self.seed = int(seed)
        self.chunk_samples = int(chunk_samples)
        self._post_along_time = bool(post_along_time)
        self._fused_pre = dict(preemph=0.0, dither=0.0, dither_first=True)
        self._host_pre: List[PreProcessor] = []
        self._device_post: List[PostProcessor] = []
        self._host_post: List[PostProcessor] = []
        self._classify(list(preprocessors), list(postprocessors))

    def _classify(self, pre, post):
        kinds = [type(p) for p in pre]
        fusable = (
            isinstance(self.computer, ShortTimeFourierTransformFrameComputer)
            and all(k in (Dither, Preemphasize) for k in kinds)
            and kinds.count(Dither) <= 1
            and kinds.count(Preemphasize) <= 1
        )
        if fusable:
            for p in pre:
                if isinstance(p, Dither):
                    self._fused_pre["dither"] = float(p.coeff)
                else:
                    self._fused_pre["preemph"] = float(p.coeff)
            if kinds:
                self._fused_pre["dither_first"] = kinds[0] is Dither
        else:
            self._host_pre = pre
        device_ok = True
        for p in post:
            on_device = (
                isinstance(p, Deltas) and p.concatenate and p._target_axis in (-1, 1) and p._pad_mode == "edge"
            ) or (isinstance(p, Standardize) and p.have_stats)
            if device_ok and on_device and self._post_along_time:
                self._device_post.append(p)
            else:
                device_ok = False
                self._host_post.append(p)

    @property
    def num_coeffs(self) -> int:
        n = self.computer.num_coeffs
        for p in self._device_post:
            if isinstance(p, Deltas):
                n *= p.num_deltas + 1
        return n

    # ---- device-resident -----------------------------------------------------------------
    def run_device(self, d_signal, offsets, lengths, utt_base: int = 0):
        """Packed CUDA signal in -> ``(rows, num_coeffs)`` CUDA features + host row offsets"""
        import torch

        computer = self.computer
        if isinstance(computer, ShortTimeFourierTransformFrameComputer):
            layout = computer.plan_batch(offsets, lengths, d_signal.device, utt_base)
            feats = computer.run_batch(layout, d_signal, seed=self.seed, **self._fused_pre)
            frame_off = layout.frame_off
        else:
            feats, frame_off = computer.compute_packed_device(d_signal, offsets, lengths)
        if self._device_post and feats.shape[0]:
            feats = self._apply_device_post(feats, frame_off)
        return feats, frame_off

    def _apply_device_post(self, feats, frame_off):
        """Device-resident post-processors; Deltas directly followed by Standardize runs as one
        fused pass (the deltas are never written un-normalised)"""
        import torch

        row_off = None
        chain = list(self._device_post)
        while chain:
            p = chain.pop(0)
            if isinstance(p, Deltas):
                if row_off is None:
                    row_off = torch.from_numpy(np.ascontiguousarray(frame_off)).to(feats.device)
                if chain and isinstance(chain[0], Standardize):
                    feats = chain.pop(0).apply_device(p.lazy_device(feats, row_off))
                else:
                    feats = p.apply_device(feats, row_off)
            else:
                feats = p.apply_device(feats, out=feats)
        return feats

    # ---- host in, host out, pipelined ----------------------------------------------------
    def _chunks(self, lengths: np.ndarray) -> List[Tuple[int, int]]:
        bounds, begin, acc = [], 0, 0
        for u, n in enumerate(lengths):
            if acc and acc + n > self.chunk_samples:
                bounds.append((begin, u))
                begin, acc = u, 0
            acc += int(n)
        if begin < len(lengths) or not bounds:
            bounds.append((begin, len(lengths)))
        return bounds

    def run_host(self, packed: PackedSignals, out: Optional[np.ndarray] = None, device=None,
                 utt_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """Features of a packed host batch; returns ``(feats, frame_off)`` on the host

        ``packed.data`` (and ``out`` if given) should live in pinned memory for the copies to
        overlap with the kernels.  Utterances are processed in chunks of about ``chunk_samples``
        samples, three streams deep.
        """
        import torch

        from ._gpu import current_device

        device = current_device() if device is None else device
        if self._host_pre or self._host_post:
            raise NotImplementedError(
                "run_host covers the device-resident part only; call the pipeline object for chains "
                "with host-side processors"
            )
        counts = np.array([self.computer.num_frames(int(n)) for n in packed.lengths], dtype=np.int64)
        frame_off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=frame_off[1:])
        width = self.num_coeffs
        if out is None:
            out = np.empty((int(frame_off[-1]), width), dtype=np.float32)
        host_in = torch.from_numpy(packed.data)
        host_out = torch.from_numpy(out)
        if isinstance(self.computer, ShortTimeFourierTransformFrameComputer):
            self._run_host_stft(packed, host_in, host_out, frame_off, device, utt_base)
            return out, frame_off
        copy_in, copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        compute = torch.cuda.current_stream(device)
        in_flight = []  # keep device buffers of the last chunks alive
        chunks = self._chunks(packed.lengths)

        def start_copy(index):
            begin, end = chunks[index]
            first = int(packed.offsets[begin]) // 4 * 4
            last = int(packed.offsets[end - 1] + packed.lengths[end - 1]) if end > begin else first
            with torch.cuda.stream(copy_in):
                d_sig = host_in[first:last].to(device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_in)
            return d_sig, ready, first

        # the sample copy of chunk k+1 is enqueued BEFORE chunk k is planned: planning uploads the
        # tile table on the compute stream (behind the wait for chunk k's samples) and blocks the
        # host, which would otherwise leave the copy engine idle between chunks
        copies = [start_copy(0)] if chunks else []
        for index, (begin, end) in enumerate(chunks):
            if index + 1 < len(chunks):
                copies.append(start_copy(index + 1))
            d_sig, ready, first = copies.pop(0)
            compute.wait_event(ready)
            feats, _ = self.run_device(d_sig, packed.offsets[begin:end] - first,
                                       packed.lengths[begin:end], utt_base + begin)
            d_sig.record_stream(compute)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                host_out[int(frame_off[begin]) : int(frame_off[end])].copy_(feats, non_blocking=True)
                feats.record_stream(copy_out)
            in_flight.append((d_sig, feats))
            if len(in_flight) > 3:
                in_flight.pop(0)
        copy_out.synchronize()
        compute.synchronize()
        return out, frame_off

    def _run_host_stft(self, packed, host_in, host_out, frame_off, device, utt_base):
        """STFT fast path of ``run_host``: nothing on the host blocks between chunks.

        The tile tables of ALL chunks are built first and uploaded with one copy; samples and
        features move through three pre-allocated device buffers each (kept across calls), so
        the loop below only enqueues: H2D copy (copy-in stream) -> fused kernel [-> device
        post-processors] (current stream) -> D2H copy (copy-out stream), ordered by events.
        """
        import torch

        computer = self.computer
        chunks = self._chunks(packed.lengths)
        if not chunks or frame_off[-1] == 0:
            return
        spans = []
        for begin, end in chunks:
            first = int(packed.offsets[begin]) // 4 * 4
            spans.append((first, int(packed.offsets[end - 1] + packed.lengths[end - 1])))
        max_samples = max(last - first for first, last in spans)
        max_rows = max(int(frame_off[end] - frame_off[begin]) for begin, end in chunks)
        nbuf = 3
        key = (str(device), host_in.dtype, max_samples, max_rows)
        cache = getattr(self, "_host_buffers", None)
        if cache is None or cache[0] != key:
            cache = (key,
                     [torch.empty(max_samples, dtype=host_in.dtype, device=device) for _ in range(nbuf)],
                     [torch.empty((max_rows, computer.num_coeffs), dtype=torch.float32, device=device)
                      for _ in range(nbuf)],
                     torch.cuda.Stream(device), torch.cuda.Stream(device))
            self._host_buffers = cache
        _, d_in, d_feat, copy_in, copy_out = cache
        compute = torch.cuda.current_stream(device)
        copy_in.wait_stream(compute)
        copy_out.wait_stream(compute)
        in_free = [None] * nbuf
        out_free = [None] * nbuf
        ready = {}

        def start_copy(index):
            slot = index % nbuf
            first, last = spans[index]
            with torch.cuda.stream(copy_in):
                if in_free[slot] is not None:
                    copy_in.wait_event(in_free[slot])
                d_in[slot][: last - first].copy_(host_in[first:last], non_blocking=True)
                ready[index] = torch.cuda.Event()
                ready[index].record(copy_in)

        # the first sample copies run while the host builds the tile tables
        for index in range(min(nbuf, len(chunks))):
            start_copy(index)
        parts, counts = [], []
        for (begin, end), (first, _) in zip(chunks, spans):
            _, tiles = computer.plan_tiles(packed.offsets[begin:end] - first, packed.lengths[begin:end],
                                           device, utt_base + begin)
            parts.append(tiles)
            counts.append(len(tiles))
        starts = np.concatenate([[0], np.cumsum(counts)])
        all_tiles = np.concatenate(parts)
        d_tiles = torch.from_numpy(all_tiles.view(np.uint8)).to(device)
        tile_bytes = all_tiles.dtype.itemsize
        for index, (begin, end) in enumerate(chunks):
            slot = index % nbuf
            first, last = spans[index]
            rows = int(frame_off[end] - frame_off[begin])
            if index not in ready:
                start_copy(index)
            d_sig = d_in[slot][: last - first]
            compute.wait_event(ready.pop(index))
            if out_free[slot] is not None:
                compute.wait_event(out_free[slot])
            layout = BatchLayout(frame_off[begin : end + 1] - frame_off[begin],
                                 d_tiles[int(starts[index]) * tile_bytes : int(starts[index + 1]) * tile_bytes],
                                 counts[index], computer.num_coeffs, device)
            feats = computer.run_batch(layout, d_sig, out=d_feat[slot][:rows], seed=self.seed, **self._fused_pre)
            in_free[slot] = torch.cuda.Event()
            in_free[slot].record(compute)
            if self._device_post and rows:
                feats = self._apply_device_post(feats, layout.frame_off)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                host_out[int(frame_off[begin]) : int(frame_off[end])].copy_(feats, non_blocking=True)
                feats.record_stream(copy_out)
                out_free[slot] = torch.cuda.Event()
                out_free[slot].record(copy_out)
        copy_out.synchronize()
        compute.synchronize()
        copy_in.synchronize()

    def __call__(self, signals: Sequence[np.ndarray]) -> List[np.ndarray]:
        return self.run_list(signals)

    def run_list(self, signals: Sequence[np.ndarray], utt_base: int = 0) -> List[np.ndarray]:
        """List of 1-D arrays in, list of ``(T, C)`` float32 arrays out (host post-processors,
        if any, are applied per utterance and may change ``T`` or ``C``).  ``utt_base`` is the
        global index of the first utterance; it keys the dither stream."""
        signals = [np.asarray(s) for s in signals]
        is_stft = isinstance(self.computer, ShortTimeFourierTransformFrameComputer)
        if self._host_pre:
            signals = [self._apply_host_pre(s.astype(np.float64)) for s in signals]
        # 16-bit PCM (what wav files hold) stays 16-bit all the way to the kernel: half the bytes
        # over PCIe, and the conversion to float32 is exact
        all_pcm = bool(signals) and all(s.dtype == np.int16 for s in signals)
        dtype = np.int16 if (all_pcm and is_stft and not self._host_pre) else np.float32
        lead = self.computer.pad_left % 4 if is_stft else 0
        packed = self._pack_pinned(signals, dtype, lead)
        counts = [self.computer.num_frames(len(s)) for s in signals]
        out = self._pinned("out", np.float32, int(sum(counts)) * self.num_coeffs)
        out = out[: int(sum(counts)) * self.num_coeffs].reshape(int(sum(counts)), self.num_coeffs)
        saved = self._host_pre, self._host_post
        self._host_pre, self._host_post = [], []
        try:
            feats, frame_off = self.run_host(packed, out=out, utt_base=utt_base)
        finally:
            self._host_pre, self._host_post = saved
        feats = feats.copy()  # the pinned buffer is reused by the next batch
        per_utt = [feats[frame_off[u] : frame_off[u + 1]] for u in range(len(signals))]
        for p in self._host_post:
            per_utt = [p.apply(f) if len(f) else f for f in per_utt]
        return per_utt

    def _pinned(self, name: str, dtype, count: int) -> np.ndarray:
        """A reusable page-locked host buffer of at least `count` elements (grown geometrically):
        copies to and from it run at PCIe speed and overlap with the kernels"""
        import torch

        cache = self.__dict__.setdefault("_pinned_buffers", {})
        key = (name, np.dtype(dtype).name)
        buf = cache.get(key)
        if buf is None or buf.numel() < count:
            size = max(int(count * 1.25), 1 << 16)
            try:
                buf = torch.empty(size, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            except RuntimeError:  # pinning refused (ulimit): pageable memory still works
                buf = torch.empty(size, dtype=getattr(torch, np.dtype(dtype).name))
            cache[key] = buf
        return buf.numpy()

    def _pack_pinned(self, signals, dtype, lead: int) -> PackedSignals:
        lengths = np.array([len(s) for s in signals], dtype=np.int64)
        offsets, total = PackedSignals.layout(lengths, lead)
        data = self._pinned("in", dtype, total)[:total]
        for sig, off, n in zip(signals, offsets, lengths):
            data[off : off + n] = sig
            data[off + n : off + (n + 3) // 4 * 4] = 0  # the padding must stay finite
        data[:lead] = 0
        data[total - 4 :] = 0
        return PackedSignals(data, offsets, lengths)

    # ---- corpus-level chain with a CMVN all-reduce ----------------------------------------
    def run_corpus(self, packed: PackedSignals, cmvn: Optional[Standardize] = None, deltas: Optional[Deltas] = None,
                   group=None, device=None, utt_base: int = 0):
        """BASELINE config 5 on this rank's shard of a corpus: features (+ `deltas` along time),
        per-coefficient statistics over ALL ranks of `group`, standardised features.

        Every rank calls this with its own utterances (``shard_utterances``).  The static features
        stay resident in HBM; the statistics are accumulated on the GPU, summed with one
        ``all_reduce`` over NCCL in stream order, and applied -- the only host synchronisation is
        the final copy of the result.  Returns ``(feats, frame_off)`` on the host; `cmvn` (created
        if None) holds the global statistics afterwards, e.g. for ``cmvn.save``.
        """
        import torch

        from ._gpu import current_device

        device = current_device() if device is None else device
        if self._host_pre or self._host_post or self._device_post:
            raise NotImplementedError("run_corpus takes its post-processors as arguments")
        cmvn = Standardize() if cmvn is None else cmvn
        d_signal = torch.from_numpy(packed.data).to(device, non_blocking=True)
        static, frame_off = self.run_device(d_signal, packed.offsets, packed.lengths, utt_base)
        rows = int(frame_off[-1])
        distributed = group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())
        source = static
        if deltas is not None:
            row_off = torch.from_numpy(np.ascontiguousarray(frame_off)).to(device, non_blocking=True)
            source = deltas.lazy_device(static, row_off)
        if rows:
            cmvn.accumulate_device(source)
        elif distributed:
            cmvn.device_stats(device, source.shape[1])  # an empty shard still takes part in the reduction
        if distributed:
            cmvn.allreduce(group)
        width = source.shape[1]
        out = np.empty((rows, width), dtype=np.float32)
        if rows:
            normed = cmvn.apply_device(source)
            torch.from_numpy(out).copy_(normed)
            cmvn.check_zero_variance()
        return out, frame_off

    def _apply_host_pre(self, signal):
        for p in self._host_pre:
            signal = p.apply(signal)
        return signal
