"""Drop-in replacements for the reference's event container and voxel encoder.

Mirrors (same names, arguments, return types, assertion behaviour):
  * EventSequence                    loader/loader_utils.py:352-397 (dup utils_luo/event_utils.py:255-300)
  * EventSequenceToVoxelGrid_Pytorch utils/transformers.py:18-124 (identical copies in
                                     utils_luo/event_utils.py:145-253, loader/loader_utils.py:429-537)

The voting and normalisation run in the sm_100a kernels of csrc/voxelize.cu through the C ABI.
There is no CPU implementation here: with gpu=False the grid is still computed on the GPU and
only the *returned tensor* is moved to the CPU, which is the device contract of the reference.
"""
from __future__ import annotations

from typing import Sequence

import numpy
import torch

from . import ops


class EventSequence(object):
    """[N,4] float64 (ts, x, y, p) events + sensor size; loader/loader_utils.py:352-397."""

    def __init__(self, dataframe, params, features=None, timestamp_multiplier=None, convert_to_relative=False):
        if _is_dataframe(dataframe):
            self.feature_names = dataframe.columns.values
            self.features = dataframe.to_numpy()
        else:
            self.feature_names = numpy.array(['ts', 'x', 'y', 'p'], dtype=object)
            if features is None:
                self.features = numpy.zeros([1, 4])
            else:
                self.features = features
        self.image_height = params['height']
        self.image_width = params['width']
        if not self.is_sorted():
            self.sort_by_timestamp()
        if timestamp_multiplier is not None:
            self.features[:, 0] *= timestamp_multiplier
        if convert_to_relative:
            self.absolute_time_to_relative()

    def get_sequence_only(self):
        return self.features

    def __len__(self):
        return len(self.features)

    def __add__(self, sequence):
        return EventSequence(dataframe=None,
                             features=numpy.concatenate([self.features, sequence.features]),
                             params={'height': self.image_height, 'width': self.image_width})

    def is_sorted(self):
        return numpy.all(self.features[:-1, 0] <= self.features[1:, 0])

    def sort_by_timestamp(self):
        if len(self.features[:, 0]) > 0:
            sort_indices = numpy.argsort(self.features[:, 0])
            self.features = self.features[sort_indices]

    def absolute_time_to_relative(self):
        """Transforms absolute time to time relative to the first event."""
        start_ts = self.features[:, 0].min()
        assert (start_ts == self.features[0, 0])
        self.features[:, 0] -= start_ts


def _is_dataframe(obj) -> bool:
    try:
        import pandas
    except ImportError:  # pandas is optional: only the DataFrame constructor path needs it
        return False
    return isinstance(obj, pandas.DataFrame)


_STAGE_CHUNK_ROWS = 1 << 17      # 4 MiB of [*,4] float64 rows per staging task
_STAGE_FLUSH_ROWS = 1 << 18      # issue the host->device copy every 8 MiB staged
_stage_pool = None
_stage_threads = None


def set_staging_threads(n: int) -> None:
    """Size of the staging pool (default: min(8, usable cores - 1)).  Multi-rank launchers give every rank its share
    of the cores (eemflow_b200.dist.pin_host_threads) so the pools of different ranks do not oversubscribe the host."""
    global _stage_pool, _stage_threads
    _stage_threads = max(1, int(n))
    if _stage_pool is not None:
        _stage_pool.shutdown(wait=True)
        _stage_pool = None


def _staging_pool():
    """Worker threads for the pageable->pinned staging copy (numpy releases the GIL while it copies)."""
    global _stage_pool
    if _stage_pool is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
        n = _stage_threads if _stage_threads is not None else max(1, min(8, usable - 1))
        _stage_pool = ThreadPoolExecutor(max_workers=n, thread_name_prefix="eem-stage")
    return _stage_pool


def _is_pinned(a: numpy.ndarray) -> bool:
    """True when the array already lives in page-locked host memory (e.g. produced by a DataLoader with
    pin_memory=True, or a view of a pinned torch tensor): the DMA engine can read it directly."""
    try:
        return bool(a.flags.c_contiguous and a.dtype == numpy.float64 and torch.from_numpy(a).is_pinned())
    except (TypeError, ValueError, RuntimeError):
        return False


class _PinnedStage:
    """Grow-only pinned staging buffer for the host->device copy of event rows.

    cudaHostAlloc costs milliseconds, so the buffer is kept between calls; an event recorded after
    each async copy guards it against being refilled while the previous copy is still in flight.
    Large uploads are staged by a few threads in 4 MiB chunks and the DMA of finished chunks is issued
    while later ones are still being staged, so the link is busy during the host memcpy instead of after it.
    """

    def __init__(self):
        self.buf = None
        self.off = None
        self.done = None
        self.dev_buf = {}         # grow-only device copy per (device, stream): reuse is ordered by that stream

    def upload(self, arrays: Sequence[numpy.ndarray], device: torch.device):
        counts = [int(a.shape[0]) for a in arrays]
        total = sum(counts)
        if self.done is not None:
            self.done.synchronize()
        if self.buf is None or self.buf.shape[0] < total:
            self.buf = torch.empty((max(total, 1024), 4), dtype=torch.float64, pin_memory=True)
        if self.off is None or self.off.numel() < len(arrays) + 1:
            self.off = torch.empty(max(len(arrays) + 1, 64), dtype=torch.int64, pin_memory=True)
        host = self.buf.numpy()
        off_host = self.off.numpy()
        off_host[0] = 0
        numpy.cumsum(counts, out=off_host[1:len(arrays) + 1])
        with torch.cuda.device(device):
            off = self.off[:len(arrays) + 1].to(device, non_blocking=True)
            if total > 0 and all(_is_pinned(a) for a in arrays):
                # page-locked caller arrays: no staging pass, one DMA per window straight from the caller's memory
                key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
                buf = self.dev_buf.get(key)
                if buf is None or buf.shape[0] < total:
                    buf = self.dev_buf[key] = torch.empty((total, 4), dtype=torch.float64, device=device)
                ev = buf[:total]
                # windows that lie back to back in the caller's pinned buffer go up as ONE copy
                pos, k = 0, 0
                while k < len(arrays):
                    j, rows = k, counts[k]
                    while (j + 1 < len(arrays) and arrays[j + 1].ctypes.data == arrays[j].ctypes.data + arrays[j].nbytes):
                        j += 1
                        rows += counts[j]
                    if j > k:       # a [rows, 4] view over the merged region of the caller's buffer
                        import ctypes
                        flat = (ctypes.c_double * (rows * 4)).from_address(arrays[k].ctypes.data)
                        src = torch.from_numpy(numpy.frombuffer(flat, dtype=numpy.float64).reshape(rows, 4))
                    else:
                        src = torch.from_numpy(arrays[k])
                    ev[pos:pos + rows].copy_(src, non_blocking=True)
                    pos += rows
                    k = j + 1
            elif total <= _STAGE_CHUNK_ROWS:
                pos = 0
                for a, n in zip(arrays, counts):
                    host[pos:pos + n] = a  # astype('float') + from_numpy of the reference, in one copy
                    pos += n
                ev = self.buf[:total].to(device, non_blocking=True)
            else:
                key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
                buf = self.dev_buf.get(key)
                if buf is None or buf.shape[0] < total:
                    buf = self.dev_buf[key] = torch.empty((total, 4), dtype=torch.float64, device=device)
                ev = buf[:total]
                tasks, pos = [], 0
                for a, n in zip(arrays, counts):
                    for lo in range(0, n, _STAGE_CHUNK_ROWS):
                        hi = min(n, lo + _STAGE_CHUNK_ROWS)
                        tasks.append((pos + lo, pos + hi, a, lo, hi))
                    pos += n

                def stage(task):
                    d0, d1, a, lo, hi = task
                    host[d0:d1] = a[lo:hi]
                    return d1

                sent = 0
                for staged in _staging_pool().map(stage, tasks):      # results arrive in submission order
                    if staged - sent >= _STAGE_FLUSH_ROWS or staged == total:
                        ev[sent:staged].copy_(self.buf[sent:staged], non_blocking=True)
                        sent = staged
            self.done = torch.cuda.Event()
            self.done.record()
        return ev, off, max(counts) if counts else 0


class EventSequenceToVoxelGrid_Pytorch(object):
    """Time-bilinear polarity voxel grid; signature of utils/transformers.py:20.

    Extra keyword arguments (all optional).  The defaults keep the reference's results for every input the
    reference accepts; the one deviation is for CORRUPT input: a vote whose flat index falls outside the grid is
    dropped silently unless strict=True (the reference's index_add_ raises IndexError there).
      deterministic  sort-by-voxel mode, bit-exact against the reference's CPU result for time-sorted input
                     (unsorted windows are argsort-ed first like EventSequence does; ties between equal stamps
                     may then be ordered differently, which changes the fp32 summation order only)
      strict         raise IndexError (after a sync) if a vote fell outside the grid, like the
                     reference's index_add_ does
      compute_device CUDA device the kernels run on when gpu=False (default cuda:gpu_nr)
    """

    def __init__(self, num_bins, gpu=False, gpu_nr=0, normalize=True, forkserver=True,
                 deterministic=False, strict=False, compute_device=None):
        if forkserver:
            try:
                torch.multiprocessing.set_start_method('forkserver')
            except RuntimeError:
                pass
        self.num_bins = num_bins
        self.normalize = normalize
        self.deterministic = deterministic
        self.strict = strict
        self.compute_device = torch.device(compute_device if compute_device is not None else 'cuda:' + str(gpu_nr))
        self._stage = _PinnedStage()
        if gpu:
            self.device = torch.device('cuda:' + str(gpu_nr))
        else:
            self.device = torch.device('cpu')

    def __call__(self, event_sequence):
        """event_sequence.features: [N x 4] NumPy array, rows [timestamp, x, y, polarity] -> float32 [bins,H,W]."""
        return self.voxelize_batch([event_sequence])[0]

    def voxelize_batch(self, event_sequences):
        """Voxelize several windows of the same sensor size in one launch -> [n, bins, H, W]."""
        assert len(event_sequences) > 0
        width = event_sequences[0].image_width
        height = event_sequences[0].image_height
        arrays = []
        for seq in event_sequences:
            events = seq.features
            assert (events.shape[1] == 4)
            assert seq.image_width == width and seq.image_height == height
            if events.shape[0] == 0:
                raise IndexError("index -1 is out of bounds for dimension 0 with size 0")  # events_torch[-1, 0]
            arrays.append(events)
        assert (self.num_bins > 0)
        assert (width > 0)
        assert (height > 0)
        with torch.no_grad():
            ev, off, max_n = self._stage.upload(arrays, self.compute_device)
            dropped = torch.zeros(1, dtype=torch.int64, device=self.compute_device) if self.strict else None
            grid = ops.voxelize(ev, off, max_n, self.num_bins, height, width, normalize=self.normalize,
                                deterministic=self.deterministic, dropped=dropped)
            if self.strict and int(dropped.item()) != 0:
                raise IndexError(f"index out of range in self ({int(dropped.item())} votes fell outside the voxel grid)")
        if grid.device != self.device:
            grid = grid.to(self.device)
        return grid


    def voxelize_concat(self, groups, height, width, timestamp_multiplier=None):
        """Voxelize windows that are CONCATENATIONS of several event arrays without concatenating them on the host:
        `groups[k]` is a list of `[N_i, 4]` float64 arrays (ts, x, y, p), e.g. the four consecutive frames the dt4 loader
        joins with `pandas.concat` (loader/MVSEC.py:245-262).  The arrays go up back to back and the group boundaries
        become the window offsets; `timestamp_multiplier` (EventSequence's argument, 1e6 in the loaders) is applied on
        the device.  Equals `__call__(EventSequence(concat, params, timestamp_multiplier=..., convert_to_relative=True))`
        -- bit for bit in deterministic mode -- for arrays whose stamps are sorted across the group, which is what the
        loader produces; an unsorted group takes the reference's host path (concatenate, sort, scale)."""
        assert len(groups) > 0
        assert (self.num_bins > 0)
        assert (width > 0)
        assert (height > 0)
        arrays, totals, slow = [], [], []
        for g, parts in enumerate(groups):
            parts = [numpy.asarray(a) for a in parts]
            assert all(a.ndim == 2 and a.shape[1] == 4 for a in parts)
            n = sum(int(a.shape[0]) for a in parts)
            if n == 0:
                raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
            parts = [a for a in parts if a.shape[0] > 0]
            ordered = all(bool(numpy.all(a[:-1, 0] <= a[1:, 0])) for a in parts) and \
                all(parts[i][-1, 0] <= parts[i + 1][0, 0] for i in range(len(parts) - 1))
            if not ordered:
                slow.append(g)
            arrays.extend(parts)
            totals.append(n)
        if slow:        # EventSequence semantics for unsorted input: concatenate + sort on the host, then the usual path
            params = {'height': height, 'width': width}
            seqs = [EventSequence(None, params, features=numpy.concatenate([numpy.asarray(a) for a in parts]).astype(numpy.float64),
                                  timestamp_multiplier=timestamp_multiplier, convert_to_relative=True) for parts in groups]
            return self.voxelize_batch(seqs)
        with torch.no_grad():
            ev, _, _ = self._stage.upload(arrays, self.compute_device)
            goff = numpy.zeros(len(totals) + 1, dtype=numpy.int64)
            numpy.cumsum(totals, out=goff[1:])
            off = torch.from_numpy(goff).to(self.compute_device, non_blocking=True)
            dropped = torch.zeros(1, dtype=torch.int64, device=self.compute_device) if self.strict else None
            grid = ops.voxelize(ev, off, max(totals), self.num_bins, height, width, normalize=self.normalize,
                                deterministic=self.deterministic, dropped=dropped,
                                timestamp_multiplier=1.0 if timestamp_multiplier is None else float(timestamp_multiplier))
            if self.strict and int(dropped.item()) != 0:
                raise IndexError(f"index out of range in self ({int(dropped.item())} votes fell outside the voxel grid)")
        if grid.device != self.device:
            grid = grid.to(self.device)
        return grid

    def voxelize_columns(self, windows, height, width):
        """Voxelize event windows given as packed columns, e.g. straight from an HREM `events{1,2}.npz`
        (x, y, t [int64 ns], p in {0,1}) without building the [N,4] float64 rows first.

        Equivalent -- bit for bit in deterministic mode -- to the reference chain
        get_compressed_events -> EventSequence(timestamp_multiplier=1e6, convert_to_relative=True) -> __call__
        (loader/loader_utils.py:26-37, 352-397; loader/HREM.py:227-232), at 13 instead of 32 bytes per event
        on the host->device link and in HBM.  `windows`: sequence of dicts (or npz files) with keys x, y, t, p;
        float64 `t` is taken as the already scaled, relative stamps of EventSequence.features[:,0].
        """
        cols = []
        counts = []
        t_dtype = None
        for wdw in windows:
            t = numpy.asarray(wdw["t"])
            if t.shape[0] == 0:
                raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
            want = numpy.int64 if numpy.issubdtype(t.dtype, numpy.integer) else numpy.float64
            assert t_dtype in (None, want), "all windows must use the same time representation"
            t_dtype = want
            x, y, p = (numpy.asarray(wdw[k]) for k in ("x", "y", "p"))
            cols.append((t, x, y, p))
            counts.append(t.shape[0])
        assert (self.num_bins > 0)
        assert (width > 0)
        assert (height > 0)
        dev = self.compute_device
        with torch.no_grad():
            d_t, d_x, d_y, d_p, off, unsorted = self._column_stage().upload(cols, counts, t_dtype, dev)
            if unsorted:                                  # EventSequence.sort_by_timestamp, then stage again (rare)
                for k in unsorted:
                    t, x, y, p = cols[k]
                    order = numpy.argsort(t)
                    cols[k] = (t[order], x[order], y[order], p[order])
                d_t, d_x, d_y, d_p, off, _ = self._column_stage().upload(cols, counts, t_dtype, dev)
            dropped = torch.zeros(1, dtype=torch.int64, device=dev) if self.strict else None
            grid = ops.voxelize_soa(d_t, d_x, d_y, d_p, off, max(counts), self.num_bins, height, width,
                                    normalize=self.normalize, deterministic=self.deterministic, dropped=dropped)
            if self.strict and int(dropped.item()) != 0:
                raise IndexError(f"index out of range in self ({int(dropped.item())} votes fell outside the voxel grid)")
        if grid.device != self.device:
            grid = grid.to(self.device)
        return grid

    def _column_stage(self):
        if getattr(self, "_col_stage", None) is None:
            self._col_stage = _ColumnStage()
        return self._col_stage


def _chunk_is_sorted(t, lo: int, hi: int) -> bool:
    """EventSequence.is_sorted restricted to stamps [lo, hi) PLUS the pair straddling `lo`, so that the chunk-wise
    results over a partition of [0, n) AND together equal numpy.all(t[:-1] <= t[1:])."""
    return bool(numpy.all(t[max(lo - 1, 0):hi - 1] <= t[max(lo, 1):hi]))


class _ColumnStage:
    """Pinned staging of packed event columns (t 8 B, x/y int16, p int8), kept between calls like _PinnedStage;
    the dtype conversions of the .npz columns (uint16 -> int16, uint8 -> int8; the reference's `.long()` of
    integer pixel coordinates) happen in the same pass that fills the pinned buffers, on the staging threads."""

    _DT = (None, numpy.int16, numpy.int16, numpy.int8)

    def __init__(self):
        self.host = None
        self.off = None
        self.done = None
        self.t_dtype = None

    @staticmethod
    def _pinned_exact(window, t_dtype) -> bool:
        """All four columns already page-locked and in the device layout (t int64 / float64, x / y int16, p int8)."""
        want = (t_dtype, numpy.int16, numpy.int16, numpy.int8)
        try:
            return all(c.dtype == d and c.flags.c_contiguous and torch.from_numpy(c).is_pinned() for c, d in zip(window, want))
        except (TypeError, ValueError, RuntimeError):
            return False

    def _upload_pinned(self, cols, counts, t_dtype, device):
        """Page-locked caller columns (a loader with pin_memory=True): no staging pass -- one DMA per column and window
        straight from the caller's memory, and the sortedness scan of EventSequence.is_sorted runs on the staging
        threads WHILE the copies are on the link (the kernels that consume them are enqueued after this returns)."""
        total = sum(counts)
        tt = torch.int64 if t_dtype == numpy.int64 else torch.float64
        if self.off is None or self.off.numel() < len(counts) + 1:
            self.off = torch.empty(max(len(counts) + 1, 64), dtype=torch.int64, pin_memory=True)
        off_host = self.off.numpy()
        off_host[0] = 0
        numpy.cumsum(counts, out=off_host[1:len(counts) + 1])
        with torch.cuda.device(device):
            off = self.off[:len(counts) + 1].to(device, non_blocking=True)
            out = [torch.empty(total, dtype=d, device=device) for d in (tt, torch.int16, torch.int16, torch.int8)]
            pos = 0
            for window, n in zip(cols, counts):
                for o, col in zip(out, window):
                    o[pos:pos + n].copy_(torch.from_numpy(col), non_blocking=True)
                pos += n
            self.done = torch.cuda.Event()
            self.done.record()
        tasks = [(k, lo, min(n, lo + 4 * _STAGE_CHUNK_ROWS)) for k, n in enumerate(counts) for lo in range(0, n, 4 * _STAGE_CHUNK_ROWS)]
        scan = (lambda task: task[0] if not _chunk_is_sorted(cols[task[0]][0], task[1], task[2]) else -1)
        results = map(scan, tasks) if total <= _STAGE_CHUNK_ROWS else _staging_pool().map(scan, tasks)
        unsorted = sorted({k for k in results if k >= 0})
        return (*out, off, unsorted)

    def upload(self, cols, counts, t_dtype, device):
        total = sum(counts)
        if self.done is not None:
            self.done.synchronize()
        if total > 0 and all(self._pinned_exact(w, t_dtype) for w in cols):
            return self._upload_pinned(cols, counts, t_dtype, device)
        if self.host is None or self.host[1].shape[0] < total or self.t_dtype != t_dtype:
            cap = max(total, 1024)
            tt = torch.int64 if t_dtype == numpy.int64 else torch.float64
            self.host = [torch.empty(cap, dtype=tt, pin_memory=True), torch.empty(cap, dtype=torch.int16, pin_memory=True),
                         torch.empty(cap, dtype=torch.int16, pin_memory=True), torch.empty(cap, dtype=torch.int8, pin_memory=True)]
            self.t_dtype = t_dtype
        if self.off is None or self.off.numel() < len(counts) + 1:
            self.off = torch.empty(max(len(counts) + 1, 64), dtype=torch.int64, pin_memory=True)
        views = [h.numpy() for h in self.host]
        tasks, pos = [], 0
        for k, (window, n) in enumerate(zip(cols, counts)):
            for lo in range(0, n, 4 * _STAGE_CHUNK_ROWS):
                hi = min(n, lo + 4 * _STAGE_CHUNK_ROWS)
                tasks.append((pos + lo, pos + hi, window, lo, hi, k))
            pos += n
        unsorted = set()

        def stage(task):
            # the sortedness scan of EventSequence.is_sorted runs here, chunk by chunk on the staging threads,
            # while the chunk is in cache anyway (one single-threaded pass over 10 M stamps costs more than staging)
            d0, d1, window, lo, hi, k = task
            if not _chunk_is_sorted(window[0], lo, hi):
                unsorted.add(k)
            for view, col in zip(views, window):
                numpy.copyto(view[d0:d1], col[lo:hi], casting="unsafe")
            return d1

        off_host = self.off.numpy()
        off_host[0] = 0
        numpy.cumsum(counts, out=off_host[1:len(counts) + 1])
        with torch.cuda.device(device):
            off = self.off[:len(counts) + 1].to(device, non_blocking=True)
            if total <= _STAGE_CHUNK_ROWS:
                for task in tasks:
                    stage(task)
                out = [h[:total].to(device, non_blocking=True) for h in self.host]
            else:
                # the DMA of finished chunks is issued while later ones are still being staged
                out = [torch.empty(total, dtype=h.dtype, device=device) for h in self.host]
                sent = 0
                for staged in _staging_pool().map(stage, tasks):      # results arrive in submission order
                    if staged - sent >= 4 * _STAGE_FLUSH_ROWS or staged == total:
                        for o, h in zip(out, self.host):
                            o[sent:staged].copy_(h[sent:staged], non_blocking=True)
                        sent = staged
            self.done = torch.cuda.Event()
            self.done.record()
        return (*out, off, sorted(unsorted))
