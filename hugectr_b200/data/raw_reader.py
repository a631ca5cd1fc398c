"""RawAsync multi-hot reader (C13): native worker threads (csrc/host/raw_reader.cpp) stream fixed-size
records of this rank's slice of each global batch into pinned staging buffers; up to
``num_threads * num_batches_per_thread`` batches are in flight.  Record layout and semantics:
HugeCTR/include/data_generator.hpp:1019-1052, samples/dlrm/train.py:468, Appendix A.6 of SURVEY."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native
from .batch import HostBatch
from .readers import IDataReader


class RawSplit:
    """Device split of a raw batch (csrc/reader_split.cu): column tables + destination offsets."""

    def __init__(self, batch, label_dim, dense_dim, hot, key_in, key_dtype, dense_float, device):
        self.batch, self.label_dim, self.dense_dim = batch, label_dim, dense_dim
        self.hot, self.key_in, self.key_dtype, self.dense_float = list(hot), key_in, key_dtype, dense_float
        self.device = device
        self.sparse_cols = sum(hot)
        self.rec_bytes = (label_dim + dense_dim) * 4 + self.sparse_cols * key_in
        feat, pos, off, o = [], [], [], 0
        for f, h in enumerate(hot):
            feat += [f] * h
            pos += list(range(h))
            off.append(o)
            o += batch * h
        self.total_keys = o
        i32 = dict(dtype=torch.int32, device=device)
        self.col_feat, self.col_pos = torch.tensor(feat or [0], **i32), torch.tensor(pos or [0], **i32)
        self.key_off = torch.tensor(off or [0], dtype=torch.int64, device=device)
        self.hot_t = torch.tensor(hot or [1], **i32)
        self.stage = None

    def run(self, raw_host: torch.Tensor, skew: int, valid: int, label, dense, keys):
        """H2D of the raw block (async, from pinned memory) + the split kernel on the current stream"""
        n = self.batch * self.rec_bytes
        if self.stage is None:
            self.stage = torch.empty(n + 16, dtype=torch.uint8, device=self.device)
        if not hasattr(self, "_lib"):
            l = _native.cuda_lib()
            vp, i = C.c_void_p, C.c_int
            l.hctr_raw_split.argtypes = [vp] * 8 + [i] * 9 + [vp]
            l.hctr_raw_split.restype = i
            self._lib = l
        nv = max(valid, 0)
        self.stage[:nv * self.rec_bytes].copy_(raw_host[skew:skew + nv * self.rec_bytes], non_blocking=True)
        rc = self._lib.hctr_raw_split(self.stage.data_ptr(), label.data_ptr(), dense.data_ptr() if dense is not None and dense.numel() else 0,
                                      keys.data_ptr(), self.col_feat.data_ptr(), self.col_pos.data_ptr(),
                                      self.key_off.data_ptr(), self.hot_t.data_ptr(), self.batch, nv, self.label_dim,
                                      self.dense_dim, self.sparse_cols, self.rec_bytes, self.key_in,
                                      8 if self.key_dtype == torch.int64 else 4, int(self.dense_float),
                                      torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            raise RuntimeError("hctr_raw_split failed")
        from ..ops import dense as D
        D._count()


class RawAsyncReader(IDataReader):
    def __init__(self, model, is_train: bool):
        rp = model.reader_params
        b = model.b_train if is_train else model.b_eval
        super().__init__(b, model.comm.rank, model.world, repeat=model.solver.repeat_dataset)
        inp = model.input
        self.path = rp.source[0] if is_train else rp.eval_source
        self.hot = [s * h for (_, s, h, _) in model.layout.blocks]
        self.label_dim, self.dense_dim = inp.label_dim, inp.dense_dim
        ap = rp.async_param
        # flag word of the native readers: bit 0 = dense features stored as floats (AsyncParam.is_dense_float,
        # the MLPerf raw format), bit 1 = labels stored as floats too (DataGenerator float_label_dense files);
        # integer labels are cast, integer dense features go through log(x + 1) (split_batch.cu:43-88)
        self.dense_float = (1 if (ap.is_dense_float or rp.float_label_dense) else 0) | \
            (2 if rp.float_label_dense else 0)
        self.key_in = 8 if model.solver.i64_input_key else 4
        self.key_dtype = model.key_dtype
        self.depth = max(2, min(64, ap.num_threads * ap.num_batches_per_thread))
        self.threads = max(1, min(8, ap.num_threads))
        self._threads_cfg = self.threads
        self.num_samples_hint = rp.num_samples if is_train else rp.eval_num_samples
        self.h = None
        self.lib = _native.host_lib()
        # device-split mode (default on CUDA): workers move raw bytes with O_DIRECT into pinned slots, ONE H2D
        # per batch, a device kernel splits label / dense / keys.  HCTR_RAW_READER=host keeps the CPU split.
        import os
        dev = getattr(model, "device", None) or torch.device("cpu")
        self.device_split = (dev.type == "cuda" and os.environ.get("HCTR_RAW_READER", "device") == "device")
        self.model_device = dev
        self.use_direct = os.environ.get("HCTR_RAW_ODIRECT", "1") == "1"
        L = self.lib
        L.hctr_rawd_open.restype = C.c_void_p
        L.hctr_rawd_open.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int]
        L.hctr_rawd_slot_bytes.restype = C.c_longlong
        L.hctr_rawd_slot_bytes.argtypes = [C.c_void_p]
        L.hctr_rawd_is_direct.argtypes = [C.c_void_p]
        L.hctr_rawd_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.hctr_rawd_next.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hctr_rawd_close.argtypes = [C.c_void_p]
        L.hctr_raw_open.restype = C.c_void_p
        L.hctr_raw_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong]
        L.hctr_raw_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.hctr_raw_next.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.hctr_raw_close.argtypes = [C.c_void_p]
        L.hctr_raw_num_samples.restype = C.c_longlong
        L.hctr_raw_num_samples.argtypes = [C.c_void_p]
        if self.device_split:
            # byte movers: each worker has ONE positional read in flight, so the IO depth the reference gets
            # from libaio (io_depth requests per thread) comes from a few more threads here
            self.threads = max(4, self._threads_cfg)
            self.split = RawSplit(b, self.label_dim, self.dense_dim, self.hot, self.key_in, self.key_dtype,
                                  self.dense_float, dev)
        else:
            self._alloc()

    def _alloc(self):
        b = self.b
        nk = b * sum(self.hot)
        pin = torch.cuda.is_available()
        self.slots = []
        for _ in range(self.depth):
            lab = torch.zeros(b, self.label_dim)
            den = torch.zeros(b, max(self.dense_dim, 1))[:, :self.dense_dim].contiguous() \
                if self.dense_dim == 0 else torch.zeros(b, self.dense_dim)
            keys = torch.zeros(max(nk, 1), dtype=self.key_dtype)
            if pin:
                lab, den, keys = lab.pin_memory(), den.pin_memory(), keys.pin_memory()
            self.slots.append((lab, den, keys))

    def _start_device(self):
        sp = self.split
        self.h = self.lib.hctr_rawd_open(self.path.encode(), sp.rec_bytes, self.b * self.world, self.b, self.rank,
                                         int(self.repeat), int(self.num_samples_hint), int(self.use_direct))
        if not self.h:
            raise FileNotFoundError(self.path)
        self.o_direct = bool(self.lib.hctr_rawd_is_direct(self.h))
        nb = int(self.lib.hctr_rawd_slot_bytes(self.h))
        if not getattr(self, "raw_slots", None) or self.raw_slots[0].numel() != nb:
            self.raw_slots = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
        bufs = (C.c_void_p * self.depth)(*[t.data_ptr() for t in self.raw_slots])
        self.lib.hctr_rawd_start(self.h, self.threads, self.depth, bufs)
        self.started = True

    def start(self):
        if self.h is not None:
            return
        if self.device_split:
            return self._start_device()
        hot = (C.c_int * len(self.hot))(*self.hot)
        self.h = self.lib.hctr_raw_open(self.path.encode(), self.label_dim, self.dense_dim, hot,
                                        len(self.hot), self.key_in, 8 if self.key_dtype == torch.int64 else 4,
                                        int(self.dense_float), self.b * self.world, self.b, self.rank,
                                        int(self.repeat), int(self.num_samples_hint))
        if not self.h:
            raise FileNotFoundError(self.path)
        vp = C.c_void_p
        n = self.depth
        la = (vp * n)(*[s[0].data_ptr() for s in self.slots])
        de = (vp * n)(*[s[1].data_ptr() for s in self.slots])
        ke = (vp * n)(*[s[2].data_ptr() for s in self.slots])
        self.lib.hctr_raw_start(self.h, self.threads, n, la, de, ke)
        self.started = True

    def set_source(self, source=None):
        if source:
            self.path = source if isinstance(source, str) else source[0]
        self.stop()
        self.start()

    def _global_valid(self) -> int:
        """valid samples of the GLOBAL batch just handed out -- the same number on every rank (a rank whose slice is
        full cannot tell from its own count that the batch is the incomplete last one of the epoch; ranks disagreeing
        on that would skip / train different steps and hang in the next collective)"""
        gb = self.b * self.world
        if getattr(self, "_nsamp", None) is None:          # (reset by stop(): a new handle starts a new epoch)
            f = self.lib.hctr_rawd_num_samples if self.device_split else self.lib.hctr_raw_num_samples
            f.restype, f.argtypes = C.c_longlong, [C.c_void_p]
            self._nsamp, self._bi = int(f(self.h)), 0
        bpe = max(1, -(-self._nsamp // gb))
        i = self._bi % bpe
        self._bi += 1
        return int(min(gb, self._nsamp - i * gb))

    def read_a_batch(self):
        if self.h is None:
            self.start()
        valid = C.c_int(0)
        if getattr(self, "_last", None) is not None:     # its slot is recycled by the call below
            self._last.wait_copied()
            self._last = None
        if self.device_split:
            skew = C.c_int(0)
            idx = self.lib.hctr_rawd_next(self.h, C.byref(valid), C.byref(skew))
            if valid.value < 0:
                return None
            nv = valid.value
            self.current_batchsize = self._global_valid()
            self._last = HostBatch(None, None, None, None, nv, raw=self.raw_slots[idx], raw_skew=skew.value,
                                   splitter=self.split)
            return self._last
        idx = self.lib.hctr_raw_next(self.h, C.byref(valid))
        if valid.value < 0:
            return None
        lab, den, keys = self.slots[idx]
        nv = valid.value
        self.current_batchsize = self._global_valid()
        self._last = HostBatch(lab, den, keys, None, nv)
        return self._last

    def stop(self):
        if self.h is not None:
            (self.lib.hctr_rawd_close if self.device_split else self.lib.hctr_raw_close)(self.h)
            self.h = None
            self.started = False
            self._nsamp = None

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass
