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
        self.dense_float = bool(ap.is_dense_float or rp.float_label_dense)
        self.key_in = 8 if model.solver.i64_input_key else 4
        self.key_dtype = model.key_dtype
        self.depth = max(2, min(64, ap.num_threads * ap.num_batches_per_thread))
        self.threads = max(1, min(8, ap.num_threads))
        self.num_samples_hint = rp.num_samples if is_train else rp.eval_num_samples
        self.h = None
        self.lib = _native.host_lib()
        L = self.lib
        L.hctr_raw_open.restype = C.c_void_p
        L.hctr_raw_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong]
        L.hctr_raw_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.hctr_raw_next.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.hctr_raw_close.argtypes = [C.c_void_p]
        L.hctr_raw_num_samples.restype = C.c_longlong
        L.hctr_raw_num_samples.argtypes = [C.c_void_p]
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

    def start(self):
        if self.h is not None:
            return
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

    def read_a_batch(self):
        if self.h is None:
            self.start()
        valid = C.c_int(0)
        if getattr(self, "_last", None) is not None:     # its slot is recycled by the call below
            self._last.wait_copied()
            self._last = None
        idx = self.lib.hctr_raw_next(self.h, C.byref(valid))
        if valid.value < 0:
            return None
        lab, den, keys = self.slots[idx]
        nv = valid.value
        # global batch size seen by all ranks (last batch may be incomplete)
        self.current_batchsize = self.b * self.world if nv == self.b else \
            max(0, min(self.b * self.world, self.rank * self.b + nv)) if nv > 0 else self.rank * self.b
        self._last = HostBatch(lab, den, keys, None, nv)
        return self._last

    def stop(self):
        if self.h is not None:
            self.lib.hctr_raw_close(self.h)
            self.h = None
            self.started = False

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass
