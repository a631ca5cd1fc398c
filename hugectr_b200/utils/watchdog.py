"""Hang diagnosis (the reference has no failure detection): a step watchdog that dumps every
thread's Python stack (and optionally aborts the rank, so the launcher can tear the job down) when one
training / evaluation step exceeds a deadline -- typically a collective some rank never entered.

Enabled by ``HCTR_STEP_TIMEOUT=<seconds>`` (``HCTR_STEP_TIMEOUT_ABORT=1`` to exit with code 124);
``Comm.init_from_env`` passes the same value as the process-group timeout so NCCL / gloo give up too.
"""
from __future__ import annotations

import faulthandler
import os
import sys
from typing import Optional


class StepWatchdog:
    def __init__(self, timeout_s: float, abort: bool = False, file=None):
        self.timeout, self.abort = float(timeout_s), abort
        self.file = file or sys.stderr
        self.armed = False

    @staticmethod
    def from_env() -> Optional["StepWatchdog"]:
        t = float(os.environ.get("HCTR_STEP_TIMEOUT", "0") or 0)
        if t <= 0:
            return None
        return StepWatchdog(t, os.environ.get("HCTR_STEP_TIMEOUT_ABORT", "0") == "1")

    def arm(self):
        """(re)start the countdown: call at the beginning of every step"""
        faulthandler.dump_traceback_later(self.timeout, repeat=False, file=self.file, exit=self.abort)
        self.armed = True

    def disarm(self):
        if self.armed:
            faulthandler.cancel_dump_traceback_later()
            self.armed = False

    def __enter__(self):
        self.arm()
        return self

    def __exit__(self, *exc):
        self.disarm()
        return False


class ExecWatchdog:
    """GIL-independent deadline (csrc/host/exec_watchdog.cpp): unless disarmed in time, a native thread writes
    ``message`` to ``message_fd`` and either replaces the process image (``argv`` given: ``execve`` with the
    environment captured at arming time plus ``env`` overrides; the PID stays, so a launcher keeps supervising
    the worker and the old CUDA context -- with whatever kernels were spinning in it -- is torn down) or leaves
    with ``exit_code``.  One watchdog per process; arming again replaces the pending deadline."""

    @staticmethod
    def _lib():
        import ctypes as C
        from .. import _native
        lib = _native.host_lib()
        if not getattr(lib, "_hctr_wd_typed", False):
            lib.hctr_exec_watchdog_arm.argtypes = [C.c_double, C.c_char_p, C.POINTER(C.c_char_p),
                                                   C.POINTER(C.c_char_p), C.c_char_p, C.c_int, C.c_int]
            lib.hctr_exec_watchdog_arm.restype = C.c_int
            lib.hctr_exec_watchdog_disarm.restype = None
            lib.hctr_exec_watchdog_armed.restype = C.c_int
            lib._hctr_wd_typed = True
        return lib

    @classmethod
    def arm(cls, seconds: float, argv=None, env=None, unset=(), message: str = "", message_fd: int = 2,
            exit_code: int = 124):
        import ctypes as C

        def arr(items):
            a = (C.c_char_p * (len(items) + 1))()
            for i, s in enumerate(items):
                a[i] = os.fsencode(s)
            a[len(items)] = None
            return a
        e = dict(os.environ)
        e.update(env or {})
        for k in unset:
            e.pop(k, None)
        path = os.fsencode(argv[0]) if argv else None
        cls._lib().hctr_exec_watchdog_arm(float(seconds), path, arr(list(argv or [])),
                                          arr([f"{k}={v}" for k, v in e.items()]),
                                          message.encode() if message else None, int(message_fd), int(exit_code))

    @classmethod
    def disarm(cls):
        cls._lib().hctr_exec_watchdog_disarm()

    @classmethod
    def armed(cls) -> bool:
        return bool(cls._lib().hctr_exec_watchdog_armed())
