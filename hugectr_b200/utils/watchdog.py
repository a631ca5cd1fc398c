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
