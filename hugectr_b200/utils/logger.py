"""Logger with HugeCTR line format and env control.

Parity: HugeCTR/core23/logger.hpp:177-300, logger.cpp:102-124. Levels ERROR(-1)..TRACE(9 in ref);
``HUGECTR_LOG_LEVEL`` filters, ``HUGECTR_LOG_TO_FILE`` tees to ``hctr_<rank>.log``; ROOT messages
are printed by rank 0 only, WORLD by every rank.
"""
from __future__ import annotations

import os
import sys
import threading
import time

LEVELS = {"ERROR": -1, "SILENCE": 0, "INFO": 1, "WARNING": 2, "DEBUG": 3, "TRACE": 4}


class HctrError(RuntimeError):
    """core23::RuntimeError equivalent carrying an Error_t code."""

    def __init__(self, code, msg):
        super().__init__(f"[HCTR][{getattr(code, 'name', code)}] {msg}")
        self.code = code


class Logger:
    _inst = None

    def __init__(self):
        self.level = int(os.environ.get("HUGECTR_LOG_LEVEL", "2"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.to_file = os.environ.get("HUGECTR_LOG_TO_FILE", "0") not in ("0", "", "false")
        self._fh = open(f"hctr_{self.rank}.log", "a") if self.to_file else None
        self._t0 = time.time()
        self._lock = threading.Lock()

    @classmethod
    def get(cls) -> "Logger":
        if cls._inst is None:
            cls._inst = Logger()
        return cls._inst

    def log(self, level: str, msg: str, world: bool = False):
        lv = LEVELS[level]
        if lv > self.level:
            return
        if not world and self.rank != 0:
            return
        line = "[HCTR][%s][%s][RK%d][%s]: %s" % (
            time.strftime("%H:%M:%S") + ".%03d" % int((time.time() % 1) * 1000), level, self.rank,
            "main" if threading.current_thread() is threading.main_thread()
            else "tid #%d" % threading.get_ident(), msg)
        with self._lock:
            stream = sys.stderr if lv < 0 else sys.stdout
            print(line, file=stream, flush=True)
            if self._fh:
                self._fh.write(line + "\n")
                self._fh.flush()


def info(msg, world=False): Logger.get().log("INFO", msg, world)
def warning(msg, world=False): Logger.get().log("WARNING", msg, world)
def error(msg, world=True): Logger.get().log("ERROR", msg, world)
def debug(msg, world=False): Logger.get().log("DEBUG", msg, world)


def check(cond, code, msg):
    """HCTR_CHECK_HINT equivalent."""
    if not cond:
        raise HctrError(code, msg)


def perf_log(key: str, value=None, **meta):
    """MLPerf-style event line (solver.perf_logging; reference model.cpp:327-330,861-1004)."""
    ms = int(time.time() * 1000)
    info(f":::MLLOG {{\"time_ms\": {ms}, \"key\": \"{key}\", \"value\": {value!r}, \"metadata\": {meta!r}}}")
