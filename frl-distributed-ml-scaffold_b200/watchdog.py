"""Hang guard for the minibatch loop.

The reference arms a fresh ``threading.Timer`` around every minibatch (reference
solver_worker.py:463, watchdog_timer.py:20-67) — a thread spawn per step on the hot path, and
its context manager leaks the timer when the body raises.  Same observable behaviour here
(thread stacks dumped, ``TimeoutError`` injected into the guarded thread) from ONE daemon thread
per loop that is merely ``kick()``-ed each step.
"""
import ctypes
import logging
import sys
import threading
import time
import traceback
from typing import Optional

logger = logging.getLogger("watchdog timer")


class StepWatchdog:
    def __init__(self, timeout_ms: int) -> None:
        self._timeout_s = timeout_ms / 1000.0
        self._tid = threading.get_ident()
        self._deadline = time.monotonic() + self._timeout_s
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.fired = False

    def __enter__(self) -> "StepWatchdog":
        if self._timeout_s > 0:
            self._thread = threading.Thread(target=self._run, name="frl-watchdog", daemon=True)
            self._thread.start()
        else:
            self._expire()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()

    def kick(self) -> None:
        self._deadline = time.monotonic() + self._timeout_s

    def _run(self) -> None:
        while not self._stop.is_set():
            remaining = self._deadline - time.monotonic()
            if remaining <= 0:
                self._expire()
                return
            self._stop.wait(min(remaining, 1.0))

    def _expire(self) -> None:
        self.fired = True
        logger.warning("Watchdog timer has expired;  dumping stacks.")
        for tid, frame in sys._current_frames().items():
            logger.warning("Thread %s stack:", tid)
            for line in traceback.format_stack(frame):
                logger.warning(line.rstrip())
        logger.warning("Watchdog timer sending TimeoutError into hanging thread.")
        ctypes.pythonapi.PyThreadState_SetAsyncExc(ctypes.c_long(self._tid),
                                                   ctypes.py_object(TimeoutError))
