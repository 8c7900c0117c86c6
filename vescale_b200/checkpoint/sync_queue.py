"""A queue whose consumers can wait until everything put so far has been PROCESSED, not merely taken (legacy
``checkpoint/utilities/sync_queue.py``).  The asynchronous checkpoint path hands write jobs to a background thread through one:
``synchronize()`` is the barrier "all pending writes of earlier saves are on storage" that the next save (and interpreter exit)
waits on."""
from __future__ import annotations

import queue
import threading
from typing import Any, Optional

__all__ = ["SynchronizedQueue"]


class SynchronizedQueue:
    def __init__(self, maxsize: int = 0):
        self._q: "queue.Queue[Any]" = queue.Queue(maxsize)
        self._cv = threading.Condition()
        self._unfinished = 0

    def put(self, item: Any, block: bool = True, timeout: Optional[float] = None) -> None:
        with self._cv:
            self._unfinished += 1
        try:
            self._q.put(item, block, timeout)
        except BaseException:
            with self._cv:
                self._unfinished -= 1
                self._cv.notify_all()
            raise

    def get(self, block: bool = True, timeout: Optional[float] = None) -> Any:
        return self._q.get(block, timeout)

    def task_done(self) -> None:
        """The consumer finished the item it got last."""
        with self._cv:
            if self._unfinished <= 0:
                raise ValueError("task_done() called more times than there were items")
            self._unfinished -= 1
            if self._unfinished == 0:
                self._cv.notify_all()

    def synchronize(self, timeout: Optional[float] = None) -> bool:
        """Block until every item put so far has been processed; ``False`` on time-out."""
        with self._cv:
            return self._cv.wait_for(lambda: self._unfinished == 0, timeout)

    def qsize(self) -> int:
        return self._q.qsize()

    def empty(self) -> bool:
        return self._q.empty()

    @property
    def pending(self) -> int:
        return self._unfinished
