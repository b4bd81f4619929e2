"""Deferred-execution executor: the reference's thread-pool fan-out as ONE GPU batch.

The reference issues every sweep as
``with ThreadPoolExecutor(max_workers=w) as ex: f = ex.submit(system, t0, tend_i, *pulses, **kw)``
followed by ``wait(futures)`` / ``f.result()`` (49 call sites; e.g.
``pyaceqd/two_time/correlations.py:153-175``).  :class:`BatchExecutor` has the same
``submit`` / context-manager surface; calls are recorded instead of run, and the whole set is
propagated in one launch when the first result is requested (or on ``__exit__`` / ``wait``).
"""
from __future__ import annotations

from typing import Callable, List

from pyaceqd_b200.general_system import general_system as _gs


class BatchFuture:
    def __init__(self, owner: "BatchExecutor", index: int):
        self._owner, self._index = owner, index
        self._callbacks: List[Callable] = []

    def result(self, timeout=None):
        self._owner.flush()
        return self._owner._results[self._index]

    def done(self):
        return self._owner._results[self._index] is not None

    def add_done_callback(self, fn):
        if self.done():
            fn(self)
        else:
            self._callbacks.append(fn)

    def exception(self, timeout=None):
        return None


class BatchExecutor:
    """Drop-in for ``concurrent.futures.ThreadPoolExecutor`` around ``system(...)`` calls."""

    def __init__(self, max_workers=None, tail_rows=0, distributed=None, tail_reduce=None, **_ignored):
        """``tail_rows`` > 0: every deferred job returns only its last ``tail_rows`` output rows
        (the workflows index results from the end, e.g. ``res[1][-n_tau:]``
        ``two_time/correlations.py:182-183``), which bounds the device->host volume of a sweep."""
        self.max_workers = max_workers
        # True: inside a torch.distributed job the sweep shards over the ranks (every rank must submit the same
        # jobs in the same order); None: only if ACEQD_DISTRIBUTED=1.  Never implicit: a script that already
        # splits its sweep per rank, or runs it on rank 0 only, must not meet a collective here.
        self.distributed = distributed
        # (pairs, spacing): every deferred job returns the trapezoid over its kept rows per output pair, reduced on the
        # device (Engine.run_jobs) -- the tau integrals of the G2 workflows without shipping the (t, tau) map
        self.tail_reduce = tail_reduce
        self.tail_rows = int(tail_rows or 0)
        self._requests = []     # (Request | immediate result, post-processing)
        self._results: list = []
        self._futures: List[BatchFuture] = []
        self._flushed_upto = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.flush()
        return False

    def shutdown(self, wait=True):
        self.flush()

    def submit_tail(self, tail_rows, fn, *args, **kwargs) -> BatchFuture:
        """``submit`` with a per-job tail length."""
        fut = self.submit(fn, *args, **kwargs)
        if self._requests and self._requests[-1][0] == fut._index:
            self._requests[-1][1].job.tail_rows = int(tail_rows or 0)
        return fut

    def submit(self, fn, *args, **kwargs) -> BatchFuture:
        """Record ``fn(*args, **kwargs)``.  A plain pass-through adapter (``tls``, ``biexciton``, ...) returns the
        deferred :class:`Request` of ``system_ace_stream`` and joins the batch.  An adapter that post-processes the
        result (unpacks it, indexes it, ...) cannot work on a placeholder: whatever it does to it -- raise or return
        something else -- the captured requests are dropped and ``fn`` runs again eagerly, so any callable the
        reference's ``ThreadPoolExecutor`` accepts is accepted here."""
        sink: list = []
        prev = getattr(_gs._capture, "sink", None)
        _gs._capture.sink = sink
        ret, failed = None, None
        try:
            ret = fn(*args, **kwargs)
        except Exception as exc:      # noqa: BLE001 - re-raised below unless a placeholder caused it
            failed = exc
        finally:
            _gs._capture.sink = prev
        fut = BatchFuture(self, len(self._results))
        self._futures.append(fut)
        self._results.append(None)
        if failed is None and isinstance(ret, _gs.Request):
            ret.job.tail_rows = self.tail_rows
            self._requests.append((len(self._results) - 1, ret))
            return fut
        if sink:      # the adapter handled (or choked on) a placeholder: run it for real
            ret = fn(*args, **kwargs)
        elif failed is not None:
            raise failed
        self._results[-1] = ret
        for cb in fut._callbacks:
            cb(fut)
        return fut

    def flush(self):
        pending = [(i, r) for (i, r) in self._requests if self._results[i] is None]
        if not pending:
            return
        res = _gs.run_requests([r for _, r in pending], distributed=self.distributed, tail_reduce=self.tail_reduce)
        for (i, _), out in zip(pending, res):
            self._results[i] = out
            for cb in self._futures[i]._callbacks:
                cb(self._futures[i])
        self._requests = []


def wait(futures, timeout=None, return_when=None):
    """``concurrent.futures.wait`` look-alike: triggers the batch."""
    for f in futures:
        if isinstance(f, BatchFuture):
            f.result()
    return set(futures), set()
