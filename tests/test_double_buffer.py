"""Host logic of deepbnd_b200.batch.DoubleBuffer (two handles, one worker thread each) without a GPU: every job
runs prepare -> solve on the handle of its parity, `store` sees the results in job order on the caller's thread,
the two workers really overlap, and an exception in a worker reaches the caller."""
import threading
import time

import pytest

from deepbnd_b200.batch import DoubleBuffer


class _Handle:
    def __init__(self, name):
        self.name, self.log = name, []

    def close(self):
        self.log.append("closed")


def test_jobs_alternate_handles_and_store_is_ordered():
    h = [_Handle("a"), _Handle("b")]
    db = DoubleBuffer(handles=h)
    main = threading.get_ident()
    stored, threads = [], set()

    def prepare(job, handle):
        handle.log.append(("prepare", job))
        return job * 10

    def solve(job, handle, prepared):
        threads.add(threading.get_ident())
        time.sleep(0.02 * (3 - job % 3))           # uneven durations: completion order != job order
        handle.log.append(("solve", job))
        return handle.name, prepared

    def store(job, res):
        assert threading.get_ident() == main       # file writes stay on the caller's thread
        stored.append((job, res))
        return job

    out = db.run(range(7), prepare, solve, store)
    assert out == list(range(7))
    assert stored == [(j, ("ab"[j % 2], 10 * j)) for j in range(7)]
    assert [j for op, j in h[0].log if op == "solve"] == [0, 2, 4, 6]
    assert [j for op, j in h[1].log if op == "solve"] == [1, 3, 5]
    assert len(threads) == 2 and main not in threads
    db.close()
    assert h[0].log[-1] == "closed" and h[1].log[-1] == "closed"


def test_workers_overlap():
    db = DoubleBuffer(handles=[_Handle("a"), _Handle("b")])
    lock = threading.Lock()
    state = {"active": 0, "peak": 0}

    def solve(job, handle, prepared):
        with lock:
            state["active"] += 1
            state["peak"] = max(state["peak"], state["active"])
        time.sleep(0.05)
        with lock:
            state["active"] -= 1

    db.run(range(6), lambda j, h: None, solve)
    assert state["peak"] == 2                      # the two handles are driven at the same time, never more than two jobs


def test_worker_exception_reaches_the_caller():
    db = DoubleBuffer(handles=[_Handle("a"), _Handle("b")])

    def solve(job, handle, prepared):
        if job == 3:
            raise ValueError("boom")
        return job

    seen = []
    with pytest.raises(ValueError, match="boom"):
        db.run(range(8), lambda j, h: None, solve, lambda j, r: seen.append(j))
    assert seen == [0, 1, 2]


def test_empty_and_single_job():
    db = DoubleBuffer(handles=[_Handle("a"), _Handle("b")])
    assert db.run([], lambda j, h: None, lambda j, h, p: j) == []
    assert db.run([5], lambda j, h: None, lambda j, h, p: j + 1) == [6]
