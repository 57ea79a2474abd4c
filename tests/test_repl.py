"""Host logic of the query loop (clipb200/repl.py, mirroring query-index.py:40-119) with a stub in
place of the GPU searcher: command parsing, state, printed lines."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))


class StubSearcher:
    """20 rows; scores 1.0, 0.95, ...; records what was asked."""

    def __init__(self):
        self.calls = []

    def features_for_text(self, text):
        return np.full((1, 512), len(text), np.float32)

    def features_for_id(self, i):
        if not 0 <= i < 20:
            raise KeyError(i)
        return np.full((1, 512), -i, np.float32)

    def path_for_id(self, i):
        return f"dir/img_{i}.jpg"

    def results(self, features, k, offset):
        self.calls.append((float(features[0, 0]), k, offset))
        n = min(k + offset + 1, 20)
        return [(1.0 - 0.05 * j, j, f"dir/img_{j}.jpg") for j in range(n) if j > offset]

    @staticmethod
    def format_row(row):
        return f"{row[0]:.4f} {row[1]} {row[2]}"


class StubIndex:
    nprobe = 32


def test_session_state_machine():
    from clipb200 import repl
    out = []
    st, ix = StubSearcher(), StubIndex()
    s = repl.QuerySession(st, ix, out=out.append)

    def feed(t):
        out.clear()
        return s.handle(t), list(out)

    assert feed("")[1] == [] and st.calls == []                 # "more" needs a text query first
    assert feed("i 3")[1][0] == "Similar to dir/img_3.jpg:"
    assert st.calls[-1] == (-3.0, 50, 0)
    assert feed("")[1] == []                                    # ... an `i` query does not count (reference quirk)
    assert feed("c 4")[1] == ["Showing 4 results."]
    ok, lines = feed("  cats  ")
    assert ok and st.calls[-1] == (4.0, 4, 0)                   # stripped text, k = 4, offset = 0
    assert lines[0].startswith("Search time: ") and lines[1:] == [
        "0.9500 1 dir/img_1.jpg", "0.9000 2 dir/img_2.jpg", "0.8500 3 dir/img_3.jpg", "0.8000 4 dir/img_4.jpg"]
    assert s.last_j == 4
    feed("")
    assert st.calls[-1] == (4.0, 4, 4) and s.last_j == 8        # next page: offset = last shown rank
    feed("c x")
    assert s.k == 50
    feed("")
    assert s.last_j == 19                                       # ran out of rows
    feed("dogs")
    assert st.calls[-1] == (4.0, 50, 0) and s.offset == 0       # a new query starts over
    assert feed("p 7")[1] == ["Set to probe 7 subsets."] and ix.nprobe == 7
    assert feed("p x")[1] == ["Invalid probe value."]
    assert feed("r 640x480")[1] == ["Set maximum resolution to 640x480."]
    assert feed("r 0x5")[1] == ["Unset maximum resolution."] and s.max_res is None
    assert feed("i 77")[1] == ["Not found."]
    assert feed("q") == (False, [])


def test_viewer_can_stop_the_listing():
    from clipb200 import repl
    seen = []
    s = repl.QuerySession(StubSearcher(), StubIndex(), out=lambda _: None,
                          show=lambda path, sess: (seen.append(path), len(seen) < 2)[1])
    s.handle("c 10")
    s.handle("birds")
    assert seen == ["dir/img_1.jpg", "dir/img_2.jpg"] and s.last_j == 2
    s.show = None
    s.handle("")
    assert s.offset == 2 and s.last_j == 12
