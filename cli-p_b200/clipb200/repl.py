"""The query loop of CLI-P's query-index.py as a line-in / lines-out state machine.

Same prompt, commands, state variables and printed lines as the reference's loop
(/root/reference/query-index.py:32-154); the two hot calls inside it - encode_text at :108
and index.search at :111 - go through clipb200 (GPU).  Kept apart from stdin and from the
OpenCV viewer so that it can be driven by tests; `query-index.py` at the repository root
wires it to a terminal.

Behaviour carried over on purpose ("surface quirks"):
  * the best hit (rank 0) is never shown and k + offset + 1 results are requested (:111-115);
  * an empty line means "more results" and only works once a text query has been made (:100-103);
  * `p N` accepts 1..100, `c N` with N < 1 resets to 50, `r WxH` / anything else unsets (:46-85);
  * result lines are "{score:.4f} {id} {path}" after a "Search time: ...s" line (:113,119).
Different on purpose: asking for more results than the index holds stops at the last real row
(the reference dies on the -1 padding id), and a non-numeric `p` / `c` / `i` argument prints a
message instead of ending the session with a traceback.
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Tuple

import numpy as np

PROMPT = "[h,q,i,r,a,c,p] >>> "

HELP = ("Enter a search query and you will receive a list of best matching\n"
        "images. The first number is the difference score, the second the\n"
        "image ID followed by the filename.\n\n"
        "Press q to stop viewing image and space for the next image.\n\n"
        "Just press enter for more results.\n\n"
        "Commands:\nq\tQuit\ni ID\tFind images similar to ID\nr [RES]\tSet maximum resolution (e.g. 1280x720)\n"
        "a\tToggle align window position\nc NUM\tSet default number of results to NUM\n"
        "p NUM\tSet number of subsets to probe (1-100, 32 default)\nh\tShow this help")


class QuerySession:
    """State of one query-index.py session.  `handle(line)` returns False when the session ends;
    everything the reference would print goes through `out` (default: print)."""

    def __init__(self, searcher, index, out: Callable[[str], None] = print,
                 show: Optional[Callable[[str, "QuerySession"], bool]] = None):
        self.searcher, self.index, self.out, self.show = searcher, index, out, show
        self.features: Optional[np.ndarray] = None
        self.have_text_query = False        # the reference's `texts is None` test
        self.k, self.offset, self.last_j = 50, 0, 0
        self.max_res: Optional[Tuple[int, int]] = None
        self.align_window = False
        self.last_rows: List[Tuple[float, int, str]] = []

    def handle(self, line: str) -> bool:
        text = line.strip()
        if text == "q":
            return False
        if text == "h":
            self.out(HELP)
            return True
        if text.startswith("p "):
            try:
                probe = int(text[2:])
            except ValueError:
                probe = 0
            if 0 < probe < 101:
                self.index.nprobe = probe           # kept for the surface; the scan is exact
                self.out(f"Set to probe {probe} subsets.")
            else:
                self.out("Invalid probe value.")
            return True
        if text == "a":
            self.align_window = not self.align_window
            self.out("Aligning window position." if self.align_window else "Not aligning window position.")
            return True
        if text.startswith("r "):
            try:
                x, y = (int(t) for t in text[2:].split("x"))
                if x > 0 and y > 0:
                    self.max_res = (x, y)
                    self.out(f"Set maximum resolution to {x}x{y}.")
                    return True
            except ValueError:
                pass
            self.max_res = None
            self.out("Unset maximum resolution.")
            return True
        if text.startswith("c "):
            try:
                k = int(text[2:])
            except ValueError:
                k = 0
            if k < 1:
                self.k = 50
                self.out("Reset number of results to 50.")
            else:
                self.k = k
                self.out(f"Showing {k} results.")
            return True
        if text.startswith("i "):
            self.offset = self.last_j = 0
            try:
                self.features = self.searcher.features_for_id(int(text[2:]))
                self.out(f"Similar to {self.searcher.path_for_id(int(text[2:]))}:")
            except Exception:
                self.out("Not found.")
                return True
        elif text == "":
            self.offset = self.last_j
            if not self.have_text_query:
                return True
        else:
            self.offset = self.last_j = 0
            self.features = self.searcher.features_for_text(text)
            self.have_text_query = True

        t0 = time.perf_counter()
        rows = self.searcher.results(self.features, k=self.k, offset=self.offset)
        self.out(f"Search time: {time.perf_counter() - t0:.4f}s")
        self.last_rows = rows
        for n, row in enumerate(rows):
            self.out(self.searcher.format_row(row))
            self.last_j = self.offset + 1 + n
            if self.show is not None and not self.show(row[2], self):
                break
        return True


def opencv_viewer():
    """The reference's image window (query-index.py:120-150) when cv2 and a display are there;
    None otherwise (results are still printed)."""
    try:
        import cv2
    except Exception:
        return None

    def show(path: str, session: QuerySession) -> bool:
        try:
            image = cv2.imread(path, cv2.IMREAD_COLOR)
            if image is None or image.shape[0] < 2:
                return True
            h, w = float(image.shape[0]), float(image.shape[1])
            if session.max_res is not None:
                scale = min(1.0, session.max_res[0] / w, session.max_res[1] / h)
                if scale < 1.0:
                    image = cv2.resize(image, (int(w * scale + 0.5), int(h * scale + 0.5)),
                                       interpolation=cv2.INTER_LANCZOS4)
            cv2.imshow("Image", image)
            if session.align_window:
                cv2.moveWindow("Image", 0, 0)
            while True:
                key = cv2.waitKey(0) & 0xFF
                if key == ord(" "):
                    return True
                if key == ord("q"):
                    cv2.destroyAllWindows()
                    return False
        except Exception:
            return True

    return show
