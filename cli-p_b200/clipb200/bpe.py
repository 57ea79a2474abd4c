"""Byte-level BPE tokenizer behind `clip.tokenize` (query-index.py:107).

Restates the published algorithm of openai/CLIP's `simple_tokenizer.py` [UPSTREAM, not
vendored in the reference]: lower-cased, whitespace-collapsed text is split by the CLIP
regex, every piece is mapped byte -> printable unicode, merged greedily by merge rank
(last symbol carries `</w>`), and looked up in a vocabulary of 256 byte symbols, the same
256 with `</w>`, the 48,894 merges and the two specials (49406 start, 49407 end).

The merges file (`bpe_simple_vocab_16e6.txt.gz`) is not available offline: point $CLIP_BPE at
it (or drop it next to this module / in ~/.cache/clip).  PARITY UNPINNED: exercised only with
a synthetic merges table (tests/test_bpe.py).  `ftfy` is not installed here either; text
repair is limited to html.unescape (applied twice, as upstream does) -- mojibake repair is
skipped, which only matters for mis-encoded input.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Optional, Tuple

import regex as re

VOCAB_SIZE = 49408
N_MERGES = 49152 - 256 - 2


def bytes_to_unicode() -> Dict[int, str]:
    """Reversible byte -> unicode map that avoids whitespace/control characters."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    chars = keep[:]
    extra = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return {b: chr(c) for b, c in zip(keep, chars)}


def _pairs(word: Tuple[str, ...]):
    return {(a, b) for a, b in zip(word, word[1:])}


def clean(text: str) -> str:
    text = html.unescape(html.unescape(text)).strip()
    return re.sub(r"\s+", " ", text).strip()


class Tokenizer:
    PATTERN = re.compile(
        r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
        re.IGNORECASE)

    def __init__(self, merges: List[Tuple[str, str]]):
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        vocab.extend("".join(m) for m in merges)
        vocab.extend(["<|startoftext|>", "<|endoftext|>"])
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        # specials keep their canonical ids even with a truncated (test) merges table
        self.encoder["<|startoftext|>"] = VOCAB_SIZE - 2
        self.encoder["<|endoftext|>"] = VOCAB_SIZE - 1
        self.decoder = {i: t for t, i in self.encoder.items()}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            merged: List[str] = []
            i = 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
            if len(word) == 1:
                break
            pairs = _pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.PATTERN.findall(clean(text).lower()):
            sym = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(sym).split(" "))
        return ids

    def decode(self, ids) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        data = bytearray(self.byte_decoder[c] for c in text)
        return data.decode("utf-8", errors="replace").replace("</w>", " ")


def load_merges(path: str) -> List[Tuple[str, str]]:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt", encoding="utf-8") as fh:
        lines = fh.read().split("\n")
    lines = lines[1:N_MERGES + 1]                   # first line is a header
    return [tuple(l.split()) for l in lines if l.strip()]


def find_merges_file() -> Optional[str]:
    cands = [os.environ.get("CLIP_BPE"),
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_simple_vocab_16e6.txt.gz"),
             os.path.expanduser("~/.cache/clip/bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    return None


@lru_cache(maxsize=1)
def default_tokenizer() -> Tokenizer:
    path = find_merges_file()
    if path is None:
        if os.environ.get("CLIPB200_BYTE_LEVEL_TOKENS"):
            import sys
            print("clipb200: no BPE merges file -- byte-level tokens only (CLIPB200_BYTE_LEVEL_TOKENS set); "
                  "results are NOT comparable with real CLIP text embeddings", file=sys.stderr)
            return Tokenizer([])
        raise FileNotFoundError(
            "clip.tokenize needs CLIP's BPE merges file bpe_simple_vocab_16e6.txt.gz: set CLIP_BPE=/path/to/it "
            "(it is not available offline); set CLIPB200_BYTE_LEVEL_TOKENS=1 to run with byte-level tokens for "
            "plumbing tests with synthetic weights")
    return Tokenizer(load_merges(path))
