"""CPU: clip.tokenize.  The real merges file is unavailable offline, so the
algorithm is checked on a synthetic merges table against an independent implementation of
the same published algorithm (transformers.CLIPTokenizer) -- parity unpinned w.r.t. openai/CLIP."""
import collections
import json
import os

import pytest

from clipb200 import bpe

CORPUS = ("a photo of a cat sitting on the sofa . a photo of a dog running on the beach . "
          "the quick brown fox jumps over the lazy dog . two cats and three dogs playing in the garden . "
          "sunset over the mountains with snow , photographed in winter . it's a bird's nest") * 3


def train_merges(text, n):
    bu = bpe.bytes_to_unicode()
    words = collections.Counter()
    for piece in bpe.Tokenizer.PATTERN.findall(text.lower()):
        sym = [bu[b] for b in piece.encode("utf-8")]
        sym[-1] += "</w>"
        words[tuple(sym)] += 1
    merges = []
    for _ in range(n):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1]); i += 2
                else:
                    out.append(w[i]); i += 1
            new[tuple(out)] += c
        words = new
    return merges


@pytest.fixture(scope="module")
def tok():
    return bpe.Tokenizer(train_merges(CORPUS, 150))


def test_vocab_layout(tok):
    assert tok.encoder["<|startoftext|>"] == 49406 and tok.encoder["<|endoftext|>"] == 49407
    assert tok.encoder["!"] == 0 and tok.encoder["!</w>"] == 256
    assert len(bpe.bytes_to_unicode()) == 256


def test_round_trip_and_merging(tok):
    text = "A photo of a cat, sitting on the   sofa."
    ids = tok.encode(text)
    assert tok.decode(ids).strip() == "a photo of a cat , sitting on the sofa ."
    assert len(ids) < len(text.replace(" ", ""))          # merges actually fire
    assert tok.encode("cat") == tok.encode("  CAT ")       # lower-casing + whitespace cleaning
    assert tok.encode("&amp;amp;") == tok.encode("&")      # html.unescape twice


def test_matches_transformers_clip_tokenizer(tok, tmp_path):
    transformers = pytest.importorskip("transformers")
    merges = [m for m, _ in sorted(tok.ranks.items(), key=lambda kv: kv[1])]
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n", encoding="utf-8")
    (tmp_path / "vocab.json").write_text(json.dumps(tok.encoder), encoding="utf-8")
    try:
        from transformers.models.clip.tokenization_clip import CLIPTokenizer
        hf = CLIPTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    except Exception as e:  # pragma: no cover
        pytest.skip(f"CLIPTokenizer unavailable: {e}")
    for text in ["a photo of a cat", "two dogs playing in the garden", "sunset over the mountains with snow",
                 "the quick brown fox", "winter bird nest on the beach"]:
        got = tok.encode(text)
        ref = hf(text)["input_ids"]
        assert ref[0] == 49406 and ref[-1] == 49407
        assert got == ref[1:-1], text


def test_clip_tokenize_shape_and_errors(tok, monkeypatch):
    import torch
    from clipb200 import clip
    monkeypatch.setattr(bpe, "default_tokenizer", lambda: tok)
    t = clip.tokenize(["a photo of a cat", "dog"])
    assert t.shape == (2, 77) and t.dtype == torch.int32
    assert t[0, 0] == 49406 and int(t[1].argmax()) == 2 and t[1, 2] == 49407 and (t[1, 3:] == 0).all()
    with pytest.raises(RuntimeError):
        clip.tokenize(["cat " * 100])
    assert clip.tokenize(["cat " * 100], truncate=True)[0, 76] == 49407


def test_missing_merges_file_is_a_loud_error(monkeypatch, tmp_path):
    monkeypatch.delenv("CLIP_BPE", raising=False)
    monkeypatch.delenv("CLIPB200_BYTE_LEVEL_TOKENS", raising=False)
    monkeypatch.setenv("HOME", str(tmp_path))
    bpe.default_tokenizer.cache_clear()
    with pytest.raises(FileNotFoundError):
        bpe.default_tokenizer()
    bpe.default_tokenizer.cache_clear()
