"""GPU, BASELINE configs[0] in miniature: synthetic JPEG/PNG folder -> batched build-index
pipeline (PIL decode/resize on the CPU, ToTensor/Normalize + ViT-B/32 + L2-norm on the GPU,
LMDB-format store, flat index) -> text query top-20, against the CPU oracle pipeline."""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make_folder(root, n=150):
    from PIL import Image
    rng = np.random.default_rng(1234)
    os.makedirs(root, exist_ok=True)
    sizes = [(224, 224), (224, 224), (320, 240), (200, 300), (640, 480)]
    for i in range(n):
        w, h = sizes[i % len(sizes)]
        base = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)
        im = Image.fromarray(base).resize((w, h), Image.BICUBIC)
        arr = np.clip(np.asarray(im, dtype=np.float32) + rng.normal(0, 8, (h, w, 3)), 0, 255).astype(np.uint8)
        ext = ".png" if i % 10 == 0 else (".JPG" if i % 7 == 0 else ".jpg")
        Image.fromarray(arr).save(os.path.join(root, f"img_{i:06d}{ext}"), quality=90)
    open(os.path.join(root, "broken.jpg"), "wb").write(b"not a jpeg at all")
    open(os.path.join(root, "notes.txt"), "w").write("ignored")


def test_build_index_and_query_against_oracle(tmp_path, monkeypatch):
    import torch
    from PIL import Image
    from clipb200 import clip, faiss, indexer, lmdb, weights
    from oracle import clip_ref, flatip_ref

    folder = str(tmp_path / "photos") + "/"
    _make_folder(folder)
    monkeypatch.chdir(tmp_path)
    sd = weights.synthetic_state_dict(0)
    model = clip.CLIPB200(sd, device=0, max_image_batch=64, max_text_batch=4)
    env = lmdb.open("vectors.lmdb", map_size=1 << 30, max_dbs=4)
    log = io.StringIO()
    ok, bad = indexer.embed_folders([folder], env, model, batch=64, out=log)
    assert (ok, bad) == (150, 1)
    text = log.getvalue()
    assert text.startswith(f"CLIPing {folder}...") and text.count(".") >= 150 and text.count("#") == 1

    # resume: nothing new the second time (build-index.py:42-44)
    log2 = io.StringIO()
    assert indexer.embed_folders([folder], env, model, batch=64, out=log2) == (0, 1)

    index = indexer.build_index(env, faiss, index_path="images.index", out=io.StringIO())
    assert index.ntotal == 150 and os.path.exists("images.index")

    # stored vectors vs the fp32 oracle run on the reference's own CPU transform
    fn_db, idx_db = env.open_db(b"fn_db"), env.open_db(b"idx_db")
    transform = clip._transform(224)
    with env.begin(db=fn_db) as txn:
        keys = [k for k, _ in txn.cursor()]
        assert keys == sorted(keys) and len(keys) == 150
        sample = keys[::17]
        x = torch.stack([transform(Image.open(k.decode())) for k in sample])
        ref = clip_ref.l2_normalize_rows(clip_ref.encode_image(sd, x)).numpy()
        got = np.stack([np.frombuffer(txn.get(k), dtype=np.float32) for k in sample])
        assert all(len(txn.get(k)) == 2048 for k in sample)
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() >= 0.999, cos.min()
    with env.begin(db=idx_db) as txn:
        assert txn.get(b"0") == keys[0] and txn.get(b"149") == keys[149]

    # text query, top-20 (c 20 -> search(k=21), rank 0 skipped: query-index.py:111-116)
    searcher = indexer.Searcher(env, index, model)
    tokens = clip_ref.synthetic_tokens(1, seed=4)
    feats = searcher.features_for_tokens(tokens)
    rows = searcher.results(feats, k=20, offset=0)
    assert len(rows) == 20
    with env.begin(db=fn_db) as txn:
        all_vecs = np.stack([np.frombuffer(txn.get(k), dtype=np.float32) for k in keys])
    Dref, Iref = flatip_ref.search(feats, all_vecs, 21)
    Dgot = np.array([[r[0] for r in rows]], dtype=np.float32)
    Igot = np.array([[r[1] for r in rows]])
    okk, _, msg = flatip_ref.ids_match_with_tolerance(Dref[:, 1:], Iref[:, 1:], Dgot, Igot)
    assert okk, msg
    assert rows[0][2] == keys[rows[0][1]].decode()
    assert indexer.Searcher.format_row(rows[0]).split(" ")[0] == f"{rows[0][0]:.4f}"
    # the text feature itself agrees with the oracle text tower
    tref = clip_ref.l2_normalize_rows(clip_ref.encode_text(sd, tokens)).numpy()
    assert float((feats * tref).sum()) >= 0.999

    # `i ID`: the stored vector's best match is itself (rank 0, skipped by the REPL)
    f7 = searcher.features_for_id(7)
    D, I = index.search(f7, 3)
    assert I[0][0] == 7 and abs(D[0][0] - 1.0) < 1e-3
    env.close()

    # reopen from disk like query-index.py:25-29
    env2 = lmdb.open("vectors.lmdb", map_size=1 << 30, max_dbs=4)
    index2 = faiss.read_index("images.index")
    s2 = indexer.Searcher(env2, index2, model)
    assert s2.results(feats, k=20, offset=0) == rows
    env2.close()


def test_gpu_resize_and_nvjpeg_paths(tmp_path, monkeypatch):
    """resize="gpu" stores bit-identical vectors to the Pillow path (the CUDA resize is
    pixel-exact); decode="nvjpeg" agrees to cosine >= 0.999 (nvjpeg vs libjpeg-turbo: +-1 LSB)."""
    import torch
    from clipb200 import clip, indexer, lmdb, weights
    folder = str(tmp_path / "photos") + "/"
    _make_folder(folder, n=40)
    sd = weights.synthetic_state_dict(0)
    model = clip.CLIPB200(sd, device=0, max_image_batch=16, max_text_batch=1)
    stores = {}
    for mode, kw in (("cpu", {}), ("gpu", {"resize": "gpu"}), ("nvjpeg", {"decode": "nvjpeg"})):
        env = lmdb.open(str(tmp_path / f"{mode}.lmdb"), map_size=1 << 30, max_dbs=4)
        try:
            ok, bad = indexer.embed_folders([folder], env, model, batch=16, out=io.StringIO(), **kw)
        except (ImportError, RuntimeError) as e:
            if mode == "nvjpeg":
                pytest.skip(f"torchvision nvjpeg decode unavailable: {e}")
            raise
        assert (ok, bad) == (40, 1), (mode, ok, bad)
        with env.begin(db=env.open_db(b"fn_db")) as txn:
            stores[mode] = {k: np.frombuffer(v, dtype=np.float32) for k, v in txn.cursor()}
        env.close()
    assert stores["cpu"].keys() == stores["gpu"].keys() == stores["nvjpeg"].keys()
    for k in stores["cpu"]:
        assert np.array_equal(stores["cpu"][k], stores["gpu"][k]), k
        assert float((stores["cpu"][k] * stores["nvjpeg"][k]).sum()) >= 0.999, k


def test_query_session_commands_and_quirks(tmp_path, monkeypatch):
    """The query loop (query-index.py:40-119) as a state machine: same lines out for the same lines
    in, with encode_text and index.search on the GPU."""
    import re
    from clipb200 import clip, faiss, indexer, lmdb, repl, weights
    from oracle import clip_ref

    folder = str(tmp_path / "photos") + "/"
    _make_folder(folder, n=30)
    monkeypatch.chdir(tmp_path)
    model = clip.CLIPB200(weights.synthetic_state_dict(0), device=0, max_image_batch=32, max_text_batch=4)
    env = lmdb.open("vectors.lmdb", map_size=1 << 30, max_dbs=4)
    indexer.embed_folders([folder], env, model, batch=32, out=io.StringIO())
    indexer.build_index(env, faiss, index_path="images.index", out=io.StringIO())
    index = faiss.read_index("images.index")           # as query-index.py:29-30
    index.nprobe = 32

    class SyntheticTextSearcher(indexer.Searcher):     # no BPE vocabulary offline: text -> seeded token row
        def features_for_text(self, text):
            return self.features_for_tokens(clip_ref.synthetic_tokens(1, seed=len(text)))

    lines = []
    s = repl.QuerySession(SyntheticTextSearcher(env, index, model), index, out=lines.append)

    def feed(text):
        lines.clear()
        alive = s.handle(text)
        return alive, list(lines)

    assert repl.PROMPT == "[h,q,i,r,a,c,p] >>> "
    assert feed("h")[1][0].startswith("Enter a search query")
    assert feed("")[1] == []                                   # "more" before any text query: nothing
    assert feed("p 100")[1] == ["Set to probe 100 subsets."] and index.nprobe == 100
    assert feed("p 101")[1] == ["Invalid probe value."] and feed("p 0")[1] == ["Invalid probe value."]
    assert feed("c 5")[1] == ["Showing 5 results."]
    assert feed("c 0")[1] == ["Reset number of results to 50."] and s.k == 50
    assert feed("r 1280x720")[1] == ["Set maximum resolution to 1280x720."] and s.max_res == (1280, 720)
    assert feed("r nonsense")[1] == ["Unset maximum resolution."] and s.max_res is None
    assert feed("a")[1] == ["Aligning window position."] and feed("a")[1] == ["Not aligning window position."]
    feed("c 5")

    alive, out = feed("a photo of a dog")
    assert alive and re.fullmatch(r"Search time: \d+\.\d{4}s", out[0]) and len(out) == 6
    first = [re.fullmatch(r"(-?\d+\.\d{4}) (\d+) (.+)", l).groups() for l in out[1:]]
    D, I = index.search(s.features, 11)
    assert [int(f[1]) for f in first] == list(I[0][1:6])       # rank 0 skipped, k + offset + 1 requested
    assert all(f[2].startswith(folder) for f in first) and s.last_j == 5
    _, out = feed("")                                          # more: ranks 6..10
    assert [int(l.split(" ")[1]) for l in out[1:]] == list(I[0][6:11]) and s.last_j == 10

    _, out = feed("i 7")                                       # similar to a stored image
    assert out[0].startswith("Similar to " + folder) and out[0].endswith(":")
    assert len(out) == 7 and int(out[2].split(" ")[1]) != 7    # rank 0 (the image itself) is skipped
    D7, I7 = index.search(s.features, 6)
    assert I7[0][0] == 7 and [int(l.split(" ")[1]) for l in out[2:]] == list(I7[0][1:6])
    assert feed("i 9999")[1] == ["Not found."]

    feed("c 100")                                              # more than the index holds: stops at the last row
    _, out = feed("another query")
    assert len(out) == 1 + 29
    shown = []
    s.show = lambda path, sess: (shown.append(path), len(shown) < 3)[1]   # viewer says "q" on the 3rd image
    _, out = feed("third")
    assert len(shown) == 3 and len(out) == 1 + 3 and s.last_j == 3
    assert feed("q")[0] is False
    env.close()


def test_two_gpus_embed_one_folder(tmp_path, monkeypatch):
    """SURVEY.md 8e, index time: the files of a folder are split over the GPUs (one replica each, no
    collective), one thread commits.  The stored vectors are those of the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from clipb200 import clip, indexer, lmdb, weights
    folder = str(tmp_path / "photos") + "/"
    _make_folder(folder, n=300)
    monkeypatch.chdir(tmp_path)
    sd = weights.synthetic_state_dict(0)
    stores = []
    for name, models in (("one", None), ("two", [0, 1])):
        if models is None:
            m = clip.CLIPB200(sd, device=0, max_image_batch=64, max_text_batch=1)
        else:
            m = [clip.CLIPB200(sd, device=d, max_image_batch=64, max_text_batch=1) for d in models]
        env = lmdb.open(f"{name}.lmdb", map_size=1 << 30, max_dbs=4)
        log = io.StringIO()
        ok, bad = indexer.embed_folders([folder], env, m, batch=64, out=log, decode="nvjpeg")
        marks = log.getvalue().split("\n", 1)[1]           # progress characters after the "CLIPing ..." line
        assert (ok, bad) == (300, 1) and marks.count(".") == 300 and marks.count("#") == 1
        with env.begin(db=env.open_db(b"fn_db")) as txn:
            stores.append({k: bytes(v) for k, v in txn.cursor()})
        env.close()
    assert stores[0].keys() == stores[1].keys() and len(stores[0]) == 300
    a = np.stack([np.frombuffer(stores[0][k], dtype=np.float32) for k in sorted(stores[0])])
    b = np.stack([np.frombuffer(stores[1][k], dtype=np.float32) for k in sorted(stores[0])])
    # same decoder, same kernels, but a row's batch neighbours differ (tile shapes) -> fp16-level differences
    assert (a * b).sum(1).min() >= 0.99999


def test_process_pool_decode_stores_the_same_bytes(tmp_path, monkeypatch):
    """The Pillow path with decode worker processes (large folders) stores byte-identical vectors to the
    in-process thread pool: same pixels in, same batches, same kernels."""
    from clipb200 import clip, indexer, lmdb, weights
    folder = str(tmp_path / "photos") + "/"
    _make_folder(folder, n=200)
    monkeypatch.chdir(tmp_path)
    model = clip.CLIPB200(weights.synthetic_state_dict(0), device=0, max_image_batch=64, max_text_batch=1)
    stores = []
    for name, nproc in (("threads", "0"), ("procs", "3")):
        monkeypatch.setenv("CLIPB200_PIL_PROCESSES", nproc)
        env = lmdb.open(f"{name}.lmdb", map_size=1 << 30, max_dbs=4)
        log = io.StringIO()
        assert indexer.embed_folders([folder], env, model, batch=64, out=log) == (200, 1)
        marks = log.getvalue().split("\n", 1)[1]
        assert marks.count(".") == 200 and marks.count("#") == 1
        with env.begin(db=env.open_db(b"fn_db")) as txn:
            stores.append({bytes(k): bytes(v) for k, v in txn.cursor()})
        env.close()
    assert stores[0] == stores[1] and len(stores[0]) == 200
