"""Generates the committed golden fixtures (run from the repo root, CPU only).

The reference ships no golden vectors and its dependencies
(faiss, openai/CLIP) are not installable offline, so the fixtures pin the oracle
restatements against each other / against independent implementations:

  flatip_golden.npz : seeded 4096x512 database (+ duplicate rows for ties), 6 queries;
                      expected (D, I) for k in {1, 21, 100} from oracle/flatip_ref.py, after
                      asserting that the independently written C heap restatement
                      (oracle/flatip_ref.c) returns the same ids.
  clip_golden.npz   : embeddings of 4 seeded synthetic images and 4 synthetic token rows
                      under the seeded synthetic ViT-B/32 weights (clipb200.weights, seed 0),
                      computed by transformers.CLIPModel -- an independent implementation of
                      the same architecture -- after asserting oracle/clip_ref.py agrees with
                      it to < 2e-5.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))

from oracle import flatip_ref as F  # noqa: E402
from clipb200 import synth  # noqa: E402


def flatip_inputs():
    xb = synth.unit_rows(4096, seed=11, clip_like=True)
    xb[100:108] = xb[7]          # exact duplicates -> exact score ties
    xb[4000] = xb[3]
    xq = synth.unit_rows(6, seed=12, clip_like=True)
    xq[5] = xb[7]                # a query identical to a stored row (image-similarity query)
    return xb.astype(np.float16), xq


def clip_inputs():
    import torch
    g = torch.Generator().manual_seed(1234)
    images = torch.randint(0, 256, (4, 224, 224, 3), generator=g, dtype=torch.uint8)
    # low-frequency structure + noise, like a photograph more than like static
    base = torch.nn.functional.interpolate(
        torch.rand((4, 3, 8, 8), generator=g), size=(224, 224), mode="bicubic", align_corners=False)
    images = (base.permute(0, 2, 3, 1) * 255 + (images.float() - 128) * 0.06).clamp(0, 255).to(torch.uint8)
    from oracle import clip_ref
    tokens = clip_ref.synthetic_tokens(4, seed=5)
    return images, tokens


def make_clip_golden():
    import torch
    from transformers import CLIPConfig, CLIPModel
    from clipb200 import weights
    from oracle import clip_ref
    torch.manual_seed(0)
    sd = weights.synthetic_state_dict(0)
    images, tokens = clip_inputs()
    x = clip_ref.preprocess_u8(images)
    img = clip_ref.encode_image(sd, x)
    txt = clip_ref.encode_text(sd, tokens)
    hf = CLIPModel(CLIPConfig()).eval()
    missing, unexpected = hf.load_state_dict(clip_ref.to_hf_state_dict(sd), strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    with torch.no_grad():
        hf_img = hf.get_image_features(pixel_values=x)
        hf_txt = hf.get_text_features(input_ids=tokens.long())
    if not torch.is_tensor(hf_img):
        hf_img, hf_txt = hf_img.pooler_output, hf_txt.pooler_output
    di = (img - hf_img).abs().max().item()
    dt = (txt - hf_txt).abs().max().item()
    print(f"oracle vs transformers.CLIPModel: image max|d|={di:.3g} text max|d|={dt:.3g}; "
          f"|img|={img.norm(dim=1).tolist()}")
    assert di < 2e-5 and dt < 2e-5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clip_golden.npz"),
                        image_features=hf_img.numpy(), text_features=hf_txt.numpy())
    print("wrote clip_golden.npz")


def main():
    if "--clip-only" not in sys.argv:
        make_flatip_golden()
    if "--flatip-only" not in sys.argv:
        make_clip_golden()


def make_flatip_golden():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_flatip.so"))
    xb, xq = flatip_inputs()
    out = {}
    for k in (1, 21, 100):
        D, I = F.search(xq, xb, k)
        D2 = np.empty_like(D)
        I2 = np.empty_like(I)
        rc = lib.oracle_flatip_search(C.c_void_p(xb.ctypes.data), 1, C.c_int64(xb.shape[0]), 512,
                                      C.c_void_p(xq.ctypes.data), C.c_int64(xq.shape[0]), C.c_int64(k),
                                      C.c_void_p(D2.ctypes.data), C.c_void_p(I2.ctypes.data), 1)
        assert rc == 0
        ok, exempt, msg = F.ids_match_with_tolerance(D, I, D2, I2)
        assert ok, msg
        assert np.abs(D - D2).max() < 1e-6
        out[f"D{k}"] = D
        out[f"I{k}"] = I
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "flatip_golden.npz"), **out)
    print("wrote flatip_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
