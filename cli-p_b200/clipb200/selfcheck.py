"""Self-check for users who have the real OpenAI checkpoint (and, optionally, the BPE vocabulary):

    python -m clipb200.selfcheck --weights ~/.cache/clip/ViT-B-32.pt
                                 [--vocab bpe_simple_vocab_16e6.txt.gz] [--images DIR] [--device 0]

Offline, clipb200's towers can only be validated with seeded synthetic weights.  Real CLIP weights
carry a few huge-magnitude residual channels, which is exactly where an fp16 residual stream and the
LayerNorm fold (include/clipb200.h, cb_clip_ln_fold_status) could lose accuracy.  This tool measures,
on the user's own weights:

  * cosine(libclipb200, fp32 torch) per image / text row -- the north star's bar is >= 0.999;
  * whether cb_clip_finalize kept the LayerNorm fold, and folded vs unfolded agreement;
  * the tcgen05 attention kernel against the mma.sync one it replaces (knob attn_tc);
  * the largest residual-stream magnitude the fp32 model sees (fp16 overflows at 65504);
  * with --vocab: that clip.tokenize round-trips a few strings through the real merges table.

The fp32 reference below is plain torch (on the same GPU, TF32 off), written against openai/CLIP's
published forward; it is a checker and is never used to produce results.  Exit code 0 = all bars met.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

_MEAN = (0.48145466, 0.4578275, 0.40821073)
_STD = (0.26862954, 0.26130258, 0.27577711)


def _block(x, sd, pre, heads, mask, peak):
    B, L, W = x.shape
    hd = W // heads
    h = F.layer_norm(x, (W,), sd[f"{pre}.ln_1.weight"], sd[f"{pre}.ln_1.bias"], 1e-5)
    q, k, v = F.linear(h, sd[f"{pre}.attn.in_proj_weight"], sd[f"{pre}.attn.in_proj_bias"]).split(W, dim=-1)
    q, k, v = (t.view(B, L, heads, hd).transpose(1, 2) for t in (q, k, v))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    a = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, W)
    x = x + F.linear(a, sd[f"{pre}.attn.out_proj.weight"], sd[f"{pre}.attn.out_proj.bias"])
    h = F.layer_norm(x, (W,), sd[f"{pre}.ln_2.weight"], sd[f"{pre}.ln_2.bias"], 1e-5)
    h = F.linear(h, sd[f"{pre}.mlp.c_fc.weight"], sd[f"{pre}.mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)
    x = x + F.linear(h, sd[f"{pre}.mlp.c_proj.weight"], sd[f"{pre}.mlp.c_proj.bias"])
    peak[0] = max(peak[0], float(x.abs().max()))
    return x


def ref_encode_image(sd, images_u8, peak):
    x = images_u8.permute(0, 3, 1, 2).float() / 255.0
    dev = x.device
    x = (x - torch.tensor(_MEAN, device=dev).view(1, 3, 1, 1)) / torch.tensor(_STD, device=dev).view(1, 3, 1, 1)
    x = F.conv2d(x, sd["visual.conv1.weight"], stride=32)
    B, W = x.shape[0], x.shape[1]
    x = x.reshape(B, W, -1).permute(0, 2, 1)
    x = torch.cat([sd["visual.class_embedding"].expand(B, 1, W), x], dim=1) + sd["visual.positional_embedding"]
    x = F.layer_norm(x, (W,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], 1e-5)
    peak[0] = max(peak[0], float(x.abs().max()))
    for i in range(12):
        x = _block(x, sd, f"visual.transformer.resblocks.{i}", 12, None, peak)
    x = F.layer_norm(x[:, 0, :], (W,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], 1e-5)
    return x @ sd["visual.proj"]


def ref_encode_text(sd, ids, peak):
    B, L = ids.shape
    x = sd["token_embedding.weight"][ids.long()] + sd["positional_embedding"]
    mask = torch.full((L, L), float("-inf"), device=x.device).triu_(1)
    peak[0] = max(peak[0], float(x.abs().max()))
    for i in range(12):
        x = _block(x, sd, f"transformer.resblocks.{i}", 8, mask, peak)
    x = F.layer_norm(x, (512,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(B, device=x.device), ids.argmax(dim=-1)] @ sd["text_projection"]


def _images(path, n, device):
    g = torch.Generator().manual_seed(7)
    imgs = [torch.randint(0, 256, (224, 224, 3), generator=g, dtype=torch.uint8) for _ in range(max(0, n - 3))]
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    imgs.append(torch.stack([(xx + yy) // 2, xx, yy], dim=-1).clamp(0, 255).to(torch.uint8))
    imgs.append(torch.full((224, 224, 3), 128, dtype=torch.uint8))
    imgs.append(((((xx // 16) + (yy // 16)) % 2) * 255).to(torch.uint8).unsqueeze(-1).expand(224, 224, 3).contiguous())
    if path:
        from .pil_transform import image_to_u8
        from PIL import Image
        for fn in sorted(os.listdir(path))[:64]:
            if fn.lower().endswith((".jpg", ".jpeg", ".png")):
                try:
                    imgs.append(torch.from_numpy(image_to_u8(Image.open(os.path.join(path, fn)), 224).copy()))
                except Exception:
                    pass
    return torch.stack(imgs).to(device)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--weights", default=os.environ.get("CLIP_WEIGHTS"),
                    help="OpenAI ViT-B-32.pt (TorchScript archive or state_dict); 'synthetic' = seeded weights")
    ap.add_argument("--vocab", default=None, help="bpe_simple_vocab_16e6.txt.gz (optional)")
    ap.add_argument("--images", default=None, help="folder of images to add to the built-in probes")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--bar", type=float, default=0.999)
    args = ap.parse_args(argv)
    if not args.weights:
        ap.error("--weights (or CLIP_WEIGHTS) is required")
    if not torch.cuda.is_available():
        print("selfcheck: no CUDA device -- clipb200 has no CPU path", file=sys.stderr)
        return 2
    from . import _native as N, clip, weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", args.device)
    sd = weights.synthetic_state_dict(0) if args.weights == "synthetic" else weights.load_state_dict(args.weights)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    ok = True

    model = clip.CLIPB200(sd, device=args.device, max_image_batch=64, max_text_batch=64)
    folded, cal = model.ln_fold_status()
    print(f"LayerNorm fold: {'kept' if folded else 'DROPPED (separate fp32 LayerNorm launches)'}; "
          f"calibration cosine folded vs unfolded = {cal:.7f}")
    with N.tuning(ln_fold=0):
        plain = clip.CLIPB200(sd, device=args.device, max_image_batch=64, max_text_batch=64)

    imgs = _images(args.images, 16, dev)
    peak = [0.0]
    with torch.no_grad():
        ref = torch.cat([ref_encode_image(sd_dev, imgs[i:i + 16], peak) for i in range(0, len(imgs), 16)])
    got = torch.cat([model.encode_image(imgs[i:i + 64]) for i in range(0, len(imgs), 64)])
    unf = torch.cat([plain.encode_image(imgs[i:i + 64]) for i in range(0, len(imgs), 64)])
    with N.tuning(attn_tc=0):                      # the mma.sync attention kernel instead of the tcgen05 one
        alt = torch.cat([model.encode_image(imgs[i:i + 64]) for i in range(0, len(imgs), 64)])
    c = F.cosine_similarity(got, ref)
    cu = F.cosine_similarity(unf, ref)
    cf = F.cosine_similarity(got, unf)
    ca = F.cosine_similarity(got, alt)
    print(f"encode_image over {len(imgs)} images: cosine vs fp32 torch  min {c.min():.6f}  mean {c.mean():.6f}   "
          f"(unfolded build: min {cu.min():.6f}); shipped vs unfolded min {cf.min():.6f}")
    print(f"  tcgen05 attention vs mma.sync attention: min cosine {ca.min():.6f}")
    print(f"  largest residual-stream magnitude in the fp32 vision tower: {peak[0]:.1f} (fp16 max 65504)")
    ok &= bool(c.min() >= args.bar) and bool(ca.min() >= args.bar) and peak[0] < 60000

    if args.vocab:
        os.environ["CLIP_BPE"] = args.vocab
        texts = ["a photo of a cat", "two dogs playing in the snow", "a diagram of a steam engine",
                 "Sunset over the harbour, long exposure", "a page of handwritten notes"]
        ids = clip.tokenize(texts).to(dev)
        print(f"tokenize: {len(texts)} strings, lengths {[int((r != 0).sum()) for r in ids]}, "
              f"sot/eot {int(ids[0, 0])}/{int(ids[0].max())}")
        ok &= int(ids[0, 0]) == 49406 and int(ids[0].max()) == 49407
    else:
        from .synth import synthetic_tokens
        ids = synthetic_tokens(16, seed=3).to(dev)
    peak = [0.0]
    with torch.no_grad():
        rt = ref_encode_text(sd_dev, ids, peak)
    gt = model.encode_text(ids)
    ut = plain.encode_text(ids)
    c = F.cosine_similarity(gt, rt)
    print(f"encode_text over {ids.shape[0]} rows: cosine vs fp32 torch  min {c.min():.6f}  mean {c.mean():.6f}   "
          f"(unfolded build: min {F.cosine_similarity(ut, rt).min():.6f})")
    print(f"  largest residual-stream magnitude in the fp32 text tower: {peak[0]:.1f}")
    ok &= bool(c.min() >= args.bar) and peak[0] < 60000
    print("selfcheck: " + ("OK -- every bar met" if ok else f"FAILED -- cosine below {args.bar} or fp16 range at risk"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
