"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def digest(t, n=64):
    """Must match oracle/make_golden.py:digest."""
    t = t.detach().to(torch.float64).reshape(-1).cpu()
    g = torch.Generator().manual_seed(t.numel() % 9973 + 17)
    idx = torch.randint(0, t.numel(), (min(n, t.numel()),), generator=g)
    return dict(numel=t.numel(), sum=t.sum().item(), norm=t.norm().item(), idx=idx, vals=t[idx].clone())


def assert_digest(t, dg, rtol, atol, what=""):
    """Compare tensor ``t`` with a stored digest: sampled elements elementwise, L2 norm relatively."""
    t = t.detach().to(torch.float64).reshape(-1).cpu()
    assert t.numel() == dg["numel"], (what, t.numel(), dg["numel"])
    got = t[dg["idx"]]
    scale = max(dg["norm"] / max(dg["numel"], 1) ** 0.5, 1e-30)   # rms of the tensor
    ok = torch.allclose(got, dg["vals"], rtol=rtol, atol=atol)
    if not ok:
        err = (got - dg["vals"]).abs().max().item()
        raise AssertionError(f"{what}: sampled elements differ, max abs err {err:.3e} (rms {scale:.3e}, rtol {rtol}, atol {atol})")
    nerr = abs(t.norm().item() - dg["norm"])
    assert nerr <= rtol * dg["norm"] + atol * dg["numel"] ** 0.5, f"{what}: norm {t.norm().item():.6e} vs {dg['norm']:.6e}"


def assert_digest_dict(d, dgs, rtol, atol, what=""):
    for k, dg in dgs.items():
        assert k in d, (what, k)
        assert_digest(d[k], dg, rtol, atol, f"{what}[{k}]")


WORDS = ["a", "cat", "dog", "bird", "on", "the", "mat", "red", "blue", "sits", "flies", "over", "house", "tree"]


def make_coco_dir(tmp_path, n_images=5, captions_per_image=2, seed=0):
    """A COCO-captions-shaped directory (images + annotations JSON, data_loader.py:46-62) and a word-piece tokenizer
    over a small cased vocabulary.  Returns ``(root, ann_file, tokenizer, [(caption, file_name)] in join order)``."""
    import json
    import random
    from PIL import Image
    from transformers import BertTokenizer
    rnd = random.Random(seed)
    g = torch.Generator().manual_seed(seed)
    root = os.path.join(str(tmp_path), "train2017")
    os.makedirs(root, exist_ok=True)
    images, annotations, rows = [], [], []
    for i in range(n_images):
        name = f"{i:012d}.png"
        h, w = rnd.choice([(48, 80), (96, 64), (70, 70)])
        arr = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).numpy()
        Image.fromarray(arr).save(os.path.join(root, name))
        images.append({"id": 100 + i, "file_name": name, "height": h, "width": w})
    for c in range(captions_per_image):          # captions interleaved over images, like the real file
        for i in range(n_images):
            cap = " ".join(rnd.choice(WORDS) for _ in range(rnd.randint(3, 8)))
            annotations.append({"id": len(annotations), "image_id": 100 + i, "caption": cap})
            rows.append((cap, images[i]["file_name"]))
    ann = os.path.join(str(tmp_path), "captions_train2017.json")
    with open(ann, "w") as f:
        json.dump({"images": images, "annotations": annotations}, f)
    vocab = os.path.join(str(tmp_path), "vocab.txt")
    with open(vocab, "w") as f:
        f.write("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS))
    return root, ann, BertTokenizer(vocab, do_lower_case=False), rows


def tiny_bert(tokenizer, layers=1, seed=0):
    """A BERT encoder with SpanBERT-base's hidden size (768 -> the projection head's input) and few layers."""
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=len(tokenizer), hidden_size=768, num_hidden_layers=layers, num_attention_heads=12,
                     intermediate_size=128, max_position_embeddings=128, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    return BertModel(cfg, add_pooling_layer=False)
