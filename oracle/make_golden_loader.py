"""Golden fixture for the caption/image loader -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference ``data_loader.get_loader`` (``/root/reference/data_loader.py:84-108``: TexttoImgCOCO +
Collate + DistributedSampler + DataLoader with 8 workers) on a small COCO-shaped directory served through the fake GCS
bucket of ``oracle/ref_harness.py``, and records what it yields into ``tests/golden/loader.pt``.  The directory is the
seeded one ``tests/_util.make_coco_dir`` writes, so the test can rebuild it anywhere and hold
``imagegenerator_b200.data_loader`` to the recorded batches.

Two accommodations to this container, neither touching the reference's code: the SpanBERT tokenizer cannot be downloaded,
so ``AutoTokenizer.from_pretrained`` is pointed at the small word-piece tokenizer of the test directory; and the installed
transformers 5.x no longer has ``batch_encode_plus`` (the reference was written against 4.x, ``data_loader.py:69``), so the
tokenizer is wrapped with an object whose ``batch_encode_plus`` is the tokenizer's ``__call__`` (its 5.x equivalent).

    python -m oracle.make_golden_loader
"""
import os
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as R  # noqa: E402

N_IMAGES, CAPTIONS_PER_IMAGE, BATCH = 5, 2, 4


class _Tok4:
    """transformers-4 surface of a transformers-5 tokenizer."""

    def __init__(self, tok):
        self._tok = tok

    def batch_encode_plus(self, texts, **kw):
        return self._tok(texts, **kw)

    def __getattr__(self, name):
        return getattr(self._tok, name)


def main():
    import torchvision.transforms as transforms
    from _util import make_coco_dir
    torch.set_num_threads(1)
    with tempfile.TemporaryDirectory() as tmp:
        root, ann, tok, rows = make_coco_dir(tmp, n_images=N_IMAGES, captions_per_image=CAPTIONS_PER_IMAGE)
        for f in os.listdir(root):                                   # "upload" the directory into the fake bucket
            with open(os.path.join(root, f), "rb") as fh:
                R._STORE[os.path.join("dataset/train2017", f)] = fh.read()
        with open(ann, "rb") as fh:
            R._STORE["dataset/annotations/captions_train2017.json"] = fh.read()
    ref = R.load("data_loader")
    ref.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda name: _Tok4(tok))
    transform = transforms.Compose([transforms.ToTensor(), transforms.Resize((64, 64)),        # train.py:40-46
                                    transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
    out = {"n_images": N_IMAGES, "captions_per_image": CAPTIONS_PER_IMAGE, "batch": BATCH, "rows": rows}
    for shuffle in (False, True):
        loader = ref.get_loader(bucket_name="data-and-checkpoints-bucket", root="dataset/train2017",
                                ann_file="dataset/annotations/captions_train2017.json", transform=transform,
                                batch_size=BATCH, shuffle=shuffle)
        batches = []
        for tokenized, imgs in loader:
            batches.append({"tokenized": {k: v.clone() for k, v in tokenized.items()}, "imgs": imgs.clone()})
        key = "shuffled" if shuffle else "ordered"
        out[key + "_len"] = len(loader)
        out[key] = [{"tokenized": b["tokenized"], "imgs": b["imgs"] if (i == 0 and not shuffle) else None,
                     "img_sum": b["imgs"].double().sum().item(), "img_abs": b["imgs"].double().abs().sum().item()}
                    for i, b in enumerate(batches)]
        out[key + "_dataset_len"] = len(loader.dataset)
    dst = os.path.join(ROOT, "tests", "golden", "loader.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes;", out["ordered_len"], "batches of", BATCH)


if __name__ == "__main__":
    main()
