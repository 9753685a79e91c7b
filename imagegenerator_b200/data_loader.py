"""Caption/image loader -- the local-disk counterpart of the reference's ``data_loader.py:16-108``.

Same items, same names, same call: ``get_loader(bucket_name, root, ann_file, transform, batch_size, shuffle)`` returns a
``DataLoader`` whose batches are ``(tokenized_texts, imgs)`` -- a dict-like of ``[B,128]`` int64 tensors from the tokenizer
(padding to ``max_length=128``, truncation; ``data_loader.py:69-75``) and a ``[B,3,H,W]`` float image batch -- which is what
``train_1`` / ``train_2`` iterate over (``stage_1_train_fn.py:93``).  What differs:

  * files are read from a local directory (COCO layout: ``root/<file_name>``, ``ann_file`` = captions JSON) instead of a
    GCS bucket; ``bucket_name`` is accepted and ignored, exactly like in the train functions;
  * the annotation join (``data_loader.py:46-62``: captions inner-joined with image file names on ``image_id``, caption
    order kept) is a dictionary look-up instead of a pandas merge;
  * the sampler shards over ``torch.distributed`` ranks instead of XLA ordinals; batches come out in pinned memory so
    the train functions' non-blocking host->device copies do not synchronise the stream;
  * the tokenizer is an argument (there is no network here: pass a tokenizer object or a local directory holding the
    SpanBERT vocabulary); transformers >= 5 dropped ``batch_encode_plus``, ``Collate`` calls whichever the tokenizer has.
"""
import json
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

TOKENIZER_NAME = "SpanBERT/spanbert-base-cased"     # data_loader.py:27
MAX_LENGTH = 128                                    # data_loader.py:73


def load_tokenizer(path_or_tokenizer=None):
    """A tokenizer object passes through; a string is a local directory (or a cached hub name) for ``AutoTokenizer``."""
    if path_or_tokenizer is not None and not isinstance(path_or_tokenizer, (str, os.PathLike)):
        return path_or_tokenizer
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(path_or_tokenizer or TOKENIZER_NAME, local_files_only=True)


def caption_table(anns):
    """``[(caption, file_name)]`` in annotation order for every caption whose image is listed (data_loader.py:46-62)."""
    names = {}
    for im in anns["images"]:
        names.setdefault(im["id"], []).append(im["file_name"])
    rows = []
    for a in anns["annotations"]:
        for f in names.get(a["image_id"], ()):       # an inner join repeats a caption for a duplicated image id
            rows.append((a["caption"], f))
    return rows


class TexttoImgCOCO(Dataset):
    def __init__(self, bucket_name, root, ann_file, transform=None, tokenizer=None):
        self.bucket_name = bucket_name
        self.img_dir = root
        with open(ann_file) as f:
            rows = caption_table(json.load(f))
        self.texts = [r[0] for r in rows]
        self.imgs = [r[1] for r in rows]
        self.transform = transform
        self.tokenizer = load_tokenizer(tokenizer)

    def __len__(self):
        return len(self.texts)

    def __getitem__(self, index):
        from PIL import Image
        with Image.open(os.path.join(self.img_dir, self.imgs[index])) as im:
            img = im.convert("RGB")
        if self.transform is not None:
            img = self.transform(img)
        return self.texts[index], img


class Collate:
    def __init__(self, tokenizer, max_length=MAX_LENGTH):
        self.tokenizer = tokenizer
        self.max_length = max_length

    def __call__(self, batch):
        texts = [item[0] for item in batch]
        encode = getattr(self.tokenizer, "batch_encode_plus", None) or self.tokenizer
        tokenized_texts = encode(texts, padding="max_length", truncation=True, max_length=self.max_length,
                                 return_tensors="pt")
        imgs = torch.stack([item[1] for item in batch], dim=0)
        return tokenized_texts, imgs


def get_loader(bucket_name, root, ann_file, transform, batch_size=64, shuffle=True, tokenizer=None, num_workers=8,
               prefetch_factor=16, seed=0):
    dataset = TexttoImgCOCO(bucket_name=bucket_name, root=root, ann_file=ann_file, transform=transform,
                            tokenizer=tokenizer)
    on = dist.is_available() and dist.is_initialized()
    sampler = DistributedSampler(dataset, num_replicas=dist.get_world_size() if on else 1,
                                 rank=dist.get_rank() if on else 0, shuffle=shuffle, seed=seed)
    return DataLoader(dataset=dataset, batch_size=batch_size, collate_fn=Collate(dataset.tokenizer), sampler=sampler,
                      drop_last=True, num_workers=num_workers, persistent_workers=False,
                      prefetch_factor=prefetch_factor if num_workers > 0 else None,
                      pin_memory=torch.cuda.is_available())
