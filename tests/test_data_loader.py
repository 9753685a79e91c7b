"""The caption/image loader (reference data_loader.py:16-108) on a small COCO-shaped directory
written by the test itself, and the text side it feeds: tokenizer -> BERT encoder CLS state -> Linear(768, 512)."""
import json
import os

import pytest
import torch

from _util import make_coco_dir, tiny_bert


def test_caption_table_is_the_inner_join_in_caption_order():
    from imagegenerator_b200.data_loader import caption_table
    anns = {"images": [{"id": 7, "file_name": "b.png"}, {"id": 3, "file_name": "a.png"}],
            "annotations": [{"image_id": 3, "caption": "x"}, {"image_id": 9, "caption": "orphan"},
                            {"image_id": 7, "caption": "y"}, {"image_id": 3, "caption": "z"}]}
    assert caption_table(anns) == [("x", "a.png"), ("y", "b.png"), ("z", "a.png")]


def test_caption_table_matches_the_pandas_merge_of_the_reference():
    """data_loader.py:54-60 builds the table as ``text_df.merge(img_df, on="image_id")``; same rows, same order."""
    import random
    import pandas as pd
    from imagegenerator_b200.data_loader import caption_table
    rnd = random.Random(3)
    images = [{"id": i, "file_name": f"{i}.jpg", "height": 1} for i in rnd.sample(range(200), 120)]
    images.append({"id": images[5]["id"], "file_name": "dup.jpg", "height": 2})          # a duplicated image id
    annotations = [{"id": k, "image_id": rnd.randrange(220), "caption": f"c{k}"} for k in range(600)]
    img_df = pd.DataFrame(images)[["id", "file_name"]].rename(columns={"id": "image_id"})
    text_df = pd.DataFrame(annotations)[["image_id", "caption"]]
    merged = text_df.merge(img_df, on="image_id")
    want = list(zip(merged["caption"], merged["file_name"]))
    assert caption_table({"images": images, "annotations": annotations}) == want


@pytest.mark.parametrize("workers", [0, 2])
def test_loader_batches_have_the_reference_shapes(tmp_path, workers):
    from imagegenerator_b200.data_loader import get_loader, MAX_LENGTH
    from imagegenerator_b200.train import image_transform
    root, ann, tok, captions = make_coco_dir(tmp_path, n_images=5, captions_per_image=2)
    loader = get_loader("ignored-bucket", root, ann, image_transform(64), batch_size=4, shuffle=False, tokenizer=tok,
                        num_workers=workers, prefetch_factor=2)
    assert len(loader.dataset) == 10 and len(loader) == 2                # drop_last (data_loader.py:101)
    batches = list(loader)
    assert len(batches) == 2
    seen = []
    for tokenized, imgs in batches:
        assert set(tokenized.keys()) >= {"input_ids", "attention_mask"}
        for v in tokenized.values():
            assert v.shape == (4, MAX_LENGTH) and v.dtype == torch.int64  # padding="max_length", max_length=128
        assert imgs.shape == (4, 3, 64, 64) and imgs.dtype == torch.float32
        assert -1.0 <= float(imgs.min()) and float(imgs.max()) <= 1.0     # Normalize(0.5, 0.5) of [0,1]
        seen += [tok.decode(r, skip_special_tokens=True) for r in tokenized["input_ids"]]
    assert seen == [c for c, _ in captions[:8]]                           # unshuffled: caption order of the join
    # the image of row 0 is its file, resized and normalised
    from PIL import Image
    import torchvision.transforms.functional as F
    want = F.normalize(F.resize(F.to_tensor(Image.open(os.path.join(root, captions[0][1])).convert("RGB")), [64, 64]),
                       [0.5] * 3, [0.5] * 3)
    assert torch.equal(batches[0][1][0], want)


def test_loader_shuffles_deterministically_and_feeds_the_text_side(tmp_path):
    from imagegenerator_b200.data_loader import get_loader
    from imagegenerator_b200.train import image_transform
    root, ann, tok, _ = make_coco_dir(tmp_path, n_images=6, captions_per_image=1)
    mk = lambda: get_loader("b", root, ann, image_transform(64), batch_size=3, shuffle=True, tokenizer=tok,
                            num_workers=0)
    a, b = list(mk()), list(mk())
    for (ta, ia), (tb, ib) in zip(a, b):
        assert torch.equal(ta["input_ids"], tb["input_ids"]) and torch.equal(ia, ib)
    enc = tiny_bert(tok)
    head = torch.nn.Linear(768, 512)
    tokenized, _ = a[0]
    perm = torch.tensor([2, 0, 1])
    mismatched = {k: v[perm] for k, v in tokenized.items()}               # stage_1_train_fn.py:123-126
    tem = head(enc(**tokenized).last_hidden_state[:, 0, :])                # :117-119
    tem_mis = head(enc(**mismatched).last_hidden_state[:, 0, :])
    assert tem.shape == (3, 512)
    assert torch.allclose(tem_mis, tem[perm], atol=1e-5)
    tem.backward(torch.ones_like(tem))                                     # the encoder trains through d lossG / d tem
    assert enc.embeddings.word_embeddings.weight.grad.abs().sum() > 0


def test_text_encoder_specs():
    from imagegenerator_b200.train import load_text_encoder, SyntheticEncoder
    assert isinstance(load_text_encoder(None, 16), SyntheticEncoder)
    m = torch.nn.Linear(2, 2)
    assert load_text_encoder(m) is m
    with pytest.raises(Exception):
        load_text_encoder("/nonexistent/spanbert")                        # no silent fallback to random weights
