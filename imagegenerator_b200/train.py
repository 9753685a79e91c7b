"""Launcher -- the B200 counterpart of the reference's ``train.py:60-169``.

    torchrun --nnodes=1 --nproc-per-node N -m imagegenerator_b200.train --stage 1 [--epochs E] [--batch-size B] ...
    python -m imagegenerator_b200.train --stage 2 --synthetic 8          (single GPU)

Same constants (``TEM_SIZE, lr, c_dim, z_dim, Nd, num_epochs, batch_size``, train.py:32-38), the same seed
(``torch.manual_seed(42)``, :66) and the same construction order of models (:68-75), optimizers (:89-102) and
StepLR(100, 0.5) schedulers (:105-113), then ``train_1`` with the argument lists of :136-164 -- and the Stage-II
driver the reference never wrote (``train_2`` is imported at :19 but not called).  What differs: one process per
GPU under torchrun with NCCL instead of ``xmp.spawn`` on a TPU (:167-169, :63); parameters are broadcast from rank 0
inside the engines (pjrt.broadcast_master_param, :78-85); checkpoints go to ``--save-dir``.

The text side (SpanBERT + COCO captions, train.py:68, data_loader.py) needs weights and data this box does not
have: ``--synthetic N`` trains on N random batches (a fixed embedding table stands in for the encoder's CLS state),
which is what the benchmarks use.  With the files on local disk, ``--coco-root DIR --ann-file captions.json
--tokenizer DIR --text-encoder DIR`` runs the reference's real pipeline (``data_loader.get_loader`` + a BERT encoder
trained through ``d lossG / d tem``, stage_1_train_fn.py:161-171); ``--text-encoder spanbert-random`` gives the same
architecture with random weights.
"""
import argparse
import os

import torch
import torch.distributed as dist
from torch import nn, optim
from torch.optim.lr_scheduler import StepLR

from .con_augment import ConditioningAugmentation
from .discrminator_1 import StageIDiscriminator
from .discriminator_2 import StageIIDiscriminator
from .generator_1 import StageIGenerator
from .generator_2 import StageIIGenerator
from .stage_1_train_fn import train_1
from .stage_2_train_fn import train_2

TEM_SIZE = 512      # train.py:32
lr = 1e-3           # :33
c_dim = 128         # :34
z_dim = 100         # :35
Nd = 128            # :36
num_epochs = 500    # :37
batch_size = 32     # :38


class SyntheticEncoder(nn.Module):
    """Stands in for ``AutoModel.from_pretrained("SpanBERT/spanbert-base-cased")`` (train.py:68): ``encoder(idx=...)``
    returns an object whose ``last_hidden_state[:, 0, :]`` is a row of a fixed 768-d table."""

    def __init__(self, n_rows, dim=768, seed=1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("table", torch.randn(n_rows, dim, generator=g))
        self.dummy = nn.Parameter(torch.zeros(1))

    def forward(self, idx):
        class _Out:
            pass
        o = _Out()
        o.last_hidden_state = self.table[idx][:, None, :] + 0.0 * self.dummy
        return o


def load_text_encoder(spec=None, n_rows=4096):
    """``train.py:68``'s ``AutoModel.from_pretrained("SpanBERT/spanbert-base-cased")`` without a network: ``None`` ->
    the synthetic table above; ``"spanbert-random"`` -> a randomly initialised BERT of SpanBERT-base-cased's shape
    (12 layers, 768 hidden, 12 heads, vocabulary 28996); anything else -> a local directory for ``AutoModel``;
    a module passes through."""
    if spec is None:
        return SyntheticEncoder(n_rows)
    if isinstance(spec, nn.Module):
        return spec
    if spec == "spanbert-random":
        from transformers import BertConfig, BertModel
        return BertModel(BertConfig(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                    intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2),
                         add_pooling_layer=False)
    from transformers import AutoModel
    return AutoModel.from_pretrained(spec, local_files_only=True)


def image_transform(hw):
    """``my_transform_1`` / ``my_transform_2`` of train.py:40-54."""
    import torchvision.transforms as transforms
    return transforms.Compose([transforms.ToTensor(), transforms.Resize((hw, hw)),
                               transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])


class SyntheticLoader:
    """``n_batches`` batches shaped like the reference loader's items (data_loader.py:64-108): a dict of tensors for
    the text encoder and an image batch normalised to (-1, 1) (train.py:40-54)."""

    def __init__(self, n_batches, batch, hw, n_rows, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.items = [({"idx": torch.randint(0, n_rows, (batch,), generator=g)},
                       torch.randn(batch, 3, hw, hw, generator=g).clamp_(-1, 1).pin_memory()) for _ in range(n_batches)]

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def build(device, text_encoder=None, n_rows=4096):
    """Models, optimizers and schedulers in the reference's order (train.py:66-113)."""
    torch.manual_seed(42)
    m = {}
    m["textEncoder"] = load_text_encoder(text_encoder, n_rows).to(device)
    m["projection_head"] = nn.Linear(768, TEM_SIZE).to(device)
    m["con_augment_1"] = ConditioningAugmentation(TEM_SIZE, 256, c_dim)
    m["critic_1"] = StageIDiscriminator(TEM_SIZE, Nd)
    m["gen_1"] = StageIGenerator(c_dim, z_dim)
    m["con_augment_2"] = ConditioningAugmentation(TEM_SIZE, 256, c_dim)
    m["critic_2"] = StageIIDiscriminator(TEM_SIZE, Nd)
    m["gen_2"] = StageIIGenerator()
    o = {"textEncoder": optim.AdamW(m["textEncoder"].parameters(), lr=5e-5)}
    for k in ("projection_head", "con_augment_1", "critic_1", "gen_1", "con_augment_2", "critic_2", "gen_2"):
        o[k] = optim.Adam(m[k].parameters(), lr=lr, betas=(0.9, 0.999))
    s = {k: StepLR(v, step_size=100, gamma=0.5) for k, v in o.items()}
    return m, o, s


def run(stage, device, epochs=num_epochs, batch=batch_size, loader=None, text_encoder=None, save_dir="./checkpoints",
        synthetic=0, log=print, use_graph=True, preview_every=100):
    n_rows = 4096
    m, o, s = build(device, text_encoder, n_rows)
    if loader is None:
        assert synthetic > 0, "no loader given: pass --synthetic N (the COCO/GCS loader of the reference is out of scope)"
        loader = SyntheticLoader(synthetic, batch, 64 if stage == 1 else 256, n_rows,
                                 seed=dist.get_rank() if dist.is_initialized() else 0)
    if stage == 1:
        keys = ["textEncoder", "projection_head", "con_augment_1", "critic_1", "gen_1"]
        return train_1([m[k] for k in keys], [o[k] for k in keys], [s[k] for k in keys], loader, epochs, device, batch,
                       save_dir=os.path.join(save_dir, "Stage1"), log=log, use_graph=use_graph), m
    keys = ["con_augment_2", "critic_2", "gen_2"]
    models = [m[k] for k in ("textEncoder", "projection_head", "con_augment_1", "con_augment_2", "gen_1", "critic_2", "gen_2")]
    return train_2(models, [o[k] for k in keys], [s[k] for k in keys], loader, epochs, device, batch,
                   save_dir=os.path.join(save_dir, "Stage2"), log=log, use_graph=use_graph, preview_every=preview_every,
                   stage1_checkpoint=os.path.join(save_dir, "Stage1", "latest_checkpoint_stage1.pth")), m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, choices=(1, 2), default=1)
    ap.add_argument("--epochs", type=int, default=num_epochs)
    ap.add_argument("--batch-size", type=int, default=batch_size)
    ap.add_argument("--save-dir", default="./checkpoints")
    ap.add_argument("--synthetic", type=int, default=0, help="train on N random batches instead of the COCO loader")
    ap.add_argument("--coco-root", help="directory with the images (train.py:118 'dataset/train2017', on local disk)")
    ap.add_argument("--ann-file", help="COCO captions JSON (train.py:119)")
    ap.add_argument("--tokenizer", help="local directory with the SpanBERT tokenizer files")
    ap.add_argument("--text-encoder", help="local directory with the SpanBERT weights, or 'spanbert-random'")
    ap.add_argument("--num-workers", type=int, default=8)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=device)
    loader = None
    if args.coco_root:
        from .data_loader import get_loader
        loader = get_loader("local", args.coco_root, args.ann_file, image_transform(64 if args.stage == 1 else 256),
                            batch_size=args.batch_size, shuffle=True, tokenizer=args.tokenizer,
                            num_workers=args.num_workers)
    run(args.stage, device, args.epochs, args.batch_size, loader=loader, text_encoder=args.text_encoder,
        save_dir=args.save_dir, synthetic=args.synthetic)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
