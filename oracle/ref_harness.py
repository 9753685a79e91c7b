"""Run the REAL reference (/root/reference) on CPU -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python on torch; it imports in the build container once
three things are stubbed:

  * ``torch_xla.core.xla_model``   (is_master_ordinal / all_reduce / optimizer_step / master_print)
  * ``google.cloud.storage``       (dict-backed fake bucket)
  * ``torch.utils.tensorboard.SummaryWriter`` (no-op; the reference opens gs:// paths)

and two one-token in-memory patches are applied to the Stage-II sources
(never copied to disk): ``stage_2_train_fn.py:67`` ``blob.`` -> ``blob_1.`` and
``discriminator_2.py:28`` ``self.down_sampler(x)`` -> ``self.down_sampler(img)``.

/root/reference does not exist on the GPU box: there the same sources come from the archive
oracle/_ref/reference_src.zip packed by ``oracle/make_ref.py`` (git-ignored, travels with gpurun).  Used by
``oracle/make_golden*.py``, by CPU tests that skip when neither is present, and by ``bench.py --impl reference`` /
its ``cpu_baseline`` leg (the unmodified reference timed on the host cores).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import random
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SG_REFERENCE_ROOT", "/root/reference")
# where /root/reference is absent (the GPU box): the archive of the same files packed by oracle/make_ref.py
REFERENCE_ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_src.zip")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "stage_1_train_fn.py")) or os.path.isfile(REFERENCE_ARCHIVE)


def reference_source(name: str):
    """(source text, origin path) of reference module ``name``: from the tree when it is here, else from the archive."""
    path = os.path.join(REFERENCE_ROOT, name + ".py")
    if os.path.isfile(path):
        with open(path, "r") as f:
            return f.read(), path
    import zipfile
    with zipfile.ZipFile(REFERENCE_ARCHIVE) as z:
        return z.read(name + ".py").decode("utf-8"), REFERENCE_ARCHIVE + "/" + name + ".py"


# --------------------------------------------------------------------------- stubs
class _Blob:
    def __init__(self, store, path):
        self._store, self._path = store, path

    def exists(self):
        return self._path in self._store

    def download_to_filename(self, fn):
        with open(fn, "wb") as f:
            f.write(self._store[self._path])

    def upload_from_filename(self, fn):
        with open(fn, "rb") as f:
            self._store[self._path] = f.read()

    def download_as_bytes(self):              # data_loader.py:38
        return self._store[self._path]

    def download_as_text(self):               # data_loader.py:50
        return self._store[self._path].decode("utf-8")


class FakeBucketStore(dict):
    """path -> bytes; shared by every storage.Client() the reference creates."""


_STORE = FakeBucketStore()


class _Bucket:
    def blob(self, path):
        return _Blob(_STORE, path)


class _Client:
    def get_bucket(self, name):
        return _Bucket()


class StepRecorder:
    """Hooked into xm.optimizer_step: snapshots grads at the moment of every step."""

    def __init__(self):
        self.events = []      # list of (optimizer_tag, {param_name: grad clone})
        self.names = {}       # id(optimizer) -> (tag, [(name, param)])
        self.prints = []

    def register(self, opt, tag, named_params):
        self.names[id(opt)] = (tag, list(named_params))

    def on_step(self, opt):
        if id(opt) in self.names:
            tag, named = self.names[id(opt)]
            self.events.append((tag, {n: (p.grad.detach().clone() if p.grad is not None else None)
                                      for n, p in named}))


RECORDER = StepRecorder()


def _install_stubs():
    if "torch_xla.core.xla_model" in sys.modules and getattr(
            sys.modules["torch_xla.core.xla_model"], "_sg_stub", False):
        return
    xm = types.ModuleType("torch_xla.core.xla_model")
    xm._sg_stub = True
    xm.is_master_ordinal = lambda: True
    xm.all_reduce = lambda op, t: t

    def optimizer_step(opt):
        RECORDER.on_step(opt)
        opt.step()

    xm.optimizer_step = optimizer_step
    xm.master_print = lambda *a, **k: RECORDER.prints.append(" ".join(str(x) for x in a))
    xm.xrt_world_size = lambda: 1
    xm.get_ordinal = lambda: 0
    core = types.ModuleType("torch_xla.core")
    core.xla_model = xm
    top = types.ModuleType("torch_xla")
    top.core = core
    sys.modules["torch_xla"] = top
    sys.modules["torch_xla.core"] = core
    sys.modules["torch_xla.core.xla_model"] = xm

    # google.cloud.storage -- do NOT shadow the top-level `google` package (protobuf lives there)
    import google  # noqa: F401  (namespace package from protobuf)
    cloud = sys.modules.get("google.cloud") or types.ModuleType("google.cloud")
    storage = types.ModuleType("google.cloud.storage")
    storage.Client = _Client
    cloud.storage = storage
    cloud.__path__ = getattr(cloud, "__path__", [])
    sys.modules["google.cloud"] = cloud
    sys.modules["google.cloud.storage"] = storage
    sys.modules["google"].cloud = cloud


class _NullWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass


_PATCHES = {
    "stage_2_train_fn": [("        blob.download_to_filename(tmp.name)\n        stage1_checkpoint",
                          "        blob_1.download_to_filename(tmp.name)\n        stage1_checkpoint")],
    "discriminator_2": [("x = self.down_sampler(x)", "x = self.down_sampler(img)")],
}

_CACHE: dict[str, types.ModuleType] = {}


def load(name: str) -> types.ModuleType:
    """Import reference module ``name`` (e.g. 'generator_1') from REFERENCE_ROOT.

    Modules are compiled from the source text in memory so the read-only tree
    never gets a __pycache__, and the two Stage-II one-token fixes are applied.
    """
    if name in _CACHE:
        return _CACHE[name]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    src, path = reference_source(name)
    for old, new in _PATCHES.get(name, []):
        if old not in src:
            raise RuntimeError("patch anchor not found in %s" % path)
        src = src.replace(old, new, 1)
    mod = types.ModuleType("sgref_" + name)
    mod.__file__ = path
    # the reference's intra-repo imports (`from utils import gradient_penalty`) resolve to our loads
    saved = {}
    for dep in ("utils",):
        if name != dep and ("from %s import" % dep) in src:
            saved[dep] = sys.modules.get(dep)
            sys.modules[dep] = load(dep)
    try:
        code = compile(src, path, "exec")
        if name.startswith("stage_"):
            import torch.utils.tensorboard as tb
            real = tb.SummaryWriter
            tb.SummaryWriter = _NullWriter
            try:
                exec(code, mod.__dict__)
            finally:
                tb.SummaryWriter = real
        else:
            exec(code, mod.__dict__)
    finally:
        for dep, old in saved.items():
            if old is None:
                sys.modules.pop(dep, None)
            else:
                sys.modules[dep] = old
    _CACHE[name] = mod
    return mod


# --------------------------------------------------------------------------- synthetic text side
class _EncOut:
    def __init__(self, h):
        self.last_hidden_state = h


class TableEncoder(torch.nn.Module):
    """Fake textEncoder: ``encoder(idx=LongTensor[B])`` -> last_hidden_state = table[idx][:, None, :].

    ``table`` is a Parameter so that the reference's lossG.backward() leaves
    d loss / d tem in ``table.grad`` (stage_1_train_fn.py:162-165)."""

    def __init__(self, table):
        super().__init__()
        self.table = torch.nn.Parameter(table.clone())

    def forward(self, idx):
        return _EncOut(self.table[idx][:, None, :])


class IdentityHead(torch.nn.Module):
    """projection_head stand-in: identity, with one dummy parameter for its optimizer."""

    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return x


# --------------------------------------------------------------------------- noise tape
class NoiseTape:
    """Records every RNG draw the reference train loop makes, in order."""

    def __init__(self):
        self.draws = []

    @contextlib.contextmanager
    def recording(self):
        names = ["randint", "randperm", "randn", "randn_like", "rand"]
        orig = {n: getattr(torch, n) for n in names}

        def wrap(n):
            def f(*a, **k):
                out = orig[n](*a, **k)
                self.draws.append((n, out.detach().clone()))
                return out
            return f

        for n in names:
            setattr(torch, n, wrap(n))
        try:
            yield self
        finally:
            for n in names:
                setattr(torch, n, orig[n])

    def by_kind(self, kind):
        return [t for (n, t) in self.draws if n == kind]


def quiet():
    """Context manager silencing the reference's debug prints."""
    return contextlib.redirect_stdout(io.StringIO())


def reset_store():
    _STORE.clear()
    RECORDER.events.clear()
    RECORDER.names.clear()
    RECORDER.prints.clear()


def store():
    return _STORE


def seed_everything(seed):
    torch.manual_seed(seed)
    random.seed(seed)
