"""Pack the reference's hot-path sources into oracle/_ref/reference_src.zip -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_ref            (build container; needs /root/reference)

/root/reference does not exist on the GPU box, and the reference is plain Python with no build or
install step (no setup.py / pyproject.toml), so the only way its UNMODIFIED code can
be timed there (``bench.py --impl reference``) is to carry the ten files of the
path along.  This recipe copies them, byte for byte, from where they lie under /root/reference into ONE
archive under oracle/_ref/ -- git-ignored (the history stays free of reference sources) but not
gpurun-ignored (it travels like the built .so).  oracle/ref_harness.py compiles the modules from the
archive's text in memory (with the torch_xla / google.cloud.storage / SummaryWriter stubs and the two
one-token Stage-II fixes) exactly as it does from /root/reference.  Nothing under imagegenerator_b200/
reads it.
"""
from __future__ import annotations

import hashlib
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF_DIR, "reference_src.zip")
SOURCE_ROOT = os.environ.get("SG_REFERENCE_ROOT", "/root/reference")
# the modules, the loss helper, both train functions, the launcher and the loader
FILES = ["con_augment.py", "generator_1.py", "generator_2.py", "discrminator_1.py", "discriminator_2.py", "utils.py",
         "stage_1_train_fn.py", "stage_2_train_fn.py", "train.py", "data_loader.py"]


def make(force=False):
    """Returns the archive path, or None when the reference tree is not here (GPU box: the prebuilt archive is used)."""
    if not os.path.isfile(os.path.join(SOURCE_ROOT, "stage_1_train_fn.py")):
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    os.makedirs(REF_DIR, exist_ok=True)
    newest = max(os.path.getmtime(os.path.join(SOURCE_ROOT, f)) for f in FILES)
    if not force and os.path.isfile(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        return ARCHIVE
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        manifest = []
        for f in FILES:
            with open(os.path.join(SOURCE_ROOT, f), "rb") as fh:
                data = fh.read()
            z.writestr(f, data)
            manifest.append(f"{hashlib.sha256(data).hexdigest()}  {f}")
        z.writestr("MANIFEST.sha256", "\n".join(manifest) + "\n")
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
