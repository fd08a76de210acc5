"""
Stage the UNMODIFIED reference (tudelft/event_flow, MIT licence) next to the repo so that it travels to the GPU box:

    python tools/stage_reference.py            # /root/reference -> baseline/_ref  (git-ignored, NOT gpurun-ignored)

The reference is pure Python without a build system (no setup.py / pyproject build section), so `pip install` has nothing
to install; what `bench.py --impl reference` and baseline/ref_probe.py need are the importable modules of its two hot
paths.  They are copied byte for byte (a sha1 manifest is written beside them); nothing under baseline/_ref is product
code, nothing in event_flow_b200/ imports it, and it never enters the git history.
`__graft_entry__.build()` calls this when /root/reference exists (the build container); the GPU box uses the staged copy.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "LICENSE",
    "models/__init__.py", "models/base.py", "models/model.py", "models/model_util.py", "models/spiking_submodules.py",
    "models/spiking_util.py", "models/submodules.py", "models/unet.py",
    "loss/__init__.py", "loss/flow.py",
    "utils/__init__.py", "utils/iwe.py",
    "dataloader/__init__.py", "dataloader/encodings.py",
    "configs/__init__.py", "configs/parser.py", "configs/train_SNN.yml",
]  # fmt: skip


def stage(src=SRC, dst=DST, quiet=False):
    if not os.path.isdir(os.path.join(src, "models")):
        if not quiet:
            print(f"stage_reference: {src} not present (GPU box?): keeping {dst} as it is")
        return os.path.isdir(os.path.join(dst, "models"))
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha1(open(d, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "tudelft/event_flow (unmodified copies)", "sha1": manifest}, f, indent=1)
    if not quiet:
        print(f"stage_reference: {len(manifest)} files -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
