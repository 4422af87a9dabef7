"""Installs the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, not gpurun-ignored: it travels to the
GPU box with the snapshot), so that `bench.py --impl reference` times the reference's own modules on the host cores there.

AntMMF has no setup.py / pyproject and `import antmmf` needs omegaconf, jsonlines, torchtext ... which are not in the image
(SURVEY.md §0.7), so `pip install --target baseline/_ref /root/reference` cannot work; the files below are plain PyTorch and import
through oracle/ref_loader.py's stub packages (B200MM_REFERENCE_ROOT=baseline/_ref). Nothing is edited: files are byte-copied and
their sha256 recorded in baseline/_ref/MANIFEST.json. No-op when /root/reference is absent (the GPU box uses the copies).
"""
import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("B200MM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "antmmf/modules/vision/backbone/clip/cn_model.py",
    "antmmf/modules/vision/backbone/clip/cn_tokenizer.py",
    "antmmf/modules/vision/backbone/clip/configuration_bert.py",
    "antmmf/modules/vision/backbone/clip/model.py",
    "antmmf/modules/vision/backbone/clip/modeling_bert.py",
    "antmmf/modules/vision/backbone/clip/vocab.txt",  # cn_tokenizer.py looks for it next to itself at import time
    "antmmf/utils/distributed_utils.py",
]


def install():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    print("installed" if install() else f"{SRC} not found: nothing installed", file=sys.stderr)
