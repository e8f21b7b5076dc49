"""Stage the reference's `whisper/` package under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

The reference is pure Python without a setup.py / pyproject ("no build system"), so
`pip install --target baseline/_ref /root/reference` has nothing to install; the equivalent - what pip would have
done - is a verbatim copy of the package directory.  Run by `__graft_entry__.build()` whenever /root/reference is
mounted (the authoring container); the GPU box only uses the staged files.  Nothing here is committed: the staged tree
is a build artefact exactly like oracle/_ref would be for a compiled reference, and bench.py only ever IMPORTS it for
the `--impl reference` arm and the GPU-eager comparator.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/whisper"
DST = os.path.join(ROOT, "baseline", "_ref", "whisper")
# live modules only (the dead model_all / model_ada / model_tmp variants are imported by nothing)
FILES = ("__init__.py", "audio.py", "decoding.py", "model.py", "resnet.py", "timing.py", "tokenizer.py",
         "transcribe.py", "triton_ops.py", "utils.py", "version.py")
DIRS = ("assets", "normalizers")


def stage(force: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force:
        same = all(os.path.exists(os.path.join(DST, f)) and
                   os.path.getsize(os.path.join(DST, f)) == os.path.getsize(os.path.join(SRC, f)) for f in FILES)
        if same:
            return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    for d in DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), ignore=shutil.ignore_patterns("__pycache__"))
    for base, dirs, files in os.walk(DST):  # the mount is read-only: make the copy writable so it can be refreshed
        for n in dirs + files:
            os.chmod(os.path.join(base, n), 0o755 if n in dirs else 0o644)
    return DST


if __name__ == "__main__":
    print(stage(force=True))
