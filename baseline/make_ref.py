"""Recipe for ``baseline/_ref/``: the UNMODIFIED reference tree, made importable on the GPU box.

The reference (DJT777/xlstm-yolo, an Ultralytics 8.3.85 fork) has no setup.py / pyproject: its repo root IS
the ``ultralytics`` package directory, so there is nothing for ``pip install --target baseline/_ref`` to build.  The
equivalent of that install is placing the tree as a package:

    /root/reference/**  ->  baseline/_ref/ultralytics/**        (``.git``, caches, docs, tests and images left out)

``baseline/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels to the
GPU box with the snapshot.  ``bench_detector.py`` puts ``baseline/_ref`` on ``sys.path`` and imports ``ultralytics`` from
there; nothing in ``xlstm_yolo_b200`` imports it.  ``__graft_entry__.build()`` runs this when ``/root/reference`` exists.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MLSTM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "ultralytics")
SKIP_DIRS = {".git", "__pycache__", "docs", "tests", "assets", ".github", "examples", "docker"}
KEEP_EXT = {".py", ".yaml", ".yml", ".txt", ".json", ".cfg", ".toml"}


def make() -> str:
    """Returns baseline/_ref (the directory to put on sys.path), or "" when neither the reference nor a copy exists."""
    if os.path.isdir(REF_ROOT):
        stamp = os.path.join(DST, ".copied_from")
        if not os.path.exists(stamp):
            for root, dirs, files in os.walk(REF_ROOT):
                dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
                rel = os.path.relpath(root, REF_ROOT)
                out = os.path.join(DST, rel) if rel != "." else DST
                os.makedirs(out, exist_ok=True)
                for f in files:
                    if os.path.splitext(f)[1] in KEEP_EXT:
                        shutil.copyfile(os.path.join(root, f), os.path.join(out, f))
            with open(stamp, "w") as fh:
                fh.write(REF_ROOT + "\n")
        return os.path.dirname(DST)
    return os.path.dirname(DST) if os.path.isdir(DST) else ""


if __name__ == "__main__":
    out = make()
    print(out or "reference tree not found; baseline/_ref not made")
    sys.exit(0 if out else 1)
