"""Recipe for ``oracle/_ref/``: the reference's own CPU implementation of the hot path, made available to
``bench.py --impl reference`` on the GPU box (where ``/root/reference`` does not exist).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is placing its one
self-contained source file for this path where the bench can import it by file path:

    /root/reference/nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py   (chunkwise_simple :149,
        parallel_stabilized_simple :9, recurrent_step_stabilized_simple :93; imports only math, typing, torch)
    -> oracle/_ref/backends.py

``oracle/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the file travels
to the GPU box with the snapshot like the built ``.so``.  ``__graft_entry__.build()`` runs this when the reference tree
is present; on the GPU box the prebuilt copy is used as is.  Nothing in ``xlstm_yolo_b200`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MLSTM_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF_ROOT, "nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py")
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "backends.py")


def make() -> str:
    """Returns the path of oracle/_ref/backends.py, or "" when neither the reference tree nor a previous copy exists."""
    if os.path.exists(SRC):
        os.makedirs(DST_DIR, exist_ok=True)
        if not os.path.exists(DST) or os.path.getmtime(DST) < os.path.getmtime(SRC):
            shutil.copyfile(SRC, DST)
        return DST
    return DST if os.path.exists(DST) else ""


def load():
    """The reference module, imported by file path (None when oracle/_ref is absent)."""
    import importlib.util
    if not os.path.exists(DST):
        return None
    spec = importlib.util.spec_from_file_location("xlstm_yolo_reference_backends", DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = make()
    print(out or "reference tree not found; oracle/_ref not made")
    sys.exit(0 if out else 1)
