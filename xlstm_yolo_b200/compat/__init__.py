"""Compatibility shims that let the UNMODIFIED reference source import and run.

``install()`` puts a package named ``mlstm_kernels`` on ``sys.path`` that exports exactly
the symbols the reference imports from the (un-vendored, un-pinned) NX-AI package of that
name — ``mlstm_kernels.torch.backend_module`` (vision_lstm2.py:1327, mlstm_large.py:20-28,
xlstm/xlstm_large/model.py:3-14) and ``mlstm_kernels.torch.chunkwise.triton_xl_chunk``
(vision_lstm2.py:801) — backed by this repo's kernels.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install() -> None:
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
