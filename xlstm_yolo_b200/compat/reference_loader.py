"""Make the UNMODIFIED reference tree importable next to this package and plug the drop-ins in.

Used by ``bench_detector.py`` and the integration tests; nothing here is on the product path of the kernels.

    V = import_reference(root)            # root contains a directory named ``ultralytics`` (baseline/_ref, or a symlink farm)
    use_b200_dropins()                    # ViLBlockPair / MatrixLSTMCell of this repo inside the reference's model builders
    apply_head_fixes(V, backend=...)      # the reference arm: HEAD's two documented breakages repaired on its own classes

HEAD breakages this handles (file:line in the reference):
  * utils/__init__.py:24 imports matplotlib at module scope (absent in this image): mocked, plotting is never reached;
  * vision_lstm2.py:801,1327 import ``mlstm_kernels`` (un-vendored): ``compat.install()`` provides the package, backed by this
    repo's kernels — or, for the reference arm, by the reference's own PyTorch ``chunkwise_simple``;
  * vision_lstm2.py:950-952: the cell's out-norm + head merge are commented out, so ViLLayer.forward fails at :498;
  * vision_lstm2.py:1438-1441: ViLBlockPair.forward returns the top-left block only (the bottom-right block never runs).
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock


def import_reference(root: str):
    """Import ``ultralytics`` from ``root`` (a directory that contains the reference tree as ``ultralytics/``) with this
    repo's ``mlstm_kernels`` shim installed.  Returns the reference's vision_lstm2 module."""
    if not os.path.isdir(os.path.join(root, "ultralytics")):
        raise FileNotFoundError(f"{root}/ultralytics not found: run baseline/make_ref.py where /root/reference exists")
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.backends",
              "matplotlib.backends.backend_agg"]:
        sys.modules.setdefault(m, MagicMock())
    from . import install
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    from ultralytics.nn.modules.vision_lstm import vision_lstm2 as V
    return V


def use_b200_dropins(pair_level: bool = True, head_compat: bool = False):
    """Put this repo's classes where the reference's builders look them up: ``ViLBlockPair`` in nn/modules/block.py
    (constructed by ViLBlockPairBlock, block.py:1815) when ``pair_level``, else only ``MatrixLSTMCell`` in vision_lstm2.py
    (constructed by ViLLayer.__init__, :439).  Call after import_reference and before building the model."""
    import ultralytics.nn.modules.block as block
    from ultralytics.nn.modules.vision_lstm import vision_lstm2 as V

    import xlstm_yolo_b200 as X
    if pair_level:
        X.ViLBlockPair.head_compat = bool(head_compat)
        block.ViLBlockPair = X.ViLBlockPair
    else:
        V.MatrixLSTMCell = X.MatrixLSTMCell
        if not head_compat:
            _compose_pair(V)


def _compose_pair(V):
    def forward(self, x, seqlens=None):   # the intended BR(TL(x)) (…checkpoint.py:1406-1408)
        return self.rowwise_from_bot_right(self.rowwise_from_top_left(x))
    V.ViLBlockPair.forward = forward


def apply_head_fixes(V, reference_backend: bool = True, compose_pair: bool = True):
    """Reference arm: the reference's own classes, with (a) the cell epilogue of vision_lstm2.py:950-952 applied to the cell's
    raw output, (b) optionally the intended pair composition, and (c) ``reference_backend``: every mLSTMBackend of the cell
    replaced by a module that calls the reference's own PyTorch ``chunkwise_simple`` (backends.py:149) on whatever device the
    tensors live on — the reference's arithmetic without the absent Triton package."""
    import torch

    head_forward = V.MatrixLSTMCell.forward

    def cell_forward(self, q, k, v):
        h = head_forward(self, q, k, v)
        if h.dim() == 4:                                # HEAD returns raw (B,NH,S,DH)
            B, NH, S, DH = h.shape
            h = self.outnorm(h).transpose(1, 2).reshape(B, S, NH * DH)
        return h

    V.MatrixLSTMCell.forward = cell_forward
    if compose_pair:
        _compose_pair(V)
    if reference_backend:
        from ultralytics.nn.modules.vision_lstm.xlstm.blocks.mlstm import backends as rb

        class RefBackend(torch.nn.Module):
            def __init__(self, config):
                super().__init__()
                self.config = config

            def forward(self, q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None, return_last_states=None,
                        mode=None):
                S, L = q.shape[2], int(self.config.chunk_size)
                while S % L:                            # chunkwise_simple needs L | S; its output is chunk-size invariant
                    L -= 1
                dt = q.dtype
                out = rb.chunkwise_simple(q.float().contiguous(), k.float().contiguous(), v.float().contiguous(),
                                          i.float().contiguous(), f.float().contiguous(), chunk_size=L, eps=self.config.eps)
                return out.to(dt)

        init = V.MatrixLSTMCell.__init__

        def cell_init(self, *a, **kw):
            init(self, *a, **kw)
            for name in ("cpu_backend", "cpu_backend_infer", "gpu_backend", "gpu_backend_infer"):
                setattr(self, name, RefBackend(getattr(self, name).config))

        V.MatrixLSTMCell.__init__ = cell_init
