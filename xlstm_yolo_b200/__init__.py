"""xlstm_yolo_b200 — B200-native mLSTM cell for DJT777/xlstm-yolo's ViL blocks.

Only the hot path lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI of
include/mlstm_b200.h), the ctypes binding, the autograd operator, and host-side mirrors of
the reference interfaces for this path (``mLSTMBackend`` seam, ``MatrixLSTMCell`` module, and the
ViL layer stack that calls it with flip-free bidirectional scans).
"""
from .backend import mLSTMBackend, mLSTMBackendConfig  # noqa: F401
from .cell import MatrixLSTMCell, MultiHeadLayerNorm  # noqa: F401
from .vil import SequenceTraversal, ViLBlock, ViLBlockPair, ViLLayer  # noqa: F401

__version__ = "0.1.0"
