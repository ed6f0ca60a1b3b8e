"""``mlstm_chunkwise__xl_chunk`` — imported (unused) by the reference at vision_lstm2.py:801."""
from xlstm_yolo_b200.backend import mLSTMBackend, mLSTMBackendConfig


def mlstm_chunkwise__xl_chunk(q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None,
                              return_last_states=False, eps=1e-6, chunk_size=128, autocast_kernel_dtype=None,
                              **kwargs):
    import torch
    kdt = {None: "bfloat16", torch.bfloat16: "bfloat16", torch.float16: "float16",
           torch.float32: "float32"}[autocast_kernel_dtype]
    be = mLSTMBackend(mLSTMBackendConfig(chunk_size=chunk_size, eps=eps, autocast_kernel_dtype=kdt))
    return be(q, k, v, i, f, c_initial, n_initial, m_initial, return_last_states)
