"""Stand-in for NX-AI ``mlstm_kernels`` backed by xlstm_yolo_b200 (see compat/__init__.py)."""
__version__ = "0.0.0+b200"
