from xlstm_yolo_b200.backend import (  # noqa: F401
    BackendModeType,
    ChunkwiseKernelType,
    DtypeType,
    SequenceKernelType,
    StepKernelType,
    mLSTMBackend,
    mLSTMBackendConfig,
)
