"""cpecan_signal: host side of the B200 banded signal pair-HMM engine (see include/cpecan_cuda.h).

The compute path is the CUDA library only (libcpecan_cuda.so); importing this package never pulls in oracle/."""
from . import em, engine, hdp, synth  # noqa: F401
from .engine import (Engine, EngineError, HostBatch, default_params, echelon_hmm, four_state_hmm, hdp_hmm,  # noqa: F401
                     three_state_hmm, vanilla_gapx, vanilla_hmm)
