"""larvanet_b200 -- B200-native (sm_100a) kernels + host glue for the LarvaNet SR forward/backward path.

The reference-facing plugin surface lives at the repo root (`models/`, `dataloaders/`, `train_larva.py`,
`validate.py`, `get_sr.py`, `runtime.py`); this package holds the CUDA sources (`csrc/`), the C-ABI loader
(`_lib.py`), thin tensor-level wrappers (`ops.py`), the fused inference/training engine (`engine.py`), autograd
wrappers for the individually-callable modules (`functional.py`) and the data-parallel plumbing (`dist.py`).
"""
__version__ = "0.1.0"
