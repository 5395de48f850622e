"""psvi -- drop-in, B200-native replacement of the PSVI hot path of souravc83/Blackbox-Coresets-VI.

Same module paths and names as the reference package (`psvi.models.neural_net`, `psvi.inference.psvi_classes`,
`psvi.inference.baselines`, `psvi.robust_higher`, `psvi.hypergrad`, `psvi.experiments.flow_psvi`); the math runs in
hand-written sm_100a CUDA kernels reached through `psvi._native` (ctypes over the C ABI in include/psvi_b200.h).
"""
__version__ = "0.1.0"
