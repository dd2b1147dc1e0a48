"""Host-side plumbing for the B200-native BPR hot path.

`rbpr.native`  — ctypes binding of the C ABI in include/rbpr.h (librbpr.so, sm_100a only).
`rbpr.engine`  — owns a native context for one (model, interaction matrix) pair.
`rbpr.synth`   — deterministic synthetic interaction matrices of the BASELINE.json shapes.

There is no CPU fallback anywhere in this package: if librbpr.so is missing or the device is
not a B200 the calls raise.
"""
