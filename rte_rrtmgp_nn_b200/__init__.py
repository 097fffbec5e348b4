"""rte_rrtmgp_nn_b200 -- B200-native NN gas optics + RTE flux solvers behind the RTE+RRTMGP-NN API.

(The distribution is called rte-rrtmgp-nn_b200; a Python package name cannot contain '-'.)
Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI, built into lib/librrnn_b200.so),
the ctypes binding (_lib), the host mirror of the reference interface (api), the spectral tables the NN
path still needs (spectral) and synthetic inputs for the benchmark configurations (synth).
"""
from . import spectral, synth  # noqa: F401  (numpy only)

__all__ = ["spectral", "synth", "api", "_lib"]


def __getattr__(name):
    if name in ("api", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
