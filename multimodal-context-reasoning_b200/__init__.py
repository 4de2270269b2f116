"""modcr_b200 -- B200-native (sm_100a) implementation of ModCR's training / scoring hot path.

Host side: Python/PyTorch modules with the reference's surface (``modeling_bert``,
``modeling_vcr_chunkalign_v10``, ``modeling_utils``).  Arithmetic: hand-written CUDA kernels in
``libmodcr_b200.so`` reached through the C ABI of ``include/modcr_b200.h`` (``_lib`` / ``ops``).
"""
__version__ = "0.1.0"

from . import synth  # noqa: F401


def __getattr__(name):
    # heavy submodules are imported lazily so that CPU-only tooling can import the package
    import importlib
    if name in ("ops", "engine", "functions", "modeling_bert", "modeling_transfomres", "modeling_utils",
                "modeling_vcr_chunkalign_v10", "_lib", "training", "parallel"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
