"""Backend routing with the reference's interface (``utils/backend.py:49-129``), for the one
route this package implements: QuanONet/HEAQNN with ``quantum_backend='torchquantum'`` →
``'pytorch_quantum'`` — served by the B200 CUDA module, so ``torchquantum`` itself is NOT required
(the reference's check at ``utils/backend.py:76-80`` is replaced by "is the CUDA library built").
Every other combination raises exactly the kind of error the reference raises when a backend is
missing, because those backends are out of this package's scope.
"""
from __future__ import annotations

import os


class BackendManager:
    quantum_models = ("QuanONet", "HEAQNN")
    classical_models = ("DeepONet", "FNN", "FNO")

    @property
    def is_torch_available(self):
        try:
            import torch  # noqa: F401
            return True
        except ImportError:
            return False

    @property
    def is_b200_library_built(self):
        from .. import _lib
        return os.path.exists(_lib.LIB_PATH)

    # kept for callers that probe it; the B200 module stands in for torchquantum
    @property
    def is_torchquantum_available(self):
        return self.is_b200_library_built

    def check_compatibility(self, model_type, quantum_backend="mindquantum", classical_backend="pytorch"):
        if model_type in self.quantum_models:
            if quantum_backend == "torchquantum":
                if not self.is_torch_available:
                    raise ImportError(f"Model '{model_type}' with torchquantum backend requires PyTorch, "
                                      "but it is not installed.")
                if not self.is_b200_library_built:
                    raise ImportError(f"Model '{model_type}' with the B200 quantum backend requires "
                                      "libquanonet_b200.so; build it with `python -m quanonet_b200.build`.")
                return "pytorch_quantum"
            if quantum_backend in ("mindquantum", "qiskit", "pennylane"):
                raise ImportError(f"Model '{model_type}' with {quantum_backend} backend is outside quanonet_b200; "
                                  "use the reference implementation for it.")
            raise ValueError(f"Unknown quantum_backend: '{quantum_backend}'. "
                             "Choose from: mindquantum, torchquantum, qiskit, pennylane")
        if model_type in self.classical_models:
            if classical_backend in ("pytorch", "mindspore"):
                raise ImportError(f"Model '{model_type}' is a classical baseline outside quanonet_b200; "
                                  "use the reference implementation for it.")
            raise ValueError(f"Unknown classical_backend: '{classical_backend}'. Choose from: pytorch, mindspore")
        return "unknown"


backend = BackendManager()
