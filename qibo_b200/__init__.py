"""qibo_b200 -- a B200-native (sm_100a) state-vector simulation backend for Qibo.

Plugin entry point: ``qibo.set_backend("qibo_b200")`` / ``qibo.set_backend("qibo-b200")`` imports this
package and calls :meth:`MetaBackend.load` (qibo/backends/__init__.py:325-350).  The qibo-independent layers
(``engine``, ``ops``, ``circuits``, the C ABI in ``include/qibo_b200.h``) can be used without qibo installed.
"""

__version__ = "0.1.0"

PLATFORMS = ("cuda-sm100a",)


class MetaBackend:
    """What Qibo's loader expects from a backend provider: ``load(**kwargs)`` and ``list_available()``."""

    @staticmethod
    def load(platform=None, **kwargs):
        if platform is not None and platform not in PLATFORMS:
            raise ValueError(f"Unsupported platform {platform} for qibo_b200, available: {PLATFORMS}.")
        from qibo_b200.backend import B200Backend  # imports qibo

        return B200Backend(**kwargs)

    def list_available(self) -> dict:
        import os

        from qibo_b200 import _lib

        try:
            import torch

            ok = os.path.exists(_lib.LIB_PATH) and torch.cuda.is_available()
        except Exception:  # pragma: no cover
            ok = False
        return {p: ok for p in PLATFORMS}
