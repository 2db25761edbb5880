"""Drop-in import path of the reference package: ``from monte_carloMPI import monte_carlo3D`` (reference
monte_carlo3D-run.py:4) resolves to the B200-native implementation in ``monte_carlompi_b200``."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)

from monte_carlompi_b200 import monte_carlo3D, parallelize  # noqa: E402,F401

sys.modules[__name__ + '.monte_carlo3D'] = monte_carlo3D
sys.modules[__name__ + '.parallelize'] = parallelize
