"""Drop-in module path of the reference (`from src.solver import Solver`, reference
demo/solo_identification.py:5): re-exports the B200-native implementation."""
from system_identification_b200.solver import Solver  # noqa: F401
