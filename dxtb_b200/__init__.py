"""B200-native GFN1-xTB single points behind the dxtb calculator API (hot path only)."""
from . import calculators, exceptions
from .calculators import Calculator, GFN1Calculator

__version__ = "0.1.0"
__all__ = ["GFN1Calculator", "Calculator", "calculators", "exceptions"]
