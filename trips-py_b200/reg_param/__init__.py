"""Host-side (NumPy) regularisation-parameter rules for the projected problems."""
from .discrepancy_principle import discrepancy_principle, discrepancy_principle_projected
from .gcv import generalized_crossvalidation
from .l_curve import l_curve

__all__ = ["generalized_crossvalidation", "discrepancy_principle", "discrepancy_principle_projected", "l_curve"]
