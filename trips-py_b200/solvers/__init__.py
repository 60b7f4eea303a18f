"""Solver drivers with the reference's signatures (trips/solvers/{CGLS,Hybrid_LSQR,Hybrid_GMRES,GKS,MMGKS}.py)."""
from .CGLS import CGLS
from .GKS import GKS
from .Hybrid_GMRES import Hybrid_GMRES
from .Hybrid_LSQR import Hybrid_LSQR
from .MMGKS import MMGKS

__all__ = ["CGLS", "Hybrid_LSQR", "Hybrid_GMRES", "GKS", "MMGKS"]
