"""drvae_b200 — B200-native (sm_100a) training hot path of DrVAE / PertVAE / VFAE.

Public surface mirrors the reference (rampasek/DrVAE, src/): the DrVAE, PVAE and VFAE classes with
their constructor arguments, state_dict keys and methods.  Everything they compute goes through the
C ABI in include/drvae_b200.h; there is no CPU fallback.
"""
from .models import DrVAE, PVAE, VFAE  # noqa: F401
from .plan import Plan  # noqa: F401
from .training import DrVAEDataset, VFAEDataset, wrap_in_DrVAEDataset, wrap_in_VFAEDataset  # noqa: F401,E402
