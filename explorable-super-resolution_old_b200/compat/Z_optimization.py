"""`from Z_optimization import Z_optimizer` -> B200 loop (reference: codes/Z_optimization.py)."""
from esr_b200.z_optimization import Z_optimizer, Optimizable_Z, ArcTanH, TV_Loss  # noqa: F401
