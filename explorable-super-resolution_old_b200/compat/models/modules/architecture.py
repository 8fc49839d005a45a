"""`models.modules.architecture.RRDBNet` -> B200 generator (reference: codes/models/modules/architecture.py:102-175)."""
from esr_b200.rrdbnet import RRDBNet  # noqa: F401
