"""`from CEM import CEMnet` -> B200 implementation (reference: codes/CEM/CEMnet.py)."""
from esr_b200.cem import (CEMnet, CEM_PyTorch, Filter_Layer, Get_CEM_Config, Return_kernel,  # noqa: F401
                          Adjust_State_Dict_Keys)
