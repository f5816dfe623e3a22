"""models.tensoRF_rotated_lights of the reference, served by tensoir_b200 (train_tensoIR.py:12)."""
from tensoir_b200.tensorbase import *            # noqa: F401,F403
from tensoir_b200.tensorbase import raw2alpha, AlphaGridMask, positional_encoding  # noqa: F401
from tensoir_b200.tensorf import TensorVMSplit   # noqa: F401
