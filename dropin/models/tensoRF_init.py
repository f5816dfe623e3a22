"""models.tensoRF_init of the reference (density/appearance-only TensoRF model)."""
from tensoir_b200.tensorbase import raw2alpha, AlphaGridMask  # noqa: F401
from tensoir_b200.tensorf_init import TensorVMSplit            # noqa: F401
