"""models.tensoRF_general_multi_lights of the reference (train_tensoIR_general_multi_lights.py:12): the same class,
selected by passing ``light_name_list=`` (tensorBase_general_multi_lights.py:361)."""
from tensoir_b200.tensorbase import raw2alpha, AlphaGridMask  # noqa: F401
from tensoir_b200.tensorf import TensorVMSplit as _VM


class TensorVMSplit(_VM):
    def __init__(self, aabb, gridSize, device, light_name_list=("sunset", "snow", "courtyard"), **kargs):
        kargs.pop("light_rotation", None)
        super().__init__(aabb, gridSize, device, light_name_list=list(light_name_list), **kargs)
