"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.skeleton3d import Serializable, tensor_to_dict, _tensor_from_dict


class TensorUtils(Serializable):
    @classmethod
    def from_dict(cls, dict_repr, *args, **kwargs):
        return _tensor_from_dict(dict_repr)

    def to_dict(self):
        return NotImplemented


__all__ = ["TensorUtils", "tensor_to_dict"]
