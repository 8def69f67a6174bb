from .camera_space import CameraSpaceFitter, guess_init_3d
from .world_space import WorldSpaceFitter, guess_init_transl_from_root

__all__ = ["CameraSpaceFitter", "WorldSpaceFitter", "guess_init_3d", "guess_init_transl_from_root"]
