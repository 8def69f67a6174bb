from .world_space import WorldSpaceFitter, guess_init_transl_from_root

__all__ = ["WorldSpaceFitter", "guess_init_transl_from_root"]
