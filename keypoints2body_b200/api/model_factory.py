"""Body-model loading (reference api/model_factory.py:19-40).

``smplx`` is only needed when the caller does not pass ``model=``; it is imported lazily so
the package works (with explicit models) where smplx is not installed.
"""

from __future__ import annotations

from pathlib import Path

from ..core.config import BodyModelConfig

MODEL_EXT_DEFAULTS = {"smpl": "pkl", "smplh": "pkl", "smplx": "npz", "mano": "pkl", "flame": "pkl"}


def load_body_model(config: BodyModelConfig, device=None):
    try:
        import smplx  # noqa: WPS433
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError("smplx is required to load body-model files; pass model= to skip it") from exc
    model_dir = Path(config.model_dir).expanduser()
    ext = config.ext or MODEL_EXT_DEFAULTS[config.model_type]
    # weights are read once on the host and uploaded by the fitter; no .to(device) needed
    return smplx.create(str(model_dir), model_type=config.model_type, gender=config.gender, ext=ext,
                        batch_size=config.batch_size)
