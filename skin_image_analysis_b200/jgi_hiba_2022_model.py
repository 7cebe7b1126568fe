"""Drop-in for ``jgi_hiba_2022_model`` -- in the reference this file is byte-identical to
``tone_bias_model.py`` (same md5), so it simply re-exports the same classes and functions."""
from .tone_bias_model import (SkinCancerListModel, SkinCancerModel, create_loss_function, create_model,  # noqa: F401
                              load_model, save_model)
