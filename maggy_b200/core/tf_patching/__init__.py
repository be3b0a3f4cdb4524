from .tf_modules import get_wrapped_model  # noqa: F401
