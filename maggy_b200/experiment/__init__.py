from .experiment import lagom, lagom_driver

__all__ = ["lagom", "lagom_driver"]
