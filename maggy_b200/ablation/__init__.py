from .ablationstudy import AblationStudy

__all__ = ["AblationStudy"]
