from .functional import pose_flip

__all__ = ["pose_flip"]
