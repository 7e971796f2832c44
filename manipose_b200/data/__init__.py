from .skeleton import Skeleton, h36m17_skeleton, skeleton_tables

__all__ = ["Skeleton", "h36m17_skeleton", "skeleton_tables"]
