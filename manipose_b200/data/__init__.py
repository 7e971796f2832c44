from .skeleton import Skeleton, h36m17_skeleton, skeleton_tables
from .windows import DeviceSequenceWindows

__all__ = ["Skeleton", "h36m17_skeleton", "skeleton_tables", "DeviceSequenceWindows"]
