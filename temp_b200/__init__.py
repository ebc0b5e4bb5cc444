"""temp_b200 -- B200-native RGCN + GRU/BiGRU/attention forward of TeMP."""
from .snapshot import Snapshot, SnapshotStore  # noqa: F401
from .planner import WindowPlan, plan_static, plan_window  # noqa: F401

__all__ = ["Snapshot", "SnapshotStore", "WindowPlan", "plan_window", "plan_static"]
