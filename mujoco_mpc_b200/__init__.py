"""B200-native rollout engine behind MJPC's planner/task surface."""
__version__ = "0.1.0"
