from .utils import Timer, TimerError, scale_to_bound

__all__ = ["Timer", "TimerError", "scale_to_bound"]
