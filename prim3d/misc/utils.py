"""`prim3d.Timer` -- the wall-clock context manager the reference's examples wrap their calls in
(reference: prim3d/misc/utils.py:41-116; used at examples/sphere.py:14-25,
examples/sphere_tetrahedra.py:15,21).  Re-implemented to the same observable behaviour:

  * `Timer(print_tmpl)` starts immediately; leaving the `with` block prints
    `print_tmpl.format(seconds since the last check)`;
  * a template without a `{:.Nf}`-style field gets `" {:.3f}"` appended (so `Timer("cpu:")`
    prints `cpu: 0.123`); no template prints `{:.3f}`;
  * `since_start()` / `since_last_check()` raise TimerError when the timer is not running.

Like the reference it measures host wall-clock time and does not synchronise CUDA.
"""
import re
import time

from ..utility.marching_cubes import scale_to_bound  # the reference keeps a duplicate here (:10-31)

__all__ = ["Timer", "TimerError", "scale_to_bound"]

_FLOAT_FIELD = re.compile(r"({:.*\df})")


class TimerError(Exception):
    def __init__(self, message):
        super().__init__(message)
        self.message = message


class Timer:
    def __init__(self, print_tmpl=None, start=True):
        if print_tmpl is None or print_tmpl == "":
            print_tmpl = "{:.3f}"
        elif not _FLOAT_FIELD.findall(print_tmpl):
            print_tmpl = print_tmpl + " {:.3f}"
        self.print_tmpl = print_tmpl
        self._is_running = False
        if start:
            self.start()

    @property
    def is_running(self):
        return self._is_running

    def start(self):
        now = time.time()
        if not self._is_running:
            self._t_start = now
            self._is_running = True
        self._t_last = time.time()

    def _require_running(self):
        if not self._is_running:
            raise TimerError("timer is not running")

    def since_start(self):
        self._require_running()
        self._t_last = time.time()
        return self._t_last - self._t_start

    def since_last_check(self):
        self._require_running()
        now = time.time()
        elapsed = now - self._t_last
        self._t_last = time.time()
        return elapsed

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, exc_type, exc, tb):
        print(self.print_tmpl.format(self.since_last_check()))
        self._is_running = False
