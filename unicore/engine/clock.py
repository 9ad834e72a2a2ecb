"""Cumulative training time across restarts (checkpoint key ``extra_state.previous_training_time``)."""
import time

import torch

from unicore import utils


class TrainingClock:
    def __init__(self):
        self._origin = time.time()
        self._carried = 0.0      # seconds trained before this process started (from the checkpoint)
        self._fleet_mean = None  # mean over ranks as of the last update (number or device scalar)

    def local(self) -> float:
        return time.time() - self._origin + self._carried

    def resume_from(self, seconds: float) -> None:
        self._carried = float(seconds)
        self._origin = time.time()

    def set_fleet_mean(self, value) -> None:
        """``value`` may be a device scalar; it is converted only when somebody asks (checkpoint, stop-time test)."""
        self._fleet_mean = value

    def cumulative(self) -> float:
        if self._fleet_mean is None:
            return self.local()
        if torch.is_tensor(self._fleet_mean):
            self._fleet_mean = float(utils.item(self._fleet_mean))
        return self._fleet_mean
