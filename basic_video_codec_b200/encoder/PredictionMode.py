"""Frame type enum (reference encoder/PredictionMode.py:4-9): the value is the container's mode byte."""
from enum import Enum


class PredictionMode(Enum):
    INTER_FRAME = 0
    INTRA_FRAME = 1

    def __str__(self):
        return self.name
