"""Fixed-length per-env ring histories (reference histories.py:10-76)."""

from collections.abc import Collection

import numpy
from numpy.typing import NDArray


class Histories:
    """``num_histories`` rows of the ``max_n`` most recent float32 events, oldest first,
    NaN where nothing has been recorded yet."""

    def __init__(self, num_histories: int, max_n: int):
        self.data = numpy.full((num_histories, max_n), numpy.nan, dtype=numpy.float32)

    def get_history(self, index: int) -> NDArray[numpy.float32]:
        row = self.data[index]
        return row[~numpy.isnan(row)]

    def most_recent_events(self) -> NDArray[numpy.float32]:
        return self.data[:, -1]

    def append_events(self, events: Collection[float], indices: NDArray[numpy.bool_] | None = None):
        """Shifts the selected rows left by one and writes one new event per selected row."""

        events = numpy.asarray(events, dtype=numpy.float32).reshape(-1)
        rows = slice(None) if indices is None else numpy.asarray(indices, dtype=bool)
        shifted = self.data[rows]
        shifted[:, :-1] = shifted[:, 1:]
        shifted[:, -1] = events
        self.data[rows] = shifted

    def reset(self, indices: Collection[bool]):
        self.data[numpy.asarray(indices, dtype=bool)] = numpy.nan
