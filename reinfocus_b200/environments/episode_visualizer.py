"""Human-facing episode visualisation (reference environments/episode_visualizer.py).

Out of the hot path, kept for API parity and because ``visualize`` renders 600-px frames
through the SHARED renderer, which advances / re-creates its RNG states and so perturbs
every later observation. matplotlib and cv2 are imported
lazily: constructing an env must work on machines without them."""

from typing import Protocol

import numpy
from numpy.typing import NDArray

from reinfocus_b200 import histories
from reinfocus_b200.graphics import render


def fading_colours(cmap, max_n: int, n: int, p: int = 2):
    """``n`` colours of ``cmap`` fading towards transparent (reference :21-40)."""

    samples = numpy.linspace(1 - (n - 1) / max_n, 1, n) ** p
    colours = cmap(samples)
    colours[:, -1] = samples
    return colours


class IEpisodeVisualizer(Protocol):
    """The interface visualizers follow (reference episode_visualizer.py:43-84)."""

    def step(self, states, observations, indices=None):
        ...

    def reset(self, states, observations, indices=None):
        ...

    def visualize(self) -> NDArray[numpy.uint8]:
        ...


class HistoryVisualizer:
    # pylint: disable=too-many-instance-attributes
    """Plots each env's recent focus positions and values next to a rendering of its scene
    (reference :87-301)."""

    def __init__(self, num_envs: int, target_index: int, focus_plane_index: int,
                 focus_value_index: int, renderer: render.FastRenderer, limits: tuple[float, float],
                 ender=None, history_length: int = 10, target_radius: float | None = None):
        # pylint: disable=too-many-arguments
        self._num_envs = num_envs
        self._target_index = target_index
        self._focus_plane_index = focus_plane_index
        self._focus_value_index = focus_value_index
        self._limits = limits
        self._history_length = history_length
        self._target_radius = target_radius
        self._ender = ender
        self._renderer = renderer
        self._current_moves = numpy.zeros(num_envs, dtype=numpy.int32)
        self._targets = numpy.zeros(num_envs, dtype=numpy.float32)
        self._move_histories = histories.Histories(num_envs, history_length)
        self._focus_histories = histories.Histories(num_envs, history_length)

    def _all(self, indices):
        return numpy.full(self._num_envs, True) if indices is None else indices

    def step(self, states, observations, indices: NDArray[numpy.bool_] | None = None):
        indices = self._all(indices)
        self._current_moves[indices] += 1
        self._move_histories.append_events(states[:, self._focus_plane_index], indices)
        self._focus_histories.append_events(observations[:, self._focus_value_index], indices)

    def reset(self, states, observations, indices: NDArray[numpy.bool_] | None = None):
        indices = self._all(indices)
        self._current_moves[indices] = 0
        self._targets[indices] = states[:, self._target_index]
        self._move_histories.reset(indices)
        self._move_histories.append_events(states[:, self._focus_plane_index], indices)
        self._focus_histories.reset(indices)
        self._focus_histories.append_events(observations[:, self._focus_value_index], indices)

    def visualize(self) -> NDArray[numpy.uint8]:
        """Renderings (600 px, through the shared renderer) beside the history graphs."""

        renderings = self._renderer.render(600)
        rows = []
        for index, rendering in enumerate(renderings):
            graph = self._visualize_single_history(index, rendering.shape[0])
            rows.append(numpy.concatenate([rendering, graph], axis=1))
        return numpy.concatenate(rows, axis=0).astype(numpy.uint8)

    def _visualize_single_history(self, env_index: int, frame_height: int = 600) -> NDArray[numpy.uint8]:
        # pylint: disable=too-many-locals,import-outside-toplevel
        try:
            import cv2
            import matplotlib
            matplotlib.use("Agg")
            from matplotlib import pyplot
        except ImportError:
            # no plotting stack: a blank panel keeps the image layout
            return numpy.full((frame_height, frame_height * 4 // 3, 3), 255, dtype=numpy.uint8)

        focus_history = self._focus_histories.get_history(env_index)
        move_history = self._move_histories.get_history(env_index)
        target = self._targets[env_index]
        figure, axes = pyplot.subplots()
        axes.set_xlim(*self._limits)
        axes.set_ylim(-1.0, 1.0)
        label = f"focus position {self._current_moves[env_index]}\n"
        if self._ender is not None:
            label += self._ender.status(env_index)
        axes.set_xlabel(label)
        axes.set_ylabel("focus value")
        axes.axvline(x=target, linestyle=":", color="darkorange", label="target")
        if self._target_radius is not None and self._target_radius > 0.0:
            axes.axvspan(target - self._target_radius, target + self._target_radius,
                         edgecolor="darkorange", facecolor=("darkorange", 0.1), linestyle=(0, (5, 10)))
        blues = fading_colours(matplotlib.colormaps["Blues"], self._history_length, len(focus_history))
        previous = None
        for i, point in enumerate(zip(move_history, focus_history)):
            pyplot.plot(*point, color=blues[i], zorder=i, marker=".",
                        label="focus" if i == len(focus_history) - 1 else "")
            if previous is not None:
                axes.annotate("", xy=point, xycoords="data", xytext=previous, textcoords="data",
                              arrowprops={"arrowstyle": "->", "color": blues[i], "shrinkA": 5,
                                          "shrinkB": 5, "connectionstyle": "arc3,rad=0.1"})
            previous = point
        figure.legend(loc="lower right")
        figure.tight_layout()
        figure.canvas.draw()
        image = numpy.array(figure.canvas.buffer_rgba())[:, :, :3]
        pyplot.close(figure)
        width = int(frame_height * image.shape[1] / image.shape[0])
        return cv2.resize(image, (width, frame_height)).astype(numpy.uint8)
