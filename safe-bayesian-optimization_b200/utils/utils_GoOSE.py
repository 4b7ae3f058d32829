"""Drop-in mirror of the reference's ``utils/utils_GoOSE.py`` (utils_GoOSE.py:14-87): the SafeOpt plot
helpers plus the target marker drawn by plot_safe_region_Benoit."""
import numpy as np

from .utils_SafeOpt import (_plt, create_data_for_plot, create_frame, create_GIF,  # noqa: F401
                            plant_outputs_drawing)
from . import utils_SafeOpt as _base


def plot_safe_region_Benoit(X, X_0, X_1, mask_safe, obj, bound, data=None):
    _base.plot_safe_region_Benoit(X, X_0, X_1, mask_safe, obj, bound, data)
    if data is not None and 'x_target_0' in data:
        plt = _plt()
        plt.plot(data['x_target_0'], data['x_target_1'], 'kx', markersize=10)
        plt.plot(np.array([data['x_0'][-1], data['x_target_0']]),
                 np.array([data['x_1'][-1], data['x_target_1']]), 'k--')
