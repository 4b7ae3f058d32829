"""Drop-in mirror of the reference's ``utils/utils_GoOSE.py`` (utils_GoOSE.py:14-87): the SafeOpt plot
helpers plus the GoOSE target marker and the dashed line from the last query to the target."""
import numpy as np

from .utils_SafeOpt import (_plt, create_data_for_plot, create_frame, create_GIF,  # noqa: F401
                            draw_safe_region, plant_outputs_drawing)


def plot_safe_region_Benoit(X, X_0, X_1, mask_safe, obj, bound, data=None):
    fig = _plt().figure()
    ax = draw_safe_region(fig.gca(), np.asarray(X), X_0, X_1, mask_safe, obj, bound, data)
    if data is not None and 'x_target_0' in data and np.all(np.isfinite([data['x_target_0'], data['x_target_1']])):
        target = (float(data['x_target_0']), float(data['x_target_1']))
        ax.plot(*target, marker='x', color='k', markersize=10, linestyle='none')
        if len(data['x_0']):
            ax.plot([data['x_0'][-1], target[0]], [data['x_1'][-1], target[1]], color='k', linestyle='--')
