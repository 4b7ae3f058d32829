"""Drop-in mirror of the reference's ``utils/utils_SafeOpt.py``.

Same functions and signatures (plot_safe_region_Benoit, create_frame, create_GIF, plant_outputs_drawing,
reference utils_SafeOpt.py:14-84).  matplotlib / imageio are imported lazily so the module imports on a
box without them.  ``create_data_for_plot`` is the grid-mask producer the reference keeps in its driver
(test/test_SafeOpt.py:324-345); here the 400x400 posterior comes from the CUDA grid pipeline.
"""
import os

import numpy as np


def _plt():
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    return plt


def create_data_for_plot(GP_m, plant_system, bound=None, n_grid=400, index=1):
    """test/test_SafeOpt.py:324-345: X_0, X_1 (n,n), mask_safe = lcb_index > 0 (n,n) bool, obj (n,n)."""
    bound = GP_m.bound if bound is None else np.asarray(bound, dtype=np.float64)
    x_0 = np.linspace(bound[0, 0], bound[0, 1], n_grid)
    x_1 = np.linspace(bound[1, 0], bound[1, 1], n_grid)
    X_0, X_1 = np.meshgrid(x_0, x_1)
    saved = (GP_m.grid_points_per_dim, GP_m._grid_set)
    GP_m.grid_points_per_dim, GP_m._grid_set = n_grid, False
    mean, var = GP_m.grid_posterior()                       # (N,G) on the B200, p = r*n + c
    GP_m.grid_points_per_dim, GP_m._grid_set = saved[0], False
    lcb = mean[:, index] - GP_m.b * np.sqrt(var[:, index])
    mask_safe = lcb.reshape(X_0.shape) > 0.
    pts = np.column_stack((X_0.ravel(), X_1.ravel()))
    obj = np.array([plant_system[0](p) for p in pts]).reshape(X_0.shape) if plant_system is not None else None
    return X_0, X_1, mask_safe, obj


# The optimum and the active constraint of the Benoit problem drawn on every frame (reference utils_SafeOpt.py:28-31)
BENOIT_OPTIMUM = (0.36845785, -0.39299271)


def _benoit_constraint_curve(n=400):
    """u_0 = 1 + u_1^2 + 2 u_1 for u_1 in [-1.5, 1.5]: the boundary of con1_system_tight."""
    u1 = np.linspace(-1.5, 1.5, n)
    return 1. + u1 * (u1 + 2.), u1


def draw_safe_region(ax, X, X_0, X_1, mask_safe, obj, bound, trajectory=None):
    """Paint one frame on a matplotlib Axes: safe / unsafe shading from the grid mask, objective contours, the true
    constraint boundary, the optimum, the initial samples and (optionally) the queried trajectory."""
    shading = dict(levels=[0., 0.5, 1.], colors=['lightcoral', 'lightblue'], alpha=0.4)
    ax.contourf(X_0, X_1, np.asarray(mask_safe, dtype=float), **shading)
    if obj is not None:
        iso = ax.contour(X_0, X_1, np.reshape(obj, np.shape(X_0)), colors='k', linestyles='dashed', linewidths=0.5)
        ax.clabel(iso, inline=True)
    ax.plot(*_benoit_constraint_curve(), color='k')
    ax.plot(*BENOIT_OPTIMUM, marker='o', color='r', linestyle='none')
    ax.plot(X[:, 0], X[:, 1], marker='o', color='b', linestyle='none')
    if trajectory is not None:
        q0, q1 = np.asarray(trajectory['x_0'], dtype=float), np.asarray(trajectory['x_1'], dtype=float)
        ax.plot(q0, q1, marker='o', color='k', markersize=5, linewidth=0.5, label='_nolegend_')
    lo_hi = np.asarray(bound, dtype=float)
    ax.set_xlim(lo_hi[0])
    ax.set_ylim(lo_hi[1])
    return ax


def plot_safe_region_Benoit(X, X_0, X_1, mask_safe, obj, bound, data=None):
    """Reference signature (utils_SafeOpt.py:14-43): opens a new current figure and draws the frame on it."""
    fig = _plt().figure()
    draw_safe_region(fig.gca(), np.asarray(X), X_0, X_1, mask_safe, obj, bound, data)


def create_frame(fun_drawing, filename):
    """Reference signature (utils_SafeOpt.py:45-48): the drawing call has already run as the argument expression;
    write the current figure to ``filename`` and release it."""
    plt = _plt()
    plt.gcf().savefig(filename)
    plt.close('all')


def create_GIF(frame_duration, filenames, GIFname, output_dir='output'):
    """Reference signature (utils_SafeOpt.py:50-60): assemble the frames into output_dir/GIFname, then delete them."""
    import imageio.v2 as iio
    frames = [iio.imread(f) for f in filenames]
    iio.mimsave(os.path.join(output_dir, GIFname), frames, duration=frame_duration, loop=0)
    for f in filenames:
        os.remove(f)


def plant_outputs_drawing(iteration, output, constraint, figname, output_dir='output'):
    """Reference signature (utils_SafeOpt.py:62-84): objective and constraint of the queried points per iteration,
    with the safety threshold g = 0."""
    fig, (top, bottom) = _plt().subplots(nrows=2, ncols=1, figsize=(5, 10))
    it = np.asarray(iteration)
    for ax, series, label in ((top, output, 'Plant Output'), (bottom, constraint, 'Plant Constraint')):
        ax.plot(it, np.asarray(series, dtype=float))
        ax.set_xlabel('Iteration', fontsize=14)
        ax.set_ylabel(label, fontsize=14)
    bottom.axhline(0., color='r', linestyle='--', label='safety threshold')
    bottom.legend()
    fig.tight_layout()
    fig.savefig(os.path.join(output_dir, figname))
