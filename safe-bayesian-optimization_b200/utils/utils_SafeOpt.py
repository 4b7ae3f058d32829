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


def plot_safe_region_Benoit(X, X_0, X_1, mask_safe, obj, bound, data=None):
    plt = _plt()
    plt.figure()
    plt.contourf(X_0, X_1, mask_safe, levels=[0., 0.5, 1.], colors=['lightcoral', 'lightblue'], alpha=0.4)
    CS1 = plt.contour(X_0, X_1, obj.reshape(X_0.shape), colors='k', linestyles='dashed', linewidths=0.5)
    plt.clabel(CS1, inline=True)
    x_0 = np.linspace(-1.5, 1.5, 400)
    plt.plot(1. + x_0 ** 2 + 2. * x_0, x_0, 'k')            # tight constraint
    plt.plot(0.36845785, -0.39299271, 'ro')                 # constrained optimum
    plt.plot(X[:, 0], X[:, 1], 'bo')
    if data is not None:
        plt.plot(data['x_0'][:], data['x_1'][:], 'ko', linewidth=1., markersize=5)
        plt.plot(data['x_0'][:], data['x_1'][:], 'k-', linewidth=0.5, label='_nolegend_')
    plt.axis((bound[0, 0], bound[0, 1], bound[1, 0], bound[1, 1]))


def create_frame(fun_drawing, filename):
    plt = _plt()
    plt.savefig(filename)
    plt.close()


def create_GIF(frame_duration, filenames, GIFname, output_dir='output'):
    import imageio.v2 as imageio
    with imageio.get_writer(os.path.join(output_dir, GIFname), mode='I', duration=frame_duration, loop=0) as writer:
        for filename in filenames:
            writer.append_data(imageio.imread(filename))
    for filename in filenames:
        os.remove(filename)


def plant_outputs_drawing(iteration, output, constraint, figname, output_dir='output'):
    plt = _plt()
    plt.figure()
    fig, axs = plt.subplots(2, 1, figsize=(5, 10))
    axs[0].plot(iteration, output)
    axs[0].set_xlabel('Iteration', fontsize=14)
    axs[0].set_ylabel('Plant Output', fontsize=14)
    axs[1].plot(iteration, constraint)
    axs[1].plot(iteration, np.array([0.] * len(iteration)), 'r--', label='safety threshold')
    axs[1].set_xlabel('Iteration', fontsize=14)
    axs[1].set_ylabel('Plant Constraint', fontsize=14)
    axs[1].legend()
    plt.tight_layout()
    plt.savefig(os.path.join(output_dir, figname))
