"""NumPy restatement of the analytic Benoit plant (reference problems/Benoit_Problem.py:14-44).
Black-box host functions, evaluated once per BO iteration -- fixtures, not part of the hot path."""
import random

import numpy as np


def Benoit_System_1(u, noise=0):
    f = u[0] ** 2 + u[1] ** 2 + u[0] * u[1]
    if noise:
        f += random.gauss(0., np.sqrt(noise))
    return f


def Benoit_System_2(u, noise=0):
    f = u[0] ** 2 + u[1] ** 2 + (1 - u[0] * u[1]) ** 2
    if noise:
        f += random.gauss(0., np.sqrt(noise))
    return f


def con1_system(u, noise=0):
    g1 = 1. - u[0] + u[1] ** 2 + 2. * u[1] - 2.
    if noise:
        g1 -= random.gauss(0., np.sqrt(noise))
    return -g1


def con1_system_tight(u, noise=0):
    g1 = 1. - u[0] + u[1] ** 2 + 2. * u[1]
    if noise:
        g1 -= random.gauss(0., np.sqrt(noise))
    return -g1
