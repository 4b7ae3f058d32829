"""Locate the parent package whether ``models`` was imported as ``sbo_b200.models`` or, reference-style,
as a top-level ``models`` package with this directory's parent on ``sys.path``."""
import importlib
import os
import sys


def package():
    if "sbo_b200" in sys.modules:
        return sys.modules["sbo_b200"]
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("sbo_b200")
