"""Catalogued storage of the products either side of the hot path (Monte Carlo samples, numerical integrals,
serialized functions): the reference's ``file_management`` package (file_management/catalog_utils.py,
file_management/data_management.py) with the same file formats -- a YAML catalog keyed by the sorted parameter
dictionary and one .npz / .npy / .pkl file per entry -- so that files written by either side load on the other.
``jetmontecarlo_b200.install_as`` also registers this package as top-level ``file_management`` when no such package
is importable."""
from . import catalog_utils, data_management, load_data      # noqa: F401
