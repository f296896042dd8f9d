"""Import shim for the reference's ``from openbabel import openbabel, pybel``
(clusttraj/openbabel_compat.py:13): XYZ-only reader.  TEST INFRASTRUCTURE."""
from . import openbabel, pybel  # noqa: F401
