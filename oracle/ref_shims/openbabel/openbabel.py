"""Placeholder for the SWIG module; the hot path never touches it."""
