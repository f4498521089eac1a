"""pvcr_b200 — B200-native (sm_100a) captioning hot path behind the reference's module API.

Import as ``pvcr_b200`` (a shim at the repository root maps that name onto this directory, whose own
name is not a valid Python identifier).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
