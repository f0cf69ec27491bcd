"""Footstep, CoP-bound and reference-speed generators (host side)."""

from .footstep_generator import Contact, generate_footsteps
from .cop_generator import CoPGenerator, State

__all__ = ["Contact", "generate_footsteps", "CoPGenerator", "State"]
