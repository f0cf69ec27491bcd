"""Utilities (plotting lives in the drivers and is imported lazily)."""
