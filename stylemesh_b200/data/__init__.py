"""GPU-resident view store and scene readers: replaces the per-step DataLoader path of the
reference (data/abstract_dataset.py) for the texture-optimisation loop."""
from .view_store import RawView, ViewStore  # noqa: F401
