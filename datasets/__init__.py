"""Dataset registry.  The disk loaders of the reference (nsvf/nerf/colmap/ngp) are out of scope
(no dataset exists offline); the synthetic Lego-shape set drives the hot path."""
from .synthetic import SyntheticLego

dataset_dict = {'synthetic': SyntheticLego}
