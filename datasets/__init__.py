"""Dataset registry.  The disk loaders of the reference (nsvf/nerf/colmap/ngp) are out of scope
(no dataset exists offline) and opt.py rejects those names; the synthetic Lego-shape set
drives the hot path and the teacher set (views of the reference's shipped trained Lego model) the PSNR protocol."""
from .synthetic import SyntheticLego
from .teacher import TeacherLego

dataset_dict = {'synthetic': SyntheticLego, 'teacher': TeacherLego}
