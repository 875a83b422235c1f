"""drmlt-mitsuba_b200: B200-native drop-in for the pssmlt / drmlt integrator hot path of
joeylitalien/drmlt-mitsuba.  CUDA kernels + C ABI live in csrc/, the host-side mirror of the
reference's plugin interface in integrator.py, procedural benchmark scenes in scenes.py."""
__version__ = "0.1.0"
