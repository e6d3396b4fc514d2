"""Drop-in for the subset of `nvdiffrast` the reference's DMTet stage binds (`import nvdiffrast.torch as dr`, nerf/renderer.py:12,
309-312, 895-931): RasterizeCudaContext / RasterizeGLContext, rasterize, interpolate, antialias — on csrc/meshrast.cu."""
__version__ = "0.3.1+sdf_b200"
