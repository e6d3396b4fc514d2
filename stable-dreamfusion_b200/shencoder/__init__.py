from .sphere_harmonics import *
