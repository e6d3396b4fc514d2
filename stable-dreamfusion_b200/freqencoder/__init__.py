from .freq import *
