from .grid import *
