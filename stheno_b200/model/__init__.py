from .fdd import *
from .gp import *
from .measure import *
from .observations import *
