from .rlecore import RunLength            # noqa: F401
from .entropycoder import EntropyCoder    # noqa: F401
from .rlemain import RLEMain              # noqa: F401
