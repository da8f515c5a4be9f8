"""point_teacher_b200 -- B200-native (sm_100a) implementation of Point Teacher's phase-2 dynamic-MIL
pseudo-box refinement path behind the reference's mmdet / mmrotate interfaces.  See DESIGN.md."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.SO_PATH
