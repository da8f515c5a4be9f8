"""point_teacher_b200 -- B200-native (sm_100a) implementation of Point Teacher's phase-2 dynamic-MIL
pseudo-box refinement path behind the reference's mmdet / mmrotate interfaces.  See DESIGN.md."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.SO_PATH


def install():
    """Register every B200 class into the reference's registries under the reference's own type names
    (``SingleRoIExtractor``, ``RotatedSingleRoIExtractor``, ``TopkAssigner``, ``FUSETopkAssigner``, ``MaxIoUAssigner``,
    ``BboxOverlaps2D``, ``BboxDistanceMetric``, ``TS_P2BFCOSHead``, ``TS_P2RBRotatedFCOSHead``) with ``force=True``.
    Call it after ``import mmdet.models`` (and ``mmrotate.models``) and before the detector is built; see
    INTEGRATION.md.  Returns {reference head type name: class now registered}."""
    from . import assigners, losses, mil_head, registry, roi_extractors  # noqa: F401  (import = register)
    return registry.install_reference_heads()
