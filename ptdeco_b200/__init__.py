"""ptdeco_b200: B200-native falor/dwain decomposition hot path.

Mirrors the reference package layout (src/ptdeco/__init__.py:1-4 imports dwain and utils; falor is
imported explicitly as `ptdeco_b200.falor`, like `ptdeco.falor`)."""
from . import dwain, utils  # noqa: F401

__version__ = "0.1.0"
