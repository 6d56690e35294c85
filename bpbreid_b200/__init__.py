"""bpbreid_b200 -- B200-native (sm_100a) implementation of the BPBReID data-parallel hot path behind the reference's
torchreid model-builder / loss / metrics API.  See DESIGN.md for the path, INTEGRATION.md for the drop-in stubs."""
from . import constants  # noqa: F401

__version__ = "0.1.0"
