"""arah_release_b200 — B200-native (sm_100a) hot path of taconite/arah-release: articulated-SDF ray tracing,
Broyden root finding, SDF/colour MLP shading and volume compositing as hand-written CUDA behind a C ABI."""
from ._lib import ArahError  # noqa: F401

__all__ = ['ArahError']
