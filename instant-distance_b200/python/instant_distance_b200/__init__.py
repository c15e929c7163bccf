"""instant-distance-b200: Blackwell-native HNSW build-and-search behind djc/instant-distance's surface."""
from . import _abi  # noqa: F401
