from .functions import bfs, mst, refine  # noqa: F401
from .modules import MinimumSpanningTree, TreeFilter2D  # noqa: F401
