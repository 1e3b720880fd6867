from .bfs import bfs  # noqa: F401
from .mst import mst  # noqa: F401
from .refine import refine  # noqa: F401
