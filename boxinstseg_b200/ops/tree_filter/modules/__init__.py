from .tree_filter import MinimumSpanningTree, TreeFilter2D  # noqa: F401
