"""``bfs(edge_index, max_adj_per_vertex)`` -- mmdet/ops/tree_filter/functions/bfs.py:9-16."""
from .. import tree_filter_cuda as _C


def bfs(edge_index, max_adj_per_vertex):
    """-> (sorted_index, sorted_parent, sorted_child); pure index work, no autograd node."""
    return _C.bfs_forward(edge_index, max_adj_per_vertex)
