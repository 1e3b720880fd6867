"""``bfs(edge_index, max_adj_per_vertex)`` -- mmdet/ops/tree_filter/functions/bfs.py:9-16."""
from .. import tree_filter_cuda as _C


def bfs(edge_index, max_adj_per_vertex, root=0):
    """-> (sorted_index, sorted_parent, sorted_child); pure index work, no autograd node.
    ``root`` (extension; the reference always roots at vertex 0): the vertex at position 0."""
    return _C.bfs_forward(edge_index, max_adj_per_vertex, root)
