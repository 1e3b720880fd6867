"""CPU: the reference's CUDA extensions, compiled unmodified by oracle/Makefile into oracle/_ref/, load without a GPU
and export the reference's pybind surface (bind.cpp:31-36, tree_filter.cpp:7-13).  They are the GPU-side oracle of
tests/test_reference_ext_gpu.py; skipped where the reference checkout was not available at build time."""
import glob
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name,symbols', [
    ('pairwise_ext_ref', ['pairwise_nlog_forward', 'pairwise_nlog_backward']),
    ('tree_filter_cuda_ref', ['mst_forward', 'bfs_forward', 'refine_forward', 'refine_backward_feature',
                              'refine_backward_weight'])])
def test_reference_extension_loads(name, symbols):
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    hits = glob.glob(os.path.join(ROOT, 'oracle', '_ref', name + '*.so'))
    if not hits:
        pytest.skip(f'oracle/_ref/{name}*.so not built')
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for s in symbols:
        assert callable(getattr(mod, s))
