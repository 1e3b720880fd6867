"""CPU: the registry surface -- classes build from the reference's own config dicts."""
import torch


def test_build_from_reference_config_kwargs():
    from boxinstseg_b200.models import HEADS, LOSSES, build_head, build_loss
    assert HEADS is LOSSES                                    # mmdet/models/builder.py:9-15: one registry
    # configs/boxinst/boxinst_r50_fpn_1x_coco.py:54-71
    head = build_head(dict(type='CondInstMaskHead', in_channels=16, in_stride=8, out_stride=4, dynamic_convs=3,
                           dynamic_channels=8, disable_rel_coors=False, bbox_head_channels=256,
                           sizes_of_interest=[64, 128, 256, 512, 1024], max_proposals=-1, topk_per_img=64,
                           boxinst_enabled=True, bottom_pixels_removed=10, pairwise_size=3, pairwise_dilation=2,
                           pairwise_color_thresh=0.3, pairwise_warmup=10000))
    assert head.num_gen_params == 233 and head.param_conv.out_channels == 233
    assert set(head.state_dict()) >= {'sizes_of_interest', '_iter', 'param_conv.weight', 'param_conv.bias'}
    # configs/boxlevelset, configs/box2mask
    assert build_loss(dict(type='BoxProjectionLoss', loss_weight=3.0)).loss_weight == 3.0
    assert build_loss(dict(type='LevelsetLoss', loss_weight=1.0)).loss_weight == 1.0
    assert build_head(dict(type='BoxSOLOv2Head', num_classes=80, in_channels=256)).num_classes == 80


def test_empty_instances_follow_reference_branch():
    from boxinstseg_b200.models import build_head
    head = build_head(dict(type='CondInstMaskHead', in_channels=16, boxinst_enabled=True, max_proposals=-1, topk_per_img=64))
    x = torch.zeros(0, 1, 8, 8, requires_grad=True)
    out = head.loss(None, [], x, torch.zeros(0, dtype=torch.long), [], None, None)     # condinst_head.py:1306-1312
    assert float(out['loss_prj']) == 0.0 and float(out['loss_pairwise']) == 0.0
    assert float(head._iter) == 1.0


def test_tree_filter_graph_construction_matches_oracle():
    from boxinstseg_b200.ops.tree_filter import MinimumSpanningTree, TreeFilter2D
    from oracle import tree as ot
    fm = torch.randn(2, 3, 5, 7)
    mst = MinimumSpanningTree(TreeFilter2D.norm2_distance)
    assert (mst._build_matrix_index(fm)[0].numpy() == ot.grid_edges(5, 7)).all()
    assert torch.equal(mst._build_feature_weight(fm), ot.grid_edge_weights(fm))


def test_partial_heads_inherit_the_reference_methods():
    """with mmdet importable the hot-path heads must not shadow the reference's other methods
    (detectors/condinst.py:69,90 call training_sample / simple_test).  Simulated with a stand-in reference class."""
    from boxinstseg_b200.models import builder as b
    reg = b.Registry('sim') if not b.HAVE_MMDET else None
    if reg is None:
        return

    class FakeHead(torch.nn.Module):                   # plays mmdet's complete class, registered first
        def __init__(self, width=3):
            super().__init__()
            self.width = width

        def training_sample(self):
            return 'ref training_sample'

        def simple_test(self):
            return 'ref simple_test'

        def loss(self):
            return 'ref loss'
    reg.register_module(module=FakeHead)

    @b.register(reg, partial=True, name='FakeHead')
    class HotPath(torch.nn.Module):
        def loss(self):
            return 'b200 loss'
    head = reg.build(dict(type='FakeHead', width=7))
    assert head.width == 7                              # the reference constructor ran
    assert head.loss() == 'b200 loss'                   # hot path overridden
    assert head.training_sample() == 'ref training_sample' and head.simple_test() == 'ref simple_test'
    for m in ('forward', 'training_sample', 'loss', 'simple_test'):     # detectors/condinst.py:54-90
        assert hasattr(head, m)


def test_discobox_head_builds_correspondence_state_from_config():
    """loss_corr of configs/discobox/discobox_solov2_coco_r50_fpn_3x.py:65-93 -> solver + object bank (f4)."""
    from boxinstseg_b200.models import build_head
    bank = dict(img_norm_cfg=None, len_object_queues=100, fg_iou_thresh=0.7, bg_iou_thresh=0.7, ratio_range=[0.9, 1.2],
                appear_thresh=0.7, min_retrieval_objs=2, max_retrieval_objs=5, feat_height=7, feat_width=7, mask_height=28,
                mask_width=28, img_height=200, img_width=200, min_size=32, num_gpu_bank=20)
    head = build_head(dict(type='DiscoBoxSOLOv2Head', num_classes=80, in_channels=256,
                           loss_corr=dict(type='InfoNCE', loss_weight=1.0, corr_exp=1.0, corr_eps=0.05, gaussian_filter_size=3,
                                          low_score=0.3, corr_num_iter=10, corr_num_smooth_iter=1, save_corr_img=False,
                                          dist_kernel=9, obj_bank=bank)))
    assert head.corr is not None and head.corr.solver.num_iter == 10 and head.corr.object_queues.len_queue == 100
    assert head.corr.objbank_min_size == 32 and len(head.corr.object_queues.queues) == 80
    plain = build_head(dict(type='DiscoBoxSOLOv2Head', num_classes=80, in_channels=256))
    assert plain.corr is None
    import pytest
    with pytest.raises(RuntimeError):
        plain.corr_loss_levels([], [], [], [], [], None, None, None)
