"""scflow_b200: B200-native (sm_100a) implementation of SCFlow's iterative pose-refinement hot path behind the
reference's registry / config API.  See DESIGN.md and INTEGRATION.md."""
from .registry import Registry, build_from_cfg, Config, ConfigDict
from .builder import (REFINERS, DECODERS, ENCODERS, HEAD, LOSSES, BACKBONES, build_refiner, build_decoder, build_encoder,
                      build_head, build_loss, build_backbone)
from .cnn import ConvModule, BaseModule
from .corr_lookup import CorrLookup, coords_grid
from .pose_head import MultiClassPoseHead, SingleClassPoseHead
from .decoder import SCFlowDecoder, RAFTDecoder, RAFTDecoderMask, CorrelationPyramid, MotionEncoder, ConvGRU, XHead
from .encoder import RAFTEncoder
from .refiner import SCFlowRefiner
from .loss import SequenceLoss, RAFTLoss, L1Loss, DisentanglePointMatchingLoss, filter_flow_by_mask, refiner_loss
from .pose import (get_pose_from_delta_pose, cal_3d_2d_corr, get_flow_from_delta_pose_and_points, unproject_dense,
                   get_flow_from_delta_pose_dense)
from . import ops
from ._lib import ScfError

__version__ = '0.1.0'

__all__ = ['Registry', 'build_from_cfg', 'Config', 'ConfigDict', 'REFINERS', 'DECODERS', 'ENCODERS', 'HEAD', 'LOSSES',
           'BACKBONES', 'build_refiner', 'build_decoder', 'build_encoder', 'build_head', 'build_loss', 'build_backbone',
           'ConvModule', 'BaseModule', 'CorrLookup', 'coords_grid', 'MultiClassPoseHead', 'SingleClassPoseHead',
           'SCFlowDecoder', 'RAFTDecoder', 'RAFTDecoderMask', 'CorrelationPyramid', 'MotionEncoder', 'ConvGRU', 'XHead', 'RAFTEncoder', 'SCFlowRefiner',
           'get_pose_from_delta_pose', 'cal_3d_2d_corr', 'get_flow_from_delta_pose_and_points', 'unproject_dense',
           'get_flow_from_delta_pose_dense', 'ops', 'ScfError', 'SequenceLoss', 'RAFTLoss', 'L1Loss',
           'DisentanglePointMatchingLoss', 'filter_flow_by_mask', 'refiner_loss']
