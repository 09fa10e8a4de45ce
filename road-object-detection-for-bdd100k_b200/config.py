"""Hot-path constants with the names and values of the reference's `config.py`
(config.py:15-16,19,30-38,74-80,87).  Kernels take every value as an explicit
parameter; these are only the defaults the drop-in functions read."""
from enum import Enum, unique

normal_anchor_range = [0.05, 0.7]      # config.py:15
special_anchor_range = [0.02, 0.03]    # config.py:16
img_size = (418, 418)                  # config.py:19

supported_backbone_name = ['vgg_16', 'mobilenet_v2']
extract_feat_name = {                  # config.py:25-31 (only the count per backbone matters here)
    'vgg_16': ['backbone/vgg_16/conv4/conv4_3', 'backbone/vgg_16/conv5/conv5_3',
               'backbone/vgg_16/block7/conv7', 'backbone/vgg_16/block8/conv3x3',
               'backbone/vgg_16/block9/conv3x3', 'backbone/vgg_16/block10/conv3x3'],
    'mobilenet_v2': ['layer_11', 'layer_15', 'layer_18', 'layer_20', 'layer_22', 'layer_24'],
}
feat_size_all_layers = {               # config.py:34-38, valid for 418x418 input
    'mobilenet_v2': {'layer_1': (53, 53), 'layer_2': (27, 27), 'layer_3': (14, 14),
                     'layer_4': (7, 7), 'layer_5': (4, 4), 'layer_6': (2, 2)},
    'vgg_16': {'layer_1': (52, 52), 'layer_2': (26, 26), 'layer_3': (13, 13),
               'layer_4': (7, 7), 'layer_5': (4, 4), 'layer_6': (2, 2)},
}
# stride 8/16/32/64/128/256 with SAME padding => 512x512 input
feat_size_512 = {'layer_1': (64, 64), 'layer_2': (32, 32), 'layer_3': (16, 16),
                 'layer_4': (8, 8), 'layer_5': (4, 4), 'layer_6': (2, 2)}


@unique
class train_range(Enum):               # config.py:52-55 (selects what the net factory returns)
    REFINE = 0
    ALL = 1


@unique
class refine_method(Enum):             # config.py:73-77
    NEAREST_NEIGHBOR = 0
    JACCARD_BIGGER = 1
    JACCARD_TOPK = 2


refine_pos_jac_val_all_layers = [0.2, 0.3, 0.4, 0.4, 0.3, 0.3]   # config.py:79
det_pos_jac_val_all_layers = [0.5, 0.6, 0.7, 0.7, 0.6, 0.6]      # config.py:80
total_obj_n = 11                                                  # config.py:87 (incl. background)
