from .corr import CorrBlock, bilinear_sampler, coords_grid, tfa_sampler, upflow8
from .extractor import BasicEncoder, SmallEncoder
from .update import (BasicMotionEncoder, BasicUpdateBlock, ConvGRU, FlowHead, SepConvGRU, SmallMotionEncoder,
                     SmallUpdateBlock)
