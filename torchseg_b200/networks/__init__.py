from .bisenet import BiSeNet, SpatialPath, BiSeNetHead
from .fcn import FCN
from .pspnet import PSPNet, PyramidPooling
from .dfn import DFN, DFNHead
from .psanet import PSANet, PointwiseSpatialAttention
