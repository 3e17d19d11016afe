from .bisenet import BiSeNet, SpatialPath, BiSeNetHead
