from .convex_net import ConvexNet, ConvexNextNet, OutBlock, SkipBlock  # noqa: F401
from .path_connected_net import (MeanStd, MinMax, NoisyPathConnectedNet, NormNet, PathConnectedNet, PixelizeNet, RealNVP, get_norm,  # noqa: F401
                                 init_realnvp, real_nvp_path_connected_net, realnvp_masks)
from .multi_prior import (BatchSizeMultiPriorModule, MultipleObjectsAwarePathConnectedNet,  # noqa: F401
                          NumberBasedMultiPriorModule)
from .star_net import StarFitter, StarShapedNet  # noqa: F401
from .convex_diffeomorphism_net import ConvexDiffeomorphismNet, NormalizingFlow1D  # noqa: F401
