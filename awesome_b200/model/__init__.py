from .convex_net import ConvexNet, ConvexNextNet, OutBlock, SkipBlock  # noqa: F401
