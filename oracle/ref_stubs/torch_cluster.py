# functional stand-in for the un-vendored torch_cluster extension (see README.md).
from oracle.cluster_ops import fps, knn  # noqa: F401
