from pantea_b200.datasets.dataset import Dataset
from pantea_b200.datasets.runner import RunnerDataSource

__all__ = ["Dataset", "RunnerDataSource"]
