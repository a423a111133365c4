from pantea_b200.descriptors.acsf.acsf import ACSF
from pantea_b200.descriptors.scaler import DescriptorScaler, ScalerParams

__all__ = ["ACSF", "DescriptorScaler", "ScalerParams"]
