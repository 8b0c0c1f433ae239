"""Sliding-window dataset over the pre-processed BasicTS pickle (history, long history and future windows)."""
from .device_loader import DeviceWindowLoader
from .forecasting_dataset import ForecastingDataset

__all__ = ["ForecastingDataset", "DeviceWindowLoader"]
