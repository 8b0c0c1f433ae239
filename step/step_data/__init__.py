"""Sliding-window datasets over the pre-processed BasicTS pickles (history, long history and future windows)."""
from .device_loader import DeviceWindowLoader
from .forecasting_dataset import ForecastingDataset
from .pretraining_dataset import PretrainingDataset

__all__ = ["PretrainingDataset", "ForecastingDataset", "DeviceWindowLoader"]
