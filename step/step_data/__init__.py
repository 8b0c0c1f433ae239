from .forecasting_dataset import ForecastingDataset

__all__ = ["ForecastingDataset"]
