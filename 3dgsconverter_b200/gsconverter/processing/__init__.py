from .data_processor import DataProcessor

__all__ = ["DataProcessor"]
