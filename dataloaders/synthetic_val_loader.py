"""Validation-flavoured alias of the synthetic loader (the reference has separate train/val loader plugins)."""
from dataloaders.synthetic_loader import SyntheticLoader


def create_loader():
    return SyntheticLoader()
