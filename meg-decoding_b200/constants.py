"""Mirror of the reference's top-level ``constants.py`` (imported by its scripts as ``from constants import device``)."""
import torch

device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
bar_format = "{desc:<5.5}{percentage:3.0f}%|{bar:10}{r_bar}"
