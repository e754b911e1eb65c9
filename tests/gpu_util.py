import pytest
import torch

requires_gpu = pytest.mark.gpu


def have_gpu():
    return torch.cuda.is_available()
