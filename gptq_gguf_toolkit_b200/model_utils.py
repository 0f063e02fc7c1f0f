"""Model plumbing -- mirror of the reference's quant/gptq/src/{model_utils,common_utils}.py (host glue)."""
from __future__ import annotations

import random
import re
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.conv import _ConvNd

LINEAR_LAYERS = (nn.Linear, _ConvNd)          # model_utils.py:12


def get_number_of_rows_and_cols(layer):      # model_utils.py:56-57
    return layer.weight.shape[0], int(np.prod(layer.weight.shape[1:]))


class ForwardInterrupt(Exception):           # model_utils.py:14-15
    pass


class InputCollector(nn.Module):             # model_utils.py:18-37
    def __init__(self, module: nn.Module, cpu_offload: bool = False):
        super().__init__()
        self.module = module
        self.cpu_offload = cpu_offload
        self.input_args = []
        self.input_kwargs = []

    def forward(self, *input_args, **input_kwargs):
        if self.cpu_offload:
            input_args = to(input_args, device="cpu")
            input_kwargs = to(input_kwargs, device="cpu")
        self.input_args.append(input_args)
        self.input_kwargs.append(input_kwargs)
        raise ForwardInterrupt


def select_layers(model: nn.Module, layer_prefix: Optional[str] = "", layer_regex: str = ".*",
                  layer_classes: Union[nn.Module, List[nn.Module]] = nn.Module) -> Dict[str, nn.Module]:
    """model_utils.py:39-53: modules of the given classes whose name matches the regex and starts with the prefix."""
    layers = {}
    for layer_name, layer in model.named_modules():
        if isinstance(layer, layer_classes) and re.search(layer_regex, layer_name) and layer_name.startswith(layer_prefix):
            layers[layer_name] = layer
    return layers


def to(data: Any, *args, **kwargs):          # common_utils.py
    if isinstance(data, torch.Tensor):
        return data.to(*args, **kwargs)
    if isinstance(data, (list, tuple)):
        return type(data)(to(v, *args, **kwargs) for v in data)
    if isinstance(data, dict):
        return {k: to(v, *args, **kwargs) for k, v in data.items()}
    return data


def maybe_first_element(x):
    if isinstance(x, (tuple, list)):
        x = x[0]
    return x


def fix_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
