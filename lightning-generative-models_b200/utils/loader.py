"""Config / model loader with the reference's convention (reference utils/loader.py:15-86):
`load_config(path)` reads the JSON and checks that img_channels / img_size agree between the "model" and "dataset"
sections; `load_model(cfg["model"])` resolves `models.generative.<family>.<name.lower()>.<name>` and instantiates
it with cfg["model"]["args"].  Only the diffusion family exists in this package (the hot path); asking
for another family fails the way the reference fails for an unknown model."""
import json
from importlib import import_module
from typing import Dict

GENERATIVE_MODELS = ["diffusion"]


def load_model(model_config: Dict, **extra):
    """`extra` = keyword-only B200 execution options (precision, device, ...) passed on to the class."""
    name = model_config["name"]
    errors = []
    for family in GENERATIVE_MODELS:
        try:
            module = import_module(f"models.generative.{family}.{name.lower()}")
            return getattr(module, name)(**model_config["args"], **extra)
        except ImportError as e:
            errors.append(str(e))
    raise ValueError(f"Failed to import {name}. Errors encountered: \n " + "\n".join(errors))


def load_config(config_path: str) -> Dict:
    try:
        with open(config_path, "r") as f:
            config = json.load(f)
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file not found at '{config_path}'.")
    except json.JSONDecodeError:
        raise ValueError(f"The file at '{config_path}' is not a valid JSON.")
    margs, dset = config.get("model", {}).get("args", {}), config.get("dataset", {})
    if margs.get("img_channels") != dset.get("img_channels"):
        raise ValueError("Mismatch in 'img_channels' between model and dataset configurations.")
    if margs.get("img_size") != dset.get("img_size"):
        raise ValueError("Mismatch in 'img_size' between model and dataset configurations.")
    return config
